// Shared by sd_kv.cu (stand-alone kv_append launches) and sd_verify.cu (the fused verify + append
// launch): parameters, helpers and the body of the column kernel.
#pragma once
#include "sd_common.cuh"

namespace sd {

struct KvParams {
    const sd_kv_layer* layers;  // device table, or nullptr -> `single` (passed by value)
    sd_kv_layer single;
    int64_t base_len;           // uniform length added to seq_lens[b] (functional single-layer form)
    int n_layers, B, H, R;
    int64_t row_bytes, capacity;
    int64_t arena_stride_b, arena_stride_h;
    int64_t new_stride_b, new_stride_h, new_stride_r;
    const int32_t* seq_lens;
    int32_t* seq_lens_out;
    const int32_t* n_keep;
    int keep_bias;
    const uint8_t* mask;
    int zero_rejected;
    int early_rows;  // column kernel: fetch the offered rows before griddepcontrol.wait
    int32_t* overflow;
    int32_t* ticket;
    const sd_exchange* xchg;  // nullable: confirm + gather the verify records of this step (first CTA, once its own rows are copied)
    int32_t* xchg_out;
    const int32_t* xchg_packed;  // nullable: this launch also PUSHES the local records (first CTA, at its start)
    int32_t* seq_ready;          // nullable: per-sequence flags set by the verify kernel's finalizers; the column
                                 // kernel then waits per sequence instead of for the whole previous grid
    unsigned total_ctas;
    int fused;  // 1 = running inside verify_append_kernel (grid-level wait already done)
};

SD_DEVINL int ld_flag(const int32_t* p) {
    int v;
    asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

SD_DEVINL int keep_count(const KvParams& p, int b) {
    int n = p.R;
    if (p.n_keep) {
        n = __ldcg(p.n_keep + b) + p.keep_bias;  // L2: may have been written by a kernel that is still running
        n = n < 0 ? 0 : (n > p.R ? p.R : n);
    }
    return n;
}

// position of the j-th set bit of mask[b, 0..R) (reference.py:149-157), or -1
SD_DEVINL int masked_source_row(const uint8_t* mask, int b, int R, int j) {
    int seen = 0;
    for (int k = 0; k < R; ++k) {
        if (mask[(int64_t)b * R + k]) {
            if (seen == j) return k;
            ++seen;
        }
    }
    return -1;
}

// last CTA to retire publishes the new lengths (so seq_lens_out may alias seq_lens)
SD_DEVINL void retire_and_commit(const KvParams& p) {
    __shared__ int s_last;
    // The ticket orders every CTA's READS of seq_lens / n_keep (complete once their values have been
    // consumed, i.e. before this barrier) ahead of the length commit; the row stores themselves need
    // no fence -- stream order publishes them to the next kernel.
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned prev = atomicAdd(reinterpret_cast<unsigned*>(p.ticket), 1u);
        s_last = prev == p.total_ctas - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    if (p.seq_ready) {
        // This launch did not wait for the previous grid as a whole (per-sequence flags instead): do it now, so
        // that "this kernel has completed" implies "the verify kernel has completed" for the next launch, and
        // re-arm the flags (every CTA has consumed its own by the time the last one retires).
        pdl_wait();
        for (int b = threadIdx.x; b < p.B; b += blockDim.x) p.seq_ready[b] = 0;
    }
    if (p.seq_lens_out) {
        for (int b = threadIdx.x; b < p.B; b += blockDim.x) {
            int len = (p.seq_lens ? p.seq_lens[b] : 0) + (int)p.base_len;
            int n = keep_count(p, b);
            if (p.mask) {  // cannot keep more rows than the mask offers
                int bits = 0;
                for (int k = 0; k < p.R; ++k) bits += p.mask[(int64_t)b * p.R + k] != 0;
                n = n < bits ? n : bits;
            }
            if ((int64_t)len + n > p.capacity) n = (int)(p.capacity - len > 0 ? p.capacity - len : 0);
            p.seq_lens_out[b] = len + n;
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) *p.ticket = 0;
}

// ---- variant 3 "column": one thread owns one 16-byte column of a (layer, b, h) for ALL R offered
// rows of K and V: up to 2*RB independent 128-bit loads in flight per thread, issued before the first
// store, and R-times fewer threads than variant 1 (one wave instead of four).
template <int RB>
SD_DEVINL void kv_col_body(const KvParams& p, const int cta_x, const int layer) {
    const int vpr = (int)(p.row_bytes >> 4);
    const int64_t slots = (int64_t)p.B * p.H * vpr;
    const int64_t slot = (int64_t)cta_x * blockDim.x + threadIdx.x;
    const bool active = slot < slots;
    int v = 0, h = 0, b = 0;
    sd_kv_layer L{nullptr, nullptr, nullptr, nullptr};
    const char* sk = nullptr;
    const char* sv = nullptr;
    uint4 kx[RB], vx[RB];
    // early_rows: the offered rows (and the layer table) were produced at least two kernels ago -- in
    // the verify step they come from the target forward, and the kernel just before this one is the
    // fused verify, which only writes accept_len -- so the first RB rows are fetched BEFORE
    // griddepcontrol.wait, while the previous kernel drains.  Only seq_lens / n_keep wait.
    if (p.early_rows && active) {
        v = (int)(slot % vpr);
        const int64_t t = slot / vpr;
        h = (int)(t % p.H);
        b = (int)(t / p.H);
        L = p.layers ? p.layers[layer] : p.single;
        sk = reinterpret_cast<const char*>(L.new_k) + (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h + ((int64_t)v << 4);
        sv = reinterpret_cast<const char*>(L.new_v) + (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h + ((int64_t)v << 4);
#pragma unroll
        for (int u = 0; u < RB; ++u) {
            if (u < p.R) {
                kx[u] = ldg_stream_v4(sk + (int64_t)u * p.new_stride_r);
                vx[u] = ldg_stream_v4(sv + (int64_t)u * p.new_stride_r);
            }
        }
    }
    if (!p.seq_ready && !p.fused) pdl_wait();  // (fused launch: the kernel waited at its top)
    if (p.xchg && p.xchg_packed && cta_x == 0 && layer == 0) {
        if (p.seq_ready) {  // the packed records of ALL sequences must be final before they are pushed
            for (int bb = threadIdx.x; bb < p.B; bb += blockDim.x)
                while (ld_flag(p.seq_ready + bb) == 0) __nanosleep(100);
            __threadfence();
            __syncthreads();
        }
        exchange_push_cta(p.xchg, p.xchg_packed, p.B);
    }
    if (active) {
        if (!p.early_rows) {
            v = (int)(slot % vpr);
            const int64_t t = slot / vpr;
            h = (int)(t % p.H);
            b = (int)(t / p.H);
            L = p.layers ? p.layers[layer] : p.single;
            sk = reinterpret_cast<const char*>(L.new_k) + (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h + ((int64_t)v << 4);
            sv = reinterpret_cast<const char*>(L.new_v) + (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h + ((int64_t)v << 4);
        }
        if (p.seq_ready) {
            // overlap with the verify kernel: sequence b's accept_len is final once its finalizer raised the flag
#ifdef SD_FLAG_ACQUIRE  // next-round A/B (off): relaxed polling + ONE acquire load (LDG.STRONG.GPU + CCTL.IVALL) instead of
                        // a volatile poll + __threadfence(), which adds MEMBAR.SC.GPU + ERRBAR per thread (5 % of the
                        // fused step's stall samples under ncu)
            for (;;) {
                int f;
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.seq_ready + b) : "memory");
                if (f != 0) break;
                __nanosleep(100);
            }
            {
                int f;
                asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(f) : "l"(p.seq_ready + b) : "memory");
            }
#else
            while (ld_flag(p.seq_ready + b) == 0) __nanosleep(100);
            __threadfence();
#endif
        }
        const int len = (p.seq_lens ? p.seq_lens[b] : 0) + (int)p.base_len;
        int n = keep_count(p, b);
        int room = (int)(p.capacity - len);
        room = room < 0 ? 0 : (room > p.R ? p.R : room);
        if (n > room) {
            if (p.overflow) *p.overflow = 1;
            n = room;
        }
        const int nstore = p.zero_rejected ? room : n;
        const int64_t doff = (int64_t)b * p.arena_stride_b + (int64_t)h * p.arena_stride_h + (int64_t)len * p.row_bytes + ((int64_t)v << 4);
        char* dk = reinterpret_cast<char*>(L.arena_k) + doff;
        char* dv = reinterpret_cast<char*>(L.arena_v) + doff;
        for (int j0 = 0; j0 < nstore; j0 += RB) {
            if (!(p.early_rows && j0 == 0)) {
#pragma unroll
                for (int u = 0; u < RB; ++u) {
                    const int j = j0 + u;
                    if (j < n) {
                        kx[u] = ldg_stream_v4(sk + (int64_t)j * p.new_stride_r);
                        vx[u] = ldg_stream_v4(sv + (int64_t)j * p.new_stride_r);
                    }
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int j = j0 + u;
                if (j < nstore) {
                    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
                    *reinterpret_cast<uint4*>(dk + (int64_t)j * p.row_bytes) = j < n ? kx[u] : z;
                    *reinterpret_cast<uint4*>(dv + (int64_t)j * p.row_bytes) = j < n ? vx[u] : z;
                }
            }
        }
    }
    // Fused result exchange (multi-GPU): CTA (0,0) -- which pushed this rank's records at its start when
    // exchange_packed is set -- gathers every rank's records of the epoch once its own rows are copied:
    // the NVLink round trips overlap the rest of the append instead of trailing it.
    if (p.xchg && cta_x == 0 && layer == 0) {
#ifndef SD_REPRO_EPOCH_RACE  // (tools/repro_epoch_race.py builds a copy without this wait to show the stall)
        if (p.seq_ready && !p.xchg_packed) {
            // The records are pushed by the verify finalizers, which read the epoch from the exchange header and
            // may still be running (this CTA only waited for ITS sequences).  The gather ends by advancing the
            // epoch, so every local finalizer must have read it first -- it raises its flag after the push.
            // lag = 0 implies this (the gather waits for the rank's own records); lag = 1 does not: without
            // this wait a late finalizer tags its record with the NEXT epoch and the step after hangs.
            for (int bb = threadIdx.x; bb < p.B; bb += blockDim.x)
                while (ld_flag(p.seq_ready + bb) == 0) __nanosleep(100);
            __syncthreads();
        }
#endif
        exchange_gather_cta(p.xchg, p.xchg_out);
    }
    retire_and_commit(p);
}


// ---- "copy-all" append for the fused step (mode 2 of sd_verify_append) -------------------------------
// The append CTAs copy ALL R offered rows of their sequences and therefore depend on nothing the verify
// part produces: they never wait, so they can sit in front of their group's verify CTAs in block order
// and stream concurrently.  What does depend on accept_len -- zeroing the rejected slots and advancing
// the length -- is done by the verify finalizer of the sequence, after the copy CTAs of its group have
// signalled completion (group counter).
struct FusedAppend {
    int enabled;
    const sd_kv_layer* layers;
    int n_layers, H, R, spc;  // spc = sequences per group (per append x-slot)
    int64_t row_bytes, capacity, arena_stride_b, arena_stride_h;
    int32_t* seq_lens;
    int32_t* overflow;
    unsigned* group_cnt;      // [2 * groups]: {copy CTAs done, finalizers done}; all zero between launches
    int zero_rejected;
};

template <int RB>
SD_DEVINL void kv_copy_all_body(const KvParams& p, const int cta_x, const int layer, unsigned* group_cnt) {
    const int vpr = (int)(p.row_bytes >> 4);
    const int64_t slots = (int64_t)p.B * p.H * vpr;
    const int64_t slot = (int64_t)cta_x * blockDim.x + threadIdx.x;
    if (slot < slots) {
        const int v = (int)(slot % vpr);
        const int64_t t = slot / vpr;
        const int h = (int)(t % p.H);
        const int b = (int)(t / p.H);
        const sd_kv_layer L = p.layers[layer];
        const int len = p.seq_lens ? p.seq_lens[b] : 0;  // the finalizer advances it only after this CTA has signalled
        int room = (int)(p.capacity - len);
        room = room < 0 ? 0 : (room > p.R ? p.R : room);
        const char* sk = reinterpret_cast<const char*>(L.new_k) + (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h + ((int64_t)v << 4);
        const char* sv = reinterpret_cast<const char*>(L.new_v) + (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h + ((int64_t)v << 4);
        const int64_t doff = (int64_t)b * p.arena_stride_b + (int64_t)h * p.arena_stride_h + (int64_t)len * p.row_bytes + ((int64_t)v << 4);
        char* dk = reinterpret_cast<char*>(L.arena_k) + doff;
        char* dv = reinterpret_cast<char*>(L.arena_v) + doff;
        for (int j0 = 0; j0 < room; j0 += RB) {
            uint4 kx[RB], vx[RB];
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int j = j0 + u;
                if (j < room) {
                    kx[u] = ldg_stream_v4(sk + (int64_t)j * p.new_stride_r);
                    vx[u] = ldg_stream_v4(sv + (int64_t)j * p.new_stride_r);
                }
            }
#pragma unroll
            for (int u = 0; u < RB; ++u) {
                const int j = j0 + u;
                if (j < room) {
                    *reinterpret_cast<uint4*>(dk + (int64_t)j * p.row_bytes) = kx[u];
                    *reinterpret_cast<uint4*>(dv + (int64_t)j * p.row_bytes) = vx[u];
                }
            }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();  // the rows are visible before the finalizer (which may zero some of them) sees the count
        atomicAdd(group_cnt + 2 * cta_x, 1u);
    }
}

// The accept_len-dependent part of the append for sequence b, run by the verify finalizer CTA (all `nthreads`
// threads): wait for the copy CTAs of b's group, zero the rejected slots, advance the length.
SD_DEVINL void fused_append_finalize(const FusedAppend& fa, const int b, const int n_keep, const int tid, const int nthreads,
                                     const bool named_barrier) {
    const int g = b / fa.spc;
    if (tid == 0) {
        unsigned v;
        for (;;) {
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(fa.group_cnt + 2 * g) : "memory");
            if (v >= (unsigned)fa.n_layers) break;
            __nanosleep(64);
        }
        __threadfence();
    }
    if (named_barrier) asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    else __syncthreads();
    const int len = fa.seq_lens[b];
    int room = (int)(fa.capacity - len);
    room = room < 0 ? 0 : (room > fa.R ? fa.R : room);
    int n = n_keep < 0 ? 0 : (n_keep > fa.R ? fa.R : n_keep);
    if (n > room) {
        if (fa.overflow && tid == 0) *fa.overflow = 1;
        n = room;
    }
    if (fa.zero_rejected && room > n) {
        const int vpr = (int)(fa.row_bytes >> 4);
        const int rej = room - n;
        const int64_t items = (int64_t)fa.n_layers * 2 * fa.H * rej * vpr;
        const uint4 z = make_uint4(0u, 0u, 0u, 0u);
        for (int64_t i = tid; i < items; i += nthreads) {
            const int v = (int)(i % vpr);
            int64_t t = i / vpr;
            const int j = (int)(t % rej);
            t /= rej;
            const int h = (int)(t % fa.H);
            t /= fa.H;
            const int kv = (int)(t & 1);
            const int layer = (int)(t >> 1);
            const sd_kv_layer L = fa.layers[layer];
            char* dst = reinterpret_cast<char*>(kv ? L.arena_v : L.arena_k) + (int64_t)b * fa.arena_stride_b +
                        (int64_t)h * fa.arena_stride_h + (int64_t)(len + n + j) * fa.row_bytes + ((int64_t)v << 4);
            *reinterpret_cast<uint4*>(dst) = z;
        }
    }
    if (named_barrier) asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    else __syncthreads();
    if (tid == 0) {
        fa.seq_lens[b] = len + n;
        // the last finalizer of the group re-arms its counters for the next launch
        const unsigned done = atomicAdd(fa.group_cnt + 2 * g + 1, 1u);
        if (done + 1u == (unsigned)fa.spc) {
            fa.group_cnt[2 * g] = 0u;
            fa.group_cnt[2 * g + 1] = 0u;
        }
    }
}

// Validates the arguments of sd_kv_append for the column kernel and fills `out`, the (x, layers) grid and the
// load batch RB.  Returns SD_OK, an sd_status, or SD_ERR_UNSUPPORTED when the column kernel cannot serve the call.
int kv_prepare_col(const sd_kv_layer* layers, int n_layers, int B, int H, int R, int64_t row_bytes, int64_t capacity,
                   int64_t arena_stride_b, int64_t arena_stride_h, int64_t new_stride_b, int64_t new_stride_h,
                   int64_t new_stride_r, const int32_t* seq_lens, int32_t* seq_lens_out, const int32_t* n_keep,
                   int keep_bias, int zero_rejected, int64_t max_seq_len_host, int32_t* overflow_flag, int32_t* ticket,
                   const sd_exchange* exchange, int32_t* exchange_out, const int32_t* exchange_packed, int32_t* seq_ready,
                   KvParams* out, unsigned* grid_x, int* rb);

}  // namespace sd
