// Shared by sd_kv.cu (stand-alone kv_append launches) and sd_verify.cu (the fused verify + append
// launch): parameters, helpers and the body of the column kernel.
#pragma once
#include "sd_common.cuh"

namespace sd {

// n / d for n < 2^31 as one 64-bit multiply + shift (m = ceil(2^(31 + s) / d), s = ceil(log2 d); exact for 31-bit n)
struct FastDiv {
    uint32_t m, sh, d, pad;
};
inline FastDiv make_fastdiv(uint32_t d) {
    FastDiv f;
    uint32_t s = 0;
    while ((1ull << s) < d) ++s;
    f.sh = 31 + s;
    f.m = (uint32_t)(((1ull << f.sh) + d - 1) / d);
    f.d = d;
    f.pad = 0;
    return f;
}
SD_DEVINL int fd_div(const FastDiv& f, int n) { return (int)(((unsigned long long)(unsigned)n * f.m) >> f.sh); }

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
    // ---- lagged append (the default fused step, verify_step_kernel in sd_verify.cu) ----
    unsigned long long* flags;  // [B] sequence words {tag:24 | n_keep:8 | old length:32}, written by the verify finalizers
    unsigned* epoch_word;       // launch counter; tag = (epoch + 1) & 0xffffff
    int lag;                    // the rows of sequence b are moved by the verify CTAs of sequence b + lag (clamped to B)
    int pieces;                 // verify CTAs per sequence = pieces the append of one sequence is cut into
    int piece_items, seq_items; // items per piece / per sequence; item = the K and the V vector of one (layer, h, row, 16-byte column)
    int tail_pieces;            // tail CTAs per sequence
    // uniform layers (optional): layer l's four pointers are layer0's + l * {ls_arena, ls_arena, ls_new, ls_new}; the
    // device table is then never read -- an item's row request depends on nothing but the thread's index
    sd_kv_layer layer0;
    int64_t ls_arena, ls_new;
    int uniform;
    int carry_ni;               // items a carrying verify CTA parks per thread (1..kCarryNI)
    // verify CTAs of the first `lag` sequences carry nothing: their chunk fills stage + park (upc_big units, S_big per row)
    int n_big, S_big, upc_big;
    int fetch;                  // A/B: 0 = chunk by one bulk copy, 1 = by cp.async, 2 = half / half
    int bulk_split;             // A/B: bytes per bulk copy of a logits chunk (0 = the whole chunk in one copy)
    int one_warp_wait;          // A/B: one warp polls the chunk's mbarrier, the rest wait in bar.sync
    int stagger_ns;             // A/B: first-wave CTAs delay their bulk copy by slot * stagger_ns
    int l2_prefetch;            // verify CTAs of the last `lag` sequences pull their sequence's rows into L2 (default on)
    int eager_rows;             // A/B: carrying CTAs fetch their rows without looking at the sequence word first
    int tail_order;             // tail CTAs: 0 = the LAST sequence's pieces are dispatched first, 1 = the earliest tail sequence's
    int pf_ahead;               // A/B: verify CTA u prefetches the logits chunk of verify CTA u + pf_ahead into L2 (0 = off)
    int n_verify;               // verify CTAs of the fused step launch
    int tail_peek;              // tail CTAs look at the sequence word while row 0 is being requested; rows 1.. only if kept
    FastDiv fd_vpr, fd_R, fd_H; // item -> (column, row, head, layer) without integer division
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

// whole CTA: wait until the verify finalizers have raised every sequence's flag (bounded)
SD_DEVINL void wait_all_seq_ready(const KvParams& p) {
    for (int bb = threadIdx.x; bb < p.B; bb += blockDim.x) {
        SpinGuard guard;
        while (ld_flag(p.seq_ready + bb) == 0) {
            if (!guard.keep_waiting(kSpinAllFlags)) break;
            __nanosleep(100);
        }
    }
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
            wait_all_seq_ready(p);
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
            SpinGuard guard;
            while (ld_flag(p.seq_ready + b) == 0) {
                if (!guard.keep_waiting(kSpinSeqFlag)) break;
                __nanosleep(100);
            }
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
            wait_all_seq_ready(p);
            __syncthreads();
        }
#endif
        exchange_gather_cta(p.xchg, p.xchg_out);
    }
    retire_and_commit(p);
}


// ---- lagged append: the accept_len-dependent row moves, spread UNDER the verify phase ----------------
// (fused step, verify_step_kernel in sd_verify.cu).  The append of one sequence -- n_layers * H * R rows of K and
// of V -- is cut into `pieces` equal runs of items (item = the K and the V vector of one 16-byte column of one
// row: one index decode, by multiply-shift, serves both), one run per verify CTA of a sequence, and
// the pieces of sequence b are carried by the verify CTAs of sequence b + lag: by the time those run, b's
// finalizer has (almost always) published its word, so they fetch exactly the kept rows while their own logits
// chunk is in flight and store them once their partial is published.  Every arena byte is written exactly
// once (kept row or zero), by a CTA that knows accept_len: no copy-then-zero pass, no fence on the data path.
// The pieces of the last `lag` sequences are carried by tail CTAs behind the verify CTAs; their rows are pulled
// into L2 early by those sequences' own verify CTAs (kv_piece_prefetch).
//
// Sequence word (one 8-byte volatile store by the finalizer, polled by the consumers):
//   bits 63..40  tag = (launch epoch + 1) & 0xffffff   -- words of earlier launches never match: no re-arming
//   bits 39..32  n_keep = accept_len + 1 (1..65)
//   bits 31..0   the sequence's length BEFORE this step
// The finalizer is the only reader and the only writer of seq_lens[b] in the launch (it stores the new length
// right away), so lengths need no ticket, and a consumer gets everything it needs from ONE load.
SD_DEVINL unsigned long long ld_word64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SD_DEVINL bool word_valid(unsigned long long w, unsigned tag) {
    return (unsigned)(w >> 40) == tag && ((unsigned)(w >> 32) & 0xffu) != 0u;
}
SD_DEVINL unsigned long long make_word(unsigned tag, int n_keep, int len) {
    return ((unsigned long long)tag << 40) | ((unsigned long long)(unsigned)(n_keep & 0xff) << 32) | (unsigned)len;
}

// item of a sequence's append -> offsets.  ONE decode serves two vectors (K and V share everything but the base).
struct PieceItem {
    int64_t src_off, dst_off;  // dst_off excludes the length term (+ len * row_bytes)
    int j, layer;
};
SD_DEVINL PieceItem piece_item(const KvParams& p, const int seq, const int it) {
    int t = fd_div(p.fd_vpr, it);
    const int v = it - t * (int)p.fd_vpr.d;
    int t2 = fd_div(p.fd_R, t);
    PieceItem r;
    r.j = t - t2 * (int)p.fd_R.d;
    r.layer = fd_div(p.fd_H, t2);
    const int h = t2 - r.layer * (int)p.fd_H.d;
    r.src_off = (int64_t)seq * p.new_stride_b + (int64_t)h * p.new_stride_h + (int64_t)r.j * p.new_stride_r + ((int64_t)v << 4);
    r.dst_off = (int64_t)seq * p.arena_stride_b + (int64_t)h * p.arena_stride_h + (int64_t)r.j * p.row_bytes + ((int64_t)v << 4);
#ifdef SD_DIAG_ALIAS_DST  // diagnostic build only (tools/diag_tlb.sh): every destination folded into 1 MB -- WRONG results
    r.dst_off &= 0xFFFF0;
    r.layer = 0;
#endif
    return r;
}
struct PtrPair {
    const char* k;
    const char* v;
};
SD_DEVINL PtrPair unpack_pair(const uint4& q) {
    PtrPair r;
    r.k = reinterpret_cast<const char*>(((unsigned long long)q.y << 32) | q.x);
    r.v = reinterpret_cast<const char*>(((unsigned long long)q.w << 32) | q.z);
    return r;
}
SD_DEVINL PtrPair layer_src(const KvParams& p, const int layer) {  // {new_k, new_v} of a layer
    if (p.uniform) {
        const int64_t o = (int64_t)layer * p.ls_new;
        return PtrPair{reinterpret_cast<const char*>(p.layer0.new_k) + o, reinterpret_cast<const char*>(p.layer0.new_v) + o};
    }
    return unpack_pair(__ldg(reinterpret_cast<const uint4*>(&p.layers[layer].new_k)));
}
SD_DEVINL PtrPair layer_dst(const KvParams& p, const int layer) {  // {arena_k, arena_v}
    if (p.uniform) {
        const int64_t o = (int64_t)layer * p.ls_arena;
        return PtrPair{reinterpret_cast<const char*>(p.layer0.arena_k) + o, reinterpret_cast<const char*>(p.layer0.arena_v) + o};
    }
    return unpack_pair(__ldg(reinterpret_cast<const uint4*>(&p.layers[layer].arena_k)));
}

// The fetched rows are PARKED in shared memory (cp.async, 16 bytes per request, L2-only): no data registers are
// held across the CTA's own reduction, and every request is in flight from the first microsecond of the CTA.
// park layout: uint4 park[2 * NI][blockDim.x] -- slot (2u + kv) of thread t at park[(2u + kv) * blockDim.x + t].
constexpr int kCarryNI = 2;  // items a carrying verify CTA parks per thread (16 KB of shared memory at 256 threads)

SD_DEVINL void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc)
                 : "memory");
}
SD_DEVINL void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

struct PieceRegs {
    unsigned long long word;       // last observed sequence word
    int64_t dst_off[kCarryNI];
    int jl[kCarryNI];              // row | layer << 8 of parked item u (-1: none)
    int seq, it, it_end;           // this thread owns items it, it + blockDim.x, ... < it_end of the sequence's append
};

// Phase 1 (before the CTA's own work): look at the sequence word; if it is already published fetch only kept
// rows, else fetch regardless (the rows are inputs of the launch, so this never waits).
SD_DEVINL void kv_piece_begin(const KvParams& p, const int seq, const int piece, const unsigned tag, PieceRegs& r, uint4* park) {
    r.seq = seq;
    const int base = piece * p.piece_items;
    r.it_end = base + p.piece_items < p.seq_items ? base + p.piece_items : p.seq_items;
    r.it = base + (int)threadIdx.x;
    r.word = ld_word64(p.flags + seq);
    const int n = (!p.eager_rows && word_valid(r.word, tag)) ? (int)((unsigned)(r.word >> 32) & 0xffu) : p.R;
#pragma unroll
    for (int u = 0; u < kCarryNI; ++u) {
        const int it = r.it + u * (int)blockDim.x;
        r.jl[u] = -1;
        if (u < p.carry_ni && it < r.it_end) {
            const PieceItem pi = piece_item(p, seq, it);
            const PtrPair src = layer_src(p, pi.layer);
            r.dst_off[u] = pi.dst_off;
            r.jl[u] = pi.j | (pi.layer << 8);
            if (pi.j < n) {
                cp_async16(park + (2 * u) * blockDim.x + threadIdx.x, src.k + pi.src_off);
                cp_async16(park + (2 * u + 1) * blockDim.x + threadIdx.x, src.v + pi.src_off);
            }
        }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");  // (its own group: the chunk's cp.async part can be waited for alone)
}

struct PieceKeep {
    int len, n, nstore;
};
// wait (bounded) for the sequence word; false = it never came (no stores)
SD_DEVINL bool piece_wait(const KvParams& p, const unsigned tag, const int seq, unsigned long long& word, PieceKeep& k) {
    SpinGuard guard;
    while (!word_valid(word, tag)) {
        if (!guard.keep_waiting(kSpinSeqFlag)) return false;
        __nanosleep(64);
        word = ld_word64(p.flags + seq);
    }
    k.len = (int)(unsigned)word;
    int n = (int)((unsigned)(word >> 32) & 0xffu);
    const int64_t room64 = p.capacity - k.len;
    const int room = room64 < 0 ? 0 : (room64 > p.R ? p.R : (int)room64);
    k.n = n > room ? room : n;
    k.nstore = p.zero_rejected ? room : k.n;
    return true;
}

// items it0, it0 + blockDim.x, ... < it_end, `ni` per thread and pass parked in shared memory: all requests of a
// pass first (every offered row while the word is unknown), then the stores.  Used by the tail CTAs (ni = what the
// CTA's shared memory holds) and for pieces longer than kCarryNI items per thread.
// `peek` (tail CTAs): the word is loaded FIRST and consumed only after row 0 of every item has been requested (row 0
// is the one row every sequence keeps), so a sequence whose word is already published costs the reads of its kept
// rows only (R = 5, 3.2 kept on average: 36 % fewer row reads) at the price of one L2 latency that overlaps row 0.
SD_DEVINL void piece_items_run(const KvParams& p, const unsigned tag, const int seq, int it0, const int it_end,
                               unsigned long long& word, uint4* park, const int ni, const bool peek = false) {
    PieceKeep keep{0, 0, 0};
    unsigned long long peeked = 0ull;
    if (peek) peeked = ld_word64(p.flags + seq);
    bool known = word_valid(word, tag);
    if (known && !piece_wait(p, tag, seq, word, keep)) return;
    bool peek_now = peek && !known;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    for (; it0 < it_end; it0 += ni * (int)blockDim.x) {
        if (peek_now) {
            for (int u = 0; u < ni; ++u) {  // row 0: no dependence on the word
                const int it = it0 + u * (int)blockDim.x;
                if (it >= it_end) break;
                const PieceItem pi = piece_item(p, seq, it);
                if (pi.j == 0) {
                    const PtrPair src = layer_src(p, pi.layer);
                    cp_async16(park + (2 * u) * blockDim.x + threadIdx.x, src.k + pi.src_off);
                    cp_async16(park + (2 * u + 1) * blockDim.x + threadIdx.x, src.v + pi.src_off);
                }
            }
            if (word_valid(peeked, tag)) {
                word = peeked;
                known = piece_wait(p, tag, seq, word, keep);  // (valid: no wait)
            }
        }
        for (int u = 0; u < ni; ++u) {
            const int it = it0 + u * (int)blockDim.x;
            if (it >= it_end) break;
            const PieceItem pi = piece_item(p, seq, it);
            if ((!known || pi.j < keep.n) && !(peek_now && pi.j == 0)) {
                const PtrPair src = layer_src(p, pi.layer);
                cp_async16(park + (2 * u) * blockDim.x + threadIdx.x, src.k + pi.src_off);
                cp_async16(park + (2 * u + 1) * blockDim.x + threadIdx.x, src.v + pi.src_off);
            }
        }
        peek_now = false;
        if (!known) {
            if (!piece_wait(p, tag, seq, word, keep)) {
                cp_async_wait_all();
                return;
            }
            known = true;
        }
        cp_async_wait_all();
        const int64_t len_off = (int64_t)keep.len * p.row_bytes;
        for (int u = 0; u < ni; ++u) {
            const int it = it0 + u * (int)blockDim.x;
            if (it >= it_end) break;
            const PieceItem pi = piece_item(p, seq, it);
            if (pi.j < keep.nstore) {
                const PtrPair dst = layer_dst(p, pi.layer);
                const bool kept = pi.j < keep.n;
                *reinterpret_cast<uint4*>(const_cast<char*>(dst.k) + pi.dst_off + len_off) =
                    kept ? park[(2 * u) * blockDim.x + threadIdx.x] : z;
                *reinterpret_cast<uint4*>(const_cast<char*>(dst.v) + pi.dst_off + len_off) =
                    kept ? park[(2 * u + 1) * blockDim.x + threadIdx.x] : z;
            }
        }
    }
}

// Phase 2 (after the CTA has published its own result): wait for the word if needed, store kept rows and zeros.
SD_DEVINL void kv_piece_finish(const KvParams& p, const unsigned tag, PieceRegs& r, uint4* park) {
    PieceKeep keep;
    const bool ok = piece_wait(p, tag, r.seq, r.word, keep);
    cp_async_wait_all();  // (also on the failure path: no request may outlive the CTA's shared memory)
    if (!ok) return;
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    const int64_t len_off = (int64_t)keep.len * p.row_bytes;
#pragma unroll
    for (int u = 0; u < kCarryNI; ++u) {
        if (r.jl[u] < 0) continue;
        const int j = r.jl[u] & 0xff;
        if (j < keep.nstore) {
            const PtrPair dst = layer_dst(p, r.jl[u] >> 8);
            const bool kept = j < keep.n;
            *reinterpret_cast<uint4*>(const_cast<char*>(dst.k) + r.dst_off[u] + len_off) =
                kept ? park[(2 * u) * blockDim.x + threadIdx.x] : z;
            *reinterpret_cast<uint4*>(const_cast<char*>(dst.v) + r.dst_off[u] + len_off) =
                kept ? park[(2 * u + 1) * blockDim.x + threadIdx.x] : z;
        }
    }
    // pieces longer than carry_ni items per thread (small vocabularies: few verify CTAs per sequence)
    if (r.it + p.carry_ni * (int)blockDim.x < r.it_end)
        piece_items_run(p, tag, r.seq, r.it + p.carry_ni * (int)blockDim.x, r.it_end, r.word, park, p.carry_ni);
}

// whole CTA: wait until every sequence word of this launch is published (bounded)
SD_DEVINL void wait_all_words(const KvParams& p, const unsigned tag) {
    for (int bb = threadIdx.x; bb < p.B; bb += blockDim.x) {
        SpinGuard guard;
        while (!word_valid(ld_word64(p.flags + bb), tag)) {
            if (!guard.keep_waiting(kSpinAllFlags)) break;
            __nanosleep(100);
        }
    }
}

// Verify CTAs of the last `lag` sequences: nobody behind them carries their rows -- tail CTAs do, after the verify
// CTAs.  A tail CTA is dispatched late, when the memory system is still queueing ~20 MB of bulk copies (a DRAM
// round trip then takes ~3.5 us), so the rows it will want are pulled into L2 HERE, early and fire-and-forget
// (no register, no shared memory, no wait): the tail's fetches become L2 hits.
SD_DEVINL void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
SD_DEVINL void kv_piece_prefetch(const KvParams& p, const int seq, const int part, const int nparts) {
    const int per = (p.seq_items + nparts - 1) / nparts;  // (part, nparts) = this verify CTA's place among its sequence's CTAs
    const int base = part * per;
    const int it_end = base + per < p.seq_items ? base + per : p.seq_items;
    // one request per 128-byte line: 8 consecutive items (16-byte columns of one row) share a line
    for (int it = base + 8 * (int)threadIdx.x; it < it_end; it += 8 * (int)blockDim.x) {
        const PieceItem pi = piece_item(p, seq, it);
        const PtrPair src = layer_src(p, pi.layer);
        prefetch_l2(src.k + pi.src_off);
        prefetch_l2(src.v + pi.src_off);
    }
}

// Tail CTA t of the fused step: tail piece (t % tail_pieces) of sequence B - 1 - t / tail_pieces -- the LAST sequence
// first: its CTAs are dispatched (and have their rows parked) before its word can possibly be there, while the
// CTAs dispatched last serve the sequences whose words have long been published and never wait.  `park_items`
// items per thread parked in the CTA's whole shared memory, every offered row requested by the first instructions of
// the CTA (no dependence on the word).  Tail CTA 0 also serves the multi-GPU exchange.  The last tail CTA to retire
// closes the launch: once every sequence word is published -- i.e. every verify CTA has published its partial,
// before which it has consumed the epoch -- it advances the launch epoch.
SD_DEVINL void kv_tail_body(const KvParams& p, int t, unsigned long long* trace, uint4* park, const int park_items) {
    // (the launch epoch only changes when THIS launch closes: the value read here is the one the closing CTA advances)
    unsigned epoch_raw;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(epoch_raw) : "l"(p.epoch_word) : "memory");
    const unsigned tag = (epoch_raw + 1u) & 0xffffffu;
    // Ticket, drawn at the START of the CTA: "this tail CTA has read the launch epoch" (the increment depends on the
    // loaded value, so the atomic cannot be issued before the load has returned).  The CTA that draws the last ticket
    // -- the last tail CTA to start -- closes the launch when it is done.  The atomic's round trip overlaps the row
    // fetch; drawn at the END of the CTA (as before), ticket + all-words check + epoch re-load were three serial L2
    // round trips (~1 us) between the last row store and the end of the launch.
    unsigned ticket_prev = 0u;
    if (threadIdx.x == 0)
        ticket_prev = atomicAdd(reinterpret_cast<unsigned*>(p.ticket), epoch_raw == 0xffffffffu ? 0u : 1u);
    // With a multi-GPU exchange the FIRST tail CTA carries no rows: it is dispatched as soon as a verify CTA retires,
    // polls the records while they arrive and is done one NVLink hop after the last finalizer of the slowest rank --
    // in parallel with the row moves, not behind the wait for the last sequence's word (N = 8: the gather of 3072
    // words behind tail CTA 0's own rows cost ~2 us per step).
    const bool exchange_cta = p.xchg != nullptr && t == 0;
    if (p.xchg != nullptr) --t;
    // tail_order 1: the pieces of the EARLIEST tail sequence are dispatched first -- their words are long published, so they
    // fetch, store and retire while the last verify CTAs still stream, and the CTAs dispatched last (~4 us before the
    // last word) serve the last sequences: the stores spread over the second half of the launch instead of piling up
    // behind the last words (order 0: no tail CTA stored anything before 23.4 us of a 30.5 us launch).
    const int tseq = t / p.tail_pieces;
    const int seq = exchange_cta ? 0 : (p.tail_order == 1 ? p.B - p.lag + tseq : p.B - 1 - tseq), tp = exchange_cta ? 0 : t % p.tail_pieces;
    if (!exchange_cta) {
        const int tail_items = (p.seq_items + p.tail_pieces - 1) / p.tail_pieces;
        const int it0 = tp * tail_items;
        const int it_end = it0 + tail_items < p.seq_items ? it0 + tail_items : p.seq_items;
        unsigned long long word = 0ull;  // "unknown": request every offered row now, look at the word afterwards
        piece_items_run(p, tag, seq, it0 + (int)threadIdx.x, it_end, word, park, park_items, p.tail_peek != 0);
    }
    if (trace && threadIdx.x == 0) trace[1] = global_timer_ns();
    if (exchange_cta) {
        // Multi-GPU exchange, served by ONE tail CTA in parallel with the others' stores (not on the closing path).
        if (p.xchg_packed) {  // records pushed from here: the finalizers released them before publishing their words
            wait_all_words(p, tag);
            __syncthreads();
            __threadfence();
            exchange_push_cta(p.xchg, p.xchg_packed, p.B);
        }
        // Wait for this rank's own finalizers FIRST (local words, cheap to poll): in lock-step the peers' records are then
        // at most one NVLink hop away, so the record words -- lines that the peers are writing over NVLink -- are polled
        // for a microsecond, not for the whole second half of the kernel (256 threads hammering them from the moment
        // this CTA starts measured SLOWER at N = 8: 37.7 us per step against 35.9 us).  It is also what the epoch
        // advance needs: every local finalizer has read the exchange epoch.
        wait_all_words(p, tag);
        __syncthreads();
        const unsigned xepoch = exchange_gather_records(p.xchg, p.xchg_out);
        exchange_advance(p.xchg, xepoch);
    }
    __shared__ int s_last;
    if (threadIdx.x == 0) s_last = ticket_prev == p.total_ctas - 1u;
    __syncthreads();
    if (!s_last) return;
    // Closing CTA.  Every tail CTA has read the epoch (all tickets drawn); every verify CTA has once every sequence
    // word is published (it consumes the epoch before it publishes its partial).  Tail CTAs still running use the
    // tag they hold in registers, and the next launch waits for this whole grid (griddepcontrol.wait).
    wait_all_words(p, tag);
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned e = epoch_raw;  // (no second load: one L2 round trip less on the closing path)
        *reinterpret_cast<volatile unsigned*>(p.epoch_word) = (e + 1u == 0xffffffffu) ? 0u : e + 1u;
        *p.ticket = 0;
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
