// Kernel 1: fused verify (argmax over V for B*(K+1) rows + accepted-prefix + bonus token).
//
// Replaces the reference's verify_prefix (src/kernels/reference.py:13-56; Triton incumbent
// src/kernels/triton/verify.py:13-98, one serial program per sequence) with ONE streaming pass.
//
// Work decomposition.  A logits row is cut into UNITS of blockDim.x 16-byte vectors (one vector per
// thread).  A CTA reduces a SEGMENT = a run of consecutive units of one row:
//   variant 3 "bulk chunk" (default): one CTA per (row, chunk of <= 12 units); thread 0 fetches the
//           whole chunk with ONE cp.async.bulk (TMA 1-D) into shared memory, every thread waits on the
//           mbarrier and folds its vectors from shared memory.
//   variant 1 "persistent": grid = SMs x resident CTAs; the B*rows*units_per_row units are
//           split into gridDim.x contiguous, equal (+-1 unit) ranges, so every SM streams the same
//           number of bytes and there is no partial last wave; a range that crosses a row boundary is
//           processed as two segments.
//   variant 0 "tiled": one CTA per (row, chunk of VPT units), register-staged 128-bit loads.
//   variant 2 "tma ring": persistent grid, producer warp + mbarrier ring of cp.async.bulk copies.
// Per segment:
//   thread: 128-bit ld.global.nc loads, VPT of them issued before first use; running packed max
//           (max.NaN.{bf16x2,f16x2,f32}) + the index of the FIRST vector at which its maximum
//           appeared -- ~1 ALU op per element, so the kernel stays HBM-bound
//   CTA   : REDUX warp reduce of (order_key, min position) -> one partial per (row, segment)
//   thread: also keeps a copy of the vector that set its maximum, so the exact element is resolved
//           from registers -- no dependent re-load at the end of a segment
//   the partial is published with ONE 64-bit store {key, pos} (0 = "not yet"); the CTA that owns the
//           last unit of a sequence polls the partial slots of its rows (their producers were
//           dispatched before it and wait on nothing), merges them, writes accept_len / mask / target
//           ids / bonus / packed record and re-zeroes the slots.  No fence or atomic on the path of an
//           ordinary CTA, no second launch, no host trip.
//
// Exact torch.argmax semantics (first maximal index; NaN maximal, first NaN; -0 == +0) are kept
// by (a) canonical-NaN propagating max, (b) folding -0 to +0 once per vector, (c) tie-breaking
// every reduction level on the smallest position.
#include <stdlib.h>

#include "sd_kv_col.cuh"

namespace sd {

constexpr int kMaxChunks = 64;  // upper bound of S (partials per row) -> workspace size


struct VerifyTuning {
    int threads;
    int vpt;
};
static VerifyTuning g_tuning = {0, 0};
static int g_variant = 3;       // default: bulk chunks (fastest measured, schedule independent of ptxas)
static unsigned long long* g_trace = nullptr;  // device buffer for per-CTA timestamps (variant 0 only)
static size_t g_trace_ctas = 0;
static int g_ctas_per_sm = 0;   // persistent grid = SMs * this (0 = occupancy)

struct VerifyParams {
    const void* logits;
    int64_t stride_b, stride_r;
    const void* ids;
    int id_dtype;
    int64_t id_stride_b;
    int32_t* accept_len;
    uint8_t* mask;
    int32_t* target_ids;
    int32_t* bonus;
    int32_t* packed;
    float* token_prob;
    float* row_max;
    float* row_sumexp;
    // partial records, [B*rows*S] x 16 bytes: {u64 kp = order key << 32 | position (0 = not yet
    // published), float m, float s}.  One array of records (not three arrays) so that a workspace
    // shared by calls of different (B, rows) never shows one call's statistics as another's kp.
    unsigned long long* part;
    int B, rows, K, V;
    int S;                // partial slots reserved per row
    int upr;              // units per row (unit = blockDim.x vectors)
    int units_per_chunk;  // tiled variant: units per CTA
    int U, q, rem;        // persistent variant: U = B*rows*upr units = gridDim.x*q + rem
    int mode;
    float threshold, scale_log2;  // scale_log2 = inv_temperature * log2(e)
    float inv_temperature;
    const sd_exchange* xchg;    // nullable: peer-memory result exchange (device memory)
    int32_t* seq_ready;         // nullable: flag b is raised once sequence b's outputs are final
    // lagged fused step (verify_step_kernel): the finalizer publishes the sequence word {tag, n_keep, old length}
    // and stores the new length; all null / zero otherwise (sd_kv_col.cuh)
    unsigned long long* words;
    unsigned* epoch_word;
    int32_t* seq_lens;
    int32_t* kv_overflow;
    int64_t kv_capacity;
    int kv_R;
    int word_release;           // 1 = a CTA reads the packed records after seeing the words (push from the append part)
    unsigned long long* trace;  // evidence runs: [grid][4] = {globaltimer at entry, after the scan, at exit, smid}
};

struct Cand {
    uint32_t key;
    int pos;
};

SD_DEVINL Cand better(Cand a, Cand b) {
    if (b.key > a.key || (b.key == a.key && b.pos < a.pos)) return b;
    return a;
}

SD_DEVINL Cand warp_best(Cand c) {
    uint32_t kmax = __reduce_max_sync(0xffffffffu, c.key);
    uint32_t p = (c.key == kmax) ? (uint32_t)c.pos : 0x7fffffffu;
    p = __reduce_min_sync(0xffffffffu, p);
    return Cand{kmax, (int)p};
}

// (m, s) pairs of an online softmax: value = s * 2^m (m already scaled to log2 domain)
SD_DEVINL void merge_ms(float& m, float& s, float m2, float s2) {
    float mn = fmaxf(m, m2);
    float a = (m == -INFINITY) ? 0.f : s * fast_exp2(m - mn);
    float b = (m2 == -INFINITY) ? 0.f : s2 * fast_exp2(m2 - mn);
    m = mn;
    s = a + b;
}

// contiguous partition of U units over G CTAs: CTA c owns [part_start(c), part_start(c+1))
SD_DEVINL int part_start(const VerifyParams& p, int c) { return c * p.q + (c < p.rem ? c : p.rem); }
SD_DEVINL int part_owner(const VerifyParams& p, int u) {
    const int big = p.rem * (p.q + 1);
    return u < big ? u / (p.q + 1) : p.rem + (u - big) / p.q;
}

// ---- row geometry: 16-byte aligned interior + scalar peel (rows of e.g. V=50257 bf16 start at
// arbitrary 2-byte offsets) ----
template <typename T>
struct RowGeom {
    const T* base;     // element 0 of the row
    const uint4* vp;   // first aligned 16-byte vector (element `head`)
    int head;          // scalar elements before the first vector
    int nvec;          // whole vectors
    int tail0;         // first scalar element after the last vector
};

template <typename T>
SD_DEVINL RowGeom<T> row_geom(const VerifyParams& p, int row) {
    constexpr int EPV = 16 / (int)sizeof(T);
    const int b = row / p.rows;
    const int j = row - b * p.rows;
    RowGeom<T> g;
    g.base = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)j * p.stride_r;
    const int mis = (int)((reinterpret_cast<uintptr_t>(g.base) & 15u) / sizeof(T));
    int head = mis ? (EPV - mis) : 0;
    if (head > p.V) head = p.V;
    g.head = head;
    g.nvec = (p.V - head) / EPV;
    g.tail0 = head + g.nvec * EPV;
    g.vp = reinterpret_cast<const uint4*>(g.base + head);
    return g;
}

// ---- per-thread running state over the vectors of one row segment ----
template <typename T, bool STATS>
struct ThreadAcc {
    uint32_t run;     // replicated scalar running max (packed)
    int first_it;     // vector index (within the row) where `run` first appeared
    uint4 bestv;      // that vector (the exact element is resolved from it at the end of the segment)
    float sm, ss;     // STATS: reference max (log2 domain) and sum of 2^(x - sm); top-k mode: sm = 0, ss = count
    float xtok;       // STATS, top-k mode: logit of the draft token (count the elements above it); NaN = softmax mode

    SD_DEVINL void reset() {
        run = Packed<T>::kNegInf;
        first_it = 0x7fffffff;
        bestv = make_uint4(0u, 0u, 0u, 0u);
        sm = -INFINITY;
        ss = 0.f;
        xtok = __uint_as_float(0x7fc00000u);
    }

    // top-k agreement mode: count logits strictly above `x` instead of accumulating the softmax sum
    SD_DEVINL void count_above(float x) {
        xtok = x;
        sm = 0.f;
    }
    SD_DEVINL bool counting() const { return xtok == xtok; }

    // fold one 16-byte vector (index `idx` in its row); `live` = the vector exists
    SD_DEVINL void add(const uint4& v, int idx, bool live, float scale_log2) {
        using P = Packed<T>;
        uint32_t t = P::vmax(P::vmax(v.x, v.y), P::vmax(v.z, v.w));
        t = P::fold(t);
        const uint32_t nw = P::vmax(run, t);
        // first live vector always claims the position (run == -inf may equal a row of -inf)
        if (live && (nw != run || first_it == 0x7fffffff)) {
            first_it = idx;
            bestv = v;
        }
        if constexpr (STATS) {
            if (counting()) {
                if (live) {
                    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if constexpr (sizeof(T) == 4) {
                            ss += __uint_as_float(w[q]) > xtok ? 1.f : 0.f;
                        } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                            ss += __uint_as_float(w[q] << 16) > xtok ? 1.f : 0.f;
                            ss += __uint_as_float(w[q] & 0xffff0000u) > xtok ? 1.f : 0.f;
                        } else {
                            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
                            ss += f.x > xtok ? 1.f : 0.f;
                            ss += f.y > xtok ? 1.f : 0.f;
                        }
                    }
                }
                run = nw;
                return;
            }
            if (live && nw != run) {
                float mn = P::scalar(nw) * scale_log2;
                ss = (sm == -INFINITY) ? 0.f : ss * fast_exp2(sm - mn);
                sm = mn;
            }
            if (live && sm != -INFINITY) {  // a thread that has only seen -inf adds nothing
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if constexpr (sizeof(T) == 4) {
                        ss += fast_exp2(fmaf(__uint_as_float(w[q]), scale_log2, -sm));
                    } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                        ss += fast_exp2(fmaf(__uint_as_float(w[q] << 16), scale_log2, -sm));
                        ss += fast_exp2(fmaf(__uint_as_float(w[q] & 0xffff0000u), scale_log2, -sm));
                    } else {
                        float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
                        ss += fast_exp2(fmaf(f.x, scale_log2, -sm));
                        ss += fast_exp2(fmaf(f.y, scale_log2, -sm));
                    }
                }
            }
        }
        run = nw;
    }
};

// top-k agreement (policies.py:300-314): logit of the draft token of global row `row`, +inf for the
// bonus row / an out-of-range id (nothing is above +inf -> count 0)
template <typename T>
SD_DEVINL float draft_logit(const VerifyParams& p, const RowGeom<T>& g, int row) {
    const int b = row / p.rows, j = row - b * p.rows;
    if (j >= p.K) return INFINITY;
    const int tok = load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + j);
    return (tok >= 0 && tok < p.V) ? to_float<T>(g.base[tok]) : INFINITY;
}

// first maximal element of one 16-byte vector whose first element has row index `pos0`
template <typename T>
SD_DEVINL Cand exact_in_vector(const uint4& v, int pos0) {
    const uint32_t w[4] = {v.x, v.y, v.z, v.w};
    Cand c{0u, 0x7fffffff};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        if constexpr (sizeof(T) == 4) {
            c = better(c, Cand{order_key(__uint_as_float(w[q])), pos0 + q});
        } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
            c = better(c, Cand{order_key(__uint_as_float(w[q] << 16)), pos0 + 2 * q});
            c = better(c, Cand{order_key(__uint_as_float(w[q] & 0xffff0000u)), pos0 + 2 * q + 1});
        } else {
            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
            c = better(c, Cand{order_key(f.x), pos0 + 2 * q});
            c = better(c, Cand{order_key(f.y), pos0 + 2 * q + 1});
        }
    }
    return c;
}

SD_DEVINL unsigned long long ld_volatile_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
SD_DEVINL unsigned ld_volatile_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
SD_DEVINL void st_volatile_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// CTA-wide barrier over the `nthreads` reducing threads: __syncthreads, or named barrier 1 when the
// CTA also holds a producer warp that must not take part
template <bool NAMED>
SD_DEVINL void cta_sync(int nthreads) {
    if constexpr (NAMED) {
        asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
    } else {
        __syncthreads();
    }
}

// Close the segment [it0, it1) of logits row `row` (global row index b*rows + j): fold the scalar peel
// elements, reduce the CTA and publish the partial in `slot` of that row.  Returns true when the
// segment holds the last unit of its sequence: that CTA must call finalize_sequence(b) -- right away
// in the tiled variants (one segment per CTA), after its whole range in the persistent ones (a CTA
// that stopped to wait in the middle of its range would stall the CTAs waiting for ITS later
// partials).  Called by all `nthreads` reducing threads (tid in [0, nthreads)); contains barriers.
template <typename T, bool STATS, bool NAMED>
SD_DEVINL bool publish_segment(const VerifyParams& p, const RowGeom<T>& g, ThreadAcc<T, STATS>& acc, const int row,
                               const int it0, const int it1, const int slot, const int tid, const int nthreads) {
    using P = Packed<T>;
    constexpr int EPV = 16 / (int)sizeof(T);
    const int b = row / p.rows;
    const int V = p.V;
    const T* base = g.base;
    float sm = acc.sm, ss = acc.ss;

    // exact (key, position) of this thread's maximum, from the vector it kept
    Cand c{0u, 0x7fffffff};  // nothing seen
    if (acc.first_it != 0x7fffffff) c = exact_in_vector<T>(acc.bestv, g.head + acc.first_it * EPV);

    // scalar peel elements (head + tail): handled by the segment that owns unit 0 of the row
    if (it0 == 0) {
        const int npeel = g.head + (V - g.tail0);
        for (int e = tid; e < npeel; e += nthreads) {
            int pos = e < g.head ? e : g.tail0 + (e - g.head);
            float x = to_float<T>(base[pos]);
            Cand c2{order_key(x), pos};
            c = better(c, c2);
            if constexpr (STATS) {
                if (acc.counting()) {
                    ss += x > acc.xtok ? 1.f : 0.f;
                } else {
                    float xm = x * p.scale_log2;
                    merge_ms(sm, ss, xm, 1.f);
                }
            }
        }
    }

    // ---- CTA reduce ----
    __shared__ uint32_t s_key[32];
    __shared__ int s_pos[32];
    __shared__ float s_m[32];
    __shared__ float s_s[32];
    const int lane = tid & 31, wid = tid >> 5, nwarps = nthreads >> 5;

    c = warp_best(c);
    if constexpr (STATS) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            float m2 = __shfl_xor_sync(0xffffffffu, sm, o);
            float s2 = __shfl_xor_sync(0xffffffffu, ss, o);
            merge_ms(sm, ss, m2, s2);
        }
    }
    if (lane == 0) {
        s_key[wid] = c.key;
        s_pos[wid] = c.pos;
        if constexpr (STATS) {
            s_m[wid] = sm;
            s_s[wid] = ss;
        }
    }
    cta_sync<NAMED>(nthreads);
    if (wid == 0) {
        Cand c2 = lane < nwarps ? Cand{s_key[lane], s_pos[lane]} : Cand{0u, 0x7fffffff};
        c2 = warp_best(c2);
        float m2 = -INFINITY, s2 = 0.f;
        if constexpr (STATS) {
            if (lane < nwarps) {
                m2 = s_m[lane];
                s2 = s_s[lane];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float m3 = __shfl_xor_sync(0xffffffffu, m2, o);
                float s3 = __shfl_xor_sync(0xffffffffu, s2, o);
                merge_ms(m2, s2, m3, s3);
            }
        }
        if (lane == 0) {
            unsigned long long* rec = p.part + 2 * ((size_t)row * p.S + slot);
            if constexpr (STATS) {
                *reinterpret_cast<float2*>(rec + 1) = make_float2(m2, s2);
                __threadfence();  // the statistics must be visible before the slot reads "published"
            }
            // {key, pos} is never 0: an empty candidate carries pos 0x7fffffff
            st_volatile_u64(rec, ((unsigned long long)c2.key << 32) | (unsigned)c2.pos);
        }
    }
    cta_sync<NAMED>(nthreads);  // s_key .. s_s are rewritten by the caller's next segment
    return (row - b * p.rows == p.rows - 1) && (it1 == p.upr);
}

// Finalize sequence b: wait for the partials of its rows, merge them, evaluate the acceptance rule,
// write the outputs and re-arm the slots.  `nslots` = partial slots in use per row for tiled launches,
// 0 = derive them from the persistent partition.  Called by all `nthreads` reducing threads.
template <typename T, bool STATS, bool NAMED>
SD_DEVINL void finalize_sequence(const VerifyParams& p, const int b, const int nslots, const int tid, const int nthreads,
                                 const unsigned tag = 0u) {
    const int V = p.V;
    const int lane = tid & 31, wid = tid >> 5, nwarps = nthreads >> 5;
    __shared__ int s_tgt[64];
    __shared__ float s_ptok[64];
    // draft ids of this sequence: issued now, consumed by warp 0 after the row merge (K <= 64)
    int my_id[2] = {-1, -1};
    // multi-GPU exchange: the descriptor fields and the exchange epoch are requested NOW (three dependent L2 round
    // trips that hide behind the partial poll below) and only used once the record is final
    int x_world = 0, x_width = 0, x_rank = 0, x_rpr = 0;
    unsigned x_epoch = 0u;
    // fused step: the sequence's length before this step, requested NOW as well (this thread is its only reader and
    // writer in the launch) -- loaded where it is used, it was one more L2 round trip between "accept_len known" and
    // "word published", i.e. on the critical path of every consumer of the word and of the launch's tail
    int len_before = 0;
    if (wid == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h * 32 + lane < p.K) my_id[h] = load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + h * 32 + lane);
        if (lane == 0 && p.words && p.seq_lens) len_before = p.seq_lens[b];
        if (p.xchg) {
            const sd_exchange* x = p.xchg;
            x_world = x->world;
            x_width = x->width;
            x_rank = x->rank;
            x_rpr = x->rows_per_rank;
            x_epoch = ld_volatile_u32(reinterpret_cast<const unsigned*>(x->peer[x_rank]));
        }
    }
    for (int r = wid; r < p.rows; r += nwarps) {
        const int grow = b * p.rows + r;
        const T* rb = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)r * p.stride_r;
        int tok = -1;
        if constexpr (STATS) {
            if (r < p.K && p.mode != SD_ACCEPT_CONFIDENCE) tok = load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + r);
        }
        int ns = nslots;
        if (ns == 0) ns = part_owner(p, (grow + 1) * p.upr - 1) - part_owner(p, grow * p.upr) + 1;
        unsigned long long* slots = p.part + 2 * ((size_t)grow * p.S);  // record i at slots[2 * i]
        // ns <= kMaxChunks = 64: two slots per lane.  Their producers were dispatched before this CTA
        // (lower block index) and wait on nothing, so the poll terminates.
        unsigned long long v0 = 1ull, v1 = 1ull;
        SpinGuard guard;  // (all lanes poll the same number of times: the verdict is warp-uniform)
        for (;;) {
            if (lane < ns) v0 = ld_volatile_u64(slots + 2 * lane);
            if (lane + 32 < ns) v1 = ld_volatile_u64(slots + 2 * (lane + 32));
            if (__all_sync(0xffffffffu, v0 != 0ull && v1 != 0ull)) break;
            if (!__all_sync(0xffffffffu, guard.keep_waiting(kSpinPartials))) break;
            __nanosleep(40);
        }
        Cand best{0u, 0x7fffffff};
        if (lane < ns) best = better(best, Cand{(uint32_t)(v0 >> 32), (int)(uint32_t)v0});
        if (lane + 32 < ns) best = better(best, Cand{(uint32_t)(v1 >> 32), (int)(uint32_t)v1});
        float fm = -INFINITY, fs = 0.f;
        if constexpr (STATS) {
            __threadfence();
            for (int sidx = lane; sidx < ns; sidx += 32) {
                const volatile float* ms = reinterpret_cast<const volatile float*>(slots + 2 * sidx + 1);
                float m2 = ms[0];
                float s2 = ms[1];
                merge_ms(fm, fs, m2, s2);
            }
        }
        // re-arm the slots for the next launch / graph replay (this CTA is their only reader)
        if (lane < ns) slots[2 * lane] = 0ull;
        if (lane + 32 < ns) slots[2 * (lane + 32)] = 0ull;
        float xtok = 0.f;
        if constexpr (STATS) {
            if (tok >= 0 && tok < V) xtok = to_float<T>(rb[tok]);
        }
        best = warp_best(best);
        if constexpr (STATS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float m2 = __shfl_xor_sync(0xffffffffu, fm, o);
                float s2 = __shfl_xor_sync(0xffffffffu, fs, o);
                merge_ms(fm, fs, m2, s2);
            }
        }
        const int idx = best.pos == 0x7fffffff ? 0 : best.pos;
        if (lane == 0) {
            s_tgt[r] = idx;
            if (p.target_ids) p.target_ids[grow] = idx;
            if constexpr (STATS) {
                if (p.row_max) p.row_max[grow] = key_to_float(best.key) * p.inv_temperature;
                if (p.row_sumexp) p.row_sumexp[grow] = fs;
                float pt = 0.f;
                if (r < p.K) {
                    if (p.mode == SD_ACCEPT_CONFIDENCE) {
                        pt = 1.0f / fs;
                    } else if (p.mode == SD_ACCEPT_TOPK) {
                        pt = (tok >= 0 && tok < V) ? fs : INFINITY;  // logits strictly above the draft token's
                    } else if (tok >= 0 && tok < V) {
                        pt = fast_exp2(xtok * p.scale_log2 - fm) / fs;
                    }
                    if (p.token_prob) p.token_prob[b * p.K + r] = pt;
                }
                s_ptok[r] = pt;
            }
        }
    }
    cta_sync<NAMED>(nthreads);
    int nacc = p.K;  // (meaningful in warp 0)
    if (wid == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = h * 32 + lane;
            bool bad = false;
            if (k < p.K) {
                if (STATS && p.mode == SD_ACCEPT_TOPK) bad = !(s_ptok[k] < p.threshold);
                else if (STATS && p.mode != SD_ACCEPT_GREEDY) bad = !(s_ptok[k] >= p.threshold);
                else bad = s_tgt[k] != my_id[h];
            }
            const unsigned ballot = __ballot_sync(0xffffffffu, bad);
            if (ballot && nacc == p.K) nacc = h * 32 + __ffs(ballot) - 1;
        }
        for (int k = lane; k < p.K; k += 32) {
            if (p.mask) p.mask[b * p.K + k] = (uint8_t)(k < nacc);
        }
        if (p.packed) {
            int32_t* rec = p.packed + (int64_t)b * (p.K + 2);
            for (int r = lane; r <= p.K; r += 32) rec[1 + r] = r < p.rows ? s_tgt[r] : -1;
            if (lane == 0) rec[0] = nacc;
        }
        if (lane == 0) {
            p.accept_len[b] = nacc;
            if (p.bonus) p.bonus[b] = p.rows > p.K ? s_tgt[nacc] : -1;
        }
    }
    // (a value that is 0 at run time but depends on the exchange epoch: folding it into the sequence word below makes the
    // word's store wait for the epoch load -- "this finalizer has READ the exchange epoch" is then implied by its word)
    const unsigned x_seen = x_epoch == 0xffffffffu ? 0x80000000u : 0u;
    if (wid == 0 && p.words) {
        // lagged fused step: publish {tag, n_keep, old length} in ONE 8-byte store and advance the length (this
        // thread is the only reader and writer of seq_lens[b] in the launch).  BEFORE the peer push: the append of
        // this sequence must not wait for NVLink stores; the exchange epoch was read above, which is all the
        // gathering CTA needs to know before it advances that epoch.
        __syncwarp();
        if (lane == 0) {
            const int len = len_before;
            const int n_keep = nacc + 1;
            int64_t room = p.kv_capacity - len;
            room = room < 0 ? 0 : (room > p.kv_R ? p.kv_R : room);
            int n = n_keep > p.kv_R ? p.kv_R : n_keep;
            if (n > room) {
                if (p.kv_overflow) *p.kv_overflow = 1;
                n = (int)room;
            }
            if (p.seq_lens) p.seq_lens[b] = len + n;
            if (p.word_release) __threadfence();
            st_volatile_u64(p.words + b, make_word(tag, n_keep > p.kv_R ? p.kv_R : n_keep, (int)((unsigned)len | x_seen)));
        }
    }
    if (wid == 0 && p.xchg) {
        // Fused result exchange: the record goes straight into every rank's buffer (NVLink stores on peer
        // pointers).  Each value travels in ONE 8-byte store together with the epoch tag (LL-style:
        // {epoch + 1, value}), so no fence, atomic or arrival counter is needed -- the consumer polls the words
        // until their tag matches.
        const sd_exchange* x = p.xchg;
        const unsigned long long xtag = (unsigned long long)(x_epoch + 1u) << 32;
        const size_t rec_row = (size_t)(x_epoch & 3u) * x_world * x_rpr + (size_t)x_rank * x_rpr + b;
        for (int i = lane; i < x_world * x_width; i += 32) {
            const int pr = i / x_width, col = i - pr * x_width;
            int val = -1;
            if (col == 0) val = nacc;
            else if (col - 1 < p.rows) val = s_tgt[col - 1];
            unsigned long long* recs =
                reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(x->peer[pr]) + kExchangeHeaderBytes);
            recs[rec_row * x_width + col] = xtag | (unsigned)val;  // weak store: posted, pipelined over NVLink
        }
    }
    if (wid == 0 && p.seq_ready) {
        // per-sequence completion flag: lets the kv_append launch that follows start on sequence b while other
        // sequences are still being verified (sd_kv_append, seq_ready)
        __syncwarp();
        if (lane == 0) {
            __threadfence();
            asm volatile("st.volatile.global.s32 [%0], %1;" ::"l"(p.seq_ready + b), "r"(1) : "memory");
        }
    }
    cta_sync<NAMED>(nthreads);  // s_tgt and the reduce scratch are reused by the caller's next segment
}

// Reduce units [it0, it1) of logits row `row` with register-staged 128-bit loads (VPT in flight per
// thread) and close the segment.  Called by all threads of the CTA.
template <typename T, int VPT, bool STATS>
SD_DEVINL bool process_segment(const VerifyParams& p, const int row, const int it0, const int it1, const int slot) {
    using P = Packed<T>;
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const RowGeom<T> g = row_geom<T>(p, row);
    ThreadAcc<T, STATS> acc;
    acc.reset();
    if constexpr (STATS) {
        if (p.mode == SD_ACCEPT_TOPK) acc.count_above(draft_logit<T>(p, g, row));
    }
    for (int i0 = it0; i0 < it1; i0 += VPT) {
        uint4 buf[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int idx = (i0 + u) * nthreads + tid;
            if (i0 + u < it1 && idx < g.nvec) {
                buf[u] = ldg_stream_v4(g.vp + idx);
            } else {
                buf[u] = make_uint4(P::kNegInf, P::kNegInf, P::kNegInf, P::kNegInf);
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) pin_loaded(buf[u]);
        __syncwarp();  // memory-ordering point: ptxas keeps all VPT loads of the batch ahead of their first use
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int idx = (i0 + u) * nthreads + tid;
            acc.add(buf[u], idx, i0 + u < it1 && idx < g.nvec, p.scale_log2);
        }
    }
    if (p.trace && tid == 0) p.trace[(size_t)blockIdx.x * 4 + 1] = global_timer_ns();
    return publish_segment<T, STATS, false>(p, g, acc, row, it0, it1, slot, tid, nthreads);
}

// variant 0: one CTA per (row, chunk)
template <typename T, int VPT, bool STATS>
// launch bounds: (512, 3) makes ptxas settle on ~40 registers and keep all VPT loads of a batch in
// flight before the first use; with a laxer or tighter bound it serialises load -> use (SASS checked)
__global__ void __launch_bounds__(512, STATS ? 2 : 3) verify_kernel(const VerifyParams p) {
    pdl_wait();
    pdl_trigger();
    const int unit = blockIdx.x;
    if (p.trace && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        p.trace[(size_t)unit * 4 + 0] = global_timer_ns();
        p.trace[(size_t)unit * 4 + 3] = smid;
    }
    const int chunk = unit % p.S;
    const int row = unit / p.S;
    const int it0 = chunk * p.units_per_chunk;
    int it1 = it0 + p.units_per_chunk;
    if (it1 > p.upr) it1 = p.upr;
    if (process_segment<T, VPT, STATS>(p, row, it0, it1, chunk))
        finalize_sequence<T, STATS, false>(p, row / p.rows, p.S, threadIdx.x, blockDim.x);
    if (p.trace && threadIdx.x == 0) p.trace[(size_t)unit * 4 + 2] = global_timer_ns();
}

// variant 1: persistent, equal contiguous unit ranges
template <typename T, int VPT, bool STATS>
__global__ void __launch_bounds__(512, STATS ? 2 : 3) verify_persist_kernel(const VerifyParams p) {
    pdl_wait();
    pdl_trigger();
    const int c = blockIdx.x;
    int u = part_start(p, c);
    const int end = part_start(p, c + 1);
    while (u < end) {
        const int row = u / p.upr;
        const int it0 = u - row * p.upr;
        int seg_end = (row + 1) * p.upr;
        if (seg_end > end) seg_end = end;
        process_segment<T, VPT, STATS>(p, row, it0, it0 + (seg_end - u), c - part_owner(p, row * p.upr));
        u = seg_end;
    }
    // sequences whose last unit lies in this CTA's range (all partials of this CTA are published)
    const int su = p.rows * p.upr;
    for (int b = part_start(p, c) / su; b <= end / su - 1; ++b)
        finalize_sequence<T, STATS, false>(p, b, 0, threadIdx.x, blockDim.x);
}

// ---- variant 2: persistent + TMA bulk staging -----------------------------------------------------
// One producer warp keeps `stages` cp.async.bulk (TMA 1-D) copies of UNIT = 256*VPU vectors in flight
// into a shared-memory ring (mbarrier full/empty pairs); 256 consumer threads fold each landed unit
// from shared memory.  Loads stay in flight while the consumers run a segment epilogue, and no
// registers are tied up by outstanding loads.
constexpr int kTmaConsumers = 256;
constexpr int kTmaMaxStages = 32;

SD_DEVINL uint32_t smem_u32v(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
SD_DEVINL void vbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32v(bar)), "r"(count) : "memory");
}
SD_DEVINL void vbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32v(bar)) : "memory");
}
SD_DEVINL void vbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32v(bar)), "r"(bytes) : "memory");
}
SD_DEVINL void vbar_wait(uint64_t* bar, unsigned parity) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32v(bar)), "r"(parity)
            : "memory");
    }
}
SD_DEVINL void vbulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32v(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(smem_u32v(bar))
                 : "memory");
}

template <typename T, int VPU, bool STATS>
__global__ void __launch_bounds__(kTmaConsumers + 32) verify_tma_kernel(const VerifyParams p, const int stages) {
    using P = Packed<T>;
    constexpr int UV = kTmaConsumers * VPU;  // vectors per unit
    extern __shared__ __align__(128) unsigned char ring[];
    __shared__ __align__(8) uint64_t full_bar[kTmaMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kTmaMaxStages];
    const int tid = threadIdx.x;
    if (tid == 0) {
        for (int s = 0; s < stages; ++s) {
            vbar_init(&full_bar[s], 1);
            vbar_init(&empty_bar[s], kTmaConsumers / 32);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();

    const int c = blockIdx.x;
    const int u_begin = part_start(p, c);
    const int u_end = part_start(p, c + 1);

    if (tid >= kTmaConsumers) {  // ---- producer warp: one elected lane issues the bulk copies ----
        if (tid == kTmaConsumers) {
            int s = 0;
            unsigned ph = 0;
            int row = -1;
            RowGeom<T> g;
            for (int u = u_begin; u < u_end; ++u) {
                const int r = u / p.upr;
                if (r != row) {
                    row = r;
                    g = row_geom<T>(p, row);
                }
                const int v0 = (u - row * p.upr) * UV;
                int nv = g.nvec - v0;
                nv = nv < 0 ? 0 : (nv > UV ? UV : nv);
                vbar_wait(&empty_bar[s], ph ^ 1u);  // slot free (first lap passes immediately)
                if (nv > 0) {
                    vbar_arrive_expect_tx(&full_bar[s], (unsigned)nv * 16u);
                    vbulk_g2s(ring + (size_t)s * (UV * 16), g.vp + v0, (unsigned)nv * 16u, &full_bar[s]);
                } else {
                    vbar_arrive(&full_bar[s]);
                }
                if (++s == stages) {
                    s = 0;
                    ph ^= 1u;
                }
            }
        }
        return;
    }

    // ---- consumers ----
    const int lane = tid & 31;
    int s = 0;
    unsigned ph = 0;
    int u = u_begin;
    while (u < u_end) {
        const int row = u / p.upr;
        const int it0 = u - row * p.upr;
        int seg_end = (row + 1) * p.upr;
        if (seg_end > u_end) seg_end = u_end;
        const RowGeom<T> g = row_geom<T>(p, row);
        ThreadAcc<T, STATS> acc;
        acc.reset();
        if constexpr (STATS) {
            if (p.mode == SD_ACCEPT_TOPK) acc.count_above(draft_logit<T>(p, g, row));
        }
        for (int uu = u; uu < seg_end; ++uu) {
            const int v0 = (uu - row * p.upr) * UV;
            vbar_wait(&full_bar[s], ph);
            const uint4* src = reinterpret_cast<const uint4*>(ring + (size_t)s * (UV * 16));
            uint4 buf[VPU];
#pragma unroll
            for (int k = 0; k < VPU; ++k) {
                const int idx = v0 + k * kTmaConsumers + tid;
                buf[k] = idx < g.nvec ? src[k * kTmaConsumers + tid] : make_uint4(P::kNegInf, P::kNegInf, P::kNegInf, P::kNegInf);
            }
#pragma unroll
            for (int k = 0; k < VPU; ++k) pin_loaded(buf[k]);
            __syncwarp();
            if (lane == 0) vbar_arrive(&empty_bar[s]);  // this warp's reads of the slot are done
#pragma unroll
            for (int k = 0; k < VPU; ++k) {
                const int idx = v0 + k * kTmaConsumers + tid;
                acc.add(buf[k], idx, idx < g.nvec, p.scale_log2);
            }
            if (++s == stages) {
                s = 0;
                ph ^= 1u;
            }
        }
        publish_segment<T, STATS, true>(p, g, acc, row, it0, it0 + (seg_end - u), c - part_owner(p, row * p.upr), tid,
                                        kTmaConsumers);
        u = seg_end;
    }
    const int su = p.rows * p.upr;
    for (int b = u_begin / su; b <= u_end / su - 1; ++b) finalize_sequence<T, STATS, true>(p, b, 0, tid, kTmaConsumers);
}

// ---- variant 3: one CTA per (row, chunk), the whole chunk fetched by ONE cp.async.bulk -------------
// Thread 0 issues a single TMA 1-D copy of the chunk (upc * blockDim.x vectors, up to ~48 KB) into
// shared memory and every thread waits on the mbarrier; the chunk is then folded from shared memory.
// All of a CTA's bytes are in flight after one instruction, independent of register allocation, and
// several CTAs per SM overlap each other's wait / fold / publish phases.
template <typename T, bool STATS>
SD_DEVINL void verify_bulk_body(const VerifyParams& p, const int unit, unsigned char* stage, uint64_t* bar) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int chunk = unit % p.S;
    const int row = unit / p.S;
    const int it0 = chunk * p.units_per_chunk;
    int it1 = it0 + p.units_per_chunk;
    if (it1 > p.upr) it1 = p.upr;
    const RowGeom<T> g = row_geom<T>(p, row);
    const int v0 = it0 * nthreads;
    int nv = g.nvec - v0;
    const int cap = (it1 - it0) * nthreads;
    nv = nv < 0 ? 0 : (nv > cap ? cap : nv);
    if (tid == 0) {
        if (nv > 0) {
            vbar_arrive_expect_tx(bar, (unsigned)nv * 16u);
            vbulk_g2s(stage, g.vp + v0, (unsigned)nv * 16u, bar);
        } else {
            vbar_arrive(bar);
        }
    }
    ThreadAcc<T, STATS> acc;
    acc.reset();
    if constexpr (STATS) {  // overlaps the bulk copy
        if (p.mode == SD_ACCEPT_TOPK) acc.count_above(draft_logit<T>(p, g, row));
    }
    vbar_wait(bar, 0u);
    const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll 4
    for (int k = tid; k < nv; k += nthreads) acc.add(src[k], v0 + k, true, p.scale_log2);
    if (publish_segment<T, STATS, false>(p, g, acc, row, it0, it1, chunk, tid, nthreads))
        finalize_sequence<T, STATS, false>(p, row / p.rows, p.S, tid, nthreads);
}

template <typename T, bool STATS>
__global__ void __launch_bounds__(256) verify_bulk_kernel(const VerifyParams p) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        vbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    verify_bulk_body<T, STATS>(p, (int)blockIdx.x, stage, &bar);
}

// ---- the whole step in ONE launch: fused verify + kv_append ------------------------------------------
// Default ("lagged append", verify_step_kernel): CTAs [0, n_verify) each reduce one logits chunk AND carry one
// piece of the append of the sequence `lag` places earlier (sd_kv_col.cuh): the piece's rows are requested right
// after the chunk's bulk copy, so both are in flight together, and stored after the CTA's partial is published.
// The append traffic therefore streams UNDER the verify phase instead of after it, and nothing on the verify
// critical path waits for an append.  CTAs [n_verify, n_verify + lag * pieces) are the tail: the pieces of the
// last `lag` sequences, the multi-GPU gather and the epoch advance.
// Progress: a CTA only ever waits for words / partials produced by CTAs with a LOWER block index (the finalizer of
// an earlier sequence; the chunks of its own sequence), which wait, transitively, only on still lower ones.
// REQUIREMENT (stated in include/sd_verify.h): the hardware dispatches the blocks of a 1-D grid in index order, so
// every CTA a resident CTA waits for is resident or finished.  Every wait is bounded (SpinGuard).
template <typename T>
SD_DEVINL void step_cta_body(const VerifyParams& p, const KvParams& kp, const int unit, unsigned char* stage, uint64_t* bar,
                             uint4* park) {
    const int tid = threadIdx.x, nthreads = blockDim.x;
    // Sequences [0, lag) have nothing to carry (no sequence `lag` places before them): their CTAs use the WHOLE shared
    // memory of the CTA (stage + park) for a bigger logits chunk -- upc_big units, S_big chunks per row.  Throughput here
    // is bytes in flight / loaded latency (~5.5 us): with 36 KB per CTA the first phase of the launch streamed at
    // ~3.4 TB/s (21 MB in flight), with stage + park = 52 KB it streams like the carrying phase.
    const bool big = unit < kp.n_big;
    const int S = big ? kp.S_big : p.S;
    const int upc = big ? kp.upc_big : p.units_per_chunk;
    const int local = big ? unit : unit - kp.n_big;
    const int lrow = local / S;
    const int chunk = local - lrow * S;
    const int row = (big ? 0 : kp.lag * p.rows) + lrow;
    const int b = row / p.rows;
    const int it0 = chunk * upc;
    int it1 = it0 + upc;
    if (it1 > p.upr) it1 = p.upr;
    const RowGeom<T> g = row_geom<T>(p, row);
    const int v0 = it0 * nthreads;
    int nv = g.nvec - v0;
    const int cap = (it1 - it0) * nthreads;
    nv = nv < 0 ? 0 : (nv > cap ? cap : nv);
    // Fetch of the chunk: kp.fetch 0 = ONE bulk copy (TMA), 1 = cp.async by every thread (LSU path), 2 = first half bulk,
    // second half cp.async.  A bulk copy progresses at ~6.3 GB/s whatever its size (a small window of outstanding requests
    // per operation: 36 KB in ~6 us, 52 KB in ~8 us), so four of them per SM stream at ~3.5 TB/s; cp.async requests have
    // no such window.
    const int nv_bulk = kp.fetch == 0 ? nv : (kp.fetch == 1 ? 0 : ((nv / 2) & ~7));
    if (tid == 0) {
        if (kp.stagger_ns > 0 && unit < 4 * 148) __nanosleep((unsigned)((unit / 148) * kp.stagger_ns));  // A/B: de-synchronise the first wave
        if (nv_bulk > 0) {
            vbar_arrive_expect_tx(bar, (unsigned)nv_bulk * 16u);
            const unsigned total = (unsigned)nv_bulk * 16u, sub = kp.bulk_split > 0 ? (unsigned)kp.bulk_split : total;
            for (unsigned off = 0; off < total; off += sub)
                vbulk_g2s(stage + off, reinterpret_cast<const unsigned char*>(g.vp + v0) + off, total - off < sub ? total - off : sub, bar);
        } else {
            vbar_arrive(bar);
        }
    }
    if (nv_bulk < nv) {
        for (int k = nv_bulk + tid; k < nv; k += nthreads) cp_async16(stage + (size_t)k * 16, g.vp + v0 + k);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    // A/B (kp.pf_ahead): pull the logits chunk of the verify CTA `pf_ahead` places behind this one into L2 with ONE
    // cp.async.bulk.prefetch.L2 -- bytes in flight towards DRAM that need no shared memory; the CTA that owns the chunk
    // then copies it from L2.  (The resident CTAs hold 592 x 51 KB = 30 MB in shared memory; at the ~6 us a chunk takes
    // under load that caps the logits phase at ~5 TB/s.)
    if (kp.pf_ahead > 0 && tid == 0) {
        const int unit2 = unit + kp.pf_ahead;
        if (unit2 < kp.n_verify) {
            const bool big2 = unit2 < kp.n_big;
            const int S2 = big2 ? kp.S_big : p.S;
            const int upc2 = big2 ? kp.upc_big : p.units_per_chunk;
            const int local2 = big2 ? unit2 : unit2 - kp.n_big;
            const int lrow2 = local2 / S2;
            const int chunk2 = local2 - lrow2 * S2;
            const int row2 = (big2 ? 0 : kp.lag * p.rows) + lrow2;
            const RowGeom<T> g2 = row_geom<T>(p, row2);
            const int v02 = chunk2 * upc2 * nthreads;
            int nv2 = g2.nvec - v02;
            nv2 = nv2 < 0 ? 0 : (nv2 > upc2 * nthreads ? upc2 * nthreads : nv2);
            if (nv2 > 0)
                asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(g2.vp + v02), "r"((unsigned)nv2 * 16u) : "memory");
        }
    }
    // The launch epoch is CONSUMED here (the branch cannot be resolved before the load returns), i.e. before this
    // CTA publishes anything: the closing CTA advances the epoch only after every partial has been published.
    unsigned epoch_raw;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(epoch_raw) : "l"(kp.epoch_word) : "memory");
    if (epoch_raw == 0xffffffffu) return;  // (never a valid epoch: the closing CTA skips it)
    const unsigned tag = (epoch_raw + 1u) & 0xffffffu;
    const int lag = kp.lag;
    const int piece = big ? 0 : local - (b - lag) * kp.pieces;
    const bool carries = !big;
    PieceRegs pr;
    if (carries) kv_piece_begin(kp, b - lag, piece, tag, pr, park);
    if (b >= p.B - lag && kp.l2_prefetch)  // rows of a tail sequence: into L2, now
        kv_piece_prefetch(kp, b, local - (big ? b : b - lag) * p.rows * S, p.rows * S);
    ThreadAcc<T, false> acc;
    acc.reset();
    // ONE warp polls the mbarrier; the others block in the hardware barrier behind it.  (256 threads spinning on
    // try_wait for the ~3.5 us a chunk takes under load cost issue slots that the epilogues of the other resident CTAs
    // need: ncu showed 55 % of the issue slots busy in a kernel that mostly waits.)
    if (kp.one_warp_wait) {
        if (tid < 32) vbar_wait(bar, 0u);
        __syncthreads();
    } else {
        vbar_wait(bar, 0u);
    }
    if (nv_bulk < nv) {  // the cp.async part of the chunk (the carried rows, committed after it, may still be in flight)
        if (carries) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
    }
    const uint4* src = reinterpret_cast<const uint4*>(stage);
#pragma unroll 4
    for (int k = tid; k < nv; k += nthreads) acc.add(src[k], v0 + k, true, p.scale_log2);
    if (p.trace && tid == 0) p.trace[(size_t)blockIdx.x * 4 + 1] = global_timer_ns();
    if (publish_segment<T, false, false>(p, g, acc, row, it0, it1, chunk, tid, nthreads))
        finalize_sequence<T, false, false>(p, b, S, tid, nthreads, tag);
    if (carries) kv_piece_finish(kp, tag, pr, park);
}

// dynamic shared memory: [stage_bytes: the logits chunk][park: kCarryNI * 2 * 16 * blockDim.x bytes of parked KV rows];
// a tail CTA parks rows in all of it (tail_park_items per thread, one pass)
template <typename T>
__global__ void __launch_bounds__(256, 4) verify_step_kernel(const VerifyParams vp, const KvParams kp, const int n_verify,
                                                          const int stage_bytes, const int tail_park_items) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        vbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    if (vp.trace && threadIdx.x == 0) {  // evidence runs (tools/trace_step.py): {entry, scan done / rows stored, exit, SM id}
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        const unsigned kind = (int)blockIdx.x < kp.n_big ? 0u : ((int)blockIdx.x < n_verify ? 1u : 2u);  // big verify / carrying verify / tail
        vp.trace[(size_t)blockIdx.x * 4 + 0] = global_timer_ns();
        vp.trace[(size_t)blockIdx.x * 4 + 3] = smid | (kind << 16);
    }
    if ((int)blockIdx.x < n_verify)
        step_cta_body<T>(vp, kp, (int)blockIdx.x, stage, &bar, reinterpret_cast<uint4*>(stage + stage_bytes));
    else
        kv_tail_body(kp, (int)blockIdx.x - n_verify, vp.trace ? vp.trace + (size_t)blockIdx.x * 4 : nullptr,
                     reinterpret_cast<uint4*>(stage), tail_park_items);
    if (vp.trace && threadIdx.x == 0) vp.trace[(size_t)blockIdx.x * 4 + 2] = global_timer_ns();
}

// Round-1 layout, kept for A/B evidence (sd_verify_append_set_mode(1, 0)): CTAs [0, n_verify) run the bulk-chunk
// verify body, CTAs after them the kv_append column body for (x, layer), each thread waiting for its sequence's
// completion flag.  The append CTAs become resident only when every verify CTA has started, so the two phases
// serialise (measured: 34.6 us = 18.9 + 14.9).
template <typename T, int RB>
__global__ void __launch_bounds__(256, 4) verify_append_kernel(const VerifyParams vp, const KvParams kp, const int n_verify,
                                                            const int kv_gx) {
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    if (threadIdx.x == 0) {
        vbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    pdl_wait();
    pdl_trigger();
    if ((int)blockIdx.x < n_verify) {
        verify_bulk_body<T, false>(vp, (int)blockIdx.x, stage, &bar);
    } else {
        const int cta = (int)blockIdx.x - n_verify;
        kv_col_body<RB>(kp, cta % kv_gx, cta / kv_gx);
    }
}

// ------------------------------------------------------------------------------------------ host
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WorkspaceLayout {
    size_t total;
};

static WorkspaceLayout layout(int B, int rows) {
    WorkspaceLayout w;
    w.total = align_up((size_t)B * rows * kMaxChunks * 16, 256);
    return w;
}

static int g_sm_count = 0;
static int sm_count() {
    if (g_sm_count == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
        g_sm_count = n;
    }
    return g_sm_count;
}

template <typename T, bool STATS>
static cudaError_t launch_t(const VerifyParams& p, int variant, unsigned grid, int threads, int vpt, cudaStream_t st) {
    if (variant == 1) {
        switch (vpt) {
            case 1: return launch_kernel(verify_persist_kernel<T, 1, STATS>, dim3(grid), dim3(threads), 0, st, p);
            case 2: return launch_kernel(verify_persist_kernel<T, 2, STATS>, dim3(grid), dim3(threads), 0, st, p);
            case 4: return launch_kernel(verify_persist_kernel<T, 4, STATS>, dim3(grid), dim3(threads), 0, st, p);
            case 8: return launch_kernel(verify_persist_kernel<T, 8, STATS>, dim3(grid), dim3(threads), 0, st, p);
            default: return launch_kernel(verify_persist_kernel<T, 16, STATS>, dim3(grid), dim3(threads), 0, st, p);
        }
    }
    switch (vpt) {
        case 1: return launch_kernel(verify_kernel<T, 1, STATS>, dim3(grid), dim3(threads), 0, st, p);
        case 2: return launch_kernel(verify_kernel<T, 2, STATS>, dim3(grid), dim3(threads), 0, st, p);
        case 4: return launch_kernel(verify_kernel<T, 4, STATS>, dim3(grid), dim3(threads), 0, st, p);
        case 5: return launch_kernel(verify_kernel<T, 5, STATS>, dim3(grid), dim3(threads), 0, st, p);
        case 6: return launch_kernel(verify_kernel<T, 6, STATS>, dim3(grid), dim3(threads), 0, st, p);
        case 7: return launch_kernel(verify_kernel<T, 7, STATS>, dim3(grid), dim3(threads), 0, st, p);
        case 8: return launch_kernel(verify_kernel<T, 8, STATS>, dim3(grid), dim3(threads), 0, st, p);
        default: return launch_kernel(verify_kernel<T, 16, STATS>, dim3(grid), dim3(threads), 0, st, p);
    }
}

template <typename T, bool STATS>
static cudaError_t launch_tma_t(const VerifyParams& p, unsigned grid, int vpu, int stages, size_t smem, cudaStream_t st) {
    cudaError_t e;
    if (vpu == 2) {
        e = cudaFuncSetAttribute(verify_tma_kernel<T, 2, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        return launch_kernel(verify_tma_kernel<T, 2, STATS>, dim3(grid), dim3(kTmaConsumers + 32), smem, st, p, stages);
    }
    e = cudaFuncSetAttribute(verify_tma_kernel<T, 4, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel(verify_tma_kernel<T, 4, STATS>, dim3(grid), dim3(kTmaConsumers + 32), smem, st, p, stages);
}

template <typename T, bool STATS>
static cudaError_t launch_bulk_t(const VerifyParams& p, unsigned grid, int threads, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(verify_bulk_kernel<T, STATS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel(verify_bulk_kernel<T, STATS>, dim3(grid), dim3(threads), smem, st, p);
}

static int g_fused_mode = 0;   // 0 = lagged append (default), 1 = round-1 layout (append CTAs behind all verify CTAs)
static int g_fused_lag = 0;    // sequences of lag for mode 0 (0 = heuristic)

template <typename T, int RB>
static cudaError_t launch_fused_t(const VerifyParams& p, const KvParams& kp, unsigned total, int n_verify, int kv_gx,
                                  int threads, size_t smem, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(verify_append_kernel<T, RB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel(verify_append_kernel<T, RB>, dim3(total), dim3(threads), smem, st, p, kp, n_verify, kv_gx);
}

template <typename T>
static cudaError_t launch_fused_rb(const VerifyParams& p, const KvParams& kp, int rb, unsigned total, int n_verify, int kv_gx,
                                   int threads, size_t smem, cudaStream_t st) {
    switch (rb) {
        case 1: return launch_fused_t<T, 1>(p, kp, total, n_verify, kv_gx, threads, smem, st);
        case 2: return launch_fused_t<T, 2>(p, kp, total, n_verify, kv_gx, threads, smem, st);
        default: return launch_fused_t<T, 3>(p, kp, total, n_verify, kv_gx, threads, smem, st);
    }
}

static cudaError_t launch_fused(const VerifyParams& p, const KvParams& kp, int logit_dtype, int rb, unsigned total,
                                int n_verify, int kv_gx, int threads, size_t smem, cudaStream_t st) {
    if (logit_dtype == SD_F32) return launch_fused_rb<float>(p, kp, rb, total, n_verify, kv_gx, threads, smem, st);
    if (logit_dtype == SD_F16) return launch_fused_rb<__half>(p, kp, rb, total, n_verify, kv_gx, threads, smem, st);
    return launch_fused_rb<__nv_bfloat16>(p, kp, rb, total, n_verify, kv_gx, threads, smem, st);
}

template <typename T>
static cudaError_t launch_step_t(const VerifyParams& p, const KvParams& kp, unsigned total, int n_verify, int threads,
                                 size_t stage_bytes, size_t smem, int tail_park_items, cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(verify_step_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return launch_kernel(verify_step_kernel<T>, dim3(total), dim3(threads), smem, st, p, kp, n_verify, (int)stage_bytes,
                         tail_park_items);
}

static cudaError_t launch_step(const VerifyParams& p, const KvParams& kp, int logit_dtype, unsigned total, int n_verify,
                               int threads, size_t stage_bytes, size_t smem, int tail_park_items, cudaStream_t st) {
    if (logit_dtype == SD_F32) return launch_step_t<float>(p, kp, total, n_verify, threads, stage_bytes, smem, tail_park_items, st);
    if (logit_dtype == SD_F16) return launch_step_t<__half>(p, kp, total, n_verify, threads, stage_bytes, smem, tail_park_items, st);
    return launch_step_t<__nv_bfloat16>(p, kp, total, n_verify, threads, stage_bytes, smem, tail_park_items, st);
}

// resident CTAs per SM of the fused step kernel (registers: 4; shared memory: chunk stage + parked rows)
static int step_occupancy(int logit_dtype, int threads, size_t smem) {
    static size_t cached_smem[3] = {0, 0, 0};
    static int cached_n[3] = {0, 0, 0};
    if (cached_n[logit_dtype] > 0 && cached_smem[logit_dtype] == smem) return cached_n[logit_dtype];
    int n = 0;
    cudaError_t e;
    if (logit_dtype == SD_F32) {
        cudaFuncSetAttribute(verify_step_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_step_kernel<float>, threads, smem);
    } else if (logit_dtype == SD_F16) {
        cudaFuncSetAttribute(verify_step_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_step_kernel<__half>, threads, smem);
    } else {
        cudaFuncSetAttribute(verify_step_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_step_kernel<__nv_bfloat16>, threads, smem);
    }
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        n = 4;
    }
    cached_smem[logit_dtype] = smem;
    cached_n[logit_dtype] = n;
    return n;
}

// resident CTAs per SM of the persistent kernel for (T, vpt, stats, threads); cached per combination
template <typename T, bool STATS>
static int persist_occupancy_t(int vpt, int threads) {
    int n = 0;
    cudaError_t e;
    switch (vpt) {
        case 1: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_persist_kernel<T, 1, STATS>, threads, 0); break;
        case 2: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_persist_kernel<T, 2, STATS>, threads, 0); break;
        case 4: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_persist_kernel<T, 4, STATS>, threads, 0); break;
        case 8: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_persist_kernel<T, 8, STATS>, threads, 0); break;
        default: e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, verify_persist_kernel<T, 16, STATS>, threads, 0); break;
    }
    if (e != cudaSuccess || n <= 0) {
        (void)cudaGetLastError();
        n = 1;
    }
    return n;
}

static int persist_occupancy(int logit_dtype, bool stats, int vpt, int threads) {
    static int cache[3][2][5][33] = {};  // dtype, stats, log2(vpt), threads/32
    int lv = 0;
    while ((1 << lv) < vpt) ++lv;
    int& slot = cache[logit_dtype][stats ? 1 : 0][lv][threads / 32];
    if (slot == 0) {
        if (logit_dtype == SD_F32) slot = stats ? persist_occupancy_t<float, true>(vpt, threads) : persist_occupancy_t<float, false>(vpt, threads);
        else if (logit_dtype == SD_F16) slot = stats ? persist_occupancy_t<__half, true>(vpt, threads) : persist_occupancy_t<__half, false>(vpt, threads);
        else slot = stats ? persist_occupancy_t<__nv_bfloat16, true>(vpt, threads) : persist_occupancy_t<__nv_bfloat16, false>(vpt, threads);
    }
    return slot;
}

static int run_verify(bool stats, const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                      const void* draft_ids, int id_dtype, int64_t id_stride_b, int mode, float threshold,
                      float inv_temperature, int32_t* accept_len, uint8_t* accepted_mask,
                      int32_t* target_ids, int32_t* bonus, int32_t* packed, float* token_prob,
                      float* row_max, float* row_sumexp, void* workspace, size_t workspace_bytes, int B,
                      int rows, int K, int V, void* stream, const sd_exchange* exchange, int32_t* seq_ready,
                      const KvParams* fused_kv = nullptr, unsigned fused_gx = 0, int fused_rb = 0,
                      int32_t* step_state = nullptr) {
    if (B < 0 || K < 0 || V <= 0 || rows <= 0) return SD_ERR_INVALID_ARG;
    if (rows != K && rows != K + 1) return SD_ERR_SHAPE;
    if (rows > 64) return SD_ERR_UNSUPPORTED;
    if (B == 0) return SD_OK;
    if (!logits || !accept_len || (K > 0 && !draft_ids)) return SD_ERR_INVALID_ARG;
    if (logit_dtype < SD_F32 || logit_dtype > SD_BF16 || (id_dtype != SD_I32 && id_dtype != SD_I64))
        return SD_ERR_INVALID_ARG;
    if (mode < SD_ACCEPT_GREEDY || mode > SD_ACCEPT_TOPK) return SD_ERR_INVALID_ARG;
    const int es = elt_size(logit_dtype);
    if (reinterpret_cast<uintptr_t>(logits) % es || reinterpret_cast<uintptr_t>(workspace) % 8) return SD_ERR_ALIGNMENT;
    WorkspaceLayout w = layout(B, rows);
    if (!workspace || workspace_bytes < w.total) return SD_ERR_WORKSPACE;

    const int epv = 16 / es;
    const int nvec_max = V / epv + 1;
    const long long total_vec = (long long)B * rows * nvec_max;
    int variant = g_variant;
    int threads = g_tuning.threads, vpt = g_tuning.vpt;
    if (threads <= 0 || vpt <= 0) {
        threads = 256;
        vpt = stats ? 4 : 8;
        if (variant == 0) {
            // ~2 waves of 256-thread CTAs with 8 vectors (128 B) in flight per thread
            const long long slots = (long long)sm_count() * 8;
            while (vpt > 1 && total_vec / ((long long)threads * vpt) < slots) vpt >>= 1;
        }
    }
    if (variant == 1 && vpt != 1 && vpt != 2 && vpt != 4 && vpt != 8) vpt = vpt > 8 ? 16 : (vpt > 4 ? 8 : 4);
    if (variant == 0 && vpt > 8 && vpt != 16) vpt = 16;
    if (variant == 0 && vpt == 3) vpt = 4;
    if (variant == 3) {  // bulk chunks: 256 threads, vpt = units (4 KB each) per chunk, <= 12 (48 KB)
        threads = 256;
        if (g_tuning.vpt > 0) {
            vpt = g_tuning.vpt > 12 ? 12 : g_tuning.vpt;
        } else {
            // 36 KB chunks (measured best at B=64, V=128256); smaller ones while the grid would leave SMs idle
            vpt = 9;
            const int upr3 = (nvec_max + threads - 1) / threads;
            while (vpt > 2 && (long long)B * rows * ((upr3 + vpt - 1) / vpt) < 2LL * sm_count()) --vpt;
        }
    }
    if (variant == 2) {  // TMA ring: 256 consumers, unit = 256 * vpu vectors
        threads = kTmaConsumers;
        vpt = (g_tuning.vpt == 4) ? 4 : 2;
    }
    const int unit_vecs = variant == 2 ? threads * vpt : threads;
    const int upr = (nvec_max + unit_vecs - 1) / unit_vecs;  // units per row
    const long long U = (long long)B * rows * upr;
    if (U > 0x3fffffffLL) return SD_ERR_UNSUPPORTED;

    VerifyParams p;
    p.logits = logits;
    p.stride_b = stride_b;
    p.stride_r = stride_r;
    p.ids = draft_ids;
    p.id_dtype = id_dtype;
    p.id_stride_b = id_stride_b;
    p.accept_len = accept_len;
    p.mask = accepted_mask;
    p.target_ids = target_ids;
    p.bonus = bonus;
    p.packed = packed;
    p.token_prob = token_prob;
    p.row_max = row_max;
    p.row_sumexp = row_sumexp;
    char* ws = reinterpret_cast<char*>(workspace);
    p.part = reinterpret_cast<unsigned long long*>(ws);
    p.B = B;
    p.rows = rows;
    p.K = K;
    p.V = V;
    p.upr = upr;
    p.S = kMaxChunks;
    p.units_per_chunk = 0;
    p.U = (int)U;
    p.q = p.rem = 0;
    p.mode = mode;
    p.threshold = threshold;
    p.inv_temperature = inv_temperature;
    p.scale_log2 = inv_temperature * 1.4426950408889634f;
    p.trace = nullptr;
    p.xchg = exchange;
    p.seq_ready = seq_ready;
    p.words = nullptr;
    p.epoch_word = nullptr;
    p.seq_lens = nullptr;
    p.kv_overflow = nullptr;
    p.kv_capacity = 0;
    p.kv_R = 0;
    p.word_release = 0;

    unsigned grid;
    int tma_stages = 0;
    size_t tma_smem = 0;
    if (variant == 1 || variant == 2) {
        int per_sm;
        int q_floor;
        if (variant == 1) {
            per_sm = persist_occupancy(logit_dtype, stats, vpt, threads);
            if (g_ctas_per_sm > 0 && g_ctas_per_sm < per_sm) per_sm = g_ctas_per_sm;
            q_floor = vpt;  // a CTA should own at least one full batch of loads
        } else {
            per_sm = g_ctas_per_sm > 0 ? (g_ctas_per_sm > 4 ? 4 : g_ctas_per_sm) : 2;
            const size_t unit_bytes = (size_t)unit_vecs * 16;
            const size_t budget = (size_t)(224 * 1024) / per_sm - 2048;  // static smem + reserve per CTA
            tma_stages = (int)(budget / unit_bytes);
            if (tma_stages > kTmaMaxStages) tma_stages = kTmaMaxStages;
            if (tma_stages < 2) return SD_ERR_UNSUPPORTED;
            tma_smem = (size_t)tma_stages * unit_bytes;
            q_floor = 2;
        }
        long long G = (long long)sm_count() * per_sm;
        // a row may be shared by at most kMaxChunks - 2 CTAs (partial slots): q >= upr / (kMaxChunks - 2)
        const int q_min = (upr + (kMaxChunks - 3)) / (kMaxChunks - 2) > q_floor ? (upr + (kMaxChunks - 3)) / (kMaxChunks - 2) : q_floor;
        if (G * q_min > U) G = U / q_min;
        if (G < 1) G = 1;
        p.q = (int)(U / G);
        p.rem = (int)(U % G);
        grid = (unsigned)G;
    } else {
        int upc = vpt;  // units per chunk
        int S = (upr + upc - 1) / upc;
        if (S > kMaxChunks) {  // very long rows: loop inside the CTA instead of adding partials
            S = kMaxChunks;
            upc = (upr + S - 1) / S;
            S = (upr + upc - 1) / upc;
            if (variant == 3 && upc > 12) variant = 0;  // chunk would not fit the shared-memory stage
        }
        p.S = S;
        p.units_per_chunk = upc;
        grid = (unsigned)((size_t)B * rows * S);
    }

    if (variant == 0 && g_trace && (size_t)grid <= g_trace_ctas) p.trace = g_trace;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (fused_kv) {
        if (variant != 3 || stats) return SD_ERR_UNSUPPORTED;
        const size_t smem = (size_t)p.units_per_chunk * threads * 16;
        if (step_state) {
            // lagged append (default): step_state = u64 words[B] | u32 epoch | u32 reserved
            KvParams kp = *fused_kv;
            kp.flags = reinterpret_cast<unsigned long long*>(step_state);
            kp.epoch_word = reinterpret_cast<unsigned*>(step_state) + 2 * (size_t)B;
            kp.seq_ready = nullptr;
            kp.pieces = rows * p.S;
            const long long seq_items = (long long)kp.n_layers * kp.H * kp.R * (kp.row_bytes / 16);
            if (seq_items > 0x3fffffffLL || kp.R > 255 || kp.n_layers > (1 << 22)) return SD_ERR_UNSUPPORTED;
            kp.seq_items = (int)seq_items;
            kp.piece_items = (int)((seq_items + kp.pieces - 1) / kp.pieces);
            kp.fd_vpr = make_fastdiv((uint32_t)(kp.row_bytes / 16));
            kp.fd_R = make_fastdiv((uint32_t)kp.R);
            kp.fd_H = make_fastdiv((uint32_t)kp.H);
            // shared memory: the chunk stage (16-byte granular) + the parked rows of kCarryNI items per thread
            const size_t stage_bytes = (smem + 15) / 16 * 16;
            static const int smem_pad = [] {  // A/B: fewer resident CTAs per SM (smaller wave -> smaller lag -> smaller tail)
                const char* e = getenv("SD_B200_STEP_SMEM_PAD");
                return e ? atoi(e) : 0;
            }();
            static const int carry_env = [] {
                const char* e = getenv("SD_B200_STEP_CARRY_NI");
                return e ? atoi(e) : 0;
            }();
            int carry_ni = (kp.piece_items + threads - 1) / threads;
            carry_ni = carry_ni < 1 ? 1 : (carry_ni > kCarryNI ? kCarryNI : carry_ni);
            if (carry_env >= 1 && carry_env <= kCarryNI) carry_ni = carry_env;
            kp.carry_ni = carry_ni;
            const size_t smem_step = stage_bytes + (size_t)carry_ni * 2 * 16 * threads + (size_t)smem_pad;
            int tail_park_items = (int)(smem_step / ((size_t)2 * 16 * threads));
            static const int tail_items_env = [] {
                const char* e = getenv("SD_B200_STEP_TAIL_ITEMS");
                return e ? atoi(e) : 0;
            }();
            if (tail_items_env > 0 && tail_items_env < tail_park_items) tail_park_items = tail_items_env;
            const int tail_park_max = tail_park_items;
            kp.tail_pieces = (int)((seq_items + (long long)tail_park_items * threads - 1) / ((long long)tail_park_items * threads));
            if (kp.tail_pieces < 1) kp.tail_pieces = 1;
            // Lag: far enough that a carrying CTA finds the word of its sequence published when it starts (one wave of
            // resident CTAs plus the finalizer's latency: >= 1.5 waves' worth of sequences).  Measured optimum with the
            // big chunks of the non-carrying sequences: ~2.4 waves (B = 64: lag 34..46 -> 30.6-30.8 us, 26 -> 31.4, 50 ->
            // 31.3; B = 128: lag 40 -> 51.4 us, 26 -> 53.7, 80 -> 53.0).
            int lag = g_fused_lag;
            if (lag <= 0) {
                static const int lag_pct = [] {  // A/B: lag as a percentage of one wave of resident CTAs
                    const char* e = getenv("SD_B200_STEP_LAG_PCT");
                    return e ? atoi(e) : 240;
                }();
                const long long resident = (long long)step_occupancy(logit_dtype, threads, smem_step) * sm_count();
                lag = (int)((resident * lag_pct / 100 + kp.pieces - 1) / kp.pieces);
                if (lag < 1) lag = 1;
            }
            if (lag > B) lag = B;
            kp.lag = lag;
            // Tail pieces: as many per sequence as keep ALL tail CTAs resident at once (one wave), i.e. the smallest
            // tail CTAs that still start together -- B = 64: 14 pieces of 5 items per thread (574 CTAs on 592 slots)
            // 27.85 us per step; 12 pieces of 6 items (what the shared memory holds) 28.2 us; 18 of 4: 28.25 us.
            if (tail_items_env <= 0) {
                const long long resident = (long long)step_occupancy(logit_dtype, threads, smem_step) * sm_count();
                long long tp = resident / lag;
                if (tp > kp.tail_pieces) {
                    int items = (int)((seq_items + tp * threads - 1) / (tp * threads));
                    if (items < 1) items = 1;
                    if (items < tail_park_max) {
                        tail_park_items = items;
                        kp.tail_pieces = (int)((seq_items + (long long)items * threads - 1) / ((long long)items * threads));
                    }
                }
            }
            // the first `lag` sequences: bigger chunks (stage + park), fewer CTAs
            static const int big_env = [] {
                const char* e = getenv("SD_B200_STEP_BIG_CHUNKS");
                return (e && e[0] == '0') ? 0 : 1;
            }();
            kp.upc_big = big_env ? (int)((stage_bytes + (size_t)carry_ni * 2 * 16 * threads) / ((size_t)threads * 16)) : p.units_per_chunk;
            if (kp.upc_big < p.units_per_chunk) kp.upc_big = p.units_per_chunk;
            kp.S_big = (upr + kp.upc_big - 1) / kp.upc_big;
            kp.n_big = lag * rows * kp.S_big;
            grid = (unsigned)(kp.n_big + (B - lag) * rows * p.S);
            const unsigned n_tail = (unsigned)lag * (unsigned)kp.tail_pieces + (kp.xchg ? 1u : 0u);  // (+ the exchange CTA)
            kp.total_ctas = n_tail;
            static const int l2pf = [] {  // A/B (default off: measured no gain, the tail is bandwidth-bound, not latency-bound)
                const char* e = getenv("SD_B200_STEP_L2_PREFETCH");
                return (e && e[0] == '1') ? 1 : 0;
            }();
            kp.l2_prefetch = l2pf;
            static const int stagger = [] {
                const char* e = getenv("SD_B200_STEP_STAGGER_NS");
                return e ? atoi(e) : 0;
            }();
            kp.stagger_ns = stagger;
            static const int oww = [] {
                const char* e = getenv("SD_B200_STEP_ONE_WARP_WAIT");
                return e ? atoi(e) : 0;
            }();
            kp.one_warp_wait = oww;
            static const int bsplit = [] {
                const char* e = getenv("SD_B200_STEP_BULK_SPLIT");
                return e ? atoi(e) : 0;
            }();
            kp.bulk_split = bsplit > 0 ? (bsplit + 15) / 16 * 16 : 0;
            static const int fetch = [] {
                const char* e = getenv("SD_B200_STEP_FETCH");
                return e ? atoi(e) : 0;
            }();
            kp.fetch = fetch;
            p.words = kp.flags;
            p.epoch_word = kp.epoch_word;
            p.seq_lens = kp.seq_lens_out;
            p.kv_overflow = kp.overflow;
            p.kv_capacity = kp.capacity;
            p.kv_R = kp.R;
            p.word_release = kp.xchg_packed ? 1 : 0;
            p.seq_ready = nullptr;
            static const int eager = [] {
                const char* e = getenv("SD_B200_STEP_EAGER_ROWS");
                return (e && e[0] == '1') ? 1 : 0;
            }();
            kp.eager_rows = eager;
            static const int tail_order = [] {
                const char* e = getenv("SD_B200_STEP_TAIL_ORDER");
                return e ? atoi(e) : 1;  // measured: 29.1 us per step against 30.0 us with the last sequence first
            }();
            kp.tail_order = tail_order;
            static const int tail_peek = [] {
                const char* e = getenv("SD_B200_STEP_TAIL_PEEK");
                return e ? atoi(e) : 0;
            }();
            kp.tail_peek = tail_peek;
            static const int pf_ahead = [] {
                const char* e = getenv("SD_B200_STEP_PF_AHEAD");
                return e ? atoi(e) : 0;
            }();
            kp.pf_ahead = pf_ahead;
            kp.n_verify = (int)grid;
            if (g_trace && (size_t)(grid + n_tail) <= g_trace_ctas) p.trace = g_trace;
            return to_status(launch_step(p, kp, logit_dtype, grid + n_tail, (int)grid, threads, stage_bytes, smem_step, tail_park_items, st));
        }
        const unsigned total = grid + fused_gx * (unsigned)fused_kv->n_layers;
        // RB <= 3 rows of K and V in flight per append thread keeps the kernel at <= 64 registers (4 CTAs per SM)
        const int rb = fused_rb > 3 ? 3 : fused_rb;
        return to_status(launch_fused(p, *fused_kv, logit_dtype, rb, total, (int)grid, (int)fused_gx, threads, smem, st));
    }
    if (variant == 3) {
        const size_t smem = (size_t)p.units_per_chunk * threads * 16;
        if (stats) {
            if (logit_dtype == SD_F32) e = launch_bulk_t<float, true>(p, grid, threads, smem, st);
            else if (logit_dtype == SD_F16) e = launch_bulk_t<__half, true>(p, grid, threads, smem, st);
            else e = launch_bulk_t<__nv_bfloat16, true>(p, grid, threads, smem, st);
        } else {
            if (logit_dtype == SD_F32) e = launch_bulk_t<float, false>(p, grid, threads, smem, st);
            else if (logit_dtype == SD_F16) e = launch_bulk_t<__half, false>(p, grid, threads, smem, st);
            else e = launch_bulk_t<__nv_bfloat16, false>(p, grid, threads, smem, st);
        }
        return to_status(e);
    }
    if (variant == 2) {
        if (stats) {
            if (logit_dtype == SD_F32) e = launch_tma_t<float, true>(p, grid, vpt, tma_stages, tma_smem, st);
            else if (logit_dtype == SD_F16) e = launch_tma_t<__half, true>(p, grid, vpt, tma_stages, tma_smem, st);
            else e = launch_tma_t<__nv_bfloat16, true>(p, grid, vpt, tma_stages, tma_smem, st);
        } else {
            if (logit_dtype == SD_F32) e = launch_tma_t<float, false>(p, grid, vpt, tma_stages, tma_smem, st);
            else if (logit_dtype == SD_F16) e = launch_tma_t<__half, false>(p, grid, vpt, tma_stages, tma_smem, st);
            else e = launch_tma_t<__nv_bfloat16, false>(p, grid, vpt, tma_stages, tma_smem, st);
        }
        return to_status(e);
    }
    if (stats) {
        if (logit_dtype == SD_F32) e = launch_t<float, true>(p, variant, grid, threads, vpt, st);
        else if (logit_dtype == SD_F16) e = launch_t<__half, true>(p, variant, grid, threads, vpt, st);
        else e = launch_t<__nv_bfloat16, true>(p, variant, grid, threads, vpt, st);
    } else {
        if (logit_dtype == SD_F32) e = launch_t<float, false>(p, variant, grid, threads, vpt, st);
        else if (logit_dtype == SD_F16) e = launch_t<__half, false>(p, variant, grid, threads, vpt, st);
        else e = launch_t<__nv_bfloat16, false>(p, variant, grid, threads, vpt, st);
    }
    return to_status(e);
}

int verify_tu_spin_error(int clear, unsigned* out) { return tu_spin_error(clear, out); }

}  // namespace sd

extern "C" {

size_t sd_verify_workspace_bytes(int B, int rows) {
    if (B <= 0 || rows <= 0) return 256;
    return sd::layout(B, rows).total;
}

int sd_verify_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace) return SD_ERR_INVALID_ARG;
    return sd::to_status(cudaMemsetAsync(workspace, 0, workspace_bytes, reinterpret_cast<cudaStream_t>(stream)));
}

int sd_verify_set_tuning(int threads, int vectors_per_thread) {
    if (threads == 0 || vectors_per_thread == 0) {
        sd::g_tuning = {0, 0};
        return SD_OK;
    }
    if (threads < 32 || threads > 1024 || threads % 32) return SD_ERR_INVALID_ARG;
    int v = vectors_per_thread;
    if (v < 1 || v > 16) return SD_ERR_INVALID_ARG;  // each variant rounds to what it was built for
    if (threads > 512) return SD_ERR_INVALID_ARG;  // register budget (stats / 8+-vector kernels are built for <= 512)
    sd::g_tuning = {threads, v};
    return SD_OK;
}

int sd_verify_set_trace(void* device_buffer, size_t capacity_ctas) {
    sd::g_trace = reinterpret_cast<unsigned long long*>(device_buffer);
    sd::g_trace_ctas = device_buffer ? capacity_ctas : 0;
    return SD_OK;
}

int sd_verify_set_variant(int variant, int ctas_per_sm) {
    if (variant < 0 || variant > 3 || ctas_per_sm < 0 || ctas_per_sm > 32) return SD_ERR_INVALID_ARG;
    sd::g_variant = variant;
    sd::g_ctas_per_sm = ctas_per_sm;
    return SD_OK;
}

int sd_verify_prefix(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                     const void* draft_ids, int id_dtype, int64_t id_stride_b, int32_t* accept_len,
                     uint8_t* accepted_mask, int32_t* target_ids, int32_t* bonus, int32_t* packed,
                     void* workspace, size_t workspace_bytes, int B, int rows, int K, int V, void* stream,
                     const sd_exchange* exchange, int32_t* seq_ready) {
    return sd::run_verify(false, logits, logit_dtype, stride_b, stride_r, draft_ids, id_dtype, id_stride_b,
                          SD_ACCEPT_GREEDY, 0.f, 1.f, accept_len, accepted_mask, target_ids, bonus, packed,
                          nullptr, nullptr, nullptr, workspace, workspace_bytes, B, rows, K, V, stream, exchange, seq_ready);
}

int sd_verify_append_set_mode(int mode, int lag) {
    if (mode < 0 || mode > 1 || lag < 0 || lag > 4096) return SD_ERR_INVALID_ARG;
    sd::g_fused_mode = mode;
    sd::g_fused_lag = lag;
    return SD_OK;
}

size_t sd_step_state_bytes(int B) { return B > 0 ? (size_t)B * 8 + 8 : 8; }

int sd_verify_append(const sd_step_args* a, void* stream) {
    using namespace sd;
    if (!a) return SD_ERR_INVALID_ARG;
    if (a->rows != a->K + 1) return SD_ERR_SHAPE;  // the append keeps accept_len + 1 of the K + 1 offered rows
    if (!a->accept_len || !a->ticket) return SD_ERR_INVALID_ARG;
    const int mode = g_fused_mode;
    if (mode == 0 && (!a->step_state || reinterpret_cast<uintptr_t>(a->step_state) % 8)) return SD_ERR_INVALID_ARG;
    if (mode == 1 && !a->seq_ready) return SD_ERR_INVALID_ARG;
    // With an exchange the records are pushed either by the append part (exchange_push_in_append) or by the verify
    // finalizers; in both cases one append CTA gathers them, after every finalizer of this launch has read
    // the epoch (sd_kv_col.cuh -- the missing wait was the multi-GPU stall of round 1, tools/repro_epoch_race.py).
    if (a->exchange && !a->packed) return SD_ERR_INVALID_ARG;
    if (a->B <= 0) return SD_OK;
    KvParams kp;
    unsigned gx = 0;
    int rb = 0;
    int rc = kv_prepare_col(a->layers, a->n_layers, a->B, a->H, a->rows, a->row_bytes, a->capacity, a->arena_stride_b,
                            a->arena_stride_h, a->new_stride_b, a->new_stride_h, a->new_stride_r, a->seq_lens, a->seq_lens,
                            a->accept_len, 1, a->zero_rejected, a->max_seq_len_host, a->overflow_flag, a->ticket, a->exchange,
                            a->exchange_out, a->exchange_push_in_append ? a->packed : nullptr, a->seq_ready, &kp, &gx, &rb);
    if (rc != SD_OK) return rc;
    if (a->layer_stride_arena != 0 || a->layer_stride_new != 0) {
        if (!a->layer0.arena_k || !a->layer0.arena_v || !a->layer0.new_k || !a->layer0.new_v) return SD_ERR_INVALID_ARG;
        if (a->layer_stride_arena % 16 || a->layer_stride_new % 16) return SD_ERR_ALIGNMENT;
        kp.uniform = 1;
        kp.layer0 = a->layer0;
        kp.ls_arena = a->layer_stride_arena;
        kp.ls_new = a->layer_stride_new;
    }
    if (mode == 1) {
        // A/B switch of the round-1 layout: request the first rows of K and V BEFORE the per-sequence flag wait
        static const int early = [] {
            const char* e = getenv("SD_B200_FUSED_EARLY_ROWS");
            return (e && e[0] == '1') ? 1 : 0;
        }();
        kp.early_rows = early;
    }
    return run_verify(false, a->logits, a->logit_dtype, a->stride_b, a->stride_r, a->draft_ids, a->id_dtype, a->id_stride_b,
                      SD_ACCEPT_GREEDY, 0.f, 1.f, a->accept_len, a->accepted_mask, a->target_ids, a->bonus, a->packed, nullptr,
                      nullptr, nullptr, a->workspace, a->workspace_bytes, a->B, a->rows, a->K, a->V, stream,
                      a->exchange_push_in_append ? nullptr : a->exchange, mode == 1 ? a->seq_ready : nullptr, &kp, gx, rb,
                      mode == 0 ? a->step_state : nullptr);
}

int sd_verify_step(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                   const void* draft_ids, int id_dtype, int64_t id_stride_b, int mode, float threshold,
                   float inv_temperature, int32_t* accept_len, uint8_t* accepted_mask, int32_t* target_ids,
                   int32_t* bonus, int32_t* packed, float* token_prob, float* row_max, float* row_sumexp,
                   void* workspace, size_t workspace_bytes, int B, int rows, int K, int V, void* stream,
                   const sd_exchange* exchange, int32_t* seq_ready) {
    if (!(inv_temperature > 0.f)) return SD_ERR_INVALID_ARG;
    return sd::run_verify(true, logits, logit_dtype, stride_b, stride_r, draft_ids, id_dtype, id_stride_b, mode,
                          threshold, inv_temperature, accept_len, accepted_mask, target_ids, bonus, packed,
                          token_prob, row_max, row_sumexp, workspace, workspace_bytes, B, rows, K, V, stream, exchange, seq_ready);
}

}  // extern "C"
