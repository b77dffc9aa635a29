// Kernel 1: fused verify (argmax over V for B*(K+1) rows + accepted-prefix + bonus token).
//
// Replaces the reference's verify_prefix (src/kernels/reference.py:13-56; Triton incumbent
// src/kernels/triton/verify.py:13-98, one serial program per sequence) with ONE streaming pass.
//
// Work decomposition.  A logits row is cut into UNITS of blockDim.x 16-byte vectors (one vector per
// thread).  A CTA reduces a SEGMENT = a run of consecutive units of one row:
//   variant 1 "persistent" (default): grid = SMs x resident CTAs; the B*rows*units_per_row units are
//           split into gridDim.x contiguous, equal (+-1 unit) ranges, so every SM streams the same
//           number of bytes and there is no partial last wave; a range that crosses a row boundary is
//           processed as two segments.
//   variant 0 "tiled": one CTA per (row, chunk of VPT units) -- the first version, kept for sweeps.
// Per segment:
//   thread: 128-bit ld.global.nc loads, VPT of them issued before first use; running packed max
//           (max.NaN.{bf16x2,f16x2,f32}) + the index of the FIRST vector at which its maximum
//           appeared -- ~1 ALU op per element, so the kernel stays HBM-bound
//   CTA   : REDUX warp reduce of (order_key, min position) -> one partial per (row, segment)
//   the CTA that retires the last unit of a sequence (ticket counter, counted in units) merges the
//           partials of its rows, resolves the exact element inside the winning vector, and writes
//           accept_len / mask / target ids / bonus / packed record.  No second launch, no host trip.
//
// Exact torch.argmax semantics (first maximal index; NaN maximal, first NaN; -0 == +0) are kept
// by (a) canonical-NaN propagating max, (b) folding -0 to +0 once per vector, (c) tie-breaking
// every reduction level on the smallest position.
#include "sd_common.cuh"

namespace sd {

constexpr int kMaxChunks = 64;  // upper bound of S (partials per row) -> workspace size

struct VerifyTuning {
    int threads;
    int vpt;
};
static VerifyTuning g_tuning = {0, 0};
static int g_variant = 0;       // 0 tiled (default: measured faster), 1 persistent
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
    uint32_t* part_key;   // [B*rows*S]
    int32_t* part_pos;    // [B*rows*S]
    float* part_m;        // [B*rows*S] (stats)
    float* part_s;        // [B*rows*S] (stats)
    uint32_t* tickets;    // [B], counts retired units of the sequence
    int B, rows, K, V;
    int S;                // partial slots reserved per row
    int upr;              // units per row (unit = blockDim.x vectors)
    int units_per_chunk;  // tiled variant: units per CTA
    int U, q, rem;        // persistent variant: U = B*rows*upr units = gridDim.x*q + rem
    int mode;
    float threshold, scale_log2;  // scale_log2 = inv_temperature * log2(e)
    float inv_temperature;
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

// Reduce units [it0, it1) of logits row `row` (global row index b*rows + j), publish the partial in
// `slot` of that row, retire `it1 - it0` units of sequence b and, if they were the last ones,
// finalize the sequence.  `nslots` = partial slots in use per row for tiled launches, 0 = derive
// them from the persistent partition.  Called by all threads of the CTA (contains barriers).
template <typename T, int VPT, bool STATS>
SD_DEVINL void process_segment(const VerifyParams& p, const int row, const int it0, const int it1, const int slot,
                               const int nslots) {
    using P = Packed<T>;
    constexpr int EPV = 16 / (int)sizeof(T);  // elements per 16-byte vector
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int b = row / p.rows;
    const int j = row - b * p.rows;

    const T* base = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)j * p.stride_r;
    const int V = p.V;
    // peel to 16-byte alignment (rows of e.g. V=50257 bf16 start at arbitrary 2-byte offsets)
    const int mis = (int)((reinterpret_cast<uintptr_t>(base) & 15u) / sizeof(T));
    int head = mis ? (EPV - mis) : 0;
    if (head > V) head = V;
    const int nvec = (V - head) / EPV;
    const int tail0 = head + nvec * EPV;  // first tail element
    const uint4* vp = reinterpret_cast<const uint4*>(base + head);

    uint32_t run = P::kNegInf;   // replicated scalar running max (packed)
    int first_it = 0x7fffffff;   // vector index (within the row) where `run` first appeared
    float sm = -INFINITY, ss = 0.f;  // STATS: reference max (log2 domain) and sum

    for (int i0 = it0; i0 < it1; i0 += VPT) {
        uint4 buf[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int idx = (i0 + u) * nthreads + tid;
            if (i0 + u < it1 && idx < nvec) {
                buf[u] = ldg_stream_v4(vp + idx);
            } else {
                buf[u] = make_uint4(P::kNegInf, P::kNegInf, P::kNegInf, P::kNegInf);
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const int idx = (i0 + u) * nthreads + tid;
            const bool live = i0 + u < it1 && idx < nvec;
            uint32_t t = P::vmax(P::vmax(buf[u].x, buf[u].y), P::vmax(buf[u].z, buf[u].w));
            t = P::fold(t);
            uint32_t nw = P::vmax(run, t);
            // first live vector always claims the position (run == -inf may equal a row of -inf)
            bool changed = live && (nw != run || first_it == 0x7fffffff);
            if (changed) first_it = idx;
            if constexpr (STATS) {
                if (live && nw != run) {
                    float mn = P::scalar(nw) * p.scale_log2;
                    ss = (sm == -INFINITY) ? 0.f : ss * fast_exp2(sm - mn);
                    sm = mn;
                }
                if (live && sm != -INFINITY) {  // a thread that has only seen -inf adds nothing
                    const uint32_t w[4] = {buf[u].x, buf[u].y, buf[u].z, buf[u].w};
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if constexpr (sizeof(T) == 4) {
                            ss += fast_exp2(fmaf(__uint_as_float(w[q]), p.scale_log2, -sm));
                        } else if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                            ss += fast_exp2(fmaf(__uint_as_float(w[q] << 16), p.scale_log2, -sm));
                            ss += fast_exp2(fmaf(__uint_as_float(w[q] & 0xffff0000u), p.scale_log2, -sm));
                        } else {
                            float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
                            ss += fast_exp2(fmaf(f.x, p.scale_log2, -sm));
                            ss += fast_exp2(fmaf(f.y, p.scale_log2, -sm));
                        }
                    }
                }
            }
            run = nw;
        }
    }

    Cand c;
    c.key = order_key(P::scalar(run));
    c.pos = first_it == 0x7fffffff ? 0x7fffffff : head + first_it * EPV;
    if (first_it == 0x7fffffff) c.key = 0u;  // thread saw nothing

    // scalar peel elements (head + tail): handled by the segment that owns unit 0 of the row
    if (it0 == 0) {
        const int npeel = head + (V - tail0);
        for (int e = tid; e < npeel; e += nthreads) {
            int pos = e < head ? e : tail0 + (e - head);
            float x = to_float<T>(base[pos]);
            Cand c2{order_key(x), pos};
            c = better(c, c2);
            if constexpr (STATS) {
                float xm = x * p.scale_log2;
                merge_ms(sm, ss, xm, 1.f);
            }
        }
    }

    // ---- CTA reduce ----
    __shared__ uint32_t s_key[32];
    __shared__ int s_pos[32];
    __shared__ float s_m[32];
    __shared__ float s_s[32];
    __shared__ int s_last;
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
    __syncthreads();
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
        // exact element inside the winning 16-byte vector (or the peel element itself): resolved per
        // segment (a 16-byte L2 hit, overlapped by the other resident CTAs) so that the sequence
        // finalizer at the end of the kernel only has to merge partials
        {
            const int e = c2.pos + lane;
            bool hit = false;
            if (lane < EPV && c2.pos != 0x7fffffff && e < V) hit = order_key(to_float<T>(base[e])) == c2.key;
            const unsigned ballot = __ballot_sync(0xffffffffu, hit);
            if (ballot) c2.pos += __ffs(ballot) - 1;
        }
        if (lane == 0) {
            const int at = row * p.S + slot;
            p.part_key[at] = c2.key;
            p.part_pos[at] = c2.pos;
            if constexpr (STATS) {
                p.part_m[at] = m2;
                p.part_s[at] = s2;
            }
            __threadfence();
            const unsigned add = (unsigned)(it1 - it0);
            unsigned prev = atomicAdd(&p.tickets[b], add);
            s_last = (prev + add == (unsigned)(p.rows * p.upr));
        }
    }
    __syncthreads();
    if (!s_last) return;

    // ---- this CTA retired the last units of sequence b: finalize the sequence ----
    __threadfence();
    __shared__ int s_tgt[64];
    __shared__ float s_ptok[64];
    // draft ids of this sequence: issued now, consumed by warp 0 after the row merge (K <= 63)
    int my_id[2] = {-1, -1};
    if (wid == 0) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
            if (h * 32 + lane < p.K) my_id[h] = load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + h * 32 + lane);
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
        Cand best{0u, 0x7fffffff};
        float fm = -INFINITY, fs = 0.f;
        for (int s = lane; s < ns; s += 32) {
            // volatile: written by other CTAs in this launch
            uint32_t k = reinterpret_cast<volatile uint32_t*>(p.part_key)[grow * p.S + s];
            int ps = reinterpret_cast<volatile int32_t*>(p.part_pos)[grow * p.S + s];
            best = better(best, Cand{k, ps});
            if constexpr (STATS) {
                float m2 = reinterpret_cast<volatile float*>(p.part_m)[grow * p.S + s];
                float s2 = reinterpret_cast<volatile float*>(p.part_s)[grow * p.S + s];
                merge_ms(fm, fs, m2, s2);
            }
        }
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
                    } else if (tok >= 0 && tok < V) {
                        pt = fast_exp2(xtok * p.scale_log2 - fm) / fs;
                    }
                    if (p.token_prob) p.token_prob[b * p.K + r] = pt;
                }
                s_ptok[r] = pt;
            }
        }
    }
    __syncthreads();
    if (wid == 0) {
        int nacc = p.K;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int k = h * 32 + lane;
            bool bad = false;
            if (k < p.K) {
                if (STATS && p.mode != SD_ACCEPT_GREEDY) bad = !(s_ptok[k] >= p.threshold);
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
            p.tickets[b] = 0u;  // self-reset: the workspace is ready for the next launch / graph replay
        }
    }
    __syncthreads();  // s_tgt / s_last are reused by the caller's next segment
}

// variant 0: one CTA per (row, chunk)
template <typename T, int VPT, bool STATS>
__global__ void __launch_bounds__((VPT >= 8 || STATS) ? 512 : 1024) verify_kernel(const VerifyParams p) {
    pdl_wait();
    pdl_trigger();
    const int unit = blockIdx.x;
    const int chunk = unit % p.S;
    const int row = unit / p.S;
    const int it0 = chunk * p.units_per_chunk;
    int it1 = it0 + p.units_per_chunk;
    if (it1 > p.upr) it1 = p.upr;
    process_segment<T, VPT, STATS>(p, row, it0, it1, chunk, p.S);
}

// variant 1: persistent, equal contiguous unit ranges
template <typename T, int VPT, bool STATS>
__global__ void __launch_bounds__(512) verify_persist_kernel(const VerifyParams p) {
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
        process_segment<T, VPT, STATS>(p, row, it0, it0 + (seg_end - u), c - part_owner(p, row * p.upr), 0);
        u = seg_end;
    }
}

// ------------------------------------------------------------------------------------------ host
static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct WorkspaceLayout {
    size_t key, pos, m, s, tickets, total;
};

static WorkspaceLayout layout(int B, int rows) {
    WorkspaceLayout w;
    size_t n = (size_t)B * rows * kMaxChunks;
    size_t off = 0;
    w.tickets = off;
    off = align_up(off + (size_t)B * 4, 256);
    w.key = off;
    off = align_up(off + n * 4, 256);
    w.pos = off;
    off = align_up(off + n * 4, 256);
    w.m = off;
    off = align_up(off + n * 4, 256);
    w.s = off;
    off = align_up(off + n * 4, 256);
    w.total = off;
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
        case 8: return launch_kernel(verify_kernel<T, 8, STATS>, dim3(grid), dim3(threads), 0, st, p);
        default: return launch_kernel(verify_kernel<T, 16, STATS>, dim3(grid), dim3(threads), 0, st, p);
    }
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
                      int rows, int K, int V, void* stream) {
    if (B < 0 || K < 0 || V <= 0 || rows <= 0) return SD_ERR_INVALID_ARG;
    if (rows != K && rows != K + 1) return SD_ERR_SHAPE;
    if (rows > 64) return SD_ERR_UNSUPPORTED;
    if (B == 0) return SD_OK;
    if (!logits || !accept_len || (K > 0 && !draft_ids)) return SD_ERR_INVALID_ARG;
    if (logit_dtype < SD_F32 || logit_dtype > SD_BF16 || (id_dtype != SD_I32 && id_dtype != SD_I64))
        return SD_ERR_INVALID_ARG;
    if (mode < SD_ACCEPT_GREEDY || mode > SD_ACCEPT_CONFIDENCE) return SD_ERR_INVALID_ARG;
    const int es = elt_size(logit_dtype);
    if (reinterpret_cast<uintptr_t>(logits) % es) return SD_ERR_ALIGNMENT;
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
    const int upr = (nvec_max + threads - 1) / threads;  // units per row
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
    p.tickets = reinterpret_cast<uint32_t*>(ws + w.tickets);
    p.part_key = reinterpret_cast<uint32_t*>(ws + w.key);
    p.part_pos = reinterpret_cast<int32_t*>(ws + w.pos);
    p.part_m = reinterpret_cast<float*>(ws + w.m);
    p.part_s = reinterpret_cast<float*>(ws + w.s);
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

    unsigned grid;
    if (variant == 1) {
        int per_sm = persist_occupancy(logit_dtype, stats, vpt, threads);
        if (g_ctas_per_sm > 0 && g_ctas_per_sm < per_sm) per_sm = g_ctas_per_sm;
        long long G = (long long)sm_count() * per_sm;
        // a CTA should own at least vpt units (one full batch of loads), and a row may be shared by
        // at most kMaxChunks - 2 CTAs (partial slots): q >= upr / (kMaxChunks - 2)
        const int q_min = (upr + (kMaxChunks - 3)) / (kMaxChunks - 2) > vpt ? (upr + (kMaxChunks - 3)) / (kMaxChunks - 2) : vpt;
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
        }
        p.S = S;
        p.units_per_chunk = upc;
        grid = (unsigned)((size_t)B * rows * S);
    }

    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
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
    if (v != 1 && v != 2 && v != 4 && v != 8 && v != 16) return SD_ERR_INVALID_ARG;
    if (threads > 512) return SD_ERR_INVALID_ARG;  // register budget (stats / 8+-vector kernels are built for <= 512)
    sd::g_tuning = {threads, v};
    return SD_OK;
}

int sd_verify_set_variant(int variant, int ctas_per_sm) {
    if (variant < 0 || variant > 1 || ctas_per_sm < 0 || ctas_per_sm > 32) return SD_ERR_INVALID_ARG;
    sd::g_variant = variant;
    sd::g_ctas_per_sm = ctas_per_sm;
    return SD_OK;
}

int sd_verify_prefix(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                     const void* draft_ids, int id_dtype, int64_t id_stride_b, int32_t* accept_len,
                     uint8_t* accepted_mask, int32_t* target_ids, int32_t* bonus, int32_t* packed,
                     void* workspace, size_t workspace_bytes, int B, int rows, int K, int V, void* stream) {
    return sd::run_verify(false, logits, logit_dtype, stride_b, stride_r, draft_ids, id_dtype, id_stride_b,
                          SD_ACCEPT_GREEDY, 0.f, 1.f, accept_len, accepted_mask, target_ids, bonus, packed,
                          nullptr, nullptr, nullptr, workspace, workspace_bytes, B, rows, K, V, stream);
}

int sd_verify_step(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                   const void* draft_ids, int id_dtype, int64_t id_stride_b, int mode, float threshold,
                   float inv_temperature, int32_t* accept_len, uint8_t* accepted_mask, int32_t* target_ids,
                   int32_t* bonus, int32_t* packed, float* token_prob, float* row_max, float* row_sumexp,
                   void* workspace, size_t workspace_bytes, int B, int rows, int K, int V, void* stream) {
    if (!(inv_temperature > 0.f)) return SD_ERR_INVALID_ARG;
    return sd::run_verify(true, logits, logit_dtype, stride_b, stride_r, draft_ids, id_dtype, id_stride_b, mode,
                          threshold, inv_temperature, accept_len, accepted_mask, target_ids, bonus, packed,
                          token_prob, row_max, row_sumexp, workspace, workspace_bytes, B, rows, K, V, stream);
}

}  // extern "C"
