// Kernel 4: batched bonus-token sampler (temperature, top-k, top-p, inverse-CDF draw) for the verify step.
//
// Replaces sample_bonus_token_from_logits (src/specdec/core/pipeline.py:48-147), which the reference
// calls once per sequence: topk(V) + sort(V) + softmax + cumsum + multinomial on a [1, V] row (the sort of
// 128 k elements dominates).  Here one CTA handles one sequence and never sorts:
//   * the row is selected on the device (row_index = the accept_len the verify kernel just wrote);
//   * top-k and top-p are both "keep a prefix of the descending order", i.e. a threshold key plus a
//     number of ties at the threshold (ties ordered by index, as a stable descending sort) -- found by
//     an 8-bit-digit radix select over the order keys: by COUNT for top-k, by probability MASS for
//     top-p (cumulative mass <= top_p, first element always kept: pipeline.py:113-115);
//   * masses are accumulated in 2^-40 fixed point with integer atomics, so every sum is independent
//     of the order of accumulation: the result is a deterministic function of (logits, u);
//   * the draw is the inverse CDF in INDEX order over the kept set: the first token t whose cumulative
//     kept mass exceeds u * total.  (torch.multinomial's RNG stream cannot be reproduced -- SURVEY.md
//     section 7 -- so the caller supplies u in [0,1); the distribution sampled is the reference's.)
//   * u == NULL is the greedy branch (pipeline.py:135-137): argmax, torch semantics (first index).
// The row (256 KB at V=128256 bf16) is re-read ~6-10 times; after the first pass it is L2-resident.
#include "sd_common.cuh"

namespace sd {

constexpr int kSampThreads = 1024;
constexpr int kSampWarps = kSampThreads / 32;
constexpr int kBins = 256;

struct SampleParams {
    const void* logits;
    int64_t stride_b, stride_r;
    const int32_t* row_index;  // nullable
    int row_bias, row_const, rows, V, B;
    const float* uniform;      // nullable -> greedy
    float inv_t_log2;          // log2(e) / temperature
    int top_k;                 // 0 = off
    float top_p;               // >= 1 = off
    int32_t* out;
    int32_t* kept_count;       // nullable: size of the final candidate set (diagnostics / tests)
    int ndigits;               // radix passes of 8 bits: bf16 2, fp16 3, fp32 4
};

struct SelectResult {
    uint32_t tau;              // threshold key
    unsigned n_ties;           // elements with key == tau that are kept (lowest indices first)
};

SD_DEVINL unsigned long long fixed_mass(float x, float m, float inv_t_log2) {
    // e = 2^((x - m) * log2e / T) in [0, 1]; 2^-40 fixed point
    const float e = fast_exp2((x - m) * inv_t_log2);
    return (unsigned long long)(e * 1099511627776.0f);
}

// kept-set membership for a (tau, n_ties) prefix: `tie_rank` = number of earlier (lower-index) elements
// with key == tau
SD_DEVINL bool in_prefix(uint32_t key, unsigned tie_rank, const SelectResult& s) {
    return key > s.tau || (key == s.tau && tie_rank < s.n_ties);
}

// Block-wide exclusive scan of (u64 mass, u32 count) over the threads in thread order; returns totals.
struct ScanPair {
    unsigned long long mass;
    unsigned cnt;
};
SD_DEVINL ScanPair block_exclusive_scan(ScanPair v, ScanPair& total, unsigned long long* s_mass, unsigned* s_cnt) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    ScanPair inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned long long m2 = __shfl_up_sync(0xffffffffu, inc.mass, o);
        unsigned c2 = __shfl_up_sync(0xffffffffu, inc.cnt, o);
        if (lane >= o) {
            inc.mass += m2;
            inc.cnt += c2;
        }
    }
    if (lane == 31) {
        s_mass[wid] = inc.mass;
        s_cnt[wid] = inc.cnt;
    }
    __syncthreads();
    if (wid == 0) {
        unsigned long long wm = s_mass[lane];
        unsigned wc = s_cnt[lane];
        unsigned long long im = wm;
        unsigned ic = wc;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned long long m2 = __shfl_up_sync(0xffffffffu, im, o);
            unsigned c2 = __shfl_up_sync(0xffffffffu, ic, o);
            if (lane >= o) {
                im += m2;
                ic += c2;
            }
        }
        s_mass[lane] = im - wm;  // exclusive prefix of the warp totals
        s_cnt[lane] = ic - wc;
        if (lane == 31) {
            s_mass[32] = im;
            s_cnt[32] = ic;
        }
    }
    __syncthreads();
    ScanPair ex;
    ex.mass = inc.mass - v.mass + s_mass[wid];
    ex.cnt = inc.cnt - v.cnt + s_cnt[wid];
    total.mass = s_mass[32];
    total.cnt = s_cnt[32];
    __syncthreads();  // scratch is reused by the next scan
    return ex;
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads) sample_bonus_kernel(const SampleParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    // per-warp private histograms: counts [warps][256] u32, masses [warps][256] u64
    unsigned* h_cnt = reinterpret_cast<unsigned*>(smem_raw);
    unsigned long long* h_mass = reinterpret_cast<unsigned long long*>(smem_raw + kSampWarps * kBins * 4);
    __shared__ unsigned long long s_mass[33];
    __shared__ unsigned s_cnt[33];
    __shared__ unsigned tot_cnt[kBins];
    __shared__ unsigned long long tot_mass[kBins];
    __shared__ uint32_t s_u32[4];
    __shared__ unsigned long long s_u64[2];
    __shared__ uint32_t s_rkey[kSampWarps];
    __shared__ int s_rpos[kSampWarps];

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.x;
    int r = p.row_const;
    if (p.row_index) {
        r = p.row_index[b] + p.row_bias;
        r = r < 0 ? 0 : (r >= p.rows ? p.rows - 1 : r);
    }
    const T* row = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)r * p.stride_r;
    const int V = p.V;

    // ---- pass A: maximum (order key) and its first index ----
    uint32_t kmax = 0u;
    int amax = 0x7fffffff;
    for (int i = tid; i < V; i += kSampThreads) {
        const uint32_t k = order_key(to_float<T>(row[i]));
        if (k > kmax) {  // strictly greater: the first index wins within a thread (indices ascend)
            kmax = k;
            amax = i;
        }
    }
    {
        uint32_t wk = __reduce_max_sync(0xffffffffu, kmax);
        unsigned wp = __reduce_min_sync(0xffffffffu, kmax == wk ? (unsigned)amax : 0x7fffffffu);
        if (lane == 0) {
            s_rkey[wid] = wk;
            s_rpos[wid] = (int)wp;
        }
        __syncthreads();
        if (wid == 0) {
            uint32_t k2 = s_rkey[lane];
            unsigned p2 = (unsigned)s_rpos[lane];
            uint32_t bk = __reduce_max_sync(0xffffffffu, k2);
            unsigned bp = __reduce_min_sync(0xffffffffu, k2 == bk ? p2 : 0x7fffffffu);
            if (lane == 0) {
                s_u32[0] = bk;
                s_u32[1] = bp;
            }
        }
        __syncthreads();
        kmax = s_u32[0];
        amax = (int)s_u32[1];
        __syncthreads();
    }
    if (!p.uniform) {  // greedy branch
        if (tid == 0) {
            p.out[b] = amax;
            if (p.kept_count) p.kept_count[b] = 1;
        }
        return;
    }
    const float m = key_to_float(kmax);

    // Radix select over the order keys of the elements inside `outer` (a previously selected prefix):
    //   by_mass == false: largest prefix of the descending order with at most `k_target` elements
    //   by_mass == true : largest prefix whose mass is <= mass_target (at least one element)
    auto radix_select = [&](bool by_mass, unsigned k_target, unsigned long long mass_target,
                            const SelectResult& outer, bool has_outer) -> SelectResult {
        uint32_t prefix = 0u, prefix_mask = 0u;
        unsigned cnt_above = 0;               // elements with key > current prefix range (all kept)
        unsigned long long mass_above = 0ull;
        for (int d = 0; d < p.ndigits; ++d) {
            const int shift = 24 - 8 * d;
            for (int i = tid; i < kSampWarps * kBins; i += kSampThreads) {
                h_cnt[i] = 0u;
                h_mass[i] = 0ull;
            }
            __syncthreads();
            // NOTE: ties of `outer` are ordered by index; an element equal to outer.tau is inside the outer
            // prefix only if its tie rank is < outer.n_ties.  Tie ranks need an index-ordered count, which
            // the histogram passes do not have; instead the outer threshold bucket is handled exactly
            // below: elements with key == outer.tau enter the histogram with weight (n_ties / n_equal),
            // realised by counting them all here and clamping the bucket totals afterwards.
            for (int i = tid; i < V; i += kSampThreads) {
                const float x = to_float<T>(row[i]);
                const uint32_t k = order_key(x);
                if (has_outer && k < outer.tau) continue;
                if ((k & prefix_mask) != prefix) continue;
                const unsigned bin = (k >> shift) & 0xffu;
                atomicAdd(&h_cnt[wid * kBins + bin], 1u);
                if (by_mass) atomicAdd(&h_mass[wid * kBins + bin], fixed_mass(x, m, p.inv_t_log2));
            }
            __syncthreads();
            if (tid < kBins) {
                unsigned c = 0;
                unsigned long long w = 0ull;
                for (int q = 0; q < kSampWarps; ++q) {
                    c += h_cnt[q * kBins + tid];
                    w += h_mass[q * kBins + tid];
                }
                tot_cnt[tid] = c;
                tot_mass[tid] = w;
            }
            __syncthreads();
            // the bucket holding outer.tau may contain more ties than the outer prefix keeps: clamp it
            if (has_outer && tid == 0 && (outer.tau & prefix_mask) == prefix && d == p.ndigits - 1) {
                const unsigned bin = (outer.tau >> shift) & 0xffu;
                if (tot_cnt[bin] > outer.n_ties) {
                    const unsigned long long each = tot_cnt[bin] ? tot_mass[bin] / tot_cnt[bin] : 0ull;
                    tot_cnt[bin] = outer.n_ties;
                    tot_mass[bin] = each * outer.n_ties;
                }
            }
            __syncthreads();
            if (tid == 0) {  // walk the 256 buckets from the top
                unsigned c = cnt_above;
                unsigned long long w = mass_above;
                int chosen = 0;
                for (int bin = kBins - 1; bin >= 0; --bin) {
                    const unsigned c2 = c + tot_cnt[bin];
                    const unsigned long long w2 = w + tot_mass[bin];
                    const bool over = by_mass ? (w2 > mass_target) : (c2 > k_target);
                    if (tot_cnt[bin] && over) {
                        chosen = bin;
                        break;
                    }
                    c = c2;
                    w = w2;
                    chosen = bin;  // everything fits: end on the lowest non-crossing bucket
                    if (bin == 0) {
                        // nothing crossed: the whole candidate range is kept; signal with bin = -1
                        chosen = -1;
                    }
                }
                s_u32[0] = (uint32_t)chosen;
                s_u32[1] = c;
                s_u64[0] = w;
            }
            __syncthreads();
            const int chosen = (int)s_u32[0];
            cnt_above = s_u32[1];
            mass_above = s_u64[0];
            __syncthreads();
            if (chosen < 0) {
                // every candidate is kept: the prefix is the outer prefix itself (or everything)
                SelectResult all;
                all.tau = has_outer ? outer.tau : 0u;
                all.n_ties = has_outer ? outer.n_ties : 0xffffffffu;
                return all;
            }
            prefix |= (uint32_t)chosen << shift;
            prefix_mask |= 0xffu << shift;
        }
        // prefix is now an exact key value (for 16-bit inputs the low 16 bits of every candidate key are a
        // constant pattern that the selected prefix does not include: recover the full key from one
        // candidate below).  Elements with key > tau: cnt_above / mass_above.  Ties at tau:
        SelectResult res;
        // full key of the threshold: any element matching the prefix (all such elements are equal in
        // the digits examined; for ndigits == 2 they are equal altogether because the inputs are 16 bit)
        if (tid == 0) s_u32[2] = 0u;
        __syncthreads();
        for (int i = tid; i < V; i += kSampThreads) {
            const uint32_t k = order_key(to_float<T>(row[i]));
            if ((k & prefix_mask) == prefix) s_u32[2] = k;  // benign race: all writers store the same value
        }
        __syncthreads();
        res.tau = s_u32[2];
        __syncthreads();
        // number of ties available at tau inside the outer prefix, and mass of one tie
        unsigned n_equal = 0;
        for (int i = tid; i < V; i += kSampThreads) n_equal += order_key(to_float<T>(row[i])) == res.tau ? 1u : 0u;
        n_equal = __reduce_add_sync(0xffffffffu, n_equal);
        if (lane == 0) s_cnt[wid] = n_equal;
        __syncthreads();
        if (tid == 0) {
            unsigned t = 0;
            for (int q = 0; q < kSampWarps; ++q) t += s_cnt[q];
            if (has_outer && res.tau == outer.tau && t > outer.n_ties) t = outer.n_ties;
            unsigned keep;
            if (by_mass) {
                const unsigned long long each = fixed_mass(key_to_float(res.tau), m, p.inv_t_log2);
                const unsigned long long room = mass_target > mass_above ? mass_target - mass_above : 0ull;
                unsigned long long fit = each ? room / each : (unsigned long long)t;
                keep = fit > t ? t : (unsigned)fit;
                if (cnt_above == 0 && keep == 0) keep = 1;  // "keep at least one token" (pipeline.py:115)
            } else {
                keep = k_target > cnt_above ? k_target - cnt_above : 0u;
                if (keep > t) keep = t;
            }
            s_u32[3] = keep;
        }
        __syncthreads();
        res.n_ties = s_u32[3];
        __syncthreads();
        return res;
    };

    SelectResult sel;
    sel.tau = 0u;
    sel.n_ties = 0xffffffffu;  // everything
    bool has_sel = false;
    if (p.top_k > 0 && p.top_k < V) {
        sel = radix_select(false, (unsigned)p.top_k, 0ull, sel, false);
        has_sel = true;
    }

    // ---- index-ordered passes over the kept set: total mass, then (optionally) the nucleus, then the draw ----
    // Tiles of 1024 x 8 consecutive elements keep the index order inside the block scan.
    auto kept_mass_total = [&](const SelectResult& s, bool restricted) -> unsigned long long {
        unsigned long long run_mass = 0ull;
        unsigned run_ties = 0;
        for (int t0 = 0; t0 < V; t0 += kSampThreads * 8) {
            const int i0 = t0 + tid * 8;
            float xs[8];
            uint32_t ks[8];
            unsigned my_ties = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + e;
                xs[e] = i < V ? to_float<T>(row[i]) : 0.f;
                ks[e] = i < V ? order_key(xs[e]) : 0u;
                if (i < V && restricted && ks[e] == s.tau) ++my_ties;
            }
            ScanPair v{0ull, my_ties}, total;
            ScanPair ex = block_exclusive_scan(v, total, s_mass, s_cnt);
            unsigned rank = run_ties + ex.cnt;
            unsigned long long mine = 0ull;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const int i = i0 + e;
                if (i >= V) continue;
                bool keep = true;
                if (restricted) {
                    keep = in_prefix(ks[e], rank, s);
                    if (ks[e] == s.tau) ++rank;
                }
                if (keep) mine += fixed_mass(xs[e], m, p.inv_t_log2);
            }
            // block sum of `mine`
            ScanPair v2{mine, 0u}, total2;
            block_exclusive_scan(v2, total2, s_mass, s_cnt);
            run_mass += total2.mass;
            run_ties += total.cnt;
        }
        return run_mass;
    };

    if (p.top_p < 1.0f) {
        const unsigned long long z1 = kept_mass_total(sel, has_sel);
        // cumulative probability <= top_p  <=>  cumulative mass <= top_p * Z1
        const unsigned long long target = (unsigned long long)((double)p.top_p * (double)z1);
        SelectResult nucleus = radix_select(true, 0u, target, sel, has_sel);
        if (has_sel && nucleus.tau == sel.tau && nucleus.n_ties > sel.n_ties) nucleus.n_ties = sel.n_ties;
        sel = nucleus;
        has_sel = true;
    }

    // ---- the draw: first index whose cumulative kept mass exceeds u * Z ----
    const unsigned long long z = kept_mass_total(sel, has_sel);
    float u = p.uniform[b];
    u = u < 0.f ? 0.f : (u > 0.99999994f ? 0.99999994f : u);
    const unsigned long long target = (unsigned long long)((double)u * (double)z);
    if (tid == 0) {
        s_u32[0] = 0xffffffffu;  // result index
        s_u32[1] = 0u;           // kept count
    }
    __syncthreads();
    unsigned long long run_mass = 0ull;
    unsigned run_ties = 0, run_kept = 0;
    bool done = false;
    for (int t0 = 0; t0 < V && !done; t0 += kSampThreads * 8) {
        const int i0 = t0 + tid * 8;
        float xs[8];
        uint32_t ks[8];
        unsigned my_ties = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = i0 + e;
            xs[e] = i < V ? to_float<T>(row[i]) : 0.f;
            ks[e] = i < V ? order_key(xs[e]) : 0u;
            if (i < V && has_sel && ks[e] == sel.tau) ++my_ties;
        }
        ScanPair v{0ull, my_ties}, total;
        ScanPair ex = block_exclusive_scan(v, total, s_mass, s_cnt);
        unsigned rank = run_ties + ex.cnt;
        unsigned long long w[8];
        unsigned long long mine = 0ull;
        unsigned mine_kept = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int i = i0 + e;
            w[e] = 0ull;
            if (i >= V) continue;
            bool keep = true;
            if (has_sel) {
                keep = in_prefix(ks[e], rank, sel);
                if (ks[e] == sel.tau) ++rank;
            }
            if (keep) {
                w[e] = fixed_mass(xs[e], m, p.inv_t_log2);
                mine += w[e];
                ++mine_kept;
            }
        }
        ScanPair v2{mine, mine_kept}, total2;
        ScanPair ex2 = block_exclusive_scan(v2, total2, s_mass, s_cnt);
        // does the target fall inside this thread's 8 elements?
        unsigned long long before = run_mass + ex2.mass;
        if (s_u32[0] == 0xffffffffu && before <= target && target < before + mine) {
            unsigned long long c = before;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (w[e] && target < c + w[e]) {
                    s_u32[0] = (uint32_t)(i0 + e);  // exactly one thread satisfies the interval test
                    break;
                }
                c += w[e];
            }
        }
        run_mass += total2.mass;
        run_ties += total.cnt;
        run_kept += total2.cnt;
        __syncthreads();
        if (s_u32[0] != 0xffffffffu && !p.kept_count) done = true;  // early exit unless the set size is wanted
    }
    if (tid == 0) {
        // rounding of the fixed-point masses can leave the target at the very end: fall back to the argmax
        p.out[b] = s_u32[0] == 0xffffffffu ? amax : (int)s_u32[0];
        if (p.kept_count) p.kept_count[b] = (int)run_kept;
    }
}

}  // namespace sd

extern "C" {

int sd_sample_bonus(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r, const int32_t* row_index,
                    int row_bias, int row_const, const float* uniform, float temperature, int top_k, float top_p,
                    int32_t* out_tokens, int32_t* kept_count, int B, int rows, int V, void* stream) {
    using namespace sd;
    if (B < 0 || rows <= 0 || V <= 0) return SD_ERR_INVALID_ARG;
    if (B == 0) return SD_OK;
    if (!logits || !out_tokens) return SD_ERR_INVALID_ARG;
    if (logit_dtype < SD_F32 || logit_dtype > SD_BF16) return SD_ERR_INVALID_ARG;
    if (uniform && !(temperature > 0.f)) return SD_ERR_INVALID_ARG;
    if (!row_index && (row_const < 0 || row_const >= rows)) return SD_ERR_SHAPE;
    SampleParams p;
    p.logits = logits;
    p.stride_b = stride_b;
    p.stride_r = stride_r;
    p.row_index = row_index;
    p.row_bias = row_bias;
    p.row_const = row_const;
    p.rows = rows;
    p.V = V;
    p.B = B;
    p.uniform = uniform;
    p.inv_t_log2 = 1.4426950408889634f / (temperature > 0.f ? temperature : 1.f);
    p.top_k = top_k > 0 ? top_k : 0;
    p.top_p = top_p > 0.f ? top_p : 1.0f;
    p.out = out_tokens;
    p.kept_count = kept_count;
    // digits that cover every information bit of the order key: fp32 32, fp16 19 (1+8+10 after widening), bf16 16
    p.ndigits = logit_dtype == SD_F32 ? 4 : (logit_dtype == SD_F16 ? 3 : 2);
    const size_t smem = (size_t)kSampWarps * kBins * (4 + 8);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (logit_dtype == SD_F32) {
        e = cudaFuncSetAttribute(sample_bonus_kernel<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = launch_kernel(sample_bonus_kernel<float>, dim3(B), dim3(kSampThreads), smem, st, p);
    } else if (logit_dtype == SD_F16) {
        e = cudaFuncSetAttribute(sample_bonus_kernel<__half>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = launch_kernel(sample_bonus_kernel<__half>, dim3(B), dim3(kSampThreads), smem, st, p);
    } else {
        e = cudaFuncSetAttribute(sample_bonus_kernel<__nv_bfloat16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = launch_kernel(sample_bonus_kernel<__nv_bfloat16>, dim3(B), dim3(kSampThreads), smem, st, p);
    }
    return to_status(e);
}

}  // extern "C"
