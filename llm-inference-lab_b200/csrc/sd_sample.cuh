// Device building blocks of the bonus-token sampler (sd_sample.cu): fixed-point masses, the block scan, the row
// accessors, the radix select (by count / by mass), the index-ordered kept-set walk and the nucleus + draw.
// All of them assume a CTA of kSampThreads threads.
#pragma once

#include <type_traits>

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
    int no_fast_path;          // 1 = skip topk_fast_path (tests compare both paths)
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
        unsigned long long wm = lane < kSampWarps ? s_mass[lane] : 0ull;
        unsigned wc = lane < kSampWarps ? s_cnt[lane] : 0u;
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

// ---- row accessors: the same selection / draw code runs over the logits row in global memory and over
// the compacted top-k candidates in shared memory ----
template <typename T>
struct GlobalRow {
    const T* row;
    int n;
    SD_DEVINL int size() const { return n; }
    SD_DEVINL int token(int i) const { return i; }
    // 8 consecutive elements starting at i0 (a multiple of 8); one 16-byte load when the address allows it
    SD_DEVINL void load8(int i0, float (&xs)[8], bool (&ok)[8]) const {
        if constexpr (sizeof(T) == 2) {
            if (i0 + 8 <= n && (reinterpret_cast<uintptr_t>(row + i0) & 15u) == 0) {
                const uint4 v = *reinterpret_cast<const uint4*>(row + i0);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                        xs[2 * q] = __uint_as_float(w[q] << 16);
                        xs[2 * q + 1] = __uint_as_float(w[q] & 0xffff0000u);
                    } else {
                        const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[q]));
                        xs[2 * q] = f.x;
                        xs[2 * q + 1] = f.y;
                    }
                    ok[2 * q] = ok[2 * q + 1] = true;
                }
                return;
            }
        }
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ok[e] = i0 + e < n;
            xs[e] = ok[e] ? to_float<T>(row[i0 + e]) : 0.f;
        }
    }
};

struct CompactRow {
    const float* x;
    const int* idx;
    int n;
    SD_DEVINL int size() const { return n; }
    SD_DEVINL int token(int i) const { return idx[i]; }
    SD_DEVINL void load8(int i0, float (&xs)[8], bool (&ok)[8]) const {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ok[e] = i0 + e < n;
            xs[e] = ok[e] ? x[i0 + e] : 0.f;
        }
    }
};

// shared scratch of one CTA
struct SampShared {
    unsigned* h_cnt;             // [warps][256]
    unsigned long long* h_mass;  // [warps][256]
    unsigned long long* s_mass;  // [33]
    unsigned* s_cnt;             // [33]
    unsigned* tot_cnt;           // [256]
    unsigned long long* tot_mass;  // [256]
    uint32_t* s_u32;             // [4]
    unsigned long long* s_u64;   // [2]
};

constexpr int kBatch = 2;                 // tiles whose loads are issued together (memory-level parallelism)
constexpr int kTile = kSampThreads * 8;  // elements per tile: thread t owns [t*8, t*8+8) of the tile (index order)

// Radix select over the order keys of the elements inside `outer` (a previously selected prefix):
//   by_mass == false: largest prefix of the descending order with at most `k_target` elements
//   by_mass == true : largest prefix whose mass is <= mass_target (at least one element)
template <typename Row>
SD_DEVINL SelectResult radix_select(const Row& R, const SampShared& sh, const SampleParams& p, float m, bool by_mass,
                                    unsigned k_target, unsigned long long mass_target, const SelectResult outer,
                                    bool has_outer) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = R.size();
    uint32_t prefix = 0u, prefix_mask = 0u;
    unsigned cnt_above = 0;  // elements with key above the current prefix range (all kept)
    unsigned long long mass_above = 0ull;
    for (int d = 0; d < p.ndigits; ++d) {
        const int shift = 24 - 8 * d;
        for (int i = tid; i < kSampWarps * kBins; i += kSampThreads) {
            sh.h_cnt[i] = 0u;
            sh.h_mass[i] = 0ull;
        }
        __syncthreads();
        // Elements equal to outer.tau are all counted here although the outer prefix keeps only n_ties of
        // them; the bucket that holds outer.tau is clamped in the last digit pass (in earlier passes the
        // over-count can only make the walk descend into the lowest bucket, which the last pass corrects).
        for (int t0 = 0; t0 < n; t0 += kTile * kBatch) {
            float xs[kBatch][8];
            bool ok[kBatch][8];
#pragma unroll
            for (int q = 0; q < kBatch; ++q) R.load8(t0 + q * kTile + tid * 8, xs[q], ok[q]);  // kBatch tiles in flight
#pragma unroll
            for (int q = 0; q < kBatch; ++q) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    unsigned bin = 0xffffffffu;
                    const uint32_t k = order_key(xs[q][e]);
                    if (ok[q][e] && !(has_outer && k < outer.tau) && (k & prefix_mask) == prefix) bin = (k >> shift) & 0xffu;
                    if (by_mass) {
                        if (bin != 0xffffffffu) {
                            atomicAdd(&sh.h_cnt[wid * kBins + bin], 1u);
                            atomicAdd(&sh.h_mass[wid * kBins + bin], fixed_mass(xs[q][e], m, p.inv_t_log2));
                        }
                    } else {
                        // warp-aggregated count: logits share their exponent byte, so most lanes hit one bin
                        const unsigned peers = __match_any_sync(0xffffffffu, bin);
                        if (bin != 0xffffffffu && lane == __ffs(peers) - 1)
                            atomicAdd(&sh.h_cnt[wid * kBins + bin], __popc(peers));
                    }
                }
            }
        }
        __syncthreads();
        if (tid < kBins) {
            unsigned c = 0;
            unsigned long long w = 0ull;
            for (int q = 0; q < kSampWarps; ++q) {
                c += sh.h_cnt[q * kBins + tid];
                w += sh.h_mass[q * kBins + tid];
            }
            sh.tot_cnt[tid] = c;
            sh.tot_mass[tid] = w;
        }
        __syncthreads();
        if (has_outer && tid == 0 && (outer.tau & prefix_mask) == prefix && d == p.ndigits - 1) {
            const unsigned bin = (outer.tau >> shift) & 0xffu;
            if (sh.tot_cnt[bin] > outer.n_ties) {
                const unsigned long long each = sh.tot_cnt[bin] ? sh.tot_mass[bin] / sh.tot_cnt[bin] : 0ull;
                sh.tot_cnt[bin] = outer.n_ties;
                sh.tot_mass[bin] = each * outer.n_ties;
            }
        }
        __syncthreads();
        if (tid == 0) {  // walk the 256 buckets from the top
            unsigned c = cnt_above;
            unsigned long long w = mass_above;
            int chosen = -1;  // -1: nothing crosses, the whole candidate range is kept
            for (int bin = kBins - 1; bin >= 0; --bin) {
                const unsigned c2 = c + sh.tot_cnt[bin];
                const unsigned long long w2 = w + sh.tot_mass[bin];
                const bool over = by_mass ? (w2 > mass_target) : (c2 > k_target);
                if (sh.tot_cnt[bin] && over) {
                    chosen = bin;
                    break;
                }
                c = c2;
                w = w2;
            }
            sh.s_u32[0] = (uint32_t)chosen;
            sh.s_u32[1] = c;
            sh.s_u64[0] = w;
        }
        __syncthreads();
        const int chosen = (int)sh.s_u32[0];
        cnt_above = sh.s_u32[1];
        mass_above = sh.s_u64[0];
        __syncthreads();
        if (chosen < 0) {
            SelectResult all;
            all.tau = has_outer ? outer.tau : 0u;
            all.n_ties = has_outer ? outer.n_ties : 0xffffffffu;
            return all;
        }
        prefix |= (uint32_t)chosen << shift;
        prefix_mask |= 0xffu << shift;
    }
    // The digits examined cover every information bit of the input type, so the chosen bucket of the last
    // pass IS one value: its count is the number of ties, and the key bits below the examined digits follow
    // from the sign (order_key: positive -> bits of x, zero-extended; negative -> complement, one-extended).
    SelectResult res;
    const uint32_t low_mask = ~prefix_mask;
    res.tau = (prefix & 0x80000000u) ? prefix : (prefix | low_mask);
    if (tid == 0) {
        unsigned t = sh.tot_cnt[(prefix >> (24 - 8 * (p.ndigits - 1))) & 0xffu];  // already clamped to the outer prefix
        if (has_outer && res.tau == outer.tau && t > outer.n_ties) t = outer.n_ties;
        unsigned keep;
        if (by_mass) {
            const unsigned long long each = fixed_mass(key_to_float(res.tau), m, p.inv_t_log2);
            const unsigned long long room = mass_target > mass_above ? mass_target - mass_above : 0ull;
            const unsigned long long fit = each ? room / each : (unsigned long long)t;
            keep = fit > t ? t : (unsigned)fit;
            if (cnt_above == 0 && keep == 0) keep = 1;  // "keep at least one token" (pipeline.py:115)
        } else {
            keep = k_target > cnt_above ? k_target - cnt_above : 0u;
            if (keep > t) keep = t;
        }
        sh.s_u32[3] = keep == t ? 0xffffffffu : keep;  // every tie kept: no index ranking needed downstream
    }
    __syncthreads();
    res.n_ties = sh.s_u32[3];
    __syncthreads();
    return res;
}

// One index-ordered pass over the row: visits every element with its membership in the (tau, n_ties)
// prefix `s` (tie ranks by index via a block scan per tile).  `visit(tile_base, xs, ok, keep)` is called by
// every thread once per tile and may itself contain block-wide collectives.
template <typename Row, typename Visit>
SD_DEVINL void for_each_kept(const Row& R, const SampShared& sh, const SelectResult& s, bool restricted, Visit&& visit) {
    const int tid = threadIdx.x;
    const bool rank_ties = restricted && s.n_ties != 0xffffffffu;  // else: keep = key >= tau, no scan needed
    unsigned run_ties = 0;
    for (int tb = 0; tb < R.size(); tb += kTile * kBatch) {
        float xs[kBatch][8];
        bool ok[kBatch][8];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) R.load8(tb + q * kTile + tid * 8, xs[q], ok[q]);  // kBatch tiles in flight
        bool go_on = true;
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
            const int t0 = tb + q * kTile;
            if (t0 >= R.size() || !go_on) continue;  // uniform across the CTA
            bool keep[8];
            uint32_t ks[8];
            unsigned my_ties = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                ks[e] = order_key(xs[q][e]);
                if (ok[q][e] && rank_ties && ks[e] == s.tau) ++my_ties;
            }
            unsigned rank = 0;
            if (rank_ties) {
                ScanPair v{0ull, my_ties}, total;
                const ScanPair ex = block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
                rank = run_ties + ex.cnt;
                run_ties += total.cnt;
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                keep[e] = ok[q][e];
                if (ok[q][e] && restricted) {
                    keep[e] = rank_ties ? in_prefix(ks[e], rank, s) : ks[e] >= s.tau;
                    if (rank_ties && ks[e] == s.tau) ++rank;
                }
            }
            go_on = visit(t0, xs[q], keep);
        }
        if (!go_on) break;
    }
}

template <typename Row>
SD_DEVINL unsigned long long kept_mass_total(const Row& R, const SampShared& sh, const SampleParams& p, float m,
                                             const SelectResult& s, bool restricted) {
    unsigned long long mine = 0ull;
    for_each_kept(R, sh, s, restricted, [&](int, const float (&xs)[8], const bool (&keep)[8]) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (keep[e]) mine += fixed_mass(xs[e], m, p.inv_t_log2);
        return true;
    });
    ScanPair v{mine, 0u}, total;
    block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
    return total.mass;
}

// top-p (optional) + inverse-CDF draw over the kept prefix `sel` of row R.  Returns the TOKEN id on thread 0
// (0xffffffff if the fixed-point rounding left the target past the end) and the kept-set size.
template <typename Row>
SD_DEVINL void nucleus_and_draw(const Row& R, const SampShared& sh, const SampleParams& p, float m, SelectResult sel,
                                bool has_sel, float u, bool want_count, uint32_t& token_out, unsigned& kept_out) {
    const int tid = threadIdx.x;
    if (p.top_p < 1.0f) {
        const unsigned long long z1 = kept_mass_total(R, sh, p, m, sel, has_sel);
        // cumulative probability <= top_p  <=>  cumulative mass <= top_p * Z1
        const unsigned long long target = (unsigned long long)((double)p.top_p * (double)z1);
        SelectResult nucleus = radix_select(R, sh, p, m, true, 0u, target, sel, has_sel);
        if (has_sel && nucleus.tau == sel.tau && nucleus.n_ties > sel.n_ties) nucleus.n_ties = sel.n_ties;
        sel = nucleus;
        has_sel = true;
    }
    const unsigned long long z = kept_mass_total(R, sh, p, m, sel, has_sel);
    u = u < 0.f ? 0.f : (u > 0.99999994f ? 0.99999994f : u);
    const unsigned long long target = (unsigned long long)((double)u * (double)z);
    if (tid == 0) sh.s_u32[0] = 0xffffffffu;
    __syncthreads();
    unsigned long long run_mass = 0ull;
    unsigned run_kept = 0;
    for_each_kept(R, sh, sel, has_sel, [&](int t0, const float (&xs)[8], const bool (&keep)[8]) {
        unsigned long long w[8];
        unsigned long long mine = 0ull;
        unsigned mine_kept = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            w[e] = keep[e] ? fixed_mass(xs[e], m, p.inv_t_log2) : 0ull;
            mine += w[e];
            mine_kept += keep[e] ? 1u : 0u;
        }
        ScanPair v{mine, mine_kept}, total;
        const ScanPair ex = block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
        const unsigned long long before = run_mass + ex.mass;
        if (before <= target && target < before + mine) {  // exactly one thread owns the target
            unsigned long long c = before;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                if (w[e] && target < c + w[e]) {
                    sh.s_u32[0] = (uint32_t)R.token(t0 + tid * 8 + e);
                    break;
                }
                c += w[e];
            }
        }
        run_mass += total.mass;
        run_kept += total.cnt;
        __syncthreads();
        return want_count || sh.s_u32[0] == 0xffffffffu;  // early exit once drawn, unless the set size is wanted
    });
    __syncthreads();
    token_out = sh.s_u32[0];
    kept_out = run_kept;
    __syncthreads();
}

constexpr int kCompactCap = 1024;  // top-k values up to this size continue on a compacted candidate list in smem

// ---- fast path for 0 < top_k <= kCompactCap: two passes over the row, everything else on <= 1024 candidates --------
// The general path above finds the top-k threshold with an 8-bit radix select over the WHOLE row (one pass per digit,
// a warp-wide match per element), compacts the survivors in index order (a block scan per tile) and runs the by-mass
// select on them: ~4 row passes and a dozen 256-bucket walks, 157 us at B = 64, V = 128256.  Here:
//   1. the k-th largest of the 1024 PER-THREAD maxima (already computed by the maximum pass) is a lower bound of the
//      k-th largest element of the row, so {key >= that bound} is a superset of the top-k -- for a row that is not
//      adversarially ordered a few elements more than k;
//   2. ONE more pass over the row (L2-resident) collects that superset, unordered, into shared memory; more than
//      kCompactCap of them (a sorted or near-constant row) -> return false, the general path takes over;
//   3. a bitonic sort by (key descending, index ascending) makes the first k entries exactly the reference's top-k
//      (ties at the threshold: lowest indices first, as a stable descending sort); top-p is a prefix of that list
//      (inclusive cumulative fixed-point mass <= top_p * Z, at least one: pipeline.py:105-121); the draw walks the kept
//      entries in INDEX order (second sort) -- the same integers as the general path, hence the same token.
// One pass over a row in global memory with NB 16-byte loads in flight per thread (raw vectors: 4 registers per 8
// sixteen-bit elements), visiting every element as f(index, value) -- per thread in ASCENDING index order.  A CTA owns a
// whole SM here, so memory-level parallelism is per thread: with the two 8-float batches of load8 a 256 KB row costs
// eight dependent HBM round trips (~25 us per pass); with 8 vectors in flight it is two.
template <int NB, typename T, typename F>
SD_DEVINL void scan_row(const GlobalRow<T>& G, F&& f) {
    const int tid = threadIdx.x, n = G.n;
    if constexpr (sizeof(T) == 2) {
        if ((reinterpret_cast<uintptr_t>(G.row) & 15u) == 0) {
            const int nvec = n >> 3;
            const uint4* vp = reinterpret_cast<const uint4*>(G.row);
            for (int v0 = 0; v0 < nvec; v0 += kSampThreads * NB) {
                uint4 v[NB];
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int vi = v0 + q * kSampThreads + tid;
                    v[q] = vi < nvec ? ldg_stream_v4(vp + vi) : make_uint4(0u, 0u, 0u, 0u);
                }
#pragma unroll
                for (int q = 0; q < NB; ++q) pin_loaded(v[q]);
#pragma unroll
                for (int q = 0; q < NB; ++q) {
                    const int vi = v0 + q * kSampThreads + tid;
                    if (vi < nvec) {
                        const uint32_t w[4] = {v[q].x, v[q].y, v[q].z, v[q].w};
#pragma unroll
                        for (int h = 0; h < 4; ++h) {
                            if constexpr (std::is_same<T, __nv_bfloat16>::value) {
                                f(vi * 8 + 2 * h, __uint_as_float(w[h] << 16));
                                f(vi * 8 + 2 * h + 1, __uint_as_float(w[h] & 0xffff0000u));
                            } else {
                                const float2 x2 = __half22float2(*reinterpret_cast<const __half2*>(&w[h]));
                                f(vi * 8 + 2 * h, x2.x);
                                f(vi * 8 + 2 * h + 1, x2.y);
                            }
                        }
                    }
                }
            }
            const int t = nvec * 8 + tid;  // < 8 elements left
            if (t < n) f(t, to_float<T>(G.row[t]));
            return;
        }
    }
    for (int tb = 0; tb < n; tb += kTile * kBatch) {
        float xs[kBatch][8];
        bool ok[kBatch][8];
#pragma unroll
        for (int q = 0; q < kBatch; ++q) G.load8(tb + q * kTile + tid * 8, xs[q], ok[q]);
#pragma unroll
        for (int q = 0; q < kBatch; ++q) {
#pragma unroll
            for (int e = 0; e < 8; ++e)
                if (ok[q][e]) f(tb + q * kTile + tid * 8 + e, xs[q][e]);
        }
    }
}

// descending bitonic sort of the first `size` entries (a power of two <= 1024); called by all 1024 threads
SD_DEVINL void bitonic_sort_desc(unsigned long long* a, const int size) {
    const int i = threadIdx.x;
    for (int k = 2; k <= size; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            const int ixj = i ^ j;
            if (ixj > i && ixj < size) {
                const unsigned long long x = a[i], y = a[ixj];
                const bool desc = (i & k) == 0;
                if (desc ? (x < y) : (x > y)) {
                    a[i] = y;
                    a[ixj] = x;
                }
            }
            __syncthreads();
        }
    }
}

// Returns false (uniformly) when the candidate superset does not fit: nothing has been written, the caller continues
// with the general path.  `my_max` = this thread's maximum order key from the maximum pass (0 if it saw no element).
template <typename T>
SD_DEVINL bool topk_fast_path(const GlobalRow<T>& R, const SampShared& sh, const SampleParams& p, float m, float u, uint32_t my_max,
                              unsigned long long* buf /* 3 x 1024 u64 of shared memory */, uint32_t& token_out, unsigned& kept_out) {
    const int tid = threadIdx.x;
    const int k = p.top_k;
    unsigned long long* A = buf;             // sort scratch / candidate list
    unsigned long long* W = buf + 1024;      // masses of the sorted candidates
    unsigned long long* I = buf + 2048;      // index-order list of the kept candidates
    // 1. lower bound of the k-th largest key: the k-th largest per-thread maximum, found bit by bit (the largest t with
    //    at least k thread maxima >= t: 32 counting barriers instead of a 55-stage sort of 1024 keys)
    uint32_t tau0 = 0u;
    for (int bit = 31; bit >= 0; --bit) {
        const uint32_t cand = tau0 | (1u << bit);
        if (__syncthreads_count(my_max >= cand) >= k) tau0 = cand;
    }
    // 2. collect {key >= tau0}
    if (tid == 0) sh.s_u32[2] = 0u;
    __syncthreads();
    scan_row<8>(R, [&](int idx, float x) {
        const uint32_t key = order_key(x);
        if (key >= tau0) {
            const unsigned pos = atomicAdd(&sh.s_u32[2], 1u);
            if (pos < (unsigned)kCompactCap) W[pos] = ((unsigned long long)key << 32) | (uint32_t)(~(uint32_t)idx);
        }
    });
    __syncthreads();
    const unsigned c = sh.s_u32[2];
    __syncthreads();
    if (c > (unsigned)kCompactCap) return false;
    // 3. exact order: key descending, index ascending
    int size = 32;
    while (size < (int)c) size <<= 1;  // the sorts run on the next power of two, not on all 1024 slots
    A[tid] = tid < (int)c ? W[tid] : 0ull;
    __syncthreads();
    bitonic_sort_desc(A, size);
    const int nk = k < (int)c ? k : (int)c;  // (c >= k whenever k < V)
    const unsigned long long ent = A[tid];
    const uint32_t my_idx = ~(uint32_t)ent;
    const unsigned long long w = tid < nk ? fixed_mass(key_to_float((uint32_t)(ent >> 32)), m, p.inv_t_log2) : 0ull;
    int n_kept = nk;
    if (p.top_p < 1.0f) {
        ScanPair total;
        const ScanPair ex = block_exclusive_scan(ScanPair{w, 0u}, total, sh.s_mass, sh.s_cnt);
        // cumulative probability <= top_p  <=>  cumulative mass <= top_p * Z1 (same integers as nucleus_and_draw)
        const unsigned long long target = (unsigned long long)((double)p.top_p * (double)total.mass);
        const unsigned ok = (tid < nk && ex.mass + w <= target) ? 1u : 0u;
        ScanPair tot2;
        block_exclusive_scan(ScanPair{0ull, ok}, tot2, sh.s_mass, sh.s_cnt);
        n_kept = tot2.cnt > 0u ? (int)tot2.cnt : 1;  // "keep at least one token" (pipeline.py:115)
    }
    const bool kept = tid < n_kept;
    W[tid] = kept ? w : 0ull;
    I[tid] = kept ? (((unsigned long long)(uint32_t)(~my_idx)) << 32) | (unsigned)tid : 0ull;  // descending ~idx == ascending idx
    __syncthreads();
    bitonic_sort_desc(I, size);
    const unsigned long long ie = I[tid];
    const unsigned long long wj = tid < n_kept ? W[(unsigned)ie & 1023u] : 0ull;
    ScanPair totz;
    const ScanPair exz = block_exclusive_scan(ScanPair{wj, 0u}, totz, sh.s_mass, sh.s_cnt);
    u = u < 0.f ? 0.f : (u > 0.99999994f ? 0.99999994f : u);
    const unsigned long long target = (unsigned long long)((double)u * (double)totz.mass);
    if (tid == 0) sh.s_u32[0] = 0xffffffffu;
    __syncthreads();
    if (wj && exz.mass <= target && target < exz.mass + wj) sh.s_u32[0] = ~(uint32_t)(ie >> 32);
    __syncthreads();
    token_out = sh.s_u32[0];
    kept_out = (unsigned)n_kept;
    __syncthreads();
    return true;
}

}  // namespace sd
