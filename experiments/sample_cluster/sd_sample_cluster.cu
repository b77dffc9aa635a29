// EXPERIMENT (not part of the product library, not measured yet): the bonus-token sampler with one ROW split over
// a thread-block CLUSTER instead of one CTA per row.
//
// Why: sd_sample_bonus (llm-inference-lab_b200/csrc/sd_sample.cu) runs one CTA per sequence and makes 4 passes
// over the 256 KB row through L2 (maximum, two radix-select digits, compaction) before it continues on the <= 1024
// top-k survivors in shared memory: 159 us at B=64 on a B200, latency-bound on 64 of 148 SMs.  Here a cluster of
// C CTAs owns a row; CTA r stages elements [r*slice, (r+1)*slice) into ITS shared memory once (one HBM/L2 read of
// the row in total) and every pass runs on shared memory; the per-pass results (maximum, digit histogram, tie /
// survivor counts) are combined across the cluster through distributed shared memory and one cluster barrier per
// pass.  The survivors are compacted, in index order, straight into the LEADER's shared memory, and the leader
// finishes with exactly the code of the product kernel (nucleus_and_draw on a CompactRow).
//
// Contract: identical outputs to sd_sample_bonus for every configuration: the (tau, n_ties) prefixes, the candidate
// list and its order are the same, and all masses are fixed-point integers (sums independent of their order), so
// the token is bit-identical.  With 0 < top_k <= 1024 the leader finishes alone on the compacted candidates; with
// no top-k or a larger one ("full-row path") the nucleus select, the kept mass and the draw also run across the
// cluster, on the slices in shared memory.
//
// Build + check on a GPU box: experiments/sample_cluster/run_gpu.py.
#include <cooperative_groups.h>

#include "../../llm-inference-lab_b200/csrc/sd_sample.cuh"

namespace cg = cooperative_groups;

namespace sd {

constexpr int kMaxCluster = 8;
constexpr unsigned kAllTies = 0xffffffffu;
constexpr size_t kScratchBytes = (size_t)kSampWarps * kBins * (4 + 8) + (size_t)kCompactCap * 8;

// What a CTA publishes for its peers (read through distributed shared memory after a cluster barrier).
struct ClusterPub {
    unsigned cnt[2][kBins];             // digit histogram of the slice, double-buffered over the radix passes
    unsigned long long mass[2][kBins];  // ... and its fixed-point masses (by-mass select)
    uint32_t max_key;
    int max_pos;
    unsigned ties, gt;                  // elements of the slice with key == tau / key > tau
    unsigned long long sum_mass;        // scratch of cluster_sum
    unsigned sum_cnt;
};

// The slice of the row held in this CTA's shared memory, in the input type.
template <typename T>
struct SliceRow {
    const T* x;
    int n;     // elements in the slice (0 for a CTA past the end of the row)
    int base;  // row index of element 0
    SD_DEVINL void load8(int i0, float (&xs)[8], bool (&ok)[8]) const {
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ok[e] = i0 + e < n;
            xs[e] = ok[e] ? to_float<T>(x[i0 + e]) : 0.f;
        }
    }
};

// ---- full-row path: the building blocks of sd_sample.cuh restated over (slice in shared memory) x (cluster) ----

// Sum of a (mass, count) pair over the cluster; `before` receives the sum over the lower ranks (index order).
// Two cluster barriers: publish -> read, and read -> the next publish of the same fields.
SD_DEVINL ScanPair cluster_sum(cg::cluster_group& cl, ClusterPub* pub, ScanPair mine, ScanPair& before, uint32_t* s_u32,
                               unsigned long long* s_u64) {
    const unsigned C = cl.num_blocks(), rank = cl.block_rank();
    if (threadIdx.x == 0) {
        pub->sum_mass = mine.mass;
        pub->sum_cnt = mine.cnt;
    }
    cl.sync();
    if (threadIdx.x == 0) {
        ScanPair tot{0ull, 0u}, bef{0ull, 0u};
        for (unsigned q = 0; q < C; ++q) {
            const ClusterPub* pq = cl.map_shared_rank(pub, q);
            const unsigned long long mq = pq->sum_mass;
            const unsigned cq = pq->sum_cnt;
            if (q < rank) {
                bef.mass += mq;
                bef.cnt += cq;
            }
            tot.mass += mq;
            tot.cnt += cq;
        }
        s_u64[0] = tot.mass;
        s_u64[1] = bef.mass;
        s_u32[0] = tot.cnt;
        s_u32[1] = bef.cnt;
    }
    __syncthreads();
    ScanPair tot{s_u64[0], s_u32[0]};
    before.mass = s_u64[1];
    before.cnt = s_u32[1];
    cl.sync();
    return tot;
}

// Number of elements of the slice with key == tau, summed over the LOWER ranks: the row-wide tie rank at which this
// slice starts (ties are kept in index order).  Only needed when the prefix keeps some, not all, of its ties.
template <typename T>
SD_DEVINL unsigned ties_before_slice(cg::cluster_group& cl, ClusterPub* pub, const SliceRow<T>& S, const SampShared& sh,
                                     const SelectResult& sel, bool restricted) {
    if (!restricted || sel.n_ties == kAllTies) return 0u;  // uniform over the cluster
    unsigned my = 0;
    for (int j = threadIdx.x; j < S.n; j += kSampThreads) my += order_key(to_float<T>(S.x[j])) == sel.tau ? 1u : 0u;
    ScanPair v{0ull, my}, total, before;
    block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
    cluster_sum(cl, pub, ScanPair{0ull, total.cnt}, before, sh.s_u32, sh.s_u64);
    return before.cnt;
}

// for_each_kept of sd_sample.cuh over the slice, tie ranks starting at `ties_before`.
template <typename T, typename Visit>
SD_DEVINL void for_each_kept_slice(const SliceRow<T>& S, const SampShared& sh, const SelectResult& s, bool restricted,
                                   unsigned ties_before, Visit&& visit) {
    const int tid = threadIdx.x;
    const bool rank_ties = restricted && s.n_ties != kAllTies;
    unsigned run_ties = ties_before;
    for (int t0 = 0; t0 < S.n; t0 += kTile) {
        float xs[8];
        bool ok[8];
        S.load8(t0 + tid * 8, xs, ok);
        bool keep[8];
        uint32_t ks[8];
        unsigned my_ties = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ks[e] = order_key(xs[e]);
            if (ok[e] && rank_ties && ks[e] == s.tau) ++my_ties;
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
            keep[e] = ok[e];
            if (ok[e] && restricted) {
                keep[e] = rank_ties ? in_prefix(ks[e], rank, s) : ks[e] >= s.tau;
                if (rank_ties && ks[e] == s.tau) ++rank;
            }
        }
        if (!visit(t0, xs, keep)) break;  // uniform across the CTA
    }
}

// Kept mass and kept count of the prefix: this slice's (returned through `mine`), the lower ranks' (`before`) and
// the cluster total (return value).
template <typename T>
SD_DEVINL ScanPair cluster_kept_mass(cg::cluster_group& cl, ClusterPub* pub, const SliceRow<T>& S, const SampShared& sh,
                                     const SampleParams& p, float m, const SelectResult& sel, bool restricted,
                                     unsigned ties_before, ScanPair& mine, ScanPair& before) {
    unsigned long long mass = 0ull;
    unsigned cnt = 0;
    for_each_kept_slice(S, sh, sel, restricted, ties_before, [&](int, const float (&xs)[8], const bool (&keep)[8]) {
#pragma unroll
        for (int e = 0; e < 8; ++e)
            if (keep[e]) {
                mass += fixed_mass(xs[e], m, p.inv_t_log2);
                ++cnt;
            }
        return true;
    });
    ScanPair v{mass, cnt};
    block_exclusive_scan(v, mine, sh.s_mass, sh.s_cnt);
    return cluster_sum(cl, pub, mine, before, sh.s_u32, sh.s_u64);
}

// radix_select of sd_sample.cuh with the per-digit histograms summed over the cluster.  `pass_no` counts the
// histogram passes of the whole kernel (double buffering of ClusterPub::cnt / mass).
template <typename T>
SD_DEVINL SelectResult cluster_radix_select(cg::cluster_group& cl, ClusterPub* pub, const SliceRow<T>& S,
                                            const SampShared& sh, const SampleParams& p, float m, bool by_mass,
                                            unsigned k_target, unsigned long long mass_target, const SelectResult outer,
                                            bool has_outer, unsigned& pass_no) {
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const unsigned C = cl.num_blocks();
    uint32_t prefix = 0u, prefix_mask = 0u;
    unsigned cnt_above = 0;
    unsigned long long mass_above = 0ull;
    for (int d = 0; d < p.ndigits; ++d) {
        const int shift = 24 - 8 * d;
        const unsigned buf = pass_no & 1u;
        ++pass_no;
        for (int i = tid; i < kSampWarps * kBins; i += kSampThreads) {
            sh.h_cnt[i] = 0u;
            sh.h_mass[i] = 0ull;
        }
        __syncthreads();
        for (int t0 = 0; t0 < S.n; t0 += kTile) {
            float xs[8];
            bool ok[8];
            S.load8(t0 + tid * 8, xs, ok);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                unsigned bin = 0xffffffffu;
                const uint32_t k = order_key(xs[e]);
                if (ok[e] && !(has_outer && k < outer.tau) && (k & prefix_mask) == prefix) bin = (k >> shift) & 0xffu;
                if (by_mass) {
                    if (bin != 0xffffffffu) {
                        atomicAdd(&sh.h_cnt[wid * kBins + bin], 1u);
                        atomicAdd(&sh.h_mass[wid * kBins + bin], fixed_mass(xs[e], m, p.inv_t_log2));
                    }
                } else {
                    const unsigned peers = __match_any_sync(0xffffffffu, bin);
                    if (bin != 0xffffffffu && lane == __ffs(peers) - 1)
                        atomicAdd(&sh.h_cnt[wid * kBins + bin], __popc(peers));
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
            pub->cnt[buf][tid] = c;
            pub->mass[buf][tid] = w;
        }
        cl.sync();
        if (tid < kBins) {
            unsigned c = 0;
            unsigned long long w = 0ull;
            for (unsigned q = 0; q < C; ++q) {
                const ClusterPub* pq = cl.map_shared_rank(pub, q);
                c += pq->cnt[buf][tid];
                w += pq->mass[buf][tid];
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
        if (tid == 0) {
            unsigned c = cnt_above;
            unsigned long long w = mass_above;
            int chosen = -1;
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
        if (chosen < 0) {  // uniform over the cluster: every CTA walked the same totals
            SelectResult all;
            all.tau = has_outer ? outer.tau : 0u;
            all.n_ties = has_outer ? outer.n_ties : kAllTies;
            return all;
        }
        prefix |= (uint32_t)chosen << shift;
        prefix_mask |= 0xffu << shift;
    }
    SelectResult res;
    const uint32_t low_mask = ~prefix_mask;
    res.tau = (prefix & 0x80000000u) ? prefix : (prefix | low_mask);
    if (tid == 0) {
        unsigned t = sh.tot_cnt[(prefix >> (24 - 8 * (p.ndigits - 1))) & 0xffu];
        if (has_outer && res.tau == outer.tau && t > outer.n_ties) t = outer.n_ties;
        unsigned keep;
        if (by_mass) {
            const unsigned long long each = fixed_mass(key_to_float(res.tau), m, p.inv_t_log2);
            const unsigned long long room = mass_target > mass_above ? mass_target - mass_above : 0ull;
            const unsigned long long fit = each ? room / each : (unsigned long long)t;
            keep = fit > t ? t : (unsigned)fit;
            if (cnt_above == 0 && keep == 0) keep = 1;
        } else {
            keep = k_target > cnt_above ? k_target - cnt_above : 0u;
            if (keep > t) keep = t;
        }
        sh.s_u32[3] = keep == t ? kAllTies : keep;
    }
    __syncthreads();
    res.n_ties = sh.s_u32[3];
    __syncthreads();
    return res;
}

// nucleus_and_draw of sd_sample.cuh across the cluster.  The CTA whose slice holds the target mass writes the token.
template <typename T>
SD_DEVINL void cluster_nucleus_and_draw(cg::cluster_group& cl, ClusterPub* pub, const SliceRow<T>& S, const SampShared& sh,
                                        const SampleParams& p, float m, SelectResult sel, bool has_sel, float u, int b,
                                        int amax, unsigned& pass_no) {
    const int tid = threadIdx.x;
    const unsigned C = cl.num_blocks(), rank = cl.block_rank();
    ScanPair mine, before;
    unsigned tb = ties_before_slice(cl, pub, S, sh, sel, has_sel);
    if (p.top_p < 1.0f) {
        const ScanPair z1 = cluster_kept_mass(cl, pub, S, sh, p, m, sel, has_sel, tb, mine, before);
        const unsigned long long target = (unsigned long long)((double)p.top_p * (double)z1.mass);
        SelectResult nucleus = cluster_radix_select(cl, pub, S, sh, p, m, true, 0u, target, sel, has_sel, pass_no);
        if (has_sel && nucleus.tau == sel.tau && nucleus.n_ties > sel.n_ties) nucleus.n_ties = sel.n_ties;
        sel = nucleus;
        has_sel = true;
        tb = ties_before_slice(cl, pub, S, sh, sel, has_sel);
    }
    const ScanPair z = cluster_kept_mass(cl, pub, S, sh, p, m, sel, has_sel, tb, mine, before);
    u = u < 0.f ? 0.f : (u > 0.99999994f ? 0.99999994f : u);
    const unsigned long long target = (unsigned long long)((double)u * (double)z.mass);
    // the slice that holds the target (the last rank also takes a target that rounding left past the end)
    const bool owner = before.mass <= target && (target < before.mass + mine.mass || rank == C - 1);
    if (owner) {  // uniform within the CTA
        if (tid == 0) sh.s_u32[0] = 0xffffffffu;
        __syncthreads();
        unsigned long long run_mass = before.mass;
        for_each_kept_slice(S, sh, sel, has_sel, tb, [&](int t0, const float (&xs)[8], const bool (&keep)[8]) {
            unsigned long long w[8];
            unsigned long long my = 0ull;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                w[e] = keep[e] ? fixed_mass(xs[e], m, p.inv_t_log2) : 0ull;
                my += w[e];
            }
            ScanPair v{my, 0u}, total;
            const ScanPair ex = block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
            const unsigned long long bef = run_mass + ex.mass;
            if (bef <= target && target < bef + my) {  // exactly one thread owns the target
                unsigned long long c = bef;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (w[e] && target < c + w[e]) {
                        sh.s_u32[0] = (uint32_t)(S.base + t0 + tid * 8 + e);
                        break;
                    }
                    c += w[e];
                }
            }
            run_mass += total.mass;
            __syncthreads();
            return sh.s_u32[0] == 0xffffffffu;
        });
        __syncthreads();
        if (tid == 0) {
            const uint32_t token = sh.s_u32[0];
            p.out[b] = token == 0xffffffffu ? amax : (int)token;
            if (p.kept_count) p.kept_count[b] = (int)z.cnt;
        }
    }
    cl.sync();  // no CTA leaves while a peer may still read its ClusterPub
}

template <typename T>
__global__ void __launch_bounds__(kSampThreads, 1) sample_cluster_kernel(const SampleParams p, const int slice) {
    cg::cluster_group cl = cg::this_cluster();
    const unsigned C = cl.num_blocks();
    const unsigned rank = cl.block_rank();
    const int b = (int)(blockIdx.x / C);

    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ ClusterPub pub;
    __shared__ unsigned long long s_mass[33];
    __shared__ unsigned s_cnt[33];
    __shared__ unsigned tot_cnt[kBins];
    __shared__ unsigned long long tot_mass[kBins];
    __shared__ uint32_t s_u32[4];
    __shared__ unsigned long long s_u64[2];
    __shared__ uint32_t s_rkey[kSampWarps];
    __shared__ int s_rpos[kSampWarps];
    SampShared sh;
    sh.h_cnt = reinterpret_cast<unsigned*>(smem_raw);
    sh.h_mass = reinterpret_cast<unsigned long long*>(smem_raw + kSampWarps * kBins * 4);
    float* cand_x = reinterpret_cast<float*>(smem_raw + kSampWarps * kBins * 12);
    int* cand_idx = reinterpret_cast<int*>(cand_x + kCompactCap);
    T* sx = reinterpret_cast<T*>(smem_raw + kScratchBytes);
    sh.s_mass = s_mass;
    sh.s_cnt = s_cnt;
    sh.tot_cnt = tot_cnt;
    sh.tot_mass = tot_mass;
    sh.s_u32 = s_u32;
    sh.s_u64 = s_u64;

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    int r = p.row_const;
    if (p.row_index) {
        r = p.row_index[b] + p.row_bias;
        r = r < 0 ? 0 : (r >= p.rows ? p.rows - 1 : r);
    }
    const int V = p.V;
    const int lo = (int)rank * slice;
    const int n = lo >= V ? 0 : (V - lo < slice ? V - lo : slice);
    const T* grow = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)r * p.stride_r + lo;

    // ---- stage the slice (the only global read of the row) ----
    {
        int done = 0;
        if (sizeof(T) == 2 && (reinterpret_cast<uintptr_t>(grow) & 15u) == 0) {
            const int nvec = n >> 3;
            for (int i = tid; i < nvec; i += kSampThreads)
                reinterpret_cast<uint4*>(sx)[i] = ldg_stream_v4(reinterpret_cast<const uint4*>(grow) + i);
            done = nvec << 3;
        }
        for (int i = done + tid; i < n; i += kSampThreads) sx[i] = grow[i];
    }
    __syncthreads();
    SliceRow<T> S;
    S.x = sx;
    S.n = n;
    S.base = lo;

    // ---- pass A: maximum (order key) and its first index, over the cluster ----
    uint32_t kmax = 0u;
    int amax = 0x7fffffff;
    for (int j = tid; j < n; j += kSampThreads) {  // indices ascend within a thread: the first one wins
        const uint32_t k = order_key(to_float<T>(sx[j]));
        if (k > kmax) {
            kmax = k;
            amax = lo + j;
        }
    }
    {
        const uint32_t wk = __reduce_max_sync(0xffffffffu, kmax);
        const unsigned wp = __reduce_min_sync(0xffffffffu, kmax == wk ? (unsigned)amax : 0x7fffffffu);
        if (lane == 0) {
            s_rkey[wid] = wk;
            s_rpos[wid] = (int)wp;
        }
        __syncthreads();
        if (wid == 0) {
            const uint32_t k2 = lane < kSampWarps ? s_rkey[lane] : 0u;
            const unsigned p2 = lane < kSampWarps ? (unsigned)s_rpos[lane] : 0x7fffffffu;
            const uint32_t bk = __reduce_max_sync(0xffffffffu, k2);
            const unsigned bp = __reduce_min_sync(0xffffffffu, k2 == bk ? p2 : 0x7fffffffu);
            if (lane == 0) {
                pub.max_key = bk;
                pub.max_pos = (int)bp;
            }
        }
    }
    cl.sync();
    if (wid == 0) {
        uint32_t k2 = 0u;
        unsigned p2 = 0x7fffffffu;
        if ((unsigned)lane < C) {
            const ClusterPub* q = cl.map_shared_rank(&pub, (unsigned)lane);
            k2 = q->max_key;
            p2 = (unsigned)q->max_pos;
        }
        const uint32_t bk = __reduce_max_sync(0xffffffffu, k2);
        const unsigned bp = __reduce_min_sync(0xffffffffu, k2 == bk ? p2 : 0x7fffffffu);
        if (lane == 0) {
            s_u32[0] = bk;
            s_u32[1] = bp;
        }
    }
    __syncthreads();
    kmax = s_u32[0];
    amax = (int)s_u32[1];
    __syncthreads();
    if (!p.uniform) {  // greedy branch
        if (rank == 0 && tid == 0) {
            p.out[b] = amax;
            if (p.kept_count) p.kept_count[b] = 1;
        }
        cl.sync();  // no CTA leaves while a peer may still read its ClusterPub
        return;
    }
    const float m = key_to_float(kmax);
    const unsigned k_target = (unsigned)p.top_k;
    unsigned pass_no = 0;
    if (!(p.top_k > 0 && p.top_k < V && p.top_k <= kCompactCap)) {
        // ---- full-row path: no top-k, or more survivors than the compacted list holds ----
        SelectResult sel;
        sel.tau = 0u;
        sel.n_ties = kAllTies;
        bool has_sel = false;
        if (p.top_k > 0 && p.top_k < V) {
            sel = cluster_radix_select(cl, &pub, S, sh, p, m, false, k_target, 0ull, sel, false, pass_no);
            has_sel = true;
        }
        cluster_nucleus_and_draw(cl, &pub, S, sh, p, m, sel, has_sel, p.uniform[b], b, amax, pass_no);
        return;
    }

    // ---- top-k: radix select by COUNT over the cluster (sd_sample.cuh radix_select, by_mass = false, no outer) ----
    uint32_t prefix = 0u, prefix_mask = 0u;
    unsigned cnt_above = 0;
    bool everything = false;
    int last_bin = 0;
    for (int d = 0; d < p.ndigits; ++d) {
        const int shift = 24 - 8 * d;
        const int buf = d & 1;
        for (int i = tid; i < kSampWarps * kBins; i += kSampThreads) sh.h_cnt[i] = 0u;
        __syncthreads();
        for (int t0 = 0; t0 < n; t0 += kTile) {  // uniform trip count across the CTA (match_any needs whole warps)
            float xs[8];
            bool ok[8];
            S.load8(t0 + tid * 8, xs, ok);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                unsigned bin = 0xffffffffu;
                const uint32_t k = order_key(xs[e]);
                if (ok[e] && (k & prefix_mask) == prefix) bin = (k >> shift) & 0xffu;
                const unsigned peers = __match_any_sync(0xffffffffu, bin);
                if (bin != 0xffffffffu && lane == __ffs(peers) - 1) atomicAdd(&sh.h_cnt[wid * kBins + bin], __popc(peers));
            }
        }
        __syncthreads();
        if (tid < kBins) {
            unsigned c = 0;
            for (int q = 0; q < kSampWarps; ++q) c += sh.h_cnt[q * kBins + tid];
            pub.cnt[buf][tid] = c;
        }
        // Double buffering makes ONE cluster barrier per pass enough: buffer `buf` is rewritten two passes later,
        // and a peer can only be there after the barrier of the pass in between, i.e. after it read this one.
        cl.sync();
        if (tid < kBins) {
            unsigned c = 0;
            for (unsigned q = 0; q < C; ++q) c += cl.map_shared_rank(&pub, q)->cnt[buf][tid];
            tot_cnt[tid] = c;
        }
        __syncthreads();
        if (tid == 0) {  // walk the 256 buckets from the top (every CTA does the same walk on the same totals)
            unsigned c = cnt_above;
            int chosen = -1;
            for (int bin = kBins - 1; bin >= 0; --bin) {
                const unsigned c2 = c + tot_cnt[bin];
                if (tot_cnt[bin] && c2 > k_target) {
                    chosen = bin;
                    break;
                }
                c = c2;
            }
            s_u32[0] = (uint32_t)chosen;
            s_u32[1] = c;
        }
        __syncthreads();
        const int chosen = (int)s_u32[0];
        cnt_above = s_u32[1];
        __syncthreads();
        if (chosen < 0) {  // fewer than top_k candidates: everything is kept (uniform over the cluster)
            everything = true;
            break;
        }
        last_bin = chosen;
        prefix |= (uint32_t)chosen << shift;
        prefix_mask |= 0xffu << shift;
    }
    SelectResult sel;
    if (everything) {
        sel.tau = 0u;
        sel.n_ties = kAllTies;
    } else {
        const uint32_t low_mask = ~prefix_mask;
        sel.tau = (prefix & 0x80000000u) ? prefix : (prefix | low_mask);
        const unsigned t = tot_cnt[last_bin];  // the last pass's bucket is ONE value: its count = number of ties
        unsigned keep = k_target > cnt_above ? k_target - cnt_above : 0u;
        if (keep > t) keep = t;
        sel.n_ties = keep == t ? kAllTies : keep;
    }

    // ---- survivors of this slice: counts, then their position in the index-ordered candidate list ----
    {
        unsigned my_gt = 0, my_ties = 0;
        for (int j = tid; j < n; j += kSampThreads) {
            const uint32_t k = order_key(to_float<T>(sx[j]));
            my_gt += k > sel.tau ? 1u : 0u;
            my_ties += k == sel.tau ? 1u : 0u;
        }
        ScanPair v{(unsigned long long)my_gt, my_ties}, total;
        block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
        if (tid == 0) {
            pub.gt = (unsigned)total.mass;
            pub.ties = total.cnt;
        }
    }
    cl.sync();
    if (tid == 0) {
        unsigned ties_run = 0, off = 0, ties_before = 0, total_kept = 0;
        for (unsigned q = 0; q < C; ++q) {
            const ClusterPub* pq = cl.map_shared_rank(&pub, q);
            const unsigned tq = pq->ties, gq = pq->gt;
            const unsigned room = sel.n_ties > ties_run ? sel.n_ties - ties_run : 0u;
            const unsigned kept_q = gq + (tq < room ? tq : room);
            if (q < rank) {
                off += kept_q;
                ties_before += tq;
            }
            total_kept += kept_q;
            ties_run += tq;
        }
        s_u32[0] = off;
        s_u32[1] = ties_before;
        s_u32[2] = total_kept;
    }
    __syncthreads();
    unsigned run_pos = s_u32[0];
    unsigned run_ties = s_u32[1];
    const unsigned total_kept = s_u32[2];
    __syncthreads();

    // ---- compaction in index order, straight into the leader's shared memory ----
    float* lead_x = cl.map_shared_rank(cand_x, 0u);
    int* lead_idx = cl.map_shared_rank(cand_idx, 0u);
    for (int t0 = 0; t0 < n; t0 += kTile) {
        float xs[8];
        bool ok[8];
        uint32_t ks[8];
        S.load8(t0 + tid * 8, xs, ok);
        unsigned my_gt = 0, my_ties = 0;
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            ks[e] = order_key(xs[e]);
            if (ok[e]) {
                my_gt += ks[e] > sel.tau ? 1u : 0u;
                my_ties += ks[e] == sel.tau ? 1u : 0u;
            }
        }
        ScanPair v{(unsigned long long)my_gt, my_ties}, total;
        const ScanPair ex = block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
        // ties are kept while their rank over the whole row (index order) is below n_ties
        const unsigned room = sel.n_ties > run_ties ? sel.n_ties - run_ties : 0u;
        unsigned tie_rank = run_ties + ex.cnt;
        unsigned pos = run_pos + (unsigned)ex.mass + (ex.cnt < room ? ex.cnt : room);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            if (!ok[e]) continue;
            bool keep = ks[e] > sel.tau;
            if (ks[e] == sel.tau) {
                keep = tie_rank < sel.n_ties;
                ++tie_rank;
            }
            if (keep && pos < (unsigned)kCompactCap) {
                lead_x[pos] = xs[e];
                lead_idx[pos] = lo + t0 + tid * 8 + e;
                ++pos;
            }
        }
        run_pos += (unsigned)total.mass + (total.cnt < room ? total.cnt : room);
        run_ties += total.cnt;
    }
    cl.sync();  // the candidate list is complete in the leader's shared memory; peers may leave
    if (rank != 0) return;

    // ---- leader: top-p + draw on the compacted candidates, the product kernel's code ----
    CompactRow Cr;
    Cr.x = cand_x;
    Cr.idx = cand_idx;
    Cr.n = (int)(total_kept < (unsigned)kCompactCap ? total_kept : (unsigned)kCompactCap);
    SelectResult all;
    all.tau = 0u;
    all.n_ties = kAllTies;
    uint32_t token = 0xffffffffu;
    unsigned kept = 0;
    nucleus_and_draw(Cr, sh, p, m, all, false, p.uniform[b], p.kept_count != nullptr, token, kept);
    if (tid == 0) {
        p.out[b] = token == 0xffffffffu ? amax : (int)token;
        if (p.kept_count) p.kept_count[b] = (int)kept;
    }
}

template <typename T>
static cudaError_t launch_cluster(const SampleParams& p, int B, int cluster, int slice, cudaStream_t st) {
    const size_t smem = kScratchBytes + (((size_t)slice * sizeof(T) + 15) & ~(size_t)15);
    cudaError_t e = cudaFuncSetAttribute(sample_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(B * cluster));
    cfg.blockDim = dim3(kSampThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, sample_cluster_kernel<T>, p, slice);
}

}  // namespace sd

extern "C" {

// Same arguments as sd_sample_bonus (include/sd_verify.h) + the cluster size (1, 2, 4 or 8).
int sd_sample_bonus_cluster(const void* logits, int logit_dtype, int64_t stride_b, int64_t stride_r,
                            const int32_t* row_index, int row_bias, int row_const, const float* uniform,
                            float temperature, int top_k, float top_p, int32_t* out_tokens, int32_t* kept_count, int B,
                            int rows, int V, int cluster, void* stream) {
    using namespace sd;
    if (B < 0 || rows <= 0 || V <= 0) return SD_ERR_INVALID_ARG;
    if (B == 0) return SD_OK;
    if (!logits || !out_tokens) return SD_ERR_INVALID_ARG;
    if (logit_dtype < SD_F32 || logit_dtype > SD_BF16) return SD_ERR_INVALID_ARG;
    if (uniform && !(temperature > 0.f)) return SD_ERR_INVALID_ARG;
    if (!row_index && (row_const < 0 || row_const >= rows)) return SD_ERR_SHAPE;
    if (cluster != 1 && cluster != 2 && cluster != 4 && cluster != kMaxCluster) return SD_ERR_INVALID_ARG;
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
    p.ndigits = logit_dtype == SD_F32 ? 4 : (logit_dtype == SD_F16 ? 3 : 2);
    const int per = (V + cluster - 1) / cluster;
    const int slice = (per + 7) & ~7;  // multiple of 8 elements: 16-byte aligned slices for 2-byte types
    const size_t elt = logit_dtype == SD_F32 ? 4 : 2;
    // the slice must fit in shared memory next to the scratch and ~10 KB of static arrays (ClusterPub, scan scratch)
    if (kScratchBytes + (size_t)slice * elt + 10240u > 227u * 1024u) return SD_ERR_UNSUPPORTED;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (logit_dtype == SD_F32) e = launch_cluster<float>(p, B, cluster, slice, st);
    else if (logit_dtype == SD_F16) e = launch_cluster<__half>(p, B, cluster, slice, st);
    else e = launch_cluster<__nv_bfloat16>(p, B, cluster, slice, st);
    return to_status(e);
}

}  // extern "C"
