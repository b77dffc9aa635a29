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
#include "sd_sample.cuh"

namespace sd {

static int g_sample_no_fast = 0;  // evidence runs / tests: 1 = always the general (radix-select) path

template <typename T>
__global__ void __launch_bounds__(kSampThreads) sample_bonus_kernel(const SampleParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
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
    sh.s_mass = s_mass;
    sh.s_cnt = s_cnt;
    sh.tot_cnt = tot_cnt;
    sh.tot_mass = tot_mass;
    sh.s_u32 = s_u32;
    sh.s_u64 = s_u64;

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b = blockIdx.x;
    int r = p.row_const;
    if (p.row_index) {
        r = p.row_index[b] + p.row_bias;
        r = r < 0 ? 0 : (r >= p.rows ? p.rows - 1 : r);
    }
    GlobalRow<T> G;
    G.row = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)r * p.stride_r;
    G.n = p.V;
    const int V = p.V;

    // ---- pass A: maximum (order key) and its first index ----
    uint32_t kmax = 0u;
    int amax = 0x7fffffff;
    scan_row<8>(G, [&](int idx, float x) {
        const uint32_t k = order_key(x);
        if (k > kmax) {  // strictly greater: indices ascend within a thread, the first one wins
            kmax = k;
            amax = idx;
        }
    });
    const uint32_t my_max = kmax;  // this thread's own maximum (the fast top-k path bounds the threshold with these)
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
    const float u = p.uniform[b];
    const bool want_count = p.kept_count != nullptr;

    SelectResult sel;
    sel.tau = 0u;
    sel.n_ties = 0xffffffffu;  // everything
    bool has_sel = false;
    uint32_t token = 0xffffffffu;
    unsigned kept = 0;
    if (p.top_k > 0 && p.top_k < V && p.top_k <= kCompactCap && !p.no_fast_path) {
        // two-pass path on <= 1024 candidates; the histogram scratch is free here and holds its three lists
        if (topk_fast_path(G, sh, p, m, u, my_max, reinterpret_cast<unsigned long long*>(smem_raw), token, kept)) {
            if (tid == 0) {
                p.out[b] = token == 0xffffffffu ? amax : (int)token;
                if (p.kept_count) p.kept_count[b] = (int)kept;
            }
            return;
        }
    }
    if (p.top_k > 0 && p.top_k < V) {
        sel = radix_select(G, sh, p, m, false, (unsigned)p.top_k, 0ull, sel, false);
        has_sel = true;
        if (p.top_k <= kCompactCap) {
            // compact the <= top_k survivors (index order) into shared memory and finish there: top-p and the
            // draw then touch <= 1024 elements instead of the whole row
            unsigned run = 0;
            for_each_kept(G, sh, sel, true, [&](int t0, const float (&xs)[8], const bool (&keep)[8]) {
                unsigned mine = 0;
#pragma unroll
                for (int e = 0; e < 8; ++e) mine += keep[e] ? 1u : 0u;
                ScanPair v{0ull, mine}, total;
                const ScanPair ex = block_exclusive_scan(v, total, sh.s_mass, sh.s_cnt);
                unsigned pos = run + ex.cnt;
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    if (keep[e] && pos < (unsigned)kCompactCap) {
                        cand_x[pos] = xs[e];
                        cand_idx[pos] = t0 + tid * 8 + e;
                        ++pos;
                    }
                }
                run += total.cnt;
                return true;
            });
            __syncthreads();
            CompactRow C;
            C.x = cand_x;
            C.idx = cand_idx;
            C.n = (int)(run < (unsigned)kCompactCap ? run : (unsigned)kCompactCap);
            SelectResult all;
            all.tau = 0u;
            all.n_ties = 0xffffffffu;
            nucleus_and_draw(C, sh, p, m, all, false, u, want_count, token, kept);
        } else {
            nucleus_and_draw(G, sh, p, m, sel, true, u, want_count, token, kept);
        }
    } else {
        nucleus_and_draw(G, sh, p, m, sel, has_sel, u, want_count, token, kept);
    }
    if (tid == 0) {
        // rounding of the fixed-point masses can leave the target at the very end: fall back to the argmax
        p.out[b] = token == 0xffffffffu ? amax : (int)token;
        if (p.kept_count) p.kept_count[b] = (int)kept;
    }
}

}  // namespace sd

extern "C" {

int sd_sample_set_fast_path(int enabled) {
    sd::g_sample_no_fast = enabled ? 0 : 1;
    return SD_OK;
}

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
    p.no_fast_path = sd::g_sample_no_fast;
    // digits that cover every information bit of the order key: fp32 32, fp16 19 (1+8+10 after widening), bf16 16
    p.ndigits = logit_dtype == SD_F32 ? 4 : (logit_dtype == SD_F16 ? 3 : 2);
    const size_t smem = (size_t)kSampWarps * kBins * (4 + 8) + (size_t)kCompactCap * 8;
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
