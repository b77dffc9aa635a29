// Kernel 1: fused verify (argmax over V for B*(K+1) rows + accepted-prefix + bonus token).
//
// Replaces the reference's verify_prefix (src/kernels/reference.py:13-56; Triton incumbent
// src/kernels/triton/verify.py:13-98, one serial program per sequence) with ONE streaming pass:
//
//   grid  = B * rows * S CTAs; CTA (row, chunk) reduces a contiguous 16-byte-aligned slice of one
//           logits row with 128-bit ld.global.nc loads, all loads of a thread issued before use
//   thread: running packed max (max.NaN.{bf16x2,f16x2,f32}) + the index of the FIRST vector at
//           which its maximum appeared -- ~1 ALU op per element, so the kernel stays HBM-bound
//   CTA   : REDUX warp reduce of (order_key, min position) -> one partial per (row, chunk)
//   last CTA of a sequence (ticket counter) merges the S partials of its rows, resolves the exact
//           element inside the winning vector, and writes accept_len / mask / target ids / bonus /
//           packed record.  No second launch, no host round trip.
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
    uint32_t* tickets;    // [B]
    int B, rows, K, V, S, chunk_vecs;
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

template <typename T, int VPT, bool STATS>
__global__ void __launch_bounds__((VPT >= 8 || STATS) ? 512 : 1024) verify_kernel(const VerifyParams p) {
    using P = Packed<T>;
    constexpr int EPV = 16 / (int)sizeof(T);  // elements per 16-byte vector
    const int tid = threadIdx.x;
    const int nthreads = blockDim.x;
    const int unit = blockIdx.x;
    const int chunk = unit % p.S;
    const int row = unit / p.S;
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

    const int v0 = chunk * p.chunk_vecs;
    int v1 = v0 + p.chunk_vecs;
    if (v1 > nvec) v1 = nvec;
    const int n = v1 > v0 ? v1 - v0 : 0;
    const uint4* vp = reinterpret_cast<const uint4*>(base + head) + v0;

    uint32_t run = P::kNegInf;  // replicated scalar running max (packed)
    int first_it = tid < n ? tid : 0x7fffffff;  // vector index (within chunk) where `run` first appeared
    float sm = -INFINITY, ss = 0.f;  // STATS: reference max (log2 domain) and sum

    for (int i0 = 0; i0 < n; i0 += nthreads * VPT) {
        uint4 buf[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            int idx = i0 + u * nthreads + tid;
            if (idx < n) {
                buf[u] = ldg_stream_v4(vp + idx);
            } else {
                buf[u] = make_uint4(P::kNegInf, P::kNegInf, P::kNegInf, P::kNegInf);
            }
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            int idx = i0 + u * nthreads + tid;
            uint32_t t = P::vmax(P::vmax(buf[u].x, buf[u].y), P::vmax(buf[u].z, buf[u].w));
            t = P::fold(t);
            uint32_t nw = P::vmax(run, t);
            bool changed = nw != run;
            if (changed) first_it = idx;
            if constexpr (STATS) {
                if (idx < n) {
                    if (changed) {
                        float mn = P::scalar(nw) * p.scale_log2;
                        ss = (sm == -INFINITY) ? 0.f : ss * fast_exp2(sm - mn);
                        sm = mn;
                    }
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
    c.pos = first_it == 0x7fffffff ? 0x7fffffff : head + (v0 + first_it) * EPV;
    if (n == 0 || first_it == 0x7fffffff) c.key = 0u;  // thread saw nothing

    // scalar peel elements (head + tail): handled by chunk 0 only, one element per thread
    if (chunk == 0) {
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
        if (lane == 0) {
            const int slot = row * p.S + chunk;
            p.part_key[slot] = c2.key;
            p.part_pos[slot] = c2.pos;
            if constexpr (STATS) {
                p.part_m[slot] = m2;
                p.part_s[slot] = s2;
            }
            __threadfence();
            unsigned prev = atomicAdd(&p.tickets[b], 1u);
            s_last = (prev == (unsigned)(p.rows * p.S) - 1u);
        }
    }
    __syncthreads();
    if (!s_last) return;

    // ---- this CTA retired the last slice of sequence b: finalize the sequence ----
    __threadfence();
    __shared__ int s_tgt[64];
    __shared__ float s_ptok[64];
    for (int r = wid; r < p.rows; r += nwarps) {
        const int grow = b * p.rows + r;
        Cand best{0u, 0x7fffffff};
        float fm = -INFINITY, fs = 0.f;
        for (int s = lane; s < p.S; s += 32) {
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
        best = warp_best(best);
        if constexpr (STATS) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                float m2 = __shfl_xor_sync(0xffffffffu, fm, o);
                float s2 = __shfl_xor_sync(0xffffffffu, fs, o);
                merge_ms(fm, fs, m2, s2);
            }
        }
        // exact element inside the winning 16-byte vector (or the peel element itself)
        const T* rb = reinterpret_cast<const T*>(p.logits) + (int64_t)b * p.stride_b + (int64_t)r * p.stride_r;
        int e = best.pos + lane;
        bool hit = false;
        if (lane < EPV && best.pos != 0x7fffffff && e < V) hit = order_key(to_float<T>(rb[e])) == best.key;
        unsigned ballot = __ballot_sync(0xffffffffu, hit);
        int idx = ballot ? best.pos + (__ffs(ballot) - 1) : 0;
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
                    } else {
                        int tok = load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + r);
                        if (tok >= 0 && tok < V) pt = fast_exp2(to_float<T>(rb[tok]) * p.scale_log2 - fm) / fs;
                    }
                    if (p.token_prob) p.token_prob[b * p.K + r] = pt;
                }
                s_ptok[r] = pt;
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        int nacc = 0;
        if constexpr (STATS) {
            if (p.mode == SD_ACCEPT_GREEDY) {
                while (nacc < p.K && s_tgt[nacc] == load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + nacc)) ++nacc;
            } else {
                while (nacc < p.K && s_ptok[nacc] >= p.threshold) ++nacc;
            }
        } else {
            while (nacc < p.K && s_tgt[nacc] == load_id(p.ids, p.id_dtype, (int64_t)b * p.id_stride_b + nacc)) ++nacc;
        }
        p.accept_len[b] = nacc;
        if (p.mask)
            for (int k = 0; k < p.K; ++k) p.mask[b * p.K + k] = (uint8_t)(k < nacc);
        const int bon = p.rows > p.K ? s_tgt[nacc] : -1;
        if (p.bonus) p.bonus[b] = bon;
        if (p.packed) {
            int32_t* rec = p.packed + (int64_t)b * (p.K + 2);
            rec[0] = nacc;
            for (int r = 0; r <= p.K; ++r) rec[1 + r] = r < p.rows ? s_tgt[r] : -1;
        }
        p.tickets[b] = 0u;  // self-reset: the workspace is ready for the next launch / graph replay
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
static cudaError_t launch_t(const VerifyParams& p, int threads, int vpt, cudaStream_t st) {
    dim3 grid((unsigned)((size_t)p.B * p.rows * p.S));
    switch (vpt) {
        case 1: verify_kernel<T, 1, STATS><<<grid, threads, 0, st>>>(p); break;
        case 2: verify_kernel<T, 2, STATS><<<grid, threads, 0, st>>>(p); break;
        case 4: verify_kernel<T, 4, STATS><<<grid, threads, 0, st>>>(p); break;
        case 8: verify_kernel<T, 8, STATS><<<grid, threads, 0, st>>>(p); break;
        default: verify_kernel<T, 16, STATS><<<grid, threads, 0, st>>>(p); break;
    }
    return cudaGetLastError();
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
    int threads = g_tuning.threads, vpt = g_tuning.vpt;
    if (threads <= 0 || vpt <= 0) {
        // heuristic: ~2 waves of 256-thread CTAs with 8 vectors (128 B) in flight per thread
        threads = 256;
        vpt = 8;
        const long long total_vec = (long long)B * rows * nvec_max;
        const long long slots = (long long)sm_count() * 8;
        while (vpt > 1 && total_vec / ((long long)threads * vpt) < slots) vpt >>= 1;
        if (stats && vpt > 4) vpt = 4;
    }
    int chunk_vecs = threads * vpt;
    int S = (nvec_max + chunk_vecs - 1) / chunk_vecs;
    if (S > kMaxChunks) {  // very long rows: loop inside the CTA instead of adding partials
        S = kMaxChunks;
        chunk_vecs = (nvec_max + S - 1) / S;
    }
    if (S < 1) S = 1;

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
    p.S = S;
    p.chunk_vecs = chunk_vecs;
    p.mode = mode;
    p.threshold = threshold;
    p.inv_temperature = inv_temperature;
    p.scale_log2 = inv_temperature * 1.4426950408889634f;

    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    cudaError_t e;
    if (stats) {
        if (logit_dtype == SD_F32) e = launch_t<float, true>(p, threads, vpt, st);
        else if (logit_dtype == SD_F16) e = launch_t<__half, true>(p, threads, vpt, st);
        else e = launch_t<__nv_bfloat16, true>(p, threads, vpt, st);
    } else {
        if (logit_dtype == SD_F32) e = launch_t<float, false>(p, threads, vpt, st);
        else if (logit_dtype == SD_F16) e = launch_t<__half, false>(p, threads, vpt, st);
        else e = launch_t<__nv_bfloat16, false>(p, threads, vpt, st);
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
