// Kernel 3: (K+1)-query causal decode attention for the verify pass (GQA aware, split-KV).
//
// Replaces K sequential single-query SDPA decode passes of the reference's verify loop
// (src/scheduler/speculative_scheduler.py:192-199; attention selected at
// src/specdec/models/hf_wrappers.py:98, computed inside `transformers` by
// torch.nn.functional.scaled_dot_product_attention) -- each of which re-reads the whole KV cache --
// with ONE pass: the R = K+1 new positions of all G = Hq/Hkv query heads that share a KV head form an
// M = G*R (<= 16*MT) row tile, so every K/V byte is fetched from HBM once per step.
//
// Roofline: 2*B*Hkv*(L+R)*D*elt bytes of KV against 4*B*Hq*R*(L+R)*D flops = 2*G*R flops/byte/elt
// (15 flop/B for Llama-3.2-3B) -- HBM-bound, but above what fp32 FMA lanes sustain at 6+ TB/s, so the
// two contractions run on the tensor cores through warp-level mma.sync.m16n8k16 (M = 16 is the whole
// query tile; a tcgen05 tile would be >= 64 rows, i.e. >= 75 % padding, for no bandwidth gain).
//
// Structure (no CTA-wide barrier in the main loop):
//   grid  = (n_splits, Hkv * head_groups, B); CTA = W warps
//   warp  : owns every W-th KT-key tile of the CTA's KV range and streams it through its PRIVATE
//           STAGES-deep cp.async ring in shared memory (16-byte chunks, XOR-swizzled so ldmatrix is
//           bank-conflict free); S = Q K^T and O += P V with fp32 accumulators and an online softmax
//           in the log2 domain; the causal mask (query i sees rows <= seq_len + i) is applied only to
//           tiles that reach past seq_len
//   CTA   : merges the W warps' (max, sum, O) in shared memory -> out (n_splits == 1) or an fp32
//           partial in the workspace, merged by attn_combine_kernel
#include "sd_common.cuh"

namespace sd {

struct AttnParams {
    const void* q;
    const void* k;
    const void* v;
    const int32_t* seq_lens;
    void* out;
    float* part_o;   // [rows_total, n_splits, D]
    float* part_ml;  // [rows_total, n_splits, 2]
    int64_t kv_stride_b, kv_stride_h;
    int capacity;          // rows per (b, kv head) of the arena: seq_lens[b] + R is clamped to it
    int32_t* overflow;     // nullable: set when a length had to be clamped (or was negative)
    int B, Hq, Hkv, R, G;  // G = Hq / Hkv
    int gsub, ngroups;     // query heads per CTA, CTAs per KV head
    int n_splits;
    float scale_log2;
};

struct AttnTuning {
    int config;  // 0 heuristic, 1 = KT16/3 stages, 2 = KT32/3 stages, 3 = KT32/2 stages
};
static AttnTuning g_attn_tuning = {0};

SD_DEVINL uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SD_DEVINL void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
SD_DEVINL void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
SD_DEVINL void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

SD_DEVINL void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}
SD_DEVINL void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr));
}

template <typename T>
SD_DEVINL void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1);
template <>
SD_DEVINL void mma16816<__nv_bfloat16>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
template <>
SD_DEVINL void mma16816<__half>(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

template <typename T>
SD_DEVINL uint32_t pack2(float lo, float hi);
template <>
SD_DEVINL uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}
template <>
SD_DEVINL uint32_t pack2<__half>(float lo, float hi) {
    __half2 v = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&v);
}

// byte offset of 16-byte chunk `chunk` of row `row` in a [rows][D] 16-bit tile, XOR swizzled
template <int D>
SD_DEVINL uint32_t swz(int row, int chunk) {
    return (uint32_t)(row * (D * 2) + ((chunk ^ (row & 7)) << 4));
}

template <int D, int MT, int KT, int STAGES, int W>
struct AttnSmem {
    static constexpr int kQBytes = MT * 16 * D * 2;
    static constexpr int kTileBytes = KT * D * 2;              // one K or V tile
    static constexpr int kStageBytes = 2 * kTileBytes;         // K + V
    static constexpr int kRingBytes = W * STAGES * kStageBytes;
    static constexpr int kMergeStride = D + 8;                 // floats per merged row (bank spread)
    static constexpr int kMergeBytes = W * MT * 16 * kMergeStride * 4 + W * MT * 16 * 2 * 4;
    static constexpr int kBodyBytes = kRingBytes > kMergeBytes ? kRingBytes : kMergeBytes;
    static constexpr int kTotal = kQBytes + kBodyBytes;
};

template <typename T, int D, int MT, int KT, int STAGES, int W>
__global__ void __launch_bounds__(W * 32) verify_attn_kernel(const AttnParams p) {
    using SM = AttnSmem<D, MT, KT, STAGES, W>;
    constexpr int CPR = D / 8;        // 16-byte chunks per row
    constexpr int NKB = KT / 16;      // 16-key blocks per tile
    constexpr int NKS = D / 16;       // k-steps of the QK^T contraction
    constexpr int NDT = D / 8;        // 8-wide n-tiles of O
    constexpr bool QREG = (MT * D <= 128);
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* q_s = smem;
    unsigned char* body = smem + SM::kQBytes;

    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int mi = lane >> 3, r8 = lane & 7;  // ldmatrix: matrix index / row within it
    const int split = blockIdx.x;
    const int kvh = blockIdx.y / p.ngroups, gs = blockIdx.y - kvh * p.ngroups;
    const int b = blockIdx.z;
    const int g0 = gs * p.gsub;
    const int gcount = (p.G - g0) < p.gsub ? (p.G - g0) : p.gsub;
    const int mrows = gcount * p.R;
    const int hq0 = kvh * p.G + g0;
    // The append kernels clamp to the capacity and raise `overflow`, so a sequence may legitimately sit AT capacity:
    // never read rows beyond the (b, head) slab -- the next head's keys would give silently wrong output, and the last
    // slab ends the allocation.  A clamped (or negative) length raises the flag; its output is not meaningful.
    int seq = p.seq_lens[b];
    if (seq < 0 || seq > p.capacity - p.R) {
        if (p.overflow && threadIdx.x == 0) *p.overflow = 1;
        seq = seq < 0 ? 0 : p.capacity - p.R;
    }
    const int Ltot = seq + p.R;
    const int tiles_total = (Ltot + KT - 1) / KT;
    const int tiles_per_split = (tiles_total + p.n_splits - 1) / p.n_splits;
    const int t0 = split * tiles_per_split;
    const int t1 = (t0 + tiles_per_split) < tiles_total ? (t0 + tiles_per_split) : tiles_total;

    // ---- Q rows of this (b, kv head, head group): contiguous [mrows, D] block ----
    const int64_t qrow0 = ((int64_t)b * p.Hq + hq0) * p.R;
    {
        const char* qg = reinterpret_cast<const char*>(p.q) + qrow0 * (D * 2);
        for (int c = tid; c < MT * 16 * CPR; c += W * 32) {
            const int row = c / CPR, col = c - row * CPR;
            const bool valid = row < mrows;
            cp_async16(smem_addr(q_s) + swz<D>(row, col), qg + (valid ? ((int64_t)row * (D * 2) + col * 16) : 0), valid ? 16 : 0);
        }
        cp_async_commit();
    }

    // ---- this warp's tile sequence and its private ring ----
    const char* kg = reinterpret_cast<const char*>(p.k) + (int64_t)b * p.kv_stride_b + (int64_t)kvh * p.kv_stride_h;
    const char* vg = reinterpret_cast<const char*>(p.v) + (int64_t)b * p.kv_stride_b + (int64_t)kvh * p.kv_stride_h;
    unsigned char* ring = body + warp * (STAGES * SM::kStageBytes);
    const uint32_t ring_a = smem_addr(ring);
    const int ntw = (t1 - t0 - warp + W - 1) / W > 0 ? (t1 - t0 - warp + W - 1) / W : 0;  // tiles of this warp

    auto load_tile = [&](int j) {  // j-th tile of this warp -> stage j % STAGES
        const int t = t0 + warp + j * W;
        const int key0 = t * KT;
        const uint32_t sk = ring_a + (j % STAGES) * SM::kStageBytes;
        const uint32_t sv = sk + SM::kTileBytes;
#pragma unroll
        for (int c = lane; c < KT * CPR; c += 32) {
            const int row = c / CPR, col = c - row * CPR;
            const int key = key0 + row;
            const bool valid = key < Ltot;
            const int64_t off = (int64_t)(valid ? key : 0) * (D * 2) + col * 16;
            const uint32_t d = swz<D>(row, col);
            cp_async16(sk + d, kg + off, valid ? 16 : 0);
            cp_async16(sv + d, vg + off, valid ? 16 : 0);  // zero-filled: 0 * garbage must stay 0
        }
    };

#pragma unroll
    for (int j = 0; j < STAGES - 1; ++j) {
        if (j < ntw) load_tile(j);
        cp_async_commit();
    }
    cp_async_wait<STAGES - 1>();  // the Q group (oldest) has landed
    __syncthreads();

    uint32_t qf[QREG ? MT : 1][QREG ? NKS : 1][4];
    if constexpr (QREG) {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int ks = 0; ks < NKS; ++ks)
                ldmatrix_x4(qf[mt][ks], smem_addr(q_s) + swz<D>(mt * 16 + (mi & 1) * 8 + r8, 2 * ks + (mi >> 1)));
    }

    float o[MT][NDT][4];
    float mrow[MT][2], lrow[MT][2];
    int lim[MT][2];
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            mrow[mt][h] = -INFINITY;
            lrow[mt][h] = 0.f;
            const int m = mt * 16 + g + 8 * h;
            lim[mt][h] = seq + (m < mrows ? m % p.R : 0);
        }
#pragma unroll
        for (int nt = 0; nt < NDT; ++nt)
#pragma unroll
            for (int e = 0; e < 4; ++e) o[mt][nt][e] = 0.f;
    }

    for (int j = 0; j < ntw; ++j) {
        cp_async_wait<STAGES - 2>();  // tile j has landed (for this lane)
        __syncwarp();                 // ... for every lane; and every lane is done reading tile j-1
        if (j + STAGES - 1 < ntw) load_tile(j + STAGES - 1);
        cp_async_commit();

        const uint32_t sk = ring_a + (j % STAGES) * SM::kStageBytes;
        const uint32_t sv = sk + SM::kTileBytes;
        const int key0 = (t0 + warp + j * W) * KT;

        // ---- S = Q K^T ----
        float s[MT][2 * NKB][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int nt = 0; nt < 2 * NKB; ++nt)
#pragma unroll
                for (int e = 0; e < 4; ++e) s[mt][nt][e] = 0.f;
#pragma unroll
        for (int ks = 0; ks < NKS; ++ks) {
            uint32_t af[MT][4];
#pragma unroll
            for (int mt = 0; mt < MT; ++mt) {
                if constexpr (QREG) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) af[mt][e] = qf[mt][ks][e];
                } else {
                    ldmatrix_x4(af[mt], smem_addr(q_s) + swz<D>(mt * 16 + (mi & 1) * 8 + r8, 2 * ks + (mi >> 1)));
                }
            }
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                uint32_t kf[4];
                ldmatrix_x4(kf, sk + swz<D>(kb * 16 + (mi >> 1) * 8 + r8, 2 * ks + (mi & 1)));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816<T>(s[mt][2 * kb], af[mt], kf[0], kf[1]);
                    mma16816<T>(s[mt][2 * kb + 1], af[mt], kf[2], kf[3]);
                }
            }
        }

        // ---- online softmax (log2 domain), causal mask only where the tile reaches past seq ----
        const bool need_mask = key0 + KT > seq + 1;
        uint32_t pa[MT][NKB][4];
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float mx = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 2 * NKB; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        float v = s[mt][nt][2 * h + e] * p.scale_log2;
                        if (need_mask && key0 + nt * 8 + 2 * tig + e > lim[mt][h]) v = -INFINITY;
                        s[mt][nt][2 * h + e] = v;
                        mx = fmaxf(mx, v);
                    }
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(mrow[mt][h], mx);
                const float m_use = m_new == -INFINITY ? 0.f : m_new;
                const float alpha = fast_exp2(mrow[mt][h] - m_use);
                mrow[mt][h] = m_new;
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 2 * NKB; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float pv = fast_exp2(s[mt][nt][2 * h + e] - m_use);
                        s[mt][nt][2 * h + e] = pv;
                        sum += pv;
                    }
                lrow[mt][h] = lrow[mt][h] * alpha + sum;
#pragma unroll
                for (int nt = 0; nt < NDT; ++nt) {
                    o[mt][nt][2 * h] *= alpha;
                    o[mt][nt][2 * h + 1] *= alpha;
                }
            }
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                pa[mt][kb][0] = pack2<T>(s[mt][2 * kb][0], s[mt][2 * kb][1]);
                pa[mt][kb][1] = pack2<T>(s[mt][2 * kb][2], s[mt][2 * kb][3]);
                pa[mt][kb][2] = pack2<T>(s[mt][2 * kb + 1][0], s[mt][2 * kb + 1][1]);
                pa[mt][kb][3] = pack2<T>(s[mt][2 * kb + 1][2], s[mt][2 * kb + 1][3]);
            }
        }

        // ---- O += P V ----
#pragma unroll
        for (int kb = 0; kb < NKB; ++kb) {
#pragma unroll
            for (int ntp = 0; ntp < NDT / 2; ++ntp) {
                uint32_t vf[4];
                ldmatrix_x4_trans(vf, sv + swz<D>(kb * 16 + (mi & 1) * 8 + r8, 2 * ntp + (mi >> 1)));
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816<T>(o[mt][2 * ntp], pa[mt][kb], vf[0], vf[1]);
                    mma16816<T>(o[mt][2 * ntp + 1], pa[mt][kb], vf[2], vf[3]);
                }
            }
        }
    }
    cp_async_wait<0>();
    __syncthreads();  // every warp is done with its ring: the body is reused for the merge

    // ---- CTA merge of the W warps ----
    float* om = reinterpret_cast<float*>(body);                          // [W][MT*16][kMergeStride]
    float* mm = om + W * MT * 16 * SM::kMergeStride;                     // [W][MT*16]
    float* lm = mm + W * MT * 16;                                        // [W][MT*16]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float l = lrow[mt][h];
            l += __shfl_xor_sync(0xffffffffu, l, 1);
            l += __shfl_xor_sync(0xffffffffu, l, 2);
            const int row = mt * 16 + g + 8 * h;
            float* orow = om + ((size_t)warp * MT * 16 + row) * SM::kMergeStride;
#pragma unroll
            for (int nt = 0; nt < NDT; ++nt)
                *reinterpret_cast<float2*>(orow + nt * 8 + 2 * tig) = make_float2(o[mt][nt][2 * h], o[mt][nt][2 * h + 1]);
            if (tig == 0) {
                mm[warp * MT * 16 + row] = mrow[mt][h];
                lm[warp * MT * 16 + row] = l;
            }
        }
    }
    __syncthreads();

    for (int idx = tid; idx < mrows * (D / 4); idx += W * 32) {
        const int m = idx / (D / 4), c4 = idx - m * (D / 4);
        float mstar = -INFINITY;
#pragma unroll
        for (int w = 0; w < W; ++w) mstar = fmaxf(mstar, mm[w * MT * 16 + m]);
        const float m_use = mstar == -INFINITY ? 0.f : mstar;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const float sc = fast_exp2(mm[w * MT * 16 + m] - m_use);
            const float4 x = *reinterpret_cast<const float4*>(om + ((size_t)w * MT * 16 + m) * SM::kMergeStride + c4 * 4);
            acc.x += sc * x.x;
            acc.y += sc * x.y;
            acc.z += sc * x.z;
            acc.w += sc * x.w;
            l += sc * lm[w * MT * 16 + m];
        }
        const int64_t grow = qrow0 + m;
        if (p.n_splits == 1) {
            const float inv = l > 0.f ? 1.f / l : 0.f;
            uint2 pk;
            pk.x = pack2<T>(acc.x * inv, acc.y * inv);
            pk.y = pack2<T>(acc.z * inv, acc.w * inv);
            *reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.out) + (grow * D + c4 * 4) * 2) = pk;
        } else {
            const int64_t slot = grow * p.n_splits + split;
            *reinterpret_cast<float4*>(p.part_o + slot * D + c4 * 4) = acc;
            if (c4 == 0) {
                p.part_ml[slot * 2] = mstar;
                p.part_ml[slot * 2 + 1] = l;
            }
        }
    }
}

// merge of the split partials: one warp per output row
template <typename T, int D>
__global__ void __launch_bounds__(256) attn_combine_kernel(const AttnParams p, int64_t rows_total) {
    pdl_wait();
    pdl_trigger();
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows_total) return;
    float mstar = -INFINITY;
    for (int s = 0; s < p.n_splits; ++s) mstar = fmaxf(mstar, p.part_ml[(row * p.n_splits + s) * 2]);
    const float m_use = mstar == -INFINITY ? 0.f : mstar;
    for (int c4 = lane; c4 < D / 4; c4 += 32) {
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f;
        for (int s = 0; s < p.n_splits; ++s) {
            const int64_t slot = row * p.n_splits + s;
            const float sc = fast_exp2(p.part_ml[slot * 2] - m_use);
            const float4 x = *reinterpret_cast<const float4*>(p.part_o + slot * D + c4 * 4);
            acc.x += sc * x.x;
            acc.y += sc * x.y;
            acc.z += sc * x.z;
            acc.w += sc * x.w;
            l += sc * p.part_ml[slot * 2 + 1];
        }
        const float inv = l > 0.f ? 1.f / l : 0.f;
        uint2 pk;
        pk.x = pack2<T>(acc.x * inv, acc.y * inv);
        pk.y = pack2<T>(acc.z * inv, acc.w * inv);
        *reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.out) + (row * D + c4 * 4) * 2) = pk;
    }
}

// ------------------------------------------------------------------------------------------ host
constexpr int kAttnDefaultMaxSplits = 8;
constexpr int kAttnWarps = 4;

static int attn_sm_count() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = 148;
    }
    return n;
}

template <typename T, int D, int MT, int KT, int STAGES>
static cudaError_t launch_attn_cfg(const AttnParams& p, dim3 grid, cudaStream_t st) {
    constexpr int W = kAttnWarps;
    using SM = AttnSmem<D, MT, KT, STAGES, W>;
    // set on every call: the attribute is per device, and one process may drive several
    cudaError_t e = cudaFuncSetAttribute(verify_attn_kernel<T, D, MT, KT, STAGES, W>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal);
    if (e != cudaSuccess) return e;
    return launch_kernel(verify_attn_kernel<T, D, MT, KT, STAGES, W>, grid, dim3(W * 32), (size_t)SM::kTotal, st, p);
}

template <int D, int MT, int KT, int STAGES>
static constexpr int attn_smem_bytes() {
    return AttnSmem<D, MT, KT, STAGES, kAttnWarps>::kTotal;
}

// tile configuration per head dim: {KT, STAGES}; config 1 keeps two CTAs per SM resident for D=128
template <typename T, int D, int MT>
static cudaError_t launch_attn(const AttnParams& p, dim3 grid, int config, cudaStream_t st) {
    if constexpr (D == 128) {
        if (config == 2) return launch_attn_cfg<T, D, MT, 32, 3>(p, grid, st);
        if (config == 3) return launch_attn_cfg<T, D, MT, 32, 2>(p, grid, st);
        return launch_attn_cfg<T, D, MT, 16, 3>(p, grid, st);
    } else {
        if (config == 2) return launch_attn_cfg<T, D, MT, 32, 4>(p, grid, st);
        if (config == 3) return launch_attn_cfg<T, D, MT, 32, 2>(p, grid, st);
        // 16-key tiles (evidence runs): half the ring per CTA.  Measured on B200 (tools/sweep_attention.py, Llama-3.2-1B,
        // B = 64, L = 1024): 32.2 us against 30.1 us for the default -- the kernel is limited to two CTAs per SM by its
        // 172 registers, not by shared memory, and forcing 128 registers (four CTAs: the whole 512-CTA grid resident)
        // spills into the main loop (41.6 us).
        if (config == 4) return launch_attn_cfg<T, D, MT, 16, 3>(p, grid, st);
        if (config == 5) return launch_attn_cfg<T, D, MT, 16, 4>(p, grid, st);
        if (config == 6) return launch_attn_cfg<T, D, MT, 16, 6>(p, grid, st);
        return launch_attn_cfg<T, D, MT, 32, 3>(p, grid, st);
    }
}

static int attn_tile_keys(int D, int config) {
    if (D == 128) return (config == 0 || config == 1) ? 16 : 32;
    return config >= 4 ? 16 : 32;
}
static int attn_ctas_per_sm(int D, int MT, int config) {
    // shared-memory bound (227 KB per SM, 1 KB reserved per CTA)
    int bytes;
    if (D == 128) bytes = config == 2 ? 196608 : config == 3 ? 131072 : 98304;
    else bytes = config == 2 ? 131072 : config == 3 ? 65536 : config == 4 ? 49152 : config == 5 ? 65536 : config == 6 ? 98304 : 98304;
    bytes += MT * 16 * D * 2 + 1024;
    int n = (227 * 1024) / bytes;
    return n < 1 ? 1 : n;
}

template <typename T>
static cudaError_t dispatch_attn(const AttnParams& p, int D, int MT, dim3 grid, int config, cudaStream_t st) {
    if (D == 128) {
        if (MT == 1) return launch_attn<T, 128, 1>(p, grid, config, st);
        return launch_attn<T, 128, 2>(p, grid, config, st);
    }
    if (MT == 1) return launch_attn<T, 64, 1>(p, grid, config, st);
    if (MT == 2) return launch_attn<T, 64, 2>(p, grid, config, st);
    return launch_attn<T, 64, 4>(p, grid, config, st);
}

}  // namespace sd

extern "C" {

size_t sd_verify_attention_workspace_bytes(int B, int Hq, int R, int D, int max_splits) {
    if (B <= 0 || Hq <= 0 || R <= 0 || D <= 0) return 256;
    if (max_splits <= 0) max_splits = sd::kAttnDefaultMaxSplits;
    if (max_splits == 1) return 256;
    const size_t rows = (size_t)B * Hq * R;
    return rows * (size_t)max_splits * ((size_t)D + 2) * sizeof(float) + 256;
}

int sd_verify_attention_set_tuning(int config) {
    if (config < 0 || config > 6) return SD_ERR_INVALID_ARG;  // 4..6: 16-key tiles at D = 64 (D = 128 treats them as 0)
    sd::g_attn_tuning.config = config;
    return SD_OK;
}

int sd_verify_attention(const void* q, const void* k_arena, const void* v_arena, const int32_t* seq_lens, void* out,
                        void* workspace, size_t workspace_bytes, int dtype, int B, int Hq, int Hkv, int R, int D,
                        int64_t kv_stride_b, int64_t kv_stride_h, int64_t max_seq_len_host, int64_t capacity,
                        int32_t* overflow_flag, float scale, int n_splits, void* stream) {
    using namespace sd;
    if (B < 0 || Hq <= 0 || Hkv <= 0 || R <= 0 || D <= 0 || n_splits < 0) return SD_ERR_INVALID_ARG;
    if (Hq % Hkv) return SD_ERR_SHAPE;
    if (dtype != SD_F16 && dtype != SD_BF16) return SD_ERR_UNSUPPORTED;
    if (D != 64 && D != 128) return SD_ERR_UNSUPPORTED;
    if (B == 0) return SD_OK;
    if (!q || !k_arena || !v_arena || !seq_lens || !out) return SD_ERR_INVALID_ARG;
    if (kv_stride_b % 16 || kv_stride_h % 16 || reinterpret_cast<uintptr_t>(k_arena) % 16 ||
        reinterpret_cast<uintptr_t>(v_arena) % 16 || reinterpret_cast<uintptr_t>(q) % 16 ||
        reinterpret_cast<uintptr_t>(out) % 8)
        return SD_ERR_ALIGNMENT;

    const int64_t es_row = (int64_t)D * 2;
    if (capacity <= 0) capacity = kv_stride_h / es_row;  // rows between two heads of the arena
    if (capacity < R || capacity > 0x7fffffffLL || capacity * es_row > kv_stride_h) return SD_ERR_CAPACITY;
    if (max_seq_len_host >= 0 && max_seq_len_host + R > capacity) return SD_ERR_CAPACITY;

    const int mt_max = 256 / D;  // accumulator budget: MT * D <= 256 fp32 columns per warp
    if (R > 16 * mt_max) return SD_ERR_UNSUPPORTED;
    const int G = Hq / Hkv;
    int gsub = G;
    while (gsub > 1 && (gsub * R + 15) / 16 > mt_max) --gsub;
    const int ngroups = (G + gsub - 1) / gsub;
    int MT = (gsub * R + 15) / 16;
    if (MT == 3) MT = 4;

    const int config = g_attn_tuning.config;
    const int KT = attn_tile_keys(D, config);

    // split heuristic: wave-quantised cost model on the host-known length bound
    int64_t lmax = max_seq_len_host >= 0 ? max_seq_len_host + R : kv_stride_h / ((int64_t)D * 2);
    if (lmax < R) lmax = R;
    const int64_t tiles = (lmax + KT - 1) / KT;
    int max_splits = kAttnDefaultMaxSplits;
    const size_t rows_total = (size_t)B * Hq * R;
    const size_t per_split = rows_total * ((size_t)D + 2) * sizeof(float);
    const int ws_splits = (workspace && workspace_bytes > 256) ? (int)((workspace_bytes - 256) / per_split) : 1;
    if (n_splits == 0) {
        const int64_t base = (int64_t)B * Hkv * ngroups;
        const int64_t slots = (int64_t)attn_sm_count() * attn_ctas_per_sm(D, MT, config);
        double best = 1e30;
        int best_n = 1;
        for (int n = 1; n <= max_splits; ++n) {
            const double per_warp = (double)tiles / n / kAttnWarps;
            if (n > 1 && per_warp < 2.0) break;
            const double waves = (double)((base * n + slots - 1) / slots);
            const double cost = waves * (per_warp + 3.0) + (n > 1 ? 1.0 : 0.0);
            if (cost < best - 1e-9) {
                best = cost;
                best_n = n;
            }
        }
        n_splits = best_n;
        if (n_splits > ws_splits) n_splits = ws_splits < 1 ? 1 : ws_splits;
    } else if (n_splits > 1 && n_splits > ws_splits) {
        return SD_ERR_WORKSPACE;
    }
    if (n_splits > 1 && reinterpret_cast<uintptr_t>(workspace) % 16) return SD_ERR_ALIGNMENT;
    if (n_splits > 65535 || (int64_t)Hkv * ngroups > 65535 || B > 65535) return SD_ERR_UNSUPPORTED;

    AttnParams p;
    p.q = q;
    p.k = k_arena;
    p.v = v_arena;
    p.seq_lens = seq_lens;
    p.out = out;
    p.part_o = n_splits > 1 ? reinterpret_cast<float*>(workspace) : nullptr;
    p.part_ml = n_splits > 1 ? p.part_o + rows_total * (size_t)n_splits * D : nullptr;
    p.kv_stride_b = kv_stride_b;
    p.kv_stride_h = kv_stride_h;
    p.capacity = (int)capacity;
    p.overflow = overflow_flag;
    p.B = B;
    p.Hq = Hq;
    p.Hkv = Hkv;
    p.R = R;
    p.G = G;
    p.gsub = gsub;
    p.ngroups = ngroups;
    p.n_splits = n_splits;
    const float sc = scale > 0.f ? scale : 1.0f / sqrtf((float)D);
    p.scale_log2 = sc * 1.4426950408889634f;

    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    dim3 grid((unsigned)n_splits, (unsigned)(Hkv * ngroups), (unsigned)B);
    cudaError_t e = dtype == SD_BF16 ? dispatch_attn<__nv_bfloat16>(p, D, MT, grid, config, st)
                                     : dispatch_attn<__half>(p, D, MT, grid, config, st);
    if (e != cudaSuccess) return to_status(e);
    if (n_splits > 1) {
        const int wpc = 8;
        const unsigned cg = (unsigned)((rows_total + wpc - 1) / wpc);
        if (dtype == SD_BF16) {
            if (D == 128) e = launch_kernel(attn_combine_kernel<__nv_bfloat16, 128>, dim3(cg), dim3(wpc * 32), 0, st, p, (int64_t)rows_total);
            else e = launch_kernel(attn_combine_kernel<__nv_bfloat16, 64>, dim3(cg), dim3(wpc * 32), 0, st, p, (int64_t)rows_total);
        } else {
            if (D == 128) e = launch_kernel(attn_combine_kernel<__half, 128>, dim3(cg), dim3(wpc * 32), 0, st, p, (int64_t)rows_total);
            else e = launch_kernel(attn_combine_kernel<__half, 64>, dim3(cg), dim3(wpc * 32), 0, st, p, (int64_t)rows_total);
        }
    }
    return to_status(e);
}

}  // extern "C"
