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
// A second, opt-in form (verify_attn_balanced_kernel, n_splits = SD_ATTN_BALANCED) runs one CTA per resident
// slot over equal spans of key tiles; it shares the tile loop and the merge with the standard form.
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
    // balanced (persistent) form only: per-CTA partial slots and per-item tile counters (zero between launches)
    float* bal_part;
    int32_t* bal_cnt;
};

struct AttnTuning {
    int config;  // 0 heuristic, 1 = KT16/3 stages, 2 = KT32/3 stages, 3 = KT32/2 stages
};
static AttnTuning g_attn_tuning = {0};

SD_DEVINL uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SD_DEVINL void cp_async16(uint32_t dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// full 16-byte copy, compile-time offsets on both addresses (no per-copy address arithmetic)
template <int DST_OFF, int SRC_OFF>
SD_DEVINL void cp_async16_at(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0+%2], [%1+%3], 16;" ::"r"(dst), "l"(src), "n"(DST_OFF), "n"(SRC_OFF) : "memory");
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
template <int OFF>
SD_DEVINL void ldmatrix_x4_at(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4+%5];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr), "n"(OFF));
}
template <int OFF>
SD_DEVINL void ldmatrix_x4_trans_at(uint32_t (&r)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4+%5];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
                 : "r"(addr), "n"(OFF));
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

// The append kernels clamp to the capacity and raise `overflow`, so a sequence may legitimately sit AT capacity:
// never read rows beyond the (b, head) slab -- the next head's keys would give silently wrong output, and the last
// slab ends the allocation.  A clamped (or negative) length raises the flag; its output is not meaningful.
SD_DEVINL int attn_clamped_len(const AttnParams& p, int b, bool report) {
    int seq = p.seq_lens[b];
    if (seq < 0 || seq > p.capacity - p.R) {
        if (p.overflow && report) *p.overflow = 1;
        seq = seq < 0 ? 0 : p.capacity - p.R;
    }
    return seq;
}

// One work segment: the key tiles [t0, t1) of item (b, y = kv head * ngroups + head group).
struct AttnSeg {
    int b, y, seq, t0, t1;
};

// Geometry of a segment's item and the shared-memory layout.  PERSIST = false (standard form): a warp's ring stages are
// adjacent and the merge area starts at `body`.  PERSIST = true (balanced form): the ring is [stage][warp], so the last
// stage of all warps is one contiguous slab -- the merge area when it fits (rows compacted to the item's mrows) --
// and the first STAGES-1 stages can already hold the NEXT segment's tiles while this one is merged.
template <typename T, int D, int MT, int KT, int STAGES, int W, bool PERSIST>
struct AttnItem {
    using SM = AttnSmem<D, MT, KT, STAGES, W>;
    static constexpr int CPR = D / 8;  // 16-byte chunks per row
    static constexpr int kMergeStride = PERSIST ? D + 4 : SM::kMergeStride;
    int mrows, hq0, Ltot, ntw;
    int64_t qrow0;
    const char *kg, *vg;

    SD_DEVINL AttnItem(const AttnParams& p, const AttnSeg& sg) {
        const int warp = threadIdx.x >> 5;
        const int kvh = sg.y / p.ngroups, gs = sg.y - kvh * p.ngroups;
        const int g0 = gs * p.gsub;
        const int gcount = (p.G - g0) < p.gsub ? (p.G - g0) : p.gsub;
        mrows = gcount * p.R;
        hq0 = kvh * p.G + g0;
        Ltot = sg.seq + p.R;
        qrow0 = ((int64_t)sg.b * p.Hq + hq0) * p.R;
        kg = reinterpret_cast<const char*>(p.k) + (int64_t)sg.b * p.kv_stride_b + (int64_t)kvh * p.kv_stride_h;
        vg = reinterpret_cast<const char*>(p.v) + (int64_t)sg.b * p.kv_stride_b + (int64_t)kvh * p.kv_stride_h;
        ntw = (sg.t1 - sg.t0 - warp + W - 1) / W > 0 ? (sg.t1 - sg.t0 - warp + W - 1) / W : 0;  // tiles of this warp
        const int lane = threadIdx.x & 31;
#pragma unroll
        for (int i = 0; i < NCP; ++i) dsw[i] = swz<D>((lane + 32 * i) / CPR, (lane + 32 * i) % CPR);
    }
    static constexpr int NCP = KT * CPR / 32;  // 16-byte chunks per lane and tile
    uint32_t dsw[NCP];                         // swizzled byte offset of this lane's i-th chunk within a K (or V) tile
    template <int I>
    SD_DEVINL void copy_full(uint32_t sk, const char* ks, const char* vs) const {
        if constexpr (I < NCP) {
            cp_async16_at<0, I * 512>(sk + dsw[I], ks);
            cp_async16_at<SM::kTileBytes, I * 512>(sk + dsw[I], vs);
            copy_full<I + 1>(sk, ks, vs);
        }
    }
    static SD_DEVINL uint32_t stage_addr(unsigned char* body, int stage) {
        const int warp = threadIdx.x >> 5;
        return smem_addr(body) + (PERSIST ? (stage * W + warp) : (warp * STAGES + stage)) * SM::kStageBytes;
    }
    // rows a warp keeps in the merge area, and whether the merge fits the last-stage slab (PERSIST only)
    SD_DEVINL int merge_rows() const { return PERSIST ? mrows : MT * 16; }
    SD_DEVINL bool merge_in_last_stage() const {
        return PERSIST && W * mrows * (kMergeStride + 2) * 4 <= W * SM::kStageBytes;
    }
    SD_DEVINL void load_tile(const AttnSeg& sg, unsigned char* body, int j) const {  // j-th tile of this warp -> stage j % STAGES
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
        const int key0 = (sg.t0 + warp + j * W) * KT;
        const uint32_t sk = stage_addr(body, j % STAGES);
        const uint32_t sv = sk + SM::kTileBytes;
        if (key0 + KT <= Ltot) {
            // full tile (all but the last one of an item): KT * CPR contiguous 16-byte chunks; lane copies chunks
            // lane + 32 i, whose swizzled destinations are per-lane constants
            const char* ks = kg + (int64_t)key0 * (D * 2) + lane * 16;
            const char* vs = vg + (int64_t)key0 * (D * 2) + lane * 16;
            copy_full<0>(sk, ks, vs);
            return;
        }
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
    }
    // Q rows (contiguous [mrows, D] block) as the oldest cp.async group, then this warp's first STAGES-1 tiles
    SD_DEVINL void issue(const AttnParams& p, const AttnSeg& sg, unsigned char* q_s, unsigned char* body) const {
        const int tid = threadIdx.x;
        const char* qg = reinterpret_cast<const char*>(p.q) + qrow0 * (D * 2);
        for (int c = tid; c < MT * 16 * CPR; c += W * 32) {
            const int row = c / CPR, col = c - row * CPR;
            const bool valid = row < mrows;
            cp_async16(smem_addr(q_s) + swz<D>(row, col), qg + (valid ? ((int64_t)row * (D * 2) + col * 16) : 0), valid ? 16 : 0);
        }
        cp_async_commit();
        __syncwarp();  // (balanced form: every lane is done with the previous segment's tiles in these stages)
#pragma unroll
        for (int j = 0; j < STAGES - 1; ++j) {
            if (j < ntw) load_tile(sg, body, j);
            cp_async_commit();
        }
    }
};

// Main loop of a segment whose Q tile and first STAGES-1 tiles were issued (AttnItem::issue): the warps' private rings,
// S = Q K^T, online softmax, O += P V.  Leaves the warp's (o, mrow, lrow) in registers.
template <typename T, int D, int MT, int KT, int STAGES, int W, bool PERSIST>
SD_DEVINL void attn_main(const AttnParams& p, const AttnItem<T, D, MT, KT, STAGES, W, PERSIST>& it, const AttnSeg& sg,
                         unsigned char* q_s, unsigned char* body, float (&o)[MT][D / 8][4], float (&mrow)[MT][2],
                         float (&lrow)[MT][2]) {
    using SM = AttnSmem<D, MT, KT, STAGES, W>;
    constexpr int NKB = KT / 16;      // 16-key blocks per tile
    constexpr int NKS = D / 16;       // k-steps of the QK^T contraction
    constexpr int NDT = D / 8;        // 8-wide n-tiles of O
    constexpr bool QREG = (MT * D <= 128);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int mi = lane >> 3, r8 = lane & 7;  // ldmatrix: matrix index / row within it
    const int seq = sg.seq, mrows = it.mrows, ntw = it.ntw;

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

    // per-lane byte offsets of the ldmatrix rows inside a tile (the 16-key block adds a compile-time offset)
    uint32_t kofs[NKS], vofs[NDT / 2];
#pragma unroll
    for (int ks = 0; ks < NKS; ++ks) kofs[ks] = swz<D>((mi >> 1) * 8 + r8, 2 * ks + (mi & 1));
#pragma unroll
    for (int ntp = 0; ntp < NDT / 2; ++ntp) vofs[ntp] = swz<D>((mi & 1) * 8 + r8, 2 * ntp + (mi >> 1));

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
        if (j + STAGES - 1 < ntw) it.load_tile(sg, body, j + STAGES - 1);
        cp_async_commit();

        const uint32_t sk = it.stage_addr(body, j % STAGES);
        const uint32_t sv = sk + SM::kTileBytes;
        const int key0 = (sg.t0 + warp + j * W) * KT;

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
            const uint32_t ka = sk + kofs[ks];
            static_assert(NKB <= 2, "16-key blocks per tile");
            {
                uint32_t kf[4];
                ldmatrix_x4_at<0>(kf, ka);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816<T>(s[mt][0], af[mt], kf[0], kf[1]);
                    mma16816<T>(s[mt][1], af[mt], kf[2], kf[3]);
                }
            }
            if constexpr (NKB == 2) {
                uint32_t kf[4];
                ldmatrix_x4_at<16 * D * 2>(kf, ka);  // (rows + 16: same swizzle phase)
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816<T>(s[mt][2], af[mt], kf[0], kf[1]);
                    mma16816<T>(s[mt][3], af[mt], kf[2], kf[3]);
                }
            }
        }

        // ---- online softmax (log2 domain) ----
        // The causal mask only exists on tiles that reach past seq (the last one or two of an item): a warp-uniform
        // branch, so the other tiles pay no compare / select.  The row maximum is taken on the raw scores and scaled
        // once (scale > 0: max commutes), the scale itself rides in the FMA that feeds exp2, and O is only rescaled
        // when some row's maximum moved (alpha == 1 exactly otherwise: ex2(0) = 1).
        if (key0 + KT > seq + 1) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nt = 0; nt < 2 * NKB; ++nt)
#pragma unroll
                        for (int e = 0; e < 2; ++e)
                            if (key0 + nt * 8 + 2 * tig + e > lim[mt][h]) s[mt][nt][2 * h + e] = -INFINITY;
        }
        uint32_t pa[MT][NKB][4];
        float alpha[MT][2];
        bool moved = false;
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float mx = -INFINITY;
#pragma unroll
                for (int nt = 0; nt < 2 * NKB; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) mx = fmaxf(mx, s[mt][nt][2 * h + e]);
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
                mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
                const float m_new = fmaxf(mrow[mt][h], mx * p.scale_log2);
                const float m_use = m_new == -INFINITY ? 0.f : m_new;
                alpha[mt][h] = fast_exp2(mrow[mt][h] - m_use);
                moved |= alpha[mt][h] != 1.f;
                mrow[mt][h] = m_new;
                float sum = 0.f;
#pragma unroll
                for (int nt = 0; nt < 2 * NKB; ++nt)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const float pv = fast_exp2(fmaf(s[mt][nt][2 * h + e], p.scale_log2, -m_use));
                        s[mt][nt][2 * h + e] = pv;
                        sum += pv;
                    }
                lrow[mt][h] = lrow[mt][h] * alpha[mt][h] + sum;
            }
#pragma unroll
            for (int kb = 0; kb < NKB; ++kb) {
                pa[mt][kb][0] = pack2<T>(s[mt][2 * kb][0], s[mt][2 * kb][1]);
                pa[mt][kb][1] = pack2<T>(s[mt][2 * kb][2], s[mt][2 * kb][3]);
                pa[mt][kb][2] = pack2<T>(s[mt][2 * kb + 1][0], s[mt][2 * kb + 1][1]);
                pa[mt][kb][3] = pack2<T>(s[mt][2 * kb + 1][2], s[mt][2 * kb + 1][3]);
            }
        }
        if (__any_sync(0xffffffffu, moved)) {
#pragma unroll
            for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                for (int h = 0; h < 2; ++h)
#pragma unroll
                    for (int nt = 0; nt < NDT; ++nt) {
                        o[mt][nt][2 * h] *= alpha[mt][h];
                        o[mt][nt][2 * h + 1] *= alpha[mt][h];
                    }
        }

        // ---- O += P V ----
#pragma unroll
        for (int ntp = 0; ntp < NDT / 2; ++ntp) {
            const uint32_t va = sv + vofs[ntp];
            {
                uint32_t vf[4];
                ldmatrix_x4_trans_at<0>(vf, va);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816<T>(o[mt][2 * ntp], pa[mt][0], vf[0], vf[1]);
                    mma16816<T>(o[mt][2 * ntp + 1], pa[mt][0], vf[2], vf[3]);
                }
            }
            if constexpr (NKB == 2) {
                uint32_t vf[4];
                ldmatrix_x4_trans_at<16 * D * 2>(vf, va);
#pragma unroll
                for (int mt = 0; mt < MT; ++mt) {
                    mma16816<T>(o[mt][2 * ntp], pa[mt][1], vf[0], vf[1]);
                    mma16816<T>(o[mt][2 * ntp + 1], pa[mt][1], vf[2], vf[3]);
                }
            }
        }
    }
}

// CTA merge of the W warps' (max, sum, O) through `area` and the segment's result: po == nullptr -> the range was the
// whole item, normalised rows go to p.out; otherwise the unnormalised fp32 partial of row m goes to po + m * po_stride
// (D floats) and its {max, sum} to pml + m * pml_stride.  Starts with the barrier that ends every warp's main loop.
template <typename T, int D, int MT, int KT, int STAGES, int W, bool PERSIST>
SD_DEVINL void attn_merge_write(const AttnParams& p, const AttnItem<T, D, MT, KT, STAGES, W, PERSIST>& it, float* area,
                                const float (&o)[MT][D / 8][4], const float (&mrow)[MT][2], const float (&lrow)[MT][2],
                                float* po, const int64_t po_stride, float* pml, const int64_t pml_stride) {
    constexpr int NDT = D / 8;
    constexpr int kStride = AttnItem<T, D, MT, KT, STAGES, W, PERSIST>::kMergeStride;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int mrows = it.mrows, rows = it.merge_rows();
    __syncthreads();  // every warp is done with its ring stages that the area overlaps

    float* om = area;                    // [W][rows][kStride]
    float* mm = om + W * rows * kStride; // [W][rows]
    float* lm = mm + W * rows;           // [W][rows]
#pragma unroll
    for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            float l = lrow[mt][h];
            l += __shfl_xor_sync(0xffffffffu, l, 1);
            l += __shfl_xor_sync(0xffffffffu, l, 2);
            const int row = mt * 16 + g + 8 * h;
            if (row < rows) {
                float* orow = om + ((size_t)warp * rows + row) * kStride;
#pragma unroll
                for (int nt = 0; nt < NDT; ++nt)
                    *reinterpret_cast<float2*>(orow + nt * 8 + 2 * tig) = make_float2(o[mt][nt][2 * h], o[mt][nt][2 * h + 1]);
                if (tig == 0) {
                    mm[warp * rows + row] = mrow[mt][h];
                    lm[warp * rows + row] = l;
                }
            }
        }
    }
    __syncthreads();

    for (int idx = tid; idx < mrows * (D / 4); idx += W * 32) {
        const int m = idx / (D / 4), c4 = idx - m * (D / 4);
        float mstar = -INFINITY;
#pragma unroll
        for (int w = 0; w < W; ++w) mstar = fmaxf(mstar, mm[w * rows + m]);
        const float m_use = mstar == -INFINITY ? 0.f : mstar;
        float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
        float l = 0.f;
#pragma unroll
        for (int w = 0; w < W; ++w) {
            const float sc = fast_exp2(mm[w * rows + m] - m_use);
            const float4 x = *reinterpret_cast<const float4*>(om + ((size_t)w * rows + m) * kStride + c4 * 4);
            acc.x += sc * x.x;
            acc.y += sc * x.y;
            acc.z += sc * x.z;
            acc.w += sc * x.w;
            l += sc * lm[w * rows + m];
        }
        if (po == nullptr) {
            const int64_t grow = it.qrow0 + m;
            const float inv = l > 0.f ? 1.f / l : 0.f;
            uint2 pk;
            pk.x = pack2<T>(acc.x * inv, acc.y * inv);
            pk.y = pack2<T>(acc.z * inv, acc.w * inv);
            *reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.out) + (grow * D + c4 * 4) * 2) = pk;
        } else {
            *reinterpret_cast<float4*>(po + m * po_stride + c4 * 4) = acc;
            if (c4 == 0) {
                pml[m * pml_stride] = mstar;
                pml[m * pml_stride + 1] = l;
            }
        }
    }
}

template <typename T, int D, int MT, int KT, int STAGES, int W>
__global__ void __launch_bounds__(W * 32) verify_attn_kernel(const AttnParams p) {
    using SM = AttnSmem<D, MT, KT, STAGES, W>;
    extern __shared__ __align__(128) unsigned char smem[];
    unsigned char* q_s = smem;
    unsigned char* body = smem + SM::kQBytes;
    pdl_wait();
    pdl_trigger();
    const int split = blockIdx.x;
    AttnSeg sg;
    sg.y = blockIdx.y;
    sg.b = blockIdx.z;
    sg.seq = attn_clamped_len(p, sg.b, threadIdx.x == 0);
    const int tiles_total = (sg.seq + p.R + KT - 1) / KT;
    const int tiles_per_split = (tiles_total + p.n_splits - 1) / p.n_splits;
    sg.t0 = split * tiles_per_split;
    sg.t1 = (sg.t0 + tiles_per_split) < tiles_total ? (sg.t0 + tiles_per_split) : tiles_total;
    const AttnItem<T, D, MT, KT, STAGES, W, false> it(p, sg);
    float o[MT][D / 8][4];
    float mrow[MT][2], lrow[MT][2];
    it.issue(p, sg, q_s, body);
    attn_main<T, D, MT, KT, STAGES, W, false>(p, it, sg, q_s, body, o, mrow, lrow);
    cp_async_wait<0>();
    if (p.n_splits == 1) {
        attn_merge_write<T, D, MT, KT, STAGES, W, false>(p, it, reinterpret_cast<float*>(body), o, mrow, lrow, nullptr, 0, nullptr, 0);
    } else {
        attn_merge_write<T, D, MT, KT, STAGES, W, false>(p, it, reinterpret_cast<float*>(body), o, mrow, lrow,
                                                         p.part_o + (it.qrow0 * p.n_splits + split) * D, (int64_t)p.n_splits * D,
                                                         p.part_ml + (it.qrow0 * p.n_splits + split) * 2, (int64_t)p.n_splits * 2);
    }
}

// ---- balanced (persistent) form ---------------------------------------------------------------------------------
// The standard form runs one CTA per item; with 512 items on 296 resident CTAs every slot serves two CTAs, all of the
// first wave end together (equal lengths), and the whole GPU goes through a merge / relaunch / ring-refill bubble at
// the same moment.  Here the grid is exactly the resident CTAs and the key tiles of all items form ONE sequence
// (item-major) cut into gridDim.x equal spans: a CTA walks its span segment by segment (the rest of an item, whole
// items, the head of an item), so every CTA streams the same number of bytes, the segment boundaries of different
// CTAs fall at different times, and a warp requests the first tiles (and the CTA the Q rows) of the NEXT segment
// before the current one is merged -- the bytes in flight per SM do not drop at a boundary.  An item cut by a span
// boundary is finished by whichever of its CTAs arrives last (per-item tile counter, no waiting anywhere): partials
// in p.bal_part[cta][2] -- a CTA's first segment uses slot 0, a later (necessarily its last) partial segment slot 1.
// Counters are zero between launches (the last arriver resets its counter);
// sd_verify_attention_balanced_workspace_init zeroes them once.
constexpr int kAttnBalancedMaxB = 512;

template <int D, int MT>
struct AttnBalSlot {
    static constexpr int kFloats = MT * 16 * (D + 2);  // [rows][D] partial O, then [rows][2] {max, sum}
};

struct AttnWalk {  // a CTA's position in the tile sequence
    long long pos, hi;
    int b;
};

// the segment at w.pos (w.pos < w.hi); advances w.pos past it and returns its slot-relevant facts
SD_DEVINL AttnSeg attn_next_segment(const AttnParams& p, const int* pre, const int Y, AttnWalk& w, int* tb_out, long long* gfirst_out) {
    while ((long long)pre[w.b + 1] * Y <= w.pos) ++w.b;
    const int tb = pre[w.b + 1] - pre[w.b];
    const long long item0 = (long long)pre[w.b] * Y;  // first tile of sequence b's first item
    AttnSeg sg;
    sg.b = w.b;
    sg.y = (int)((w.pos - item0) / tb);
    sg.t0 = (int)(w.pos - item0 - (long long)sg.y * tb);
    sg.t1 = (w.hi - w.pos) < (long long)(tb - sg.t0) ? sg.t0 + (int)(w.hi - w.pos) : tb;
    sg.seq = attn_clamped_len(p, w.b, false);  // (the flag is raised by the scan)
    *tb_out = tb;
    *gfirst_out = item0 + (long long)sg.y * tb;
    w.pos += sg.t1 - sg.t0;
    return sg;
}

template <typename T, int D, int MT, int KT, int STAGES, int W>
__global__ void __launch_bounds__(W * 32, 2) verify_attn_balanced_kernel(const AttnParams p) {
    using SM = AttnSmem<D, MT, KT, STAGES, W>;
    using Item = AttnItem<T, D, MT, KT, STAGES, W, true>;
    constexpr int kSlot = AttnBalSlot<D, MT>::kFloats;
    extern __shared__ __align__(128) unsigned char smem[];
    // Q tile, double-buffered across segments: buffer i at smem + i * kQBytes
    unsigned char* body = smem + 2 * SM::kQBytes;
    int* pre = reinterpret_cast<int*>(body + SM::kBodyBytes);  // [B + 1] tiles before sequence b (per item column)
    int* last_flag = pre + kAttnBalancedMaxB + 1;
    pdl_wait();
    pdl_trigger();
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int Y = p.Hkv * p.ngroups;

    for (int b = tid; b < p.B; b += W * 32) pre[b + 1] = (attn_clamped_len(p, b, true) + p.R + KT - 1) / KT;
    if (tid == 0) pre[0] = 0;
    __syncthreads();
    if (warp == 0) {  // inclusive scan of pre[1..B] in place
        int carry = 0;
        for (int base = 1; base <= p.B; base += 32) {
            const int i = base + lane;
            int v = i <= p.B ? pre[i] : 0;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int n = __shfl_up_sync(0xffffffffu, v, o);
                if (lane >= o) v += n;
            }
            if (i <= p.B) pre[i] = carry + v;
            carry += __shfl_sync(0xffffffffu, v, 31);
        }
    }
    __syncthreads();

    const long long total = (long long)pre[p.B] * Y;
    const long long span = (total + gridDim.x - 1) / gridDim.x;
    AttnWalk w;
    w.pos = (long long)blockIdx.x * span;
    w.hi = w.pos + span < total ? w.pos + span : total;
    w.b = 0;
    if (w.pos >= w.hi) return;

    int tb, seg = 0, qi = 0;
    long long gfirst;
    AttnSeg sg = attn_next_segment(p, pre, Y, w, &tb, &gfirst);
    Item(p, sg).issue(p, sg, smem, body);
    for (;;) {
        const Item it(p, sg);
        const bool more = w.pos < w.hi;
        const bool whole = sg.t0 == 0 && sg.t1 == tb;
        float o[MT][D / 8][4];
        float mrow[MT][2], lrow[MT][2];
        attn_main<T, D, MT, KT, STAGES, W, true>(p, it, sg, smem + qi * SM::kQBytes, body, o, mrow, lrow);

        // the next segment: known now, requested before this one is merged when the merge stays out of its stages
        AttnSeg nsg = sg;
        int ntb = tb;
        long long ngfirst = gfirst;
        if (more) nsg = attn_next_segment(p, pre, Y, w, &ntb, &ngfirst);
        const bool early = it.merge_in_last_stage();
        if (more && early) Item(p, nsg).issue(p, nsg, smem + (qi ^ 1) * SM::kQBytes, body);

        float* slot = p.bal_part + ((size_t)blockIdx.x * 2 + (seg ? 1 : 0)) * kSlot;
        float* area = reinterpret_cast<float*>(early ? body + (STAGES - 1) * W * SM::kStageBytes : body);
        attn_merge_write<T, D, MT, KT, STAGES, W, true>(p, it, area, o, mrow, lrow, whole ? nullptr : slot, D,
                                                        slot + MT * 16 * D, 2);
        if (!whole) {
            const int item = sg.b * Y + sg.y;
            __threadfence();  // this thread's part of the partial is visible device-wide ...
            __syncthreads();
            if (tid == 0) {   // ... before the count says so
                const int old = atomicAdd(p.bal_cnt + item, sg.t1 - sg.t0);
                const int is_last = old + (sg.t1 - sg.t0) == tb;
                if (is_last) p.bal_cnt[item] = 0;  // nobody else touches it in this launch: ready for the next one
                *last_flag = is_last;
            }
            __syncthreads();
            if (*last_flag) {
                __threadfence();
                const int c_first = (int)(gfirst / span), c_last = (int)((gfirst + tb - 1) / span);  // the item's tiles: [gfirst, gfirst + tb)
                const int first_slot = gfirst == (long long)c_first * span ? 0 : 1;
                for (int idx = tid; idx < it.mrows * (D / 4); idx += W * 32) {
                    const int m = idx / (D / 4), c4 = idx - m * (D / 4);
                    float mstar = -INFINITY;
                    for (int c = c_first; c <= c_last; ++c) {
                        const float* sl = p.bal_part + ((size_t)c * 2 + (c == c_first ? first_slot : 0)) * kSlot;
                        mstar = fmaxf(mstar, __ldcg(sl + MT * 16 * D + m * 2));
                    }
                    const float m_use = mstar == -INFINITY ? 0.f : mstar;
                    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
                    float l = 0.f;
                    for (int c = c_first; c <= c_last; ++c) {
                        const float* sl = p.bal_part + ((size_t)c * 2 + (c == c_first ? first_slot : 0)) * kSlot;
                        const float sc = fast_exp2(__ldcg(sl + MT * 16 * D + m * 2) - m_use);
                        const float4 x = __ldcg(reinterpret_cast<const float4*>(sl + m * D + c4 * 4));
                        acc.x += sc * x.x;
                        acc.y += sc * x.y;
                        acc.z += sc * x.z;
                        acc.w += sc * x.w;
                        l += sc * __ldcg(sl + MT * 16 * D + m * 2 + 1);
                    }
                    const float inv = l > 0.f ? 1.f / l : 0.f;
                    uint2 pk;
                    pk.x = pack2<T>(acc.x * inv, acc.y * inv);
                    pk.y = pack2<T>(acc.z * inv, acc.w * inv);
                    *reinterpret_cast<uint2*>(reinterpret_cast<char*>(p.out) + ((it.qrow0 + m) * D + c4 * 4) * 2) = pk;
                }
            }
        }
        if (!more) break;
        if (!early) {
            __syncthreads();  // the merge area was the front of the ring: everyone is done reading it
            Item(p, nsg).issue(p, nsg, smem + (qi ^ 1) * SM::kQBytes, body);
        }
        sg = nsg;
        tb = ntb;
        gfirst = ngfirst;
        qi ^= 1;
        ++seg;
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

// ---- balanced form: default tile configuration per head dim only (D = 128: 16 keys x 3 stages, D = 64: 32 x 3) ----
template <int D>
struct AttnBalCfg {
    static constexpr int KT = D == 128 ? 16 : 32;
    static constexpr int STAGES = 3;
};
constexpr int kAttnBalExtraSmem = ((kAttnBalancedMaxB + 2) * 4 + 15) / 16 * 16;

template <typename T, int D, int MT>
static cudaError_t attn_balanced_setup(int* ctas_per_sm) {
    constexpr int W = kAttnWarps;
    using SM = AttnSmem<D, MT, AttnBalCfg<D>::KT, AttnBalCfg<D>::STAGES, W>;
    auto kern = verify_attn_balanced_kernel<T, D, MT, AttnBalCfg<D>::KT, AttnBalCfg<D>::STAGES, W>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SM::kTotal + SM::kQBytes + kAttnBalExtraSmem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kern, W * 32, (size_t)(SM::kTotal + SM::kQBytes + kAttnBalExtraSmem));
}
template <typename T, int D, int MT>
static cudaError_t attn_balanced_launch(const AttnParams& p, unsigned grid, cudaStream_t st) {
    constexpr int W = kAttnWarps;
    using SM = AttnSmem<D, MT, AttnBalCfg<D>::KT, AttnBalCfg<D>::STAGES, W>;
    return launch_kernel(verify_attn_balanced_kernel<T, D, MT, AttnBalCfg<D>::KT, AttnBalCfg<D>::STAGES, W>, dim3(grid),
                         dim3(W * 32), (size_t)(SM::kTotal + SM::kQBytes + kAttnBalExtraSmem), st, p);
}

// dispatch over (dtype, D, MT); op 0 = setup (attribute + occupancy), 1 = launch
static cudaError_t attn_balanced_dispatch(int op, int dtype, int D, int MT, const AttnParams* p, unsigned grid,
                                          cudaStream_t st, int* ctas_per_sm) {
#define SD_BAL_CASE(TT, DD, MM) \
    return op == 0 ? attn_balanced_setup<TT, DD, MM>(ctas_per_sm) : attn_balanced_launch<TT, DD, MM>(*p, grid, st)
    if (dtype == SD_BF16) {
        if (D == 128) { if (MT == 1) { SD_BAL_CASE(__nv_bfloat16, 128, 1); } SD_BAL_CASE(__nv_bfloat16, 128, 2); }
        if (MT == 1) { SD_BAL_CASE(__nv_bfloat16, 64, 1); }
        if (MT == 2) { SD_BAL_CASE(__nv_bfloat16, 64, 2); }
        SD_BAL_CASE(__nv_bfloat16, 64, 4);
    }
    if (D == 128) { if (MT == 1) { SD_BAL_CASE(__half, 128, 1); } SD_BAL_CASE(__half, 128, 2); }
    if (MT == 1) { SD_BAL_CASE(__half, 64, 1); }
    if (MT == 2) { SD_BAL_CASE(__half, 64, 2); }
    SD_BAL_CASE(__half, 64, 4);
#undef SD_BAL_CASE
}

// query-tile geometry shared by both forms
struct AttnGeom {
    int G, gsub, ngroups, MT;
};
static bool attn_geometry(int Hq, int Hkv, int R, int D, AttnGeom* o) {
    const int mt_max = 256 / D;  // accumulator budget: MT * D <= 256 fp32 columns per warp
    if (R > 16 * mt_max) return false;
    o->G = Hq / Hkv;
    int gsub = o->G;
    while (gsub > 1 && (gsub * R + 15) / 16 > mt_max) --gsub;
    o->gsub = gsub;
    o->ngroups = (o->G + gsub - 1) / gsub;
    o->MT = (gsub * R + 15) / 16;
    if (o->MT == 3) o->MT = 4;
    return true;
}

// resident CTAs of the balanced kernel on the current device (cached per instantiation)
static int attn_balanced_grid(int dtype, int D, int MT, cudaError_t* err) {
    static int cache[2][2][5] = {};
    int& c = cache[dtype == SD_BF16 ? 0 : 1][D == 128 ? 0 : 1][MT];
    *err = cudaSuccess;
    if (c == 0) {
        int per_sm = 0;
        *err = attn_balanced_dispatch(0, dtype, D, MT, nullptr, 0, nullptr, &per_sm);
        if (*err != cudaSuccess) return 0;
        c = per_sm < 1 ? -1 : per_sm * attn_sm_count();
    }
    return c;
}

static size_t attn_balanced_counter_bytes(int B, int Hkv, int ngroups) {
    return (((size_t)B * Hkv * ngroups * sizeof(int32_t)) + 255) / 256 * 256;
}

}  // namespace sd

extern "C" {

size_t sd_verify_attention_balanced_workspace_bytes(int B, int Hq, int Hkv, int R, int D) {
    using namespace sd;
    AttnGeom g;
    if (B <= 0 || Hq <= 0 || Hkv <= 0 || R <= 0 || (D != 64 && D != 128) || Hq % Hkv || !attn_geometry(Hq, Hkv, R, D, &g)) return 256;
    // slots for the largest grid any device of this generation can hold (4 CTAs per SM is beyond what the tiles allow)
    const size_t slots = (size_t)attn_sm_count() * 4 * 2;
    return attn_balanced_counter_bytes(B, Hkv, g.ngroups) + slots * (size_t)g.MT * 16 * (D + 2) * sizeof(float);
}

int sd_verify_attention_balanced_workspace_init(void* workspace, size_t workspace_bytes, void* stream) {
    if (!workspace || workspace_bytes == 0) return SD_ERR_INVALID_ARG;
    return sd::to_status(cudaMemsetAsync(workspace, 0, workspace_bytes, reinterpret_cast<cudaStream_t>(stream)));
}

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
    if (B < 0 || Hq <= 0 || Hkv <= 0 || R <= 0 || D <= 0 || (n_splits < 0 && n_splits != SD_ATTN_BALANCED)) return SD_ERR_INVALID_ARG;
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

    AttnGeom geom;
    if (!attn_geometry(Hq, Hkv, R, D, &geom)) return SD_ERR_UNSUPPORTED;
    const int G = geom.G, gsub = geom.gsub, ngroups = geom.ngroups, MT = geom.MT;
    const float sc0 = scale > 0.f ? scale : 1.0f / sqrtf((float)D);

    if (n_splits == SD_ATTN_BALANCED) {
        // persistent form: one CTA per resident slot, equal spans of key tiles (see verify_attn_balanced_kernel)
        if (B > kAttnBalancedMaxB) return SD_ERR_UNSUPPORTED;
        if (!workspace || reinterpret_cast<uintptr_t>(workspace) % 16) return SD_ERR_ALIGNMENT;
        cudaError_t ce;
        const int grid = attn_balanced_grid(dtype, D, MT, &ce);
        if (ce != cudaSuccess) return to_status(ce);
        if (grid <= 0) return SD_ERR_UNSUPPORTED;
        const size_t cnt_bytes = attn_balanced_counter_bytes(B, Hkv, ngroups);
        if (workspace_bytes < cnt_bytes + (size_t)grid * 2 * MT * 16 * (D + 2) * sizeof(float)) return SD_ERR_WORKSPACE;
        AttnParams p;
        p.q = q;
        p.k = k_arena;
        p.v = v_arena;
        p.seq_lens = seq_lens;
        p.out = out;
        p.part_o = nullptr;
        p.part_ml = nullptr;
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
        p.n_splits = 1;
        p.scale_log2 = sc0 * 1.4426950408889634f;
        p.bal_cnt = reinterpret_cast<int32_t*>(workspace);
        p.bal_part = reinterpret_cast<float*>(reinterpret_cast<char*>(workspace) + cnt_bytes);
        return to_status(attn_balanced_dispatch(1, dtype, D, MT, &p, (unsigned)grid, reinterpret_cast<cudaStream_t>(stream), nullptr));
    }

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
    p.scale_log2 = sc0 * 1.4426950408889634f;
    p.bal_part = nullptr;
    p.bal_cnt = nullptr;

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
