// Shared device helpers for the sm_100a verify-path kernels.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sd_verify.h"

#define SD_DEVINL __device__ __forceinline__

namespace sd {

constexpr int kWarp = 32;

// ---- programmatic dependent launch (PDL): every kernel of this library is launched with the
// programmatic-stream-serialization attribute, so its CTAs may become resident while the previous
// kernel on the stream drains; griddepcontrol.wait blocks until that kernel has completed and its
// writes are visible, so it must precede the first global access.  Both instructions are no-ops for a
// kernel launched without the attribute (sd_set_pdl(0), or the previous work not being a kernel).
SD_DEVINL void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
SD_DEVINL void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

int pdl_enabled();  // sd_api.cu

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                 Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

SD_DEVINL unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// ---- bounded device-side waits ------------------------------------------------------------------
// Every cross-CTA / cross-GPU poll of this library (finalizer -> partial slots, append -> sequence
// flags, gather -> peer words) goes through a SpinGuard: after kSpinBudgetNs of %globaltimer it gives
// up, records WHAT it was waiting for in a per-translation-unit device word and lets the kernel run to
// its end (results of that launch are garbage, memory accesses stay in bounds).  The host reads the
// word with sd_device_error() (sd_api.cu) -- a dead peer, a stale workspace or a broken dispatch-order
// assumption becomes an error code instead of an unkillable GPU.
enum SpinSite : unsigned {
    kSpinPartials = 1u,   // verify finalizer waiting for the partial slots of its rows
    kSpinSeqFlag = 2u,    // append waiting for a sequence's completion flag
    kSpinAllFlags = 4u,   // gather / last CTA waiting for every sequence flag of the launch
    kSpinPeerWords = 8u,  // exchange gather waiting for a peer's tagged words
    kSpinStartGate = 16u  // sd_exchange_start_gate waiting for the peers' arrival words
};
constexpr unsigned long long kSpinBudgetNs = 2000000000ull;  // 2 s: five orders of magnitude above any legitimate wait
static __device__ unsigned g_spin_error = 0u;                // one copy per translation unit (no -rdc)

struct SpinGuard {
    unsigned long long t0 = 0ull;
    unsigned polls = 0u;
    // true = keep waiting; false = budget exhausted (error recorded)
    SD_DEVINL bool keep_waiting(unsigned site) {
        if ((++polls & 127u) != 0u) return true;
        const unsigned long long now = global_timer_ns();
        if (t0 == 0ull) {
            t0 = now;
            return true;
        }
        if (now - t0 < kSpinBudgetNs) return true;
#ifdef SD_SPIN_TWO_STRIKE
        // Default-off (-DSD_SPIN_TWO_STRIKE): the start gate's rule (sd_exchange.cu, GateSpinGuard) for every wait -- one
        // step of %globaltimer re-arms the window instead of failing the launch; the strike lives in the top bit of
        // `polls` (no register).  Compiles to the same register counts; not measured on a GPU yet,
        // so the shipped kernels are built without it.
        if (!(polls >> 31)) {
            polls |= 0x80000000u;
            t0 = now;
            return true;
        }
#endif
        atomicOr(&g_spin_error, site);
        return false;
    }
};

// ---- streaming global loads: read-once data must not displace the KV working set in L1 ----
SD_DEVINL uint4 ldg_stream_v4(const void* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// Scheduling fence for a batch of loads: placed after the LAST load of a batch (one call per loaded
// vector), it keeps the compiler from sinking a load next to its first use -- which would serialise
// load -> use -> load and leave one request in flight per thread instead of the whole batch.
SD_DEVINL void pin_loaded(uint4& v) { asm volatile("" : "+r"(v.x), "+r"(v.y), "+r"(v.z), "+r"(v.w)); }

SD_DEVINL void stg_stream_v4(void* p, const uint4& v) {
    asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(v.x), "r"(v.y),
                 "r"(v.z), "r"(v.w)
                 : "memory");
}

// ---- order-preserving key for a float under torch.argmax semantics ----
// larger key == "more maximal": NaN > +inf > ... > +0 == -0 > ... > -inf.
SD_DEVINL uint32_t order_key(float f) {
    uint32_t u = __float_as_uint(f);
    if ((u & 0x7fffffffu) > 0x7f800000u) return 0xffffffffu;  // NaN: maximal, all NaNs equal
    if (u == 0x80000000u) u = 0u;                             // -0 == +0
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

// inverse of order_key for non-NaN keys (NaN key -> quiet NaN)
SD_DEVINL float key_to_float(uint32_t k) {
    if (k == 0xffffffffu) return __uint_as_float(0x7fc00000u);
    uint32_t u = (k & 0x80000000u) ? (k & 0x7fffffffu) : ~k;
    return __uint_as_float(u);
}

template <typename T>
SD_DEVINL float to_float(T v);
template <>
SD_DEVINL float to_float<float>(float v) { return v; }
template <>
SD_DEVINL float to_float<__half>(__half v) { return __half2float(v); }
template <>
SD_DEVINL float to_float<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

// ---- packed max with NaN propagation (canonical NaN result), per 32-bit word ----
template <typename T>
struct Packed;  // operations on one 32-bit word holding 1 (fp32) or 2 (16-bit) elements

template <>
struct Packed<float> {
    static constexpr int kPerWord = 1;
    static constexpr uint32_t kNegInf = 0xff800000u;
    SD_DEVINL static uint32_t vmax(uint32_t a, uint32_t b) {
        uint32_t d;
        asm("max.NaN.f32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
    // fold the word to a scalar replicated in every lane of the word, with -0 -> +0
    SD_DEVINL static uint32_t fold(uint32_t w) {
        uint32_t d;
        asm("add.rn.f32 %0, %1, 0f00000000;" : "=r"(d) : "r"(w));
        return d;
    }
    SD_DEVINL static float scalar(uint32_t w) { return __uint_as_float(w); }
};

template <>
struct Packed<__half> {
    static constexpr int kPerWord = 2;
    static constexpr uint32_t kNegInf = 0xfc00fc00u;
    SD_DEVINL static uint32_t vmax(uint32_t a, uint32_t b) {
        uint32_t d;
        asm("max.NaN.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
    SD_DEVINL static uint32_t fold(uint32_t w) {
        uint32_t sw = __byte_perm(w, w, 0x1032);
        uint32_t m = vmax(w, sw), d, z = 0u;
        asm("add.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(m), "r"(z));
        return d;
    }
    SD_DEVINL static float scalar(uint32_t w) {
        return __half2float(__ushort_as_half((unsigned short)(w & 0xffffu)));
    }
};

template <>
struct Packed<__nv_bfloat16> {
    static constexpr int kPerWord = 2;
    static constexpr uint32_t kNegInf = 0xff80ff80u;
    SD_DEVINL static uint32_t vmax(uint32_t a, uint32_t b) {
        uint32_t d;
        asm("max.NaN.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
        return d;
    }
    SD_DEVINL static uint32_t fold(uint32_t w) {
        uint32_t sw = __byte_perm(w, w, 0x1032);
        uint32_t m = vmax(w, sw), d, z = 0u;
        asm("add.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(m), "r"(z));
        return d;
    }
    SD_DEVINL static float scalar(uint32_t w) { return __uint_as_float(w << 16); }
};

// 2^x, MUFU.EX2 (approx, flush-to-zero): relative error ~2^-22, exact for x == 0
SD_DEVINL float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

SD_DEVINL int load_id(const void* ids, int id_dtype, int64_t i) {
    // ids outside int32 can never equal an argmax index; clamp them to -1 so equality fails
    if (id_dtype == SD_I32) return reinterpret_cast<const int32_t*>(ids)[i];
    long long v = reinterpret_cast<const long long*>(ids)[i];
    return (v < 0 || v > 0x7fffffffLL) ? -1 : (int)v;
}

// ---- peer-memory result exchange, consumer side (include/sd_verify.h: sd_exchange) ----
// Exchange buffer: 256-byte header {u32 epoch} + u64 words[4][world * rows_per_rank][width], each word
// {epoch + 1, value} written by ONE 8-byte store of the producing verify kernel into slot epoch & 3.
// Called by a whole CTA: gathers the records of epoch (current - lag) -- polls every expected word until
// its tag matches, writes the values to `out` ([world * rows_per_rank, width] int32, nullable; rows
// beyond a rank's batch are left untouched) -- then advances the epoch.  lag = 0: the step's own
// records (all ranks run in lockstep); lag = 1: the previous step's (a rank may run one step ahead of
// the slowest; nothing is gathered at epoch 0).  4 slots cover lag <= 1: a producer is never more
// than 2 epochs ahead of a slot that is still being read.
// Rank r holds total_rows / world (+1 for the first total_rows % world ranks) rows.
constexpr int kExchangeHeaderBytes = 256;
constexpr int kExchangeSlots = 4;
// Producer side when the records are pushed by a kernel other than the verify kernel: one CTA stores the
// `rows` local records packed[rows][width] (int32, already complete in local memory) into every rank's
// slot of the current epoch, one tagged 8-byte word per value.
SD_DEVINL void exchange_push_cta(const sd_exchange* x, const int32_t* packed, int rows) {
    char* own = reinterpret_cast<char*>(x->peer[x->rank]);
    unsigned epoch;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(epoch) : "l"(own) : "memory");
    const unsigned long long tag = (unsigned long long)(epoch + 1u) << 32;
    const int world = x->world, rpr = x->rows_per_rank, width = x->width;
    const size_t slot_base = ((size_t)(epoch & 3u) * world + x->rank) * rpr * width;
    const int n = rows * width;
    for (int i = threadIdx.x; i < n * world; i += blockDim.x) {
        const int pr = i / n, j = i - pr * n;
        unsigned long long* recs = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(x->peer[pr]) + 256);
        recs[slot_base + j] = tag | (unsigned)packed[j];
    }
}

// Gather, split in two so that a caller can put a wait between the data movement and the epoch advance:
// exchange_gather_records polls + copies the records of epoch (current - lag) and returns the current epoch;
// exchange_advance publishes epoch + 1.  Both are called by a whole CTA.
SD_DEVINL unsigned exchange_gather_records(const sd_exchange* x, int32_t* out) {
    char* own = reinterpret_cast<char*>(x->peer[x->rank]);
    unsigned epoch;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(epoch) : "l"(own) : "memory");
    const unsigned lag = (unsigned)x->lag;
    if (epoch >= lag) {
        const unsigned src = epoch - lag;
        const unsigned want = src + 1u;
        const int world = x->world, rpr = x->rows_per_rank, width = x->width;
        const int base = x->total_rows / world, extra = x->total_rows % world;
        const unsigned long long* words =
            reinterpret_cast<const unsigned long long*>(own + kExchangeHeaderBytes) + (size_t)(src & 3u) * world * rpr * width;
        const int n = world * rpr * width;
        // kU words per thread and pass (world = 8: a thread owns 12 words -> ONE pass).  Every poll round re-loads ALL words
        // of the pass whose tag is still stale, back to back, and only then looks at them: the gather ends one L2 round trip
        // after the last record has landed.  (Polled one word after the other -- the previous form -- a thread that owned
        // the records of the last sequences of four ranks needed four serial round trips after the last arrival: the
        // words a thread owns are 256 apart, so at world = 8 every third one falls into the same part of a rank's batch.)
        constexpr int kU = 12;
        for (int i0 = threadIdx.x; i0 < n; i0 += blockDim.x * kU) {
            unsigned long long w[kU];
            bool live[kU];
#pragma unroll
            for (int u = 0; u < kU; ++u) {
                const int i = i0 + u * blockDim.x;
                const int row = i / width;
                const int r = row / rpr;
                live[u] = i < n && (row - r * rpr) < base + (r < extra ? 1 : 0);  // not a padding row of an uneven shard
                w[u] = 0ull;  // (tag 0 is never wanted: want = epoch + 1 >= 1)
            }
            SpinGuard guard;
            for (;;) {
#pragma unroll
                for (int u = 0; u < kU; ++u)
                    if (live[u] && (unsigned)(w[u] >> 32) != want)
                        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w[u]) : "l"(words + i0 + u * blockDim.x) : "memory");
                bool stale = false;
#pragma unroll
                for (int u = 0; u < kU; ++u) stale |= live[u] && (unsigned)(w[u] >> 32) != want;
                if (!stale) break;
                if (!guard.keep_waiting(kSpinPeerWords)) break;
                __nanosleep(100);
            }
#pragma unroll
            for (int u = 0; u < kU; ++u)
                if (live[u] && out) out[i0 + u * blockDim.x] = (int)(unsigned)w[u];
        }
    }
    return epoch;
}

SD_DEVINL void exchange_advance(const sd_exchange* x, const unsigned epoch) {
    __syncthreads();
    if (threadIdx.x == 0) *reinterpret_cast<volatile unsigned*>(reinterpret_cast<char*>(x->peer[x->rank])) = epoch + 1u;
}

SD_DEVINL void exchange_gather_cta(const sd_exchange* x, int32_t* out) {
    const unsigned epoch = exchange_gather_records(x, out);
    exchange_advance(x, epoch);
}

// host side: OR of the spin sites that ran out of budget in THIS translation unit's kernels since the last clear
// (synchronises the device: not for the step path)
static inline int tu_spin_error(int clear, unsigned* out) {
    unsigned v = 0u;
    cudaError_t e = cudaMemcpyFromSymbol(&v, g_spin_error, sizeof(v));
    if (e == cudaSuccess && clear && v) {
        const unsigned z = 0u;
        e = cudaMemcpyToSymbol(g_spin_error, &z, sizeof(z));
    }
    *out = v;
    return e == cudaSuccess ? 0 : (int)e;
}
int verify_tu_spin_error(int clear, unsigned* out);    // sd_verify.cu
int kv_tu_spin_error(int clear, unsigned* out);        // sd_kv.cu
int exchange_tu_spin_error(int clear, unsigned* out);  // sd_exchange.cu

inline int elt_size(int dtype) { return dtype == SD_F32 ? 4 : 2; }

inline int to_status(cudaError_t e) { return e == cudaSuccess ? 0 : (int)e; }

}  // namespace sd
