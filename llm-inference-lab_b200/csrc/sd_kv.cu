// Kernel 2: kv_append into a preallocated arena + rollback of rejected slots, all layers in one
// launch, driven by the accept lengths the verify kernel left on the device.
//
// Replaces torch.cat growth (reference: src/kernels/reference.py:59-93, O(L) bytes per layer per
// step), the masked compaction (reference.py:96-159; src/kernels/cuda/kv_cache.cu:14-95, which
// re-materialises the whole cache and returns zeros on zero acceptance) and the slice/clone/pad/stack
// rollback (src/specdec/cache/kv_cache_manager.py:353-479) by O(R) traffic per (layer, sequence, head).
//
// Two variants, selected by measurement (sd_kv_set_variant):
//   1  "vec":  one thread per 16-byte vector of an offered row, K and V loads both in flight,
//              ld.global.nc / st.global 128-bit.
//   3  "column" (default when rows are 16-byte aligned and unmasked): one thread per 16-byte column of
//              a (layer, b, h), all R rows of K and V in flight before the first store.
//   2  "bulk": one thread per (layer, K|V, b, h) unit issues ONE cp.async.bulk (TMA 1-D)
//              global->shared load of the accepted rows and ONE shared->global bulk store of the R
//              offered slots (accepted rows followed by zeros) -- no per-element instructions.
#include "sd_kv_col.cuh"

namespace sd {

static int g_kv_variant = 0;
static int g_kv_rows_in_flight = 0;  // column kernel: rows per load batch (0 = heuristic)

template <typename VecT>
__global__ void __launch_bounds__(256) kv_append_vec_kernel(const KvParams p) {
    pdl_wait();
    pdl_trigger();
    if (p.xchg && p.xchg_packed && blockIdx.x == 0 && blockIdx.y == 0) exchange_push_cta(p.xchg, p.xchg_packed, p.B);
    const int layer = blockIdx.y;
    const int vpr = (int)(p.row_bytes / (int64_t)sizeof(VecT));
    const int64_t slots = (int64_t)p.B * p.H * p.R * vpr;
    const int64_t slot = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot < slots) {
        const int v = (int)(slot % vpr);
        int64_t t = slot / vpr;
        const int j = (int)(t % p.R);
        t /= p.R;
        const int h = (int)(t % p.H);
        const int b = (int)(t / p.H);
        const sd_kv_layer L = p.layers ? p.layers[layer] : p.single;
        const int len = (p.seq_lens ? p.seq_lens[b] : 0) + (int)p.base_len;
        int n = keep_count(p, b);
        const int64_t dst_row = (int64_t)len + j;
        if (dst_row < p.capacity) {
            const int64_t dst_off = (int64_t)b * p.arena_stride_b + (int64_t)h * p.arena_stride_h +
                                    dst_row * p.row_bytes + (int64_t)v * (int64_t)sizeof(VecT);
            int src_row = j;
            if (p.mask && j < n) {
                src_row = masked_source_row(p.mask, b, p.R, j);
                if (src_row < 0) n = 0;  // fewer set bits than accept_len: nothing more to copy
            }
            if (j < n) {
                const int64_t src_off = (int64_t)b * p.new_stride_b + (int64_t)h * p.new_stride_h +
                                        (int64_t)src_row * p.new_stride_r + (int64_t)v * (int64_t)sizeof(VecT);
                VecT kx, vx;
                if constexpr (sizeof(VecT) == 16) {
                    kx = ldg_stream_v4(reinterpret_cast<const char*>(L.new_k) + src_off);
                    vx = ldg_stream_v4(reinterpret_cast<const char*>(L.new_v) + src_off);
                } else {
                    kx = *reinterpret_cast<const VecT*>(reinterpret_cast<const char*>(L.new_k) + src_off);
                    vx = *reinterpret_cast<const VecT*>(reinterpret_cast<const char*>(L.new_v) + src_off);
                }
                *reinterpret_cast<VecT*>(reinterpret_cast<char*>(L.arena_k) + dst_off) = kx;
                *reinterpret_cast<VecT*>(reinterpret_cast<char*>(L.arena_v) + dst_off) = vx;
            } else if (p.zero_rejected) {
                VecT z;
                memset(&z, 0, sizeof(VecT));
                *reinterpret_cast<VecT*>(reinterpret_cast<char*>(L.arena_k) + dst_off) = z;
                *reinterpret_cast<VecT*>(reinterpret_cast<char*>(L.arena_v) + dst_off) = z;
            }
        } else if (j < n && p.overflow) {
            *p.overflow = 1;
        }
    }
    // Fused result exchange (multi-GPU): CTA (0,0) -- which pushed this rank's records at its start when
    // exchange_packed is set -- gathers every rank's records of the epoch once its own rows are copied:
    // the NVLink round trips overlap the rest of the append instead of trailing it.
    if (p.xchg && blockIdx.x == 0 && blockIdx.y == 0) exchange_gather_cta(p.xchg, p.xchg_out);
    retire_and_commit(p);
}

// ---- variant 3 "column" (body in sd_kv_col.cuh) ----
template <int RB>
__global__ void __launch_bounds__(256) kv_append_col_kernel(const KvParams p) {
    pdl_trigger();
    kv_col_body<RB>(p, (int)blockIdx.x, (int)blockIdx.y);
}

// ---- cp.async.bulk (TMA 1-D) variant -------------------------------------------------------------
SD_DEVINL uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

SD_DEVINL void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
SD_DEVINL void mbar_arrive_expect_tx(uint64_t* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
SD_DEVINL void mbar_wait(uint64_t* bar, unsigned phase) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
SD_DEVINL void bulk_g2s(void* smem_dst, const void* gsrc, unsigned bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
SD_DEVINL void bulk_s2g(void* gdst, const void* smem_src, unsigned bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst),
                 "r"(smem_u32(smem_src)), "r"(bytes)
                 : "memory");
}

// One thread per unit (layer, K|V, b, h).  Dynamic smem: blockDim.x * R * row_bytes staging bytes.
__global__ void __launch_bounds__(128) kv_append_bulk_kernel(const KvParams p) {
    pdl_wait();
    pdl_trigger();
    if (p.xchg && p.xchg_packed && blockIdx.x == 0 && blockIdx.y == 0) exchange_push_cta(p.xchg, p.xchg_packed, p.B);
    extern __shared__ __align__(128) unsigned char stage[];
    __shared__ __align__(8) uint64_t bar;
    const int64_t units = (int64_t)p.n_layers * 2 * p.B * p.H;
    const int64_t unit = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned slot_bytes = (unsigned)(p.R * p.row_bytes);
    unsigned char* mine = stage + (size_t)threadIdx.x * slot_bytes;

    if (threadIdx.x == 0) {
        mbar_init(&bar, blockDim.x);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    bool active = unit < units;
    char* dst = nullptr;
    unsigned store_bytes = 0, load_bytes = 0;
    if (active) {
        int64_t t = unit;
        const int h = (int)(t % p.H);
        t /= p.H;
        const int b = (int)(t % p.B);
        t /= p.B;
        const int kv = (int)(t & 1);
        const int layer = (int)(t >> 1);
        const sd_kv_layer L = p.layers ? p.layers[layer] : p.single;
        const int len = (p.seq_lens ? p.seq_lens[b] : 0) + (int)p.base_len;
        int n = keep_count(p, b);
        int room = (int)(p.capacity - len);
        room = room < 0 ? 0 : (room > p.R ? p.R : room);
        if (n > room) {
            if (p.overflow) *p.overflow = 1;
            n = room;
        }
        const char* src = reinterpret_cast<const char*>(kv ? L.new_v : L.new_k) + (int64_t)b * p.new_stride_b +
                          (int64_t)h * p.new_stride_h;
        dst = reinterpret_cast<char*>(kv ? L.arena_v : L.arena_k) + (int64_t)b * p.arena_stride_b +
              (int64_t)h * p.arena_stride_h + (int64_t)len * p.row_bytes;
        load_bytes = (unsigned)(n * p.row_bytes);
        store_bytes = p.zero_rejected ? (unsigned)(room * p.row_bytes) : load_bytes;
        // zero the rejected tail of this thread's staging slot (generic proxy), then make it
        // visible to the async proxy that performs the bulk store
        for (unsigned o = load_bytes; o < store_bytes; o += 16) *reinterpret_cast<uint4*>(mine + o) = make_uint4(0, 0, 0, 0);
        mbar_arrive_expect_tx(&bar, load_bytes);
        if (load_bytes) bulk_g2s(mine, src, load_bytes, &bar);
    } else {
        mbar_arrive_expect_tx(&bar, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    mbar_wait(&bar, 0);
    if (active && store_bytes) bulk_s2g(dst, mine, store_bytes);
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    // Fused result exchange (multi-GPU): CTA (0,0) -- which pushed this rank's records at its start when
    // exchange_packed is set -- gathers every rank's records of the epoch once its own rows are copied:
    // the NVLink round trips overlap the rest of the append instead of trailing it.
    if (p.xchg && blockIdx.x == 0 && blockIdx.y == 0) exchange_gather_cta(p.xchg, p.xchg_out);
    retire_and_commit(p);
}

// Strides AND (where the host can see them) base pointers: `host_layer` is the by-value record of the single-layer
// form.  The pointers of a device table cannot be read here without a synchronising copy: for sd_kv_append /
// sd_verify_append they are the caller's contract (16-byte aligned; the Python side checks every tensor it puts in a
// table -- KVArena._build_table, SafeKVCacheManager).  A misaligned 16-byte access is a sticky context error, so the
// single-layer form, which serves arbitrary views (a last-dim slice of a fused QKV buffer ...), drops to a narrower
// vector instead.
static bool aligned_to(const KvParams& p, const sd_kv_layer* host_layer, int64_t a) {
    if (p.row_bytes % a || p.arena_stride_b % a || p.arena_stride_h % a || p.new_stride_b % a ||
        p.new_stride_h % a || p.new_stride_r % a)
        return false;
    if (host_layer) {
        const uintptr_t bits = reinterpret_cast<uintptr_t>(host_layer->arena_k) | reinterpret_cast<uintptr_t>(host_layer->arena_v) |
                               reinterpret_cast<uintptr_t>(host_layer->new_k) | reinterpret_cast<uintptr_t>(host_layer->new_v);
        if (bits % (uintptr_t)a) return false;
    }
    return true;
}


int kv_prepare_col(const sd_kv_layer* layers, int n_layers, int B, int H, int R, int64_t row_bytes, int64_t capacity,
                   int64_t arena_stride_b, int64_t arena_stride_h, int64_t new_stride_b, int64_t new_stride_h,
                   int64_t new_stride_r, const int32_t* seq_lens, int32_t* seq_lens_out, const int32_t* n_keep,
                   int keep_bias, int zero_rejected, int64_t max_seq_len_host, int32_t* overflow_flag, int32_t* ticket,
                   const sd_exchange* exchange, int32_t* exchange_out, const int32_t* exchange_packed, int32_t* seq_ready,
                   KvParams* out, unsigned* grid_x, int* rb) {
    if (n_layers <= 0 || B <= 0 || H <= 0 || R <= 0 || row_bytes <= 0 || capacity < 0 || !layers || !ticket) return SD_ERR_INVALID_ARG;
    if (max_seq_len_host >= 0 && max_seq_len_host + R > capacity) return SD_ERR_CAPACITY;
    KvParams p;
    p.layers = layers;
    p.single = sd_kv_layer{nullptr, nullptr, nullptr, nullptr};
    p.base_len = 0;
    p.n_layers = n_layers;
    p.B = B;
    p.H = H;
    p.R = R;
    p.row_bytes = row_bytes;
    p.capacity = capacity;
    p.arena_stride_b = arena_stride_b;
    p.arena_stride_h = arena_stride_h;
    p.new_stride_b = new_stride_b;
    p.new_stride_h = new_stride_h;
    p.new_stride_r = new_stride_r;
    p.seq_lens = seq_lens;
    p.seq_lens_out = seq_lens_out;
    p.n_keep = n_keep;
    p.keep_bias = keep_bias;
    p.mask = nullptr;
    p.zero_rejected = zero_rejected;
    p.early_rows = 0;
    p.overflow = overflow_flag;
    p.ticket = ticket;
    p.xchg = exchange;
    p.xchg_out = exchange_out;
    p.xchg_packed = exchange_packed;
    p.seq_ready = seq_ready;
    p.fused = 1;
    p.flags = nullptr;
    p.epoch_word = nullptr;
    p.lag = p.pieces = p.piece_items = p.seq_items = p.tail_pieces = p.eager_rows = p.l2_prefetch = p.uniform = p.stagger_ns = p.one_warp_wait = 0;
    p.n_big = p.S_big = p.upc_big = p.bulk_split = p.fetch = p.tail_order = p.tail_peek = p.pf_ahead = p.n_verify = 0;
    p.carry_ni = kCarryNI;
    p.ls_arena = p.ls_new = 0;
    p.layer0 = sd_kv_layer{nullptr, nullptr, nullptr, nullptr};
    p.fd_vpr = p.fd_R = p.fd_H = make_fastdiv(1);
    if (!aligned_to(p, nullptr, 16)) return SD_ERR_UNSUPPORTED;
    const int64_t slots = (int64_t)B * H * (row_bytes / 16);
    *grid_x = (unsigned)((slots + 255) / 256);
    p.total_ctas = *grid_x * (unsigned)n_layers;
    int r = R <= 6 ? R : ((R + 1) / 2 <= 6 ? (R + 1) / 2 : 6);
    *rb = r;
    *out = p;
    return SD_OK;
}

int kv_tu_spin_error(int clear, unsigned* out) { return tu_spin_error(clear, out); }

}  // namespace sd

extern "C" {

int sd_kv_set_variant(int variant) {
    // variant 3 + 16 * rows_in_flight selects the column kernel's load batch (evidence runs)
    const int rif = variant >> 4;
    variant &= 15;
    if (variant < 0 || variant > 3 || rif < 0 || rif > 6) return SD_ERR_INVALID_ARG;
    sd::g_kv_variant = variant;
    sd::g_kv_rows_in_flight = rif;
    return SD_OK;
}

static int kv_append_impl(const sd_kv_layer* layers, const sd_kv_layer* single, int64_t base_len, int n_layers,
                          int B, int H, int R, int64_t row_bytes, int64_t capacity, int64_t arena_stride_b,
                          int64_t arena_stride_h, int64_t new_stride_b, int64_t new_stride_h,
                          int64_t new_stride_r, const int32_t* seq_lens, int32_t* seq_lens_out,
                          const int32_t* n_keep, int keep_bias, const uint8_t* accepted_mask, int zero_rejected,
                          int64_t max_seq_len_host, int32_t* overflow_flag, int32_t* ticket, int early_rows, void* stream,
                          const sd_exchange* exchange = nullptr, int32_t* exchange_out = nullptr,
                          const int32_t* exchange_packed = nullptr, int32_t* seq_ready = nullptr) {
    using namespace sd;
    if (n_layers < 0 || B < 0 || H < 0 || R < 0 || row_bytes <= 0 || capacity < 0) return SD_ERR_INVALID_ARG;
    if (n_layers == 0 || B == 0 || H == 0) return SD_OK;
    if ((!layers && !single) || !ticket) return SD_ERR_INVALID_ARG;
    if (accepted_mask && R > 64) return SD_ERR_UNSUPPORTED;
    if (max_seq_len_host >= 0 && max_seq_len_host + R > capacity) return SD_ERR_CAPACITY;
    if (row_bytes % 2) return SD_ERR_ALIGNMENT;

    KvParams p;
    p.layers = layers;
    if (single) p.single = *single;
    else p.single = sd_kv_layer{nullptr, nullptr, nullptr, nullptr};
    p.base_len = base_len;
    p.n_layers = n_layers;
    p.B = B;
    p.H = H;
    p.R = R;
    p.row_bytes = row_bytes;
    p.capacity = capacity;
    p.arena_stride_b = arena_stride_b;
    p.arena_stride_h = arena_stride_h;
    p.new_stride_b = new_stride_b;
    p.new_stride_h = new_stride_h;
    p.new_stride_r = new_stride_r;
    p.seq_lens = seq_lens;
    p.seq_lens_out = seq_lens_out;
    p.n_keep = n_keep;
    p.keep_bias = keep_bias;
    p.mask = accepted_mask;
    p.zero_rejected = zero_rejected;
    p.early_rows = early_rows ? 1 : 0;
    p.overflow = overflow_flag;
    p.ticket = ticket;
    p.xchg = exchange;
    p.xchg_out = exchange_out;
    p.xchg_packed = exchange_packed;
    p.seq_ready = seq_ready;
    p.fused = 0;
    p.flags = nullptr;
    p.epoch_word = nullptr;
    p.lag = p.pieces = p.piece_items = p.seq_items = p.tail_pieces = p.eager_rows = p.l2_prefetch = p.uniform = p.stagger_ns = p.one_warp_wait = 0;
    p.n_big = p.S_big = p.upc_big = p.bulk_split = p.fetch = p.tail_order = p.tail_peek = p.pf_ahead = p.n_verify = 0;
    p.carry_ni = kCarryNI;
    p.ls_arena = p.ls_new = 0;
    p.layer0 = sd_kv_layer{nullptr, nullptr, nullptr, nullptr};
    p.fd_vpr = p.fd_R = p.fd_H = make_fastdiv(1);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);

    if (R == 0) {  // nothing offered: only the length commit (identity) is required
        if (seq_lens_out && seq_lens && seq_lens_out != seq_lens)
            return to_status(cudaMemcpyAsync(seq_lens_out, seq_lens, (size_t)B * 4, cudaMemcpyDeviceToDevice, st));
        return SD_OK;
    }

    // Strides are host-known; so are the pointers of the single-layer form (checked in aligned_to).
    const bool bulk_ok = !accepted_mask && aligned_to(p, single, 16) && new_stride_r == row_bytes &&
                         (size_t)R * row_bytes * 32 <= 200 * 1024;
    const bool col_ok = !accepted_mask && aligned_to(p, single, 16);
    int variant = g_kv_variant;
    if (variant == 0) variant = col_ok ? 3 : 1;
    if (variant == 2 && !bulk_ok) variant = 1;
    if (variant == 3 && !col_ok) variant = 1;
    if (seq_ready && (variant != 3 || !n_keep)) return SD_ERR_UNSUPPORTED;  // the per-sequence wait lives in the column kernel

    if (variant == 3) {
        const int64_t slots = (int64_t)B * H * (row_bytes / 16);
        const int threads = 256;
        dim3 grid((unsigned)((slots + threads - 1) / threads), (unsigned)n_layers);
        p.total_ctas = grid.x * grid.y;
        int rb = g_kv_rows_in_flight > 0 ? g_kv_rows_in_flight : (R <= 6 ? R : (R + 1) / 2 <= 6 ? (R + 1) / 2 : 6);
        if (rb > R) rb = R;
        switch (rb) {
            case 1: return to_status(launch_kernel(kv_append_col_kernel<1>, grid, dim3(threads), 0, st, p));
            case 2: return to_status(launch_kernel(kv_append_col_kernel<2>, grid, dim3(threads), 0, st, p));
            case 3: return to_status(launch_kernel(kv_append_col_kernel<3>, grid, dim3(threads), 0, st, p));
            case 4: return to_status(launch_kernel(kv_append_col_kernel<4>, grid, dim3(threads), 0, st, p));
            case 5: return to_status(launch_kernel(kv_append_col_kernel<5>, grid, dim3(threads), 0, st, p));
            default: return to_status(launch_kernel(kv_append_col_kernel<6>, grid, dim3(threads), 0, st, p));
        }
    }

    if (variant == 2) {
        const int64_t units = (int64_t)n_layers * 2 * B * H;
        int threads = 128;
        while (threads > 32 && (size_t)threads * R * row_bytes > 100 * 1024) threads >>= 1;
        const size_t smem = (size_t)threads * R * row_bytes;
        unsigned ctas = (unsigned)((units + threads - 1) / threads);
        p.total_ctas = ctas;
        cudaError_t e = cudaFuncSetAttribute(kv_append_bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return to_status(e);
        return to_status(launch_kernel(kv_append_bulk_kernel, dim3(ctas), dim3(threads), smem, st, p));
    }

    int vec = 16;
    while (vec > 2 && !aligned_to(p, single, vec)) vec >>= 1;
    if (!aligned_to(p, single, vec)) return SD_ERR_ALIGNMENT;  // not even element-aligned
    const int64_t slots = (int64_t)B * H * R * (row_bytes / vec);
    const int threads = 256;
    dim3 grid((unsigned)((slots + threads - 1) / threads), (unsigned)n_layers);
    p.total_ctas = grid.x * grid.y;
    switch (vec) {
        case 16: return to_status(launch_kernel(kv_append_vec_kernel<uint4>, grid, dim3(threads), 0, st, p));
        case 8: return to_status(launch_kernel(kv_append_vec_kernel<uint2>, grid, dim3(threads), 0, st, p));
        case 4: return to_status(launch_kernel(kv_append_vec_kernel<uint32_t>, grid, dim3(threads), 0, st, p));
        default: return to_status(launch_kernel(kv_append_vec_kernel<uint16_t>, grid, dim3(threads), 0, st, p));
    }
}

int sd_kv_append(const sd_kv_layer* layers, int n_layers, int B, int H, int R, int64_t row_bytes,
                 int64_t capacity, int64_t arena_stride_b, int64_t arena_stride_h, int64_t new_stride_b,
                 int64_t new_stride_h, int64_t new_stride_r, const int32_t* seq_lens, int32_t* seq_lens_out,
                 const int32_t* n_keep, int keep_bias, const uint8_t* accepted_mask, int zero_rejected,
                 int64_t max_seq_len_host, int32_t* overflow_flag, int32_t* ticket, int rows_ready_early, void* stream,
                 const sd_exchange* exchange, int32_t* exchange_out, const int32_t* exchange_packed,
                 int32_t* seq_ready) {
    if (n_layers > 0 && !layers) return SD_ERR_INVALID_ARG;
    if (exchange && (n_layers == 0 || B == 0 || H == 0 || R == 0)) return SD_ERR_UNSUPPORTED;  // nothing would launch
    return kv_append_impl(layers, nullptr, 0, n_layers, B, H, R, row_bytes, capacity, arena_stride_b, arena_stride_h,
                          new_stride_b, new_stride_h, new_stride_r, seq_lens, seq_lens_out, n_keep, keep_bias,
                          accepted_mask, zero_rejected, max_seq_len_host, overflow_flag, ticket, rows_ready_early, stream,
                          exchange, exchange_out, exchange_packed, seq_ready);
}

int sd_kv_append_layer(void* arena_k, void* arena_v, const void* new_k, const void* new_v, int B, int H, int R,
                       int64_t row_bytes, int64_t capacity, int64_t arena_stride_b, int64_t arena_stride_h,
                       int64_t new_stride_b, int64_t new_stride_h, int64_t new_stride_r, int64_t base_len,
                       const int32_t* n_keep, int keep_bias, const uint8_t* accepted_mask, int zero_rejected,
                       int32_t* ticket, void* stream) {
    if (!arena_k || !arena_v || ((!new_k || !new_v) && R > 0) || base_len < 0) return SD_ERR_INVALID_ARG;
    sd_kv_layer one{arena_k, arena_v, new_k, new_v};
    return kv_append_impl(nullptr, &one, base_len, 1, B, H, R, row_bytes, capacity, arena_stride_b, arena_stride_h,
                          new_stride_b, new_stride_h, new_stride_r, nullptr, nullptr, n_keep, keep_bias,
                          accepted_mask, zero_rejected, base_len, nullptr, ticket, 0, stream);
}

}  // extern "C"
