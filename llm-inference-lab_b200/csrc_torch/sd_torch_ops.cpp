// Torch-side shim of the C ABI (include/sd_verify.h): the hot registry drop-ins as native code.
//
// Why: the reference calls verify_prefix once PER SEQUENCE (src/specdec/policies/policies.py:128-134 from
// src/specdec/core/pipeline.py:3018) and kv_append once PER LAYER (src/specdec/models/hf_wrappers.py:1009), so the
// host cost of a call matters more than the kernel.  Through Python + ctypes a call costs ~20 us of host time
// (tools/host_overhead.py); here the argument checks, the output allocation, the stream lookup and the C-ABI call are
// a few microseconds of C++.  The same functions are registered as torch custom ops (TORCH_LIBRARY_FRAGMENT sd_b200,
// CUDA dispatch key only: there is no CPU kernel to fall back to).
//
// This file contains NO kernels and no arithmetic: it only marshals at::Tensor arguments into the extern "C" entry
// points of libsd_verify_b200.so, to which it is linked.  A fast path returns an empty optional when a call is not
// eligible (wrong device / dtype / layout): the Python caller (kernels/cuda_ops.py) then takes its own path, which
// raises the reference's errors.
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>
#include <torch/extension.h>
#include <torch/library.h>

#include <mutex>
#include <optional>
#include <tuple>
#include <unordered_map>

#include "sd_verify.h"

namespace {

int dtype_code(c10::ScalarType t) {
    switch (t) {
        case at::kFloat: return SD_F32;
        case at::kHalf: return SD_F16;
        case at::kBFloat16: return SD_BF16;
        default: return -1;
    }
}

void check_rc(int rc, const char* what) {
    TORCH_CHECK(rc == 0, what, ": ", rc, " (", sd_error_string(rc), ")");
}

// ---- per (device, stream) scratch: verify workspace (zeroed once; the kernels leave it zeroed) and kv ticket ----
struct Scratch {
    at::Tensor workspace, ticket;
};
std::mutex g_mu;
std::unordered_map<uint64_t, Scratch> g_scratch;

Scratch& scratch_for(const c10::Device& dev, cudaStream_t st, size_t need_ws) {
    const uint64_t key = (reinterpret_cast<uint64_t>(st) << 6) ^ (uint64_t)(dev.index() & 63);
    std::lock_guard<std::mutex> lock(g_mu);
    Scratch& s = g_scratch[key];
    if (!s.ticket.defined()) s.ticket = at::zeros({1}, at::TensorOptions().dtype(at::kInt).device(dev));
    if (need_ws && (!s.workspace.defined() || (size_t)s.workspace.numel() < need_ws))
        s.workspace = at::zeros({(int64_t)std::max<size_t>(need_ws, 1u << 20)}, at::TensorOptions().dtype(at::kByte).device(dev));
    return s;
}

bool capturing(cudaStream_t st) {
    cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
    return cudaStreamIsCapturing(st, &status) == cudaSuccess && status != cudaStreamCaptureStatusNone;
}

// ---- verify_prefix(logits [B,K,V], draft_ids [B,K]) -> (accept_len [B] i32, accepted_mask [B,K] u8) ----
// reference: src/kernels/reference.py:13-56
bool verify_eligible(const at::Tensor& logits, const at::Tensor& ids, int rows_extra) {
    return logits.is_cuda() && ids.is_cuda() && logits.device() == ids.device() && logits.dim() == 3 && ids.dim() == 2 &&
           logits.size(0) == ids.size(0) && logits.size(1) == ids.size(1) + rows_extra && dtype_code(logits.scalar_type()) >= 0 &&
           (ids.scalar_type() == at::kLong || ids.scalar_type() == at::kInt) && logits.size(2) > 0 && logits.stride(2) == 1 &&
           (ids.size(1) == 0 || ids.stride(1) == 1) && logits.size(0) > 0 && ids.size(1) > 0;
}

std::optional<std::tuple<at::Tensor, at::Tensor>> verify_prefix_fast(const at::Tensor& logits, const at::Tensor& ids) {
    if (!verify_eligible(logits, ids, 0)) return std::nullopt;
    const int64_t B = logits.size(0), K = logits.size(1), V = logits.size(2);
    c10::cuda::OptionalCUDAGuard guard;
    if (logits.get_device() != c10::cuda::current_device()) guard.set_device(logits.device());
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(logits.get_device()).stream();
    auto i32 = at::TensorOptions().dtype(at::kInt).device(logits.device());
    at::Tensor accept_len = at::empty({B}, i32);
    at::Tensor mask = at::empty({B, K}, i32.dtype(at::kByte));
    const size_t need = sd_verify_workspace_bytes((int)B, (int)K);
    at::Tensor ws = capturing(st) ? at::zeros({(int64_t)need}, i32.dtype(at::kByte)) : scratch_for(logits.device(), st, need).workspace;
    const int rc = sd_verify_prefix(logits.data_ptr(), dtype_code(logits.scalar_type()), logits.stride(0), logits.stride(1),
                                    ids.data_ptr(), ids.scalar_type() == at::kLong ? SD_I64 : SD_I32, ids.stride(0),
                                    accept_len.data_ptr<int32_t>(), mask.data_ptr<uint8_t>(), nullptr, nullptr, nullptr,
                                    ws.data_ptr(), (size_t)ws.numel(), (int)B, (int)K, (int)K, (int)V, st, nullptr, nullptr);
    check_rc(rc, "sd_verify_prefix");
    return std::make_tuple(accept_len, mask);
}

// verify_prefix_bonus(logits [B,K+1,V], ids [B,K]) -> (accept_len, mask, target_ids [B,K+1], bonus [B], packed [B,K+2])
std::optional<std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor>> verify_prefix_bonus_fast(const at::Tensor& logits,
                                                                                                            const at::Tensor& ids) {
    if (!verify_eligible(logits, ids, 1)) return std::nullopt;
    const int64_t B = logits.size(0), rows = logits.size(1), K = rows - 1, V = logits.size(2);
    c10::cuda::OptionalCUDAGuard guard;
    if (logits.get_device() != c10::cuda::current_device()) guard.set_device(logits.device());
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(logits.get_device()).stream();
    auto i32 = at::TensorOptions().dtype(at::kInt).device(logits.device());
    at::Tensor accept_len = at::empty({B}, i32), mask = at::empty({B, K}, i32.dtype(at::kByte)), target = at::empty({B, rows}, i32),
               bonus = at::empty({B}, i32), packed = at::empty({B, K + 2}, i32);
    const size_t need = sd_verify_workspace_bytes((int)B, (int)rows);
    at::Tensor ws = capturing(st) ? at::zeros({(int64_t)need}, i32.dtype(at::kByte)) : scratch_for(logits.device(), st, need).workspace;
    const int rc = sd_verify_prefix(logits.data_ptr(), dtype_code(logits.scalar_type()), logits.stride(0), logits.stride(1),
                                    ids.data_ptr(), ids.scalar_type() == at::kLong ? SD_I64 : SD_I32, ids.stride(0),
                                    accept_len.data_ptr<int32_t>(), mask.data_ptr<uint8_t>(), target.data_ptr<int32_t>(),
                                    bonus.data_ptr<int32_t>(), packed.data_ptr<int32_t>(), ws.data_ptr(), (size_t)ws.numel(), (int)B,
                                    (int)rows, (int)K, (int)V, st, nullptr, nullptr);
    check_rc(rc, "sd_verify_prefix");
    return std::make_tuple(accept_len, mask, target, bonus, packed);
}

// ---- kv_append: rows of `new` written behind the first base_len rows of a [B,H,C,D] arena (in place, out-variant) ----
// reference: src/kernels/reference.py:59-159.  The functional drop-in (arena growth, adoption of a foreign base)
// lives in kernels/cuda_ops.py; this is the launch it issues per layer.
void kv_append_rows_(at::Tensor arena_k, at::Tensor arena_v, const at::Tensor& new_k, const at::Tensor& new_v, int64_t base_len,
                     const std::optional<at::Tensor>& n_keep, int64_t keep_bias, const std::optional<at::Tensor>& mask,
                     bool zero_rejected) {
    TORCH_CHECK(arena_k.is_cuda() && arena_v.is_cuda() && new_k.is_cuda() && new_v.is_cuda(), "kv_append_rows_: CUDA tensors only");
    TORCH_CHECK(arena_k.dim() == 4 && new_k.dim() == 4 && arena_k.sizes() == arena_v.sizes() && new_k.sizes() == new_v.sizes(),
                "kv_append_rows_: [B,H,C,D] arenas and [B,H,R,D] rows");
    TORCH_CHECK(arena_k.size(0) == new_k.size(0) && arena_k.size(1) == new_k.size(1) && arena_k.size(3) == new_k.size(3), "B / H / D mismatch");
    TORCH_CHECK(arena_k.scalar_type() == new_k.scalar_type() && arena_v.scalar_type() == new_k.scalar_type() &&
                new_v.scalar_type() == new_k.scalar_type(), "dtype mismatch");
    TORCH_CHECK(arena_k.stride(3) == 1 && arena_k.stride(2) == arena_k.size(3) && arena_k.strides() == arena_v.strides(), "arena layout");
    TORCH_CHECK(new_k.size(3) == 0 || (new_k.stride(3) == 1 && new_k.strides() == new_v.strides()), "rows layout");
    const int64_t B = arena_k.size(0), H = arena_k.size(1), C = arena_k.size(2), D = arena_k.size(3), R = new_k.size(2);
    if (B * H * D * R == 0) return;
    const int64_t es = arena_k.element_size();
    c10::cuda::OptionalCUDAGuard guard;
    if (arena_k.get_device() != c10::cuda::current_device()) guard.set_device(arena_k.device());
    cudaStream_t st = c10::cuda::getCurrentCUDAStream(arena_k.get_device()).stream();
    at::Tensor ticket = capturing(st) ? at::zeros({1}, at::TensorOptions().dtype(at::kInt).device(arena_k.device()))
                                      : scratch_for(arena_k.device(), st, 0).ticket;
    const int rc = sd_kv_append_layer(arena_k.data_ptr(), arena_v.data_ptr(), new_k.data_ptr(), new_v.data_ptr(), (int)B, (int)H, (int)R,
                                      D * es, C, arena_k.stride(0) * es, arena_k.stride(1) * es, new_k.stride(0) * es,
                                      new_k.stride(1) * es, new_k.stride(2) * es, base_len,
                                      n_keep.has_value() ? n_keep->data_ptr<int32_t>() : nullptr, (int)keep_bias,
                                      mask.has_value() ? mask->data_ptr<uint8_t>() : nullptr, zero_rejected ? 1 : 0,
                                      ticket.data_ptr<int32_t>(), st);
    check_rc(rc, "sd_kv_append_layer");
}

// Arena bookkeeping of the functional drop-in: which storages are arenas of ours, how many rows are valid.  Keyed by
// the StorageImpl; a weak reference detects a freed (and possibly re-used) address.
struct ArenaInfo {
    c10::weak_intrusive_ptr<c10::StorageImpl> storage;
    int64_t used, capacity;
};
std::unordered_map<const c10::StorageImpl*, ArenaInfo> g_arenas;

void arena_register(const at::Tensor& arena, int64_t used) {
    std::lock_guard<std::mutex> lock(g_mu);
    if (g_arenas.size() > 4096) {  // drop entries whose storage is gone
        for (auto it = g_arenas.begin(); it != g_arenas.end();)
            it = it->second.storage.expired() ? g_arenas.erase(it) : std::next(it);
    }
    c10::StorageImpl* impl = arena.storage().unsafeGetStorageImpl();
    g_arenas.insert_or_assign(impl, ArenaInfo{c10::weak_intrusive_ptr<c10::StorageImpl>(
                                                   c10::intrusive_ptr<c10::StorageImpl>::reclaim_copy(impl)),
                                               used, arena.size(2)});
}

// (used, capacity) of the arena behind `t`, or (-1, -1)
std::tuple<int64_t, int64_t> arena_info(const at::Tensor& t) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_arenas.find(t.storage().unsafeGetStorageImpl());
    if (it == g_arenas.end()) return {-1, -1};
    if (it->second.storage.expired()) {
        g_arenas.erase(it);
        return {-1, -1};
    }
    return {it->second.used, it->second.capacity};
}

// `t` [B,H,L,D] is the [:, :, :L] prefix view of one of our arenas whose tip is at L, with room for R more rows:
// returns the capacity, else -1.
int64_t tip_capacity(const at::Tensor& t, int64_t R) {
    if (t.dim() != 4 || t.storage_offset() != 0) return -1;
    const auto [used, cap] = arena_info(t);
    if (cap < 0 || used != t.size(2) || used + R > cap) return -1;
    const int64_t H = t.size(1), D = t.size(3);
    if (t.numel() > 0 && (t.stride(3) != 1 || t.stride(2) != D || t.stride(1) != cap * D || t.stride(0) != H * cap * D)) return -1;
    if ((int64_t)(t.storage().nbytes() / t.element_size()) != t.size(0) * H * cap * D) return -1;
    return cap;
}

at::Tensor arena_view(const at::Tensor& like, int64_t cap, int64_t rows) {
    at::Tensor v = at::empty({0}, like.options());
    const int64_t B = like.size(0), H = like.size(1), D = like.size(3);
    v.set_(like.storage(), 0, {B, H, rows, D}, {H * cap * D, cap * D, D, 1});
    return v;
}

// kv_append(base_k, base_v, new_k, new_v) when base_* are the tips of two of our arenas: O(Kn) bytes, results are views
// of the arenas.  Empty optional = not eligible (the Python drop-in adopts the base into a fresh arena).
std::optional<std::tuple<at::Tensor, at::Tensor>> kv_append_fast(const at::Tensor& base_k, const at::Tensor& base_v, const at::Tensor& new_k,
                                                                  const at::Tensor& new_v) {
    if (!(base_k.is_cuda() && base_v.is_cuda() && new_k.is_cuda() && new_v.is_cuda())) return std::nullopt;
    if (base_k.dim() != 4 || new_k.dim() != 4 || base_k.sizes() != base_v.sizes() || new_k.sizes() != new_v.sizes()) return std::nullopt;
    if (base_k.size(0) != new_k.size(0) || base_k.size(1) != new_k.size(1) || base_k.size(3) != new_k.size(3)) return std::nullopt;
    const auto dt = base_k.scalar_type();
    if (base_v.scalar_type() != dt || new_k.scalar_type() != dt || new_v.scalar_type() != dt) return std::nullopt;
    const int64_t L = base_k.size(2), R = new_k.size(2);
    if (R == 0 || base_k.numel() == 0 || new_k.stride(3) != 1 || new_k.strides() != new_v.strides()) return std::nullopt;
    const int64_t cap = tip_capacity(base_k, R);
    if (cap < 0 || tip_capacity(base_v, R) != cap) return std::nullopt;
    at::Tensor ak = arena_view(base_k, cap, cap), av = arena_view(base_v, cap, cap);
    kv_append_rows_(ak, av, new_k, new_v, L, std::nullopt, 0, std::nullopt, false);
    {
        std::lock_guard<std::mutex> lock(g_mu);
        for (const at::Tensor* t : {&base_k, &base_v}) {
            auto it = g_arenas.find(t->storage().unsafeGetStorageImpl());
            if (it != g_arenas.end()) it->second.used = L + R;
        }
    }
    return std::make_tuple(arena_view(base_k, cap, L + R), arena_view(base_v, cap, L + R));
}

void arena_set_used(const at::Tensor& arena, int64_t used) {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_arenas.find(arena.storage().unsafeGetStorageImpl());
    if (it != g_arenas.end()) it->second.used = used;
}

// ---- dispatcher ops (non-optional signatures) ----
std::tuple<at::Tensor, at::Tensor> op_verify_prefix(const at::Tensor& logits, const at::Tensor& ids) {
    auto r = verify_prefix_fast(logits, ids.scalar_type() == at::kLong || ids.scalar_type() == at::kInt ? ids : ids.to(at::kLong));
    TORCH_CHECK(r.has_value(), "sd_b200::verify_prefix: logits [B,K,V] float32/float16/bfloat16 (unit stride on V) and draft_ids [B,K] on the same CUDA device");
    return *r;
}
std::tuple<at::Tensor, at::Tensor, at::Tensor, at::Tensor, at::Tensor> op_verify_prefix_bonus(const at::Tensor& logits, const at::Tensor& ids) {
    auto r = verify_prefix_bonus_fast(logits, ids.scalar_type() == at::kLong || ids.scalar_type() == at::kInt ? ids : ids.to(at::kLong));
    TORCH_CHECK(r.has_value(), "sd_b200::verify_prefix_bonus: logits [B,K+1,V] and draft_ids [B,K] on the same CUDA device");
    return *r;
}

}  // namespace

TORCH_LIBRARY_FRAGMENT(sd_b200, m) {
    m.def("verify_prefix(Tensor logits, Tensor draft_ids) -> (Tensor, Tensor)");
    m.def("verify_prefix_bonus(Tensor logits, Tensor draft_ids) -> (Tensor, Tensor, Tensor, Tensor, Tensor)");
    m.def("kv_append_rows_(Tensor(a!) arena_k, Tensor(b!) arena_v, Tensor new_k, Tensor new_v, int base_len, Tensor? n_keep, "
          "int keep_bias, Tensor? accepted_mask, bool zero_rejected) -> ()");
}

TORCH_LIBRARY_IMPL(sd_b200, CUDA, m) {
    m.impl("verify_prefix", &op_verify_prefix);
    m.impl("verify_prefix_bonus", &op_verify_prefix_bonus);
    m.impl("kv_append_rows_", &kv_append_rows_);
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
    m.doc() = "sd_b200 torch shim: native fast paths of the registry drop-ins over the C ABI of libsd_verify_b200.so";
    m.def("verify_prefix_fast", &verify_prefix_fast);
    m.def("verify_prefix_bonus_fast", &verify_prefix_bonus_fast);
    m.def("kv_append_fast", &kv_append_fast);
    m.def("kv_append_rows_", &kv_append_rows_);
    m.def("arena_register", &arena_register);
    m.def("arena_info", &arena_info);
    m.def("arena_set_used", &arena_set_used);
    m.def("abi_version", []() { return sd_abi_version(); });
}
