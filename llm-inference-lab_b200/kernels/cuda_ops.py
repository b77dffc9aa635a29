"""Host side of the B200 verify path: torch tensors in, C-ABI launches out.

Everything here is plumbing around ``libsd_verify_b200.so`` (include/sd_verify.h): argument
checking with the reference's error behaviour, output allocation with the torch allocator, and
launching on torch's CURRENT stream (the reference's own .cu files launched on the legacy default
stream -- src/kernels/cuda/verify.cu:196 -- which breaks ``torch.cuda.stream`` callers).  No function
in this module synchronises the device or falls back to a CPU / PyTorch implementation: inputs that
are not CUDA tensors raise.

The functions registered in the kernel registry are the ``*_cuda`` ones; their ``__name__`` contains
"cuda" because ``get_kernel_info`` classifies backends by name
(reference: src/kernels/__init__.py:129-147).  The same entry points are also exposed as torch custom
ops under the ``sd_b200::`` namespace (see ``_register_torch_ops``).
"""

from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import _capi
from . import build as _build

_DTYPE = {torch.float32: _capi.DTYPE_F32, torch.float16: _capi.DTYPE_F16, torch.bfloat16: _capi.DTYPE_BF16}
_ID_DTYPE = {torch.int32: _capi.ID_I32, torch.int64: _capi.ID_I64}

ACCEPT_MODES = {"greedy": _capi.ACCEPT_GREEDY, "typical": _capi.ACCEPT_TYPICAL, "confidence": _capi.ACCEPT_CONFIDENCE,
                "topk": _capi.ACCEPT_TOPK}


def _require_cuda(t: torch.Tensor, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(
            f"{what} must be a CUDA tensor: the sd_b200 kernels have no CPU implementation "
            f"(got device {t.device})"
        )


def _stream_ptr(device: torch.device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class _OnDevice:
    """``with _OnDevice(dev):`` == ``with torch.cuda.device(dev):`` but free when ``dev`` already is the current
    device (the common case; torch's context manager costs ~4 us of host time per call, which matters for a
    drop-in that the reference calls once per sequence / once per layer)."""

    __slots__ = ("ctx",)

    def __init__(self, device: torch.device) -> None:
        idx = device.index
        self.ctx = None if (idx is None or idx == torch.cuda.current_device()) else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            return self.ctx.__exit__(*exc)
        return False


class _Scratch:
    """Per (device, stream) scratch: verify workspaces and kv tickets.

    A workspace must not be shared by launches that can overlap, so the cache key includes the
    stream.  While a CUDA graph is being captured nothing is cached: the allocation belongs to the
    graph's private pool and stays alive with the graph."""

    def __init__(self) -> None:
        self._ws: Dict[Tuple[int, int, int], torch.Tensor] = {}
        self._ticket: Dict[Tuple[int, int], torch.Tensor] = {}

    def workspace(self, device: torch.device, B: int, rows: int, stream_ptr: Optional[int] = None) -> torch.Tensor:
        need = B * rows * 1024 + 256 if B > 0 and rows > 0 else 256  # == sd_verify_workspace_bytes (checked by the library)
        if torch.cuda.is_current_stream_capturing():
            return torch.zeros(need, dtype=torch.uint8, device=device)
        key = (device.index or 0, stream_ptr if stream_ptr is not None else _stream_ptr(device), 0)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.zeros(max(need, 1 << 20), dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def attn_workspace(self, device: torch.device, need: int) -> torch.Tensor:
        if torch.cuda.is_current_stream_capturing():
            return torch.empty(need, dtype=torch.uint8, device=device)
        key = (device.index or 0, _stream_ptr(device), 1)
        ws = self._ws.get(key)
        if ws is None or ws.numel() < need:
            ws = torch.empty(need, dtype=torch.uint8, device=device)
            self._ws[key] = ws
        return ws

    def ticket(self, device: torch.device, stream_ptr: Optional[int] = None) -> torch.Tensor:
        if torch.cuda.is_current_stream_capturing():
            return torch.zeros(1, dtype=torch.int32, device=device)
        key = (device.index or 0, stream_ptr if stream_ptr is not None else _stream_ptr(device))
        t = self._ticket.get(key)
        if t is None:
            t = torch.zeros(1, dtype=torch.int32, device=device)
            self._ticket[key] = t
        return t


_scratch = _Scratch()

_shim_mod = None


def _shim():
    """The native torch shim (csrc_torch/sd_torch_ops.cpp): fast paths of the registry drop-ins + the arena registry."""
    global _shim_mod
    if _shim_mod is None:
        _shim_mod = _build.load_torch_shim()
    return _shim_mod


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


# ------------------------------------------------------------------------------------------ verify
def _check_verify_inputs(logits: torch.Tensor, draft_ids: torch.Tensor, rows_extra: int):
    # same assertions as the reference (src/kernels/reference.py:27-31) -> AssertionError
    ls, ds = logits.shape, draft_ids.shape
    assert len(ls) == 3, "logits must be 3D tensor [B][K][V]"
    assert len(ds) == 2, "draft_ids must be 2D tensor [B][K]"
    assert ls[0] == ds[0], "Batch size mismatch"
    assert ls[1] == ds[1] + rows_extra, "K dimension mismatch"
    if not logits.is_cuda:
        _require_cuda(logits, "logits")
    if logits.dtype not in _DTYPE:
        raise TypeError(f"logits dtype {logits.dtype} not supported (float32 / float16 / bfloat16)")
    if draft_ids.device != logits.device:
        draft_ids = draft_ids.to(logits.device)
    if draft_ids.dtype not in _ID_DTYPE:
        draft_ids = draft_ids.to(torch.int64)
    if ls[2] > 0 and logits.stride(2) != 1:
        logits = logits.contiguous()
    if ds[1] > 0 and draft_ids.stride(1) != 1:
        draft_ids = draft_ids.contiguous()
    return logits, draft_ids


def verify_launch(
    logits: torch.Tensor,
    draft_ids: torch.Tensor,
    accept_len: torch.Tensor,
    accepted_mask: Optional[torch.Tensor] = None,
    target_ids: Optional[torch.Tensor] = None,
    bonus: Optional[torch.Tensor] = None,
    packed: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
    *,
    mode: int = _capi.ACCEPT_GREEDY,
    threshold: float = 0.0,
    inv_temperature: float = 1.0,
    token_prob: Optional[torch.Tensor] = None,
    row_max: Optional[torch.Tensor] = None,
    row_sumexp: Optional[torch.Tensor] = None,
    stats: bool = False,
    exchange: Optional[int] = None,
    seq_ready: Optional[torch.Tensor] = None,
) -> None:
    """Enqueue one fused-verify launch into caller-owned buffers (no allocation when a workspace is
    passed).  This is the call the graph-captured step uses.  ``exchange`` = device pointer of an
    ``sd_exchange`` descriptor (scheduler.batch_shard.PeerExchange.desc_ptr) for the fused multi-GPU
    result exchange."""
    B, rows, V = logits.shape
    K = draft_ids.shape[1]
    dev = logits.device
    sp = _stream_ptr(dev)
    if workspace is None:
        workspace = _scratch.workspace(dev, B, rows, sp)
    lib = _capi.lib()
    common = (
        logits.data_ptr(), _DTYPE[logits.dtype], logits.stride(0), logits.stride(1),
        draft_ids.data_ptr(), _ID_DTYPE[draft_ids.dtype], draft_ids.stride(0),
    )
    with _OnDevice(dev):
        if stats or mode != _capi.ACCEPT_GREEDY:
            rc = lib.sd_verify_step(
                *common, mode, float(threshold), float(inv_temperature),
                accept_len.data_ptr(), _ptr(accepted_mask), _ptr(target_ids), _ptr(bonus), _ptr(packed),
                _ptr(token_prob), _ptr(row_max), _ptr(row_sumexp),
                workspace.data_ptr(), workspace.numel(), B, rows, K, V, sp, exchange, _ptr(seq_ready),
            )
        else:
            rc = lib.sd_verify_prefix(
                *common, accept_len.data_ptr(), _ptr(accepted_mask), _ptr(target_ids), _ptr(bonus), _ptr(packed),
                workspace.data_ptr(), workspace.numel(), B, rows, K, V, sp, exchange, _ptr(seq_ready),
            )
    if rc:
        lib.check(rc, "sd_verify")


def verify_prefix_cuda(logits: torch.Tensor, draft_ids: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Drop-in for the registry op ``verify_prefix``.

    reference: src/kernels/reference.py:13-56 (semantics), src/kernels/triton/verify.py:101-169
    (the incumbent it outranks), call site src/specdec/policies/policies.py:128-134.
    logits [B,K,V] fp32/fp16/bf16 (batch slices welcome), draft_ids [B,K] int -> (accept_len [B]
    int32, accepted_mask [B,K] uint8) on logits.device, enqueued on the current stream, no sync.
    The call is served natively (csrc_torch/sd_torch_ops.cpp) whenever the arguments are eligible; anything else takes
    the path below, which raises the reference's errors."""
    fast = _shim().verify_prefix_fast(logits, draft_ids)
    if fast is not None:
        return fast
    logits, draft_ids = _check_verify_inputs(logits, draft_ids, 0)
    B, K, _ = logits.shape
    accept_len = torch.empty(B, dtype=torch.int32, device=logits.device)
    mask = torch.empty((B, K), dtype=torch.uint8, device=logits.device)
    if B == 0:
        return accept_len, mask
    if K == 0:
        return accept_len.zero_(), mask
    if logits.size(2) == 0:
        raise RuntimeError("verify_prefix: empty vocabulary")
    verify_launch(logits, draft_ids, accept_len, mask)
    return accept_len, mask


def verify_prefix_bonus_cuda(logits: torch.Tensor, draft_ids: torch.Tensor):
    """Fused verify over K+1 target rows (new registry op ``verify_prefix_bonus``).

    logits [B,K+1,V], draft_ids [B,K] -> (accept_len [B] i32, accepted_mask [B,K] u8,
    target_ids [B,K+1] i32, bonus [B] i32, packed [B,K+2] i32).  bonus[b] is the greedy token of row
    accept_len[b] (reference: src/specdec/core/pipeline.py:3140-3163, :3174-3206, :3341-3362)."""
    fast = _shim().verify_prefix_bonus_fast(logits, draft_ids)
    if fast is not None:
        return fast
    logits, draft_ids = _check_verify_inputs(logits, draft_ids, 1)
    B, rows, _ = logits.shape
    K = rows - 1
    dev = logits.device
    accept_len = torch.empty(B, dtype=torch.int32, device=dev)
    mask = torch.empty((B, K), dtype=torch.uint8, device=dev)
    target_ids = torch.empty((B, rows), dtype=torch.int32, device=dev)
    bonus = torch.empty(B, dtype=torch.int32, device=dev)
    packed = torch.empty((B, K + 2), dtype=torch.int32, device=dev)
    if B > 0:
        if logits.size(2) == 0:
            raise RuntimeError("verify_prefix_bonus: empty vocabulary")
        verify_launch(logits, draft_ids, accept_len, mask, target_ids, bonus, packed)
    return accept_len, mask, target_ids, bonus, packed


def verify_step_cuda(
    logits: torch.Tensor,
    draft_ids: torch.Tensor,
    mode: str = "typical",
    threshold: float = 0.9,
    temperature: float = 1.0,
):
    """Sampling-mode verify (new registry op ``verify_step``): the same single pass also yields the
    softmax statistics and evaluates the acceptance rule on device.

    mode "typical":    accept while softmax(target)[draft] >= threshold (policies.py:358-373)
    mode "confidence": accept while max softmax(logits) >= threshold     (policies.py:240-250)
    mode "topk":       accept while fewer than ``threshold`` (= k) target logits lie strictly above the
                       draft token's (policies.py:300-314); ``token_prob`` then holds that count
    mode "greedy":     argmax match, statistics still returned
    logits [B,K or K+1,V].  Returns dict(accept_len, accepted_mask, target_ids, bonus, token_prob,
    row_max, row_sumexp)."""
    rows_extra = logits.size(1) - draft_ids.size(1)
    assert rows_extra in (0, 1), "logits must carry K or K+1 rows"
    logits, draft_ids = _check_verify_inputs(logits, draft_ids, rows_extra)
    if temperature <= 0:
        raise ValueError("temperature must be > 0")
    B, rows, _ = logits.shape
    K = draft_ids.shape[1]
    dev = logits.device
    out = {
        "accept_len": torch.empty(B, dtype=torch.int32, device=dev),
        "accepted_mask": torch.empty((B, K), dtype=torch.uint8, device=dev),
        "target_ids": torch.empty((B, rows), dtype=torch.int32, device=dev),
        "bonus": torch.empty(B, dtype=torch.int32, device=dev),
        "token_prob": torch.empty((B, K), dtype=torch.float32, device=dev),
        "row_max": torch.empty((B, rows), dtype=torch.float32, device=dev),
        "row_sumexp": torch.empty((B, rows), dtype=torch.float32, device=dev),
    }
    if B > 0:
        verify_launch(
            logits, draft_ids, out["accept_len"], out["accepted_mask"], out["target_ids"], out["bonus"], None,
            mode=ACCEPT_MODES[mode], threshold=threshold, inv_temperature=1.0 / temperature,
            token_prob=out["token_prob"], row_max=out["row_max"], row_sumexp=out["row_sumexp"], stats=True,
        )
    return out


# ------------------------------------------------------------------------------------------ bonus sampler
def sample_bonus_cuda(
    logits: torch.Tensor,
    row_index: Optional[torch.Tensor] = None,
    uniform: Optional[torch.Tensor] = None,
    temperature: float = 1.0,
    top_k: Optional[int] = None,
    top_p: Optional[float] = None,
    row: int = 0,
    row_bias: int = 0,
    out: Optional[torch.Tensor] = None,
    kept_count: Optional[torch.Tensor] = None,
) -> torch.Tensor:
    """Batched ``sample_bonus_token_from_logits`` (reference: src/specdec/core/pipeline.py:48-147; new
    registry op ``sample_bonus``).

    logits [B, rows, V] (or [B, V]); ``row_index`` [B] int32 on device selects the row per sequence
    (typically the ``accept_len`` the verify kernel wrote), else ``row``.  ``uniform`` [B] float32 in [0,1)
    drives the draw (``do_sample=True``); ``None`` is the greedy branch.  Returns tokens [B] int32.  One
    launch, no sort, no host sync."""
    if logits.dim() == 2:
        logits = logits.unsqueeze(1)
    assert logits.dim() == 3, "logits must be [B, rows, V] or [B, V]"
    _require_cuda(logits, "logits")
    if logits.dtype not in _DTYPE:
        raise TypeError(f"logits dtype {logits.dtype} not supported")
    if logits.stride(2) != 1:
        logits = logits.contiguous()
    B, rows, V = logits.shape
    dev = logits.device
    if out is None:
        out = torch.empty(B, dtype=torch.int32, device=dev)
    if B == 0:
        return out
    if uniform is not None:
        if temperature <= 0:
            raise ValueError("temperature must be > 0 when sampling")
        uniform = uniform.to(device=dev, dtype=torch.float32).contiguous()
        assert uniform.numel() == B
    if row_index is not None:
        row_index = row_index.to(device=dev, dtype=torch.int32).contiguous()
        assert row_index.numel() == B
    lib = _capi.lib()
    with torch.cuda.device(dev):
        rc = lib.sd_sample_bonus(
            logits.data_ptr(), _DTYPE[logits.dtype], logits.stride(0), logits.stride(1), _ptr(row_index), row_bias, row,
            _ptr(uniform), float(temperature), int(top_k or 0), float(top_p if top_p is not None else 1.0),
            out.data_ptr(), _ptr(kept_count), B, rows, V, _stream_ptr(dev),
        )
    lib.check(rc, "sd_sample_bonus")
    return out


# ------------------------------------------------------------------------------------------ kv append
def arena_info(t: torch.Tensor):
    """(rows valid, capacity) of the kv_append arena behind ``t`` (a result of ``kv_append_cuda``), or None.  The
    registry lives in the native shim, keyed by the tensor's storage (a weak reference detects a freed arena)."""
    used, cap = _shim().arena_info(t)
    return None if cap < 0 else (used, cap)


def _arena_of(t: torch.Tensor, L: int):
    """If ``t`` is the [:, :, :L] prefix view of one of our arenas AND L is the arena's tip, return (arena, capacity).
    In place ONLY at the tip: a handle with L < used is stale (a rolled-back prefix, or a base that has already been
    appended to) -- writing at L would overwrite rows still visible through the newer result tensors, which torch.cat
    never does.  Such a call gets a fresh arena (one adoption copy, then O(Kn) again)."""
    if t.dim() != 4 or t.storage_offset() != 0:
        return None
    info = arena_info(t)
    if info is None or info[0] != L:
        return None
    cap = info[1]
    B, H, _, D = t.shape
    if t.numel() > 0 and t.stride() != (H * cap * D, cap * D, D, 1):
        return None
    if t.untyped_storage().nbytes() != B * H * cap * D * t.element_size():
        return None
    arena = torch.empty(0, dtype=t.dtype, device=t.device).set_(t.untyped_storage(), 0, (B, H, cap, D), (H * cap * D, cap * D, D, 1))
    return arena, cap


def _grow_capacity(need: int) -> int:
    # geometric growth: amortised O(1) bytes copied per appended row (torch.cat costs O(L) each call)
    return max(64, 1 << (max(need, 1) * 2 - 1).bit_length())


def _kv_layer_launch(arena_k, arena_v, new_k, new_v, base_len, n_keep, keep_bias, mask, zero_rejected) -> None:
    """One sd_kv_append_layer launch on the current stream (native marshalling: sd_b200::kv_append_rows_)."""
    _shim().kv_append_rows_(arena_k, arena_v, new_k, new_v, base_len, n_keep, keep_bias, mask, bool(zero_rejected))


def _same_layout(a: torch.Tensor, b: torch.Tensor) -> bool:
    return a.shape == b.shape and a.stride() == b.stride()


def _rows_operand(t: torch.Tensor) -> torch.Tensor:
    """Offered rows must have a contiguous head dim; other strides are passed through."""
    if t.size(3) > 0 and t.stride(3) != 1:
        t = t.contiguous()
    return t


def _check_kv_inputs(base_k, base_v, new_k, new_v) -> None:
    # reference assertions: src/kernels/reference.py:77-87
    sb, sn = base_k.shape, new_k.shape
    assert len(sb) == 4, "base_k must be 4D tensor [B][H][L][D]"
    assert base_v.dim() == 4, "base_v must be 4D tensor [B][H][L][D]"
    assert len(sn) == 4, "new_k must be 4D tensor [B][H][K][D]"
    assert new_v.dim() == 4, "new_v must be 4D tensor [B][H][K][D]"
    assert sb[0] == sn[0], "Batch size mismatch"
    assert sb[1] == sn[1], "Num heads mismatch"
    assert sb[3] == sn[3], "Head dim mismatch"
    assert sb == base_v.shape and sn == new_v.shape, "K/V shape mismatch"
    if not (base_k.is_cuda and base_v.is_cuda and new_k.is_cuda and new_v.is_cuda):
        for t, n in ((base_k, "base_k"), (base_v, "base_v"), (new_k, "new_k"), (new_v, "new_v")):
            _require_cuda(t, n)
    dt = base_k.dtype
    assert dt == base_v.dtype and dt == new_k.dtype and dt == new_v.dtype, "dtype mismatch"


def _append_functional(base_k, base_v, new_k, new_v, n_keep, mask, zero_rejected):
    _check_kv_inputs(base_k, base_v, new_k, new_v)
    B, H, L, D = base_k.shape
    R = new_k.shape[2]
    new_k, new_v = _rows_operand(new_k), _rows_operand(new_v)
    if not _same_layout(new_k, new_v):
        new_k, new_v = new_k.contiguous(), new_v.contiguous()
    hit_k, hit_v = _arena_of(base_k, L), _arena_of(base_v, L)
    arena_k = arena_v = None
    if hit_k and hit_v and hit_k[1] == hit_v[1] and L + R <= hit_k[1]:
        arena_k, arena_v = hit_k[0], hit_v[0]
    if arena_k is None:
        cap = _grow_capacity(L + R)
        arena_k = torch.empty((B, H, cap, D), dtype=base_k.dtype, device=base_k.device)
        arena_v = torch.empty((B, H, cap, D), dtype=base_k.dtype, device=base_k.device)
        _shim().arena_register(arena_k, 0)
        _shim().arena_register(arena_v, 0)
        if L > 0 and B * H * D > 0:  # one-time adoption copy of the existing rows (our kernel, R = L)
            bk, bv = _rows_operand(base_k), _rows_operand(base_v)
            if not _same_layout(bk, bv):
                bk, bv = bk.contiguous(), bv.contiguous()
            _kv_layer_launch(arena_k, arena_v, bk, bv, 0, None, 0, None, False)
    if R > 0 and B * H * D > 0:
        _kv_layer_launch(arena_k, arena_v, new_k, new_v, L, n_keep, 0, mask, zero_rejected)
    _shim().arena_set_used(arena_k, L + R)
    _shim().arena_set_used(arena_v, L + R)
    return arena_k[:, :, : L + R], arena_v[:, :, : L + R]


def kv_append_cuda(base_k, base_v, new_k, new_v):
    """Drop-in for the registry op ``kv_append`` (reference: src/kernels/reference.py:59-93; call site
    src/specdec/models/hf_wrappers.py:1009).

    Returns (out_k, out_v) = rows of base followed by rows of new, [B,H,L+Kn,D].  The results are
    views of a geometrically grown arena: when ``base_*`` is itself the LATEST result of this function (the
    reference's usage: ``cache = kv_append(cache, chunk)``, hf_wrappers.py:941-943), only the Kn new rows move --
    O(Kn) bytes instead of torch.cat's O(L) -- and the result shares storage with ``base_*``, whose own rows
    [0, L) are never rewritten.  Any other base (a stale handle: a prefix that has already been appended to, or a
    rolled-back slice) is copied into a fresh arena first, so two appends to the SAME base give two independent
    results exactly like torch.cat (tests/test_kv_gpu.py::test_two_appends_to_the_same_base).
    The tip append -- the steady state of the reference's per-layer loop -- is served natively
    (csrc_torch/sd_torch_ops.cpp::kv_append_fast, ~4 us of host time instead of ~20 us through Python + ctypes)."""
    fast = _shim().kv_append_fast(base_k, base_v, new_k, new_v)
    if fast is not None:
        return fast
    return _append_functional(base_k, base_v, new_k, new_v, None, None, False)


def kv_append_with_mask_cuda(base_k, base_v, draft_k, draft_v, accepted_mask, accept_len, offset: int = 0):
    """Registry op ``kv_append_with_mask`` (reference: src/kernels/reference.py:96-159): masked rows
    compacted after the base rows, at most accept_len[b] of them, zero tail; ``offset`` ignored as in
    the reference.  Zero acceptance keeps the base rows (the reference's dead CUDA kernel zeroed
    them: src/kernels/cuda/kv_cache.cu:39-40)."""
    assert accepted_mask.dim() == 2, "accepted_mask must be 2D tensor [B][K]"
    assert accept_len.dim() == 1, "accept_len must be 1D tensor [B]"
    dev = base_k.device
    mask = accepted_mask.to(device=dev, dtype=torch.uint8).contiguous()
    alen = accept_len.to(device=dev, dtype=torch.int32).contiguous()
    return _append_functional(base_k, base_v, draft_k, draft_v, alen, mask, True)


def kv_arena_append_launch(
    layer_table: torch.Tensor,
    n_layers: int,
    B: int,
    H: int,
    R: int,
    row_bytes: int,
    capacity: int,
    arena_strides: Tuple[int, int],
    new_strides: Tuple[int, int, int],
    seq_lens: torch.Tensor,
    seq_lens_out: Optional[torch.Tensor],
    n_keep: Optional[torch.Tensor],
    keep_bias: int,
    accepted_mask: Optional[torch.Tensor],
    zero_rejected: bool,
    max_seq_len_host: int,
    overflow_flag: Optional[torch.Tensor],
    ticket: torch.Tensor,
    rows_ready_early: bool = False,
    exchange: Optional[int] = None,
    exchange_out: Optional[torch.Tensor] = None,
    exchange_packed: Optional[torch.Tensor] = None,
    seq_ready: Optional[torch.Tensor] = None,
) -> None:
    """All-layer arena append (sd_kv_append).  ``layer_table`` is a device int64 [n_layers, 4] tensor
    of (arena_k, arena_v, new_k, new_v) pointers == sd_kv_layer[n_layers]."""
    dev = layer_table.device
    lib = _capi.lib()
    with torch.cuda.device(dev):
        rc = lib.sd_kv_append(
            layer_table.data_ptr(), n_layers, B, H, R, row_bytes, capacity, arena_strides[0], arena_strides[1],
            new_strides[0], new_strides[1], new_strides[2], _ptr(seq_lens), _ptr(seq_lens_out), _ptr(n_keep),
            keep_bias, _ptr(accepted_mask), int(zero_rejected), max_seq_len_host, _ptr(overflow_flag),
            ticket.data_ptr(), int(rows_ready_early), _stream_ptr(dev), exchange, _ptr(exchange_out), _ptr(exchange_packed), _ptr(seq_ready),
        )
    lib.check(rc, "sd_kv_append")


# ------------------------------------------------------------------------------------------ fused step
def fused_step_available(arena) -> bool:
    """sd_verify_append needs the default bulk-chunk verify variant and 16-byte aligned, unmasked KV rows."""
    return (_verify_variant_state["variant"] == DEFAULT_VERIFY_VARIANT and arena.row_bytes % 16 == 0
            and _kv_variant_state["variant"] in (0, 3))


def verify_append_launch(
    logits: torch.Tensor,
    draft_ids: torch.Tensor,
    accept_len: torch.Tensor,
    accepted_mask: Optional[torch.Tensor],
    target_ids: Optional[torch.Tensor],
    bonus: Optional[torch.Tensor],
    packed: Optional[torch.Tensor],
    workspace: torch.Tensor,
    seq_ready: torch.Tensor,
    arena,
    zero_rejected: bool = True,
    check_capacity: bool = False,
    exchange=None,
    exchange_push_in_append: bool = False,
) -> None:
    """The whole greedy verify step in ONE launch (sd_verify_append): fused verify over ``logits``
    [B,K+1,V] + kv_append of ``arena``'s staged rows for every layer (accept_len + 1 rows kept).  ``arena`` is a
    specdec.cache.kv_arena.KVArena; ``exchange`` a scheduler.batch_shard.PeerExchange (its records are pushed by
    the verify finalizers, each as soon as its sequence is verified -- or, with ``exchange_push_in_append``, by
    the first append CTA once all are final, measured slower -- and gathered by the append part into
    ``exchange.out``)."""
    import ctypes as C

    B, rows, V = logits.shape
    K = draft_ids.shape[1]
    es = arena.k.element_size()
    a = _capi.StepArgs()
    a.logits, a.logit_dtype, a.id_dtype = logits.data_ptr(), _DTYPE[logits.dtype], _ID_DTYPE[draft_ids.dtype]
    a.stride_b, a.stride_r, a.id_stride_b = logits.stride(0), logits.stride(1), draft_ids.stride(0)
    a.draft_ids, a.accept_len = draft_ids.data_ptr(), accept_len.data_ptr()
    a.accepted_mask, a.target_ids, a.bonus, a.packed = _ptr(accepted_mask), _ptr(target_ids), _ptr(bonus), _ptr(packed)
    a.workspace, a.workspace_bytes = workspace.data_ptr(), workspace.numel()
    a.B, a.rows, a.K, a.V = B, rows, K, V
    a.layers, a.n_layers, a.H = arena.layer_table.data_ptr(), arena.n_layers, arena.n_kv_heads
    a.zero_rejected, a.exchange_push_in_append = int(zero_rejected), int(exchange_push_in_append and exchange is not None)
    a.row_bytes, a.capacity = arena.row_bytes, arena.capacity
    a.arena_stride_b, a.arena_stride_h = arena.k.stride(1) * es, arena.k.stride(2) * es
    a.new_stride_b, a.new_stride_h, a.new_stride_r = (arena.new_k.stride(1) * es, arena.new_k.stride(2) * es,
                                                      arena.new_k.stride(3) * es)
    a.max_seq_len_host = arena._max_len_host if check_capacity else -1
    a.seq_lens, a.overflow_flag, a.ticket = arena.seq_lens.data_ptr(), arena.overflow.data_ptr(), arena._ticket.data_ptr()
    a.seq_ready = seq_ready.data_ptr()
    a.exchange = exchange.desc_ptr if exchange is not None else None
    a.exchange_out = exchange.out.data_ptr() if exchange is not None else None
    a.step_state = arena.step_state().data_ptr()
    uni = arena.uniform_layers()
    if uni is not None:  # one [layers, ...] allocation: layer pointers by arithmetic, no table lookup on the device
        (a.layer0.arena_k, a.layer0.arena_v, a.layer0.new_k, a.layer0.new_v), a.layer_stride_arena, a.layer_stride_new = uni
    lib = _capi.lib()
    dev = logits.device
    with torch.cuda.device(dev):
        rc = lib.sd_verify_append(C.byref(a), _stream_ptr(dev))
    lib.check(rc, "sd_verify_append")
    arena._max_len_host = min(arena.capacity, arena._max_len_host + rows)


# ------------------------------------------------------------------------------------------ attention
ATTN_BALANCED = -1  # SD_ATTN_BALANCED (include/sd_verify.h)


class BalancedAttentionWorkspace:
    """Workspace of the balanced (persistent) form of the verify attention: per-item tile counters + per-CTA partial
    slots, zeroed ONCE here (every launch leaves the counters zero again -> graph-replayable).  One workspace serves one
    launch at a time: keep one per stream / per captured graph node chain.  ``reset()`` after an aborted launch."""

    def __init__(self, B: int, Hq: int, Hkv: int, R: int, D: int, device: torch.device):
        lib = _capi.lib()
        self.shape = (B, Hq, Hkv, R, D)
        self.device = torch.device(device)
        self.buf = torch.empty(int(lib.sd_verify_attention_balanced_workspace_bytes(B, Hq, Hkv, R, D)), dtype=torch.uint8,
                               device=self.device)
        self.reset()

    def reset(self) -> None:
        lib = _capi.lib()
        with torch.cuda.device(self.device):
            lib.check(lib.sd_verify_attention_balanced_workspace_init(self.buf.data_ptr(), self.buf.numel(),
                                                                      _stream_ptr(self.device)),
                      "sd_verify_attention_balanced_workspace_init")

    def fits(self, B: int, Hq: int, Hkv: int, R: int, D: int) -> bool:
        return (B, Hq, Hkv, R, D) == self.shape


def verify_attention_cuda(
    q: torch.Tensor,
    k_arena: torch.Tensor,
    v_arena: torch.Tensor,
    seq_lens: torch.Tensor,
    scale: Optional[float] = None,
    max_seq_len: int = -1,
    n_splits: int = 0,
    out: Optional[torch.Tensor] = None,
    workspace: Optional[torch.Tensor] = None,
    overflow: Optional[torch.Tensor] = None,
    balanced: Optional[BalancedAttentionWorkspace] = None,
) -> torch.Tensor:
    """(K+1)-query causal decode attention over the KV arena (new registry op ``verify_attention``).

    ``balanced``: a :class:`BalancedAttentionWorkspace` selects the persistent form (one CTA per resident slot, equal
    spans of key tiles; B <= 512) -- same results up to fp32 summation order.

    q [B,Hq,R,D]; k_arena/v_arena [B,Hkv,C,D] with the R new rows already appended at
    seq_lens[b]..; seq_lens [B] int32 = rows valid before them.  Replaces K sequential SDPA decode
    steps of the reference (src/scheduler/speculative_scheduler.py:192-199,
    src/specdec/models/hf_wrappers.py:98).  The arena capacity C is passed down: lengths are clamped to it on the
    device (``overflow`` [1] int32, when given, is set if that happened)."""
    assert q.dim() == 4 and k_arena.dim() == 4 and v_arena.dim() == 4
    _require_cuda(q, "q")
    B, Hq, R, D = q.shape
    Hkv, C = k_arena.shape[1], k_arena.shape[2]
    assert k_arena.shape == v_arena.shape and k_arena.shape[0] == B and k_arena.shape[3] == D
    assert Hq % Hkv == 0, "query heads must be a multiple of KV heads"
    assert q.dtype in (torch.float16, torch.bfloat16) and k_arena.dtype == q.dtype == v_arena.dtype
    assert k_arena.stride(3) == 1 and k_arena.stride(2) == D and k_arena.stride() == v_arena.stride()
    q = q.contiguous()
    seq_lens = seq_lens.to(device=q.device, dtype=torch.int32)
    if out is None:
        out = torch.empty_like(q)
    lib = _capi.lib()
    es = q.element_size()
    if balanced is not None:
        assert balanced.fits(B, Hq, Hkv, R, D), f"balanced workspace was made for {balanced.shape}"
        workspace, n_splits = balanced.buf, ATTN_BALANCED
    elif workspace is None:
        need = int(lib.sd_verify_attention_workspace_bytes(B, Hq, R, D, n_splits))
        workspace = _scratch.attn_workspace(q.device, need)
    with torch.cuda.device(q.device):
        rc = lib.sd_verify_attention(
            q.data_ptr(), k_arena.data_ptr(), v_arena.data_ptr(), seq_lens.data_ptr(), out.data_ptr(),
            workspace.data_ptr(), workspace.numel(), _DTYPE[q.dtype], B, Hq, Hkv, R, D,
            k_arena.stride(0) * es, k_arena.stride(1) * es, max_seq_len, C, _ptr(overflow),
            float(scale) if scale is not None else 0.0, n_splits, _stream_ptr(q.device),
        )
    lib.check(rc, "sd_verify_attention")
    return out


# ------------------------------------------------------------------------------------------ torch ops
_TORCH_LIB = None
_TORCH_OP_KERNELS: list = []  # the Python kernels behind the ops defined below (tests call them directly)


def _register_torch_ops() -> None:
    """Expose the entry points as torch custom ops (CUDA dispatch key only -> CPU tensors raise
    NotImplementedError from the dispatcher: there is no CPU kernel to fall back to)."""
    global _TORCH_LIB
    if _TORCH_LIB is not None:
        return
    _shim()  # C++: sd_b200::verify_prefix, ::verify_prefix_bonus, ::kv_append_rows_ (TORCH_LIBRARY_FRAGMENT, CUDA key)
    lib = torch.library.Library("sd_b200", "FRAGMENT")
    lib.define(
        "verify_step(Tensor logits, Tensor draft_ids, int mode, float threshold, float temperature)"
        " -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)"
    )
    lib.define("kv_append(Tensor base_k, Tensor base_v, Tensor new_k, Tensor new_v) -> (Tensor, Tensor)")
    lib.define(
        "kv_append_with_mask(Tensor base_k, Tensor base_v, Tensor draft_k, Tensor draft_v, Tensor accepted_mask,"
        " Tensor accept_len) -> (Tensor, Tensor)"
    )
    lib.define(
        "verify_attention(Tensor q, Tensor k_arena, Tensor v_arena, Tensor seq_lens, float scale, int max_seq_len)"
        " -> Tensor"
    )

    lib.define(
        "sample_bonus(Tensor logits, Tensor? row_index, Tensor? uniform, float temperature, int top_k, float top_p)"
        " -> Tensor"
    )

    def _sample_bonus_op(logits, row_index, uniform, temperature, top_k, top_p):
        # top_k <= 0 / top_p >= 1 = filter off (reference defaults: pipeline.py:52-53); uniform None = greedy branch
        return sample_bonus_cuda(logits, row_index=row_index, uniform=uniform, temperature=float(temperature),
                                 top_k=int(top_k) if top_k > 0 else None, top_p=float(top_p) if top_p < 1.0 else None)

    def _verify_step_op(logits, draft_ids, mode, threshold, temperature):
        name = {v: k for k, v in ACCEPT_MODES.items()}[int(mode)]
        o = verify_step_cuda(logits, draft_ids, name, threshold, temperature)
        return (o["accept_len"], o["accepted_mask"], o["target_ids"], o["bonus"], o["token_prob"], o["row_max"],
                o["row_sumexp"])

    def _fresh_append(base_k, base_v, new_k, new_v, n_keep, mask, zero_rejected):
        # A functional torch op may not return views of its inputs: the result is written, by the append kernel
        # itself (out-variant sd_b200::kv_append_rows_), into fresh tensors of exactly [B,H,L+R,D] -- no clone of a view.
        _check_kv_inputs(base_k, base_v, new_k, new_v)
        B, H, L, D = base_k.shape
        R = new_k.shape[2]
        out_k = torch.zeros((B, H, L + R, D), dtype=base_k.dtype, device=base_k.device) if mask is not None else \
            torch.empty((B, H, L + R, D), dtype=base_k.dtype, device=base_k.device)
        out_v = torch.zeros_like(out_k) if mask is not None else torch.empty_like(out_k)
        if B * H * D > 0:
            if L > 0:
                _kv_layer_launch(out_k, out_v, _rows_operand(base_k).contiguous(), _rows_operand(base_v).contiguous(), 0, None, 0, None, False)
            if R > 0:
                _kv_layer_launch(out_k, out_v, _rows_operand(new_k).contiguous(), _rows_operand(new_v).contiguous(), L, n_keep, 0, mask, zero_rejected)
        return out_k, out_v

    def _kv_append_op(base_k, base_v, new_k, new_v):
        return _fresh_append(base_k, base_v, new_k, new_v, None, None, False)

    def _kv_append_mask_op(base_k, base_v, draft_k, draft_v, accepted_mask, accept_len):
        mask = accepted_mask.to(device=base_k.device, dtype=torch.uint8).contiguous()
        alen = accept_len.to(device=base_k.device, dtype=torch.int32).contiguous()
        return _fresh_append(base_k, base_v, draft_k, draft_v, alen, mask, True)

    def _attn_op(q, k_arena, v_arena, seq_lens, scale, max_seq_len):
        return verify_attention_cuda(q, k_arena, v_arena, seq_lens, scale if scale > 0 else None, max_seq_len)

    lib.impl("verify_step", _verify_step_op, "CUDA")
    lib.impl("kv_append", _kv_append_op, "CUDA")
    lib.impl("kv_append_with_mask", _kv_append_mask_op, "CUDA")
    lib.impl("verify_attention", _attn_op, "CUDA")
    lib.impl("sample_bonus", _sample_bonus_op, "CUDA")
    _TORCH_OP_KERNELS[:] = [_verify_step_op, _kv_append_op, _kv_append_mask_op, _attn_op, _sample_bonus_op]
    _TORCH_LIB = lib


def set_verify_tuning(threads: int = 0, vectors_per_thread: int = 0) -> None:
    _capi.lib().check(_capi.lib().sd_verify_set_tuning(threads, vectors_per_thread), "sd_verify_set_tuning")


DEFAULT_VERIFY_VARIANT = 3  # bulk-chunk kernel (csrc/sd_verify.cu)


_verify_variant_state = {"variant": DEFAULT_VERIFY_VARIANT}
_kv_variant_state = {"variant": 0}


def set_verify_variant(variant: int = DEFAULT_VERIFY_VARIANT, ctas_per_sm: int = 0) -> None:
    _verify_variant_state["variant"] = variant
    _capi.lib().check(_capi.lib().sd_verify_set_variant(variant, ctas_per_sm), "sd_verify_set_variant")


def set_attention_tuning(config: int = 0) -> None:
    _capi.lib().check(_capi.lib().sd_verify_attention_set_tuning(config), "sd_verify_attention_set_tuning")


def set_fused_mode(mode: int = 0, lag: int = 0) -> None:
    """Layout of the fused step (include/sd_verify.h: sd_verify_append_set_mode): 0 = lagged append (default),
    1 = the round-1 layout (append CTAs behind all verify CTAs), kept for A/B evidence."""
    _capi.lib().check(_capi.lib().sd_verify_append_set_mode(int(mode), int(lag)), "sd_verify_append_set_mode")


class StartGate:
    """Device-side start gate (include/sd_verify.h: sd_start_gate).  ``arm()`` enqueues the gate kernel on the current
    stream; work enqueued after it starts only when ``release()`` has been called on the host AND (with a
    ``scheduler.batch_shard.PeerExchange``) every rank has reached its own gate."""

    def __init__(self, device, exchange=None) -> None:
        self.device = torch.device(device)
        self.exchange = exchange
        self.flag = torch.zeros(1, dtype=torch.int32).pin_memory()
        self.value = 0

    def arm(self, hold: bool = True) -> None:
        self.value += 1
        if not hold:
            self.flag[0] = self.value
        with torch.cuda.device(self.device):
            rc = _capi.lib().sd_start_gate(self.exchange.desc_ptr if self.exchange is not None else None,
                                           self.flag.data_ptr(), self.value, _stream_ptr(self.device))
        _capi.lib().check(rc, "sd_start_gate")

    def release(self) -> None:
        self.flag[0] = self.value


def device_error(clear: bool = True) -> int:
    """OR of the bounded device-side waits that expired since the last clear (include/sd_verify.h:
    sd_device_error); 0 = none.  Synchronises the device."""
    import ctypes as C

    v = C.c_uint(0)
    _capi.lib().check(_capi.lib().sd_device_error(int(clear), C.byref(v)), "sd_device_error")
    return int(v.value)


def set_sample_fast_path(enabled: bool = True) -> None:
    """Bonus sampler: two-pass candidate path for 0 < top_k <= 1024 (default on); off = always the radix-select path."""
    _capi.lib().check(_capi.lib().sd_sample_set_fast_path(int(enabled)), "sd_sample_set_fast_path")


def set_pdl(enabled: bool = True) -> None:
    _capi.lib().check(_capi.lib().sd_set_pdl(int(enabled)), "sd_set_pdl")


def set_kv_variant(variant: int = 0) -> None:
    _kv_variant_state["variant"] = variant & 15
    _capi.lib().check(_capi.lib().sd_kv_set_variant(variant), "sd_kv_set_variant")
