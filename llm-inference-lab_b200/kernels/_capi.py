"""ctypes binding of include/sd_verify.h (the C ABI of libsd_verify_b200.so).

This is the "thin C-ABI extension": plain pointers and sizes, no torch types.  ``cuda_ops.py``
wraps these calls as torch custom ops; ``INTEGRATION.md`` shows the same binding written for the
reference tree.  Importing this module never falls back to another implementation: a missing or
stale library is rebuilt with nvcc, and failure raises.
"""

from __future__ import annotations

import ctypes as C
import re
from pathlib import Path
from typing import List

from . import build as _build

DTYPE_F32, DTYPE_F16, DTYPE_BF16 = 0, 1, 2
ID_I32, ID_I64 = 0, 1
ACCEPT_GREEDY, ACCEPT_TYPICAL, ACCEPT_CONFIDENCE, ACCEPT_TOPK = 0, 1, 2, 3
ABI_VERSION = 3

_vp = C.c_void_p
_i = C.c_int
_i64 = C.c_int64
_f = C.c_float
_sz = C.c_size_t

# name -> (restype, argtypes); must list every function include/sd_verify.h declares
SIGNATURES = {
    "sd_abi_version": (_i, []),
    "sd_struct_layout": (_sz, [_i]),
    "sd_error_string": (C.c_char_p, [_i]),
    "sd_set_pdl": (_i, [_i]),
    "sd_device_info": (_i, [C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)]),
    "sd_verify_workspace_bytes": (_sz, [_i, _i]),
    "sd_verify_workspace_init": (_i, [_vp, _sz, _vp]),
    "sd_verify_prefix": (_i, [_vp, _i, _i64, _i64, _vp, _i, _i64, _vp, _vp, _vp, _vp, _vp, _vp, _sz,
                              _i, _i, _i, _i, _vp, _vp, _vp]),
    "sd_verify_append": (_i, [_vp, _vp]),
    "sd_verify_append_set_mode": (_i, [_i, _i]),
    "sd_step_state_bytes": (_sz, [_i]),
    "sd_device_error": (_i, [_i, C.POINTER(C.c_uint)]),
    "sd_verify_step": (_i, [_vp, _i, _i64, _i64, _vp, _i, _i64, _i, _f, _f, _vp, _vp, _vp, _vp, _vp,
                            _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _vp, _vp, _vp]),
    "sd_exchange_bytes": (_sz, [_i, _i, _i]),
    "sd_exchange_alloc": (_i, [_sz, C.POINTER(_vp), _vp]),
    "sd_exchange_open": (_i, [_vp, C.POINTER(_vp)]),
    "sd_exchange_close": (_i, [_vp]),
    "sd_exchange_free": (_i, [_vp]),
    "sd_exchange_wait": (_i, [_vp, _vp, _vp]),
    "sd_start_gate": (_i, [_vp, _vp, C.c_uint, _vp]),
    "sd_selftest_gate_clock": (_i, [_vp, _i, C.c_uint64]),
    "sd_verify_set_tuning": (_i, [_i, _i]),
    "sd_verify_set_variant": (_i, [_i, _i]),
    "sd_verify_set_trace": (_i, [_vp, _sz]),
    "sd_kv_append": (_i, [_vp, _i, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _vp, _vp, _vp,
                          _i, _vp, _i, _i64, _vp, _vp, _i, _vp, _vp, _vp, _vp, _vp]),
    "sd_kv_append_layer": (_i, [_vp, _vp, _vp, _vp, _i, _i, _i, _i64, _i64, _i64, _i64, _i64, _i64, _i64, _i64,
                                _vp, _i, _vp, _i, _vp, _vp]),
    "sd_kv_set_variant": (_i, [_i]),
    "sd_sample_set_fast_path": (_i, [_i]),
    "sd_sample_bonus": (_i, [_vp, _i, _i64, _i64, _vp, _i, _i, _vp, _f, _i, _f, _vp, _vp, _i, _i, _i, _vp]),
    "sd_verify_attention_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sd_verify_attention_balanced_workspace_bytes": (_sz, [_i, _i, _i, _i, _i]),
    "sd_verify_attention_balanced_workspace_init": (_i, [_vp, _sz, _vp]),
    "sd_verify_attention_set_tuning": (_i, [_i]),
    "sd_verify_attention": (_i, [_vp, _vp, _vp, _vp, _vp, _vp, _sz, _i, _i, _i, _i, _i, _i, _i64, _i64,
                                 _i64, _i64, _vp, _f, _i, _vp]),
}


def declared_symbols(header: Path | None = None) -> List[str]:
    """Function names declared in include/sd_verify.h (used by the symbol-export test)."""
    header = header or (_build.INCLUDE / "sd_verify.h")
    text = re.sub(r"/\*.*?\*/", "", header.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(sd_[a-z0-9_]+)\s*\(", text)))


MAX_PEERS = 16


class ExchangeDesc(C.Structure):
    """struct sd_exchange (include/sd_verify.h)."""

    _fields_ = [("world", C.c_int32), ("rank", C.c_int32), ("rows_per_rank", C.c_int32), ("width", C.c_int32),
                ("total_rows", C.c_int32), ("lag", C.c_int32), ("reserved", C.c_int32 * 2), ("peer", C.c_void_p * MAX_PEERS)]


class KvLayer(C.Structure):
    """struct sd_kv_layer (include/sd_verify.h)."""

    _fields_ = [("arena_k", C.c_void_p), ("arena_v", C.c_void_p), ("new_k", C.c_void_p), ("new_v", C.c_void_p)]


class StepArgs(C.Structure):
    """struct sd_step_args (include/sd_verify.h)."""

    _fields_ = [
        ("logits", C.c_void_p), ("logit_dtype", C.c_int32), ("id_dtype", C.c_int32),
        ("stride_b", C.c_int64), ("stride_r", C.c_int64), ("id_stride_b", C.c_int64),
        ("draft_ids", C.c_void_p), ("accept_len", C.c_void_p), ("accepted_mask", C.c_void_p), ("target_ids", C.c_void_p),
        ("bonus", C.c_void_p), ("packed", C.c_void_p), ("workspace", C.c_void_p), ("workspace_bytes", C.c_size_t),
        ("B", C.c_int32), ("rows", C.c_int32), ("K", C.c_int32), ("V", C.c_int32),
        ("layers", C.c_void_p), ("n_layers", C.c_int32), ("H", C.c_int32), ("zero_rejected", C.c_int32),
        ("exchange_push_in_append", C.c_int32),
        ("row_bytes", C.c_int64), ("capacity", C.c_int64), ("arena_stride_b", C.c_int64), ("arena_stride_h", C.c_int64),
        ("new_stride_b", C.c_int64), ("new_stride_h", C.c_int64), ("new_stride_r", C.c_int64), ("max_seq_len_host", C.c_int64),
        ("seq_lens", C.c_void_p), ("overflow_flag", C.c_void_p), ("ticket", C.c_void_p), ("seq_ready", C.c_void_p),
        ("exchange", C.c_void_p), ("exchange_out", C.c_void_p), ("step_state", C.c_void_p),
        ("layer0", KvLayer), ("layer_stride_arena", C.c_int64), ("layer_stride_new", C.c_int64),
    ]


class SdError(RuntimeError):
    def __init__(self, code: int, what: str):
        super().__init__(f"{what}: {code} ({_error_string(code)})")
        self.code = code


class Lib:
    """Loaded library with typed entry points."""

    def __init__(self, path: Path):
        self.path = Path(path)
        self.cdll = C.CDLL(str(path))
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(self.cdll, name)  # AttributeError here == symbol not exported: loud
            fn.restype = res
            fn.argtypes = args
            setattr(self, name, fn)
        got = self.sd_abi_version()
        if got != ABI_VERSION:
            raise RuntimeError(f"{path}: ABI version {got}, binding expects {ABI_VERSION}")

    def check(self, code: int, what: str) -> None:
        if code != 0:
            raise SdError(code, what)


_LIB: Lib | None = None


def lib() -> Lib:
    global _LIB
    if _LIB is None:
        _LIB = Lib(_build.build_cuda_kernels())
    return _LIB


def _error_string(code: int) -> str:
    try:
        return lib().sd_error_string(code).decode()
    except Exception:  # pragma: no cover
        return "?"
