"""ctypes loader for oracle/_build/libsd_oracle.so (TEST INFRASTRUCTURE ONLY -- see oracle.c)."""

from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB = _HERE / "_build" / "libsd_oracle.so"
_DT = {"fp32": 0, "fp16": 1, "bf16": 2}


def build(force: bool = False) -> Path:
    if force or not _LIB.exists() or _LIB.stat().st_mtime < (_HERE / "oracle.c").stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = C.CDLL(str(build()))
        _lib.sd_oracle_verify.restype = C.c_int
        _lib.sd_oracle_kv_append.restype = C.c_int
        _lib.sd_oracle_kv_arena_append.restype = C.c_int
        _lib.sd_oracle_num_threads.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().sd_oracle_num_threads())


def _ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


def verify(logits: np.ndarray, dtype: str, draft_ids: np.ndarray, rows: int | None = None):
    """logits [B,R,V] contiguous (np.float32 / np.float16 / np.uint16 bf16 bits); ids [B,K] int32/int64."""
    logits = np.ascontiguousarray(logits)
    ids = np.ascontiguousarray(draft_ids)
    B, R, V = logits.shape
    K = ids.shape[1]
    id_dt = 0 if ids.dtype == np.int32 else 1
    acc = np.zeros(B, np.int32)
    mask = np.zeros((B, K), np.uint8)
    tgt = np.zeros((B, R), np.int32)
    bonus = np.zeros(B, np.int32)
    rc = lib().sd_oracle_verify(
        _ptr(logits), _DT[dtype], C.c_int64(R * V), C.c_int64(V), _ptr(ids), id_dt, C.c_int64(K),
        B, R, K, V, _ptr(acc), _ptr(mask), _ptr(tgt), _ptr(bonus),
    )
    if rc != 0:
        raise RuntimeError(f"sd_oracle_verify rc={rc}")
    return acc, mask, tgt, bonus


def kv_append(base: np.ndarray, draft: np.ndarray, mask=None, accept_len=None) -> np.ndarray:
    """One of K/V: base [B,H,L,D], draft [B,H,K,D] (any 2/4-byte dtype) -> [B,H,L+K,D]."""
    base = np.ascontiguousarray(base)
    draft = np.ascontiguousarray(draft)
    B, H, L, D = base.shape
    K = draft.shape[2]
    out = np.empty((B, H, L + K, D), base.dtype)
    m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
    a = None if accept_len is None else np.ascontiguousarray(accept_len, dtype=np.int32)
    rc = lib().sd_oracle_kv_append(
        _ptr(base), _ptr(draft), None if m is None else _ptr(m), None if a is None else _ptr(a),
        _ptr(out), B, H, L, K, C.c_int64(D * base.dtype.itemsize),
    )
    if rc != 0:
        raise RuntimeError(f"sd_oracle_kv_append rc={rc}")
    return out


def kv_arena_append(arena: np.ndarray, seq_lens, fresh: np.ndarray, n_keep, zero_rejected=True):
    """In place on ``arena`` [B,H,C,D]; returns new lengths."""
    assert arena.flags.c_contiguous
    fresh = np.ascontiguousarray(fresh)
    B, H, Cap, D = arena.shape
    R = fresh.shape[2]
    lens = np.ascontiguousarray(seq_lens, dtype=np.int32)
    keep = np.ascontiguousarray(n_keep, dtype=np.int32)
    new_lens = np.zeros(B, np.int32)
    rc = lib().sd_oracle_kv_arena_append(
        _ptr(arena), _ptr(lens), _ptr(fresh), _ptr(keep), _ptr(new_lens), B, H, Cap, R,
        C.c_int64(D * arena.dtype.itemsize), int(bool(zero_rejected)),
    )
    if rc != 0:
        raise RuntimeError(f"sd_oracle_kv_arena_append rc={rc}")
    return new_lens
