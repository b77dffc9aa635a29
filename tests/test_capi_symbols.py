"""CPU: the C-ABI library builds for sm_100a, loads, and exports every symbol include/sd_verify.h
declares.  No compute calls (no GPU here)."""
import ctypes
import subprocess

import pytest

from kernels import _capi, build


def test_library_builds_and_is_current():
    path = build.build_cuda_kernels()
    assert path.exists() and path.stat().st_size > 0
    assert build.is_current()
    assert build.get_cuda_arch() == "sm_100a"


def test_every_declared_symbol_is_exported_and_bound():
    declared = _capi.declared_symbols()
    assert len(declared) >= 12
    cdll = ctypes.CDLL(str(build.LIB_PATH))
    for name in declared:
        assert hasattr(cdll, name), f"{name} declared in include/sd_verify.h but not exported"
    assert sorted(_capi.SIGNATURES) == declared, "ctypes binding and header disagree"


def test_abi_version_and_error_strings():
    lib = _capi.lib()
    assert lib.sd_abi_version() == _capi.ABI_VERSION
    assert lib.sd_error_string(0) == b"ok"
    assert b"workspace" in lib.sd_error_string(-3)
    assert lib.sd_verify_workspace_bytes(64, 5) >= 64 * 5 * 64 * 8


def test_argument_errors_are_reported_without_a_gpu():
    lib = _capi.lib()
    # rows must be K or K+1 -> SD_ERR_SHAPE, detected before any CUDA call
    rc = lib.sd_verify_prefix(1, _capi.DTYPE_BF16, 0, 0, 1, _capi.ID_I64, 0, 1, None, None, None, None, 1, 1 << 30,
                              2, 7, 4, 100, None, None, None)
    assert rc == -2
    rc = lib.sd_verify_prefix(None, _capi.DTYPE_BF16, 0, 0, 1, _capi.ID_I64, 0, 1, None, None, None, None, 1, 1 << 30,
                              2, 5, 4, 100, None, None, None)
    assert rc == -1
    assert lib.sd_verify_set_tuning(100, 4) == -1  # threads must be a multiple of 32
    assert lib.sd_verify_set_tuning(256, 17) == -1
    assert lib.sd_verify_set_tuning(0, 0) == 0
    assert lib.sd_kv_set_variant(7) == -1
    # capacity check on host-known sizes
    rc = lib.sd_kv_append(1, 1, 1, 1, 5, 256, 8, 0, 0, 0, 0, 0, None, None, None, 0, None, 1, 4, None, 1, 0, None, None, None, None, None)
    assert rc == -6


def test_attention_and_sampler_argument_errors_without_a_gpu():
    """Host-known argument errors of kernels 3 and 4 and of the per-layer append are rejected before any CUDA call
    (include/sd_verify.h: sd_status), so they can be pinned here: fake, suitably aligned, never dereferenced pointers."""
    lib = _capi.lib()
    P, BF16, F32 = 0x10000, _capi.DTYPE_BF16, 0  # SD_F32 = 0

    def attn(q=P, k=P, v=P, lens=P, out=P, dtype=BF16, B=2, Hq=8, Hkv=2, R=5, D=128, sb=2 * 64 * 256, sh=64 * 256,
             max_len=-1, cap=0, splits=0):
        return lib.sd_verify_attention(q, k, v, lens, out, None, 0, dtype, B, Hq, Hkv, R, D, sb, sh, max_len, cap, None,
                                       0.0, splits, None)

    assert attn(Hq=0) == -1 and attn(R=0) == -1 and attn(B=-1) == -1 and attn(splits=-7) == -1
    assert attn(Hq=7, Hkv=2) == -2                       # query heads must be a multiple of KV heads
    assert attn(dtype=F32) == -5 and attn(D=96) == -5    # f16 / bf16, head_dim 64 or 128 only
    assert attn(B=0) == 0                                # an empty batch is a no-op
    assert attn(q=None) == -1 and attn(lens=None) == -1
    assert attn(k=P + 8) == -4 and attn(sh=64 * 256 + 2) == -4 and attn(q=P + 4) == -4
    assert attn(sh=4 * 256, R=5) == -6                   # 4 rows between two heads < R offered rows
    assert attn(max_len=60, R=5) == -6                   # host-known length + R beyond the 64-row capacity
    assert attn(cap=1000) == -6                          # stated capacity does not fit the head stride

    def sample(logits=P, out=P, dtype=BF16, uniform=None, T=1.0, row_index=None, row=0, B=4, rows=5, V=1000):
        return lib.sd_sample_bonus(logits, dtype, rows * V, V, row_index, 0, row, uniform, T, 0, 1.0, out, None, B, rows, V, None)

    assert sample(rows=0) == -1 and sample(V=0) == -1 and sample(B=-1) == -1
    assert sample(B=0) == 0
    assert sample(logits=None) == -1 and sample(out=None) == -1 and sample(dtype=9) == -1
    assert sample(uniform=P, T=0.0) == -1                # sampling needs a positive temperature (pipeline.py:91-92)
    assert sample(row=5) == -2 and sample(row=-1) == -2  # constant row outside [0, rows)

    # per-layer append: null arenas / rows, negative base length
    def layer(ak=P, av=P, nk=P, nv=P, R=4, base_len=0):
        return lib.sd_kv_append_layer(ak, av, nk, nv, 1, 1, R, 256, 64, 64 * 256, 64 * 256, R * 256, R * 256, 256, base_len,
                                      None, 0, None, 0, P, None)

    assert layer(ak=None) == -1 and layer(nk=None) == -1 and layer(base_len=-1) == -1
    assert layer(base_len=62) == -6                      # 62 + 4 rows beyond the 64-row capacity


def test_fused_step_argument_errors_without_a_gpu():
    """sd_verify_append (the whole step in one launch): the checks on host-known fields of sd_step_args."""
    import ctypes as C

    lib = _capi.lib()
    P = 0x10000

    def args(**kw):
        a = _capi.StepArgs()
        a.logits, a.logit_dtype, a.id_dtype = P, _capi.DTYPE_BF16, _capi.ID_I64
        a.B, a.rows, a.K, a.V = 4, 5, 4, 1000
        a.stride_b, a.stride_r, a.id_stride_b = 5 * 1000, 1000, 4
        a.draft_ids = a.accept_len = a.accepted_mask = a.target_ids = a.bonus = a.packed = a.workspace = P
        a.workspace_bytes = 1 << 20
        a.layers, a.n_layers, a.H, a.zero_rejected = P, 2, 2, 1
        a.row_bytes, a.capacity = 256, 64
        a.arena_stride_b, a.arena_stride_h = 2 * 64 * 256, 64 * 256
        a.new_stride_b, a.new_stride_h, a.new_stride_r = 2 * 5 * 256, 5 * 256, 256
        a.max_seq_len_host = -1
        a.seq_lens = a.ticket = a.step_state = P
        for k, v in kw.items():
            setattr(a, k, v)
        return a

    call = lambda a: lib.sd_verify_append(C.byref(a), None)
    assert lib.sd_verify_append(None, None) == -1
    assert call(args(rows=4)) == -2                       # the step needs the K + 1 rows of the teacher-forced pass
    assert call(args(accept_len=None)) == -1 and call(args(ticket=None)) == -1
    assert call(args(step_state=None)) == -1 and call(args(step_state=P + 4)) == -1   # 8-byte sequence words
    assert call(args(exchange=P, packed=None)) == -1      # an exchange needs the packed records
    assert call(args(B=0)) == 0                           # an empty batch is a no-op
    assert call(args(max_seq_len_host=60)) == -6          # 60 + 5 offered rows beyond the 64-row capacity
    assert lib.sd_step_state_bytes(64) == 64 * 8 + 8 and lib.sd_step_state_bytes(0) == 8


def test_start_gate_clock_rule_on_the_host():
    """The start gate's bounded wait (csrc/sd_exchange.cu: gate_clock_step, the same inline function on host and
    device): a single step of the clock -- forwards by more than the budget, or backwards -- must not open the gate
    (profiles/r2j_gate_probe.md); a wait that really outlasts the budget twice over must."""
    import ctypes as C

    lib = _capi.lib()
    budget = 2_000_000_000
    base = 1_760_000_000_000_000_000  # %globaltimer counts ns since the epoch

    def gives_up_at(readings):
        arr = (C.c_uint64 * len(readings))(*readings)
        return lib.sd_selftest_gate_clock(arr, len(readings), budget)

    us = 1000
    steady = [base + i * 40 * us for i in range(1000)]
    assert gives_up_at(steady) == -1
    # what the probe saw: the first reading lags by seconds, the following ones are consistent
    assert gives_up_at([base - 5 * budget] + steady) == -1
    # a backwards step, then steady
    assert gives_up_at([base + 10 * budget] + steady) == -1
    # one forward step in the middle of a long wait
    assert gives_up_at(steady[:500] + [t + 3 * budget for t in steady[500:]]) == -1
    # a dead peer: readings 0.5 s apart -> gives up once the budget has been exceeded twice (t0 at reading 0, first
    # strike at reading 4 = 2 s later, second strike at reading 8)
    dead = [base + i * (budget // 4) for i in range(40)]
    assert gives_up_at(dead) == 8
    assert gives_up_at([]) == -1 and lib.sd_selftest_gate_clock(None, 3, budget) == -2


def test_sass_is_sm100a_and_uses_128bit_loads_and_tma():
    """cuobjdump evidence checked here so a wrong -gencode is caught on CPU."""
    out = subprocess.run(["cuobjdump", "-lelf", str(build.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", str(build.LIB_PATH)], capture_output=True, text=True).stdout
    import re

    assert re.search(r"LDG\.E[.A-Z]*\.128", sass), "128-bit global loads expected in the verify kernel"
    assert "HMNMX2.BF16_V2.NAN" in sass  # packed NaN-propagating bf16x2 max
    assert "REDUX.MAX" in sass           # warp reduce in hardware
    assert "UBLKCP" in sass  # cp.async.bulk (TMA 1-D) in the kv_append bulk variant


def test_ctypes_mirrors_match_the_c_struct_layout():
    """The ctypes Structures of kernels/_capi.py must have the layout the compiler gave the C structs."""
    import ctypes as C

    lib = _capi.lib()
    assert lib.sd_struct_layout(0) == 4 * C.sizeof(C.c_void_p)
    assert lib.sd_struct_layout(1) == C.sizeof(_capi.ExchangeDesc)
    assert lib.sd_struct_layout(2) == _capi.ExchangeDesc.peer.offset
    assert lib.sd_struct_layout(3) == C.sizeof(_capi.StepArgs)
    assert lib.sd_struct_layout(4) == _capi.StepArgs.layers.offset
    assert lib.sd_struct_layout(5) == _capi.StepArgs.row_bytes.offset
    assert lib.sd_struct_layout(6) == _capi.StepArgs.seq_lens.offset
    assert lib.sd_struct_layout(7) == _capi.StepArgs.step_state.offset
    assert lib.sd_struct_layout(99) == 0
