"""GPU parity of the arena-backed SafeKVCacheManager / KVCache (drop-ins for src/specdec/cache/*) against
the reference's semantics: growth == torch.cat along dim 2, subset filtering == index_select, rollback ==
the reference's realign outputs stored in tests/golden/kv_realign.npz.  Copies -> bit-exact."""
import numpy as np
import pytest
import torch

from specdec.cache.kv_cache_manager import SafeKVCacheManager
from specdec.cache.kv_types import KVCache, validate_kv_compatibility
from helpers import TORCH_DT, to_np

pytestmark = pytest.mark.gpu


def _kv(layers, B, H, S, D, dtype, seed):
    g = torch.Generator().manual_seed(seed)
    return tuple((torch.randn(B, H, S, D, generator=g).to(TORCH_DT[dtype]).cuda(),
                  torch.randn(B, H, S, D, generator=g).to(TORCH_DT[dtype]).cuda()) for _ in range(layers))


def _cat(a, b):
    return tuple((torch.cat([ka, kb], 2), torch.cat([va, vb], 2)) for (ka, va), (kb, vb) in zip(a, b))


def _same(got, want):
    assert len(got) == len(want)
    for (gk, gv), (wk, wv) in zip(got, want):
        assert gk.shape == wk.shape
        assert torch.equal(gk, wk) and torch.equal(gv, wv)


@pytest.mark.parametrize("dtype", ["bf16", "fp32"])
def test_growth_matches_torch_cat_and_regrows(dtype):
    layers, B, H, D = 3, 4, 2, 64
    m = SafeKVCacheManager("cuda", capacity=32)
    m.set_batch_size(B)
    want = None
    for step, S in enumerate([7, 5, 1, 40, 3]):  # 40 forces a regrowth of the arena
        new = _kv(layers, B, H, S, D, dtype, seed=step)
        m.update_base_cache(new, list(range(B)))
        m.update_draft_cache(new)
        want = new if want is None else _cat(want, new)
        _same(m.get_base_past_kv(torch.arange(B, device="cuda")), want)
        _same(m.get_draft_past_kv(torch.arange(B, device="cuda")), want)
    shapes = m.get_kv_shapes("base")
    assert shapes["cache_exists"] and shapes["num_layers"] == layers
    assert shapes["shapes"]["layer_0"]["key_shape"] == [B, H, 56, D]


def test_subset_filtering_then_append():
    """reference: kv_cache_manager.py:202-217 -- new rows for the active subset only: the cache is filtered
    to that subset (index_select) and then grows."""
    layers, B, H, D = 2, 5, 2, 32
    m = SafeKVCacheManager("cuda")
    m.set_batch_size(B)
    first = _kv(layers, B, H, 6, D, "bf16", 1)
    m.update_base_cache(first, list(range(B)))
    active = [0, 2, 3]
    idx = torch.tensor(active, device="cuda")
    _same(m.get_base_past_kv(idx), tuple((k.index_select(0, idx), v.index_select(0, idx)) for k, v in first))
    nxt = _kv(layers, len(active), H, 4, D, "bf16", 2)
    m.update_base_cache(nxt, active)
    want = _cat(tuple((k.index_select(0, idx), v.index_select(0, idx)) for k, v in first), nxt)
    _same(m.get_base_past_kv(torch.arange(len(active), device="cuda")), want)
    assert m.get_base_past_kv(torch.tensor([], dtype=torch.long, device="cuda")) is None
    assert m.filter_cache_layer(first[0][0], torch.tensor([9], device="cuda")) is None  # out of range -> None (:100-107)


def test_realign_matches_reference_outputs(golden_dir):
    z = np.load(golden_dir / "kv_realign.npz")
    for name in z["__cases__"]:
        name = str(name)
        k, v = torch.from_numpy(z[f"{name}.k"]).cuda(), torch.from_numpy(z[f"{name}.v"]).cuda()
        new_lengths = [int(n) for n in z[f"{name}.new_lengths"]]
        B = k.shape[0]
        m = SafeKVCacheManager("cuda")
        m.set_batch_size(B)
        m.set_sequence_metadata(list(range(100, 100 + B)), [k.shape[2]] * B)
        m.update_base_cache(((k, v), (v, k)), list(range(B)))
        m.realign_kv_cache(list(range(100, 100 + B)), new_lengths, "base")
        got = m.get_base_past_kv(torch.arange(B, device="cuda"))
        np.testing.assert_array_equal(to_np(got[0][0]), z[f"{name}.rk"], err_msg=name)
        np.testing.assert_array_equal(to_np(got[0][1]), z[f"{name}.rv"], err_msg=name)
        np.testing.assert_array_equal(to_np(got[1][0]), z[f"{name}.rv"], err_msg=name)
        assert m.get_sequence_lengths(list(range(100, 100 + B))) == new_lengths
        # growth after a rollback continues at the new length
        S = 2
        new = _kv(2, B, k.shape[1], S, k.shape[3], "fp32", 5)
        new = tuple((a.to(k.dtype), b.to(k.dtype)) for a, b in new)
        m.update_base_cache(new, list(range(B)))
        got2 = m.get_base_past_kv(torch.arange(B, device="cuda"))
        assert got2[0][0].shape[2] == max(new_lengths) + S
        assert torch.equal(got2[0][0][:, :, max(new_lengths):], new[0][0])


def test_kvcache_container():
    """reference: kv_types.py:33-156 and tests/test_phase4a_optimizations.py:184-250."""
    kv = _kv(2, 1, 2, 9, 16, "fp32", 3)
    c = KVCache.from_hf_output(kv)
    assert c.seq_len == 9 and c.get_num_layers() == 2 and c.get_shapes()[0] == (1, 2, 9, 16)
    s = c.slice_prefix(4)
    assert s.seq_len == 4 and s.past_key_values[0][0].data_ptr() == kv[0][0].data_ptr()  # a view
    with pytest.raises(ValueError):
        c.slice_prefix(10)
    with pytest.raises(ValueError):
        KVCache.from_hf_output(())
    validate_kv_compatibility(c, s)
    with pytest.raises(ValueError):
        validate_kv_compatibility(c, KVCache.from_hf_output(_kv(3, 1, 2, 9, 16, "fp32", 4)))

    class _Layer:  # transformers >= 4.36 DynamicCache shape: .layers[i].keys / .values
        def __init__(self, k, v):
            self.keys, self.values = k, v

    class _Dyn:
        def __init__(self, kv):
            self.layers = [_Layer(k, v) for k, v in kv]

        def __len__(self):
            return len(self.layers)

    d = KVCache.from_hf_output(_Dyn(kv))
    assert d.seq_len == 9 and torch.equal(d.past_key_values[1][1], kv[1][1])
