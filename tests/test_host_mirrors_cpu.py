"""CPU tests of the host-side drop-ins (no kernel launches): API surface, metadata bookkeeping and the paths
that need no device.  The reference's own tests for these classes: tests/test_phase4a_optimizations.py:184-250
(KVCache), tests/specdec/test_policies.py (policy construction / info)."""
import pytest
import torch

from scheduler.batch_shard import BatchShard, advance_sequences
from specdec.cache.kv_cache_manager import SafeKVCacheManager
from specdec.cache.kv_types import KVCache, validate_kv_compatibility
from specdec.policies import create_policy
from specdec.policies.policies import (ConfidenceThresholdPolicy, LongestPrefixPolicy, TopKAgreementPolicy,
                                       TypicalAcceptancePolicy)
from specdec.utils import graph_capture


def test_policy_factory_names_and_info():
    assert isinstance(create_policy("longest_prefix"), LongestPrefixPolicy)
    assert create_policy("conf_threshold", tau=0.7).get_info() == {"policy": "conf_threshold", "tau": 0.7}
    assert create_policy("topk_agree", k=3).get_info() == {"policy": "topk_agree", "k": 3}
    assert create_policy("typical", p=0.2).get_info() == {"policy": "typical", "p": 0.2}
    assert isinstance(create_policy("conf_threshold"), ConfidenceThresholdPolicy)
    assert isinstance(create_policy("topk_agree"), TopKAgreementPolicy)
    assert isinstance(create_policy("typical"), TypicalAcceptancePolicy)
    with pytest.raises(ValueError, match="Unknown policy"):
        create_policy("best_of_n")


def test_token_comparison_fallback_needs_no_device():
    """reference: policies.py:181-196 -- without logits the policies compare token ids."""
    prop = torch.tensor([[4, 5, 6, 7], [1, 2, 3, 4]])
    base = torch.tensor([[4, 5, 9, 7], [1, 2, 3, 4]])
    out = create_policy("longest_prefix").accept_batch(prop, base_tokens=base)
    assert out["accept_len"].tolist() == [2, 4]
    assert out["accepted_mask"].tolist() == [[1, 1, 0, 0], [1, 1, 1, 1]]
    n, info = create_policy("typical").accept_tokens(prop[:1], base[:1])  # no logits -> longest prefix (:353-356)
    assert n == 2 and info["proposed_len"] == 4


def test_cpu_logits_are_refused_not_computed():
    with pytest.raises(RuntimeError, match="CUDA"):
        create_policy("longest_prefix").accept_batch(torch.zeros(1, 2, dtype=torch.long), base_logits=torch.zeros(1, 2, 8))


def test_kvcache_container_on_cpu():
    kv = tuple((torch.randn(2, 3, 7, 4), torch.randn(2, 3, 7, 4)) for _ in range(2))
    c = KVCache.from_hf_output(kv)
    assert (c.seq_len, c.dtype, c.device.type, c.get_num_layers()) == (7, torch.float32, "cpu", 2)
    s = c.slice_prefix(3)
    assert s.get_shapes() == ((2, 3, 3, 4), (2, 3, 3, 4)) and s.to(torch.device("cpu")) is s
    validate_kv_compatibility(c, s)
    other = KVCache.from_hf_output(tuple((k.half(), v.half()) for k, v in kv))
    with pytest.raises(ValueError, match="Dtype"):
        validate_kv_compatibility(c, other)


def test_cache_manager_metadata():
    m = SafeKVCacheManager("cpu")
    m.set_batch_size(3)
    m.set_sequence_metadata([10, 11, 12], [5, 6, 7])
    assert m.batch_to_global_map == {0: 10, 1: 11, 2: 12}
    assert m.get_sequence_lengths([12, 10, 99]) == [7, 5, 0]
    m.realign_kv_cache([10, 11, 12], [4, 4, 2], "draft")  # no cache yet: only the lengths move (reference :385-389)
    assert m.get_sequence_lengths([10, 11, 12], "draft") == [4, 4, 2] and m.get_sequence_lengths([10], "base") == [5]
    with pytest.raises(ValueError):
        m.set_sequence_metadata([1, 2], [3])
    m.set_active_indices([0, 2])
    assert m.active_indices.tolist() == [0, 2]
    assert m.get_base_past_kv() is None and m.get_kv_shapes()["cache_exists"] is False
    t = torch.arange(12.0).view(3, 4)
    assert torch.equal(m.filter_cache_layer(t), t[[0, 2]])
    m.set_batch_size(5)  # a batch-size change resets everything (reference :51-55)
    assert m.base_sequence_lengths == {} and m.active_indices is None


def test_graph_capture_disabled_without_cuda():
    g = graph_capture.GraphCapture(enabled=True)
    assert g.enabled == torch.cuda.is_available()
    if not torch.cuda.is_available():
        assert not g.capture_once(lambda x: x, None) and not g.is_ready()
        with pytest.raises(RuntimeError):
            g.replay()
    assert graph_capture.can_capture() == torch.cuda.is_available()
    graph_capture.reset()
    assert not graph_capture.is_ready()


def test_advance_sequences_with_eos_and_capacity():
    """pipeline.py:3059 / :3120-3131 / :3572-3608 as tensor ops: accepted tokens + bonus, EOS truncation, buffer end."""
    tokens = torch.zeros(3, 8, dtype=torch.long)
    lengths = torch.tensor([2, 6, 7])
    packed = torch.tensor([[2, 11, 12, 13, 14, 15],     # accept 2 -> tokens 11 12 + bonus 13
                           [4, 21, 99, 23, 24, 25],     # EOS (99) at the second new token
                           [1, 31, 32, 33, 34, 35]], dtype=torch.int32)  # only one slot left in the buffer
    new_len, n_new = advance_sequences(tokens, lengths, packed, eos_id=99)
    assert n_new.tolist() == [3, 2, 1] and new_len.tolist() == [5, 8, 8]
    assert tokens[0, 2:5].tolist() == [11, 12, 13] and tokens[1, 6:8].tolist() == [21, 99] and tokens[2, 7].item() == 31
    assert BatchShard(10, 4, 3).bounds() == (8, 10)
