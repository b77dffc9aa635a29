"""CPU: batch sharding + the packed-record all-gather over gloo (world_size 2), and the batched
sequence advance against a per-sequence restatement of the reference's host loop."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import specdec_oracle as O
from scheduler.batch_shard import BatchShard, StepResultExchange, advance_sequences


def test_bounds_cover_batch_evenly():
    for B in (0, 1, 7, 64, 512, 513):
        for W in (1, 2, 4, 8):
            spans = [BatchShard(B, W, r).bounds() for r in range(W)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1
    assert BatchShard(512, 8, 3).bounds() == (192, 256)
    with pytest.raises(ValueError):
        BatchShard(8, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, B, K, V, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1234)
        logits = torch.randn(B, K + 1, V, generator=g)
        ids = torch.argmax(logits[:, :K], -1)
        ids[::3, 1] += 1
        shard = BatchShard(B, world, rank)
        # each rank verifies only its shard (oracle stands in for the CUDA kernel on CPU)
        acc, _, tgt, _ = O.verify_prefix_bonus(shard.take(logits).numpy(), shard.take(ids).numpy())
        packed_local = torch.from_numpy(O.pack_step_result(acc, tgt))
        ex = StepResultExchange(shard, K, "cpu")
        packed = ex.all_gather(packed_local)
        full_acc, _, full_tgt, _ = O.verify_prefix_bonus(logits.numpy(), ids.numpy())
        want = O.pack_step_result(full_acc, full_tgt)
        np.save(os.path.join(out_dir, f"ok{rank}.npy"), np.array([int(np.array_equal(packed.numpy(), want))]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_allgather_of_packed_records_gloo_world2(tmp_path, B):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, B, 4, 257, str(tmp_path)), nprocs=2, join=True)
    for r in range(2):
        assert np.load(tmp_path / f"ok{r}.npy")[0] == 1


def test_advance_sequences_matches_per_sequence_loop():
    g = torch.Generator().manual_seed(3)
    B, K, T = 9, 4, 40
    tokens = torch.randint(5, 100, (B, T), generator=g)
    lengths = torch.randint(3, 20, (B,), generator=g)
    acc = torch.randint(0, K + 1, (B,), generator=g)
    tgt = torch.randint(5, 100, (B, K + 1), generator=g)
    tgt[2, 1] = 2  # EOS inside the accepted run
    tgt[5, 0] = 2
    packed = torch.cat([acc[:, None], tgt], 1).int()
    for eos in (-1, 2):
        want = tokens.clone()
        want_len = lengths.clone()
        for b in range(B):  # restatement of pipeline.py:3059-3608 for one sequence
            new = tgt[b, : int(acc[b]) + 1].tolist()
            if eos >= 0 and eos in new:
                new = new[: new.index(eos) + 1]
            want[b, int(lengths[b]) : int(lengths[b]) + len(new)] = torch.tensor(new)
            want_len[b] += len(new)
        got = tokens.clone()
        new_len, n_new = advance_sequences(got, lengths.clone(), packed, eos_id=eos)
        assert torch.equal(got, want) and torch.equal(new_len, want_len)
        assert torch.equal(n_new, want_len - lengths)
