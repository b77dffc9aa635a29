#!/usr/bin/env python3
"""Multi-process check of the fused peer-memory exchange (run under torchrun on >= 2 GPUs):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29533 tools/check_peer_exchange.py

Every rank verifies its shard with VerifyStep(exchange=PeerExchange) -- eagerly and as a replayed CUDA
graph -- and compares the gathered records with a NCCL all-gather of the same packed tensors.  Also
times both exchanges (device events, max over ranks)."""
import json
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT / "llm-inference-lab_b200"), str(ROOT)]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from scheduler.batch_shard import BatchShard, PeerExchange, StepResultExchange  # noqa: E402
from specdec.verify_step import VerifyStep  # noqa: E402


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    B, K, V = 64, 4, 128256
    shard = BatchShard(B * world, world, rank)
    peer = PeerExchange(shard, K, dev)
    nccl = StepResultExchange(shard, K, dev)
    st = VerifyStep(B, K, V, device=dev, exchange=peer)
    st.exchange_wait = False  # this script issues peer.wait() / the fused append itself
    g = torch.Generator(device=dev).manual_seed(1234 + rank)
    bad = 0
    for it in range(6):
        st.logits.copy_(torch.randn(B, K + 1, V, generator=g, device=dev).to(torch.bfloat16))
        st.draft_ids.copy_(torch.argmax(st.logits[:, :K].float(), -1))
        st.draft_ids[it % B:, (it + rank) % K:] = 7  # break the prefix at different places per rank/step
        st.run()
        got = peer.wait().clone()
        want = nccl.all_gather(st.packed).clone()
        torch.cuda.synchronize()
        if not torch.equal(got, want):
            bad += 1
            print(f"[rank {rank}] step {it}: MISMATCH", flush=True)
    # graph replay
    side = torch.cuda.Stream(dev)
    gr = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        with torch.cuda.graph(gr, stream=side):
            st.run()
            peer.wait()
    for it in range(5):
        st.draft_ids[:, (it + rank) % K] = 11 + it
        gr.replay()
        got = peer.out.clone()
        want = nccl.all_gather(st.packed).clone()
        torch.cuda.synchronize()
        if not torch.equal(got, want):
            bad += 1
            print(f"[rank {rank}] graph step {it}: MISMATCH", flush=True)

    def timed(fn, n=200):
        with torch.cuda.stream(side):
            g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g2, stream=side):
                for _ in range(4):
                    fn()
        torch.cuda.synchronize()
        for _ in range(5):
            g2.replay()
        dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(n // 4):
            g2.replay()
        b.record()
        torch.cuda.synchronize()
        t = torch.tensor([a.elapsed_time(b) / n * 1e3], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def step_peer():
        st.exchange = peer
        st.run()
        peer.wait()

    def step_nccl():
        st.exchange = None
        st.run()
        nccl.all_gather(st.packed)

    def step_none():
        st.exchange = None
        st.run()

    from specdec.cache.kv_arena import KVArena

    arena = KVArena(28, B, 8, 128, 1024 + 5 * 400, K + 1, torch.bfloat16, dev)
    arena.set_lengths([1024] * B)
    lens0 = torch.full((B,), 1024, dtype=torch.int32, device=dev)

    def step_fused():  # verify (pushes records) -> kv_append (first CTA confirms + gathers)
        st.exchange = peer
        st.run()
        arena.seq_lens.copy_(lens0)
        arena.append(st.accept_len, keep_bias=1, check_capacity=False, exchange=peer)

    def step_fused_kv_push():  # verify (no exchange) -> kv_append: first CTA pushes at its start and gathers at its end
        st.exchange = None
        st.run()
        arena.seq_lens.copy_(lens0)
        arena.append(st.accept_len, keep_bias=1, check_capacity=False, exchange=peer, exchange_packed=st.packed)

    def step_plain():
        st.exchange = None
        st.run()
        arena.seq_lens.copy_(lens0)
        arena.append(st.accept_len, keep_bias=1, check_capacity=False)

    def step_plain_nccl():
        step_plain()
        nccl.all_gather(st.packed)

    t_peer, t_nccl, t_none = timed(step_peer), timed(step_nccl), timed(step_none)
    t_fused, t_plain, t_plain_nccl = timed(step_fused), timed(step_plain), timed(step_plain_nccl)
    t_fused_kv = timed(step_fused_kv_push)
    step_fused_kv_push()
    got = peer.out.clone()
    want = nccl.all_gather(st.packed).clone()
    torch.cuda.synchronize()
    if not torch.equal(got, want):
        bad += 1
        print(f"[rank {rank}] kv-side push: MISMATCH", flush=True)
    # correctness of the fused path once more after the timing loops
    st.exchange = peer
    st.run()
    arena.seq_lens.copy_(lens0)
    arena.append(st.accept_len, keep_bias=1, check_capacity=False, exchange=peer)
    got = peer.out.clone()
    want = nccl.all_gather(st.packed).clone()
    torch.cuda.synchronize()
    if not torch.equal(got, want):
        bad += 1
        print(f"[rank {rank}] fused append: MISMATCH", flush=True)
    ok = torch.tensor([bad], device=dev)
    dist.all_reduce(ok)
    if rank == 0:
        print(json.dumps({"world": world, "mismatches": int(ok.item()), "verify_only_us": t_none,
                          "verify_plus_peer_exchange_us": t_peer, "verify_plus_nccl_allgather_us": t_nccl,
                          "step_verify_kv_us": t_plain, "step_with_fused_exchange_us": t_fused,
                          "step_with_nccl_allgather_us": t_plain_nccl, "step_with_kv_side_push_us": t_fused_kv}), flush=True)
    torch.cuda.synchronize()
    sys.stdout.flush()
    os._exit(0 if int(ok.item()) == 0 else 1)


if __name__ == "__main__":
    main()
