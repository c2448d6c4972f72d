"""N > 1 host logic on CPU: world_size-2 `gloo` run of the sharding plan + per-step rank averaging, checked
against the single-process oracle following the same protocol (oracle/scoring_ref.py, world=2)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from dcfp_b200.scorer import average_over_ranks, shard_plan

K, H, W, N_IMG, MB = 19, 32, 64, 8, 2


def test_shard_plan_is_rank_count_invariant():
    ref = [tuple(range(lo, hi)) for lo, hi in shard_plan(24, 2, 1, 0)]
    for world in (2, 3, 4):
        got = {}
        for rank in range(world):
            for step, (lo, hi) in enumerate(shard_plan(24, 2, world, rank)):
                got[step * world + rank] = tuple(range(lo, hi))
        assert [got[i] for i in sorted(got)] == ref[:len(got)]
        assert len(got) == (24 // (2 * world)) * world
    assert shard_plan(3, 2, 2, 0) == []
    with pytest.raises(ValueError):
        shard_plan(8, 2, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _batches():
    from dcfp_b200.workloads.synthetic import synthetic_batch
    return [synthetic_batch(list(range(i, i + MB)), K, H, W) for i in range(0, N_IMG, MB)]


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(1)
        from dcfp_b200.workloads.segnets import build_segnet
        from dcfp_b200.workloads.synthetic import synthetic_batch
        from oracle import eic_ref, scoring_ref
        model = build_segnet("deeplabv3", "resnet50", K, seed=0)
        model.train()
        layers = scoring_ref.scored_bn_layers(model)
        eic = {n: 0 for n, _ in layers}
        for lo, hi in shard_plan(N_IMG, MB, world, rank):
            x, y = synthetic_batch(list(range(lo, hi)), K, H, W)
            torch.manual_seed(77 + lo // MB)                            # dropout keyed by the GLOBAL micro-batch index
            grads, _ = scoring_ref.gamma_grads(model, x, y)           # stand-in for K1 on this rank's micro-batch
            flat = torch.cat([grads[n] for n, _ in layers])
            average_over_ranks(flat)                                    # product code path: one all-reduce per step
            pos = 0
            for n, m in layers:
                c = m.weight.numel()
                eic[n] = eic_ref.eic_step(eic[n], flat[pos:pos + c].numpy(), m.weight.detach().numpy(), 0.999)
                pos += c
        # end-of-pass combine of a statistics arena: ONE all-reduce (SUM, fp64)
        arena = torch.full((5,), float(rank + 1), dtype=torch.float64)
        dist.all_reduce(arena)
        if rank == 0:
            torch.save({"eic": eic, "arena": arena}, out)
    finally:
        dist.destroy_process_group()


def test_two_rank_scoring_matches_single_process_oracle(tmp_path):
    from dcfp_b200.workloads.segnets import build_segnet
    from oracle import scoring_ref
    out = str(tmp_path / "rank0.pt")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = torch.load(out, weights_only=False)
    assert torch.equal(got["arena"], torch.full((5,), 3.0, dtype=torch.float64))
    threads = torch.get_num_threads()
    torch.set_num_threads(1)
    try:
        model = build_segnet("deeplabv3", "resnet50", K, seed=0)
        exp, _ = scoring_ref.score(model, _batches(), r=0.999, world=2, seed=77)
    finally:
        torch.set_num_threads(threads)
    for n, v in exp["eic"].items():
        g = got["eic"][n]
        # (a + b) / 2 is computed in the same order by gloo's 2-rank sum and by the oracle: bit-exact
        assert np.array_equal(np.asarray(g).view(np.uint32), v.view(np.uint32)), n
    assert average_over_ranks(torch.ones(3)).tolist() == [1.0, 1.0, 1.0]  # no process group: no-op
