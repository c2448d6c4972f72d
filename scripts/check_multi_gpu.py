"""N-rank scoring vs the same protocol emulated on one GPU (run under torchrun on an N-GPU box).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/check_multi_gpu.py

Every rank scores its round-robin share of the micro-batches (scorer.shard_plan); each step's dgamma vector is averaged
over the ranks with one NCCL all-reduce before the sign gate; at the end ONE all-reduce combines the class-statistics
arena.  Rank 0 then replays the identical protocol alone (the N micro-batches of a step one after the other, dgamma
averaged on the device) and compares: EIC scores (tolerance: cuDNN's backward is not bit-reproducible), keep masks at
global_percent 0.5, and the all-reduced class statistics against the sum of the per-micro-batch statistics."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from dcfp_b200 import ops
from dcfp_b200.scorer import CalibrationRun, score_calibration_set, shard_plan
from dcfp_b200.workloads.segnets import build_segnet
from dcfp_b200.workloads.synthetic import synthetic_batch

K, H, W, MB = 19, 256, 512, 2
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
N_IMG = max(16, 4 * MB * world)  # at least 4 steps at every world size (after one step half of the scores are exactly 0)
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
torch.backends.cudnn.allow_tf32 = False  # IEEE fp32 convolutions: smaller run-to-run differences for the comparison
model = build_segnet("deeplabv3", "resnet50", K, seed=0).to(dev).to(memory_format=torch.channels_last)
x, y = synthetic_batch(list(range(N_IMG)), K, H, W)
out = score_calibration_set(model, x, y, K, micro_batch=MB, r=0.999, return_class_stats=True, seed=0)
eic_dist = torch.cat([v for v in out["eic"].values()]).numpy()
stats_dist = out["class_stats"]
if rank == 0:
    plan_steps = N_IMG // (MB * world)
    run = CalibrationRun(model, K, r=0.999, seed=0, keep_totals=True)
    sc = run.scorer
    for s in range(plan_steps):
        acc = None
        for r in range(world):
            lo, hi = shard_plan(N_IMG, MB, world, r)[s]
            xb = x[lo:hi].to(dev).contiguous(memory_format=torch.channels_last)
            yb = y[lo:hi].to(dev)
            torch.manual_seed(0 + lo // MB)
            sc.set_labels(yb)
            model.zero_grad(set_to_none=True)
            model(xb, yb.long(), deepsup=True)["loss"].backward()
            dg = sc.fold_step()
            acc = dg if acc is None else acc + dg
        sc.eic_step(acc / world)
    eic_one = sc.eic.cpu().numpy()
    close = np.abs(eic_dist - eic_one) <= 1e-3 * np.abs(eic_one) + 1e-3 * np.abs(eic_one).mean()
    print("EIC: %d channels, %.4f %% within 1e-3 of the single-GPU replay, max abs diff %.3g (mean |eic| %.3g)" %
          (eic_one.size, 100 * close.mean(), np.abs(eic_dist - eic_one).max(), np.abs(eic_one).mean()))
    assert close.mean() > 0.995
    # masks at global_percent 0.5 through K2 on both score vectors
    sizes = [m.weight.numel() for _, m in sc.layers]
    offs = torch.tensor(np.concatenate([[0], np.cumsum(sizes)]), dtype=torch.int32, device=dev)
    groups = torch.tensor([0 if n.startswith("backbone") else 1 for n in sc.names], dtype=torch.int32, device=dev)
    mk = torch.tensor([max(int(c * 0.02), 1) for c in sizes], dtype=torch.int32, device=dev)
    k = [int(sum(c for c, n in zip(sizes, sc.names) if n.startswith("backbone") == (g == 0)) * 0.5) for g in (0, 1)]
    m1, t1, _ = ops.thresh_mask(torch.from_numpy(eic_dist).to(dev), offs, groups, mk, k[0], k[1])
    m2, t2, _ = ops.thresh_mask(torch.from_numpy(eic_one).to(dev), offs, groups, mk, k[0], k[1])
    same = (m1 == m2).float().mean().item()
    print("masks @0.5: %.4f %% of channels agree; thresholds %s vs %s" % (100 * same, t1.tolist(), t2.tolist()))
    assert same > 0.995
    # threshold margins: how far the nearest score is from each threshold, and -- if a keep bit differs
    # between the N-rank pass and the single-GPU replay -- how close to its threshold that channel sits
    from dcfp_b200.pruners.margin import compare_masks, format_margins
    off_h, grp_h = offs.cpu().tolist(), groups.cpu().tolist()
    cmp = compare_masks(eic_dist, eic_one, off_h, grp_h, 0.5)
    print("margin min|score - thresh|/thresh  N ranks: %s | replay: %s" % (format_margins(cmp["margins_a"]), format_margins(cmp["margins_b"])))
    print("keep bits that differ: %d of %d; they lie within %.3g of their threshold (relative); score discrepancy near the "
          "thresholds: %.3g" % (cmp["flipped"], cmp["n"], cmp["flip_band"], cmp["disc_near"]))
    assert cmp["flip_band"] <= max(4 * cmp["disc_near"], 1e-6), "a mask bit flipped outside the band the score discrepancy explains"
    # the all-reduced class statistics: sum over ALL micro-batches of S1 == totals of the replay
    tot1 = sc.totals[0].cpu()
    name = sc.names[7]
    a, b = sc.offsets[7], sc.offsets[8]
    ref = tot1[:K, a:b]  # class rows; row K holds the pixels outside [0, K)
    got = stats_dist[name][0]
    scale = ref.abs().mean()
    err = (got - ref).abs().max().item()
    print("class stats of %s: max abs diff %.3g at scale %.3g" % (name, err, scale))
    assert err <= 5e-2 * scale
    cnt_d = sum(v.sum().item() for v in out["class_counts"].values())
    cnt_1 = sc.cnt.sum().item()
    assert cnt_d == cnt_1, (cnt_d, cnt_1)
    print("pixel counts agree exactly: %d" % cnt_1)
    run.close()
    print("multi-GPU check OK (world %d)" % world)
dist.barrier()
dist.destroy_process_group()
