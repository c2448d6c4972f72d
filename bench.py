#!/usr/bin/env python
"""bench.py -- DCFP scoring throughput (images/s) on B200, with the K1 roofline and the CPU reference beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one micro-batch (2 images per GPU, fixed by global image index) of
synthetic Cityscapes-shaped input: forward + backward of the random-init segmentation net (torch / cuDNN produce the
conv/BN feature maps), the label-keyed segmented reduction K1 over every scored BN layer (hand-written sm_100a kernel,
deferred into grouped launches), the end-of-step fold, the per-step all-reduce of the BN-gamma gradient (N > 1) and the
K2 EIC update.  Workload at N = 1: BASELINE.json configs[1] (DeepLabV3-ResNet101, 19 classes, 512x1024).

Printed JSON line (rank 0):
  value     images/s over all ranks, inputs already resident in HBM (CUDA events, max over ranks)
  e2e       the same through the public HOST API dcfp_b200.scorer.score_calibration_set: per step the micro-batch is
            copied host->device from pinned memory and the loss is read back; scorer set-up and the final score
            read-back are inside the timed region
  roofline  K1 only: algorithmic bytes of its launches / their CUDA-event durations inside the timed region,
            against the measured HBM copy bandwidth in MEASURED_PEAKS.json
  cpu_baseline  the oracle port of the reference's CPU path (oracle/scoring_ref.py) on this box's host cores
`--impl reference` times that CPU path alone (the reference is pure Python + torch CPU ops; there is no
compiled reference, so the arm is the oracle port, kind "port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "dcfp_scoring_images_per_s"
UNIT = "images/s"
FALLBACK_HBM_GBS = 6650.0  # B200_PROFILING.md fallback when MEASURED_PEAKS.json is absent


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4"])
    p.add_argument("--micro-batch", type=int, default=2)
    p.add_argument("--flush-mb", type=int, default=16384,
                   help="pending feature-map MiB that trigger a grouped K1 launch before the end of the backward pass "
                        "(0 = one launch per layer; the default defers a whole c2 step, 8 GB, into ONE launch)")
    p.add_argument("--prime", type=int, default=3,
                   help="set-up steps before the W warm-up steps (cuDNN autotuning, caching-allocator growth); untimed")
    p.add_argument("--conv-precision", default="tf32", choices=["fp32", "tf32"],
                   help="cuDNN convolution math of the feature-map PRODUCER (not part of the path): tf32 = torch's default "
                        "(torch.backends.cudnn.allow_tf32=True), which is what the reference's train.py runs on any Ampere+ GPU "
                        "since it never touches torch.backends; fp32 = IEEE fp32 convolutions")
    p.add_argument("--layout", default="channels_last", choices=["channels_last", "nchw"],
                   help="memory format of the model / feature maps: channels_last is cuDNN's native tensor-core layout "
                        "(no per-conv transposes) and takes K1's NHWC path; nchw takes K1's TMA-tile path")
    p.add_argument("--scores-only", action="store_true",
                   help="freeze every non-BN parameter during scoring: no weight-gradient convolutions (the scores do not "
                        "need them); default off = the reference's full backward")
    p.add_argument("--no-fused", action="store_true",
                   help="score with torch/cuDNN BatchNorm + ReLU and the hook-fed, deferred K1 launches (round 1's path) instead of "
                        "the fused BN(+ReLU) kernels whose backward yields the class-keyed sums (SURVEY 8 f1)")
    p.add_argument("--no-forward-functor", action="store_true",
                   help="skip the extra forward-only region (north_star-literal statistics: K1 with v = BN output)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--cpu-budget-s", type=float, default=240.0, help="wall-clock cap of the reference arm")
    return p.parse_args()


def hbm_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 50 ms DURING the timed region."""
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu_index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append((time.time(), line.strip()))

    def stop(self, t0=None, t1=None):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        for ts, line in self.lines:
            if t0 is not None and not (t0 <= ts <= t1 + 0.25):
                continue
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def workload(cfg_name):
    from dcfp_b200.workloads.segnets import CONFIGS
    c = dict(CONFIGS[cfg_name])
    names = {"c1": "DeepLabV3-ResNet50", "c2": "DeepLabV3-ResNet101", "c3": "PSPNet-ResNet101", "c4": "DeepLabV3+-ResNet101"}
    c["label"] = "%s random-init, %d classes, %dx%d synthetic images (BASELINE.json %s)" % (
        names[cfg_name], c["num_classes"], c["height"], c["width"], {"c1": "configs[0]", "c2": "configs[1]", "c3": "configs[2]",
                                                                      "c4": "configs[3]"}[cfg_name])
    return c


def make_batches(c, indices_per_step, pin):
    from dcfp_b200.workloads.synthetic import synthetic_batch
    out = []
    for idx in indices_per_step:
        x, y = synthetic_batch(idx, c["num_classes"], c["height"], c["width"])
        out.append((x.pin_memory(), y.pin_memory()) if pin else (x, y))
    return out


# ------------------------------------------------------------------------------------------ CPU reference arm
def cpu_reference(c, micro_batch, steps, warmup, budget_s):
    """The reference's CPU path (restated loop of train.py:255-268 over torch CPU autograd + pruners/dcfp_pruner.py:15-20
    in oracle/scoring_ref.py), all host threads.  A step = one micro-batch.  Returns (images/s, steps done, s/step)."""
    import torch

    from dcfp_b200.workloads.segnets import build_segnet
    from oracle import eic_ref, scoring_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0)
    model.train()
    layers = scoring_ref.scored_bn_layers(model)
    eic = {n: 0 for n, _ in layers}
    batches = make_batches(c, [list(range(i * micro_batch, (i + 1) * micro_batch)) for i in range(2)], pin=False)

    def one(i):
        x, y = batches[i % len(batches)]
        torch.manual_seed(i)
        grads, _ = scoring_ref.gamma_grads(model, x, y)
        for n, m in layers:
            eic[n] = eic_ref.eic_step(eic[n], grads[n].numpy(), m.weight.detach().numpy(), 0.999)

    t_begin, per, done_w = time.time(), None, 0
    for i in range(warmup):  # warm-up may use at most half of the wall-clock budget
        if per is not None and (time.time() - t_begin) + per > 0.5 * budget_s:
            break
        t = time.time()
        one(i)
        per = time.time() - t
        done_w += 1
    done, t0 = 0, time.time()
    while done < steps:  # the timed steps stop early (and say so) rather than overrun the budget
        if per is not None and done >= 1 and (time.time() - t_begin) + per > budget_s:
            break
        t = time.time()
        one(done_w + done)
        per = time.time() - t
        done += 1
    dt = time.time() - t0
    return done * micro_batch / dt, done, dt / done, cores, done_w


def run_reference_arm(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, done, s_per, cores, done_w = cpu_reference(c, args.micro_batch, args.steps, args.warmup, args.cpu_budget_s)
    sample = "%d timed micro-batches of %d images (%d warm-up) of the %s workload, fp32, %d host threads" % (
        done, args.micro_batch, done_w, args.config, cores)
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": done_w, "ms_per_step": s_per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": c["label"], "micro_batch": args.micro_batch, "protocol": "zero_grad -> loss(x, y, deepsup) -> "
                       "backward -> dcfp_pruning.step, no optimizer step", "requested_steps": args.steps,
                       "requested_warmup": args.warmup},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
def run_b200_arm(args, c):
    import torch
    import torch.distributed as dist

    from dcfp_b200 import ops
    from dcfp_b200.scorer import CalibrationRun, score_calibration_set
    from dcfp_b200.workloads.segnets import build_segnet

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, world))
    if args.gpus > 1 and world == 1:
        raise SystemExit("--gpus %d needs torchrun: python -m torch.distributed.run --nproc-per-node %d bench.py ..." % (args.gpus, args.gpus))
    ops.require_gpu()  # no CPU fallback: fail loudly
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    tf32 = args.conv_precision == "tf32"
    torch.backends.cudnn.allow_tf32 = tf32
    torch.backends.cuda.matmul.allow_tf32 = tf32
    torch.backends.cudnn.benchmark = os.environ.get("DCFP_BENCH_CUDNN_BENCHMARK", "1") == "1"

    K, W, mb = args.steps, args.warmup, args.micro_batch
    model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0).to(dev)
    nhwc = args.layout == "channels_last"
    if nhwc:
        model = model.to(memory_format=torch.channels_last)
    n_steps_total = W + K
    # global micro-batch index of (step s, rank r) = s * world + r  -- the plan of scorer.shard_plan
    idx = [list(range((s * world + rank) * mb, (s * world + rank + 1) * mb)) for s in range(n_steps_total)]
    host = make_batches(c, idx, pin=True)
    resident = [(x.to(dev).contiguous(memory_format=torch.channels_last) if nhwc else x.to(dev), y.to(dev)) for x, y in host]
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- phase A: inputs resident in HBM -------------------------------------------------------------------
    run = CalibrationRun(model, c["num_classes"], r=0.999, flush_bytes=args.flush_mb << 20, keep_totals=True, timing=True, seed=0,
                         scores_only=args.scores_only, fused=not args.no_fused)
    sc = run.scorer
    # nvidia-smi attaches to the driver when it starts (stalls launches for ~0.2 s): start it before the set-up steps
    sampler = ClockSampler(local_rank).start() if rank == 0 else None
    for s in range(args.prime):  # set-up: cuDNN benchmark autotuning + allocator growth, not part of W
        run.step(*resident[s % len(resident)], mb_index=s * world + rank)
    # cuDNN's autotuning trials leave >100 GB of workspace blocks cached; with a few GB of deferred gradients on top the
    # caching allocator would garbage-collect inside timed steps.  Release them once; steady state needs ~20 GB.
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    torch.cuda.reset_peak_memory_stats(dev)  # "allocated_peak_gb" below = warm-up + timed steps, not the autotuning trials
    # the EIC state must not see the priming steps: restart the accumulator
    sc.steps = 0
    sc.eic.zero_()
    if sc.total_arena is not None:
        sc.total_arena.zero_()
    for s in range(W):
        run.step(*resident[s], mb_index=s * world + rank)
    barrier()
    sc.k1_events.clear()
    sc.phase_events.clear()
    launches0 = ops.launch_count()
    mem0 = torch.cuda.memory_stats(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall0 = time.time()
    barrier()
    # process-wide start/end range (the backward kernels are launched from the autograd thread, which a push/pop
    # range of this thread would miss): ncu --nvtx --nvtx-include "timed" profiles exactly the timed launches
    nvtx_range = torch.cuda.nvtx.range_start("timed")
    e0.record()
    marks = [e0]
    debug = os.environ.get("DCFP_BENCH_DEBUG") == "1"
    dbg = []
    for s in range(W, W + K):
        run.step(*resident[s], mb_index=s * world + rank)
        marks.append(torch.cuda.Event(enable_timing=True))
        marks[-1].record()
        if debug:
            m = torch.cuda.memory_stats(dev)
            dbg.append((m.get("num_device_alloc", 0), m.get("num_device_free", 0), m.get("num_alloc_retries", 0),
                        round(m.get("reserved_bytes.all.current", 0) / 1e9, 2), round(time.time() - t_wall0, 3)))
    e1.record()
    torch.cuda.nvtx.range_end(nvtx_range)
    barrier()
    t_wall1 = time.time()
    step_ms = [round(a.elapsed_time(b), 3) for a, b in zip(marks[:-1], marks[1:])]
    ms_a = max_over_ranks(e0.elapsed_time(e1))
    launches = ops.launch_count() - launches0
    mem1 = torch.cuda.memory_stats(dev)
    alloc = {"cudaMalloc_calls_in_timed_region": mem1.get("num_device_alloc", 0) - mem0.get("num_device_alloc", 0),
             "cudaFree_calls_in_timed_region": mem1.get("num_device_free", 0) - mem0.get("num_device_free", 0),
             "alloc_retries_in_timed_region": mem1.get("num_alloc_retries", 0) - mem0.get("num_alloc_retries", 0),
             "debug": dbg, "reserved_gb": mem1.get("reserved_bytes.all.current", 0) / 1e9, "allocated_peak_gb": mem1.get("allocated_bytes.all.peak", 0) / 1e9}
    k1_ms, k1_bytes, k1_launches = sc.k1_time_ms()
    region_ms = max(e0.elapsed_time(e1), 1e-9)
    phases = {k: {"ms_per_step": ms / K, "algorithmic_gb_per_step": nb / K / 1e9, "calls_per_step": n / K,
                  "achieved_gbs": (nb / (ms * 1e-3) / 1e9) if ms > 0 else None, "share_of_step": ms / region_ms}
              for k, (ms, nb, n) in sc.phase_times().items()}
    fused_on = "bn_bwd_reduce" in phases
    if k1_launches:
        phases["k1_deferred"] = {"ms_per_step": k1_ms / K, "algorithmic_gb_per_step": k1_bytes / K / 1e9, "calls_per_step": k1_launches / K,
                                 "achieved_gbs": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None, "share_of_step": k1_ms / region_ms}
    if fused_on:  # the dominant kernel of the path is now B1: the class-keyed reduction inside the fused BN backward
        k1_ms, k1_bytes, k1_launches = sc.phase_times()["bn_bwd_reduce"]
    share_a = k1_ms / region_ms
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    sc.all_reduce_totals()  # the single end-of-pass statistics all-reduce (outside the per-step timing, reported below)
    if world > 1:
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        sc.all_reduce_totals()
        eb.record()
        torch.cuda.synchronize()
        allreduce_ms = max_over_ranks(ea.elapsed_time(eb))
    else:
        allreduce_ms = 0.0
    arena_bytes = sc.total_arena.numel() * 8
    sum_c = sc.total_channels
    run.close()
    del run, sc
    value = K * mb * world / (ms_a * 1e-3)

    # ---- forward functor (north_star-literal): forward pass only, K1 with v = BN output, deferred into grouped launches
    forward = None
    if not args.no_forward_functor:
        run_f = CalibrationRun(model, c["num_classes"], mode="fwd", flush_bytes=args.flush_mb << 20, keep_totals=True, timing=True, seed=0)
        for s in range(max(W, 1)):
            run_f.step(*resident[s], mb_index=s * world + rank)
        barrier()
        run_f.scorer.k1_events.clear()
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        f0.record()
        for s in range(W, W + K):
            run_f.step(*resident[s], mb_index=s * world + rank)
        f1.record()
        barrier()
        ms_f = max_over_ranks(f0.elapsed_time(f1))
        fk_ms, fk_bytes, fk_launches = run_f.scorer.k1_time_ms()
        run_f.close()
        peak_f, _ = hbm_peak()
        forward = {"what": "forward pass + class statistics of every scored BN OUTPUT (v = y, pre-ReLU), inputs resident in HBM",
                   "value": K * mb * world / (ms_f * 1e-3), "unit": UNIT, "ms_per_step": ms_f / K,
                   "k1_achieved_gbs": fk_bytes / (fk_ms * 1e-3) / 1e9 if fk_ms > 0 else None,
                   "k1_frac_of_hbm_peak": (fk_bytes / (fk_ms * 1e-3) / 1e9 / peak_f) if fk_ms > 0 else None,
                   "k1_launches": fk_launches, "k1_algorithmic_bytes_per_image": fk_bytes / (K * mb),
                   "k1_images_per_s_of_kernel_time": K * mb / (fk_ms * 1e-3) if fk_ms > 0 else None}
        del run_f
    del resident
    torch.cuda.empty_cache()

    # ---- phase B: public host API, H2D + D2H inside the timed region -----------------------------------------
    e2e = None
    if not args.no_e2e:
        # score_calibration_set shards by rank itself (shard_plan): hand it the images in GLOBAL order.  This rank
        # only ever reads its own slices, so the other ranks' slots repeat its own micro-batch as placeholders.
        def global_order(lo, hi):
            xs = [host[s][0] for s in range(lo, hi) for _ in range(world)]
            ys = [host[s][1] for s in range(lo, hi) for _ in range(world)]
            return torch.cat(xs).pin_memory(), torch.cat(ys).pin_memory()
        if W:
            xw, yw = global_order(0, W)
            score_calibration_set(model, xw, yw, c["num_classes"], micro_batch=mb, flush_bytes=args.flush_mb << 20,
                                  scores_only=args.scores_only, fused=not args.no_fused)
        xk, yk = global_order(W, W + K)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = score_calibration_set(model, xk, yk, c["num_classes"], micro_batch=mb, flush_bytes=args.flush_mb << 20,
                                    scores_only=args.scores_only, fused=not args.no_fused)
        e1.record()
        barrier()
        ms_b = max_over_ranks(e0.elapsed_time(e1))
        st = out["_stats"]
        assert st["steps"] == K
        e2e = {"value": K * mb * world / (ms_b * 1e-3), "unit": UNIT, "h2d_bytes_per_step": st["h2d_bytes"] // K,
               "d2h_bytes_per_step": st["d2h_bytes"] / K, "ms_per_step": ms_b / K,
               "api": "dcfp_b200.scorer.score_calibration_set(model, host_images, host_labels, K)"}

    peak, peak_src = hbm_peak()
    achieved = k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    roofline = {"kernel": ("dcfp::class_stats_nhwc_kernel<float, BWD, FUSED> (B1 of the fused BN backward: label-keyed segmented reduction "
                           "of v = dz * xhat with the ReLU gate recomputed, + sum dz, sum v; one launch per BN layer)") if fused_on else
                          "dcfp::%s<float, BWD> (K1, label-keyed segmented reduction, v = dy * xhat)" %
                          ("class_stats_nhwc_kernel" if nhwc else "class_stats_kernel"),
                "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "peak_source": peak_src, "traffic": traffic["dram_bytes_per_launch"] if traffic else None,
                "traffic_note": traffic.get("note") if traffic else "no ncu --set full capture recorded yet",
                "launches": k1_launches, "algorithmic_bytes_per_launch": k1_bytes / max(k1_launches, 1),
                "avg_launch_ms": k1_ms / max(k1_launches, 1), "share_of_step": share_a,
                "algorithmic_bytes_per_image": k1_bytes / (K * mb),
                "images_per_s_of_kernel_time": K * mb / (k1_ms * 1e-3) if k1_ms > 0 else None}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, done, s_per, cores, done_w = cpu_reference(c, mb, 1, 0, 120.0)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": "1 micro-batch of %d images of the same workload (%.1f s), fp32, oracle/scoring_ref.py, no warm-up" % (mb, s_per)}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": ms_a / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": c["label"], "micro_batch_per_gpu": mb, "calibration_images_nominal": 500,
                           "images_timed": K * mb * world, "sum_scored_channels": sum_c,
                           "conv_math": ("tf32 (torch default cudnn.allow_tf32=True, as the reference's train.py runs its convolutions)"
                                         if tf32 else "fp32 (cudnn.allow_tf32=False)") + "; K1/K2/K3 arithmetic is fp32 (fp64 across CTAs)",
                           "protocol": "zero_grad -> loss(x, y, deepsup) -> backward [K1 on every scored BN: S[k,c] += dy*xhat] -> "
                                       "fold -> all-reduce(dgamma)/N -> EIC update; no optimizer step",
                           "bn": ("fused: dcfp BN(+ReLU) forward / backward kernels, class-keyed sums inside the BN backward" if fused_on else
                                  "torch/cuDNN BatchNorm + ReLU, hook-fed deferred K1"),
                           "l2": "per-step feature maps (%.1f GB read by K1) exceed the 126 MB L2; no explicit flush" % (k1_bytes / K / 1e9),
                           "layout": args.layout, "backward": "scores_only (no weight-gradient convolutions)" if args.scores_only else "full (all gradients, as the reference's training step)", "k1_flush_mib": args.flush_mb, "priming_steps": args.prime, "parallelism": "dp%d (micro-batches dealt round-robin)" % world},
                "step_ms": step_ms, "allocator": alloc, "roofline": roofline, "path_phases": phases, "forward_functor": forward, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
                "stats_allreduce": {"bytes": arena_bytes, "ms": allreduce_ms, "what": "one all-reduce of the [2,K+1,sumC] fp64 totals (K classes + the pixels outside [0,K)) + counts at the end of the pass"}}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    args = parse_args()
    c = workload(args.config)
    if args.impl == "reference":
        run_reference_arm(args, c)
    else:
        run_b200_arm(args, c)


if __name__ == "__main__":
    main()
