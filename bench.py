#!/usr/bin/env python
"""bench.py -- DCFP scoring throughput (images/s) on B200, with the path's roofline and the CPU reference beside it.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--config c2] [--sweep]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A *step* is one pass of the hot path over one micro-batch (2 images per GPU, fixed by global image index) of synthetic
Cityscapes-shaped input: forward + backward of the random-init segmentation net -- convolutions by torch / cuDNN, every scored
BatchNorm(+ReLU) layer by the library's own fused sm_100a kernels, whose backward reads (x, dy) once for the class-keyed sums
S[k, c] += dz * xhat (the label-keyed segmented reduction K1), dgamma (= their row sum: the reference's bn.weight.grad) and
dbeta, and once more for dx -- then the end-of-step fold, the per-step all-reduce of the BN-gamma gradient (N > 1) and the K2
EIC update.  Workload at N = 1: BASELINE.json configs[1] (DeepLabV3-ResNet101, 19 classes, 512x1024).

Printed JSON line (rank 0):
  value        images/s over all ranks, inputs already resident in HBM (CUDA events, max over ranks), no per-kernel timers
  roofline     the dominant kernel of the path (the class-keyed reduction inside the fused BN backward): algorithmic bytes of
               its launches / their CUDA-event durations, measured in a separate timed region of the same K steps whose CUDA
               graph carries event-record nodes around every launch of the path (on the launching stream); `path_phases` has the other passes; `by_label_fragmentation` the same
               kernel on coarser / finer label maps
  e2e          the same metric through the public HOST API dcfp_b200.scorer.score_calibration_set: per step the micro-batch is
               copied host->device from pinned memory and the loss is read back; scorer set-up and the final score read-back
               are inside the timed region
  cpu_baseline the reference's CPU path on this box's host cores (bounded sample)
  extras       (N = 1) scores_only, bf16_autocast, unfused (round-1 path: cuDNN BN + hook-fed deferred K1), producer_only (plain torch
               fwd+bwd: what the path costs on top), conv_fp32, other_configs (c3, c4), sweep (calibration-set sizes)
`--impl reference` times the reference's own CPU implementation: the UNMODIFIED reference modules from baseline/_ref
(scripts/vendor_reference.py; kind "reference") when present, else the oracle port (kind "port").
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
PROTOCOL = ("per micro-batch (fixed by global image index): zero_grad -> loss(x, y, deepsup=True) [CE + 0.4 CE(deepsup)] -> backward -> "
            "EIC step on this step's BN-gamma gradient (averaged over ranks when N > 1); BN in train mode, no optimizer step, r = 0.999")
NAMES = {"c1": "DeepLabV3-ResNet50", "c2": "DeepLabV3-ResNet101", "c3": "PSPNet-ResNet101", "c4": "DeepLabV3+-ResNet101"}
BASELINE_CFG = {"c1": "configs[0]", "c2": "configs[1]", "c3": "configs[2]", "c4": "configs[3]"}


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=10)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="b200", choices=["b200", "reference"])
    p.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4"])
    p.add_argument("--micro-batch", type=int, default=2)
    p.add_argument("--labels", default="street", choices=["coarse", "street", "fine"],
                   help="fragmentation of the synthetic label maps (dcfp_b200/workloads/synthetic.py): coarse = 48 blobs per image "
                        "(round 1's workload), street = 400 blobs + thin structures (Cityscapes-like; default), fine = 3000 blobs")
    p.add_argument("--flush-mb", type=int, default=16384,
                   help="(unfused path) pending feature-map MiB that trigger a grouped K1 launch before the end of the backward pass")
    p.add_argument("--prime", type=int, default=3,
                   help="set-up steps before the W warm-up steps (cuDNN autotuning, caching-allocator growth); untimed")
    p.add_argument("--conv-precision", default="tf32", choices=["fp32", "tf32"],
                   help="cuDNN convolution math of the feature-map PRODUCER (not part of the path): tf32 = torch's default "
                        "(torch.backends.cudnn.allow_tf32=True), which is what the reference's train.py runs on any Ampere+ GPU "
                        "since it never touches torch.backends; fp32 = IEEE fp32 convolutions")
    p.add_argument("--layout", default="channels_last", choices=["channels_last", "nchw"],
                   help="memory format of the model / feature maps: channels_last is cuDNN's native tensor-core layout and what "
                        "the fused BN kernels take; nchw scores through torch's BN + the hook-fed K1 (TMA-tile NCHW path)")
    p.add_argument("--scores-only", action="store_true",
                   help="freeze every non-BN parameter during scoring: no weight-gradient convolutions (the scores do not "
                        "need them); default off = the reference's full backward")
    p.add_argument("--no-fused", action="store_true",
                   help="score with torch/cuDNN BatchNorm + ReLU and the hook-fed, deferred K1 launches (round 1's path) instead of "
                        "the fused BN(+ReLU) kernels whose backward yields the class-keyed sums (SURVEY 8 f1)")
    p.add_argument("--no-forward-functor", action="store_true",
                   help="skip the extra forward-only region (north_star-literal statistics: K1 with v = BN output)")
    p.add_argument("--no-extras", action="store_true", help="skip the extra keys (scores_only, unfused, producer_only, conv_fp32, "
                                                            "other_configs, label fragmentation, sweep)")
    p.add_argument("--no-graph", action="store_true", help="launch every step eagerly instead of replaying it from a CUDA graph "
                                                           "(for ncu captures of single launches)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--sweep", action="store_true",
                   help="BASELINE.json configs[4]: calibration sets of 100..5000 synthetic images through score_calibration_set "
                        "(set-up, per-step H2D, final D2H inside the timing), sharded over the N ranks")
    p.add_argument("--sweep-sizes", default="100,200,500,1000,2000,5000")
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
    c["name"] = cfg_name
    c["label"] = "%s random-init, %d classes, %dx%d synthetic images (BASELINE.json %s)" % (
        NAMES[cfg_name], c["num_classes"], c["height"], c["width"], BASELINE_CFG[cfg_name])
    return c


def shared_config(c, args):
    """`config` of the JSON line: the SAME dict from both arms (what is measured, not how)."""
    return {"workload": c["label"], "micro_batch_per_gpu": args.micro_batch, "labels": "synthetic, fragmentation '%s'" % args.labels,
            "calibration_images_nominal": 500, "protocol": PROTOCOL}


def make_batches(c, indices_per_step, pin, labels="street"):
    from dcfp_b200.workloads.synthetic import synthetic_batch
    out = []
    for idx in indices_per_step:
        x, y = synthetic_batch(idx, c["num_classes"], c["height"], c["width"], fragmentation=labels)
        out.append((x.pin_memory(), y.pin_memory()) if pin else (x, y))
    return out


# ------------------------------------------------------------------------------------------ CPU reference arm
def _reference_model(c):
    """(step_fn, kind): the UNMODIFIED reference modules (networks.<arch>.Seg_Model + loss.criterion.CriterionDSN +
    pruners.dcfp_pruning, train.py:192-199,215-216) when a reference tree is present (baseline/_ref on the GPU box), else the
    oracle port (bit-identical nets of dcfp_b200/workloads/segnets.py + oracle/eic_ref.py)."""
    import types

    import torch

    from oracle import ref_compat
    if ref_compat.available():
        from dcfp_b200.workloads.segnets import BACKBONE_PARA
        ref = ref_compat.load_reference()
        torch.manual_seed(0)
        crit = ref.crit.CriterionDSN(dataset=types.SimpleNamespace(ignore_label=255))
        net = getattr(ref.networks, c["arch"]).Seg_Model(backbone=c["backbone"], backbone_para=dict(BACKBONE_PARA), model_para={},
                                                         num_classes=c["num_classes"], align_corner=True, criterion=crit, deepsup=True)
        net.train()
        tp = ref.dp.dcfp_pruning(net, 0.999)

        def one(x, y, i):
            torch.manual_seed(i)
            net.zero_grad()
            loss = net(x, y.long(), deepsup=True)  # train.py:259
            loss["loss"].backward()                # train.py:265
            tp.step(net)                           # train.py:267-268 -> pruners/dcfp_pruner.py:15-20
        return one, "reference"
    from dcfp_b200.workloads.segnets import build_segnet
    from oracle import eic_ref, scoring_ref
    model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0)
    model.train()
    layers = scoring_ref.scored_bn_layers(model)
    eic = {n: 0 for n, _ in layers}

    def one(x, y, i):
        torch.manual_seed(i)
        grads, _ = scoring_ref.gamma_grads(model, x, y)
        for n, m in layers:
            eic[n] = eic_ref.eic_step(eic[n], grads[n].numpy(), m.weight.detach().numpy(), 0.999)
    return one, "port"


def cpu_reference(c, micro_batch, steps, warmup, budget_s, labels="street"):
    """The reference's CPU path (train.py:255-268 around torch CPU autograd + pruners/dcfp_pruner.py:15-20), all host threads.
    A step = one micro-batch.  Returns (images/s, steps done, s/step, threads, warm-up steps done, kind)."""
    import torch
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    one, kind = _reference_model(c)
    batches = make_batches(c, [list(range(i * micro_batch, (i + 1) * micro_batch)) for i in range(2)], pin=False, labels=labels)
    t_begin, per, done_w = time.time(), None, 0
    for i in range(warmup):  # warm-up may use at most half of the wall-clock budget
        if per is not None and (time.time() - t_begin) + per > 0.5 * budget_s:
            break
        t = time.time()
        one(*batches[i % len(batches)], i)
        per = time.time() - t
        done_w += 1
    done, t0 = 0, time.time()
    while done < steps:  # the timed steps stop early (and say so) rather than overrun the budget
        if per is not None and done >= 1 and (time.time() - t_begin) + per > budget_s:
            break
        t = time.time()
        one(*batches[(done_w + done) % len(batches)], done_w + done)
        per = time.time() - t
        done += 1
    dt = time.time() - t0
    return done * micro_batch / dt, done, dt / done, cores, done_w, kind


def run_reference_arm(args, c):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, done, s_per, cores, done_w, kind = cpu_reference(c, args.micro_batch, args.steps, args.warmup, args.cpu_budget_s, args.labels)
    sample = "%d timed micro-batches of %d images (%d warm-up) of the %s workload, fp32, %d host threads, %s" % (
        done, args.micro_batch, done_w, args.config, cores,
        "the unmodified reference modules (baseline/_ref)" if kind == "reference" else "oracle port (no reference tree found)")
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": done,
            "warmup": done_w, "ms_per_step": s_per * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": shared_config(c, args),
            "arm": {"requested_steps": args.steps, "requested_warmup": args.warmup, "device": "host CPU, torch %d threads" % cores},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------ B200 arm
class Ctx:
    """Distributed plumbing + timing helpers shared by the measured regions."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        if args.gpus != self.world and self.world > 1:
            raise SystemExit("--gpus %d but WORLD_SIZE=%d" % (args.gpus, self.world))
        if args.gpus > 1 and self.world == 1:
            raise SystemExit("--gpus %d needs torchrun: python -m torch.distributed.run --nproc-per-node %d bench.py ..." % (args.gpus, args.gpus))
        torch.cuda.set_device(self.local_rank)
        self.dev = torch.device("cuda", self.local_rank)
        if self.world > 1:
            os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
            dist.init_process_group("nccl", device_id=self.dev)
        self.nhwc = args.layout == "channels_last"

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms):
        if self.world == 1:
            return ms
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def set_conv_math(self, tf32):
        self.torch.backends.cudnn.allow_tf32 = tf32
        self.torch.backends.cuda.matmul.allow_tf32 = tf32

    def build(self, c):
        from dcfp_b200.workloads.segnets import build_segnet
        model = build_segnet(c["arch"], c["backbone"], c["num_classes"], seed=0).to(self.dev)
        return model.to(memory_format=self.torch.channels_last) if self.nhwc else model

    def resident(self, host):
        t = self.torch
        return [(x.to(self.dev).contiguous(memory_format=t.channels_last) if self.nhwc else x.to(self.dev), y.to(self.dev)) for x, y in host]

    def timed(self, step_fn, first, n, nvtx=None, finish=None):
        """n steps bracketed by barrier + synchronize on both sides, CUDA events on the current stream, max over ranks.
        finish: called before the end event is recorded (joins side-stream work into the timed region)."""
        t = self.torch
        self.barrier()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        marks = [e0]
        rng = t.cuda.nvtx.range_start(nvtx) if nvtx else None
        e0.record()
        for s in range(first, first + n):
            step_fn(s)
            marks.append(t.cuda.Event(enable_timing=True))
            marks[-1].record()
        if finish is not None:
            finish()
        e1.record()
        if rng is not None:
            t.cuda.nvtx.range_end(rng)
        self.barrier()
        return self.max_over_ranks(e0.elapsed_time(e1)), e0.elapsed_time(e1), [round(a.elapsed_time(b), 3) for a, b in zip(marks[:-1], marks[1:])]


def phase_table(sc, K, region_ms):
    out = {k: {"ms_per_step": ms / K, "algorithmic_gb_per_step": nb / K / 1e9, "launches_per_step": n / K,
               "achieved_gbs": (nb / (ms * 1e-3) / 1e9) if ms > 0 and nb else None, "share_of_step": ms / region_ms}
           for k, (ms, nb, n) in sc.phase_times().items()}
    k1_ms, k1_bytes, k1_n = sc.k1_time_ms()
    if k1_n:
        out["k1_deferred"] = {"ms_per_step": k1_ms / K, "algorithmic_gb_per_step": k1_bytes / K / 1e9, "launches_per_step": k1_n / K,
                              "achieved_gbs": k1_bytes / (k1_ms * 1e-3) / 1e9 if k1_ms > 0 else None, "share_of_step": k1_ms / region_ms}
    return out


def dominant(sc):
    """(ms, algorithmic bytes, launches, fused?) of the path's dominant kernel: B1 of the fused BN backward, else deferred K1."""
    ph = sc.phase_times()
    if "bn_bwd_reduce" in ph:
        return ph["bn_bwd_reduce"] + (True,)
    return sc.k1_time_ms() + (False,)


def event_pair_ms(sc):
    """Mean elapsed time of an EMPTY event bracket recorded right in front of every bracketed B1 launch (scorer.py): the part
    of a bracketed kernel time that is the timing events' own node transition, not the kernel."""
    ms, _, n = sc.phase_times().get("event_pair", (0.0, 0, 0))
    return ms / n if n else 0.0


def measure_run(ctx, model, c, resident, K, W, prime, gidx, timing_region=True, graph=True, **run_kw):
    """[timing run: prime + W + K steps launched eagerly with CUDA events around every launch of the path]  then the measured
    run: prime + W warm-up + K timed steps, no per-kernel timers, the step replayed from a CUDA graph (graph=True: the first
    two steps launch eagerly, the third is captured).  Returns a dict."""
    from dcfp_b200 import ops
    from dcfp_b200.scorer import CalibrationRun
    t = ctx.torch
    n = len(resident)
    res = {}
    kw = dict(r=0.999, flush_bytes=ctx.args.flush_mb << 20, keep_totals=True, seed=0, **run_kw)
    if timing_region:
        # the same step captured WITH event-record nodes around every launch of the path: in-graph kernel durations,
        # free of the host's launch rate (eager launches from Python left gaps that the event pairs counted as kernel time)
        run = CalibrationRun(model, c["num_classes"], timing=True, graph=graph, **kw)
        sc = run.scorer

        def tstep(s):
            run.step(*resident[s % n], mb_index=gidx(s))
        for s in range(max(prime, 3)):
            tstep(s)
        if prime:
            # cuDNN's autotuning trials leave >100 GB of workspace blocks cached: release them once (steady state needs ~20 GB)
            t.cuda.synchronize()
            t.cuda.empty_cache()
        for s in range(W):
            tstep(s)
        t.cuda.synchronize()
        sc.reset_timing(drop_events=run._graph is None)
        ms_t, ms_local, _ = ctx.timed(tstep, W, K, finish=run.sync_scores)
        t.cuda.synchronize()
        d_ms, d_bytes, d_n, fused_on = dominant(sc)
        pair = event_pair_ms(sc) if fused_on else 0.0
        res.update(timing_ms=ms_t, phases=phase_table(sc, K, max(ms_local, 1e-9)), dom_ms_raw=d_ms, dom_ms=d_ms - pair * d_n,
                   event_pair_us=pair * 1e3, dom_bytes=d_bytes, dom_launches=d_n, fused_on=fused_on, share=d_ms / max(ms_local, 1e-9))
        run.close()
        prime = 0 if prime == 0 else max(prime, 3)
    run = CalibrationRun(model, c["num_classes"], timing=False, graph=graph, **kw)
    sc = run.scorer

    def step(s):
        run.step(*resident[s % n], mb_index=gidx(s))
    for s in range(max(prime, 3 if graph else 2)):
        step(s)
    t.cuda.reset_peak_memory_stats(ctx.dev)
    sc.steps = 0  # the EIC state must not see the priming steps
    sc.eic.zero_()
    if sc.total_arena is not None:
        sc.total_arena.zero_()
    for s in range(W):
        step(s)
    launches0 = ops.launch_count()
    replays0 = run.graph_replays
    mem0 = t.cuda.memory_stats(ctx.dev)
    t_wall0 = time.time()
    ms, _, step_ms = ctx.timed(step, W, K, nvtx="timed", finish=run.sync_scores)
    t_wall1 = time.time()
    # kernels of this library inside the timed region: launched eagerly + (graph replays x kernels captured per replay)
    launches = ops.launch_count() - launches0 + (run.graph_replays - replays0) * run.graph_launches
    mem1 = t.cuda.memory_stats(ctx.dev)
    res.update({"ms": ms, "step_ms": step_ms, "launches": int(launches), "graph_replays": run.graph_replays - replays0,
                "wall": (t_wall0, t_wall1), "run": run, "scorer": sc,
                "allocator": {"cudaMalloc_calls_in_timed_region": mem1.get("num_device_alloc", 0) - mem0.get("num_device_alloc", 0),
                              "cudaFree_calls_in_timed_region": mem1.get("num_device_free", 0) - mem0.get("num_device_free", 0),
                              "alloc_retries_in_timed_region": mem1.get("num_alloc_retries", 0) - mem0.get("num_alloc_retries", 0),
                              "reserved_gb": mem1.get("reserved_bytes.all.current", 0) / 1e9,
                              "allocated_peak_gb": mem1.get("allocated_bytes.all.peak", 0) / 1e9}})
    return res


def roofline_of(res, K, mb, nhwc):
    peak, peak_src = hbm_peak()
    achieved = res["dom_bytes"] / (res["dom_ms"] * 1e-3) / 1e9 if res["dom_ms"] > 0 else 0.0
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            traffic = json.load(f)
    except Exception:
        pass
    kernel = ("dcfp::class_stats_nhwc_kernel<float, BWD, FUSED> -- B1 of the fused BN backward: label-keyed segmented reduction of "
              "v = dz * xhat (ReLU gate recomputed) + sum dz, sum v; one launch per BN layer") if res["fused_on"] else \
        "dcfp::%s<float, BWD> (K1, label-keyed segmented reduction, v = dy * xhat; deferred grouped launches)" % (
            "class_stats_nhwc_kernel" if nhwc else "class_stats_kernel")
    tr = (traffic or {}).get("fused" if res["fused_on"] else "deferred")
    return {"kernel": kernel, "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "peak_source": peak_src,
            # DRAM bytes per launch from the committed ncu --set full capture: where the capture covers a SUBSET of the step's
            # launches (the per-layer launches of the fused path) its measured traffic / algorithmic ratio is applied to this
            # run's average algorithmic bytes per launch, so that the two figures refer to the same launches
            "traffic": (tr["ratio"] * res["dom_bytes"] / max(res["dom_launches"], 1) if "ratio" in tr else tr["dram_bytes_per_launch"]) if tr else None,
            "traffic_note": tr.get("note") if tr else "no ncu --set full capture recorded for this path yet",
            "launches": res["dom_launches"], "algorithmic_bytes_per_launch": res["dom_bytes"] / max(res["dom_launches"], 1),
            "avg_launch_ms": res["dom_ms"] / max(res["dom_launches"], 1), "share_of_step": res["share"],
            # every bracketed launch is [event record | kernel | event record]; an EMPTY bracket recorded in front of each one
            # measures what the events themselves add (one graph-node transition), and `achieved` / `avg_launch_ms` are net of
            # it -- kernel + ONE launch transition, what the launch costs the un-instrumented step.  *_raw: nothing subtracted
            "event_pair_overhead_us": res.get("event_pair_us", 0.0),
            "avg_launch_ms_raw": res.get("dom_ms_raw", res["dom_ms"]) / max(res["dom_launches"], 1),
            "achieved_raw": res["dom_bytes"] / (res.get("dom_ms_raw", res["dom_ms"]) * 1e-3) / 1e9 if res["dom_ms"] > 0 else 0.0,
            "frac_raw": (res["dom_bytes"] / (res.get("dom_ms_raw", res["dom_ms"]) * 1e-3) / 1e9 / peak) if res["dom_ms"] > 0 else 0.0,
            "algorithmic_bytes_per_image": res["dom_bytes"] / (K * mb),
            "images_per_s_of_kernel_time": K * mb / (res["dom_ms"] * 1e-3) if res["dom_ms"] > 0 else None,
            "measured": "a separate timed region of the same %d steps, the step captured into a CUDA graph WITH event-record nodes "
                        "around every launch of the path (in-graph kernel durations, read after each synchronised replay); the mean "
                        "of the empty brackets recorded beside them is subtracted once per launch (see event_pair_overhead_us)" % K}


def run_sweep(ctx, model, c, sizes, fused=True):
    """Calibration sets of n images through score_calibration_set (host loader -> pinned H2D per step -> scores on the host)."""
    from dcfp_b200.scorer import score_calibration_set
    t, args = ctx.torch, ctx.args
    mb = args.micro_batch
    pool = make_batches(c, [list(range(i * mb, (i + 1) * mb)) for i in range(8)], pin=True, labels=args.labels)

    def fetch(lo, hi):  # a loader: global image range -> pinned host tensors (8 distinct micro-batches, cycled)
        return pool[(lo // mb) % len(pool)]
    out = []
    score_calibration_set(model, fetch, None, c["num_classes"], micro_batch=mb, n_images=4 * mb * ctx.world, fused=fused)
    for n in sizes:
        ctx.barrier()
        e0, e1 = t.cuda.Event(enable_timing=True), t.cuda.Event(enable_timing=True)
        e0.record()
        res = score_calibration_set(model, fetch, None, c["num_classes"], micro_batch=mb, n_images=n, fused=fused)
        e1.record()
        ctx.barrier()
        ms = ctx.max_over_ranks(e0.elapsed_time(e1))
        used = res["_stats"]["steps"] * mb * ctx.world
        out.append({"images": n, "images_scored": used, "seconds": ms * 1e-3, "images_per_s": used / (ms * 1e-3),
                    "h2d_bytes": res["_stats"]["h2d_bytes"], "d2h_bytes": res["_stats"]["d2h_bytes"]})
    return out


def run_b200_arm(args, c):
    import torch

    from dcfp_b200 import ops
    from dcfp_b200.scorer import CalibrationRun, score_calibration_set
    from dcfp_b200.workloads.synthetic import label_run_stats
    ops.require_gpu()  # no CPU fallback: fail loudly
    ctx = Ctx(args)
    world, rank = ctx.world, ctx.rank
    tf32 = args.conv_precision == "tf32"
    ctx.set_conv_math(tf32)
    torch.backends.cudnn.benchmark = os.environ.get("DCFP_BENCH_CUDNN_BENCHMARK", "1") == "1"
    K, W, mb = args.steps, args.warmup, args.micro_batch
    model = ctx.build(c)

    def gidx(s):  # global micro-batch index of (step s, this rank): scorer.shard_plan's deal
        return s * world + rank

    if args.sweep:
        sizes = [int(v) for v in args.sweep_sizes.split(",")]
        sw = run_sweep(ctx, model, c, sizes, fused=not args.no_fused)
        if rank == 0:
            nominal = min(sw, key=lambda r: abs(r["images"] - 500))
            steps = max(nominal["images_scored"] // (mb * world), 1)
            line = {"metric": METRIC, "value": nominal["images_per_s"], "unit": UNIT, "n_gpus": world, "steps": steps,
                    "warmup": 4, "ms_per_step": nominal["seconds"] * 1e3 / steps,
                    "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                    "config": dict(shared_config(c, args), sweep="BASELINE.json configs[4]: whole calibration sets through "
                                   "score_calibration_set, set-up + per-step H2D + final D2H inside the timing; value = the 500-image set"),
                    "sweep": sw}
            print(json.dumps(line), flush=True)
        if world > 1:
            ctx.dist.destroy_process_group()
        return

    host = make_batches(c, [list(range(gidx(s) * mb, (gidx(s) + 1) * mb)) for s in range(W + K)], pin=True, labels=args.labels)
    resident = ctx.resident(host)
    torch.cuda.synchronize()
    run_len, n_cls = label_run_stats(torch.stack([y for _, y in host[:4]]).flatten(0, 1))

    # ---- phase A: inputs resident in HBM ---------------------------------------------------------------------------
    # nvidia-smi attaches to the driver when it starts (stalls launches for ~0.2 s): start it before the set-up steps
    sampler = ClockSampler(ctx.local_rank).start() if rank == 0 else None
    main = measure_run(ctx, model, c, resident, K, W, args.prime, gidx, scores_only=args.scores_only, fused=not args.no_fused,
                       graph=not args.no_graph)
    clocks = sampler.stop(*main["wall"]) if sampler else None
    sc = main["scorer"]
    main["run"].sync_scores()
    value = K * mb * world / (main["ms"] * 1e-3)
    roofline = roofline_of(main, K, mb, ctx.nhwc)
    # the path as a whole: algorithmic bytes of ALL its passes (F1 + F2 + B1 + B2 [+ deferred K1]) over the sum of their bracketed
    # times -- raw brackets, nothing subtracted.  Second reads of a map that the 126 MB L2 still holds count as algorithmic bytes.
    ph = {k: v for k, v in main["phases"].items() if k != "event_pair"}
    p_gb, p_ms = sum(v["algorithmic_gb_per_step"] for v in ph.values()), sum(v["ms_per_step"] for v in ph.values())
    if p_ms > 0:
        roofline["whole_path"] = {"algorithmic_gb_per_step": p_gb, "ms_per_step": p_ms, "achieved_gbs": p_gb / (p_ms * 1e-3),
                                  "frac": p_gb / (p_ms * 1e-3) / roofline["peak"], "passes": sorted(ph)}
    sc.all_reduce_totals()  # the single end-of-pass statistics all-reduce (outside the per-step timing, reported below)
    allreduce_ms = 0.0
    if world > 1:
        torch.cuda.synchronize()
        ea, eb = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ea.record()
        sc.all_reduce_totals()
        eb.record()
        torch.cuda.synchronize()
        allreduce_ms = ctx.max_over_ranks(ea.elapsed_time(eb))
    arena_bytes = sc.total_arena.numel() * 8
    sum_c = sc.total_channels
    fused_layers, fused_tails = sc.fused_layer_calls, sc.fused_tail_calls
    main["run"].close()
    extras = {}

    def short(label, Ks=5, Ws=2, batches=None, timing_region=False, graph=False, **kw):
        r = measure_run(ctx, model, c, batches or resident, Ks, Ws, 0, gidx, timing_region=timing_region, graph=graph, **kw)
        out = {"value": Ks * mb * world / (r["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms"] / Ks, "steps": Ks, "warmup": Ws}
        if timing_region:
            rf = roofline_of(r, Ks, mb, ctx.nhwc)
            out.update(roofline_frac=rf["frac"], roofline_achieved_gbs=rf["achieved"], dominant_kernel_share_of_step=rf["share_of_step"],
                       path_phases=r["phases"])
        r["run"].close()
        extras[label] = out
        return out

    do_extras = not args.no_extras and not args.no_fused and not args.scores_only and ctx.nhwc
    if do_extras and world == 1:
        # (1) the same kernel on coarser / finer label maps (slot-cache evictions, short class runs)
        frag = {}
        for lab in ("coarse", "street", "fine"):
            hb = make_batches(c, [list(range(i * mb, (i + 1) * mb)) for i in range(2)], pin=False, labels=lab)
            rl, nc = label_run_stats(torch.stack([y for _, y in hb]).flatten(0, 1))
            o = short("_frag", Ks=4, Ws=2, batches=ctx.resident(hb), timing_region=True, graph=True)
            frag[lab] = {"roofline_frac": o["roofline_frac"], "achieved_gbs": o["roofline_achieved_gbs"], "value": o["value"],
                         "mean_class_run_px_at_stride8": rl, "classes_per_image": nc}
        extras.pop("_frag", None)
        roofline["by_label_fragmentation"] = frag
        # (2) what the path costs on top of the bare producer, and the alternatives
        short("scores_only", scores_only=True, graph=True)
        extras["scores_only"]["what"] = "non-BN parameters frozen: no weight-gradient convolutions; same scores (tests/test_gpu_scorer.py)"
        short("no_residual_fusion", graph=True, fuse_residual=False)
        extras["no_residual_fusion"]["what"] = ("fused BN(+ReLU) kernels, but bn3 -> (+ shortcut) -> ReLU of every bottleneck left as BN kernel + "
                                                "torch add + torch ReLU (the round-2 path before the fused tail)")
        short("unfused", fused=False, timing_region=True, graph=True)
        extras["unfused"]["what"] = "round-1 path: torch/cuDNN BatchNorm + ReLU, hook-fed K1 deferred into grouped launches"

        def bare(tag, steps=5, warm=2):
            model.train()

            def st(s):
                x, y = resident[s % len(resident)]
                model.zero_grad(set_to_none=True)
                out = model(x, y.long(), deepsup=True)
                (out["loss"] if isinstance(out, dict) else out).backward()
            for s in range(warm):
                st(s)
            ms, _, _ = ctx.timed(st, warm, steps)
            model.zero_grad(set_to_none=True)
            extras[tag] = {"value": steps * mb / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms / steps, "steps": steps, "warmup": warm}
        bare("producer_only")
        extras["producer_only"]["what"] = "plain torch fwd + bwd of the same net (cuDNN BN, no scorer): the feature-map producer alone"
        extras["producer_only"]["path_overhead_ms_per_step"] = main["ms"] / K - extras["producer_only"]["ms_per_step"]
        short("bf16_autocast", graph=True, autocast_dtype=torch.bfloat16)
        extras["bf16_autocast"]["what"] = ("forward under torch.autocast(bfloat16): bf16 convolutions and bf16 feature maps through the fused BN "
                                           "kernels (fp32 statistics / sums); NOT the reference's fp32 arithmetic -- scores differ at bf16 level")
        ctx.set_conv_math(False)
        short("conv_fp32", Ks=3, Ws=2)
        extras["conv_fp32"]["what"] = "IEEE fp32 convolutions (cudnn.allow_tf32=False), channels_last, fused path"
        ctx.set_conv_math(tf32)
    if do_extras:
        # (3) the other BASELINE configs (c3 is the one BASELINE.json shards across 2/4/8 GPUs: measured at every N)
        del resident
        torch.cuda.empty_cache()
        others = {}
        for name in (["c3", "c4"] if world == 1 else ["c3"]):
            if name == c["name"]:
                continue
            oc = workload(name)
            om = ctx.build(oc)
            ob = ctx.resident(make_batches(oc, [list(range(gidx(s) * mb, (gidx(s) + 1) * mb)) for s in range(4)], pin=False, labels=args.labels))
            r = measure_run(ctx, om, oc, ob, 5, 2, 2, gidx, timing_region=True, graph=True)
            rf = roofline_of(r, 5, mb, ctx.nhwc)
            others[name] = {"workload": oc["label"], "value": 5 * mb * world / (r["ms"] * 1e-3), "unit": UNIT, "ms_per_step": r["ms"] / 5,
                            "steps": 5, "warmup": 2, "roofline_frac": rf["frac"], "roofline_achieved_gbs": rf["achieved"],
                            "dominant_kernel_share_of_step": rf["share_of_step"]}
            r["run"].close()
            del om, ob, r
            torch.cuda.empty_cache()
        extras["other_configs"] = others
        resident = ctx.resident(host)

    # ---- forward functor (north_star-literal): forward pass only, K1 with v = BN output, deferred into grouped launches
    forward = None
    if not args.no_forward_functor:
        run_f = CalibrationRun(model, c["num_classes"], mode="fwd", flush_bytes=args.flush_mb << 20, keep_totals=True, timing=True, seed=0)

        def fstep(s):
            run_f.step(*resident[s % len(resident)], mb_index=gidx(s))
        for s in range(max(W, 1)):
            fstep(s)
        ctx.barrier()
        run_f.scorer.k1_events.clear()
        ms_f, _, _ = ctx.timed(fstep, W, K)
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
        kw = dict(micro_batch=mb, flush_bytes=args.flush_mb << 20, scores_only=args.scores_only, fused=not args.no_fused,
                  graph=not args.no_graph)
        if W:
            xw, yw = global_order(0, W)
            score_calibration_set(model, xw, yw, c["num_classes"], **kw)
        xk, yk = global_order(W, W + K)
        ctx.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = score_calibration_set(model, xk, yk, c["num_classes"], **kw)
        e1.record()
        ctx.barrier()
        ms_b = ctx.max_over_ranks(e0.elapsed_time(e1))
        st = out["_stats"]
        assert st["steps"] == K
        e2e = {"value": K * mb * world / (ms_b * 1e-3), "unit": UNIT, "h2d_bytes_per_step": st["h2d_bytes"] // K,
               "d2h_bytes_per_step": st["d2h_bytes"] / K, "ms_per_step": ms_b / K,
               "api": "dcfp_b200.scorer.score_calibration_set(model, host_images, host_labels, K)"}
        if do_extras:  # a short calibration-set sweep (the full one: bench.py --sweep)
            extras["sweep"] = run_sweep(ctx, model, c, [100, 500])

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, done, s_per, cores, done_w, kind = cpu_reference(c, mb, 3, 1, 60.0, args.labels)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": kind,
               "sample": "%d timed micro-batches of %d images after %d warm-up (%.1f s each) of the same workload, fp32, %s" % (
                   done, mb, done_w, s_per, "unmodified reference modules from baseline/_ref" if kind == "reference" else "oracle/scoring_ref.py")}

    if rank == 0:
        fused_on = main["fused_on"]
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": main["ms"] / K,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": shared_config(c, args),
                "arm": {"images_timed": K * mb * world, "sum_scored_channels": sum_c, "fused_bn_layer_calls_total": fused_layers, "fused_bn_residual_tail_calls_total": fused_tails,
                        "labels_mean_class_run_px_at_stride8": run_len, "labels_classes_per_image": n_cls,
                        "conv_math": ("tf32 (torch default cudnn.allow_tf32=True, as the reference's train.py runs its convolutions)"
                                      if tf32 else "fp32 (cudnn.allow_tf32=False)") + "; the path's own arithmetic is fp32 (fp64 across CTAs)",
                        "bn": ("fused: dcfp BN(+ReLU) forward (statistics + normalise launches; bn3 + shortcut + ReLU of a bottleneck in "
                               "the same normalise pass) / backward kernels, class-keyed sums inside the BN backward") if fused_on else "torch/cuDNN BatchNorm + ReLU, hook-fed deferred K1",
                        "l2": "per-step feature maps (%.1f GB read by the dominant kernel) exceed the 126 MB L2; no explicit flush" % (
                            roofline["algorithmic_bytes_per_image"] * mb / 1e9),
                        "layout": args.layout, "backward": "scores_only (no weight-gradient convolutions)" if args.scores_only else
                        "full (all gradients, as the reference's training step)", "priming_steps": args.prime,
                        "parallelism": "dp%d (micro-batches dealt round-robin)" % world},
                "step_ms": main["step_ms"], "allocator": main["allocator"], "roofline": roofline, "path_phases": main["phases"],
                "timed_with_per_launch_events": {"ms_per_step": main["timing_ms"] / K, "value": K * mb * world / (main["timing_ms"] * 1e-3)},
                "forward_functor": forward, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": main["launches"],
                "cuda_graph": {"replays_in_timed_region": main["graph_replays"], "dcfp_kernels_per_replay": main["run"].graph_launches},
                "clocks": clocks,
                "stats_allreduce": {"bytes": arena_bytes, "ms": allreduce_ms, "what": "one all-reduce of the [2,K+1,sumC] fp64 totals (K classes "
                                    "+ the pixels outside [0,K)) + counts at the end of the pass"}}
        line.update(extras)
        print(json.dumps(line), flush=True)
    if world > 1:
        ctx.dist.destroy_process_group()


def main():
    args = parse_args()
    c = workload(args.config)
    if args.impl == "reference":
        run_reference_arm(args, c)
    else:
        run_b200_arm(args, c)


if __name__ == "__main__":
    main()
