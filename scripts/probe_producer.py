"""How fast is the feature-map PRODUCER (torch/cuDNN fwd+bwd of the c2 net) in NCHW vs channels_last? (development aid)"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from dcfp_b200.workloads.segnets import build_segnet, CONFIGS
from dcfp_b200.workloads.synthetic import synthetic_batch
cfg = CONFIGS[os.environ.get("CFG", "c2")]
torch.backends.cudnn.benchmark = True
dev = "cuda"
x, y = synthetic_batch([0, 1], cfg["num_classes"], cfg["height"], cfg["width"])
x, y = x.to(dev), y.to(dev).long()
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    for cl in (False, True):
        model = build_segnet(cfg["arch"], cfg["backbone"], cfg["num_classes"]).to(dev).train()
        xi = x
        if cl:
            model = model.to(memory_format=torch.channels_last)
            xi = x.contiguous(memory_format=torch.channels_last)
        def step():
            model.zero_grad(set_to_none=True)
            loss = model(xi, y, deepsup=True)["loss"]
            loss.backward()
        for _ in range(3):
            step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t = time.time()
        e0.record()
        for _ in range(5):
            step()
        e1.record()
        torch.cuda.synchronize()
        print("tf32=%s channels_last=%s: %.1f ms/step (gpu) %.1f ms/step (wall) peak mem %.1f GB" %
              (tf32, cl, e0.elapsed_time(e1) / 5, (time.time() - t) / 5 * 1e3, torch.cuda.max_memory_allocated() / 1e9), flush=True)
        del model
        torch.cuda.empty_cache()
        if not tf32:
            break
