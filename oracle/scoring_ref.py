"""CPU restatement of the reference's score-accumulation loop -- TEST INFRASTRUCTURE ONLY.

The reference produces `score.pth` inside its training loop (train.py:255-270, run on CUDA there);
`train.py` itself cannot be executed without datasets on disk and unconditional `.cuda()` calls
(train.py:246-253), so the six lines that matter are restated here around torch's CPU autograd:

    optimizer.zero_grad()                              train.py:255
    loss = model(images, labels, deepsup=True)         train.py:259
    loss['loss'].backward()                            train.py:265
    train_pruning.step(seg_model)                      train.py:267-268  (pruners/dcfp_pruner.py:15-20)
    [optimizer.step()]                                 train.py:270      (left out: weights stay fixed
                                                                           during calibration scoring)

`eic_step` is the numpy restatement in oracle/eic_ref.py (pinned by tests/golden/eic_steps.npz);
the model's forward/backward is torch on the host cores -- exactly what the reference's CPU path
executes.  Used by tests (checker), __graft_entry__.smoke() (checker) and bench.py's `cpu_baseline`
/ `--impl reference` legs (the timed CPU arm).  Nothing in dcfp_b200 imports this.

Multi-rank protocol restated as well (engine.py:66 DDP averages gradients before `step`):
`world` micro-batches form one step, their BN-gamma gradients are averaged before the sign gate.
"""
import numpy as np
import torch
import torch.nn as nn

from . import eic_ref


def scored_bn_layers(model):
    """(name, module) of every BN the reference scores (pruners/dcfp_pruner.py:11-13)."""
    ignore = getattr(model, "ignore_prune_layer", [])
    return [(n, m) for n, m in model.named_modules() if isinstance(m, (nn.BatchNorm2d, nn.SyncBatchNorm)) and n not in ignore]


def gamma_grads(model, images, labels):
    """One forward/backward of the calibration loss on the host; returns ({bn: dgamma fp32}, loss)."""
    model.zero_grad(set_to_none=True)
    out = model(images, labels.long(), deepsup=True)
    loss = out["loss"] if isinstance(out, dict) else out
    loss.backward()
    return {n: m.weight.grad.detach().clone() for n, m in scored_bn_layers(model)}, float(loss.detach())


def score(model, micro_batches, r=0.999, world=1, restore_bn_stats=True, seed=None):
    """EIC after len(micro_batches) // world steps.  micro_batches: sequence of (images, labels) CPU tensors.
    seed: when given, the torch RNG is re-seeded with seed + global micro-batch index before each forward
    (the deep-supervision head holds a Dropout2d, deeplabv3.py:40), making the result independent of the
    number of ranks -- the product's CalibrationRun follows the same rule.

    Returns ({'eic': {bn_name: np.float32[C]}}, [loss per micro-batch])."""
    layers = scored_bn_layers(model)
    saved = None
    if restore_bn_stats:
        saved = [(m, m.running_mean.clone(), m.running_var.clone(), m.num_batches_tracked.clone())
                 for m in model.modules() if isinstance(m, nn.modules.batchnorm._BatchNorm) and m.running_mean is not None]
    was_training = model.training
    model.train()
    eic = {n: 0 for n, _ in layers}
    losses = []
    try:
        n_steps = len(micro_batches) // world
        for step in range(n_steps):
            acc = None
            for rnk in range(world):
                x, y = micro_batches[step * world + rnk]
                if seed is not None:
                    torch.manual_seed(seed + step * world + rnk)
                g, loss = gamma_grads(model, x, y)
                losses.append(loss)
                acc = g if acc is None else {k: acc[k] + g[k] for k in acc}
            if world > 1:
                acc = {k: v / world for k, v in acc.items()}
            for n, m in layers:
                eic[n] = eic_ref.eic_step(eic[n], acc[n].numpy(), m.weight.detach().numpy(), r)
    finally:
        model.train(was_training)
        if saved is not None:
            with torch.no_grad():
                for m, mean, var, nbt in saved:
                    m.running_mean.copy_(mean)
                    m.running_var.copy_(var)
                    m.num_batches_tracked.copy_(nbt)
    return {"eic": {n: np.asarray(v, dtype=np.float32) for n, v in eic.items()}}, losses
