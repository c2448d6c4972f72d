"""CPU restatement (numpy fp32) of the EIC score update -- TEST INFRASTRUCTURE ONLY.

Follows pruners/dcfp_pruner.py:15-20 of the reference, operation by operation:

    flag     = (grad * gamma > 0)
    grad_tmp = flag * abs(grad) + logical_not(flag) * eic            (:19)
    eic      = eic * r + grad_tmp * (1 - r)                          (:20)

`r` is a Python float in the reference; multiplying an fp32 tensor by it rounds the scalar to
fp32, `(1 - r)` is evaluated in double first (0.0010000000000000009 for r = 0.999).  The state
starts as the Python int 0 (:13), which behaves as an fp32 zero in these expressions.
Pinned against the unmodified reference by tests/golden/eic_steps.npz.
"""
import numpy as np


def eic_step(eic_prev, grad, gamma, r):
    grad = np.asarray(grad, dtype=np.float32)
    gamma = np.asarray(gamma, dtype=np.float32)
    prev = np.zeros_like(grad) if isinstance(eic_prev, int) else np.asarray(eic_prev, dtype=np.float32)
    with np.errstate(invalid="ignore", over="ignore"):
        flag = (grad * gamma) > 0
        grad_tmp = flag.astype(np.float32) * np.abs(grad) + (~flag).astype(np.float32) * prev
        return prev * np.float32(r) + grad_tmp * np.float32(1.0 - r)


def eic_run(grads, gammas, r):
    """grads/gammas: sequences over steps of fp32 vectors -> final eic."""
    eic = 0
    for g, w in zip(grads, gammas):
        eic = eic_step(eic, g, w, r)
    return eic
