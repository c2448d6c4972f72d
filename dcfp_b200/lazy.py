"""Deferred BatchNorm outputs: how the tail of a bottleneck block -- bn3 -> (+ shortcut) -> ReLU
(networks/backbone/resnet.py:49-56) -- becomes ONE pass over the map without touching the block's code.

The shortcut is only known when the block executes `out = out + residual`, after bn3 has returned.  So a BatchNorm that
the scorer has seen feeding such a tail returns a `PendingBN`: a tensor-shaped placeholder (shape / dtype / device of the
result, no arithmetic done yet) that remembers how to run the layer.  What happens next decides the kernel:

  pending + residual        -> still pending, the residual recorded                        (no launch)
  ReLU module / F.relu      -> run(residual, relu=True): y = relu(bn(x) + residual)        (statistics + ONE element-wise pass)
  anything else             -> the placeholder is replaced by its real value -- run(None, False) [+ residual] -- and the
                               operation proceeds on plain tensors: results are those of the unfused program

`__torch_function__` sees every torch-level use of the placeholder, so the deferral is a pure scheduling decision: no
sequence of torch operations can observe a value that differs from the eager one.  (Code that bypasses torch -- reading
`data_ptr()` in an extension -- would see the un-normalised input; nothing in the scored networks does.  A program that
uses BOTH the BN output and the rectified sum makes the layer run twice -- same values, but the train-mode running
statistics advance twice; the scored networks use each once.)

No CUDA dependency here: scorer.py supplies `run`; tests/test_lazy_bn_cpu.py drives the protocol with a torch `run`."""
import torch
from torch.utils._pytree import tree_map

_ADD = {"add", "__add__", "__radd__", "add_", "__iadd__"}
_ADD_INPLACE = {"add_", "__iadd__"}
_RELU = {"relu", "relu_"}
# attribute reads and shape queries answered by the placeholder itself (it has the result's shape, dtype, device, strides)
_META = {"shape", "dtype", "device", "ndim", "layout", "is_cuda", "is_cpu", "is_meta", "is_sparse", "is_quantized", "is_mkldnn",
         "is_nested", "requires_grad", "is_leaf", "names", "size", "dim", "ndimension", "stride", "numel", "nelement", "element_size",
         "is_contiguous", "is_floating_point", "is_complex", "storage_offset", "get_device", "__len__", "is_same_size", "type"}


def _name(func):
    n = getattr(func, "__name__", "")
    if n == "__get__":  # getset descriptor of a tensor attribute (`t.shape`)
        return getattr(getattr(func, "__self__", None), "__name__", "")
    return n


class PendingBN(torch.Tensor):
    """Placeholder for a BatchNorm output that has not been computed yet (see the module docstring).

    run(residual | None, relu: bool) -> the layer's output as a plain tensor, differentiable w.r.t. its input and
    the residual."""

    @staticmethod
    def make(x, run, requires_grad, tag=None, residual=None, parent=None):
        t = torch.Tensor._make_subclass(PendingBN, x.detach(), bool(requires_grad))
        t._x, t._run, t._residual, t._parent, t._alias, t._value, t._dcfp_bn = x, run, residual, parent, None, None, tag
        return t

    # -- resolution ------------------------------------------------------------------------------------------------
    def materialize(self):
        """The eager value: bn(x) [+ residual], computed once."""
        if self._alias is not None:  # `p += residual` happened: this object IS the sum now
            return self._alias.materialize()
        if self._value is None:
            if self._residual is None:
                self._value = self._run(None, False)
            else:  # bn(x) + residual: the BN part is whatever the parent placeholder resolves to
                base = self._parent.materialize() if self._parent is not None else self._run(None, False)
                self._value = base + self._residual
        return self._value

    def relu(self, inplace):
        """ReLU of the pending value.  Deferred so far: ONE fused pass.  Already materialised: a plain ReLU on the value
        (in place when asked, exactly as the module would have done)."""
        if self._alias is not None:
            return self._alias.relu(inplace)
        if self._value is None and self._parent is not None and self._parent._value is not None:
            self.materialize()  # the BN output already exists: add and rectify it like the eager program
        if self._value is not None:
            return torch.relu_(self._value) if inplace else torch.relu(self._value)
        y = self._run(self._residual, True)
        if inplace:  # the tensor the program holds now contains the rectified values
            self._value = y
        return y

    # -- interception ----------------------------------------------------------------------------------------------
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = _name(func)
        if name in _META and not (name == "type" and (len(args) > 1 or kwargs)):  # t.type(dtype) converts VALUES: resolve first
            with torch._C.DisableTorchFunctionSubclass():
                return func(*args, **kwargs)
        if name in _ADD and len(args) == 2 and (not kwargs or (list(kwargs) == ["alpha"] and kwargs["alpha"] == 1)):
            a, b = args
            p, other = (a, b) if isinstance(a, PendingBN) else (b, a)
            if (isinstance(p, PendingBN) and p._residual is None and p._value is None and p._alias is None and type(other) in (torch.Tensor, torch.nn.Parameter)
                    and (p is a or name not in _ADD_INPLACE)
                    and other.shape == p.shape and other.dtype == p.dtype and other.device == p.device
                    and other.dim() == 4 and other.is_contiguous(memory_format=torch.channels_last)):
                with torch._C.DisableTorchFunctionSubclass():
                    req = p.requires_grad or other.requires_grad
                if name in _ADD_INPLACE:
                    p._alias = q = PendingBN.make(p._x, p._run, req, p._dcfp_bn, other)
                    return q
                return PendingBN.make(p._x, p._run, req, p._dcfp_bn, other, p)
        if name in _RELU and len(args) == 1 and isinstance(args[0], PendingBN) and set(kwargs) <= {"inplace"}:
            return args[0].relu(name == "relu_" or bool(kwargs.get("inplace", False)))

        def real(t):
            return t.materialize() if isinstance(t, PendingBN) else t

        args, kwargs = tree_map(real, (tuple(args), dict(kwargs)))
        return func(*args, **kwargs)


# ---- the same tail in the backward pass: the shortcut gradient handed straight to the block that produced the shortcut -------
#
# y_prev --+--> conv1 -> ... -> bn3 --(+)--> ReLU --> y          backward:  dz = relu'(y) * dy  is BOTH bn3's gradient and the
#          +---------- shortcut ------^                                    shortcut's; autograd adds the latter to the gradient
#                                                                          arriving at y_prev through conv1 (read 2, write 1),
# and the previous block then gates that sum with relu'(y_prev) (read 2, write 1).  When y_prev was itself produced by a fused
# tail, the consumer can instead DEPOSIT dz on the producer's backward node and report "no gradient" for the shortcut; the
# producer folds the deposit into its own gate kernel: dz_prev = relu'(y_prev) * (dy + deposit) -- read 3, write 1, no sum tensor.
# autograd runs a node only after every consumer of its outputs has run, so the deposit is always there in time.

def deposit_shortcut_grad(producer_node, grad, ledger):
    """Called by the consumer's backward instead of returning `grad` for its shortcut input.  `ledger`: a one-element list
    counting open deposits, checked at the end of the step (an unconsumed deposit would be a silently lost gradient)."""
    held = getattr(producer_node, "_dcfp_deposit", None)
    producer_node._dcfp_deposit = grad if held is None else held + grad
    if held is None:
        ledger[0] += 1


def take_shortcut_grad(node, ledger):
    """Called at the top of the producer's backward: the deposited gradient, or None."""
    held = getattr(node, "_dcfp_deposit", None)
    if held is not None:
        node._dcfp_deposit = None
        ledger[0] -= 1
    return held
