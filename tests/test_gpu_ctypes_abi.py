"""The C ABI called the way a non-torch host would (ctypes, raw device pointers, explicit stream) -- the binding shown in
INTEGRATION.md section 3.  torch only provides the device memory here."""
import ctypes

import numpy as np
import pytest
import torch

from oracle import class_stats_ref as ref
from oracle import eic_ref

pytestmark = pytest.mark.gpu


def test_class_stats_and_eic_through_ctypes(native):
    from dcfp_b200 import abi
    lib = abi.load()
    dev = torch.device("cuda")
    N, C, h, w, K = 2, 96, 32, 64, 19
    g = torch.Generator().manual_seed(3)
    x = torch.randn(N, C, h, w, generator=g)
    dy = torch.randn(N, C, h, w, generator=g) * 1e-3
    label = torch.randint(0, K, (N, 4 * h, 4 * w), generator=g).to(torch.uint8)
    label[0, :7] = 255
    mean = x.mean((0, 2, 3))
    invstd = torch.rsqrt(x.var((0, 2, 3), unbiased=False) + 1e-5)
    xd, dyd, ld_, md, sd = x.to(dev), dy.to(dev), label.to(dev), mean.to(dev), invstd.to(dev)
    keys = torch.empty(N, h, w, dtype=torch.uint8, device=dev)
    cnt = torch.zeros(K, dtype=torch.float64, device=dev)
    S1 = torch.zeros(K, C, dtype=torch.float64, device=dev)
    S2 = torch.zeros_like(S1)
    stream = torch.cuda.Stream()
    sp = ctypes.c_void_p(stream.cuda_stream)
    stream.wait_stream(torch.cuda.current_stream())
    rc = lib.dcfp_label_keys(ld_.data_ptr(), abi.LABEL_U8, N, 4 * h, 4 * w, h, w, K, keys.data_ptr(), cnt.data_ptr(), sp)
    assert rc == 0, abi.last_error()
    d = abi.LayerDesc(x=xd.data_ptr(), dy=dyd.data_ptr(), scale=sd.data_ptr(), shift=md.data_ptr(), keys=keys.data_ptr(),
                      S1=S1.data_ptr(), S2=S2.data_ptr(), N=N, C=C, h=h, w=w, K=K, dtype=abi.F32, layout=abi.NCHW, ld=C,
                      affine_mode=1, hints=0)
    rc = lib.dcfp_class_stats(ctypes.byref(d), sp)
    assert rc == 0, abi.last_error()
    dgamma = torch.empty(C, dtype=torch.float32, device=dev)
    assert lib.dcfp_reduce_classes(S1.data_ptr(), K, C, dgamma.data_ptr(), sp) == 0
    gamma = (torch.rand(C, generator=g) + 0.5).to(dev)
    eic = torch.full((C,), float("nan"), device=dev)
    r32, omr32 = float(np.float32(0.999)), float(np.float32(1.0 - 0.999))
    assert lib.dcfp_eic_update_flat(dgamma.data_ptr(), gamma.data_ptr(), eic.data_ptr(), C, r32, omr32, 1, sp) == 0
    stream.synchronize()
    rc_, r1, r2 = ref.class_stats_bwd(x, dy, mean, invstd, label, K)
    mass = ref.abs_mass(ref.functor_bwd(x, dy, invstd, -mean * invstd), label, K)
    assert torch.equal(cnt.cpu(), rc_)
    assert ((S1.cpu() - r1).abs() <= 2e-5 * mass + 1e-30).all() and ((S2.cpu() - r2).abs() <= 2e-5 * r2 + 1e-30).all()
    exp = eic_ref.eic_step(0, dgamma.cpu().numpy(), gamma.cpu().numpy(), 0.999)
    assert np.array_equal(eic.cpu().numpy().view(np.uint32), exp.view(np.uint32))
    # a rejected call leaves a message and touches nothing
    d.K = 0
    assert lib.dcfp_class_stats(ctypes.byref(d), sp) == -3 and b"K=0" in lib.dcfp_last_error()
    assert lib.dcfp_launch_count(0) > 0
