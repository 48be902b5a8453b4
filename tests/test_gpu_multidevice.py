"""One process, several devices, one host thread per device (the jax.pmap driving pattern): the C ABI keeps no
per-process launch state -- kernels with more than 48 KB of dynamic shared memory are configured per (function,
device), see tc_common.cuh:configure_smem -- so both devices must produce the single-device result."""
import threading

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _work(dev, out, err):
    try:
        from eqnn_jax_b200 import ops
        from eqnn_jax_b200.nequip import _mm, normalize_constant
        torch.cuda.set_device(dev)
        g = torch.Generator().manual_seed(0)
        E = 8192
        d = (torch.rand(E, generator=g) * 0.5 + 0.01)
        geom = torch.zeros((E, 4), device=f"cuda:{dev}")
        geom[:, 3] = d.to(geom.device)
        dims = [64, 128, 128, 128, 8]
        Ws = [torch.randn(dims[i], dims[i + 1], generator=g).to(geom.device) for i in range(4)]
        wl = [(W, 0, W.shape[1]) for W in Ws]
        act = 1.0 / normalize_constant("gelu")
        res = []
        for rep in range(3):
            w_t = torch.empty((8, E), device=geom.device)
            stash = ops.radial_mlp_stash(E, 3, geom.device)
            ops.radial_mlp_fwd(geom, 64, 0.6, wl, 8, act, w_t, z_stash=stash)             # rmlp_fwd_kernel: 168 KB smem
            radial = ops.edge_features(torch.cat([geom[:, 3:4], torch.zeros_like(geom[:, :2])], 1).contiguous(), 64, 0.6)
            gw = [torch.empty_like(W) for W in Ws]
            dw = torch.randn((8, E), generator=g).to(geom.device)
            ops.radial_mlp_bwd(geom, 64, 0.6, wl, 8, act, dw, [(t, 0, t.shape[1]) for t in gw], z_stash=stash, radial=radial)
            A = torch.randn(E, 128, generator=g).to(geom.device)                          # tc_rowgemm_kernel
            C = torch.empty((E, 128), device=geom.device)
            _mm(E, 128, 128, 0.5, A, 0, 128, Ws[1], 0, 128, C, 0, 128, tf32x3=True)
            res = [w_t.cpu(), C.cpu()] + [t.cpu() for t in gw]
        torch.cuda.synchronize(dev)
        out[dev] = res
    except Exception as e:          # surfaced in the main thread
        err[dev] = e


def test_two_devices_from_two_threads_match_one_device():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs in one process")
    out, err = {}, {}
    _work(0, out, err)
    assert not err, err
    ref = out[0]
    out, err = {}, {}
    ts = [threading.Thread(target=_work, args=(d, out, err)) for d in (0, 1)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not err, err
    for d in (0, 1):
        for a, b in zip(out[d], ref):
            assert torch.equal(a, b), f"device {d} differs from the single-device result"
