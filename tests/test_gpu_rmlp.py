"""Fused radial MLP (eqnn_radial_mlp_fwd / _bwd) against the fp64 oracle restatement of models/nequip.py:55-68.

Tolerance: 1e-5 of the largest magnitude of each output tensor (north_star: fp32 features within 1e-5 relative);
gradients additionally get 4x the deviation of a plain fp32 evaluation from fp64."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

TOL = 1e-5


def _oracle_forward(d, Ws, r_cut, act_scale):
    """fp64 restatement: bessel with the reference's fp32 phase rounding, Dense / sqrt(fan_in), gelu / c."""
    from oracle import e3nn_ref as R
    a = R.bessel(torch.tensor(d, dtype=torch.float64), Ws[0].shape[0], r_cut)
    zs, acts = [], [a]
    for l, W in enumerate(Ws):
        z = a @ torch.tensor(W, dtype=torch.float64) / math.sqrt(W.shape[0])
        zs.append(z)
        if l < len(Ws) - 1:
            a = R.gelu_tanh(z) * act_scale
            acts.append(a)
    return zs, acts


def _inputs(E, n_rb, n_hidden, n_out, seed, zero_frac=0.02):
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.005, 1.5, size=E).astype(np.float32)
    d[rng.random(E) < zero_frac] = 0.0             # self-loops: the d = 0 limit of the basis
    dims = [n_rb] + [128] * n_hidden + [n_out]
    Ws = [rng.normal(size=(dims[i], dims[i + 1])).astype(np.float32) for i in range(len(dims) - 1)]
    return d, Ws


@pytest.mark.parametrize("E,n_rb,n_hidden,n_out", [
    (128 * 148 * 2 + 77, 64, 3, 8),     # cfg-2 model, two tiles per CTA + a ragged one
    (4096, 64, 3, 8),
    (1000, 32, 2, 11),
    (129, 16, 1, 16),
    (5, 64, 3, 1),
    (128 * 148 * 3, 64, 2, 8),          # odd number of tiles per CTA (one slot idle in the last pair)
])
def test_radial_mlp_fwd_matches_oracle(E, n_rb, n_hidden, n_out):
    from eqnn_jax_b200 import ops
    from eqnn_jax_b200.nequip import normalize_constant
    r_cut = 0.6
    act_scale = 1.0 / normalize_constant("gelu")
    d, Ws = _inputs(E, n_rb, n_hidden, n_out, seed=E % 1000)
    assert ops.radial_mlp_supported(n_rb, 128, n_hidden, n_out)
    geom = torch.zeros((E, 4), dtype=torch.float32, device="cuda")
    geom[:, 3] = torch.tensor(d, device="cuda")
    Wd = [torch.tensor(W, device="cuda") for W in Ws]
    w_t = torch.full((n_out, E), float("nan"), dtype=torch.float32, device="cuda")
    ops.radial_mlp_fwd(geom, n_rb, r_cut, [(W, 0, W.shape[1]) for W in Wd], n_out, act_scale, w_t)
    torch.cuda.synchronize()
    zs, _ = _oracle_forward(d, Ws, r_cut, act_scale)
    want = zs[-1].T
    got = w_t.double().cpu()
    assert torch.isfinite(got).all()
    err = float((got - want).abs().max() / want.abs().max())
    assert err < TOL, f"radial MLP forward: {err:.2e}"


def test_radial_mlp_fwd_rejects_unsupported_shapes():
    from eqnn_jax_b200 import ops
    assert not ops.radial_mlp_supported(8, 128, 3, 8)
    assert not ops.radial_mlp_supported(64, 64, 3, 8)
    assert not ops.radial_mlp_supported(64, 128, 4, 8)
    assert not ops.radial_mlp_supported(64, 128, 3, 17)


@pytest.mark.parametrize("E,n_rb,n_hidden,n_out,gmag", [
    (128 * 148 * 2 + 77, 64, 3, 8, 1e-6), # cfg-2 model; tiny gradients exercise the power-of-two scale
    (128 * 148 * 20 + 5, 64, 3, 8, 1.0),  # 20 tiles per CTA: the fused layer kernel's weight-gradient accumulator rolls over
    (4096, 64, 3, 8, 1.0),
    (5000, 32, 2, 11, 3e3),
    (4200, 16, 1, 16, 1e-3),
    (300, 64, 3, 8, 1.0),                  # below the tensor-core weight-gradient threshold (CUDA-core GEMMs)
])
@pytest.mark.parametrize("variant", ["stash", "recompute", "stash-chain", "stash-bessel"])
def test_radial_mlp_bwd_matches_oracle(E, n_rb, n_hidden, n_out, gmag, variant, monkeypatch):
    from eqnn_jax_b200 import ops
    no_radial = variant == "stash-bessel"    # first layer's weight gradient generates the Bessel basis in the kernel
    if no_radial:
        variant = "stash"
    if variant == "stash-chain":             # the chain kernel + separate weight gradients instead of rmlp_lbwd_kernel
        monkeypatch.setenv("EQNN_RMLP_NO_LBWD", "1")
        variant = "stash"
    from eqnn_jax_b200.nequip import normalize_constant
    r_cut = 0.6
    act_scale = 1.0 / normalize_constant("gelu")
    d, Ws = _inputs(E, n_rb, n_hidden, n_out, seed=7 + E % 1000)
    rng = np.random.default_rng(E)
    # heavy-tailed cotangent: a few edges carry most of the gradient, most are orders of magnitude smaller.  The kernels
    # bring the gradients into fp16 range with ONE power-of-two scale per launch, so rows more than ~2^13 below the largest
    # row keep only the hi half of the (hi, lo) split: the log-normal tail is exp(2 sigma) (max / median ~ 4e3 at 4e4
    # edges) for the small cases and exp(sigma) (~1e2) for the 379 k-edge one, which is there for the accumulator rollover.
    tail = 1.0 if E > 100000 else 2.0
    dw = (rng.normal(size=(n_out, E)) * np.exp(rng.normal(size=(1, E)) * tail) * gmag).astype(np.float32)
    geom = torch.zeros((E, 4), dtype=torch.float32, device="cuda")
    geom[:, 3] = torch.tensor(d, device="cuda")
    Wd = [torch.tensor(W, device="cuda") for W in Ws]
    gW = [torch.full_like(W, float("nan")) for W in Wd]
    wl = [(W, 0, W.shape[1]) for W in Wd]
    stash = radial = None
    if variant == "stash":
        # what the forward pass leaves behind: the pre-activations (fused kernel) and the Bessel basis (eqnn_edge_geom)
        stash = ops.radial_mlp_stash(E, n_hidden, geom.device)
        w_t = torch.empty((n_out, E), dtype=torch.float32, device="cuda")
        ops.radial_mlp_fwd(geom, n_rb, r_cut, wl, n_out, act_scale, w_t, z_stash=stash)
        radial = ops.edge_features(torch.cat([geom[:, 3:4], torch.zeros_like(geom[:, :2])], dim=1).contiguous(), n_rb, r_cut)
        if no_radial:
            if E < 4096 or n_rb not in (32, 64):
                pytest.skip("the in-kernel basis covers n_rb 32 / 64 from 4096 edges on")
            radial = None
    ops.radial_mlp_bwd(geom, n_rb, r_cut, wl, n_out, act_scale, torch.tensor(dw, device="cuda"),
                       [(g, 0, g.shape[1]) for g in gW], z_stash=stash, radial=radial)
    torch.cuda.synchronize()
    # oracle: autograd through the fp64 restatement
    from oracle import e3nn_ref as R
    Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in Ws]
    a = R.bessel(torch.tensor(d, dtype=torch.float64), n_rb, r_cut)
    for l, W in enumerate(Wt):
        a = a @ W / math.sqrt(W.shape[0])
        if l < len(Wt) - 1:
            a = R.gelu_tanh(a) * act_scale
    (a * torch.tensor(dw, dtype=torch.float64).T).sum().backward()
    # conditioning floor: the same arithmetic in plain fp32.  The d = 0 and small-d rows of the basis reach 1e2..1e3,
    # so the pre-activations do too, and gelu' turns their rounding into gradient errors above 1e-5 in any fp32 code.
    W32 = [torch.tensor(W, dtype=torch.float32, requires_grad=True) for W in Ws]
    a = R.bessel(torch.tensor(d, dtype=torch.float32), n_rb, r_cut)
    for l, W in enumerate(W32):
        a = a @ W / math.sqrt(W.shape[0])
        if l < len(W32) - 1:
            a = R.gelu_tanh(a) * act_scale
    (a * torch.tensor(dw, dtype=torch.float32).T).sum().backward()
    for l, (g, W, w32) in enumerate(zip(gW, Wt, W32)):
        got, want = g.double().cpu(), W.grad
        assert torch.isfinite(got).all(), f"layer {l}: non-finite gradient"
        scale = float(want.abs().max())
        err = float((got - want).abs().max())
        floor = float((w32.grad.double() - want).abs().max())
        if os.environ.get("EQNN_TEST_VERBOSE"):
            print(f"layer {l}: err/scale {err / scale:.2e}  floor/scale {floor / scale:.2e}")
        assert err <= TOL * scale + 4 * floor, \
            f"radial MLP weight gradient, layer {l}: err {err / scale:.2e}, fp32 floor {floor / scale:.2e} (of max|want|)"
