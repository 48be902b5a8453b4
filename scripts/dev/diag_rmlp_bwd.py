"""Per-layer error of eqnn_radial_mlp_bwd against fp64 autograd for a few cotangent distributions."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eqnn_jax_b200 import ops
from eqnn_jax_b200.nequip import normalize_constant
from oracle import e3nn_ref as R

def run(E, n_rb, n_hidden, n_out, gmag, tail, seed=0):
    rng = np.random.default_rng(seed)
    d = rng.uniform(0.005, 1.5, size=E).astype(np.float32)
    d[rng.random(E) < 0.02] = 0.0
    dims = [n_rb] + [128] * n_hidden + [n_out]
    Ws = [rng.normal(size=(dims[i], dims[i + 1])).astype(np.float32) for i in range(len(dims) - 1)]
    dw = (rng.normal(size=(n_out, E)) * np.exp(rng.normal(size=(1, E)) * tail) * gmag).astype(np.float32)
    act = 1.0 / normalize_constant("gelu")
    geom = torch.zeros((E, 4), device="cuda"); geom[:, 3] = torch.tensor(d, device="cuda")
    Wd = [torch.tensor(W, device="cuda") for W in Ws]
    gW = [torch.full_like(W, float("nan")) for W in Wd]
    ops.radial_mlp_bwd(geom, n_rb, 0.6, [(W, 0, W.shape[1]) for W in Wd], n_out, act, torch.tensor(dw, device="cuda"),
                       [(g, 0, g.shape[1]) for g in gW])
    torch.cuda.synchronize()
    Wt = [torch.tensor(W, dtype=torch.float64, requires_grad=True) for W in Ws]
    a = R.bessel(torch.tensor(d, dtype=torch.float64), n_rb, 0.6)
    for l, W in enumerate(Wt):
        a = a @ W / math.sqrt(W.shape[0])
        if l < len(Wt) - 1:
            a = R.gelu_tanh(a) * act
    (a * torch.tensor(dw, dtype=torch.float64).T).sum().backward()
    errs = [float((g.double().cpu() - W.grad).abs().max() / W.grad.abs().max()) for g, W in zip(gW, Wt)]
    # the same computation in plain fp32 (what an fp32 reference implementation gets): the conditioning floor
    W32 = [torch.tensor(W, dtype=torch.float32, requires_grad=True) for W in Ws]
    a = R.bessel(torch.tensor(d, dtype=torch.float32), n_rb, 0.6)
    for l, W in enumerate(W32):
        a = a @ W / math.sqrt(W.shape[0])
        if l < len(W32) - 1:
            a = R.gelu_tanh(a) * act
    (a * torch.tensor(dw, dtype=torch.float32).T).sum().backward()
    floor = [float((w32.grad.double() - W.grad).abs().max() / W.grad.abs().max()) for w32, W in zip(W32, Wt)]
    print(f"E={E} n_rb={n_rb} H={n_hidden} n_out={n_out} gmag={gmag:g} tail={tail}: " + " ".join(f"{e:.1e}" for e in errs) + "   fp32 floor: " + " ".join(f"{e:.1e}" for e in floor), flush=True)

for E in (4096, 19021, 19072, 200000):
    run(E, 64, 3, 8, 1.0, 0.0)
