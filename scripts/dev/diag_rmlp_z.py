"""Error of the stashed pre-activations z_l (fused forward kernel) against the fp64 oracle, per layer."""
import math, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from eqnn_jax_b200 import ops
from eqnn_jax_b200.nequip import normalize_constant
from oracle import e3nn_ref as R

def unblock(stash, E, H):
    ep = (E + 127) // 128 * 128
    out = []
    for l in range(H):
        b = stash[l * ep * 128:(l + 1) * ep * 128].view(ep // 32, 32, 32, 4)   # [chunk][c4][row][4]
        out.append(b.permute(0, 2, 1, 3).reshape(ep, 128)[:E])
    return out

def run(E, seed=0, dlo=0.005):
    rng = np.random.default_rng(seed)
    n_rb, H, n_out = 64, 3, 8
    d = rng.uniform(dlo, 1.5, size=E).astype(np.float32)
    d[rng.random(E) < 0.02] = 0.0
    dims = [n_rb] + [128] * H + [n_out]
    Ws = [rng.normal(size=(dims[i], dims[i + 1])).astype(np.float32) for i in range(len(dims) - 1)]
    act = 1.0 / normalize_constant("gelu")
    geom = torch.zeros((E, 4), device="cuda"); geom[:, 3] = torch.tensor(d, device="cuda")
    Wd = [torch.tensor(W, device="cuda") for W in Ws]
    wl = [(W, 0, W.shape[1]) for W in Wd]
    stash = ops.radial_mlp_stash(E, H, geom.device)
    w_t = torch.empty((n_out, E), device="cuda")
    ops.radial_mlp_fwd(geom, n_rb, 0.6, wl, n_out, act, w_t, z_stash=stash)
    torch.cuda.synchronize()
    zs = unblock(stash, E, H)
    a = R.bessel(torch.tensor(d, dtype=torch.float64), n_rb, 0.6)
    msg = []
    for l, W in enumerate(Ws):
        z = a @ torch.tensor(W, dtype=torch.float64) / math.sqrt(W.shape[0])
        if l < H:
            got = zs[l].double().cpu()
            err = (got - z).abs()
            rowrel = (err.max(dim=1).values / z.abs().max(dim=1).values)
            msg.append(f"z{l+1}: max {float(err.max()/z.abs().max()):.1e} rowrel med {float(rowrel.median()):.1e} max {float(rowrel.max()):.1e} |z|max {float(z.abs().max()):.0f}")
            a = R.gelu_tanh(z) * act
        else:
            got = w_t.double().cpu().T
            msg.append(f"out: {float((got - z).abs().max()/z.abs().max()):.1e}")
    print(f"E={E} dlo={dlo}: " + " | ".join(msg), flush=True)

run(19021)
run(19021, dlo=0.1)
run(4096)
