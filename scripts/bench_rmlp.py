"""Kernel-level timing of the fused radial-MLP kernels at the cfg-2 size (3.2 M edges, 64 -> 128 x3 -> 8).

python scripts/bench_rmlp.py [E]   (CUDA events on the launching stream, 3 warm-ups, L2 flushed between launches)"""
import math
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqnn_jax_b200 import ops  # noqa: E402
from eqnn_jax_b200.nequip import normalize_constant  # noqa: E402


def timeit(fn, n=5, flush=None):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        if flush is not None:
            flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main():
    E = int(sys.argv[1]) if len(sys.argv) > 1 else 3_200_000
    n_rb, n_hidden, n_out, r_cut = 64, 3, 8, 0.6
    act = 1.0 / normalize_constant("gelu")
    g = torch.Generator(device="cuda").manual_seed(0)
    geom = torch.zeros((E, 4), device="cuda")
    geom[:, 3] = torch.rand(E, device="cuda", generator=g) * 0.5 + 0.01
    dims = [n_rb] + [128] * n_hidden + [n_out]
    Ws = [torch.randn(dims[i], dims[i + 1], device="cuda", generator=g) for i in range(len(dims) - 1)]
    w_t = torch.empty((n_out, E), device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    wl = [(W, 0, W.shape[1]) for W in Ws]
    ms = timeit(lambda: ops.radial_mlp_fwd(geom, n_rb, r_cut, wl, n_out, act, w_t), flush=flush)
    fl = 2.0 * sum(dims[i] * dims[i + 1] for i in range(len(dims) - 1)) * E
    print(f"radial_mlp_fwd: {ms:.3f} ms for {E} edges  ({E / ms / 1e6:.2f} G edges/s, {fl / ms / 1e9:.1f} TFLOP/s algorithmic, "
          f"{3 * fl / ms / 1e9:.1f} executed fp16)")
    dw_t = torch.randn((n_out, E), device="cuda", generator=g) * 1e-4
    gw = [(torch.empty_like(W), 0, W.shape[1]) for W in Ws]
    stash = ops.radial_mlp_stash(E, n_hidden, geom.device)
    ms = timeit(lambda: ops.radial_mlp_fwd(geom, n_rb, r_cut, wl, n_out, act, w_t, z_stash=stash), flush=flush)
    print(f"radial_mlp_fwd + z stash: {ms:.3f} ms")
    radial = ops.edge_features(torch.cat([geom[:, 3:4], torch.zeros_like(geom[:, :2])], dim=1).contiguous(), n_rb, r_cut)
    ms = timeit(lambda: ops.radial_mlp_bwd(geom, n_rb, r_cut, wl, n_out, act, dw_t, gw, z_stash=stash, radial=radial), flush=flush)
    print(f"radial_mlp_bwd (stash): {ms:.3f} ms for {E} edges")
    ms = timeit(lambda: ops.radial_mlp_bwd(geom, n_rb, r_cut, wl, n_out, act, dw_t, gw), flush=flush)
    print(f"radial_mlp_bwd (recompute): {ms:.3f} ms for {E} edges")


if __name__ == "__main__":
    main()
