"""Kernel-level timing of the interaction-block kernels (eqnn_tpconv_fwd / _bwd + the sender-side reduction) at the
cfg-2 size (32 x 5000-node kNN-20 periodic clouds = 3.2 M edges), both generations, with a bit-equality check.

python scripts/bench_tpconv.py [B] [N] [k] [lmax]   (CUDA events, 3 warm-ups, L2 flushed between launches)"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from eqnn_jax_b200 import graph_utils as GU  # noqa: E402
from eqnn_jax_b200 import ops  # noqa: E402


def timeit(fn, n=7, flush=None):
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
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    N = int(sys.argv[2]) if len(sys.argv) > 2 else 5000
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 20
    lmax = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    agg = ops.AGG[os.environ.get("AGG", "mean")]
    rng = np.random.default_rng(0)
    x = torch.tensor(rng.uniform(0.0, 1.0, size=(B, N, 3)).astype(np.float32)).cuda()
    g = GU.build_graph(x, None, k, apply_pbc=GU.get_apply_pbc(None, cell=np.eye(3)))
    rowptr, perm, src = ops.csr_build(g.senders.reshape(B, -1).contiguous(), g.receivers.reshape(B, -1).contiguous(), N)
    rowptr_s, perm_s, _ = ops.csr_build(g.receivers.reshape(B, -1).contiguous(), g.senders.reshape(B, -1).contiguous(), N)
    Nt, E = B * N, B * N * k
    geom, _ = ops.edge_geom(x.reshape(Nt, 3).contiguous(), 3, Nt, rowptr, src, 0, 1.0)
    tp_id = ops.tp_lookup(f"x=1x0e+2x1o;lmax={lmax};live=0e,1o")
    dxp, ny, n_live, d_live = ops.tp_dims(tp_id)
    gen = torch.Generator(device="cuda").manual_seed(0)
    xt = torch.randn((Nt, dxp), device="cuda", generator=gen)
    xt[:, 7:] = 0
    w_t = torch.randn((n_live, E), device="cuda", generator=gen)
    dA = torch.randn((Nt, d_live), device="cuda", generator=gen)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    amax = torch.empty((Nt, d_live), dtype=torch.int32, device="cuda") if agg == 2 else None
    res = {}
    for gen_name, env in (("v1", "1"), ("v2", "0")):
        os.environ["EQNN_TPCONV_V1"] = env
        A = torch.zeros((Nt, d_live), device="cuda")
        dw_t = torch.zeros((n_live, E), device="cuda")
        dxe = torch.zeros((E, dxp), device="cuda")
        dxt = torch.zeros((Nt, dxp), device="cuda")
        f = lambda: ops.tpconv_fwd(tp_id, rowptr, src, geom, w_t, xt, Nt, agg, A, d_live, amax)
        b = lambda: ops.tpconv_bwd(tp_id, rowptr, src, perm, geom, w_t, xt, Nt, agg, dA, d_live, dw_t, dxe, amax)
        r = lambda: ops.segment_gather_sum(dxe, rowptr_s, perm_s, Nt, dxp, dxt)

        def br():
            b()
            r()
        if os.environ.get("ONCE"):          # profiling runs: one launch of each kernel per generation
            f(); b(); r()
            torch.cuda.synchronize()
            continue
        tf, tb, tr, tbr = timeit(f, flush=flush), timeit(b, flush=flush), timeit(r, flush=flush), timeit(br, flush=flush)
        res[gen_name] = (A.clone(), dw_t.clone(), dxe.clone(), dxt.clone())
        tot = tf + tbr
        print(f"{gen_name}: fwd {tf * 1e3:.1f} us  bwd {tb * 1e3:.1f} us  sender reduce {tr * 1e3:.1f} us  bwd+reduce {tbr * 1e3:.1f} us  "
              f"total {tot * 1e3:.1f} us = {E / tot / 1e6:.2f} G edge-steps/s  ({E} edges, n_live {n_live}, d_live {d_live})")
    if os.environ.get("ONCE"):
        return
    for name, a, b_ in zip(("A", "dw_t", "dxe", "dxt"), res["v1"], res["v2"]):
        print(f"  {name}: bit-identical {torch.equal(a, b_)}  max|diff| {float((a - b_).abs().max()):.3e}")


if __name__ == "__main__":
    main()
