import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np, torch
import test_gpu_egnn as T
kw = dict(positions_only=True, soft_edges=False, n_layers=3, d_hidden=128, n_radial_basis=32, r_max=0.6,
          message_passing_steps=int(sys.argv[1]) if len(sys.argv) > 1 else 3, message_passing_agg="mean", task="graph", n_outputs=2)
for tc in (True, False):
    import eqnn_jax_b200.egnn as EGm
    orig = EGm.EGNN.__init__
    got, want, params, p, x = T._run(dict(kw, tensor_cores=tc) if False else kw, pbc=True)
    print("tc", tc, "out err", np.abs(got.cpu().numpy() - want).max() / np.abs(want).max())
    gt = params.grad_tree
    for path, w in T._leaves(p):
        g = gt
        for q in path: g = g[q]
        wn = w.grad.numpy(); e = np.abs(g.cpu().numpy() - wn).max()
        print("  ", "/".join(path), "rel-to-self %.2e" % (e / max(np.abs(wn).max(), 1e-30)), "max %.2e" % np.abs(wn).max())
    break
