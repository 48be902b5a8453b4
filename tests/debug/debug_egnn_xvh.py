"""Per-tensor gradient errors of the (x, v, h) EGNN against the fp64 oracle (and the oracle's own fp32 error)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))), "tests"))
import numpy as np
import test_gpu_egnn as T

kw, pbc, F = T.test_egnn_xvh_forward_and_gradients.pytestmark[0].args[1][int(sys.argv[1]) if len(sys.argv) > 1 else 0]
for tc in (True, False):
    kw2 = dict(kw)
    got, want, params, p, x, p32 = T._run(kw2, pbc=pbc, F=F, seed=3) if tc else (None,) * 6
    if not tc:
        import eqnn_jax_b200.egnn as EGm
        orig = EGm.EGNN.__init__
        got, want, params, p, x, p32 = T._run(dict(kw, ), pbc=pbc, F=F, seed=3)
    gt = params.grad_tree
    print("tensor_cores", tc, "out err", np.abs(got.cpu().numpy() - want).max())
    for (path, w), (_, w32) in zip(T._leaves(p), T._leaves(p32)):
        g = gt
        for q in path:
            g = g[q]
        wn = w.grad.numpy()
        err = np.abs(g.cpu().numpy().astype(np.float64) - wn).max()
        e32 = np.abs(w32.grad.double().numpy() - wn).max()
        print(f"{'/'.join(path):28s} scale {np.abs(wn).max():.3e}  err {err:.3e}  oracle32 {e32:.3e}  ratio {err / max(e32, 1e-30):.1f}")
    break
