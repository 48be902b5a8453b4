"""GPU parity: SEGNN (models/segnn.py) forward and gradients against the CPU oracle (fp64), 1e-5 relative
(per tensor against the largest gradient of its Flax module, as in tests/test_gpu_nequip.py)."""
import numpy as np
import pytest
import torch

from oracle import graph_ref as G
from oracle import nequip_ref as NQ
from oracle import segnn_ref as S
from oracle.e3nn_ref import IrrepsArray as OIA
from oracle.so3 import Irreps as OI

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _leaves(tree, pre=()):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from _leaves(v, pre + (k,))
        else:
            yield pre + (k,), v


def _close(got, want, what, rtol=RTOL, scale=None):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = np.abs(want).max() if scale is None else scale
    err = np.abs(got - want).max() if got.size else 0.0
    assert err <= rtol * max(scale, 1e-30), f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / max(scale, 1e-30):.2e})"


def _run(kw, irreps, x, k, lmax_attr, n_rb, pbc, vel, n_glob=0, lowering="dense", tc=True):
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
    from eqnn_jax_b200.irreps import Irreps
    from eqnn_jax_b200.nequip import IrrepsArray
    from eqnn_jax_b200.segnn import SEGNN
    B, N, D = x.shape
    cell = np.eye(3, dtype=np.float32) * 3.4
    opbc = G.get_apply_pbc(cell=cell) if pbc else None
    gpbc = GU.get_apply_pbc(cell=cell) if pbc else None
    g = G.build_graph(x, None, k, apply_pbc=opbc)
    v = x[..., 3:6] if vel else None
    glob = np.random.default_rng(11).normal(size=(B, n_glob)).astype(np.float32) if n_glob else None
    ost = G.get_equivariant_graph(OIA(irreps, torch.tensor(x, dtype=torch.float64)), x[..., :3], v, g.senders, g.receivers,
                                  g.n_node, g.n_edge, None, glob, lmax_attr, steerable_velocities=vel, apply_pbc=opbc,
                                  n_radial_basis=n_rb, r_max=1.5)
    cfg = S.SEGNNConfig(**kw)
    n_add = max(n_rb, 1)
    tree = S.init_params(cfg, irreps, OI.spherical_harmonics(lmax_attr), f"{n_add}x0e", seed=1, n_globals=n_glob)
    for _, w in _leaves(tree):                                         # non-zero biases: exercise their gradient path
        pass
    rng = np.random.default_rng(7)
    for path, w in _leaves(tree):
        if path[-1].startswith("b[") or path[-1] == "bias":
            w += 0.1 * rng.normal(size=w.shape)
    p = NQ.to_torch(tree, torch.float64, requires_grad=True)
    out = S.apply_batched(p, cfg, ost)
    val = out.array if cfg.task == "node" else out
    cot = rng.normal(size=tuple(val.shape))
    (val * torch.tensor(cot)).sum().backward()

    xt = torch.tensor(x).cuda()
    gst = get_equivariant_graph(IrrepsArray(Irreps(irreps), xt), xt[..., :3].contiguous(),
                                xt[..., 3:6].contiguous() if vel else None, torch.tensor(g.senders).cuda(),
                                torch.tensor(g.receivers).cuda(), g.n_node, g.n_edge, None,
                                None if glob is None else torch.tensor(glob).cuda(), lmax_attr,
                                steerable_velocities=vel, apply_pbc=gpbc, n_radial_basis=n_rb, r_max=1.5)
    model = SEGNN(lowering=lowering, tensor_cores=tc, **kw)
    params = model.init(0, gst)
    params.load_tree(tree)
    cott = torch.tensor(cot, dtype=torch.float32).cuda()
    got = model.forward_backward(params, gst, lambda o: cott)
    return got.cpu().numpy(), val.detach().numpy(), params, p


@pytest.mark.parametrize("kw,irreps,F,lmax_attr,n_rb,pbc,vel,n_glob", [
    # tests/test_equivariance.py:288-310 (node) and :313-335 (graph): 7 features "1o+1o+1x0e", attributes l <= 1
    (dict(message_passing_steps=2, d_hidden=32, task="node", residual=True, output_irreps="1o + 1o + 1x0e"),
     "1o+1o+1x0e", 7, 1, 0, False, True, 0),
    (dict(message_passing_steps=3, d_hidden=32, task="graph", residual=True, output_irreps="1x0e", n_outputs=2),
     "1o+1o+1x0e", 7, 1, 0, False, True, 0),
    # cfg-3 shape (train_velocities.py:364-378,454-458): positions only, l_max_hidden 2, sum, node task -> 1x1o, PBC
    (dict(d_hidden=64, n_layers=3, message_passing_steps=2, message_passing_agg="sum", task="node", output_irreps="1x1o",
          l_max_hidden=2, residual=True), "1o", 3, 1, 0, True, False, 0),
    # train_cosmology.py:185 style: attributes l <= 2, Bessel messages, mean aggregation, graph read-out, no residual
    (dict(d_hidden=64, n_layers=2, message_passing_steps=2, message_passing_agg="mean", task="graph", output_irreps="1x0e",
          l_max_hidden=2, n_outputs=2, residual=False), "1o", 3, 2, 8, True, False, 0),
    # 12 Bessel messages: more additional messages than the node-level first-edge-layer path takes (<= 8), so the
    # concat(add, h_s, h_r) route runs
    (dict(d_hidden=32, n_layers=2, message_passing_steps=2, message_passing_agg="sum", task="node", output_irreps="1x1o",
          l_max_hidden=1, residual=True), "1o", 3, 1, 12, True, False, 0),
    # attributes up to l = 3 (16 components): the widest attribute slot of the node-level edge-layer kernels
    (dict(d_hidden=32, n_layers=2, message_passing_steps=1, message_passing_agg="mean", task="node", output_irreps="1x1o",
          l_max_hidden=2, residual=True), "1o", 3, 3, 0, False, False, 0),
    # scalar_activation = "silu": the class default of TensorProductLinearGate (segnn.py:27), e3nn-normalised
    (dict(d_hidden=32, n_layers=2, message_passing_steps=2, message_passing_agg="mean", task="node", output_irreps="1x1o",
          l_max_hidden=1, residual=True, scalar_activation="silu"), "1o", 3, 1, 4, True, False, 0),
    # jraph.segment_max over the receivers + jnp.max read-out (segnn.py:155,218,265-270): arg-max routed backward
    (dict(d_hidden=32, n_layers=2, message_passing_steps=2, message_passing_agg="max", readout_agg="max", task="graph",
          output_irreps="1x0e", l_max_hidden=1, n_outputs=2, residual=True), "1o", 3, 1, 4, True, False, 0),
    (dict(d_hidden=32, n_layers=3, message_passing_steps=2, message_passing_agg="max", task="node", output_irreps="1x1o",
          l_max_hidden=2, residual=True), "1o", 3, 2, 0, False, False, 0),
    # graph globals (TPCF vector) -> concat + LayerNorm before the read-out MLP (segnn.py:282-286)
    (dict(d_hidden=32, n_layers=2, message_passing_steps=1, message_passing_agg="mean", task="graph", output_irreps="1x0e",
          n_outputs=2), "1o", 3, 1, 0, False, False, 12),
])
@pytest.mark.parametrize("lowering,tc", [("dense", True), ("blocked", True), ("dense", False)])
def test_segnn_forward_and_gradients(kw, irreps, F, lmax_attr, n_rb, pbc, vel, n_glob, lowering, tc):
    rng = np.random.default_rng(len(irreps) + lmax_attr)
    B, N, k = 2, 60, 6
    x = rng.normal(size=(B, N, F)).astype(np.float32)
    x[..., :3] = rng.uniform(0.0, 3.4, size=(B, N, 3))
    got, want, params, p = _run(kw, irreps, x, k, lmax_attr, n_rb, pbc, vel, n_glob, lowering=lowering, tc=tc)
    _close(got, want, "output")
    gt = params.grad_tree
    scales = {}
    for path, w in _leaves(p):
        scales[path[0]] = max(scales.get(path[0], 0.0), float(w.grad.abs().max()) if w.grad is not None else 0.0)
    n_dead = 0
    gmax = max(scales.values())
    for path, w in _leaves(p):
        g = gt
        for q in path:
            g = g[q]
        wn = np.zeros(tuple(w.shape)) if w.grad is None else w.grad.numpy()
        if scales[path[0]] == 0.0:                      # the discarded node-function layers (segnn.py:84-89)
            assert float(g.abs().max()) == 0.0, path
            n_dead += 1
            continue
        # 1e-5 of the module's own gradient scale; modules whose gradients sit orders of magnitude below the dominant
        # flow of the same backward pass inherit that flow's fp32 rounding (1e-7 of it), as in tests/test_gpu_egnn.py
        err = np.abs(g.cpu().numpy().astype(np.float64) - wn).max()
        assert err <= RTOL * scales[path[0]] + 1e-7 * gmax, \
            f"grad {'/'.join(path)}: err {err:.3e}, module scale {scales[path[0]]:.3e}, largest gradient {gmax:.3e}"
    assert n_dead > 0


@pytest.mark.parametrize("lowering", ["dense", "blocked"])
def test_segnn_row_blocks(monkeypatch, lowering):
    """Edge rows are processed in blocks of ROW_BLOCK (weight gradients accumulate across blocks): same result with
    blocks of 100 rows (8 blocks, a ragged last one) as the oracle."""
    from eqnn_jax_b200 import segnn as SG
    monkeypatch.setattr(SG, "ROW_BLOCK", 100)
    rng = np.random.default_rng(5)
    kw = dict(d_hidden=32, n_layers=2, message_passing_steps=1, message_passing_agg="sum", task="node",
              output_irreps="1x1o", l_max_hidden=2, residual=True)
    x = rng.uniform(0.0, 3.4, size=(2, 60, 3)).astype(np.float32)
    got, want, params, p = _run(kw, "1o", x, 6, 1, 0, True, False, 0, lowering=lowering)
    _close(got, want, "output")
    gt = params.grad_tree
    gmax = max(float(w.grad.abs().max()) for _, w in _leaves(p) if w.grad is not None)
    for path, w in _leaves(p):
        if w.grad is None or not path[0].startswith("TensorProductLinearGate"):
            continue
        g = gt
        for q in path:
            g = g[q]
        err = np.abs(g.cpu().numpy().astype(np.float64) - w.grad.numpy()).max()
        assert err <= RTOL * max(float(w.grad.abs().max()), 1e-30) + 1e-6 * gmax, ("/".join(path), err)


def test_segnn_routes_agree_at_cloud_size(monkeypatch):
    """A 3000-node kNN-20 PBC cloud (too large for the oracle in a test): the default route (dense lowering on tensor
    cores, first and last edge layer of every step at the nodes) against the plain per-edge route on the CUDA cores and
    against the irrep-blocked lowering -- outputs and every parameter gradient."""
    from eqnn_jax_b200 import graph_utils as GU, ops
    from eqnn_jax_b200 import segnn as SG
    from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
    from eqnn_jax_b200.irreps import Irreps
    from eqnn_jax_b200.nequip import IrrepsArray
    rng = np.random.default_rng(3)
    B, N, k = 1, 3000, 20
    x = torch.tensor(rng.uniform(0.0, 1.0, size=(B, N, 3)).astype(np.float32)).cuda()
    pbc = GU.get_apply_pbc(cell=np.eye(3, dtype=np.float32))
    s, r, _ = ops.knn_pbc(x, k, pbc.cell, pbc.cell_inv, want_dr=False)
    g = get_equivariant_graph(IrrepsArray(Irreps("1o"), x), x, None, s, r, None, None, None, None, 1, apply_pbc=pbc)
    kw = dict(d_hidden=64, n_layers=3, message_passing_steps=2, message_passing_agg="mean", task="node",
              output_irreps="1x1o", l_max_hidden=2, residual=True)
    cot = torch.tensor(rng.normal(size=(B, N, 3)).astype(np.float32)).cuda()

    def run(per_edge=False, **extra):
        model = SG.SEGNN(**kw, **extra)
        if per_edge:
            monkeypatch.setattr(SG.SEGNN, "_edge0_ok", lambda self, *a: False)
            monkeypatch.setattr(SG.SEGNN, "_edgeL_ok", lambda self, *a: False)
        params = model.init(0, g)
        out = model.forward_backward(params, g, lambda o: cot)
        monkeypatch.undo()
        return out.double().cpu().numpy(), params.grad.double().cpu().numpy(), params

    out0, g0, params = run()
    scale = {}
    for path, shape, off in params.layout:
        n = int(np.prod(shape))
        scale[path[0]] = max(scale.get(path[0], 0.0), float(np.abs(g0[off:off + n]).max()))
    gmax = max(scale.values())
    for what, (out1, g1, _) in {"per-edge, CUDA cores": run(per_edge=True, tensor_cores=False),
                                "blocked lowering": run(lowering="blocked")}.items():
        assert np.abs(out1 - out0).max() <= 2e-5 * np.abs(out0).max(), what
        for path, shape, off in params.layout:
            n = int(np.prod(shape))
            err = np.abs(g1[off:off + n] - g0[off:off + n]).max()
            assert err <= 2e-5 * scale[path[0]] + 2e-7 * gmax, (what, "/".join(path), err, scale[path[0]], gmax)


def test_segnn_reference_equivariance_property():
    """tests/test_equivariance.py:288-310 on the B200 path: f(R x) = R f(x) for the 7-feature node task."""
    from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.irreps import Irreps
    from eqnn_jax_b200.nequip import IrrepsArray
    from eqnn_jax_b200.segnn import SEGNN
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 5, 7)).astype(np.float32)
    axis = np.array([0, 1 / np.sqrt(2), 1 / np.sqrt(2)])
    xr = np.stack([G.rotate_representation(x[b], 45.0, axis) for b in range(2)]).astype(np.float32)
    model = SEGNN(message_passing_steps=2, d_hidden=32, task="node", residual=True, output_irreps="1o + 1o + 1x0e")

    def graph(a):
        t = torch.tensor(a).cuda()
        g = GU.build_graph(t, None, 5)
        return get_equivariant_graph(IrrepsArray(Irreps("1o+1o+1x0e"), t), t[..., :3].contiguous(), t[..., 3:6].contiguous(),
                                     g.senders, g.receivers, g.n_node, g.n_edge, None, None, 1)
    g0, g1 = graph(x), graph(xr)
    params = model.init(0, g0)
    o0 = model.apply(params, g0).nodes.array.cpu().numpy()
    o1 = model.apply(params, g1).nodes.array.cpu().numpy()
    want = np.stack([G.rotate_representation(o0[i], 45.0, axis) for i in range(2)])
    assert not np.allclose(o0, o1, rtol=0.05)
    assert np.allclose(o1, want, rtol=0.05, atol=1e-5)
    _close(o1, want, "equivariance", rtol=1e-4)


def test_segnn_unsupported_fails_loudly():
    from eqnn_jax_b200.segnn import SEGNN, _Plan
    from eqnn_jax_b200.irreps import Irreps
    with pytest.raises(NotImplementedError):
        _Plan(SEGNN(scalar_activation="relu"), Irreps("1o"), 1, 1)
    with pytest.raises(ValueError):
        _Plan(SEGNN(message_passing_agg="median"), Irreps("1o"), 1, 1)
