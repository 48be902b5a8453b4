"""GPU parity: EGNN (positions-only coordinate layers) forward and gradients against the CPU oracle.

Outputs: 1e-5 relative.  Gradients: EGNN re-derives the Bessel features from the UPDATED fp32 positions in every
layer (phase derivative j*pi/r_max ~ 170 per unit length), so any fp32 evaluation -- including the reference's
own -- sits 2e-5 .. 2e-4 away from the fp64 value in the gradients of the early layers.  The oracle is therefore
run in fp64 AND in fp32 (the reference's precision); a gradient tensor must be within
    1e-5 * scale  +  4 * max|oracle_fp32 - oracle_fp64|      (both maxima over the Flax module the tensor lives in)
of the fp64 oracle, i.e. as close to the truth as the reference's arithmetic allows (measured: same size)."""
import numpy as np
import pytest
import torch

from oracle import egnn_ref as EG
from oracle import graph_ref as G
from oracle import nequip_ref as NQ

pytestmark = pytest.mark.gpu
RTOL = 1e-5


def _close(got, want, what, scale=None, rtol=RTOL):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, (what, got.shape, want.shape)
    scale = np.abs(want).max() if scale is None else scale
    err = np.abs(got - want).max()
    assert err <= rtol * max(scale, 1e-30), f"{what}: max err {err:.3e} vs scale {scale:.3e} (rel {err / max(scale, 1e-30):.2e})"


def _leaves(tree, pre=()):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from _leaves(v, pre + (k,))
        else:
            yield pre + (k,), v


def _run(kw, B=2, N=256, k=8, pbc=False, seed=0, F=3, n_glob=0):
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.egnn import EGNN
    rng = np.random.default_rng(seed)
    x = rng.uniform(-1.0, 1.7, size=(B, N, 3)).astype(np.float32)      # non-zero mean: the read-out pools positions
    if F > 3 and kw.get("positions_only", False):                      # (x, h): O(1) scalars
        x = np.concatenate([x, rng.normal(size=(B, N, F - 3))], -1).astype(np.float32)
    elif F > 3:                                                        # (x, v, h): small velocities, O(1) scalars
        x = np.concatenate([x, 0.05 * rng.normal(size=(B, N, 3)), rng.normal(size=(B, N, F - 6))], -1).astype(np.float32)
    cell = np.eye(3, dtype=np.float32) * 3.4
    opbc = G.get_apply_pbc(cell=cell) if pbc else None
    gpbc = GU.get_apply_pbc(cell=cell) if pbc else None
    glob = rng.normal(size=(B, n_glob)).astype(np.float32) if n_glob else None
    g = G.build_graph(x, glob, k, apply_pbc=opbc)
    cfg = EG.EGNNConfig(**kw, apply_pbc=opbc)
    tree = EG.init_params(cfg, F, seed=seed + 1, n_globals=n_glob)
    # make the coordinate updates O(1e-3..1e-2) instead of O(1e-4) so that every path carries signal, without
    # turning the layer stack into a chaotic map (large moves amplify fp32 rounding through the Bessel phase)
    for s in range(cfg.message_passing_steps):
        tree[f"Dense_{s}"]["kernel"] = tree[f"Dense_{s}"]["kernel"] * (4.0 if cfg.message_passing_agg == "mean" else 0.5)
    model = EGNN(**kw, apply_pbc=gpbc)
    gg = GU.GraphsTuple(nodes=torch.tensor(x).cuda(), edges=None, receivers=torch.tensor(g.receivers).cuda(),
                        senders=torch.tensor(g.senders).cuda(), globals=None if glob is None else torch.tensor(glob).cuda(),
                        n_node=None, n_edge=None)
    params = model.init(0, gg)
    params.load_tree(tree)
    p = NQ.to_torch(tree, torch.float64, requires_grad=True)
    out = EG.apply_batched(p, cfg, g._replace(nodes=torch.tensor(x, dtype=torch.float64)))
    val = out.nodes if cfg.task == "node" else out
    cot = rng.normal(size=tuple(val.shape))
    (val * torch.tensor(cot)).sum().backward()
    p32 = NQ.to_torch(tree, torch.float32, requires_grad=True)
    out32 = EG.apply_batched(p32, cfg, g._replace(nodes=torch.tensor(x)))
    ((out32.nodes if cfg.task == "node" else out32) * torch.tensor(cot, dtype=torch.float32)).sum().backward()
    cott = torch.tensor(cot, dtype=torch.float32).cuda()
    got = model.forward_backward(params, gg, lambda o: cott)
    return got, val.detach().numpy(), params, p, x, p32


@pytest.mark.parametrize("kw,pbc", [
    (dict(positions_only=True, soft_edges=False, n_layers=3, d_hidden=128, n_radial_basis=32, r_max=0.6,
          message_passing_steps=3, message_passing_agg="mean", task="graph", n_outputs=2), True),     # cfg-4 model
    (dict(positions_only=True, soft_edges=False, n_layers=3, d_hidden=128, n_radial_basis=32, r_max=0.6,
          message_passing_steps=2, message_passing_agg="sum", task="node", tanh_out=True), False),
    (dict(positions_only=True, soft_edges=False, n_layers=4, d_hidden=128, n_radial_basis=64, r_max=0.6,
          message_passing_steps=2, message_passing_agg="mean", task="graph", n_outputs=2, tanh_out=True), True),
    (dict(positions_only=True, soft_edges=False, n_layers=2, d_hidden=32, n_radial_basis=0, r_max=0.6,
          message_passing_steps=2, message_passing_agg="mean", task="node"), False),                  # CUDA-core path
    (dict(positions_only=True, soft_edges=True, n_layers=4, d_hidden=128, n_radial_basis=64, r_max=0.6,
          message_passing_steps=3, message_passing_agg="mean", task="graph", n_outputs=2, tanh_out=True), False),  # notebook cell 14
    (dict(positions_only=True, soft_edges=True, n_layers=3, d_hidden=32, n_radial_basis=8, r_max=0.6,
          message_passing_steps=2, message_passing_agg="mean", task="node", tanh_out=True), True),    # CUDA-core path
    # jraph.segment_max of the translations + jnp.max read-out (egnn.py:282-284,326): arg-max routed backward
    (dict(positions_only=True, soft_edges=False, n_layers=3, d_hidden=128, n_radial_basis=32, r_max=0.6,
          message_passing_steps=2, message_passing_agg="max", readout_agg="max", task="graph", n_outputs=2), True),
    (dict(positions_only=True, soft_edges=True, n_layers=3, d_hidden=32, n_radial_basis=8, r_max=0.6,
          message_passing_steps=2, message_passing_agg="max", task="node", tanh_out=True), False),
])
def test_egnn_forward_and_gradients(kw, pbc):
    got, want, params, p, x, p32 = _run(kw, pbc=pbc)
    got = got.cpu().numpy()
    if kw["task"] == "node":
        # outputs are positions (O(1), fp32 resolution ~1e-7): 1e-5 on the positions, and the UPDATE x' - x itself,
        # which is what the layers compute, to 1e-3 of its own size (it is a difference of fp32 positions)
        _close(got, want, "positions")
        _close(got - x, want - x, "coordinate update", rtol=1e-3)
    else:
        _close(got, want, "output")
    gt = params.grad_tree
    scales = {}
    for path, w in _leaves(p):
        scales[path[0]] = max(scales.get(path[0], 0.0), float(w.grad.abs().max()))
    floors = {}
    for (path, w), (_, w32) in zip(_leaves(p), _leaves(p32)):
        floors[path[0]] = max(floors.get(path[0], 0.0), float((w32.grad.double() - w.grad).abs().max()))
    for path, w in _leaves(p):
        g = gt
        for q in path:
            g = g[q]
        wn = w.grad.numpy()
        floor = 4.0 * floors[path[0]]                                     # what fp32 arithmetic itself costs here
        err = np.abs(g.cpu().numpy().astype(np.float64) - wn).max()
        assert err <= RTOL * scales[path[0]] + floor, \
            f"grad {'/'.join(path)}: err {err:.3e}, scale {scales[path[0]]:.3e}, fp32 floor {floor:.3e}"


XVH = dict(positions_only=False, n_layers=3, r_max=0.6, message_passing_steps=2, n_outputs=2)


@pytest.mark.parametrize("kw,pbc,F,n_glob", [
    # graph globals (TPCF vector, train_cosmology.py --use_tpcf): broadcast to edges and nodes, LayerNorm read-out
    (dict(positions_only=True, soft_edges=False, n_layers=3, d_hidden=128, n_radial_basis=32, r_max=0.6,
          message_passing_steps=2, message_passing_agg="mean", task="graph", n_outputs=2), True, 3, 10),
    (dict(XVH, positions_only=True, soft_edges=True, d_hidden=32, n_radial_basis=8, message_passing_agg="mean",
          task="graph"), False, 6, 4),
    (dict(XVH, soft_edges=False, decouple_pos_vel_updates=True, d_hidden=32, n_radial_basis=8,
          message_passing_agg="sum", task="graph"), False, 8, 5),
    (dict(XVH, soft_edges=True, tanh_out=True, decouple_pos_vel_updates=True, d_hidden=32, n_radial_basis=16,
          message_passing_agg="mean", task="node", message_passing_steps=3), False, 7, 0),   # tests/test_equivariance.py:224-250
    (dict(XVH, soft_edges=False, decouple_pos_vel_updates=False, d_hidden=128, n_radial_basis=32,
          message_passing_agg="mean", task="graph"), True, 9, 0),                            # tensor-core edge MLPs
    (dict(XVH, soft_edges=True, tanh_out=True, decouple_pos_vel_updates=True, d_hidden=64, n_radial_basis=0,
          message_passing_agg="sum", task="graph"), False, 8, 0),
    # (x, h): train_cosmology.py --feats all with EGNN_PARAMS (positions_only=True, velocities as scalars)
    (dict(XVH, positions_only=True, soft_edges=False, d_hidden=128, n_radial_basis=32, message_passing_agg="mean",
          task="graph"), True, 6, 0),
    (dict(XVH, positions_only=True, soft_edges=True, tanh_out=True, d_hidden=32, n_radial_basis=8,
          message_passing_agg="sum", task="node", message_passing_steps=3), False, 5, 0),
    # (x, v): train_cosmology.py --feats all with positions_only=False; the node width grows to 6 + d_hidden (egnn.py:196)
    (dict(XVH, soft_edges=False, decouple_pos_vel_updates=True, d_hidden=32, n_radial_basis=8,
          message_passing_agg="mean", task="node", message_passing_steps=3), False, 6, 0),
    (dict(XVH, soft_edges=True, tanh_out=True, decouple_pos_vel_updates=False, d_hidden=64, n_radial_basis=16,
          message_passing_agg="sum", task="graph"), True, 6, 0),
    # segment_max of the messages m_i (read by the (x, h), (x, v) and (x, v, h) node functions) and of the translations
    (dict(XVH, positions_only=True, soft_edges=True, d_hidden=32, n_radial_basis=8, message_passing_agg="max",
          readout_agg="max", task="graph"), False, 6, 0),
    (dict(XVH, soft_edges=False, decouple_pos_vel_updates=True, d_hidden=32, n_radial_basis=8,
          message_passing_agg="max", task="node", message_passing_steps=3), False, 6, 0),
    (dict(XVH, soft_edges=True, tanh_out=True, decouple_pos_vel_updates=True, d_hidden=32, n_radial_basis=16,
          message_passing_agg="max", task="graph"), True, 7, 0),
])
def test_egnn_xvh_forward_and_gradients(kw, pbc, F, n_glob):
    """(x, v, h) node layout, egnn.py:47-60 and :198-229; (x, h) layout, :42-46 and :152-166."""
    got, want, params, p, x, p32 = _run(kw, pbc=pbc, F=F, seed=3, n_glob=n_glob)
    got = got.cpu().numpy()
    if kw["task"] == "node":
        _close(got, want, "nodes")
        _close(got[..., :3] - x[..., :3], want[..., :3] - x[..., :3], "coordinate update", rtol=1e-3)
        if not kw.get("positions_only", False):
            assert np.array_equal(got[..., 3:6], x[..., 3:6])                  # the old velocity is returned (:229)
    else:
        _close(got, want, "output")
    gt = params.grad_tree
    scales, floors = {}, {}
    for (path, w), (_, w32) in zip(_leaves(p), _leaves(p32)):
        scales[path[0]] = max(scales.get(path[0], 0.0), float(w.grad.abs().max()))
        floors[path[0]] = max(floors.get(path[0], 0.0), float((w32.grad.double() - w.grad).abs().max()))
    gmax = max(scales.values())
    for path, w in _leaves(p):
        g = gt
        for q in path:
            g = g[q]
        err = np.abs(g.cpu().numpy().astype(np.float64) - w.grad.numpy()).max()
        # Dense_s (one 1-column kernel per step) is a module of its own, so its floor is a single sample of the
        # reference-precision rounding noise instead of a maximum over many tensors: 10 x that sample here
        # ... and modules whose gradients are orders of magnitude below the dominant flow of the same backward pass
        # (here: the position path of a read-out that pools O(1) scalars) inherit that flow's fp32 rounding: 1e-7 of it
        assert err <= RTOL * scales[path[0]] + 10.0 * floors[path[0]] + 1e-7 * gmax, \
            f"grad {'/'.join(path)}: err {err:.3e}, scale {scales[path[0]]:.3e}, fp32 floor {10 * floors[path[0]]:.3e}"


def test_egnn_xvh_equivariance_property():
    """The reference's own EGNN test (tests/test_equivariance.py:224-250): 7 features, soft edges, tanh,
    decoupled position/velocity updates, node task; positions and velocities rotate, the scalar does not."""
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.egnn import EGNN
    rng = np.random.default_rng(4)
    x = rng.normal(size=(2, 64, 7)).astype(np.float32)
    R = G.rotation_matrix(45.0, np.array([0, 1 / np.sqrt(2), 1 / np.sqrt(2)]))
    xr = x.copy()
    xr[..., :3] = x[..., :3] @ R.T
    xr[..., 3:6] = x[..., 3:6] @ R.T
    model = EGNN(message_passing_steps=3, d_hidden=32, n_layers=3, activation="gelu", tanh_out=True, soft_edges=True,
                 positions_only=False, task="node", decouple_pos_vel_updates=True)

    def graph(a):
        a = torch.tensor(a).cuda()
        return GU.build_graph(a, None, 5)._replace(edges=None)
    g0, g1 = graph(x), graph(xr)
    params = model.init(0, g0)
    o0 = model.forward_backward(params, g0).cpu().numpy()
    o1 = model.forward_backward(params, g1).cpu().numpy()
    want = o0.copy()
    want[..., :3] = o0[..., :3] @ R.T
    want[..., 3:6] = o0[..., 3:6] @ R.T
    assert not np.allclose(o0[..., :3], o1[..., :3], rtol=0.05)
    assert np.allclose(o1, want, rtol=0.05, atol=1e-4)
    _close(o1, want, "equivariance", rtol=1e-4)


def test_egnn_equivariance_property():
    """tests/test_equivariance.py:224-250 style check on the B200 path: f(Rx) = R f(x)."""
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.egnn import EGNN
    rng = np.random.default_rng(3)
    x = rng.normal(size=(2, 64, 3)).astype(np.float32)
    R = G.rotation_matrix(45.0, np.array([0, 1 / np.sqrt(2), 1 / np.sqrt(2)]))
    xr = (x @ R.T).astype(np.float32)
    model = EGNN(positions_only=True, soft_edges=False, tanh_out=True, d_hidden=128, n_layers=3, n_radial_basis=32,
                 r_max=1.0, message_passing_steps=3, task="node")

    def graph(a):
        a = torch.tensor(a).cuda()
        g = GU.build_graph(a, None, 6)
        return g._replace(edges=None)
    g0, g1 = graph(x), graph(xr)
    params = model.init(0, g0)
    o0 = model.forward_backward(params, g0).cpu().numpy()
    o1 = model.forward_backward(params, g1).cpu().numpy()
    assert not np.allclose(o0, o1, rtol=0.05)
    assert np.allclose(o1, o0 @ R.T, rtol=0.05)
    _close(o1, o0 @ R.T, "equivariance", rtol=1e-4)


def test_egnn_unsupported_variants_fail_loudly():
    from eqnn_jax_b200.egnn import EGNN
    with pytest.raises(NotImplementedError):
        EGNN(positions_only=True, activation="relu").init(0, type("G", (), {"nodes": torch.zeros(1, 4, 3).cuda()})())
    with pytest.raises(ValueError):
        EGNN(positions_only=True, message_passing_agg="median").init(0, type("G", (), {"nodes": torch.zeros(1, 4, 3).cuda()})())
