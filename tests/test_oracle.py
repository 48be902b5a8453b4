"""Pins for the CPU oracle (no GPU): closed forms, independent CG, equivariance, param counts,
and the reference's own PBC sanity tests (tests/test_pbcs.py) re-run on the restatement."""
import dataclasses

import numpy as np
import pytest
import torch

from oracle import e3nn_ref as E
from oracle import graph_ref as G
from oracle import nequip_ref as NQ
from oracle import so3

IRREPS_IN = "1o+1o+1x0e"


def test_cg_closed_forms():
    # C(1,1,0) = delta/sqrt3, C(1,1,1) = +eps/sqrt6
    np.testing.assert_allclose(so3.clebsch_gordan(1, 1, 0)[:, :, 0], np.eye(3) / np.sqrt(3), atol=1e-14)
    c = so3.clebsch_gordan(1, 1, 1) * np.sqrt(6)
    eps = np.zeros((3, 3, 3))
    for i, j, k in [(0, 1, 2), (1, 2, 0), (2, 0, 1)]:
        eps[i, j, k], eps[j, i, k] = 1, -1
    np.testing.assert_allclose(c, eps, atol=1e-14)
    for l1, l2, l3 in [(1, 2, 3), (2, 2, 2), (1, 3, 4), (3, 2, 1)]:
        assert abs(np.sum(so3.clebsch_gordan(l1, l2, l3) ** 2) - 1) < 1e-13


def test_product_cg_matches_sympy_oracle():
    from eqnn_jax_b200 import so3 as p
    for l1 in range(4):
        for l2 in range(5):
            for l3 in range(abs(l1 - l2), min(l1 + l2, 4) + 1):
                np.testing.assert_allclose(p.real_cg(l1, l2, l3), so3.clebsch_gordan(l1, l2, l3), atol=1e-13)
    for l in (2, 3, 4):
        assert abs(p.sh_recursion_constant(l) - so3._sh_recursion_constant(l)) < 1e-12


def test_sh_closed_forms_and_scipy():
    rng = np.random.default_rng(0)
    v = rng.normal(size=(50, 3))
    u = v / np.linalg.norm(v, axis=-1, keepdims=True)
    x, y, z = u.T
    Y = so3.spherical_harmonics(3, v, True, "norm")
    np.testing.assert_allclose(Y[:, 0], 1.0)
    np.testing.assert_allclose(Y[:, 1:4], u, atol=1e-14)
    s3 = np.sqrt(3)
    l2 = np.stack([s3 * x * z, s3 * x * y, y * y - 0.5 * (x * x + z * z), s3 * y * z, s3 / 2 * (z * z - x * x)], -1)
    np.testing.assert_allclose(Y[:, 4:9], l2, atol=1e-13)
    # l = 3 against scipy's complex harmonics: e3nn axes (x,y,z) = standard (y,z,x)
    from scipy.special import sph_harm_y
    xs, ys, zs = z, x, y                      # standard coordinates of the same points
    theta, phi = np.arccos(np.clip(zs, -1, 1)), np.arctan2(ys, xs)
    for l in (1, 2, 3):
        cols = []
        for m in range(-l, l + 1):
            Yc = sph_harm_y(l, abs(m), theta, phi)
            if m < 0:
                r = np.sqrt(2) * (-1) ** m * Yc.imag
            elif m == 0:
                r = Yc.real
            else:
                r = np.sqrt(2) * (-1) ** m * Yc.real
            cols.append(r * np.sqrt(4 * np.pi / (2 * l + 1)))
        np.testing.assert_allclose(Y[:, l * l:(l + 1) ** 2], np.stack(cols, -1), atol=1e-12)
    # zero-vector guard and normalisations
    Y0 = so3.spherical_harmonics(2, np.zeros((1, 3)), True, "integral")[0]
    assert Y0[0] == pytest.approx(np.sqrt(1 / (4 * np.pi))) and np.all(Y0[1:] == 0)
    Yc = so3.spherical_harmonics(2, v, True, "component")
    np.testing.assert_allclose(Yc[:, 4:9], np.sqrt(5) * Y[:, 4:9], atol=1e-13)


def test_tensor_product_equivariance():
    rng = np.random.default_rng(1)
    R = G.rotation_matrix(37.0, np.array([1.0, -2.0, 0.5]))
    for Rm in (R, -R):
        irx, iry = so3.Irreps("1x0e+2x1o+1x2e"), so3.Irreps.spherical_harmonics(2)
        x = torch.tensor(rng.normal(size=(4, irx.dim)))
        y = torch.tensor(rng.normal(size=(4, iry.dim)))
        Dx, Dy = so3.irreps_rotation_matrix(irx, Rm), so3.irreps_rotation_matrix(iry, Rm)
        out = E.tensor_product(E.IrrepsArray(irx, x), E.IrrepsArray(iry, y))
        out_r = E.tensor_product(E.IrrepsArray(irx, x @ torch.tensor(Dx.T)), E.IrrepsArray(iry, y @ torch.tensor(Dy.T)))
        Do = so3.irreps_rotation_matrix(out.irreps, Rm)
        np.testing.assert_allclose(out_r.array.numpy(), out.array.numpy() @ Do.T, atol=1e-12)


def test_message_irreps_table():
    # SURVEY.md section 3.2
    for lmax, want, n_irr in [(1, "3x0e+3x1o+2x1e+2x2e", 10),
                              (2, "3x0e+5x1o+2x1e+3x2e+2x2o+2x3o", 17),
                              (3, "3x0e+5x1o+2x1e+5x2e+2x2o+3x3o+2x3e+2x4e", 24)]:
        m = NQ._message_irreps(so3.Irreps(IRREPS_IN), lmax)
        assert m == so3.Irreps(want) and m.num_irreps == n_irr
    assert so3.balanced_irreps(2, 128) == "88x0e+14x1o+8x2e"
    assert so3.balanced_irreps(3, 128) == "100x0e+10x1o+6x2e+4x3o"


def test_param_counts():
    # hand count of the current models/nequip.py (the notebook's 391,622 predates it)
    cfg = NQ.NequIPConfig(n_outputs=2, n_radial_basis=64, r_cutoff=0.6)
    assert NQ.count_params(NQ.init_params(cfg, "1o")) == 388_676
    cfg2 = NQ.NequIPConfig(d_hidden=128, l_max=2, sphharm_norm="integral", message_passing_steps=4, n_layers=3,
                           n_outputs=2, n_radial_basis=64, r_cutoff=0.6)
    assert NQ.count_params(NQ.init_params(cfg2, IRREPS_IN)) == 439_342
    assert NQ.count_params(NQ.init_params(dataclasses.replace(cfg2, l_max=3), IRREPS_IN)) == 443_062


def test_egnn_param_count_matches_notebook():
    # notebooks/examples.ipynb cells 14-15: EGNN(positions_only=True, n_outputs=2, n_layers=4, 64 Bessel, tanh_out) -> 441778
    from oracle import egnn_ref as EG
    cfg = EG.EGNNConfig(positions_only=True, n_outputs=2, n_layers=4, n_radial_basis=64, r_max=0.6, tanh_out=True)
    assert NQ.count_params(EG.init_params(cfg, 3)) == 441_778
    from eqnn_jax_b200.egnn import EGNN
    m = EGNN(positions_only=True, n_outputs=2, n_layers=4, n_radial_basis=64, r_max=0.6, tanh_out=True)
    assert m._layout()[3] == 441_778
    tree = EG.init_params(EG.EGNNConfig(positions_only=True, soft_edges=False, n_outputs=2), 3)
    m2 = EGNN(positions_only=True, soft_edges=False, n_outputs=2)
    flat = {}

    def walk(d, pre=()):
        for k, v in d.items():
            if isinstance(v, dict):
                walk(v, pre + (k,))
            else:
                flat[pre + (k,)] = tuple(np.shape(v))
    walk(tree)
    assert {p: s for p, s, _ in m2._layout()[0]} == flat


def test_normalize_constants():
    # SURVEY.md A.6 (grid values)
    assert E.normalization_constant("gelu") == pytest.approx(0.652059099, abs=2e-8)
    assert E.normalization_constant("sigmoid") == pytest.approx(0.541644565, abs=2e-8)
    from eqnn_jax_b200.nequip import normalize_constant
    assert normalize_constant("gelu") == pytest.approx(E.normalization_constant("gelu"), rel=1e-10)
    assert normalize_constant("sigmoid") == pytest.approx(E.normalization_constant("sigmoid"), rel=1e-10)


def _graph(x, k=5):
    g = G.build_graph(x, None, k)
    return g._replace(nodes=torch.tensor(g.nodes, dtype=torch.float64))


@pytest.mark.parametrize("task", ["node", "graph"])
def test_reference_equivariance_tests_on_oracle(task):
    """tests/test_equivariance.py:337-388 (NequIP node / graph), rtol 0.05, same rotation."""
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 5, 7)).astype(np.float32)
    axis = np.array([0, 1 / np.sqrt(2), 1 / np.sqrt(2)])
    xr = np.stack([G.rotate_representation(x[b], 45.0, axis) for b in range(2)]).astype(np.float32)
    cfg = NQ.NequIPConfig(message_passing_steps=2, d_hidden=32, n_layers=3, task=task,
                          irreps_out="1o + 1o + 1x0e" if task == "node" else "1x0e")
    p = NQ.to_torch(NQ.init_params(cfg, IRREPS_IN, seed=0))
    o1 = NQ.apply_batched(p, cfg, _graph(x), IRREPS_IN)
    o2 = NQ.apply_batched(p, cfg, _graph(xr), IRREPS_IN)
    if task == "node":
        a, b = o1.nodes.numpy(), o2.nodes.numpy()
        ar = np.stack([G.rotate_representation(a[i], 45.0, axis) for i in range(2)])
        assert not np.allclose(a, b, rtol=0.05)
        assert np.allclose(b, ar, rtol=0.05)
        np.testing.assert_allclose(b, ar, atol=2e-5 * np.abs(a).max())
    else:
        np.testing.assert_allclose(o1.numpy(), o2.numpy(), rtol=1e-4)


@pytest.mark.parametrize("F,kw", [(3, dict(positions_only=True)), (5, dict(positions_only=True)),
                                  (6, dict(positions_only=False)),
                                  (7, dict(positions_only=False, decouple_pos_vel_updates=True))])
def test_reference_egnn_equivariance_test_on_oracle(F, kw):
    """tests/test_equivariance.py:224-250 (EGNN, 7 features, soft edges, tanh, decoupled updates) on the oracle,
    and the other three node layouts of models/egnn.py:38-60: vectors rotate, scalars are invariant."""
    from oracle import egnn_ref as EG
    rng = np.random.default_rng(1)
    x = rng.normal(size=(2, 5, F)).astype(np.float32)
    axis = np.array([0, 1 / np.sqrt(2), 1 / np.sqrt(2)])
    R = G.rotation_matrix(45.0, axis)
    n_vec = 1 if kw["positions_only"] else 2

    def rot(a):
        a = np.array(a, dtype=np.float64)
        for v in range(n_vec):
            a[..., 3 * v:3 * v + 3] = a[..., 3 * v:3 * v + 3] @ R.T
        return a
    cfg = EG.EGNNConfig(message_passing_steps=3, d_hidden=32, n_layers=3, tanh_out=True, soft_edges=True, task="node", **kw)
    p = NQ.to_torch(EG.init_params(cfg, F, seed=0))
    g0 = G.build_graph(x, None, 5)
    g1 = G.build_graph(rot(x).astype(np.float32), None, 5)
    assert np.array_equal(g0.senders, g1.senders)
    o0 = EG.apply_batched(p, cfg, g0._replace(nodes=torch.tensor(x, dtype=torch.float64))).nodes.numpy()
    o1 = EG.apply_batched(p, cfg, g1._replace(nodes=torch.tensor(rot(x)))).nodes.numpy()
    assert not np.allclose(o0[..., :3], o1[..., :3], rtol=0.05)
    np.testing.assert_allclose(o1, rot(o0), rtol=0.05, atol=1e-6)
    # exact up to the reference's `+ 1e-7` on every component of r_ij (egnn.py:82-83), which is not a vector
    np.testing.assert_allclose(o1, rot(o0), atol=2e-5 * max(1.0, np.abs(o0).max()))


def test_steerable_attributes_on_oracle():
    """equivariant_graph_utils.py:27-102 restated in oracle/graph_ref.py: the attributes are spherical tensors
    (Y(R r) = D(R) Y(r)), the node attributes are receiver-side means, self-loop edges (r = 0) give (Y_0, 0, ...),
    and the radial messages are rotation invariant."""
    from oracle import so3
    rng = np.random.default_rng(2)
    x = rng.normal(size=(2, 20, 3)).astype(np.float32)
    v = rng.normal(size=(2, 20, 3)).astype(np.float32)
    R = G.rotation_matrix(33.0, np.array([0.3, -0.5, 0.81]) / np.linalg.norm([0.3, -0.5, 0.81]))
    g = G.build_graph(x, None, 4)
    kw = dict(steerable_velocities=True, spherical_harmonics_norm="integral", n_radial_basis=6, r_max=2.0)
    a = G.get_equivariant_graph(None, x, v, g.senders, g.receivers, g.n_node, g.n_edge, None, None, 2, **kw)
    b = G.get_equivariant_graph(None, (x @ R.T).astype(np.float32), (v @ R.T).astype(np.float32), g.senders, g.receivers,
                                g.n_node, g.n_edge, None, None, 2, **kw)
    D = so3.irreps_rotation_matrix(so3.Irreps.spherical_harmonics(2), R)
    np.testing.assert_allclose(b.steerable_edge_attrs, a.steerable_edge_attrs @ D.T, atol=2e-6)
    np.testing.assert_allclose(b.steerable_node_attrs, a.steerable_node_attrs @ D.T, atol=2e-6)
    np.testing.assert_allclose(b.additional_messages, a.additional_messages, atol=2e-5)
    self_loops = g.senders == g.receivers
    assert self_loops.any()
    y0 = 1.0 / np.sqrt(4 * np.pi)
    np.testing.assert_allclose(a.steerable_edge_attrs[self_loops], np.array([y0] + [0.0] * 8)[None].repeat(self_loops.sum(), 0))
    # node attribute of node 0 of graph 0 = mean over its incoming edges + Y(v)
    inc = g.receivers[0] == 0
    want = a.steerable_edge_attrs[0][inc].mean(0) + so3.spherical_harmonics(2, v[0, 0].astype(np.float64), True, "integral")
    np.testing.assert_allclose(a.steerable_node_attrs[0, 0], want, atol=1e-12)


def _steerable(x, k=5):
    from oracle.e3nn_ref import IrrepsArray
    g = G.build_graph(x, None, k)
    return G.get_equivariant_graph(IrrepsArray("1o+1o+1x0e", torch.tensor(x, dtype=torch.float64)), x[..., :3], x[..., 3:6],
                                   g.senders, g.receivers, g.n_node, g.n_edge, None, None, 1)


def test_segnn_param_count_matches_notebook():
    """notebooks/examples.ipynb cells 8-9: 573 253 parameters (d_hidden 128, l_max 1, 3 x 3 layers, 64 radial)."""
    from oracle import segnn_ref as S
    from oracle.so3 import Irreps
    cfg = S.SEGNNConfig(d_hidden=128, l_max_hidden=1, n_layers=3, message_passing_steps=3, task="graph",
                        output_irreps="1x0e", message_passing_agg="mean", n_outputs=2, residual=False)
    p = S.init_params(cfg, "1o", Irreps.spherical_harmonics(1), "64x0e")

    def count(t):
        return sum(count(v) if isinstance(v, dict) else int(np.prod(v.shape)) for v in t.values())
    assert count(p) == 573253


@pytest.mark.parametrize("task", ["node", "graph"])
def test_reference_segnn_equivariance_tests_on_oracle(task):
    """tests/test_equivariance.py:288-335 (SEGNN node: f(Rx) = R f(x); SEGNN graph: f(Rx) = f(x)), rtol 0.05."""
    from oracle import segnn_ref as S
    from oracle.so3 import Irreps
    rng = np.random.default_rng(0)
    x = rng.normal(size=(2, 5, 7)).astype(np.float32)
    axis = np.array([0, 1 / np.sqrt(2), 1 / np.sqrt(2)])
    xr = np.stack([G.rotate_representation(x[b], 45.0, axis) for b in range(2)]).astype(np.float32)
    cfg = S.SEGNNConfig(message_passing_steps=2 if task == "node" else 3, d_hidden=32, task=task, residual=True,
                        output_irreps="1o + 1o + 1x0e" if task == "node" else "1x0e")
    p = NQ.to_torch(S.init_params(cfg, "1o+1o+1x0e", Irreps.spherical_harmonics(1), "1x0e", seed=0))
    o1 = S.apply_batched(p, cfg, _steerable(x))
    o2 = S.apply_batched(p, cfg, _steerable(xr))
    if task == "node":
        a, b = o1.array.numpy(), o2.array.numpy()
        ar = np.stack([G.rotate_representation(a[i], 45.0, axis) for i in range(2)])
        assert not np.allclose(a, b, rtol=0.05)
        assert np.allclose(b, ar, rtol=0.05, atol=1e-6)
        np.testing.assert_allclose(b, ar, atol=2e-5 * np.abs(a).max())
    else:
        np.testing.assert_allclose(o1.numpy(), o2.numpy(), rtol=1e-4, atol=1e-7)


def test_segnn_discarded_node_layers_have_zero_gradient():
    """segnn.py:84-89: the first n_layers-1 TPLGs of the node function own parameters but never reach the output."""
    from oracle import segnn_ref as S
    from oracle.so3 import Irreps
    rng = np.random.default_rng(1)
    x = rng.normal(size=(1, 6, 7)).astype(np.float32)
    cfg = S.SEGNNConfig(message_passing_steps=1, d_hidden=16, n_layers=3, task="graph", output_irreps="1x0e")
    p = NQ.to_torch(S.init_params(cfg, "1o+1o+1x0e", Irreps.spherical_harmonics(1), "1x0e", seed=0), requires_grad=True)
    S.apply_batched(p, cfg, _steerable(x, 4)).sum().backward()
    # creation order: embed (0), edge fn (1, 2, 3), node fn (4, 5 discarded; 6 used), read-out (7)
    for name, dead in [("TensorProductLinearGate_4", True), ("TensorProductLinearGate_5", True),
                       ("TensorProductLinearGate_6", False), ("TensorProductLinearGate_3", False)]:
        g = max(float(w.grad.abs().max()) if w.grad is not None else 0.0 for w in p[name]["Linear_0"].values())
        assert (g == 0.0) == dead, (name, g)


def test_pbc_sanity_reference_tests():
    """tests/test_pbcs.py:8-33 and :35-62."""
    rng = np.random.default_rng(0)
    pos = rng.uniform(size=(1, 10, 3)).astype(np.float32)
    g0 = G.build_graph(pos, None, k=10, apply_pbc=None)
    g1 = G.build_graph(pos, None, k=10, apply_pbc=G.get_apply_pbc())
    assert g0.edges.mean() > g1.edges.mean()
    assert g0.edges.max() > np.sqrt(3) / 2 and g1.edges.max() < np.sqrt(3) / 2
    pos = rng.uniform(size=(1, 100, 3)).astype(np.float32)
    std = pos.std(axis=(0, 1))
    spos = (pos - pos.mean(axis=(0, 1))) / std
    gs = G.build_graph(spos, None, k=10, apply_pbc=G.get_apply_pbc(std))
    g = G.build_graph(pos, None, k=10, apply_pbc=G.get_apply_pbc())
    assert gs.edges.mean() * std[0] == pytest.approx(g.edges.mean(), rel=0.05)


def test_knn_self_is_first_and_ties_are_stable():
    x = np.zeros((6, 3), dtype=np.float32)
    x[3:] = 1.0                      # two clusters of coincident points: all-tie rows
    s, t, dr = G.nearest_neighbors(x, 3)
    assert t.reshape(6, 3).tolist() == [[0, 1, 2]] * 3 + [[3, 4, 5]] * 3
    assert s.tolist() == np.repeat(np.arange(6), 3).tolist()


def test_radius_graph_restatement_properties():
    """oracle.graph_ref.radius_graph (graph_utils.py:266-277, the branch that cannot run upstream): symmetric pairs, no
    self loops, counts equal to a brute-force count in float64 away from the cut-off, np.where ordering."""
    from oracle import graph_ref as G
    rng = np.random.default_rng(1)
    x = rng.uniform(0, 1, size=(2, 60, 3)).astype(np.float32)
    cell = np.diag([1.0, 1.0, 1.0]).astype(np.float32)
    s, t, d, n = G.radius_graph(x, 0.3, G.get_apply_pbc(cell=cell))
    assert n.sum() == len(s) == len(t) == len(d)
    off = 0
    for b in range(2):
        sb, tb = s[off:off + n[b]], t[off:off + n[b]]
        pairs = set(zip(sb.tolist(), tb.tolist()))
        assert all((j, i) in pairs for i, j in pairs) and all(i != j for i, j in pairs)
        assert np.all(np.diff(sb.astype(np.int64) * 60 + tb) > 0)           # np.where order: row-major, strictly increasing
        dr = x[b, :, None, :].astype(np.float64) - x[b, None, :, :]
        dr -= np.round(dr)
        d64 = np.sqrt((dr ** 2).sum(-1))
        sure = (np.abs(d64 - 0.3) > 1e-5)
        want = (d64 < 0.3) & (d64 > 0)
        got = np.zeros_like(want)
        got[sb, tb] = True
        assert np.array_equal(got[sure], want[sure])
        off += n[b]
    assert np.all((d > 0) & (d < 0.3))
