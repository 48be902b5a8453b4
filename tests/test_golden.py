"""Golden vectors (oracle-generated, tests/golden/make_golden.py): oracle regression on CPU, product on GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import graph_ref as G
from oracle import nequip_ref as NQ

HERE = os.path.dirname(os.path.abspath(__file__))
IRR = "1o+1o+1x0e"
CFGS = {
    "graph_l2": dict(d_hidden=32, l_max=2, sphharm_norm="integral", message_passing_steps=2, n_layers=2, n_outputs=2,
                     n_radial_basis=8, r_cutoff=0.6, task="graph"),
    "node_l1": dict(d_hidden=32, l_max=1, message_passing_steps=2, n_layers=3, n_radial_basis=4, task="node",
                    irreps_out=IRR),
}


@pytest.fixture(scope="module")
def knn():
    return np.load(os.path.join(HERE, "golden", "knn_pbc.npz"))


@pytest.fixture(scope="module")
def neq():
    return np.load(os.path.join(HERE, "golden", "nequip_small.npz"))


def test_oracle_reproduces_knn_golden(knn):
    pbc = G.get_apply_pbc(knn["std_over_box"])
    for k in (8, 20):
        for b in range(2):
            s, t, dr = G.nearest_neighbors(knn["x"][b], k, pbc)
            np.testing.assert_array_equal(t, knn[f"receivers_k{k}"][b])
            assert np.array_equal(dr.view(np.uint32), knn[f"dr_k{k}"][b].view(np.uint32))
    np.testing.assert_array_equal(G.nearest_neighbors(knn["lattice"], 7)[1], knn["lattice_receivers_k7"])


@pytest.mark.parametrize("name", list(CFGS))
def test_oracle_reproduces_nequip_golden(neq, name):
    cfg = NQ.NequIPConfig(**CFGS[name])
    tree = NQ.init_params(cfg, IRR, seed=11)
    g = G.build_graph(neq["x"], None, int(neq["k"]))
    o = NQ.apply_batched(NQ.to_torch(tree), cfg, g._replace(nodes=torch.tensor(neq["x"], dtype=torch.float64)), IRR)
    val = (o.nodes if cfg.task == "node" else o).numpy()
    np.testing.assert_allclose(val, neq[f"{name}_out"], rtol=1e-10, atol=1e-12)


@pytest.mark.gpu
def test_gpu_knn_matches_golden(knn):
    from eqnn_jax_b200 import graph_utils as GU, ops
    pbc = GU.get_apply_pbc(knn["std_over_box"])
    x = torch.tensor(knn["x"]).cuda()
    for k in (8, 20):
        s, t, dr = ops.knn_pbc(x, k, pbc.cell, pbc.cell_inv)
        np.testing.assert_array_equal(t.cpu().numpy(), knn[f"receivers_k{k}"])
        assert np.array_equal(dr.cpu().numpy().view(np.uint32), knn[f"dr_k{k}"].view(np.uint32))
    s, t, dr = GU.nearest_neighbors(torch.tensor(knn["lattice"]).cuda(), 7)
    np.testing.assert_array_equal(t.cpu().numpy(), knn["lattice_receivers_k7"])


@pytest.mark.gpu
@pytest.mark.parametrize("name", list(CFGS))
def test_gpu_nequip_matches_golden(neq, name):
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.nequip import IrrepsArray, NequIP
    kw = CFGS[name]
    x = torch.tensor(neq["x"]).cuda()
    g = GU.build_graph(x, None, int(neq["k"]))
    gg = g._replace(nodes=IrrepsArray(IRR, x), edges=None)
    model = NequIP(**kw)
    params = model.init(0, gg)
    params.load_tree(NQ.init_params(NQ.NequIPConfig(**kw), IRR, seed=11))
    cot = torch.tensor(neq[f"{name}_cot"], dtype=torch.float32).cuda()
    out = model.forward_backward(params, gg, lambda o: cot)
    want = neq[f"{name}_out"]
    got = out.cpu().numpy().astype(np.float64)
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()
    gt = params.grad_tree
    w = neq[f"{name}_grad_mlp0_dense0"]
    got = gt["MultiLayerPerceptron_0"]["Dense_0"]["kernel"].cpu().numpy()
    assert np.abs(got - w).max() <= 1e-5 * np.abs(w).max()
    w = neq[f"{name}_grad_linear1"]
    got = np.concatenate([v.cpu().numpy().reshape(-1) for v in gt["Linear_1"].values()])
    assert np.abs(got - w).max() <= 1e-5 * np.abs(w).max()


# ---- EGNN (x, v, h) and SEGNN: the configurations of the reference's own equivariance tests --------------------
EGNN_KW = dict(message_passing_steps=2, d_hidden=32, n_layers=3, tanh_out=True, soft_edges=True, positions_only=False,
               task="node", decouple_pos_vel_updates=True, n_radial_basis=8, r_max=2.0)
SEGNN_KW = dict(message_passing_steps=2, d_hidden=32, task="node", residual=True, output_irreps="1o + 1o + 1x0e")


@pytest.fixture(scope="module")
def es():
    return np.load(os.path.join(HERE, "golden", "egnn_segnn_small.npz"))


def test_oracle_reproduces_egnn_segnn_golden(es):
    from oracle import egnn_ref as EG, segnn_ref as SG
    from oracle.e3nn_ref import IrrepsArray as OIA
    from oracle.so3 import Irreps as OI
    x = es["x"]
    g = G.build_graph(x, None, int(es["k"]))
    cfg = EG.EGNNConfig(**EGNN_KW)
    o = EG.apply_batched(NQ.to_torch(EG.init_params(cfg, 7, seed=5)), cfg, g._replace(nodes=torch.tensor(x, dtype=torch.float64)))
    np.testing.assert_allclose(o.nodes.numpy(), es["egnn_out"], rtol=1e-9, atol=1e-11)
    st = G.get_equivariant_graph(OIA(IRR, torch.tensor(x, dtype=torch.float64)), x[..., :3], x[..., 3:6], g.senders,
                                 g.receivers, g.n_node, g.n_edge, None, None, 1)
    np.testing.assert_allclose(np.asarray(st.steerable_node_attrs), es["steerable_node_attrs"], rtol=1e-10, atol=1e-12)
    scfg = SG.SEGNNConfig(**SEGNN_KW)
    so = SG.apply_batched(NQ.to_torch(SG.init_params(scfg, IRR, OI.spherical_harmonics(1), "1x0e", seed=6)), scfg, st)
    np.testing.assert_allclose(so.array.numpy(), es["segnn_out"], rtol=1e-9, atol=1e-11)


@pytest.mark.gpu
def test_gpu_egnn_segnn_match_golden(es):
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.egnn import EGNN
    from eqnn_jax_b200.equivariant_graph_utils import get_equivariant_graph
    from eqnn_jax_b200.irreps import Irreps
    from eqnn_jax_b200.nequip import IrrepsArray
    from eqnn_jax_b200.segnn import SEGNN
    from oracle import egnn_ref as EG, segnn_ref as SG
    from oracle.so3 import Irreps as OI
    x = torch.tensor(es["x"]).cuda()
    g = GU.build_graph(x, None, int(es["k"]))
    model = EGNN(**EGNN_KW)
    gg = g._replace(edges=None)
    params = model.init(0, gg)
    params.load_tree(EG.init_params(EG.EGNNConfig(**EGNN_KW), 7, seed=5))
    got = model.forward_backward(params, gg).cpu().numpy().astype(np.float64)
    assert np.abs(got - es["egnn_out"]).max() <= 1e-5 * np.abs(es["egnn_out"]).max()
    st = get_equivariant_graph(IrrepsArray(Irreps(IRR), x), x[..., :3].contiguous(), x[..., 3:6].contiguous(), g.senders,
                               g.receivers, g.n_node, g.n_edge, None, None, 1)
    a = st.steerable_node_attrs.array.cpu().numpy().astype(np.float64)
    assert np.abs(a - es["steerable_node_attrs"]).max() <= 1e-5 * np.abs(es["steerable_node_attrs"]).max()
    smodel = SEGNN(**SEGNN_KW)
    sparams = smodel.init(0, st)
    sparams.load_tree(SG.init_params(SG.SEGNNConfig(**SEGNN_KW), IRR, OI.spherical_harmonics(1), "1x0e", seed=6))
    sgot = smodel.forward_backward(sparams, st).cpu().numpy().astype(np.float64)
    assert np.abs(sgot - es["segnn_out"]).max() <= 1e-5 * np.abs(es["segnn_out"]).max()
