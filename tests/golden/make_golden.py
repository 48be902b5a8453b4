#!/usr/bin/env python
"""Generates tests/golden/*.npz.

These vectors are produced by the CPU ORACLE (oracle/), not by the reference: the reference's
jax / e3nn-jax stack cannot be installed in this image, and its own tests carry no
golden vectors.  They pin the oracle against regressions and give the GPU tests fixed inputs/outputs that
do not depend on the oracle being importable at the same version.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))

from oracle import graph_ref as G          # noqa: E402
from oracle import nequip_ref as NQ        # noqa: E402
from oracle import egnn_ref as EG          # noqa: E402
from oracle import segnn_ref as SG         # noqa: E402
from oracle.e3nn_ref import IrrepsArray as OIA   # noqa: E402
from oracle.so3 import Irreps as OI        # noqa: E402

EGNN_KW = dict(message_passing_steps=2, d_hidden=32, n_layers=3, tanh_out=True, soft_edges=True, positions_only=False,
               task="node", decouple_pos_vel_updates=True, n_radial_basis=8, r_max=2.0)
SEGNN_KW = dict(message_passing_steps=2, d_hidden=32, task="node", residual=True, output_irreps="1o + 1o + 1x0e")

IRR = "1o+1o+1x0e"


def knn_fixture():
    rng = np.random.default_rng(123)
    pos = rng.uniform(0, 1000, size=(2, 96, 3))
    std = np.array([288.71092533309235, 288.7525818573022, 288.70234893905575])
    x = ((pos - 500.0) / std).astype(np.float32)
    pbc = G.get_apply_pbc((std / 1000.0).astype(np.float32))
    out = {"x": x, "std_over_box": (std / 1000.0).astype(np.float32)}
    for k in (8, 20):
        res = [G.nearest_neighbors(x[b], k, pbc) for b in range(2)]
        out[f"receivers_k{k}"] = np.stack([r[1] for r in res])
        out[f"dr_k{k}"] = np.stack([r[2] for r in res])
    # a lattice with exact ties, no PBC
    g = np.stack(np.meshgrid(*[np.arange(4)] * 3, indexing="ij"), -1).reshape(-1, 3).astype(np.float32)
    out["lattice"] = g
    out["lattice_receivers_k7"] = G.nearest_neighbors(g, 7)[1]
    np.savez_compressed(os.path.join(HERE, "knn_pbc.npz"), **out)


def nequip_fixture():
    rng = np.random.default_rng(7)
    B, N, k = 2, 24, 6
    x = rng.normal(size=(B, N, 7)).astype(np.float32)
    out = {"x": x, "k": np.array(k)}
    for name, kw in {
        "graph_l2": dict(d_hidden=32, l_max=2, sphharm_norm="integral", message_passing_steps=2, n_layers=2,
                         n_outputs=2, n_radial_basis=8, r_cutoff=0.6, task="graph"),
        "node_l1": dict(d_hidden=32, l_max=1, message_passing_steps=2, n_layers=3, n_radial_basis=4, task="node",
                        irreps_out=IRR),
    }.items():
        cfg = NQ.NequIPConfig(**kw)
        tree = NQ.init_params(cfg, IRR, seed=11)
        g = G.build_graph(x, None, k)
        p = NQ.to_torch(tree, torch.float64, requires_grad=True)
        o = NQ.apply_batched(p, cfg, g._replace(nodes=torch.tensor(x, dtype=torch.float64)), IRR)
        val = o.nodes if kw["task"] == "node" else o
        cot = np.random.default_rng(5).normal(size=tuple(val.shape))
        (val * torch.tensor(cot)).sum().backward()
        out[f"{name}_out"] = val.detach().numpy()
        out[f"{name}_cot"] = cot
        out[f"{name}_grad_mlp0_dense0"] = p["MultiLayerPerceptron_0"]["Dense_0"]["kernel"].grad.numpy()
        out[f"{name}_grad_linear1"] = np.concatenate([v.grad.numpy().reshape(-1) for v in p["Linear_1"].values()])
    np.savez_compressed(os.path.join(HERE, "nequip_small.npz"), **out)


def egnn_segnn_fixture():
    """EGNN on (x, v, h) nodes and SEGNN on "1o+1o+1x0e" nodes (the configurations of the reference's own
    equivariance tests, tests/test_equivariance.py:224-310), oracle outputs in fp64."""
    rng = np.random.default_rng(21)
    B, N, k = 2, 20, 5
    x = rng.normal(size=(B, N, 7)).astype(np.float32)
    g = G.build_graph(x, None, k)
    out = {"x": x, "k": np.array(k)}
    cfg = EG.EGNNConfig(**EGNN_KW)
    p = NQ.to_torch(EG.init_params(cfg, 7, seed=5))
    out["egnn_out"] = EG.apply_batched(p, cfg, g._replace(nodes=torch.tensor(x, dtype=torch.float64))).nodes.numpy()
    st = G.get_equivariant_graph(OIA(IRR, torch.tensor(x, dtype=torch.float64)), x[..., :3], x[..., 3:6], g.senders,
                                 g.receivers, g.n_node, g.n_edge, None, None, 1)
    scfg = SG.SEGNNConfig(**SEGNN_KW)
    sp = NQ.to_torch(SG.init_params(scfg, IRR, OI.spherical_harmonics(1), "1x0e", seed=6))
    out["segnn_out"] = SG.apply_batched(sp, scfg, st).array.numpy()
    out["steerable_node_attrs"] = np.asarray(st.steerable_node_attrs)
    out["steerable_edge_attrs"] = np.asarray(st.steerable_edge_attrs)
    np.savez_compressed(os.path.join(HERE, "egnn_segnn_small.npz"), **out)


if __name__ == "__main__":
    knn_fixture()
    nequip_fixture()
    egnn_segnn_fixture()
    print("wrote", [f for f in os.listdir(HERE) if f.endswith(".npz")])
