#!/usr/bin/env python
"""Pins the oracle (and through it the CUDA path) against the REFERENCE ITSELF.

Run on any machine that has the reference checkout and its stack (jax, flax, jraph, e3nn-jax):

    python tests/golden/make_reference_golden.py --reference /path/to/eqnn-jax

It imports the reference's modules UNMODIFIED (models/nequip.py:96-217, models/segnn.py:151-297,
models/egnn.py:236-347, models/utils/graph_utils.py:13-72, models/utils/equivariant_graph_utils.py:27-102), builds
seeded inputs, lets Flax initialise the parameters, and writes per case

    tests/golden/ref_<case>.npz :  x, k, cell_std, senders, receivers, dr (kNN), cot,
                                   param::<flax path>      every parameter leaf
                                   out                     model output (fp32, jitted on CPU)
                                   grad::<flax path>       d sum(out * cot) / d param

tests/test_reference_golden.py consumes every ref_*.npz it finds: the oracle (fp64 and fp32) is evaluated on the
stored parameters and inputs and compared with `out` / `grad::*`, and the `-m gpu` half does the same with the CUDA
path.  THIS IMAGE HAS NO JAX, so no ref_*.npz is committed yet and this script has not been
executed here; until someone runs it the parity claim stays "unpinned".  The cases mirror BASELINE.json's configs at
sizes a CPU handles in seconds.
"""
import argparse
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

NEQUIP_CFG1 = dict(message_passing_steps=3, d_hidden=32, n_layers=3, l_max=1, n_radial_basis=4, task="node",
                   irreps_out="1o + 1o + 1x0e")                                  # BASELINE configs[0]
NEQUIP_CFG2 = dict(d_hidden=128, l_max=2, sphharm_norm="integral", message_passing_steps=4, n_layers=3,
                   message_passing_agg="mean", readout_agg="mean", mlp_readout_widths=(4, 2, 2), task="graph",
                   n_outputs=2, n_radial_basis=64, r_cutoff=0.6)                 # train_cosmology.py:123-134
NEQUIP_CFG5 = dict(NEQUIP_CFG2, l_max=3)
SEGNN_CFG3 = dict(d_hidden=128, n_layers=3, message_passing_steps=3, message_passing_agg="sum", task="node",
                  output_irreps="1x1o", l_max_hidden=2, residual=True)           # train_velocities.py:364-378
EGNN_CFG4 = dict(message_passing_steps=4, d_hidden=128, n_layers=3, activation="gelu", soft_edges=False,
                 positions_only=True, tanh_out=False, decouple_pos_vel_updates=True, message_passing_agg="mean",
                 readout_agg="mean", mlp_readout_widths=(4, 2, 2), task="graph", n_outputs=2, n_radial_basis=32,
                 r_max=0.6)                                                       # train_cosmology.py:67-81
STD_OVER_BOX = np.array([0.28871092, 0.28875258, 0.28870235], dtype=np.float32)   # dataset.py:16-17 / 1000


def flatten(tree, pre=()):
    for k, v in tree.items():
        if hasattr(v, "items"):
            yield from flatten(v, pre + (str(k),))
        else:
            yield "/".join(pre + (str(k),)), np.asarray(v)


def strip_wrappers(tree):
    """Drops the {'params': {'model': ...}} levels the vmap wrapper adds; returns the module's own parameter dict."""
    names = ("Linear_0", "MLP_0", "MultiLayerPerceptron_0", "TensorProductLinearGate_0", "Dense_0")
    while hasattr(tree, "items") and not any(n in tree for n in names) and len(tree) == 1:
        tree = next(iter(tree.values()))
    return tree


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    ap.add_argument("--out", default=HERE)
    args = ap.parse_args()
    sys.path.insert(0, args.reference)
    os.environ.setdefault("JAX_PLATFORMS", "cpu")

    import e3nn_jax as e3nn
    import flax.linen as nn
    import jax
    import jax.numpy as jnp
    import jraph

    from models.egnn import EGNN
    from models.nequip import NequIP
    from models.segnn import SEGNN
    from models.utils.equivariant_graph_utils import get_equivariant_graph
    from models.utils.graph_utils import get_apply_pbc, nearest_neighbors

    class GraphWrapper(nn.Module):            # tests/test_equivariance.py:24-30
        model: nn.Module

        @nn.compact
        def __call__(self, x):
            return jax.vmap(self.model)(x)

    def halos(seed, B, N):
        rng = np.random.default_rng(seed)
        pos = rng.uniform(0.0, 1000.0, size=(B, N, 3))
        vel = rng.normal(0.0, 344.0, size=(B, N, 3))
        x = np.concatenate([(pos - 500.0) / (1000.0 * STD_OVER_BOX), vel / 344.0, np.ones((B, N, 1))], -1)
        return x.astype(np.float32)

    def knn(x, k, pbc):
        apply_pbc = get_apply_pbc(std=STD_OVER_BOX) if pbc else None
        s, r, dr = jax.vmap(nearest_neighbors, in_axes=(0, None, None))(jnp.asarray(x[..., :3]), k, apply_pbc)
        return np.asarray(s), np.asarray(r), np.asarray(dr), apply_pbc

    def save(case, model, graph, fixed, out_of):
        """init -> forward -> gradient of sum(out * cot) w.r.t. the parameters (train_cosmology.py:245)."""
        wrapped = GraphWrapper(model)
        variables = wrapped.init(jax.random.PRNGKey(0), graph)
        out = jax.jit(wrapped.apply)(variables, graph)
        val = np.asarray(out_of(out))
        cot = np.random.default_rng(99).normal(size=val.shape).astype(np.float32)
        grads = jax.jit(jax.grad(lambda v: jnp.sum(out_of(wrapped.apply(v, graph)) * cot)))(variables)
        rec = dict(fixed, out=val, cot=cot)
        for name, leaf in flatten(strip_wrappers(variables)):
            rec["param::" + name] = leaf
        for name, leaf in flatten(strip_wrappers(grads)):
            rec["grad::" + name] = leaf
        path = os.path.join(args.out, f"ref_{case}.npz")
        np.savez_compressed(path, **rec)
        print(f"wrote {path}: out {val.shape}, {sum(1 for n in rec if n.startswith('param::'))} parameter leaves")

    def nequip_case(case, kw, B, N, k, pbc, seed):
        x = halos(seed, B, N)
        s, r, dr, _ = knn(x, k, pbc)
        g = jraph.GraphsTuple(n_node=np.array([[N]] * B), n_edge=np.array([[k]] * B),
                              nodes=e3nn.IrrepsArray("1o + 1o + 1x0e", jnp.asarray(x)), edges=None, globals=None,
                              senders=jnp.asarray(s), receivers=jnp.asarray(r))
        task = kw["task"]
        save(case, NequIP(**kw), g, dict(x=x, k=np.array(k), pbc=np.array(int(pbc)), cell_std=STD_OVER_BOX, senders=s,
                                         receivers=r, dr=dr),
             (lambda o: o.nodes.array) if task == "node" else (lambda o: o))

    def segnn_case(case, kw, lmax_attr, B, N, k, seed):
        x = halos(seed, B, N)[..., :3]
        s, r, dr, apply_pbc = knn(x, k, True)
        st = get_equivariant_graph(node_features=e3nn.IrrepsArray("1o", jnp.asarray(x)), positions=jnp.asarray(x),
                                   velocities=None, senders=jnp.asarray(s), receivers=jnp.asarray(r),
                                   n_node=np.array([[N]] * B), n_edge=np.array([[k]] * B), edges=None, globals=None,
                                   lmax_attributes=lmax_attr, apply_pbc=apply_pbc)
        save(case, SEGNN(**kw), st, dict(x=x, k=np.array(k), pbc=np.array(1), cell_std=STD_OVER_BOX, senders=s,
                                         receivers=r, dr=dr, lmax_attributes=np.array(lmax_attr)),
             lambda o: o.nodes.array)

    def egnn_case(case, kw, B, N, k, seed):
        x = halos(seed, B, N)[..., :3]
        s, r, dr, apply_pbc = knn(x, k, True)
        g = jraph.GraphsTuple(n_node=np.array([[N]] * B), n_edge=np.array([[k]] * B), nodes=jnp.asarray(x), edges=None,
                              globals=None, senders=jnp.asarray(s), receivers=jnp.asarray(r))
        save(case, EGNN(apply_pbc=apply_pbc, **kw), g,
             dict(x=x, k=np.array(k), pbc=np.array(1), cell_std=STD_OVER_BOX, senders=s, receivers=r, dr=dr), lambda o: o)

    nequip_case("nequip_cfg1", NEQUIP_CFG1, B=2, N=5, k=5, pbc=False, seed=0)
    nequip_case("nequip_cfg2_small", NEQUIP_CFG2, B=2, N=256, k=20, pbc=True, seed=1)     # 10 240 edges: tensor-core route
    nequip_case("nequip_cfg5_small", NEQUIP_CFG5, B=1, N=160, k=32, pbc=True, seed=2)
    segnn_case("segnn_cfg3_small", SEGNN_CFG3, 1, B=2, N=96, k=20, seed=3)
    egnn_case("egnn_cfg4_small", EGNN_CFG4, B=2, N=256, k=20, seed=4)


if __name__ == "__main__":
    main()
