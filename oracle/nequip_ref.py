"""Oracle (test infrastructure only): NequIP forward in torch (CPU).

Line-by-line restatement of models/nequip.py (reference):
  get_edge_mlp_updates.update_fn   :39-71
  get_node_mlp_updates.update_fn   :77-91
  NequIP.__call__                  :113-217
plus the batching wrapper benchmarks/galaxies/train_cosmology.py:192-210
(``jax.vmap`` over graphs sharing parameters).  jraph.GraphNetwork semantics per
SURVEY.md A.9.  Parameters use the Flax auto-names the reference would create
(Linear_i, MultiLayerPerceptron_s, MLP_0, LayerNorm_0) so the product and the
oracle can exchange one parameter tree.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import e3nn_ref as E
from .graph_ref import GraphsTuple
from .so3 import Irreps, balanced_irreps, spherical_harmonics


@dataclasses.dataclass
class NequIPConfig:
    """Field names and defaults of models/nequip.py:97-111."""
    d_hidden: int = 128
    n_layers: int = 3
    message_passing_steps: int = 3
    message_passing_agg: str = "mean"
    activation: str = "gelu"
    task: str = "graph"
    n_outputs: int = 1
    readout_agg: str = "mean"
    mlp_readout_widths: Sequence[int] = (4, 2, 2)
    n_radial_basis: int = 4
    r_cutoff: float = 1.0
    l_max: int = 1
    sphharm_norm: str = "component"
    irreps_out: Optional[str] = None
    residual: bool = True


def _node_update_irreps(cfg: NequIPConfig, msg_irreps: Irreps):
    irreps_hidden = balanced_irreps(lmax=cfg.l_max, feature_size=cfg.d_hidden, use_sh=True)
    irreps = irreps_hidden.filter(keep=msg_irreps)                       # nequip.py:81
    gate_irreps = Irreps([(irreps.num_irreps - irreps.count("0e"), 0, 1)])  # :82
    return (gate_irreps + irreps).regroup()                              # :83


def _message_irreps(irreps_in: Irreps, l_max: int) -> Irreps:
    _, out = E.tensor_product_paths(Irreps(irreps_in).regroup(), Irreps.spherical_harmonics(l_max).regroup())
    return out.regroup()


def init_params(cfg: NequIPConfig, irreps_in, seed: int = 0, n_globals: int = 0) -> Dict:
    """Creates a parameter tree of the reference's shapes (values: the library initialisers'
    distributions drawn from numpy, not JAX's PRNG stream)."""
    rng = np.random.default_rng(seed)
    irreps_in = Irreps(irreps_in)
    msg = _message_irreps(irreps_in, cfg.l_max)
    upd = _node_update_irreps(cfg, msg)
    gated_out = Irreps(E.gate(E.IrrepsArray(upd, torch.zeros(1, upd.dim))).irreps)
    params, li = {}, 0
    for s in range(cfg.message_passing_steps):
        params[f"Linear_{li}"] = E.linear_init(rng, irreps_in, irreps_in); li += 1
        n_rb = cfg.n_radial_basis if cfg.n_radial_basis > 0 else 1
        params[f"MultiLayerPerceptron_{s}"] = E.e3nn_mlp_init(
            rng, n_rb, cfg.n_layers * (cfg.d_hidden,) + (msg.num_irreps,))
        params[f"Linear_{li}"] = E.linear_init(rng, msg, upd); li += 1
        if cfg.residual:
            params[f"Linear_{li}"] = E.linear_init(rng, irreps_in, upd, force_irreps_out=True); li += 1
        params[f"Linear_{li}"] = E.linear_init(rng, gated_out, irreps_in); li += 1
    irreps_out = Irreps(cfg.irreps_out) if cfg.irreps_out is not None else irreps_in
    if cfg.task == "node":
        if irreps_out != irreps_in:
            params[f"Linear_{li}"] = E.linear_init(rng, irreps_in, irreps_out); li += 1
    else:
        params[f"Linear_{li}"] = E.linear_init(rng, msg, Irreps([(cfg.d_hidden, 0, 1)])); li += 1
        n_pool = cfg.d_hidden + n_globals
        if n_globals > 0:
            params["LayerNorm_0"] = {"scale": np.ones((n_pool,)), "bias": np.zeros((n_pool,))}
        widths = [cfg.mlp_readout_widths[0] * n_pool] + \
                 [w * cfg.d_hidden for w in cfg.mlp_readout_widths[1:]] + [cfg.n_outputs]
        params["MLP_0"] = E.flax_mlp_init(rng, n_pool, widths)
    return params


def count_params(params) -> int:
    if isinstance(params, dict):
        return sum(count_params(v) for v in params.values())
    return int(np.prod(np.shape(params)))


def to_torch(params, dtype=torch.float64, requires_grad=False):
    if isinstance(params, dict):
        return {k: to_torch(v, dtype, requires_grad) for k, v in params.items()}
    t = torch.as_tensor(np.asarray(params), dtype=dtype).clone()
    return t.requires_grad_(requires_grad)


def apply_single(params, cfg: NequIPConfig, graph: GraphsTuple, irreps_in, intermediates: dict = None):
    """NequIP.__call__ on ONE un-batched graph (nodes (N,D), senders/receivers (E,), n_edge (1,))."""
    irreps_in = Irreps(irreps_in)
    dtype = graph.nodes.dtype
    nodes = E.IrrepsArray(irreps_in, graph.nodes)
    senders = torch.as_tensor(np.asarray(graph.senders)).long()
    receivers = torch.as_tensor(np.asarray(graph.receivers)).long()
    n_edge = torch.as_tensor(np.asarray(graph.n_edge), dtype=dtype).reshape(-1)
    N = nodes.array.shape[0]
    agg = E.SEGMENT[cfg.message_passing_agg]
    irreps_hidden = balanced_irreps(lmax=cfg.l_max, feature_size=cfg.d_hidden, use_sh=True)
    irreps_out = Irreps(cfg.irreps_out) if cfg.irreps_out is not None else irreps_in

    # nequip.py:129-134.  The reference holds positions in fp32; r_ij and |r_ij| are formed with its
    # fp32 rounding points (they feed the ill-conditioned Bessel phase, see e3nn_ref.bessel), then promoted.
    pos32 = nodes.first_irrep_copy().array.to(torch.float32)
    r32 = pos32[senders] - pos32[receivers]
    sq = r32 * r32
    d32 = torch.sqrt((sq[:, 0] + sq[:, 1]) + sq[:, 2])          # jnp.linalg.norm (:55), sequential fp32 sum
    r_ij = r32.to(dtype)
    li = 0
    for s in range(cfg.message_passing_steps):
        # ---- update_edge_fn (:39-71), called by jraph.GraphNetwork with nodes[senders]
        sent = E.IrrepsArray(irreps_in, nodes.array[senders])
        m = E.linear_apply(params[f"Linear_{li}"], sent, irreps_in); li += 1         # :43
        a = E.IrrepsArray(Irreps.spherical_harmonics(cfg.l_max),
                          spherical_harmonics(cfg.l_max, r_ij, True, cfg.sphharm_norm))  # :46-51
        m = E.tensor_product(m, a)                                                    # :52
        d = d32.to(dtype)                                                             # :55
        R = E.bessel(d, cfg.n_radial_basis, cfg.r_cutoff) if cfg.n_radial_basis > 0 else d[..., None]
        W = E.e3nn_mlp_apply(params[f"MultiLayerPerceptron_{s}"], R, cfg.activation, False)  # :62-66
        m = m.mul_by_scalars(W)                                                       # :68
        # ---- aggregation over receivers (:118-120)
        A = E.IrrepsArray(m.irreps, agg(m.array, receivers, N))
        # ---- update_node_fn (:77-91)
        m_i = E.IrrepsArray(A.irreps, A.array / torch.sqrt(n_edge))                   # :80
        upd = _node_update_irreps(cfg, m_i.irreps)
        z = E.linear_apply(params[f"Linear_{li}"], m_i, upd); li += 1                 # :85
        if cfg.residual:
            z2 = E.linear_apply(params[f"Linear_{li}"], nodes, upd, force_irreps_out=True); li += 1
            z = E.IrrepsArray(z.irreps, z.array + z2.array)
        g = E.gate(z)                                                                 # :88
        nodes = E.linear_apply(params[f"Linear_{li}"], g, irreps_in); li += 1         # :162
        if intermediates is not None:
            intermediates[f"W_{s}"] = W
            intermediates[f"messages_{s}"] = m.array
            intermediates[f"agg_{s}"] = A.array
            intermediates[f"nodes_{s}"] = nodes.array

    if cfg.task == "node":
        if irreps_out != irreps_in:
            nodes = E.linear_apply(params[f"Linear_{li}"], nodes, irreps_out); li += 1
        return graph._replace(nodes=nodes.array, edges=m.array)
    if cfg.task != "graph":
        raise ValueError(f"Invalid task {cfg.task}")
    a = E.IrrepsArray(Irreps.spherical_harmonics(cfg.l_max),
                      spherical_harmonics(cfg.l_max, r_ij, True, cfg.sphharm_norm))   # :180-185
    attrs = E.IrrepsArray(a.irreps, E.segment_sum(a.array, receivers, N) / n_edge)     # :186-191
    tp = E.tensor_product(nodes, attrs)
    pre = E.linear_apply(params[f"Linear_{li}"], tp, Irreps([(cfg.d_hidden, 0, 1)])).array  # :197
    if cfg.readout_agg == "mean":
        pooled = pre.mean(0)
    elif cfg.readout_agg == "sum":
        pooled = pre.sum(0)
    else:
        pooled = pre.max(0).values
    if graph.globals is not None:                                                     # :201-205
        pooled = torch.cat([pooled, torch.as_tensor(graph.globals, dtype=dtype)])
        mu = pooled.mean()
        var = ((pooled - mu) ** 2).mean()
        pooled = (pooled - mu) / torch.sqrt(var + 1e-6)
        pooled = pooled * params["LayerNorm_0"]["scale"] + params["LayerNorm_0"]["bias"]
    return E.flax_mlp_apply(params["MLP_0"], pooled)                                  # :208-214


def apply_batched(params, cfg: NequIPConfig, graphs: GraphsTuple, irreps_in, intermediates: list = None):
    """``jax.vmap(NequIP(**cfg))(graphs)`` (train_cosmology.py:210)."""
    B = graphs.nodes.shape[0]
    outs = []
    for b in range(B):
        g = GraphsTuple(nodes=graphs.nodes[b], edges=None, receivers=graphs.receivers[b],
                        senders=graphs.senders[b],
                        globals=None if graphs.globals is None else graphs.globals[b],
                        n_node=graphs.n_node[b], n_edge=graphs.n_edge[b])
        inter = {} if intermediates is not None else None
        outs.append(apply_single(params, cfg, g, irreps_in, inter))
        if intermediates is not None:
            intermediates.append(inter)
    if cfg.task == "node":
        return graphs._replace(nodes=torch.stack([o.nodes for o in outs]),
                               edges=torch.stack([o.edges for o in outs]))
    return torch.stack(outs)
