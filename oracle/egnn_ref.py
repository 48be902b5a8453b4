"""Oracle (test infrastructure only): EGNN forward in torch (CPU).

Restatement of models/egnn.py (reference):
  get_edge_mlp_updates.update_fn  :23-116   r_ij (+1e-7, PBC), |r|, e3nn.bessel, phi_e, soft edges, phi_x, clip
  get_node_mlp_updates.update_fn  :129-231  x' = PBC(x + agg), optional phi_h / phi_v / phi_xx
  EGNN.__call__                   :257-347  message passing, graph read-out (mean over nodes -> MLP)
and of models/mlp.py (Dense + bias, xavier-uniform, gelu-tanh).  jraph.GraphNetwork semantics per
SURVEY.md A.9 (the tuple (x_ij_trans, m_ij) is aggregated over receivers leaf by leaf).
Flax auto-names: inside one step the edge function creates phi_e, phi_x, Dense_{s} (last layer of phi_x)
and phi_a (its MLP index is reserved, parameters only when soft_edges), then the node function its MLPs
(phi_v, phi_h, phi_xx as the node layout requires); MLP indices run on across steps.  The notebook's parameter count (441 778, cell 15) is reproduced by this bookkeeping.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Dict, Optional, Sequence

import numpy as np
import torch

from . import e3nn_ref as E
from .graph_ref import GraphsTuple


@dataclasses.dataclass
class EGNNConfig:
    """Field names and defaults of models/egnn.py:240-255."""
    message_passing_steps: int = 3
    d_hidden: int = 128
    n_layers: int = 3
    activation: str = "gelu"
    soft_edges: bool = True
    positions_only: bool = False
    tanh_out: bool = False
    n_radial_basis: int = 16
    r_max: float = 0.3
    decouple_pos_vel_updates: bool = False
    message_passing_agg: str = "sum"
    readout_agg: str = "mean"
    mlp_readout_widths: Sequence[int] = (4, 2, 2)
    task: str = "graph"
    n_outputs: int = 1
    apply_pbc: Optional[Callable] = None


def _mlp_init(rng, n_in, widths):
    return E.flax_mlp_init(rng, n_in, widths)


def _variant(cfg: EGNNConfig, F: int):
    """(has_vel, d_scalar) of a node-feature width, egnn.py:38-60 / :145-229."""
    if cfg.positions_only:
        if F < 3:
            raise ValueError("positions_only needs >= 3 node features")
        return False, F - 3
    if F < 6:
        raise ValueError("positions + velocities need >= 6 node features")
    return True, F - 6


def init_params(cfg: EGNNConfig, n_feat: int, seed: int = 0, n_globals: int = 0) -> Dict:
    """All four node-feature layouts of the reference: (x), (x, h), (x, v), (x, v, h), with optional graph globals
    (broadcast to every edge and node by jraph; egnn.py:41,46,53,60,163-164,181-182,211-212,332-336).
    Flax auto-names follow construction order: phi_e, phi_x, [Dense], phi_a, then the node function's phi_v / phi_h /
    phi_xx (egnn.py:63-79,100,161,177-189,210-222)."""
    rng = np.random.default_rng(seed)
    p: Dict = {}
    d, L = cfg.d_hidden, cfg.n_layers
    n_msg = cfg.n_radial_basis if cfg.n_radial_basis > 0 else 1
    mi, F, G_ = 0, n_feat, n_globals
    for s in range(cfg.message_passing_steps):
        vel, dh = _variant(cfg, F)
        p[f"MLP_{mi}"] = _mlp_init(rng, n_msg + 2 * dh + G_, [d] * (L - 1) + [d]); mi += 1   # phi_e
        p[f"MLP_{mi}"] = _mlp_init(rng, d, [d] * (L - 1)); mi += 1                      # phi_x
        lim = math.sqrt(3.0 * 1e-2 / d)                                                 # variance_scaling(1e-2, fan_in, uniform)
        p[f"Dense_{s}"] = {"kernel": rng.uniform(-lim, lim, size=(d, 1))}
        if cfg.soft_edges:
            p[f"MLP_{mi}"] = _mlp_init(rng, d, [d])                                     # phi_a
        mi += 1
        if not vel:
            if dh > 0:
                p[f"MLP_{mi}"] = _mlp_init(rng, dh + d + G_, [d] * (L - 1) + [dh]); mi += 1  # phi_h([h_i, m_i, globals])
        else:
            n_in = (d if dh == 0 else dh) + G_                                          # concats = [m_i | h_i (, globals)]
            p[f"MLP_{mi}"] = _mlp_init(rng, n_in, [d] * (L - 1) + [1]); mi += 1         # phi_v
            if dh > 0:
                p[f"MLP_{mi}"] = _mlp_init(rng, n_in, [d] * (L - 1) + [dh]); mi += 1    # phi_h
            if cfg.decouple_pos_vel_updates:
                p[f"MLP_{mi}"] = _mlp_init(rng, n_in, [d] * (L - 1) + [1]); mi += 1     # phi_xx
            if dh == 0:
                F = 6 + d + G_                                                          # returns [x', v, concats] (:196)
    if cfg.task == "graph":
        n_pool = F + G_
        if G_ > 0:
            p["LayerNorm_0"] = {"scale": np.ones((n_pool,)), "bias": np.zeros((n_pool,))}
        w = [cfg.mlp_readout_widths[0] * n_pool] + [x * d for x in cfg.mlp_readout_widths[1:]] + [cfg.n_outputs]
        p[f"MLP_{mi}"] = _mlp_init(rng, n_pool, w)
    return p


def _mlp(params, x, activate_final):
    return E.flax_mlp_apply(params, x, activate_final=activate_final)


def apply_single(params, cfg: EGNNConfig, graph: GraphsTuple):
    """EGNN.__call__ on one un-batched graph; nodes (N, F)."""
    nodes = graph.nodes
    dtype = nodes.dtype
    senders = torch.as_tensor(np.asarray(graph.senders)).long()
    receivers = torch.as_tensor(np.asarray(graph.receivers)).long()
    N = nodes.shape[0]
    agg = E.SEGMENT[cfg.message_passing_agg]
    pbc = cfg.apply_pbc
    if pbc is not None:
        cell = torch.as_tensor(pbc.cell, dtype=dtype)
        cell_inv = torch.as_tensor(pbc.cell_inv, dtype=dtype)
    wrap = (lambda a: a - torch.round(a @ cell_inv) @ cell) if pbc is not None else (lambda a: a)
    glob = None if graph.globals is None else torch.as_tensor(graph.globals, dtype=dtype).reshape(1, -1)   # :268-271
    g_e = None if glob is None else glob.expand(senders.shape[0], -1)
    g_n = None if glob is None else glob.expand(N, -1)
    mi = 0
    for s in range(cfg.message_passing_steps):
        vel, dh = _variant(cfg, nodes.shape[1])
        x = nodes[:, :3]
        v = nodes[:, 3:6] if vel else None
        h = nodes[:, (6 if vel else 3):] if dh > 0 else None
        # ---- edge update, egnn.py:38-116 (jraph: "senders" = nodes[senders], "receivers" = nodes[receivers])
        r_ij = wrap(x[senders] - x[receivers] + 1e-7)
        d = torch.sqrt((r_ij ** 2).sum(1))
        feats = E.bessel(d, cfg.n_radial_basis, cfg.r_max, reference_rounding=False) if cfg.n_radial_basis > 0 \
            else d[:, None]
        if dh > 0:
            feats = torch.cat([feats, h[senders], h[receivers]], -1)
        if g_e is not None:
            feats = torch.cat([feats, g_e], -1)
        m = _mlp(params[f"MLP_{mi}"], feats, True); mi += 1
        phi_x = params[f"MLP_{mi}"]; mi += 1
        if cfg.soft_edges:
            m = m * torch.sigmoid(_mlp(params[f"MLP_{mi}"], m, False))
        mi += 1
        t = _mlp(phi_x, m, True) @ params[f"Dense_{s}"]["kernel"]
        if cfg.tanh_out:
            t = torch.tanh(t)
        xt = torch.clamp(r_ij * t, -100.0, 100.0)
        agg_x, m_i = agg(xt, receivers, N), agg(m, receivers, N)
        # ---- node update, egnn.py:143-229
        if not vel:
            x_p = wrap(x + agg_x)
            if dh == 0:
                nodes = x_p
            else:
                c = torch.cat([h, m_i] + ([g_n] if g_n is not None else []), -1)
                h_p = h + _mlp(params[f"MLP_{mi}"], c, False); mi += 1
                nodes = torch.cat([x_p, h_p], -1)
        else:
            c = m_i if dh == 0 else h
            if g_n is not None:
                c = torch.cat([c, g_n], -1)
            v_p = _mlp(params[f"MLP_{mi}"], c, False) * v + agg_x; mi += 1
            if dh > 0:
                h_p = h + _mlp(params[f"MLP_{mi}"], c, False); mi += 1
            if cfg.decouple_pos_vel_updates:
                x_p = _mlp(params[f"MLP_{mi}"], c, False) * x + v_p; mi += 1
            else:
                x_p = x + v_p
            x_p = wrap(x_p)
            nodes = torch.cat([x_p, v, c if dh == 0 else h_p], -1)      # the OLD velocity is returned (:196,229)
    if cfg.task == "node":
        return graph._replace(nodes=nodes)
    pooled = {"mean": nodes.mean(0), "sum": nodes.sum(0), "max": nodes.max(0).values}[cfg.readout_agg]
    if glob is not None:                                                # :332-336
        pooled = torch.cat([pooled, glob[0]])
        mu = pooled.mean()
        var = ((pooled - mu) ** 2).mean()
        pooled = (pooled - mu) / torch.sqrt(var + 1e-6) * params["LayerNorm_0"]["scale"] + params["LayerNorm_0"]["bias"]
    return _mlp(params[f"MLP_{mi}"], pooled, False)


def apply_batched(params, cfg: EGNNConfig, graphs: GraphsTuple):
    outs = []
    for b in range(graphs.nodes.shape[0]):
        g = GraphsTuple(nodes=graphs.nodes[b], edges=None, receivers=graphs.receivers[b], senders=graphs.senders[b],
                        globals=None if graphs.globals is None else graphs.globals[b], n_node=None, n_edge=None)
        outs.append(apply_single(params, cfg, g))
    if cfg.task == "node":
        return graphs._replace(nodes=torch.stack([o.nodes for o in outs]))
    return torch.stack(outs)
