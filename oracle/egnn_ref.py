"""Oracle (test infrastructure only): EGNN forward in torch (CPU).

Restatement of models/egnn.py (reference):
  get_edge_mlp_updates.update_fn  :23-116   r_ij (+1e-7, PBC), |r|, e3nn.bessel, phi_e, soft edges, phi_x, clip
  get_node_mlp_updates.update_fn  :129-231  x' = PBC(x + agg), optional phi_h / phi_v / phi_xx
  EGNN.__call__                   :257-347  message passing, graph read-out (mean over nodes -> MLP)
and of models/mlp.py (Dense + bias, xavier-uniform, gelu-tanh).  jraph.GraphNetwork semantics per
SURVEY.md A.9 (the tuple (x_ij_trans, m_ij) is aggregated over receivers leaf by leaf).
Flax auto-names: inside one step the edge function creates MLP_{3s} (phi_e), MLP_{3s+1} (phi_x),
Dense_{s} (last layer of phi_x), MLP_{3s+2} (phi_a, parameters only when soft_edges), then the node
function its MLPs.  The notebook's parameter count (441 778, cell 15) is reproduced by this bookkeeping.
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


def init_params(cfg: EGNNConfig, n_feat: int, seed: int = 0) -> Dict:
    """Positions-only clouds without scalar attributes (n_feat == 3): the benchmarked configuration
    (train_cosmology.py:67-81 with feats='pos')."""
    if not (cfg.positions_only and n_feat == 3):
        raise NotImplementedError("oracle covers positions_only=True with 3-d nodes")
    rng = np.random.default_rng(seed)
    p: Dict = {}
    d, L = cfg.d_hidden, cfg.n_layers
    n_msg = cfg.n_radial_basis if cfg.n_radial_basis > 0 else 1
    mi = 0
    for s in range(cfg.message_passing_steps):
        p[f"MLP_{mi}"] = _mlp_init(rng, n_msg, [d] * (L - 1) + [d]); mi += 1           # phi_e
        p[f"MLP_{mi}"] = _mlp_init(rng, d, [d] * (L - 1)); mi += 1                      # phi_x
        lim = math.sqrt(3.0 * 1e-2 / d)                                                 # variance_scaling(1e-2, fan_in, uniform)
        p[f"Dense_{s}"] = {"kernel": rng.uniform(-lim, lim, size=(d, 1))}
        if cfg.soft_edges:
            p[f"MLP_{mi}"] = _mlp_init(rng, d, [d])                                     # phi_a
        mi += 1
    if cfg.task == "graph":
        w = [cfg.mlp_readout_widths[0] * n_feat] + [x * d for x in cfg.mlp_readout_widths[1:]] + [cfg.n_outputs]
        p[f"MLP_{mi}"] = _mlp_init(rng, n_feat, w)
    return p


def _mlp(params, x, activate_final):
    return E.flax_mlp_apply(params, x, activate_final=activate_final)


def apply_single(params, cfg: EGNNConfig, graph: GraphsTuple):
    """EGNN.__call__ on one un-batched graph; nodes (N,3)."""
    x = graph.nodes
    dtype = x.dtype
    senders = torch.as_tensor(np.asarray(graph.senders)).long()
    receivers = torch.as_tensor(np.asarray(graph.receivers)).long()
    N = x.shape[0]
    agg = E.SEGMENT[cfg.message_passing_agg]
    pbc = cfg.apply_pbc
    mi = 0
    for s in range(cfg.message_passing_steps):
        x_i, x_j = x[senders], x[receivers]
        # egnn.py:82-85; the reference forms r_ij in fp32: keep its rounding points for the ill-conditioned Bessel phase
        r_ij = x_i - x_j + 1e-7
        if pbc is not None:
            cell = torch.as_tensor(pbc.cell, dtype=dtype)
            cell_inv = torch.as_tensor(pbc.cell_inv, dtype=dtype)
            r_ij = r_ij - torch.round(r_ij @ cell_inv) @ cell
        d = torch.sqrt((r_ij ** 2).sum(1))
        feats = E.bessel(d, cfg.n_radial_basis, cfg.r_max, reference_rounding=False) if cfg.n_radial_basis > 0 \
            else d[:, None]
        m = _mlp(params[f"MLP_{mi}"], feats, True); mi += 1
        phi_x = params[f"MLP_{mi}"]; mi += 1
        if cfg.soft_edges:
            a = torch.sigmoid(_mlp(params[f"MLP_{mi}"], m, False))
            m = m * a
        mi += 1
        t = _mlp(phi_x, m, True) @ params[f"Dense_{s}"]["kernel"]
        if cfg.tanh_out:
            t = torch.tanh(t)
        xt = torch.clamp(r_ij * t, -100.0, 100.0)
        agg_x = agg(xt, receivers, N)
        x = x + agg_x                                                    # egnn.py:146-150
        if pbc is not None:
            x = x - torch.round(x @ cell_inv) @ cell
    if cfg.task == "node":
        return graph._replace(nodes=x)
    pooled = {"mean": x.mean(0), "sum": x.sum(0), "max": x.max(0).values}[cfg.readout_agg]
    return _mlp(params[f"MLP_{mi}"], pooled, False)


def apply_batched(params, cfg: EGNNConfig, graphs: GraphsTuple):
    outs = []
    for b in range(graphs.nodes.shape[0]):
        g = GraphsTuple(nodes=graphs.nodes[b], edges=None, receivers=graphs.receivers[b], senders=graphs.senders[b],
                        globals=None, n_node=None, n_edge=None)
        outs.append(apply_single(params, cfg, g))
    if cfg.task == "node":
        return graphs._replace(nodes=torch.stack([o.nodes for o in outs]))
    return torch.stack(outs)
