"""Oracle (test infrastructure only): SEGNN forward in torch (CPU).

Restatement of models/segnn.py (reference):
  TensorProductLinearGate.__call__   :30-64    Linear(out [+ gate scalars], biases, "element"/"element")(TP(x, y)), e3nn.gate
  get_node_mlp_updates.update_fn     :73-93    concat(h, agg) -> TPLG; the first n_layers-1 TPLGs are applied to the SAME
                                               input and overwritten (:84-89): they own parameters, never reach the output
  get_edge_mlp_updates.update_fn     :105-126  concat(additional_messages, h_s, h_r) -> n_layers chained TPLGs
  SEGNN.__call__                     :198-297  embed, message passing (jraph), residual, decode / read-out
e3nn pieces (tensor_product, Linear with biases and "element" gradient normalisation, gate) are the restatements of
oracle/e3nn_ref.py.  Flax auto-names: TensorProductLinearGate_i in creation order, each
holding one Linear_0; the notebook's parameter count (573 253, examples.ipynb cell 9) is reproduced by this
bookkeeping (tests/test_oracle.py).
"""
from __future__ import annotations

import dataclasses
from typing import Dict, Optional, Sequence

import numpy as np
import torch

from . import e3nn_ref as E
from .e3nn_ref import IrrepsArray
from .so3 import Irreps, balanced_irreps


@dataclasses.dataclass
class SEGNNConfig:
    """Field names and defaults of models/segnn.py:152-165."""
    d_hidden: int = 64
    n_layers: int = 3
    message_passing_steps: int = 3
    message_passing_agg: str = "sum"
    scalar_activation: str = "gelu"
    gate_activation: str = "sigmoid"
    task: str = "graph"
    n_outputs: int = 1
    output_irreps: Optional[object] = None
    readout_agg: str = "mean"
    mlp_readout_widths: Sequence[int] = (4, 2, 2)
    l_max_hidden: int = 1
    hidden_irreps: Optional[object] = None
    residual: bool = True


# --------------------------------------------------------------------------------------------------
# TensorProductLinearGate (segnn.py:21-64)
# --------------------------------------------------------------------------------------------------
ONE = Irreps("1x0e")


def tplg_linear_irreps(output_irreps, activation: bool) -> Irreps:
    output_irreps = Irreps(output_irreps)
    if not activation:
        return output_irreps
    n_gates = output_irreps.num_irreps - sum(m for m, l, p in output_irreps if (l, p) == (0, 1))
    return (Irreps([(n_gates, 0, 1)]) + output_irreps).regroup()                    # :39-46


def tp_out_irreps(x_irreps, y_irreps) -> Irreps:
    _, out = E.tensor_product_paths(Irreps(x_irreps).regroup(), Irreps(y_irreps).regroup())
    return out.regroup()


def tplg_init(rng, x_irreps, y_irreps, output_irreps, activation: bool) -> Dict:
    lin = E.linear_init(rng, tp_out_irreps(x_irreps, y_irreps if y_irreps is not None else ONE),
                        tplg_linear_irreps(output_irreps, activation), gradient_normalization=0.0, biases=True)
    return {"Linear_0": lin}


def tplg_apply(params, x: IrrepsArray, y: Optional[IrrepsArray], output_irreps, activation: bool,
               scalar_activation: str = "gelu") -> IrrepsArray:
    if y is None:                                                                    # :35-36
        y = IrrepsArray(ONE, torch.ones(x.array.shape[:-1] + (1,), dtype=x.array.dtype))
    out = E.linear_apply(params["Linear_0"], E.tensor_product(x, y), tplg_linear_irreps(output_irreps, activation),
                         gradient_normalization=0.0)
    if activation:
        out = E.gate(out, even_act=scalar_activation)        # even gates keep the default sigmoid (:59-63)
    return out


def concatenate(arrays) -> IrrepsArray:
    """e3nn.concatenate along the last axis: irreps are appended, not regrouped."""
    irreps = Irreps(sum((list(Irreps(a.irreps).items) for a in arrays), []))
    return IrrepsArray(irreps, torch.cat([a.array for a in arrays], dim=-1))


# --------------------------------------------------------------------------------------------------
# model
# --------------------------------------------------------------------------------------------------
def _hidden(cfg: SEGNNConfig) -> Irreps:
    return Irreps(cfg.hidden_irreps) if cfg.hidden_irreps is not None else \
        balanced_irreps(cfg.l_max_hidden, cfg.d_hidden, True)


def _gate_out_irreps(output_irreps) -> Irreps:
    """Irreps of e3nn.gate's output for a gated TPLG: activated scalars first, then the gated irreps."""
    o = tplg_linear_irreps(output_irreps, True)
    n_gated = sum(m for m, l, p in o if l > 0)
    n_scal = sum(m for m, l, p in o if l == 0)
    return Irreps([(n_scal - n_gated, 0, 1)] + [(m, l, p) for m, l, p in o if l > 0]).remove_zero_multiplicities()


def init_params(cfg: SEGNNConfig, irreps_in, attr_irreps, add_irreps, seed: int = 0, n_globals: int = 0) -> Dict:
    rng = np.random.default_rng(seed)
    hidden = _hidden(cfg)
    irreps_in = Irreps(irreps_in)
    out_irreps = Irreps(cfg.output_irreps) if cfg.output_irreps is not None else irreps_in
    p: Dict = {}
    ti = 0

    def add(x_irreps, y_irreps, o_irreps, act):
        nonlocal ti
        p[f"TensorProductLinearGate_{ti}"] = tplg_init(rng, x_irreps, y_irreps, o_irreps, act)
        ti += 1

    add(irreps_in, attr_irreps, hidden, False)                                       # _embed :167-177
    h_irreps = hidden
    gated_hidden = _gate_out_irreps(hidden)
    for _ in range(cfg.message_passing_steps):
        m_irreps = Irreps(add_irreps) + h_irreps + h_irreps                          # edge fn :111-126
        for _ in range(cfg.n_layers - 1):
            add(m_irreps, attr_irreps, hidden, True)
            m_irreps = gated_hidden
        add(m_irreps, attr_irreps, hidden, False)
        n_in = h_irreps + hidden                                                     # node fn :79-93
        for _ in range(cfg.n_layers - 1):
            add(n_in, attr_irreps, hidden, True)                                     # applied to m_i, result discarded
        add(n_in, attr_irreps, hidden, False)
        h_irreps = hidden
    if cfg.task == "node":                                                           # _decode :179-196
        x_irreps = h_irreps
        for _ in range(cfg.n_layers):
            add(x_irreps, None, hidden, True)
            x_irreps = gated_hidden
        add(x_irreps, attr_irreps, out_irreps, False)
    else:
        add(h_irreps, attr_irreps, Irreps([(cfg.d_hidden, 0, 1)]), False)            # :273-277
        lim = np.sqrt(6.0 / (cfg.d_hidden + cfg.d_hidden))                           # nn.Dense default: lecun_normal
        p["Dense_0"] = {"kernel": rng.normal(size=(cfg.d_hidden, cfg.d_hidden)) / np.sqrt(cfg.d_hidden),
                        "bias": np.zeros((cfg.d_hidden,))}
        n_pool = cfg.d_hidden + n_globals
        if n_globals > 0:
            p["LayerNorm_0"] = {"scale": np.ones((n_pool,)), "bias": np.zeros((n_pool,))}
        w = [cfg.mlp_readout_widths[0] * n_pool] + [x * cfg.d_hidden for x in cfg.mlp_readout_widths[1:]] + [cfg.n_outputs]
        p["MLP_0"] = E.flax_mlp_init(rng, n_pool, w)
    return p


def apply_single(params, cfg: SEGNNConfig, nodes: IrrepsArray, senders, receivers, node_attrs: IrrepsArray,
                 edge_attrs: IrrepsArray, additional_messages: IrrepsArray, globals_=None):
    """SEGNN.__call__ on one un-batched steerable graph."""
    hidden = _hidden(cfg)
    out_irreps = Irreps(cfg.output_irreps) if cfg.output_irreps is not None else Irreps(nodes.irreps)
    senders = torch.as_tensor(np.asarray(senders)).long()
    receivers = torch.as_tensor(np.asarray(receivers)).long()
    N = nodes.array.shape[0]
    agg = E.SEGMENT[cfg.message_passing_agg]
    act = cfg.scalar_activation
    ti = 0

    def tplg(x, y, o_irreps, activation):
        nonlocal ti
        out = tplg_apply(params[f"TensorProductLinearGate_{ti}"], x, y, o_irreps, activation, act)
        ti += 1
        return out

    h = tplg(nodes, node_attrs, hidden, False)
    for _ in range(cfg.message_passing_steps):
        m = concatenate([additional_messages, IrrepsArray(h.irreps, h.array[senders]),
                         IrrepsArray(h.irreps, h.array[receivers])])
        for _ in range(cfg.n_layers - 1):
            m = tplg(m, edge_attrs, hidden, True)
        m = tplg(m, edge_attrs, hidden, False)
        a = IrrepsArray(m.irreps, agg(m.array, receivers, N))
        m_i = concatenate([h, a])
        for _ in range(cfg.n_layers - 1):
            tplg(m_i, node_attrs, hidden, True)                                      # discarded (:84-89)
        new = tplg(m_i, node_attrs, hidden, False)
        h = IrrepsArray(new.irreps, new.array + h.array) if cfg.residual else new
    if cfg.task == "node":
        x = h
        for _ in range(cfg.n_layers):
            x = tplg(x, None, hidden, True)
        return tplg(x, node_attrs, out_irreps, False)
    pre = tplg(h, node_attrs, Irreps([(cfg.d_hidden, 0, 1)]), False).array
    pre = pre @ params["Dense_0"]["kernel"] + params["Dense_0"]["bias"]
    pooled = {"mean": pre.mean(0), "sum": pre.sum(0), "max": pre.max(0).values}[cfg.readout_agg]
    if globals_ is not None:
        pooled = torch.cat([pooled, torch.as_tensor(globals_, dtype=pooled.dtype)])
        mu = pooled.mean()
        var = ((pooled - mu) ** 2).mean()
        pooled = (pooled - mu) / torch.sqrt(var + 1e-6) * params["LayerNorm_0"]["scale"] + params["LayerNorm_0"]["bias"]
    return E.flax_mlp_apply(params["MLP_0"], pooled)


def apply_batched(params, cfg: SEGNNConfig, st_graph, dtype=torch.float64):
    """vmap over the leading batch axis of an oracle SteerableGraphsTuple (oracle/graph_ref.get_equivariant_graph)."""
    outs = []
    nodes = st_graph.nodes
    B = nodes.array.shape[0]
    attr_irreps = Irreps.spherical_harmonics(int(round(np.sqrt(np.asarray(st_graph.steerable_edge_attrs).shape[-1]))) - 1)
    add = torch.as_tensor(np.asarray(st_graph.additional_messages), dtype=dtype)
    add_irreps = Irreps([(add.shape[-1], 0, 1)])
    for b in range(B):
        outs.append(apply_single(
            params, cfg, IrrepsArray(nodes.irreps, torch.as_tensor(nodes.array[b], dtype=dtype)),
            st_graph.senders[b], st_graph.receivers[b],
            IrrepsArray(attr_irreps, torch.as_tensor(np.asarray(st_graph.steerable_node_attrs)[b], dtype=dtype)),
            IrrepsArray(attr_irreps, torch.as_tensor(np.asarray(st_graph.steerable_edge_attrs)[b], dtype=dtype)),
            IrrepsArray(add_irreps, add[b]),
            None if st_graph.globals is None else st_graph.globals[b]))
    if cfg.task == "node":
        return IrrepsArray(outs[0].irreps, torch.stack([o.array for o in outs]))
    return torch.stack(outs)
