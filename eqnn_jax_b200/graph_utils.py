"""Graph construction with the reference's call signatures, on B200 kernels.

Drop-in mirror of ``models/utils/graph_utils.py`` (reference) for the hot-path
functions ``get_apply_pbc`` (:13-37), ``nearest_neighbors`` (:40-72) and
``build_graph`` (:248-324, kNN branch).  Tensors are torch CUDA tensors; the
returned ``GraphsTuple`` has jraph's field names so model code reads the same.
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np
import torch

from . import ops

EPS = 1e-7  # graph_utils.py:11


class GraphsTuple(NamedTuple):
    """Field-for-field mirror of ``jraph.GraphsTuple``."""
    nodes: Optional[torch.Tensor]
    edges: Optional[torch.Tensor]
    receivers: Optional[torch.Tensor]
    senders: Optional[torch.Tensor]
    globals: Optional[torch.Tensor]
    n_node: Optional[torch.Tensor]
    n_edge: Optional[torch.Tensor]


class ApplyPBC:
    """Callable returned by :func:`get_apply_pbc`; carries the fp32 cell and its inverse
    (``np.linalg.inv`` in fp32, as the reference evaluates it at graph_utils.py:36)."""

    def __init__(self, cell: np.ndarray):
        self.cell = np.ascontiguousarray(cell, dtype=np.float32)
        self.cell_inv = np.ascontiguousarray(np.linalg.inv(self.cell), dtype=np.float32)

    def __call__(self, dr: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError(
            "apply_pbc is fused into the kNN kernel (eqnn_knn_pbc); stand-alone application is not on the NequIP path")


def get_apply_pbc(std=None, cell=None) -> ApplyPBC:
    """graph_utils.py:13-37: ``cell = cell / std[:3]`` when ``std`` is given."""
    if cell is None:
        cell = np.eye(3, dtype=np.float32)
    cell = np.asarray(cell, dtype=np.float32)
    if std is not None:
        std = np.asarray(std.detach().cpu() if isinstance(std, torch.Tensor) else std, dtype=np.float32)
        cell = (cell / std[:3]).astype(np.float32)
    return ApplyPBC(cell)


def nearest_neighbors(x: torch.Tensor, k: int, apply_pbc: Optional[ApplyPBC] = None):
    """graph_utils.py:40-72 for one graph: x [N, 3] -> (sources [N*k], targets [N*k], dr [N*k, 3])."""
    if x.dim() != 2:
        raise ValueError("nearest_neighbors expects positions of one graph, shape (N, 3)")
    x3 = x[:, :3].contiguous().unsqueeze(0)
    s, t, dr = ops.knn_pbc(x3, k, None if apply_pbc is None else apply_pbc.cell,
                           None if apply_pbc is None else apply_pbc.cell_inv)
    return s[0], t[0], dr[0]


def build_graph(node_feats: torch.Tensor, global_feats, k: int, use_edges: bool = True,
                apply_pbc: Optional[ApplyPBC] = None, n_radial_basis: int = 0, r_max: float = 1.0, radius=None,
                use_3d_distances: bool = False, use_invariant_edges: bool = False) -> GraphsTuple:
    """graph_utils.py:248-324 (kNN branch).  node_feats [B, N, F] fp32 on the GPU."""
    if use_invariant_edges and n_radial_basis > 0:
        raise ValueError("Cannot use both invariant edges and radial basis functions!")
    if use_invariant_edges:
        raise NotImplementedError("use_invariant_edges is not on the equivariant hot path")
    B, N, _ = node_feats.shape
    node_feats = node_feats.contiguous()
    dev = node_feats.device
    if radius is not None:
        return _build_radius_graph(node_feats, global_feats, float(radius), use_edges, apply_pbc, n_radial_basis, r_max,
                                   use_3d_distances)
    senders, receivers, dr = ops.knn_pbc(node_feats, k, None if apply_pbc is None else apply_pbc.cell,
                                         None if apply_pbc is None else apply_pbc.cell_inv)
    n_node = torch.full((B, 1), N, dtype=torch.int32)                  # host-side metadata, like jraph
    n_edge = torch.full((B, 1), k, dtype=torch.int32)                  # k, not N*k (graph_utils.py:284)
    if use_edges:
        if use_3d_distances:
            edges = dr
            if n_radial_basis > 0:
                raise NotImplementedError("bessel of 3-d displacements is not on the NequIP path")
        else:
            edges = ops.edge_features(dr, n_radial_basis, r_max)
    else:
        raise AttributeError("'NoneType' object has no attribute 'shape'")  # the reference fails at :305 too
    return GraphsTuple(nodes=node_feats, edges=edges, receivers=receivers, senders=senders, globals=global_feats,
                       n_node=n_node, n_edge=n_edge)


def _build_radius_graph(node_feats, global_feats, radius, use_edges, apply_pbc, n_radial_basis, r_max, use_3d_distances):
    """The ``radius is not None`` branch of build_graph (graph_utils.py:266-277).  Upstream it cannot run (``x`` is
    undefined at :277, np.where of the batched mask yields three arrays, the distances are already scalars at :298); this
    is what it evidently means: every pair with 0 < |apply_pbc(x_i - x_j)| < radius, compute_distances' arithmetic
    (:209-226, no EPS), in np.where order.  The edge list is ragged, so -- unlike the kNN branch -- senders / receivers
    are flat (n_total,) arrays of per-graph LOCAL node ids, graph b owning the next n_edge[b] entries, and ``edges`` is
    (1, n_total, 1 | n_radial_basis) as :319-320 leaves it.  The batched models of this package take equal-sized graphs
    (kNN); the radius graph is a graph-construction utility."""
    if use_3d_distances:
        raise NotImplementedError("radius graphs carry scalar distances (compute_distances, graph_utils.py:209-226)")
    B, N, _ = node_feats.shape
    senders, receivers, dist, n_edge = ops.radius_graph(node_feats, radius, None if apply_pbc is None else apply_pbc.cell,
                                                        None if apply_pbc is None else apply_pbc.cell_inv)
    if not use_edges:
        raise AttributeError("'NoneType' object has no attribute 'shape'")  # the reference fails at :319 too
    if dist.numel() == 0:
        edges = torch.empty((1, 0, max(n_radial_basis, 1)), dtype=torch.float32, device=dist.device)
    else:
        d3 = torch.zeros((dist.numel(), 3), dtype=torch.float32, device=dist.device)
        d3[:, 0] = dist                                                # |(d, 0, 0)| = d exactly
        edges = ops.edge_features(d3, n_radial_basis, r_max).unsqueeze(0)
    return GraphsTuple(nodes=node_feats, edges=edges, receivers=receivers, senders=senders, globals=global_feats,
                       n_node=torch.full((B, 1), N, dtype=torch.int32), n_edge=n_edge.to(torch.int32))
