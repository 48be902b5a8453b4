"""Oracle (test infrastructure only): periodic-boundary kNN graph construction.

NumPy float32 restatement of models/utils/graph_utils.py (reference):
  get_apply_pbc      :13-37   minimum image  dr - round(dr @ inv(cell)) @ cell
  nearest_neighbors  :40-72   dense (N,N,3) difference (+1e-7), PBC, d^2,
                              stable argsort, first k
  build_graph        :248-324 vmap over the batch, n_edge = [[k]]*B (k, not E)
NumPy evaluates every float32 operation with a separate IEEE rounding (no FMA
contraction), which is the arithmetic SURVEY.md C.3 specifies.  Memory is
O(N^2): use only at sizes that finish in seconds (N <= ~5000).
"""
from __future__ import annotations

from typing import Callable, NamedTuple, Optional

import numpy as np

EPS = np.float32(1e-7)  # graph_utils.py:11


class GraphsTuple(NamedTuple):
    """Field-for-field mirror of jraph.GraphsTuple."""
    nodes: object
    edges: object
    receivers: object
    senders: object
    globals: object
    n_node: object
    n_edge: object


def get_apply_pbc(std=None, cell=None) -> Callable:
    """graph_utils.py:13-37."""
    if cell is None:
        cell = np.eye(3, dtype=np.float32)
    cell = np.asarray(cell, dtype=np.float32)
    if std is not None:
        cell = (cell / np.asarray(std, dtype=np.float32)[:3]).astype(np.float32)
    cell_inv = np.linalg.inv(cell).astype(np.float32)

    def apply_pbc(dr):
        dr = np.asarray(dr, dtype=np.float32)
        n = np.round(_dot3(dr, cell_inv))          # np.round == rint, ties to even
        return (dr - _dot3(n, cell)).astype(np.float32)

    apply_pbc.cell = cell
    apply_pbc.cell_inv = cell_inv
    return apply_pbc


def _dot3(a, m):
    """(...,3) @ (3,3) in float32 with the fixed order ((a0*m0 + a1*m1) + a2*m2).

    For the diagonal cells every reference caller uses, two of the three
    products are exact zeros, so the order is immaterial."""
    a = a.astype(np.float32)
    m = m.astype(np.float32)
    return ((a[..., 0:1] * m[0] + a[..., 1:2] * m[1]) + a[..., 2:3] * m[2]).astype(np.float32)


def nearest_neighbors(x, k: int, apply_pbc: Optional[Callable] = None):
    """graph_utils.py:40-72 -> (sources, targets, dr[sources, targets])."""
    x = np.asarray(x, dtype=np.float32)
    n = x.shape[0]
    dr = (x[:, None, :] - x[None, :, :]) + EPS
    if apply_pbc is not None:
        dr = apply_pbc(dr)
    sq = dr * dr
    dist = (sq[..., 0] + sq[..., 1]) + sq[..., 2]
    idx = np.argsort(dist, axis=-1, kind="stable")[:, :k]
    sources = np.repeat(np.arange(n, dtype=np.int32), k)
    targets = idx.reshape(-1).astype(np.int32)
    return sources, targets, dr[sources, targets]


def compute_distances(x, apply_pbc: Optional[Callable] = None):
    """graph_utils.py:209-226: pairwise distances of one graph, no EPS."""
    x = np.asarray(x, dtype=np.float32)
    dr = x[:, None, :] - x[None, :, :]
    if apply_pbc is not None:
        dr = apply_pbc(dr)
    sq = dr * dr
    return np.sqrt((sq[..., 0] + sq[..., 1]) + sq[..., 2]).astype(np.float32)


def radius_graph(x, radius: float, apply_pbc: Optional[Callable] = None):
    """The `radius is not None` branch of build_graph (graph_utils.py:266-277) as it evidently means to work (upstream:
    `x` undefined at :277, np.where of the batched mask returns three arrays): per graph, mask = (d < radius) & (d > 0),
    sources, targets = np.where(mask), distances[sources, targets]; concatenated over the batch with per-graph counts."""
    x = np.asarray(x, dtype=np.float32)
    src, dst, dist, n_edge = [], [], [], []
    for b in range(x.shape[0]):
        d = compute_distances(x[b, :, :3], apply_pbc)
        mask = (d < np.float32(radius)) & (d > 0)
        s, t = np.where(mask)
        src.append(s.astype(np.int32)); dst.append(t.astype(np.int32)); dist.append(d[s, t]); n_edge.append(int(mask.sum()))
    return np.concatenate(src), np.concatenate(dst), np.concatenate(dist), np.array(n_edge)


def bessel_np(x, n, x_max):
    x = np.asarray(x, dtype=np.float32)[..., None]
    j = np.arange(1, n + 1, dtype=np.float32)
    xs = np.where(x == 0, np.float32(1), x)
    val = np.where(x == 0, j * np.float32(np.pi) / np.float32(x_max),
                   np.sin(j * np.float32(np.pi) / np.float32(x_max) * xs) / xs)
    return (np.sqrt(np.float32(2.0) / np.float32(x_max)) * val).astype(np.float32)


def build_graph(node_feats, global_feats, k, use_edges=True, apply_pbc=None, n_radial_basis=0,
                r_max=1.0, use_3d_distances=False) -> GraphsTuple:
    """graph_utils.py:248-324, kNN branch (radius branch is broken upstream: `x` undefined :277)."""
    node_feats = np.asarray(node_feats, dtype=np.float32)
    B, N = node_feats.shape[:2]
    n_node = np.array([[N]] * B)
    res = [nearest_neighbors(node_feats[b, :, :3], k, apply_pbc) for b in range(B)]
    sources = np.stack([r[0] for r in res])
    targets = np.stack([r[1] for r in res])
    distances = np.stack([r[2] for r in res])
    n_edge = np.array(B * [[k]])
    if use_edges:
        if use_3d_distances:
            edges = distances
        else:
            edges = np.sqrt(np.sum(distances ** 2, axis=-1, keepdims=True))
        if n_radial_basis > 0:
            edges = np.squeeze(bessel_np(edges, n_radial_basis, r_max))
        if use_3d_distances:
            edges = edges.reshape(edges.shape[0], edges.shape[1], -1)
    else:
        edges = None
    if edges is not None and edges.ndim < 3:
        edges = edges[None]
    return GraphsTuple(nodes=node_feats, edges=edges, receivers=targets, senders=sources,
                       globals=global_feats, n_node=n_node, n_edge=n_edge)


# rotation helpers used by the reference's tests (graph_utils.py:334-364)
class SteerableGraphsTuple(NamedTuple):
    """models/utils/equivariant_graph_utils.py:14-24."""
    nodes: object
    edges: object
    receivers: object
    senders: object
    additional_messages: object
    steerable_node_attrs: object
    steerable_edge_attrs: object
    globals: object
    n_node: object
    n_edge: object


def get_equivariant_graph(node_features, positions, velocities, senders, receivers, n_node, n_edge, edges, globals,
                          lmax_attributes, additional_messages=None, steerable_velocities=False,
                          spherical_harmonics_norm="integral", apply_pbc=None, n_radial_basis=0, r_max=1.0):
    """equivariant_graph_utils.py:27-102 in numpy fp32 (test infrastructure only).

    r_ij = x[senders] - x[receivers] (minimum image if apply_pbc; no EPS, :46-52); edge attributes = normalised
    spherical harmonics up to lmax_attributes (:56-61); node attributes = e3nn.scatter_mean of the edge attributes
    over the receivers, graph by graph (:64-74), plus the harmonics of the velocities if steerable_velocities
    (:76-83); additional_messages = e3nn.bessel(|r_ij|) as n_rb x 0e or |r_ij| as 1x0e (:85-90)."""
    from .so3 import spherical_harmonics
    pos = np.asarray(positions, dtype=np.float32)
    B, N = pos.shape[:2]
    snd, rcv = np.asarray(senders), np.asarray(receivers)
    bi = np.arange(B)[:, None]
    r_ij = (pos[bi, snd] - pos[bi, rcv]).astype(np.float32)
    if apply_pbc is not None:
        r_ij = np.asarray(apply_pbc(r_ij), dtype=np.float32)
    e_attr = spherical_harmonics(lmax_attributes, r_ij.astype(np.float64), True, spherical_harmonics_norm)
    D = e_attr.shape[-1]
    n_attr = np.zeros((B, N, D))
    cnt = np.zeros((B, N))
    for b in range(B):
        np.add.at(n_attr[b], rcv[b], e_attr[b])
        np.add.at(cnt[b], rcv[b], 1.0)
    n_attr = n_attr / np.maximum(cnt, 1.0)[..., None]
    if steerable_velocities:
        n_attr = n_attr + spherical_harmonics(lmax_attributes, np.asarray(velocities, dtype=np.float64), True,
                                              spherical_harmonics_norm)
    d = np.sqrt((r_ij.astype(np.float32) ** 2).sum(-1, dtype=np.float32)).astype(np.float32)
    add = bessel_np(d, n_radial_basis, r_max) if n_radial_basis > 0 else d[..., None]
    return SteerableGraphsTuple(nodes=node_features, edges=edges, receivers=receivers, senders=senders,
                                additional_messages=add, steerable_node_attrs=n_attr, steerable_edge_attrs=e_attr,
                                globals=globals, n_node=n_node, n_edge=n_edge)


def rotation_matrix(angle_deg, axis):
    a_rad = np.radians(angle_deg)
    axis = np.asarray(axis, dtype=np.float64)
    axis = axis / np.linalg.norm(axis)
    a = np.cos(a_rad / 2)
    b, c, d = -axis * np.sin(a_rad / 2)
    return np.array([
        [a * a + b * b - c * c - d * d, 2 * (b * c - a * d), 2 * (b * d + a * c)],
        [2 * (b * c + a * d), a * a + c * c - b * b - d * d, 2 * (c * d - a * b)],
        [2 * (b * d - a * c), 2 * (c * d + a * b), a * a + d * d - b * b - c * c],
    ])


def rotate_representation(data, angle_deg, axis):
    R = rotation_matrix(angle_deg, axis)
    pos, vel, sc = data[:, :3], data[:, 3:6], data[:, 6:]
    return np.concatenate([(R @ pos.T).T, (R @ vel.T).T, sc], axis=1)
