"""Steerable graph attributes on B200 kernels: mirror of ``models/utils/equivariant_graph_utils.py`` (reference).

``SteerableGraphsTuple`` (:14-24) and ``get_equivariant_graph`` (:27-102) with the reference's argument names and
order.  Tensors are torch CUDA tensors with a leading batch axis (the reference indexes ``positions[b, senders[b]]``
itself, :46-47); ``IrrepsArray`` is the (irreps, array) pair of ``eqnn_jax_b200.nequip``.  One C-ABI call
(``eqnn_steerable_attrs``) produces r_ij, the edge harmonics and their receiver-side mean; the radial embedding of
``additional_messages`` is ``eqnn_edge_features`` (the same e3nn.bessel kernel ``build_graph`` uses).
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

from typing import Any, Callable, NamedTuple, Optional

import torch

from . import ops
from .graph_utils import ApplyPBC
from .irreps import Irreps
from .nequip import IrrepsArray

_SH_NORM = {"norm": 0, "component": 1, "integral": 2}


class SteerableGraphsTuple(NamedTuple):
    """Field names and order of the reference's NamedTuple (equivariant_graph_utils.py:14-24)."""
    nodes: Any
    edges: Any
    receivers: Optional[torch.Tensor]
    senders: Optional[torch.Tensor]
    additional_messages: Any
    steerable_node_attrs: Any
    steerable_edge_attrs: Any
    globals: Any
    n_node: Any
    n_edge: Any


def get_equivariant_graph(node_features, positions, velocities, senders, receivers, n_node, n_edge, edges, globals,
                          lmax_attributes: int, additional_messages=None, steerable_velocities: bool = False,
                          spherical_harmonics_norm: str = "integral", apply_pbc: Optional[ApplyPBC] = None,
                          n_radial_basis: int = 0, r_max: float = 1.0) -> SteerableGraphsTuple:
    """equivariant_graph_utils.py:27-102.  positions / velocities: (B, N, 3) tensors or IrrepsArray("1o", ...)."""
    pos = positions.array if isinstance(positions, IrrepsArray) else positions
    vel = velocities.array if isinstance(velocities, IrrepsArray) else velocities
    if not isinstance(pos, torch.Tensor) or not pos.is_cuda:
        raise RuntimeError("get_equivariant_graph: positions must be a CUDA tensor (this path has no CPU implementation)")
    if spherical_harmonics_norm not in _SH_NORM:
        raise ValueError(f"unknown spherical-harmonics normalization {spherical_harmonics_norm!r}")
    if pos.dim() != 3 or pos.shape[-1] != 3:
        raise ValueError("positions must have shape (B, N, 3)")
    B, N, _ = pos.shape
    pos = pos.to(torch.float32).contiguous()
    snd = senders.reshape(B, -1).to(torch.int32).contiguous()
    rcv = receivers.reshape(B, -1).to(torch.int32).contiguous()
    E = snd.shape[1]
    v = None
    if steerable_velocities:
        if vel is None:
            raise ValueError("steerable_velocities=True needs velocities")
        v = vel.to(torch.float32).contiguous().reshape(B * N, 3)
    rowptr, perm, _ = ops.csr_build(snd, rcv, N)
    cell = None if apply_pbc is None else apply_pbc.cell
    cinv = None if apply_pbc is None else apply_pbc.cell_inv
    r_ij, e_attr, n_attr = ops.steerable_attrs(pos.reshape(B * N, 3), 3, v, 3, B, N, E, snd, rcv, rowptr, perm, cell, cinv,
                                               int(lmax_attributes), _SH_NORM[spherical_harmonics_norm])
    add = ops.edge_features(r_ij.view(B, E, 3), int(n_radial_basis), float(r_max))
    attr_irreps = Irreps.spherical_harmonics(int(lmax_attributes))
    ny = (int(lmax_attributes) + 1) ** 2
    return SteerableGraphsTuple(
        nodes=node_features, edges=edges, receivers=receivers, senders=senders,
        additional_messages=IrrepsArray(Irreps(f"{max(int(n_radial_basis), 1)}x0e"), add),
        steerable_node_attrs=IrrepsArray(attr_irreps, n_attr.view(B, N, ny)),
        steerable_edge_attrs=IrrepsArray(attr_irreps, e_attr.view(B, E, ny)),
        globals=globals, n_node=n_node, n_edge=n_edge)
