"""Static layer plans: everything the kernels need that is known at trace time.

* ``TPPlan``      — path/column tables of ``e3nn.tensor_product(x, Y)`` as used at
                    models/nequip.py:52 and :197 (x = node irreps, Y = spherical
                    harmonics up to ``lmax``), restricted to the *live* output
                    columns (those whose irrep survives the following Linear).
* ``LinearPlan``  — ``e3nn.flax.Linear`` (nequip.py:43,85,162,170,197) lowered to
                    one dense matrix  y = x @ Wfull  whose entries are gathered
                    from the Flax parameters (path weight 1/sqrt(fan_in) folded in).

Feature layouts on the device
  "stored"     the reference's own layout (chunks in declaration order);
  "regrouped"  chunks stably sorted in e3nn order and merged — what
               tensor_product / Linear use internally; the gathered sender record
               ``xt`` uses this layout zero-padded to a multiple of 4 floats.
"""
from __future__ import annotations

import dataclasses
from math import sqrt
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import so3
from .irreps import Irreps, irrep_sort_key


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


@dataclasses.dataclass(frozen=True)
class TPColumn:
    """One irrep copy of the regrouped tensor-product output."""
    l1: int
    l2: int
    l3: int
    p3: int
    x_off: int        # offset of the (2 l1+1) input components in the regrouped x
    y_off: int        # offset of Y_{l2} in the concatenated harmonics
    msg_off: int      # offset in the full regrouped message
    col: int          # index among all n_irr columns (== index into the radial MLP output)
    live: int         # index among live columns, or -1
    live_off: int     # component offset among live components, or -1


@dataclasses.dataclass(frozen=True)
class TPPlan:
    irreps_x: Irreps            # regrouped x irreps
    lmax: int
    columns: Tuple[TPColumn, ...]
    irreps_msg: Irreps          # full regrouped output irreps
    irreps_live: Irreps         # regrouped output restricted to live irreps
    live_types: Tuple[Tuple[int, int], ...]

    @property
    def dx(self) -> int:
        return self.irreps_x.dim

    @property
    def dxp(self) -> int:
        return pad4(self.irreps_x.dim)

    @property
    def ny(self) -> int:
        return (self.lmax + 1) ** 2

    @property
    def n_irr(self) -> int:
        return len(self.columns)

    @property
    def n_live(self) -> int:
        return sum(1 for c in self.columns if c.live >= 0)

    @property
    def d_live(self) -> int:
        return self.irreps_live.dim

    @property
    def live_cols(self) -> List[int]:
        return [c.col for c in self.columns if c.live >= 0]

    @property
    def signature(self) -> str:
        live = ",".join(f"{l}{'e' if p > 0 else 'o'}" for l, p in self.live_types)
        return f"x={self.irreps_x};lmax={self.lmax};live={live}"


def make_tp_plan(irreps_x, lmax: int, live_irreps=None) -> TPPlan:
    """``tensor_product(x, Y_{0..lmax})`` with e3nn's defaults (regrouped inputs and
    output, 'component' irrep normalisation).  ``live_irreps`` = irreps kept by
    the consumer (None keeps everything)."""
    xr = Irreps(irreps_x).regroup()
    x_offs = xr.offsets()
    chunks = []  # (key, gen_index, mul, l1, l2, l3, p3, x_off)
    g = 0
    for (m1, l1, p1), xo in zip(xr, x_offs):
        for l2 in range(lmax + 1):
            p2 = (-1) ** l2
            for l3 in range(abs(l1 - l2), l1 + l2 + 1):
                chunks.append((irrep_sort_key(l3, p1 * p2), g, m1, l1, l2, l3, p1 * p2, xo))
                g += 1
    chunks.sort(key=lambda c: (c[0], c[1]))
    live_set = None if live_irreps is None else {(l, p) for _, l, p in Irreps(live_irreps)}
    cols: List[TPColumn] = []
    msg_off = live_off = n_live = 0
    msg_items, live_items = [], []
    for _, _, mul, l1, l2, l3, p3, xo in chunks:
        is_live = live_set is None or (l3, p3) in live_set
        msg_items.append((mul, l3, p3))
        if is_live:
            live_items.append((mul, l3, p3))
        for u in range(mul):
            cols.append(TPColumn(l1, l2, l3, p3, xo + u * (2 * l1 + 1), l2 * l2, msg_off, len(cols),
                                 n_live if is_live else -1, live_off if is_live else -1))
            msg_off += 2 * l3 + 1
            if is_live:
                n_live += 1
                live_off += 2 * l3 + 1
    irreps_live = Irreps(live_items).simplify()
    live_types = tuple(sorted({(l, p) for _, l, p in irreps_live}, key=lambda t: irrep_sort_key(*t)))
    return TPPlan(xr, lmax, tuple(cols), Irreps(msg_items).simplify(), irreps_live, live_types)


# ----------------------------------------------------------------------------
# Linear -> dense matrix
# ----------------------------------------------------------------------------
@dataclasses.dataclass
class LinearPlan:
    irreps_in: Irreps            # stored input irreps
    irreps_out: Irreps           # stored (unsimplified) output irreps
    param_shapes: Dict[str, Tuple[int, int]]
    # dense matrix (rows = stored input comps [+pad], cols = stored output comps [+pad])
    rows: int
    cols: int
    # gather table: for every non-zero dense entry (row, col) -> (param name, flat index, scale)
    entries: List[Tuple[int, int, str, int, float]]


def make_linear_plan(irreps_in, irreps_out, force_irreps_out: bool = False,
                     in_layout: str = "stored", out_layout: str = "stored",
                     rows_pad: Optional[int] = None, cols_pad: Optional[int] = None,
                     col_scale: Optional[Sequence[float]] = None) -> LinearPlan:
    """e3nn.flax.Linear with library defaults (path_normalization 'element',
    gradient_normalization 'path': forward multiplier 1/sqrt(fan_in), N(0,1) init).

    in_layout / out_layout choose whether the dense matrix addresses the stored
    layout of the operand or its regrouped layout (used for the padded sender
    record).  ``col_scale`` multiplies dense column j (used to fold constants)."""
    irreps_in = Irreps(irreps_in)
    irreps_out_full = Irreps(irreps_out)
    xin = irreps_in.regroup()
    out_unsimpl = irreps_out_full if force_irreps_out else irreps_out_full.filter(keep=xin)
    out_simpl = out_unsimpl.simplify()
    # component index maps
    in_map = irreps_in.regroup_index_map() if in_layout == "stored" else list(range(xin.dim))
    if out_layout == "stored":
        out_map = list(range(out_unsimpl.dim))          # rechunk keeps memory order
    else:
        om = out_unsimpl.regroup_index_map()            # regrouped comp -> stored offset
        out_map = [0] * len(om)
        for r, s in enumerate(om):
            out_map[s] = r                              # stored offset -> regrouped comp
    rows = rows_pad if rows_pad is not None else irreps_in.dim
    cols = cols_pad if cols_pad is not None else out_unsimpl.dim
    shapes: Dict[str, Tuple[int, int]] = {}
    entries: List[Tuple[int, int, str, int, float]] = []
    in_offs, out_offs = xin.offsets(), out_simpl.offsets()
    for i_out, (mo, lo, po) in enumerate(out_simpl):
        fan_in = sum(mi for mi, li, pi in xin if (li, pi) == (lo, po))
        for i_in, (mi, li, pi) in enumerate(xin):
            if (li, pi) != (lo, po):
                continue
            name = f"w[{i_in},{i_out}] {Irreps([(mi, li, pi)])},{Irreps([(mo, lo, po)])}"
            shapes[name] = (mi, mo)
            pw = 1.0 / sqrt(fan_in)
            d = 2 * lo + 1
            for u in range(mi):
                for w in range(mo):
                    for i in range(d):
                        r = in_map[in_offs[i_in] + u * d + i]
                        c = out_map[out_offs[i_out] + w * d + i]
                        sc = pw * (col_scale[c] if col_scale is not None else 1.0)
                        entries.append((r, c, name, u * mo + w, sc))
    return LinearPlan(irreps_in, out_unsimpl, shapes, rows, cols, entries)
