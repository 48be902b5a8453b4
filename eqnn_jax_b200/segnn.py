"""SEGNN on B200 kernels, behind the reference's module interface (models/segnn.py).

Same dataclass fields and defaults (:152-165), a ``SteerableGraphsTuple`` (``equivariant_graph_utils``) in,
``GraphsTuple`` with ``IrrepsArray`` nodes (task="node") or ``(B, n_outputs)`` (task="graph") out, Flax parameter
names (``TensorProductLinearGate_i/Linear_0/w[i,j] ...``, ``Dense_0``, ``MLP_0``).

Every layer of the model is a ``TensorProductLinearGate`` (:21-64): ``gate(Linear(TP(x, y)))`` with y the steerable
attributes (spherical harmonics, one copy per degree).  ``Linear(TP(x, y))`` is bilinear, so it is lowered at plan
time into one dense matrix per attribute component, ``M_cat`` (rows: components of x in their
stored order; columns: Linear output x attribute component, attribute fastest), whose entries are Clebsch-Gordan coefficients times
single Linear weights (``eqnn_weights_expand`` gathers them from the flat parameter buffer; the transposed table
contracts dense gradients back).  Per layer and row block:
    Xbig = x @ M_cat                      eqnn_gemm_f32                              segnn.py:54  (TP + Linear)
    z    = sum_j y_j Xbig[:, :, j] + b    eqnn_segnn_ycontract                        segnn.py:48-54 (biases on 0e)
    out  = e3nn.gate(z)                   eqnn_gate_fwd                              segnn.py:55-63
Edges are processed in receiver-sorted order: ``eqnn_egnn_edge_input_fwd`` builds concat(additional_messages, h_s,
h_r) (:112), ``eqnn_egnn_segment_rows_fwd`` is jraph's segment_sum / segment_mean over receivers (:218), and the
backward passes gather sender / receiver ends through the two CSRs (no atomics).  The first n_layers-1 layers of
the node function never reach the output (:84-89): they own parameters (exact zero gradients) and are not
computed.  The layer GEMMs run on the tcgen05 tensor cores with 3xTF32 operand splitting (csrc/tc_gemm.cu, any
strides / alignment, K split over CTAs for the weight gradients); ``tensor_cores=False`` keeps them on the fp32
CUDA-core kernel.  ``lowering="blocked"`` is a second, irrep-blocked route (multiplicity GEMMs per stored irrep
block + a Clebsch-Gordan table kernel, 4x fewer FLOPs but overhead-bound K <= 89 GEMMs).
Two of the three edge layers of a message-passing step do not run per edge: the first one takes
concat(add_e, h_sender, h_receiver), so its wide products are formed per node and gathered (``_edge0_fwd``); the last
one has no gate and is only aggregated over the receivers, so the aggregation is pulled in front of it and its GEMMs
run on the node rows (``_edgeL_fwd``).  With attributes l <= 1 the forward GEMM of the remaining layers contracts the
attribute axis in its epilogue (``eqnn_segnn_tplg_fwd``).
Graph globals (concatenated to the pooled scalars, LayerNorm, :282-286) are supported.
Scalar activations: gelu and silu (models/segnn.py:27,56).  Aggregations: sum, mean, max (jraph.segment_max; the node-level evaluation of the last edge layer needs a linear
aggregation and is skipped for max).
"""
from __future__ import annotations

import dataclasses
import math
from typing import Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from . import so3
from .equivariant_graph_utils import SteerableGraphsTuple
from .graph_utils import GraphsTuple
from .irreps import Irreps, balanced_irreps
from .nequip import IrrepsArray, NequIPParams as FlatParams, _mm, normalize_constant

ROW_BLOCK = 1 << 18      # rows per Xbig block: bounds the (rows, Dy*Do) temporaries (0.8 GB each at Dy*Do = 768)


# --------------------------------------------------------------------------------------------------
# static plan of one TensorProductLinearGate
# --------------------------------------------------------------------------------------------------
class _TPLG:
    def __init__(self, owner: "_Plan", name: str, x_irreps: Irreps, y_lmax: Optional[int], out_irreps: Irreps,
                 activation: bool, live: bool = True):
        x_irreps, out_irreps = Irreps(x_irreps), Irreps(out_irreps)
        self.name, self.activation, self.live = name, activation, live
        self.Dx = x_irreps.dim
        y_irreps = Irreps.spherical_harmonics(y_lmax) if y_lmax is not None else Irreps("1x0e")
        self.has_y = y_lmax is not None
        self.Dy = y_irreps.dim
        # Linear output irreps: gate scalars are prepended, then regrouped (:39-46)
        lin = out_irreps
        if activation:
            n_gate = out_irreps.num_irreps - out_irreps.count("1x0e")
            lin = ((Irreps([(n_gate, 0, 1)]) + out_irreps) if n_gate else out_irreps).regroup()
        xin = x_irreps.regroup()
        in_map = x_irreps.regroup_index_map()                     # regrouped component -> stored column of x
        x_off = xin.offsets()
        # tensor-product paths in e3nn's enumeration order and the regrouped layout of its output
        paths = []
        for i1, (m1, l1, p1) in enumerate(xin):
            for (_, l2, p2) in y_irreps:
                for l3 in range(abs(l1 - l2), l1 + l2 + 1):
                    paths.append((i1, l2, l3, p1 * p2, m1))
        tp_unsorted = Irreps([(m1, l3, p3) for (_, _, l3, p3, m1) in paths])
        tp_out = tp_unsorted.regroup()
        lin_f = lin.filter(keep=tp_out)
        if lin_f != lin:
            raise NotImplementedError(f"{name}: the Linear drops output irreps absent from TP(x, y) ({lin} -> {lin_f})")
        lin = lin.simplify()
        self.Do = lin.dim
        out_off = lin.offsets()
        chunk_of = {(l, p): i for i, (_, l, p) in enumerate(tp_out)}        # regrouped: one chunk per irrep type
        row_in_chunk: Dict[Tuple[int, int], int] = {}
        # parameters (names of e3nn.flax.Linear) -- created whether or not the layer is computed
        self.w_off: Dict[int, int] = {}
        for i_out, (mo, lo, po) in enumerate(lin):
            i_in = chunk_of[(lo, po)]
            mi = tp_out[i_in][0]
            wname = f"w[{i_in},{i_out}] {Irreps([(mi, lo, po)])},{Irreps([(mo, lo, po)])}"
            self.w_off[i_out] = owner.add_param((name, "Linear_0", wname), (mi, mo))
        self.n_bias, self.b_off, self.b_col = 0, None, 0
        for i_out, (mo, lo, po) in enumerate(lin):
            if (lo, po) == (0, 1):
                if self.n_bias:
                    raise NotImplementedError("more than one 0e block in the Linear output")
                self.b_off = owner.add_param((name, "Linear_0", f"b[{i_out}] {Irreps([(mo, lo, po)])}"), (mo,))
                self.n_bias, self.b_col = mo, out_off[i_out]
        self.init_std = {self.w_off[i]: 1.0 / math.sqrt(tp_out[chunk_of[(lo, po)]][0]) for i, (_, lo, po) in enumerate(lin)}
        # gate layout (e3nn.gate): leading scalars activated, last num_gated scalars gate the l>0 irreps
        self.out_irreps = out_irreps
        if activation:
            gated = [(m, l, p) for m, l, p in lin if l > 0]
            self.n_gates = sum(m for m, _, _ in gated)
            self.n_extra = sum(m for m, l, _ in lin if l == 0) - self.n_gates
            self.gate_mul = [m for m, _, _ in gated]
            self.gate_l = [l for _, l, _ in gated]
            self.d_g = self.n_extra + sum(m * (2 * l + 1) for m, l, _ in gated)
            self.out_irreps = Irreps([(self.n_extra, 0, 1)] + gated) if self.n_extra else Irreps(gated)
        self.D_out = self.d_g if activation else self.Do
        if not live:
            return
        if owner.lowering == "blocked":
            self._lower_blocked(owner, x_irreps, xin, in_map, x_off, paths, lin, out_off)
            return
        # dense lowering: M_cat[x column, o * Dy + j] (the Dy attribute components of one output are adjacent, so a
        # GEMM tile holds whole outputs and the attribute contraction can run in its epilogue)
        self.M_off = owner.new_dense(self.Dx * self.Dy * self.Do)
        ncols = self.Dy * self.Do
        for (i1, l2, l3, p3, m1) in paths:
            key = (l3, p3)
            r0 = row_in_chunk.get(key, 0)
            row_in_chunk[key] = r0 + m1
            i_out = next((i for i, (_, lo, po) in enumerate(lin) if (lo, po) == key), None)
            if i_out is None:
                continue                                          # pseudo-irreps (1e, 2o, ...) are dropped by the Linear
            mo = lin[i_out][0]
            l1 = xin[i1][1]
            cg = so3.real_cg(l1, l2, l3) * math.sqrt(2 * l3 + 1)  # e3nn "component" normalisation of the product
            a, b, c = np.nonzero(cg)
            if a.size == 0:
                continue
            u = np.arange(m1)[:, None, None]
            w = np.arange(mo)[None, :, None]
            t = np.arange(a.size)[None, None, :]
            rows = np.asarray(in_map)[x_off[i1] + u * (2 * l1 + 1) + a[t]]
            cols = (out_off[i_out] + w * (2 * l3 + 1) + c[t]) * self.Dy + (l2 * l2 + b[t])
            dst = self.M_off + rows * ncols + cols
            par = self.w_off[i_out] + (r0 + u) * mo + w + 0 * t
            val = cg[a, b, c][t] + 0.0 * (u + w)
            owner.add_entries(np.broadcast_to(dst, val.shape).ravel(), np.broadcast_to(par, val.shape).ravel(), val.ravel())


    def _lower_blocked(self, owner, x_irreps, xin, in_map, x_off, paths, lin, out_off):
        """Irrep-blocked lowering: the multiplicity contraction first, T[(path, w, m1)] = sum_u x[(u, m1)] W_path[u, w]
        (one GEMM per run of stored x columns of one irrep type; ``runs``), then the Clebsch-Gordan terms
        z[(w, m3)] += sqrt(2 l3 + 1) C[m1, m2, m3] y[m2] T[(path, w, m1)] from CSR tables (``cg_f`` by Linear output
        for the forward, ``cg_b`` by T column for the backward)."""
        row_in_chunk: Dict[Tuple[int, int], int] = {}
        kept, DT = [], 0
        for (i1, l2, l3, p3, m1) in paths:
            key = (l3, p3)
            r0 = row_in_chunk.get(key, 0)
            row_in_chunk[key] = r0 + m1
            i_out = next((i for i, (_, lo, po) in enumerate(lin) if (lo, po) == key), None)
            if i_out is None:
                continue
            l1 = xin[i1][1]
            cg = so3.real_cg(l1, l2, l3) * math.sqrt(2 * l3 + 1)
            if not np.any(cg):
                continue
            mo = lin[i_out][0]
            kept.append(dict(i1=i1, l1=l1, l2=l2, l3=l3, i_out=i_out, r0=r0, mo=mo, toff=DT, cg=cg))
            DT += mo * (2 * l1 + 1)
        self.DT = DT
        t_lo = {i1: min(k["toff"] for k in kept if k["i1"] == i1) for i1 in {k["i1"] for k in kept}}
        t_hi = {i1: max(k["toff"] + k["mo"] * (2 * k["l1"] + 1) for k in kept if k["i1"] == i1) for i1 in t_lo}
        # runs of stored x columns of one irrep type
        inv = np.empty((self.Dx,), dtype=np.int64)
        inv[np.asarray(in_map, dtype=np.int64)] = np.arange(self.Dx)
        type_of = {(l, p): i for i, (_, l, p) in enumerate(xin)}
        stored, c = [], 0
        for (m, l, p) in x_irreps:
            w = m * (2 * l + 1)
            if w:
                if stored and stored[-1][0] == (l, p) and stored[-1][1] + stored[-1][2] == c:
                    stored[-1][2] += w
                else:
                    stored.append([(l, p), c, w])
            c += w
        self.runs, seen, covered = [], set(), 0
        for (l, p), c0, K in stored:
            i1 = type_of[(l, p)]
            if i1 not in t_lo:
                continue                                          # no path from this irrep type reaches the Linear
            d1 = 2 * l + 1
            Nr = t_hi[i1] - t_lo[i1]
            W_off = owner.new_dense(K * Nr)
            comp = inv[c0:c0 + K] - x_off[i1]
            u, m1 = comp // d1, comp % d1
            for k in (k for k in kept if k["i1"] == i1):
                wv = np.arange(k["mo"])[None, :]
                dst = W_off + np.arange(K)[:, None] * Nr + (k["toff"] - t_lo[i1]) + wv * d1 + m1[:, None]
                par = self.w_off[k["i_out"]] + (k["r0"] + u[:, None]) * k["mo"] + wv
                owner.add_entries(dst.ravel(), par.ravel(), np.ones(dst.size))
            self.runs.append((c0, K, t_lo[i1], Nr, W_off, i1 in seen))      # last: accumulate into T
            seen.add(i1)
            covered += K
        self.dx_full = covered == self.Dx                         # else: uncovered dx columns are exact zeros
        # Clebsch-Gordan terms
        o_l, t_l, j_l, c_l = [], [], [], []
        for k in kept:
            a, b, cc = np.nonzero(k["cg"])
            d1, d3 = 2 * k["l1"] + 1, 2 * k["l3"] + 1
            wv = np.arange(k["mo"])[:, None]
            o_l.append((out_off[k["i_out"]] + wv * d3 + cc[None, :]).ravel())
            t_l.append((k["toff"] + wv * d1 + a[None, :]).ravel())
            j_l.append(np.broadcast_to(k["l2"] ** 2 + b[None, :], (k["mo"], a.size)).ravel())
            c_l.append(np.broadcast_to(k["cg"][a, b, cc][None, :], (k["mo"], a.size)).ravel())
        o_a, t_a, j_a, c_a = (np.concatenate(v) for v in (o_l, t_l, j_l, c_l))
        self.cg_f = owner.add_cg(o_a, t_a, j_a, c_a, self.Do)     # rows: Linear outputs, sources: T columns
        self.cg_b = owner.add_cg(t_a, o_a, j_a, c_a, DT)          # rows: T columns, sources: Linear outputs


class _Plan:
    def __init__(self, cfg: "SEGNN", irreps_in: Irreps, lmax_attr: int, n_add: int, n_globals: int = 0):
        self.n_globals = int(n_globals)
        self.lowering = cfg.lowering
        if self.lowering not in ("blocked", "dense"):
            raise ValueError(f"lowering must be 'blocked' or 'dense', not {cfg.lowering!r}")
        self._cg_ptr: List[np.ndarray] = []
        self._cg_idx: List[np.ndarray] = []
        self._cg_coef: List[np.ndarray] = []
        self._n_cg_ptr = 0
        self._n_cg = 0
        self.layout: List[Tuple[Tuple[str, ...], Tuple[int, ...], int]] = []
        self._n_params = 0
        self._n_exp = 0
        self._dst: List[np.ndarray] = []
        self._par: List[np.ndarray] = []
        self._val: List[np.ndarray] = []
        if cfg.scalar_activation not in ("gelu", "silu"):
            raise NotImplementedError("scalar_activation must be 'gelu' or 'silu' on the B200 path (models/segnn.py:27,56)")
        if cfg.message_passing_agg not in ("sum", "mean", "max") or (cfg.task == "graph" and
                                                                      cfg.readout_agg not in ("sum", "mean", "max")):
            raise ValueError("aggregations must be 'sum', 'mean' or 'max' (models/segnn.py:155,218,265-270)")
        hidden = Irreps(cfg.hidden_irreps) if cfg.hidden_irreps is not None else \
            balanced_irreps(cfg.l_max_hidden, cfg.d_hidden, True)
        self.hidden, self.irreps_in = hidden, irreps_in
        self.out_irreps = Irreps(cfg.output_irreps) if cfg.output_irreps is not None else irreps_in
        self.act = cfg.scalar_activation
        self.inv_c_act = 1.0 / normalize_constant(cfg.scalar_activation)
        self.inv_c_gate = 1.0 / normalize_constant("sigmoid")
        ti = 0

        self.tplgs: List[_TPLG] = []

        def tplg(x_irreps, y_lmax, o_irreps, act, live=True):
            nonlocal ti
            t = _TPLG(self, f"TensorProductLinearGate_{ti}", x_irreps, y_lmax, o_irreps, act, live)
            ti += 1
            self.tplgs.append(t)
            return t

        self.embed = tplg(irreps_in, lmax_attr, hidden, False)                      # :167-177
        self.steps = []
        add_irreps = Irreps([(n_add, 0, 1)])
        h_irreps = self.embed.out_irreps
        for _ in range(cfg.message_passing_steps):
            st = {"edge": [], "h_irreps": h_irreps}
            m_irreps = add_irreps + h_irreps + h_irreps                              # :111-113
            for _ in range(cfg.n_layers - 1):
                st["edge"].append(tplg(m_irreps, lmax_attr, hidden, True))
                m_irreps = st["edge"][-1].out_irreps
            st["edge"].append(tplg(m_irreps, lmax_attr, hidden, False))
            n_in = h_irreps + st["edge"][-1].out_irreps                              # :79-81
            for _ in range(cfg.n_layers - 1):
                tplg(n_in, lmax_attr, hidden, True, live=False)                      # discarded (:84-89)
            st["node"] = tplg(n_in, lmax_attr, hidden, False)
            if cfg.residual and st["node"].out_irreps != h_irreps:
                raise NotImplementedError("residual connection needs matching irreps")
            h_irreps = st["node"].out_irreps
            self.steps.append(st)
        self.h_irreps = h_irreps
        if cfg.task == "node":                                                       # _decode :179-196
            self.decode = []
            x_irreps = h_irreps
            for _ in range(cfg.n_layers):
                self.decode.append(tplg(x_irreps, None, hidden, True))
                x_irreps = self.decode[-1].out_irreps
            self.decode.append(tplg(x_irreps, lmax_attr, self.out_irreps, False))
        elif cfg.task == "graph":                                                    # :265-295
            d = cfg.d_hidden
            self.pre = tplg(h_irreps, lmax_attr, Irreps([(d, 0, 1)]), False)
            self.dense = (self.add_param(("Dense_0", "kernel"), (d, d)), self.add_param(("Dense_0", "bias"), (d,)))
            n_pool = d + self.n_globals                                              # globals + LayerNorm, :282-286
            self.ln = None
            if self.n_globals > 0:
                self.ln = (self.add_param(("LayerNorm_0", "scale"), (n_pool,)),
                           self.add_param(("LayerNorm_0", "bias"), (n_pool,)))
            ws = [cfg.mlp_readout_widths[0] * n_pool] + [w * d for w in cfg.mlp_readout_widths[1:]] + [cfg.n_outputs]
            self.ro_mlp, fan = [], n_pool
            for l, w in enumerate(ws):
                self.ro_mlp.append((self.add_param(("MLP_0", f"Dense_{l}", "kernel"), (fan, w)),
                                    self.add_param(("MLP_0", f"Dense_{l}", "bias"), (w,)), fan, w))
                fan = w
        else:
            raise ValueError(f"Invalid task {cfg.task}")
        self._finalize()

    # ---- bookkeeping --------------------------------------------------------------------------
    def add_param(self, path, shape) -> int:
        off = self._n_params
        self.layout.append((tuple(path), tuple(shape), off))
        self._n_params += int(np.prod(shape))
        return off

    def new_dense(self, n: int) -> int:
        off = self._n_exp
        self._n_exp += n
        return off

    def add_cg(self, rows, src, j, coef, n_rows: int) -> int:
        """Append one CSR table (sorted by row, stable) to the plan's Clebsch-Gordan tables; returns its ptr offset."""
        order = np.argsort(rows, kind="stable")
        counts = np.bincount(rows, minlength=n_rows)
        ptr = self._n_cg + np.concatenate([[0], np.cumsum(counts)])
        off = self._n_cg_ptr
        self._cg_ptr.append(ptr.astype(np.int32))
        self._cg_idx.append((src[order].astype(np.int64) | (j[order].astype(np.int64) << 24)).astype(np.int32))
        self._cg_coef.append(coef[order].astype(np.float32))
        self._n_cg_ptr += n_rows + 1
        self._n_cg += rows.size
        return off

    def add_entries(self, dst, par, val):
        self._dst.append(np.asarray(dst, dtype=np.int64))
        self._par.append(np.asarray(par, dtype=np.int64))
        self._val.append(np.asarray(val, dtype=np.float64))

    def _finalize(self):
        dst = np.concatenate(self._dst) if self._dst else np.zeros((0,), np.int64)
        par = np.concatenate(self._par) if self._par else np.zeros((0,), np.int64)
        val = np.concatenate(self._val) if self._val else np.zeros((0,), np.float64)
        if np.unique(dst).size != dst.size:
            raise AssertionError("dense lowering: an entry received two sources")
        idx = np.full((self._n_exp,), -1, dtype=np.int32)
        scale = np.zeros((self._n_exp,), dtype=np.float32)
        idx[dst] = par
        scale[dst] = val
        order = np.argsort(par, kind="stable")
        ps = par[order]
        pidx, start = np.unique(ps, return_index=True)
        ptr = np.concatenate([start, [ps.size]]).astype(np.int32)
        self._np = dict(exp_idx=idx, exp_scale=scale, con_pidx=pidx.astype(np.int32), con_ptr=ptr,
                        con_tidx=dst[order].astype(np.int32), con_scale=val[order].astype(np.float32))
        cat = lambda parts, dt: np.concatenate(parts).astype(dt) if parts else np.zeros((1,), dt)
        self._np.update(cg_ptr=cat(self._cg_ptr, np.int32), cg_idx=cat(self._cg_idx, np.int32),
                        cg_coef=cat(self._cg_coef, np.float32))
        self._dev: Dict[str, Dict[str, torch.Tensor]] = {}
        del self._dst, self._par, self._val, self._cg_ptr, self._cg_idx, self._cg_coef

    def device_tables(self, device):
        key = str(device)
        if key not in self._dev:
            self._dev[key] = {k: torch.from_numpy(v).to(device) for k, v in self._np.items()}
        return self._dev[key]


# --------------------------------------------------------------------------------------------------
# the module
# --------------------------------------------------------------------------------------------------
@dataclasses.dataclass
class SEGNN:
    """Field names and defaults of the reference module (models/segnn.py:152-165)."""
    d_hidden: int = 64
    n_layers: int = 3
    message_passing_steps: int = 3
    message_passing_agg: str = "sum"
    scalar_activation: str = "gelu"
    gate_activation: str = "sigmoid"
    task: str = "graph"
    n_outputs: int = 1
    output_irreps: Optional[Union[str, Irreps]] = None
    readout_agg: str = "mean"
    mlp_readout_widths: Sequence[int] = (4, 2, 2)
    l_max_hidden: int = 1
    hidden_irreps: Optional[Union[str, Irreps]] = None
    residual: bool = True
    # not a reference field: "dense" = one GEMM on M_cat per layer; "blocked" = irrep-blocked lowering (multiplicity
    # GEMMs per stored irrep block + Clebsch-Gordan table kernel): 4x fewer FLOPs, but its K <= 89 GEMMs run at
    # 4-10 TFLOP/s on the CUDA-core kernel, so it is 14 % slower end to end today
    lowering: str = "dense"
    # not a reference field: run the layer GEMMs on tcgen05 tensor cores with 3xTF32 operand splitting
    # (fp32-class accuracy); False = fp32 FMA on the CUDA cores
    tensor_cores: bool = True

    def __post_init__(self):
        self._plans: Dict = {}
        self._gemm_kw = dict(tag="segnn_gemm", tc_any=bool(self.tensor_cores))

    # ---- plan / parameters ----------------------------------------------------------------------
    def _unpack(self, g: SteerableGraphsTuple):
        for f in ("nodes", "steerable_node_attrs", "steerable_edge_attrs", "additional_messages"):
            if not isinstance(getattr(g, f), IrrepsArray):
                raise TypeError(f"SEGNN expects st_graphs.{f} to be an IrrepsArray (reference: e3nn.IrrepsArray)")
        nodes = g.nodes.array
        if not nodes.is_cuda:
            raise RuntimeError("SEGNN: expected CUDA tensors (this path has no CPU implementation)")
        lmax_attr = int(round(math.sqrt(g.steerable_edge_attrs.array.shape[-1]))) - 1
        if Irreps(g.steerable_edge_attrs.irreps) != Irreps.spherical_harmonics(lmax_attr):
            raise NotImplementedError("steerable attributes must be spherical harmonics 0..lmax (get_equivariant_graph)")
        n_glob = 0 if (g.globals is None or self.task != "graph") else int(g.globals.shape[-1])
        return Irreps(g.nodes.irreps), lmax_attr, int(g.additional_messages.array.shape[-1]), n_glob

    def plan(self, irreps_in, lmax_attr: int, n_add: int, n_globals: int = 0) -> _Plan:
        key = (Irreps(irreps_in), lmax_attr, n_add, n_globals)
        if key not in self._plans:
            self._plans[key] = _Plan(self, Irreps(irreps_in), lmax_attr, n_add, n_globals)
        return self._plans[key]

    def init(self, rng: Union[int, np.random.Generator], st_graphs: SteerableGraphsTuple) -> FlatParams:
        """Flax-style ``model.init`` with the library initialisers' distributions: e3nn Linear with
        gradient_normalization='element' N(0, 1/fan_in), zero biases; nn.Dense lecun-normal; models/mlp.py xavier-uniform."""
        plan = self.plan(*self._unpack(st_graphs))
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        host = np.zeros((plan._n_params,), dtype=np.float32)
        stds = {}
        for t in plan.tplgs:
            stds.update(t.init_std)
        for path, shape, off in plan.layout:
            n = int(np.prod(shape))
            if path[0] == "LayerNorm_0":
                host[off:off + n] = 1.0 if path[-1] == "scale" else 0.0
                continue
            if path[-1] == "bias" or path[-1].startswith("b["):
                continue
            if path[0].startswith("TensorProductLinearGate"):
                v = rng.normal(size=shape) * stds[off]
            elif path[0] == "Dense_0":
                v = rng.normal(size=shape) / math.sqrt(shape[0])
            else:
                lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                v = rng.uniform(-lim, lim, size=shape)
            host[off:off + n] = v.reshape(-1)
        return FlatParams(plan.layout, torch.from_numpy(host).to(st_graphs.nodes.array.device))

    # ---- forward / backward ------------------------------------------------------------------------
    def _prepare(self, g: SteerableGraphsTuple):
        irreps_in, lmax_attr, n_add, n_glob = self._unpack(g)
        nodes = g.nodes.array
        if nodes.dim() == 2:
            raise ValueError("SEGNN on B200 expects a leading batch axis (B, N, D)")
        B, N, D = nodes.shape
        snd = g.senders.reshape(B, -1).to(torch.int32).contiguous()
        rcv = g.receivers.reshape(B, -1).to(torch.int32).contiguous()
        E = snd.shape[1]
        rowptr, perm, src = ops.csr_build(snd, rcv, N)
        rowptr_s, perm_s, _ = ops.csr_build(rcv, snd, N)
        spos = ops.perm_compose(perm, perm_s)             # sender-sorted position -> receiver-sorted position
        ny = (lmax_attr + 1) ** 2
        f32c = lambda t, w: t.to(torch.float32).contiguous().reshape(-1, w)
        y_node = f32c(g.steerable_node_attrs.array, ny)
        y_edge = ops.gather_rows(f32c(g.steerable_edge_attrs.array, ny), perm, ny)       # receiver-sorted order
        add = ops.gather_rows(f32c(g.additional_messages.array, n_add), perm, n_add)
        plan = self.plan(irreps_in, lmax_attr, n_add, n_glob)
        glob = None
        if n_glob:
            if not isinstance(g.globals, torch.Tensor) or not g.globals.is_cuda:
                raise RuntimeError("st_graphs.globals must be a CUDA tensor (this path has no CPU implementation)")
            glob = g.globals.reshape(B, n_glob).to(torch.float32).contiguous()
        return plan, f32c(nodes, D), B, N, E, (rowptr, perm, src, rowptr_s, perm_s, spos), y_node, y_edge, add, glob

    def forward_backward(self, params: FlatParams, st_graphs: SteerableGraphsTuple, grad_fn=None):
        """Forward (and, if ``grad_fn(out) -> dL/dout`` is given, backward into ``params.grad``).  Returns the raw
        output tensor: (B, N, D_out) for task='node', (B, n_outputs) for task='graph'."""
        prep = self._prepare(st_graphs)
        out, saved = self._forward(params.flat, *prep)
        if grad_fn is not None:
            self._backward(params.flat, params.grad, saved, grad_fn(out).contiguous(), *prep)
        return out

    def apply(self, params: FlatParams, st_graphs: SteerableGraphsTuple):
        out = _SEGNNFunction.apply(params.flat, self, st_graphs)
        if self.task == "graph":
            return out
        plan = self.plan(*self._unpack(st_graphs))
        return GraphsTuple(nodes=IrrepsArray(plan.out_irreps, out), edges=st_graphs.edges, receivers=st_graphs.receivers,
                           senders=st_graphs.senders, globals=st_graphs.globals, n_node=st_graphs.n_node,
                           n_edge=st_graphs.n_edge)

    __call__ = apply

    # -- one TensorProductLinearGate ------------------------------------------------------------------
    def _tplg_fwd(self, plan, t: _TPLG, theta, wexp, x, ldx, y, R):
        dev = theta.device
        f32 = dict(dtype=torch.float32, device=dev)
        z = torch.empty((R, t.Do), **f32)
        blk = min(R, ROW_BLOCK)
        if plan.lowering == "blocked":
            tab = plan.device_tables(dev)
            T = torch.empty((blk, t.DT), **f32)
            for r0 in range(0, R, blk):
                n = min(blk, R - r0)
                for (c0, K, tb, Nr, W_off, acc) in t.runs:
                    _mm(n, Nr, K, 1.0, x, r0 * ldx + c0, ldx, wexp, W_off, Nr, T, tb, t.DT, accumulate=acc, **self._gemm_kw)
                ops.segnn_cg_apply(T, t.DT, t.DT, y if t.has_y else None, t.Dy, t.Dy, tab["cg_ptr"], tab["cg_idx"],
                                   tab["cg_coef"], theta if t.n_bias else None, t.b_off or 0, t.b_col, t.n_bias, n, t.Do,
                                   z, t.Do, y_off=r0 * t.Dy, out_off=r0 * t.Do, ptr_off=t.cg_f)
        elif self.tensor_cores and t.has_y and ops.segnn_fwd_fused_ok(R, t.Dx, t.Dy, t.Do):
            # attribute contraction in the GEMM epilogue: x @ M_cat never leaves tensor memory, no row blocks
            ops.segnn_tplg_fwd(x, ldx, t.Dx, wexp, t.M_off, y, t.Dy, t.Do, theta if t.n_bias else None, t.b_off or 0,
                               t.b_col, t.n_bias, R, z)
        else:
            nc = t.Dy * t.Do
            Xbig = torch.empty((blk, nc), **f32)
            for r0 in range(0, R, blk):
                n = min(blk, R - r0)
                _mm(n, nc, t.Dx, 1.0, x, r0 * ldx, ldx, wexp, t.M_off, nc, Xbig, 0, nc, **self._gemm_kw)
                ops.segnn_ycontract(Xbig, y if t.has_y else None, t.Dy, theta if t.n_bias else None, t.b_off or 0, t.b_col,
                                    t.n_bias, n, t.Dy, t.Do, z, y_off=r0 * t.Dy, z_off=r0 * t.Do)
        if not t.activation:
            return z, (x, ldx, y, z)
        g = torch.empty((R, t.d_g), **f32)
        ops.gate_fwd(z, R, t.n_extra, t.n_gates, t.gate_mul, t.gate_l, plan.inv_c_act, plan.inv_c_gate, g, plan.act)
        return g, (x, ldx, y, z)

    def _tplg_bwd(self, plan, t: _TPLG, theta, gtheta, wexp, gwexp, sv, dout, R, need_dx=True):
        x, ldx, y, z = sv
        dev = theta.device
        f32 = dict(dtype=torch.float32, device=dev)
        dz = dout
        if t.activation:
            dz = torch.empty((R, t.Do), **f32)
            ops.gate_bwd(z, dout, R, t.n_extra, t.n_gates, t.gate_mul, t.gate_l, plan.inv_c_act, plan.inv_c_gate, dz, plan.act)
        if t.n_bias:
            tmp = torch.empty((t.Do,), **f32)
            ops.colsum_tall(dz, R, t.Do, tmp)
            gtheta[t.b_off:t.b_off + t.n_bias].copy_(tmp[t.b_col:t.b_col + t.n_bias])
        blk = min(R, ROW_BLOCK)
        if plan.lowering == "blocked":
            tab = plan.device_tables(dev)
            dT = torch.empty((blk, t.DT), **f32)
            dx = None
            if need_dx:
                dx = torch.empty((R, t.Dx), **f32) if t.dx_full else torch.zeros((R, t.Dx), **f32)
            for r0 in range(0, R, blk):
                n = min(blk, R - r0)
                ops.segnn_cg_apply(dz, t.Do, t.Do, y if t.has_y else None, t.Dy, t.Dy, tab["cg_ptr"], tab["cg_idx"],
                                   tab["cg_coef"], None, 0, 0, 0, n, t.DT, dT, t.DT, in_off=r0 * t.Do, y_off=r0 * t.Dy,
                                   ptr_off=t.cg_b)
                for (c0, K, tb, Nr, W_off, _) in t.runs:
                    _mm(K, Nr, n, 1.0, x, r0 * ldx + c0, ldx, dT, tb, t.DT, gwexp, W_off, Nr, ta=True, accumulate=r0 > 0,
                        **self._gemm_kw)
                    if need_dx:
                        _mm(n, K, Nr, 1.0, dT, tb, t.DT, wexp, W_off, Nr, dx, r0 * t.Dx + c0, t.Dx, tb=True,
                            **self._gemm_kw)
            return dx
        nc = t.Dy * t.Do
        dXbig = torch.empty((blk, nc), **f32)
        dx = torch.empty((R, t.Dx), **f32) if need_dx else None
        for r0 in range(0, R, blk):
            n = min(blk, R - r0)
            ops.segnn_yexpand(dz, y if t.has_y else None, t.Dy, n, t.Dy, t.Do, dXbig, dz_off=r0 * t.Do, y_off=r0 * t.Dy)
            _mm(t.Dx, nc, n, 1.0, x, r0 * ldx, ldx, dXbig, 0, nc, gwexp, t.M_off, nc, ta=True, accumulate=r0 > 0,
                **self._gemm_kw)
            if need_dx:
                _mm(n, t.Dx, nc, 1.0, dXbig, 0, nc, wexp, t.M_off, nc, dx, r0 * t.Dx, t.Dx, tb=True, **self._gemm_kw)
        return dx

    # -- first edge layer of a step, evaluated at the nodes ----------------------------------------------------
    def _edge0_ok(self, plan, t: _TPLG, n_add: int) -> bool:
        return plan.lowering == "dense" and t.has_y and 1 <= n_add <= 8

    def _edge0_fwd(self, plan, t: _TPLG, theta, wexp, h, Dh, add, n_add, y, csr, Nt, Et):
        """concat(add, h_s, h_r) @ M_cat = add @ M_a + (h @ M_s)[sender] + (h @ M_r)[receiver]: the two wide products
        are formed per node (1/k of the per-edge rows), the per-edge part is a gather + attribute contraction."""
        rowptr, _, src = csr[:3]
        f32 = dict(dtype=torch.float32, device=theta.device)
        nc = t.Dy * t.Do
        P = torch.empty((2, Nt, nc), **f32)
        Ms = t.M_off + n_add * nc
        _mm(Nt, nc, Dh, 1.0, h, 0, Dh, wexp, Ms, nc, P, 0, nc, **self._gemm_kw)
        _mm(Nt, nc, Dh, 1.0, h, 0, Dh, wexp, Ms + Dh * nc, nc, P, Nt * nc, nc, **self._gemm_kw)
        z = torch.empty((Et, t.Do), **f32)
        ops.segnn_edge0_fwd(P[0], P[1], add, n_add, wexp, t.M_off, y, t.Dy, t.Do, theta if t.n_bias else None, t.b_off or 0,
                            t.b_col, t.n_bias, rowptr, src, Nt, z)
        if not t.activation:
            return z, (h, y, z)
        g = torch.empty((Et, t.d_g), **f32)
        ops.gate_fwd(z, Et, t.n_extra, t.n_gates, t.gate_mul, t.gate_l, plan.inv_c_act, plan.inv_c_gate, g, plan.act)
        return g, (h, y, z)

    def _edge0_bwd(self, plan, t: _TPLG, theta, gtheta, wexp, gwexp, sv, dout, Dh, add, n_add, csr, Nt, Et, dh_prev):
        """Parameter gradients of the layer and the gradient of h (accumulated into dh_prev) from dL/dout [Et, D_out]."""
        h, y, z = sv
        rowptr, _, _, rowptr_s, _, spos = csr
        f32 = dict(dtype=torch.float32, device=theta.device)
        dz = dout
        if t.activation:
            dz = torch.empty((Et, t.Do), **f32)
            ops.gate_bwd(z, dout, Et, t.n_extra, t.n_gates, t.gate_mul, t.gate_l, plan.inv_c_act, plan.inv_c_gate, dz, plan.act)
        if t.n_bias:
            tmp = torch.empty((t.Do,), **f32)
            ops.colsum_tall(dz, Et, t.Do, tmp)
            gtheta[t.b_off:t.b_off + t.n_bias].copy_(tmp[t.b_col:t.b_col + t.n_bias])
        nc = t.Dy * t.Do
        dP = torch.empty((2, Nt, nc), **f32)
        ops.segnn_edge0_bwd(dz, y, t.Dy, t.Do, rowptr, rowptr_s, spos, Nt, dP[0], dP[1])
        Ms = t.M_off + n_add * nc
        for k in range(2):                                   # sender block, receiver block of M_cat
            _mm(Dh, nc, Nt, 1.0, h, 0, Dh, dP, k * Nt * nc, nc, gwexp, Ms + k * Dh * nc, nc, ta=True, **self._gemm_kw)
            _mm(Nt, Dh, nc, 1.0, dP, k * Nt * nc, nc, wexp, Ms + k * Dh * nc, nc, dh_prev, 0, Dh, tb=True, accumulate=True,
                **self._gemm_kw)
        ops.segnn_edge0_wadd(add, n_add, y, t.Dy, dz, t.Do, Et, gwexp, t.M_off)

    # -- last edge layer of a step, evaluated at the nodes -----------------------------------------------------
    def _edgeL_ok(self, plan, st) -> bool:
        t = st["edge"][-1]
        # (moving the aggregation in front of the layer needs a linear aggregation: sum or mean, not max)
        return plan.lowering == "dense" and len(st["edge"]) >= 2 and t.has_y and not t.activation and t.Dy <= 16 and \
            self.message_passing_agg != "max"

    def _edgeL_fwd(self, plan, t: _TPLG, theta, wexp, m, y, rowptr, Nt, agg, Cc, ldc, col_off):
        """The last edge layer has no gate and is only aggregated over the receivers, so the aggregation moves in front
        of it: S_j[n] = f sum_{p in n} y[p, j] m[p] (one pass over the edges), agg[n] = sum_j S_j[n] @ M_j + b w(n)
        (Dy GEMMs on the node rows, written straight into the node function's input Cc[:, col_off:])."""
        f32 = dict(dtype=torch.float32, device=theta.device)
        Dx, nc = t.Dx, t.Dy * t.Do
        S = torch.empty((t.Dy, Nt, Dx), **f32)
        ops.segnn_yagg_fwd(m, Dx, y, t.Dy, rowptr, Nt, agg, S)
        for j in range(t.Dy):
            ops.gemm(Nt, t.Do, Dx, 1.0, S, Dx, 1, wexp, nc, t.Dy, Cc, ldc, 1, accumulate=j > 0, a_off=j * Nt * Dx,
                     b_off=t.M_off + j, c_off=col_off, **self._gemm_kw)
        if t.n_bias:
            ops.segnn_bias_deg_fwd(Cc, col_off + t.b_col, ldc, theta, t.b_off, t.n_bias, rowptr, Nt, agg)
        return (S,)

    def _edgeL_bwd(self, plan, t: _TPLG, theta, gtheta, wexp, gwexp, sv, dC, ldc, col_off, y, rowptr, Nt, Et, agg):
        """-> dL/dm [Et, Dx] from dL/dagg = dC[:, col_off:col_off + Do]; parameter gradients on the node rows."""
        (S,) = sv
        f32 = dict(dtype=torch.float32, device=theta.device)
        Dx, nc = t.Dx, t.Dy * t.Do
        if t.n_bias:
            ops.segnn_bias_deg_bwd(dC, col_off + t.b_col, ldc, rowptr, Nt, agg, t.n_bias, gtheta, t.b_off)
        dS = torch.empty((t.Dy, Nt, Dx), **f32)
        for j in range(t.Dy):
            # dS_j = dagg @ M_j^T ;  gM_j = S_j^T @ dagg  (M_j[k, o] = M_cat[k, o*Dy + j])
            ops.gemm(Nt, Dx, t.Do, 1.0, dC, ldc, 1, wexp, t.Dy, nc, dS, Dx, 1, a_off=col_off, b_off=t.M_off + j,
                     c_off=j * Nt * Dx, **self._gemm_kw)
            ops.gemm(Dx, t.Do, Nt, 1.0, S, 1, Dx, dC, ldc, 1, gwexp, nc, t.Dy, a_off=j * Nt * Dx, b_off=col_off,
                     c_off=t.M_off + j, **self._gemm_kw)
        dm = torch.empty((Et, Dx), **f32)
        ops.segnn_yagg_bwd(dS, Dx, y, t.Dy, rowptr, Nt, agg, dm)
        return dm

    # -- whole model ------------------------------------------------------------------------------------
    def _forward(self, theta, plan, nodes, B, N, E, csr, y_node, y_edge, add, glob=None):
        rowptr, perm, src, rowptr_s, perm_s, spos = csr
        dev = theta.device
        f32 = dict(dtype=torch.float32, device=dev)
        Nt, Et = B * N, B * E
        tab = plan.device_tables(dev)
        wexp = torch.empty((plan._n_exp,), **f32)
        ops.weights_expand(theta, tab["exp_idx"], tab["exp_scale"], wexp)
        agg = ops.AGG[self.message_passing_agg]
        n_add = add.shape[1]
        saved = {"wexp": wexp, "steps": []}
        h, saved["embed"] = self._tplg_fwd(plan, plan.embed, theta, wexp, nodes, nodes.shape[1], y_node, Nt)
        for st in plan.steps:
            Dh = st["h_irreps"].dim
            es = []
            rest = st["edge"]
            if self._edge0_ok(plan, rest[0], n_add):
                m, s_ = self._edge0_fwd(plan, rest[0], theta, wexp, h, Dh, add, n_add, y_edge, csr, Nt, Et)
                es.append(s_)
                rest = rest[1:]
            else:
                m = torch.empty((Et, n_add + 2 * Dh), **f32)
                ops.egnn_edge_input_fwd(add, n_add, h, Dh, 0, Dh, rowptr, src, Nt, m)       # concat(add, h_s, h_r) :112
            ld = m.shape[1]
            fastL = self._edgeL_ok(plan, st)
            for t in (rest[:-1] if fastL else rest):
                m, s_ = self._tplg_fwd(plan, t, theta, wexp, m, ld, y_edge, Et)
                ld = m.shape[1]
                es.append(s_)
            Dm = st["edge"][-1].D_out
            Cc = torch.empty((Nt, Dh + Dm), **f32)                                        # concat(h, agg) :79-81
            ops.copy_cols(h, Dh, 0, Dh, Nt, Cc, Dh + Dm, 0)
            if fastL:
                es.append(self._edgeL_fwd(plan, st["edge"][-1], theta, wexp, m, y_edge, rowptr, Nt, agg, Cc, Dh + Dm, Dh))
            else:
                amax = torch.empty((Nt, Dm), dtype=torch.int32, device=dev) if agg == 2 else None   # segment_max winners
                ops.egnn_segment_rows_fwd(m, Dm, 0, rowptr, Nt, agg, Cc, Dh + Dm, col_off=Dh, argmax=amax)
            new, ns = self._tplg_fwd(plan, st["node"], theta, wexp, Cc, Dh + Dm, y_node, Nt)
            if self.residual:
                ops.axpy(1.0, h, new)
            saved["steps"].append(dict(edge=es, node=ns, Dh=Dh, Dm=Dm, amax=None if fastL else amax))
            h = new
        if self.task == "node":
            x, ld, ds = h, h.shape[1], []
            for t in plan.decode:
                x, s_ = self._tplg_fwd(plan, t, theta, wexp, x, ld, y_node if t.has_y else None, Nt)
                ld = x.shape[1]
                ds.append(s_)
            saved["decode"] = ds
            return x.view(B, N, -1), saved
        d = self.d_hidden
        pre, saved["pre"] = self._tplg_fwd(plan, plan.pre, theta, wexp, h, h.shape[1], y_node, Nt)
        pre2 = torch.empty((Nt, d), **f32)
        _mm(Nt, d, d, 1.0, pre, 0, d, theta, plan.dense[0], d, pre2, 0, d, epi=1, aux=theta, aux_off=plan.dense[1])
        pooled = torch.empty((B, d), **f32)
        ro_arg = None
        if self.readout_agg == "max":                        # jnp.max over the nodes: value + arg-max (gradient routing)
            ro_arg = torch.empty((B, d), dtype=torch.int32, device=dev)
            ops.pool_max_fwd(pre2, B, N, d, pooled, ro_arg)
        else:
            ops.pool_fwd(pre2, B, N, d, 1 if self.readout_agg == "mean" else 0, pooled)
        a, ln_stats = pooled, None
        if plan.ln is not None:
            a = torch.empty((B, d + plan.n_globals), **f32)
            ln_stats = torch.empty((B, 2), **f32)
            ops.layernorm_fwd(pooled, glob, theta, plan.ln[0], plan.ln[1], 1e-6, a, ln_stats)
        ro_in = a
        zs, pro = [], 0
        for (ko, bo, fan, w) in plan.ro_mlp:
            zl = torch.empty((B, w), **f32)
            _mm(B, w, fan, 1.0, a, 0, fan, theta, ko, w, zl, 0, w, pro=pro, pro_scale=1.0, epi=1, aux=theta, aux_off=bo)
            zs.append(zl)
            a, pro = zl, 1
        saved.update(pre_out=pre, pooled=pooled, ro_z=zs, ro_in=ro_in, ln_stats=ln_stats, ro_arg=ro_arg)
        return zs[-1], saved

    def _backward(self, theta, gtheta, saved, gout, plan, nodes, B, N, E, csr, y_node, y_edge, add, glob=None):
        rowptr, perm, src, rowptr_s, perm_s, spos = csr
        dev = theta.device
        f32 = dict(dtype=torch.float32, device=dev)
        Nt, Et = B * N, B * E
        tab = plan.device_tables(dev)
        wexp = saved["wexp"]
        gwexp = torch.zeros((plan._n_exp,), **f32)
        gtheta.zero_()
        agg = ops.AGG[self.message_passing_agg]
        n_add = add.shape[1]
        if self.task == "node":
            dx = gout.reshape(Nt, -1).contiguous()
            for t, s_ in zip(reversed(plan.decode), reversed(saved["decode"])):
                dx = self._tplg_bwd(plan, t, theta, gtheta, wexp, gwexp, s_, dx, Nt)
            dh = dx
        else:
            d = self.d_hidden
            zs = saved["ro_z"]
            dz = gout.reshape(B, -1).contiguous()
            for l in range(len(plan.ro_mlp) - 1, -1, -1):
                ko, bo, fan, w = plan.ro_mlp[l]
                a_prev = zs[l - 1] if l > 0 else saved["ro_in"]
                _mm(fan, w, B, 1.0, a_prev, 0, fan, dz, 0, w, gtheta, ko, w, ta=True, pro=1 if l > 0 else 0, pro_scale=1.0)
                ops.colsum(dz, B, w, gtheta, y_off=bo)
                dprev = torch.empty((B, fan), **f32)
                if l > 0:
                    _mm(B, fan, w, 1.0, dz, 0, w, theta, ko, w, dprev, 0, fan, tb=True, epi=2, aux=zs[l - 1], ld_aux=fan,
                        epi_scale=1.0)
                else:
                    _mm(B, fan, w, 1.0, dz, 0, w, theta, ko, w, dprev, 0, fan, tb=True)
                dz = dprev
            if plan.ln is not None:
                d_pool = torch.empty((B, d), **f32)
                ops.layernorm_bwd(saved["pooled"], glob, theta, plan.ln[0], saved["ln_stats"], dz, d_pool, gtheta,
                                  plan.ln[0], plan.ln[1])
                dz = d_pool
            dpre2 = torch.empty((Nt, d), **f32)
            if self.readout_agg == "max":
                ops.pool_max_bwd(dz, saved["ro_arg"], B, N, d, dpre2)
            else:
                ops.pool_bwd(dz, B, N, d, 1 if self.readout_agg == "mean" else 0, dpre2)
            _mm(d, d, Nt, 1.0, saved["pre_out"], 0, d, dpre2, 0, d, gtheta, plan.dense[0], d, ta=True)     # Dense_0
            ops.colsum_tall(dpre2, Nt, d, gtheta, y_off=plan.dense[1])
            dpre = torch.empty((Nt, d), **f32)
            _mm(Nt, d, d, 1.0, dpre2, 0, d, theta, plan.dense[0], d, dpre, 0, d, tb=True)
            dh = self._tplg_bwd(plan, plan.pre, theta, gtheta, wexp, gwexp, saved["pre"], dpre, Nt)
        for st, sv in zip(reversed(plan.steps), reversed(saved["steps"])):
            Dh, Dm = sv["Dh"], sv["Dm"]
            dC = self._tplg_bwd(plan, st["node"], theta, gtheta, wexp, gwexp, sv["node"], dh, Nt)     # (Nt, Dh + Dm)
            dh_prev = torch.empty((Nt, Dh), **f32)
            ops.copy_cols(dC, Dh + Dm, 0, Dh, Nt, dh_prev, Dh, 0)
            if self.residual:
                ops.axpy(1.0, dh, dh_prev)
            fast0 = self._edge0_ok(plan, st["edge"][0], n_add)
            fastL = self._edgeL_ok(plan, st)
            lo, hi = (1 if fast0 else 0), len(st["edge"]) - (1 if fastL else 0)
            if fastL:
                dm = self._edgeL_bwd(plan, st["edge"][-1], theta, gtheta, wexp, gwexp, sv["edge"][-1], dC, Dh + Dm, Dh, y_edge,
                                     rowptr, Nt, Et, agg)
            else:
                dm = torch.zeros((Et, Dm), **f32)
                ops.egnn_segment_rows_bwd(dC, Dh + Dm, Dh, Dm, None, 0, rowptr, Nt, agg, dm, argmax=sv["amax"])
            for t, s_ in zip(reversed(st["edge"][lo:hi]), reversed(sv["edge"][lo:hi])):
                dm = self._tplg_bwd(plan, t, theta, gtheta, wexp, gwexp, s_, dm, Et)
            if fast0:
                self._edge0_bwd(plan, st["edge"][0], theta, gtheta, wexp, gwexp, sv["edge"][0], dm, Dh, add, n_add, csr,
                                Nt, Et, dh_prev)
            else:
                dhs_e = torch.empty((Et, Dh), **f32)                                       # dm is now dL/dU
                ops.egnn_edge_input_bwd(dm, n_add, Dh, rowptr, perm, rowptr_s, perm_s, Nt, None, dhs_e, dh_prev, Dh, 0)
            dh = dh_prev
        self._tplg_bwd(plan, plan.embed, theta, gtheta, wexp, gwexp, saved["embed"], dh, Nt, need_dx=False)
        ops.weights_contract(gwexp, tab["con_pidx"], tab["con_ptr"], tab["con_tidx"], tab["con_scale"], gtheta)
        return None


class _SEGNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, model, st_graphs):
        prep = model._prepare(st_graphs)
        out, saved = model._forward(theta.detach(), *prep)
        ctx.stuff = (model, saved, prep, theta.detach())
        return out

    @staticmethod
    def backward(ctx, gout):
        model, saved, prep, theta = ctx.stuff
        g = torch.zeros_like(theta)
        model._backward(theta, g, saved, gout.contiguous(), *prep)
        return g, None, None
