"""NequIP on B200 kernels, behind the reference's module interface.

Drop-in for ``models/nequip.py`` (reference): same dataclass fields and defaults
(:97-111), ``GraphsTuple`` in -> ``GraphsTuple`` (task="node") or ``(B, n_outputs)``
array (task="graph") out.  The reference module sees one un-batched graph and is
batched by an outer ``jax.vmap`` (benchmarks/galaxies/train_cosmology.py:210);
here the leading batch axis is native.  Parameters carry the Flax auto-names the
reference creates (Linear_i / MultiLayerPerceptron_s / MLP_0) and live in one
flat fp32 buffer (one all-reduce bucket, one fused optimiser launch).

Data flow per message-passing step, every box a sm_100a kernel:
    xt  = h @ W0x                      Linear of nequip.py:43 (commutes with the gather)
    W   = radial MLP(bessel(|r_ij|))   nequip.py:55-66, edges in receiver-sorted order
    A   = tpconv(xt, Y(r_ij), W)       nequip.py:46-52,68 + jraph segment reduce :118-120
    z   = A @ W1 / sqrt(k) + h @ W2    nequip.py:80-85
    h'  = gate(z) @ W3                 nequip.py:88, :162
Only the message irreps that the following Linear reads ("live" columns) are computed;
the dropped ones cannot influence outputs or gradients (their weights get exact zeros).
"""
from __future__ import annotations

import dataclasses
import functools
import math
from typing import Dict, List, NamedTuple, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops, so3
from .graph_utils import GraphsTuple
from .irreps import Irreps, balanced_irreps
from .plan import LinearPlan, make_linear_plan, make_tp_plan


class IrrepsArray(NamedTuple):
    """Minimal stand-in for ``e3nn_jax.IrrepsArray``: declared irreps + dense array."""
    irreps: Irreps
    array: torch.Tensor


@functools.lru_cache(maxsize=None)
def normalize_constant(name: str) -> float:
    """``e3nn.normalize_function`` second moment on its deterministic grid."""
    u = torch.linspace(-1.0, 1.0, 1_000_001 + 2, dtype=torch.float64)[1:-1]
    x = math.sqrt(2.0) * torch.special.erfinv(u)
    if name == "gelu":
        y = 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))
    elif name == "sigmoid":
        y = torch.sigmoid(x)
    elif name == "silu":
        y = x * torch.sigmoid(x)
    else:
        raise NotImplementedError(f"activation {name!r} has no B200 kernel on this path")
    return float(torch.sqrt(torch.mean(y * y)))


# --------------------------------------------------------------------------------------
# parameters
# --------------------------------------------------------------------------------------
class NequIPParams:
    """Flat fp32 parameter buffer + Flax-style nested views."""

    def __init__(self, layout: List[Tuple[Tuple[str, ...], Tuple[int, ...], int]], flat: torch.Tensor):
        self.layout = layout
        self.flat = flat
        self.grad = torch.zeros_like(flat)

    @property
    def tree(self) -> Dict:
        return self._tree_of(self.flat)

    @property
    def grad_tree(self) -> Dict:
        return self._tree_of(self.grad)

    def _tree_of(self, buf: torch.Tensor) -> Dict:
        out: Dict = {}
        for path, shape, off in self.layout:
            d = out
            for p in path[:-1]:
                d = d.setdefault(p, {})
            d[path[-1]] = buf[off:off + int(np.prod(shape))].view(*shape)
        return out

    def load_tree(self, tree: Dict) -> None:
        """Copy values from a nested dict of arrays with the same names / shapes."""
        host = np.empty((self.flat.numel(),), dtype=np.float32)
        for path, shape, off in self.layout:
            v = tree
            for p in path:
                v = v[p]
            v = np.asarray(v.detach().cpu() if isinstance(v, torch.Tensor) else v, dtype=np.float32)
            if tuple(v.shape) != tuple(shape):
                raise ValueError(f"{'/'.join(path)}: shape {v.shape} != {shape}")
            host[off:off + v.size] = v.reshape(-1)
        self.flat.copy_(torch.from_numpy(host))

    def num_params(self) -> int:
        return int(self.flat.numel())


# --------------------------------------------------------------------------------------
# static plan
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class _Dense:
    off: int          # offset in the expanded-weight buffer
    rows: int
    cols: int


class _Plan:
    def __init__(self, cfg: "NequIP", irreps_in: Irreps, n_globals: int = 0):
        self.irreps_in = irreps_in
        self.n_globals = int(n_globals)
        D = irreps_in.dim
        if len(irreps_in) == 0 or irreps_in[0][1] != 1:
            raise ValueError("the first irrep of the node features must be a vector (positions), nequip.py:129-132")
        if cfg.activation not in ("gelu", "silu"):
            raise NotImplementedError("activation must be 'gelu' or 'silu' on the B200 path (models/nequip.py:61)")
        if cfg.message_passing_agg not in ("sum", "mean", "max"):
            raise ValueError(f"Invalid message_passing_agg {cfg.message_passing_agg}")
        if cfg.task == "graph" and cfg.readout_agg not in ("sum", "mean", "max"):
            raise ValueError(f"Invalid global aggregation function {cfg.readout_agg}")    # models/nequip.py:174-177
        hidden = balanced_irreps(cfg.l_max, cfg.d_hidden, True)
        tp_full = make_tp_plan(irreps_in, cfg.l_max, None)
        self.tp_full = tp_full
        self.tp_full_spec = (str(irreps_in), cfg.l_max,
                             "+".join(f"{l}{'e' if p > 0 else 'o'}" for l, p in tp_full.live_types))
        msg = tp_full.irreps_msg
        irr = hidden.filter(keep=msg)
        n_gated = irr.num_irreps - irr.count("0e")
        upd = (Irreps([(n_gated, 0, 1)]) + irr).regroup()
        if upd.filter(keep=msg) != upd:
            raise NotImplementedError("hidden irreps absent from the messages are not supported")
        # Live message columns: irreps that (a) the node Linear reads (they occur in `upd`) and (b) can still
        # reach the step's output.  Every step re-projects to the input irreps (nequip.py:162), so hidden
        # irreps that do not occur in the input (e.g. 2e, 3o for "1o+1o+1x0e") are dropped by that Linear:
        # their messages, gates and weights cannot influence outputs, and their gradients are exact zeros
        # in the reference as well.
        # (0e is always kept: scalar messages feed the gate scalars of the surviving l>0 irreps.)
        live = upd.filter(keep=irreps_in + Irreps("1x0e"))
        self.tp = make_tp_plan(irreps_in, cfg.l_max, live)
        self.tp_sig = self.tp.signature
        # (x irreps, lmax, live irreps): lets a signature outside codegen.PREBUILT be generated and compiled on first use
        self.tp_spec = (str(irreps_in), cfg.l_max, str(live))
        self.upd = upd
        scal = [(m, l, p) for m, l, p in upd if l == 0]
        gated = [(m, l, p) for m, l, p in upd if l > 0]
        if any(p == -1 for _, _, p in scal) or len(scal) != 1:
            raise NotImplementedError("odd scalars are not on the reference's path")
        n_s = scal[0][0]
        n_gates_full = sum(m for m, _, _ in gated)
        self.n_extra = n_s - n_gates_full
        self.g_irreps = Irreps(([(self.n_extra, 0, 1)] if self.n_extra else []) + gated)   # reference layout
        # Node update restricted to the live irreps (same argument as for the messages): gated chunks whose
        # irrep does not occur in the input are dropped together with their gate scalars.  zmap / gmap send a
        # component index of the reference's z / gate-output layout to the compact layout (or -1).
        live_types = {(l, p) for _, l, p in irreps_in}
        z_full, g_full = upd.dim, self.g_irreps.dim
        self.zmap, self.gmap = [-1] * z_full, [-1] * g_full
        for c in range(self.n_extra):
            self.zmap[c] = self.gmap[c] = c
        self.gate_mul, self.gate_l = [], []
        zg, zc, gate_pos, comp_pos = self.n_extra, 0, [], []
        n_gates_live = sum(m for m, l, pp in gated if (l, pp) in live_types)
        goff_full, coff_full, glive, clive = 0, 0, 0, 0
        for m, l, pp in gated:
            d = 2 * l + 1
            if (l, pp) in live_types:
                self.gate_mul.append(m)
                self.gate_l.append(l)
                for u in range(m):
                    self.zmap[self.n_extra + goff_full + u] = self.n_extra + glive + u
                for q in range(m * d):
                    self.zmap[self.n_extra + n_gates_full + coff_full + q] = self.n_extra + n_gates_live + clive + q
                    self.gmap[self.n_extra + coff_full + q] = self.n_extra + clive + q
                glive += m
                clive += m * d
            goff_full += m
            coff_full += m * d
        self.n_gates = n_gates_live
        self.d_z = self.n_extra + n_gates_live + clive
        self.d_g = self.n_extra + clive
        self.d_in = D
        self.n_rad_in = cfg.n_radial_basis if cfg.n_radial_basis > 0 else 1
        self.shn = [so3.sh_norm_factor(l, cfg.sphharm_norm) for l in range(cfg.l_max + 1)]
        # the radial MLP uses `activation` (nequip.py:61); e3nn.gate keeps its defaults, gelu / sigmoid (nequip.py:88-90)
        self.inv_c_mlp = 1.0 / normalize_constant(cfg.activation)
        self.inv_c_act = 1.0 / normalize_constant("gelu")
        self.inv_c_gate = 1.0 / normalize_constant("sigmoid")

        # ---- parameter layout (Flax creation order of the reference module) ----
        self.layout: List[Tuple[Tuple[str, ...], Tuple[int, ...], int]] = []
        self._poff: Dict[Tuple[str, ...], int] = {}
        self._n_params = 0
        # ---- expanded dense matrices ----
        self._entries: List[Tuple[int, int, float]] = []   # (dst index, flat param index or -2, scale)
        self._n_exp = 0
        self.steps: List[Dict] = []

        full_to_live = {}
        for c in self.tp.columns:
            if c.live >= 0:
                for i in range(2 * c.l3 + 1):
                    full_to_live[c.msg_off + i] = c.live_off + i
        li = 0
        widths = cfg.n_layers * (cfg.d_hidden,) + (tp_full.n_irr,)
        for s in range(cfg.message_passing_steps):
            st: Dict = {}
            lp = make_linear_plan(irreps_in, irreps_in, out_layout="regrouped", cols_pad=self.tp.dxp)
            st["W0x"] = self._add_linear(f"Linear_{li}", lp); li += 1
            mlp_name = f"MultiLayerPerceptron_{s}"
            st["mlp"] = []
            fan = self.n_rad_in
            for l, h in enumerate(widths):
                st["mlp"].append((self._add_param((mlp_name, f"Dense_{l}", "kernel"), (fan, h)), fan, h))
                fan = h
            # live columns of the last radial layer, SH normalisation of each column folded in
            koff, kfan, kh = st["mlp"][-1]
            st["W4L"] = self._new_dense(kfan, self.tp.n_live)
            for c in self.tp.columns:
                if c.live >= 0:
                    for r in range(kfan):
                        self._entries.append((st["W4L"].off + r * self.tp.n_live + c.live, koff + r * kh + c.col,
                                              self.shn[c.l2]))
            lp = make_linear_plan(msg, upd)
            zcols = {c: v for c, v in enumerate(self.zmap) if v >= 0}
            st["W1"] = self._add_linear(f"Linear_{li}", lp, row_map=full_to_live, rows=self.tp.d_live,
                                        skip_dead=True, col_map=zcols, cols=self.d_z); li += 1
            if cfg.residual:
                lp = make_linear_plan(irreps_in, upd, force_irreps_out=True)
                st["W2"] = self._add_linear(f"Linear_{li}", lp, skip_dead=True, col_map=zcols, cols=self.d_z); li += 1
            lp = make_linear_plan(self.g_irreps, irreps_in)
            st["W3"] = self._add_linear(f"Linear_{li}", lp, row_map={r: v for r, v in enumerate(self.gmap) if v >= 0},
                                        rows=self.d_g, skip_dead=True); li += 1
            self.steps.append(st)

        self.irreps_out = Irreps(cfg.irreps_out) if cfg.irreps_out is not None else irreps_in
        self.Wout = None
        if cfg.task == "node":
            if self.irreps_out != irreps_in:
                lp = make_linear_plan(irreps_in, self.irreps_out)
                self.Wout = self._add_linear(f"Linear_{li}", lp); li += 1
                self.d_out = lp.cols
            else:
                self.d_out = D
        elif cfg.task == "graph":
            self.tp_ro = make_tp_plan(irreps_in, cfg.l_max, "0e")
            self.tp_ro_spec = (str(irreps_in), cfg.l_max, "0e")
            # constant regroup + pad matrix  h (stored) -> x (regrouped, padded)
            self.P = self._new_dense(D, self.tp_ro.dxp)
            for r, s_off in enumerate(irreps_in.regroup_index_map()):
                self._entries.append((self.P.off + s_off * self.tp_ro.dxp + r, -2, 1.0))
            lp = make_linear_plan(msg, Irreps([(cfg.d_hidden, 0, 1)]))
            ro_map, ro_scale = {}, {}
            for c in self.tp_ro.columns:
                if c.live >= 0:
                    for i in range(2 * c.l3 + 1):
                        ro_map[c.msg_off + i] = c.live_off + i
                        ro_scale[c.live_off + i] = self.shn[c.l2]
            self.Wro = self._add_linear(f"Linear_{li}", lp, row_map=ro_map, rows=self.tp_ro.d_live,
                                        row_scale=ro_scale); li += 1
            # graph globals are concatenated to the pooled scalars and LayerNorm-ed (nequip.py:201-205)
            n_pool = cfg.d_hidden + self.n_globals
            self.ln = None
            if self.n_globals > 0:
                self.ln = (self._add_param(("LayerNorm_0", "scale"), (n_pool,)),
                           self._add_param(("LayerNorm_0", "bias"), (n_pool,)))
            ws = [cfg.mlp_readout_widths[0] * n_pool] + [w * cfg.d_hidden for w in cfg.mlp_readout_widths[1:]] + \
                 [cfg.n_outputs]
            self.ro_mlp = []
            fan = n_pool
            for l, h in enumerate(ws):
                ko = self._add_param(("MLP_0", f"Dense_{l}", "kernel"), (fan, h))
                bo = self._add_param(("MLP_0", f"Dense_{l}", "bias"), (h,))
                self.ro_mlp.append((ko, bo, fan, h))
                fan = h
        else:
            raise ValueError(f"Invalid task {cfg.task}")

        # ---- device tables ----
        n = self._n_exp
        idx = np.full((n,), -1, dtype=np.int32)
        scale = np.zeros((n,), dtype=np.float32)
        for d, p, sc in self._entries:
            idx[d] = p
            scale[d] = sc
        self.exp_idx_np, self.exp_scale_np = idx, scale
        # contraction CSR: for every parameter that feeds dense entries
        by_param: Dict[int, List[Tuple[int, float]]] = {}
        for d, p, sc in self._entries:
            if p >= 0:
                by_param.setdefault(p, []).append((d, sc))
        pidx = np.array(sorted(by_param), dtype=np.int32)
        ptr = np.zeros((len(pidx) + 1,), dtype=np.int32)
        tidx, tsc = [], []
        for j, p in enumerate(pidx):
            for d, sc in by_param[int(p)]:
                tidx.append(d)
                tsc.append(sc)
            ptr[j + 1] = len(tidx)
        self.con_pidx_np, self.con_ptr_np = pidx, ptr
        self.con_tidx_np = np.array(tidx, dtype=np.int32)
        self.con_scale_np = np.array(tsc, dtype=np.float32)
        self._dev: Dict[str, Dict[str, torch.Tensor]] = {}

    # -- helpers ------------------------------------------------------------------------
    def _add_param(self, path: Tuple[str, ...], shape: Tuple[int, ...]) -> int:
        off = self._n_params
        self.layout.append((path, tuple(shape), off))
        self._poff[path] = off
        self._n_params += int(np.prod(shape))
        return off

    def _new_dense(self, rows: int, cols: int) -> _Dense:
        d = _Dense(self._n_exp, rows, cols)
        self._n_exp += rows * cols
        return d

    def _add_linear(self, name: str, lp: LinearPlan, row_map=None, rows=None, row_scale=None,
                    skip_dead: bool = False, col_map=None, cols=None) -> _Dense:
        """Registers the Linear's Flax parameters and its dense lowering.  row_map / col_map re-index the dense
        matrix into a compact (live-only) layout; with skip_dead, entries that touch a dropped row or column are
        omitted: the parameter still exists (and receives an exact zero gradient), its term is never computed."""
        offs = {pn: self._add_param((name, pn), shp) for pn, shp in lp.param_shapes.items()}
        ncols = cols if cols is not None else lp.cols
        d = self._new_dense(rows if rows is not None else lp.rows, ncols)
        for r, c, pn, flat, sc in lp.entries:
            if row_map is not None:
                if r not in row_map:
                    if skip_dead:
                        continue
                    raise AssertionError("Linear reads a component that is not live")
                r = row_map[r]
            if col_map is not None:
                if c not in col_map:
                    if skip_dead:
                        continue
                    raise AssertionError("Linear writes a component that is not live")
                c = col_map[c]
            if row_scale is not None:
                sc = sc * row_scale[r]
            self._entries.append((d.off + r * ncols + c, offs[pn] + flat, sc))
        return d

    def device_tables(self, device) -> Dict[str, torch.Tensor]:
        key = str(device)
        if key not in self._dev:
            t = lambda a: torch.from_numpy(a).to(device)
            self._dev[key] = dict(exp_idx=t(self.exp_idx_np), exp_scale=t(self.exp_scale_np),
                                  con_pidx=t(self.con_pidx_np), con_ptr=t(self.con_ptr_np),
                                  con_tidx=t(self.con_tidx_np), con_scale=t(self.con_scale_np))
        return self._dev[key]


# --------------------------------------------------------------------------------------
# prepared graph (receiver-sorted CSR + geometry), built once per batch
# --------------------------------------------------------------------------------------
class PreparedGraph(NamedTuple):
    B: int
    N: int
    E: int
    rowptr: torch.Tensor
    perm: torch.Tensor
    src: torch.Tensor
    rowptr_s: torch.Tensor
    perm_s: torch.Tensor
    geom: torch.Tensor
    radial: Optional[torch.Tensor]
    pairs: Optional["ops.EdgePairs"] = None     # shared radial evaluations (NequIP.share_radial)


def _mm(M, N, K, alpha, A, a_off, lda, Bm, b_off, ldb, Cm, c_off, ldc, ta=False, tb=False, tc=False, **kw):
    """C (M x N) = alpha * op(A) (M x K) @ op(B) (K x N).  ta: A stored K x M; tb: B stored N x K; tc: C stored N x M."""
    sam, sak = (1, lda) if ta else (lda, 1)
    sbk, sbn = (1, ldb) if tb else (ldb, 1)
    scm, scn = (1, ldc) if tc else (ldc, 1)
    ops.gemm(M, N, K, alpha, A, sam, sak, Bm, sbk, sbn, Cm, scm, scn, a_off=a_off, b_off=b_off, c_off=c_off, **kw)


# --------------------------------------------------------------------------------------
# the module
# --------------------------------------------------------------------------------------
@dataclasses.dataclass
class NequIP:
    """Field names and defaults of the reference module (models/nequip.py:97-111)."""
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
    irreps_out: Optional[Union[str, Irreps]] = None
    residual: bool = True
    # not a reference field: run the radial-MLP GEMMs on tcgen05 tensor cores with 3xTF32 operand
    # splitting (fp32-class accuracy); False = fp32 FMA on the CUDA cores
    tensor_cores: bool = True
    # not a reference field: Linear + self-connection + gate + re-projection as one fused N-row kernel
    fuse_node_update: bool = True
    # not a reference field: the whole radial MLP (Bessel basis -> n_layers Dense + gelu -> radial weights) as ONE
    # tcgen05 kernel per direction with the activations in tensor memory (eqnn_radial_mlp_fwd / _bwd); shapes the
    # fused kernels do not cover (d_hidden != 128, n_radial_basis not in {16, 32, 64}, n_layers > 3, more than 16 live
    # message columns) keep the per-layer GEMMs.  "stash": the forward kernel keeps the pre-activations for the
    # backward pass (1.5 KB per edge and step); "recompute": nothing is kept, the backward recomputes them from d.
    fuse_radial_mlp: bool = True
    radial_mlp_backward: str = "stash"
    # not a reference field: evaluate the radial MLP once per UNIQUE edge length instead of once per edge.  The MLP sees an
    # edge only through d = |x_s - x_r| (models/nequip.py:55-68) and the two directions of a reciprocated edge have
    # bit-identical lengths (85 % of the edges of a kNN-20 graph are reciprocated), so both read one evaluation and their
    # cotangents are summed before the MLP's backward pass: the same arithmetic on the same numbers, 0.58 of the radial-MLP
    # work (csrc/pairs.cu).  Applies on the fused stash route (gelu, >= 2 hidden layers, >= 4096 edges); False = one
    # evaluation per edge.
    share_radial: bool = True
    # not a reference field: the reference's node task returns the processed GraphsTuple, whose `edges` are the last
    # step's messages (models/nequip.py:154-171); no caller reads them, so they are only computed on request (every message
    # column, un-reduced, (B, E, D_msg) in the reference's edge order; not differentiated through)
    return_edges: bool = False

    def __post_init__(self):
        self._plans: Dict[Irreps, _Plan] = {}

    # -- plans / params -----------------------------------------------------------------
    def plan(self, irreps_in, n_globals: int = 0) -> _Plan:
        irreps_in = Irreps(irreps_in)
        n_globals = int(n_globals) if self.task == "graph" else 0
        key = (irreps_in, n_globals)
        if key not in self._plans:
            self._plans[key] = _Plan(self, irreps_in, n_globals)
        return self._plans[key]

    @staticmethod
    def _nodes_of(graphs: GraphsTuple) -> Tuple[Irreps, torch.Tensor]:
        nodes = graphs.nodes
        if not isinstance(nodes, IrrepsArray):
            raise TypeError("NequIP expects graphs.nodes to be an IrrepsArray (reference: e3nn.IrrepsArray)")
        return Irreps(nodes.irreps), nodes.array

    def init(self, rng: Union[int, np.random.Generator], graphs: GraphsTuple) -> NequIPParams:
        """Flax-style ``model.init``: draws parameters with the library initialisers' distributions
        (e3nn Linear / MultiLayerPerceptron: N(0,1); models/mlp.py Dense: xavier-uniform, zero bias)."""
        irreps_in, arr = self._nodes_of(graphs)
        plan = self.plan(irreps_in, _n_globals_of(graphs))
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        host = np.empty((plan._n_params,), dtype=np.float32)
        for path, shape, off in plan.layout:
            n = int(np.prod(shape))
            if path[0] == "LayerNorm_0":
                v = np.ones(shape) if path[-1] == "scale" else np.zeros(shape)
            elif path[0] == "MLP_0":
                if path[-1] == "bias":
                    v = np.zeros(shape)
                else:
                    lim = math.sqrt(6.0 / (shape[0] + shape[1]))
                    v = rng.uniform(-lim, lim, size=shape)
            else:
                v = rng.normal(size=shape)
            host[off:off + n] = v.reshape(-1)
        flat = torch.from_numpy(host).to(arr.device)
        return NequIPParams(plan.layout, flat)

    # -- graph preparation ----------------------------------------------------------------
    def prepare(self, graphs: GraphsTuple) -> PreparedGraph:
        irreps_in, arr = self._nodes_of(graphs)
        if arr.dim() == 2:
            arr = arr.unsqueeze(0)
        B, N, D = arr.shape
        senders = graphs.senders.reshape(B, -1).contiguous()
        receivers = graphs.receivers.reshape(B, -1).contiguous()
        E = senders.shape[1]
        rowptr, perm, src = ops.csr_build(senders, receivers, N)
        rowptr_s, perm_s, _ = ops.csr_build(receivers, senders, N)
        arr = arr.contiguous()
        share = self._share_radial(self.plan(irreps_in, _n_globals_of(graphs)), B * E)
        geom, radial = ops.edge_geom(arr, D, B * N, rowptr, src, self.n_radial_basis, self.r_cutoff, want_radial=not share)
        pairs = ops.edge_pairs(rowptr, src, geom, B * N, self.n_radial_basis, self.r_cutoff) if share else None
        return PreparedGraph(B, N, E, rowptr, perm, src, rowptr_s, perm_s, geom, radial, pairs)

    def _share_radial(self, plan: "_Plan", n_edges: int) -> bool:
        """One radial-MLP evaluation per unique edge length (share_radial) applies to this model and batch."""
        return bool(self.share_radial and not self.return_edges and self.activation == "gelu"
                    and self.radial_mlp_backward == "stash" and n_edges >= 4096 and len(plan.steps) > 0
                    and len(plan.steps[0]["mlp"]) - 1 >= 2 and _fused_radial_ok(self, plan))

    # -- forward / backward -----------------------------------------------------------------
    def apply(self, params: NequIPParams, graphs: GraphsTuple, prepared: Optional[PreparedGraph] = None):
        """Flax-style ``model.apply``.  Differentiable w.r.t. ``params.flat`` through torch autograd."""
        irreps_in, arr = self._nodes_of(graphs)
        unbatched = arr.dim() == 2
        if unbatched:
            arr = arr.unsqueeze(0)
        plan = self.plan(irreps_in, _n_globals_of(graphs))
        pg = prepared if prepared is not None else self.prepare(graphs)
        k = _n_edge_of(graphs)
        out, edges = _NequIPFunction.apply(params.flat, self, plan, pg, arr.contiguous(), k,
                                           _globals_of(graphs, arr.shape[0]))
        if self.task == "graph":
            return out[0] if unbatched else out
        nodes = out[0] if unbatched else out
        res = graphs._replace(nodes=IrrepsArray(plan.irreps_out if plan.Wout is None else
                                                Irreps(self.irreps_out).filter(keep=irreps_in), nodes))
        if self.return_edges:                                       # models/nequip.py:154-171: the last step's messages
            res = res._replace(edges=IrrepsArray(plan.tp_full.irreps_msg, edges[0] if unbatched else edges))
        return res

    __call__ = apply

    def forward_backward(self, params: NequIPParams, graphs: GraphsTuple, grad_fn, prepared=None):
        """Training fast path without torch autograd: runs forward, calls ``grad_fn(out) -> dL/dout``
        (device tensors), runs backward; gradient lands in ``params.grad``.  Returns the output."""
        irreps_in, arr = self._nodes_of(graphs)
        if arr.dim() == 2:
            arr = arr.unsqueeze(0)
        plan = self.plan(irreps_in, _n_globals_of(graphs))
        pg = prepared if prepared is not None else self.prepare(graphs)
        k = _n_edge_of(graphs)
        out, saved = _forward(self, plan, params.flat, pg, arr.contiguous(), k, save=True,
                              globals_=_globals_of(graphs, arr.shape[0]))
        gout = grad_fn(out)
        _backward(self, plan, params.flat, pg, saved, gout.contiguous(), k, params.grad)
        return out


def globals_present(graphs: GraphsTuple) -> bool:
    return graphs.globals is not None


def _n_globals_of(graphs: GraphsTuple) -> int:
    return 0 if graphs.globals is None else int(graphs.globals.shape[-1])


def _globals_of(graphs: GraphsTuple, B: int) -> Optional[torch.Tensor]:
    """(B, G) fp32 device tensor of graph-level features, or None."""
    if graphs.globals is None:
        return None
    g = graphs.globals
    if not isinstance(g, torch.Tensor) or not g.is_cuda:
        raise RuntimeError("graphs.globals must be a CUDA tensor (this path has no CPU implementation)")
    return g.reshape(B, -1).to(torch.float32).contiguous()


def _n_edge_of(graphs: GraphsTuple) -> float:
    ne = graphs.n_edge
    if isinstance(ne, torch.Tensor):
        v = ne.reshape(-1)
        k = float(v[0].item())
        return k
    v = np.asarray(ne).reshape(-1)
    return float(v[0])


def _fused_radial_ok(model: "NequIP", plan: _Plan) -> bool:
    """The fused radial-MLP kernels cover this model (tensor-core path requested, supported shape)."""
    if not (model.fuse_radial_mlp and model.tensor_cores):
        if model.activation != "gelu":
            raise NotImplementedError("activation='silu' runs on the fused radial-MLP kernels only")
        return False
    if model.radial_mlp_backward not in ("stash", "recompute"):
        raise ValueError("radial_mlp_backward must be 'stash' or 'recompute'")
    ok = ops.radial_mlp_supported(model.n_radial_basis, model.d_hidden, model.n_layers, plan.tp.n_live)
    if not ok and model.activation != "gelu":
        raise NotImplementedError("activation='silu' needs a radial MLP the fused kernels cover "
                                  "(d_hidden 128, n_radial_basis 16/32/64, n_layers <= 3)")
    return ok


def _radial_weights(theta: torch.Tensor, wexp: torch.Tensor, st: Dict, n_live: int):
    """[(buffer, element offset, leading dimension)] of the radial MLP's Dense kernels: the hidden layers are Flax
    parameters in the flat buffer, the output layer is its live-column expansion (works for the gradient buffers too)."""
    ws = [(theta, koff, hdim) for (koff, fan, hdim) in st["mlp"][:-1]]
    ws.append((wexp, st["W4L"].off, n_live))
    return ws


def _edge_messages(model: "NequIP", plan: _Plan, theta: torch.Tensor, pg: PreparedGraph, st: Dict, xt: torch.Tensor):
    """models/nequip.py:39-71 for every edge and EVERY message column, un-reduced: what jraph.GraphNetwork leaves in
    `edges` after the last step (:154-171).  The radial MLP runs layer by layer here (the fused kernel emits at most 16
    columns); (B, E, D_msg) in the reference's edge order."""
    if model.activation != "gelu":
        raise NotImplementedError("return_edges: the layer-by-layer radial MLP has the gelu kernels only")
    dev = theta.device
    f32 = dict(dtype=torch.float32, device=dev)
    Et = pg.B * pg.E
    a, pro = pg.radial, 0
    for (koff, fan, hdim) in st["mlp"][:-1]:
        Z = torch.empty((Et, hdim), **f32)
        _mm(Et, hdim, fan, 1.0 / math.sqrt(fan), a, 0, fan, theta, koff, hdim, Z, 0, hdim, pro=pro,
            pro_scale=plan.inv_c_act, tf32x3=model.tensor_cores)
        a, pro = Z, 1
    koff, fan, n_irr = st["mlp"][-1]
    w_rows = torch.empty((Et, n_irr), **f32)
    _mm(Et, n_irr, fan, 1.0 / math.sqrt(fan), a, 0, fan, theta, koff, n_irr, w_rows, 0, n_irr, pro=pro,
        pro_scale=plan.inv_c_act)
    tpf = plan.tp_full
    tp_id = ops.tp_lookup(tpf.signature, plan.tp_full_spec)
    out = torch.empty((Et, tpf.irreps_msg.dim), **f32)
    ops.tp_messages(tp_id, pg.src, pg.perm, pg.geom, w_rows, [c.col for c in tpf.columns],
                    [plan.shn[c.l2] for c in tpf.columns], xt, out)
    return out.view(pg.B, pg.E, -1)


class _NequIPFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, model, plan, pg, nodes, k, globals_=None):
        out, saved = _forward(model, plan, theta.detach(), pg, nodes, k, save=True, globals_=globals_)
        ctx.model, ctx.plan, ctx.pg, ctx.saved, ctx.k = model, plan, pg, saved, k
        ctx.theta = theta.detach()
        edges = saved.pop("edges", None)
        if edges is None:
            edges = out.new_empty((0,))
        ctx.mark_non_differentiable(edges)
        return out, edges

    @staticmethod
    def backward(ctx, gout, _gedges):
        g = torch.zeros_like(ctx.theta)
        _backward(ctx.model, ctx.plan, ctx.theta, ctx.pg, ctx.saved, gout.contiguous(), ctx.k, g)
        return g, None, None, None, None, None, None


def _forward(model: NequIP, plan: _Plan, theta: torch.Tensor, pg: PreparedGraph, nodes: torch.Tensor, k: float,
             save: bool, globals_: Optional[torch.Tensor] = None):
    dev = nodes.device
    B, N, D = nodes.shape
    Nt, Et = B * N, B * pg.E
    f32 = dict(dtype=torch.float32, device=dev)
    tab = plan.device_tables(dev)
    wexp = torch.empty((plan._n_exp,), **f32)
    ops.weights_expand(theta, tab["exp_idx"], tab["exp_scale"], wexp)
    tp_id = ops.tp_lookup(plan.tp_sig, plan.tp_spec)
    dxp, n_live, d_live = plan.tp.dxp, plan.tp.n_live, plan.tp.d_live
    agg = ops.AGG[model.message_passing_agg]
    inv_sqrt_k = 1.0 / math.sqrt(k)
    h = nodes.reshape(Nt, D)
    saved = {"steps": [], "wexp": wexp}
    for st in plan.steps:
        xt = torch.empty((Nt, dxp), **f32)
        _mm(Nt, dxp, D, 1.0, h, 0, D, wexp, st["W0x"].off, dxp, xt, 0, dxp)
        Zs, a, fan_prev, pro = [], pg.radial, plan.n_rad_in, 0
        fused = _fused_radial_ok(model, plan)
        stash = None
        if fused and pg.pairs is not None:
            # one evaluation per unique edge length; the interaction kernel gathers the rows through pg.pairs.uidx
            w_t = torch.empty((Et, (n_live + 3) // 4 * 4), **f32)
            if save:
                stash = ops.radial_mlp_stash(Et, len(st["mlp"]) - 1, dev)
            ops.radial_mlp_fwd_rows(pg.pairs, model.n_radial_basis, model.r_cutoff, _radial_weights(theta, wexp, st, n_live),
                                    n_live, plan.inv_c_mlp, w_t, z_stash=stash, activation=model.activation)
        elif fused:
            w_t = torch.empty((n_live, Et), **f32)
            if save and model.radial_mlp_backward == "stash" and model.activation == "gelu":
                stash = ops.radial_mlp_stash(Et, len(st["mlp"]) - 1, dev)
            ops.radial_mlp_fwd(pg.geom, model.n_radial_basis, model.r_cutoff, _radial_weights(theta, wexp, st, n_live),
                               n_live, plan.inv_c_mlp, w_t, z_stash=stash, activation=model.activation)
        for (koff, fan, hdim) in ([] if fused else st["mlp"][:-1]):
            Z = torch.empty((Et, hdim), **f32)
            _mm(Et, hdim, fan, 1.0 / math.sqrt(fan), a, 0, fan, theta, koff, hdim, Z, 0, hdim,
                pro=pro, pro_scale=plan.inv_c_act, tag="radial_mlp_fwd", tf32x3=model.tensor_cores)
            Zs.append(Z)
            a, pro = Z, 1
        _, fan, _ = st["mlp"][-1]
        if not fused:
            w_t = torch.empty((n_live, Et), **f32)
            _mm(Et, n_live, fan, 1.0 / math.sqrt(fan), a, 0, fan, wexp, st["W4L"].off, n_live, w_t, 0, Et, tc=True,
                pro=pro, pro_scale=plan.inv_c_act, tag="radial_mlp_fwd", tf32x3=model.tensor_cores)
        A = torch.empty((Nt, d_live), **f32)
        amax = torch.empty((Nt, d_live), dtype=torch.int32, device=dev) if agg == 2 else None   # segment_max winners
        if fused and pg.pairs is not None:
            ops.tpconv_fwd_shared(tp_id, pg.rowptr, pg.src, pg.geom, pg.pairs, w_t, xt, Nt, agg, A, d_live, amax)
        else:
            ops.tpconv_fwd(tp_id, pg.rowptr, pg.src, pg.geom, w_t, xt, Nt, agg, A, d_live, amax)
        z = torch.empty((Nt, plan.d_z), **f32)
        hn = torch.empty((Nt, D), **f32)
        g = None
        if model.fuse_node_update:
            ops.node_update_fwd(A, d_live, d_live, h, D, wexp, st["W1"].off, inv_sqrt_k,
                                st["W2"].off if model.residual else None, plan.n_extra, plan.n_gates, plan.gate_mul,
                                plan.gate_l, plan.inv_c_act, plan.inv_c_gate, st["W3"].off, Nt, z, hn)
        else:
            _mm(Nt, plan.d_z, d_live, inv_sqrt_k, A, 0, d_live, wexp, st["W1"].off, plan.d_z, z, 0, plan.d_z)
            if model.residual:
                _mm(Nt, plan.d_z, D, 1.0, h, 0, D, wexp, st["W2"].off, plan.d_z, z, 0, plan.d_z, accumulate=True)
            g = torch.empty((Nt, plan.d_g), **f32)
            ops.gate_fwd(z, Nt, plan.n_extra, plan.n_gates, plan.gate_mul, plan.gate_l, plan.inv_c_act,
                         plan.inv_c_gate, g)
            _mm(Nt, D, plan.d_g, 1.0, g, 0, plan.d_g, wexp, st["W3"].off, D, hn, 0, D)
        if save:
            saved["steps"].append(dict(h=h, xt=xt, Zs=Zs, w_t=w_t, A=A, z=z, g=g, stash=stash, amax=amax))
        if model.return_edges and model.task == "node" and st is plan.steps[-1]:
            saved["edges"] = _edge_messages(model, plan, theta, pg, st, xt)
        h = hn
    saved["h_final"] = h
    if model.task == "node":
        if plan.Wout is not None:
            o = torch.empty((Nt, plan.d_out), **f32)
            _mm(Nt, plan.d_out, D, 1.0, h, 0, D, wexp, plan.Wout.off, plan.d_out, o, 0, plan.d_out)
        else:
            o = h
        return o.view(B, N, -1), saved
    # ---- graph read-out (nequip.py:172-214)
    tpr = plan.tp_ro
    ro_id = ops.tp_lookup(tpr.signature, plan.tp_ro_spec)
    attrs = torch.empty((Nt, tpr.ny), **f32)
    ops.sh_scatter(model.l_max, pg.rowptr, pg.geom, Nt, 1.0 / k, attrs)
    hr = torch.empty((Nt, tpr.dxp), **f32)
    _mm(Nt, tpr.dxp, D, 1.0, h, 0, D, wexp, plan.P.off, tpr.dxp, hr, 0, tpr.dxp)
    T0 = torch.empty((Nt, tpr.d_live), **f32)
    ops.node_tp_fwd(ro_id, hr, attrs, Nt, T0)
    dh = model.d_hidden
    # The Linear to d_hidden scalars and the sum / mean over the nodes commute (nequip.py:195-199): pool the
    # D_live-wide tensor-product output first, then apply the Linear to B rows instead of B*N rows.
    pooled = torch.empty((B, dh), **f32)
    ro_arg = None
    if model.readout_agg == "max":
        # jnp.max does not commute with the Linear: per-node scalars first, then value + arg-max over the nodes
        T0p = torch.empty((Nt, dh), **f32)
        _mm(Nt, dh, tpr.d_live, 1.0, T0, 0, tpr.d_live, wexp, plan.Wro.off, dh, T0p, 0, dh)
        ro_arg = torch.empty((B, dh), dtype=torch.int32, device=dev)
        ops.pool_max_fwd(T0p, B, N, dh, pooled, ro_arg)
        T0p = T0                                       # the weight gradient's left operand
    else:
        T0p = torch.empty((B, tpr.d_live), **f32)
        ops.pool_fwd(T0, B, N, tpr.d_live, 1 if model.readout_agg == "mean" else 0, T0p)
        _mm(B, dh, tpr.d_live, 1.0, T0p, 0, tpr.d_live, wexp, plan.Wro.off, dh, pooled, 0, dh)
    ln_stats = None
    a = pooled
    if plan.ln is not None:
        if globals_ is None or globals_.shape != (B, plan.n_globals):
            raise ValueError(f"graphs.globals must have shape ({B}, {plan.n_globals})")
        a = torch.empty((B, dh + plan.n_globals), **f32)
        ln_stats = torch.empty((B, 2), **f32)
        ops.layernorm_fwd(pooled, globals_, theta, plan.ln[0], plan.ln[1], 1e-6, a, ln_stats)
    zs, pro = [], 0
    ro_in = a
    for (ko, bo, fan, hdim) in plan.ro_mlp:
        zl = torch.empty((B, hdim), **f32)
        _mm(B, hdim, fan, 1.0, a, 0, fan, theta, ko, hdim, zl, 0, hdim, pro=pro, pro_scale=1.0, epi=1, aux=theta,
            aux_off=bo)
        zs.append(zl)
        a, pro = zl, 1
    saved.update(attrs=attrs, hr=hr, T0p=T0p, pooled=pooled, ro_z=zs, ro_in=ro_in, ln_stats=ln_stats, globals=globals_,
                 ro_arg=ro_arg)
    return zs[-1], saved


def _backward(model: NequIP, plan: _Plan, theta: torch.Tensor, pg: PreparedGraph, saved, gout: torch.Tensor,
              k: float, gtheta: torch.Tensor):
    dev = theta.device
    B, N = pg.B, pg.N
    Nt, Et = B * N, B * pg.E
    D = plan.d_in
    f32 = dict(dtype=torch.float32, device=dev)
    tab = plan.device_tables(dev)
    wexp = saved["wexp"]
    gwexp = torch.zeros((plan._n_exp,), **f32)
    gtheta.zero_()
    tp_id = ops.tp_lookup(plan.tp_sig)
    dxp, n_live, d_live = plan.tp.dxp, plan.tp.n_live, plan.tp.d_live
    agg = ops.AGG[model.message_passing_agg]
    inv_sqrt_k = 1.0 / math.sqrt(k)
    h_final = saved["h_final"]

    if model.task == "node":
        go = gout.reshape(Nt, -1)
        if plan.Wout is not None:
            dh = torch.empty((Nt, D), **f32)
            _mm(Nt, D, plan.d_out, 1.0, go, 0, plan.d_out, wexp, plan.Wout.off, plan.d_out, dh, 0, D, tb=True)
            _mm(D, plan.d_out, Nt, 1.0, h_final, 0, D, go, 0, plan.d_out, gwexp, plan.Wout.off, plan.d_out, ta=True)
        else:
            dh = go.clone()
    else:
        tpr = plan.tp_ro
        ro_id = ops.tp_lookup(tpr.signature)
        zs = saved["ro_z"]
        dz = gout.reshape(B, -1).contiguous()
        nl = len(plan.ro_mlp)
        for l in range(nl - 1, -1, -1):
            ko, bo, fan, hdim = plan.ro_mlp[l]
            a_prev = zs[l - 1] if l > 0 else saved["ro_in"]
            # dK = act(z_{l-1})^T dz ; db = colsum(dz)
            _mm(fan, hdim, B, 1.0, a_prev, 0, fan, dz, 0, hdim, gtheta, ko, hdim, ta=True, pro=1 if l > 0 else 0,
                pro_scale=1.0)
            ops.colsum(dz, B, hdim, gtheta, y_off=bo)
            dprev = torch.empty((B, fan), **f32)
            if l > 0:
                _mm(B, fan, hdim, 1.0, dz, 0, hdim, theta, ko, hdim, dprev, 0, fan, tb=True, epi=2, aux=zs[l - 1],
                    ld_aux=fan, epi_scale=1.0)
            else:
                _mm(B, fan, hdim, 1.0, dz, 0, hdim, theta, ko, hdim, dprev, 0, fan, tb=True)
            dz = dprev
        dhid = model.d_hidden
        if plan.ln is not None:
            d_pool = torch.empty((B, dhid), **f32)
            ops.layernorm_bwd(saved["pooled"], saved["globals"], theta, plan.ln[0], saved["ln_stats"], dz, d_pool,
                              gtheta, plan.ln[0], plan.ln[1])
            dz = d_pool
        # pooled = pool(T0) @ Wro: weight gradient and data gradient on B rows, then un-pool
        T0p = saved["T0p"]
        dT0 = torch.empty((Nt, tpr.d_live), **f32)
        if model.readout_agg == "max":
            # the gradient of every pooled scalar goes to the node that attained the maximum
            dL = torch.empty((Nt, dhid), **f32)
            ops.pool_max_bwd(dz, saved["ro_arg"], B, N, dhid, dL)
            _mm(tpr.d_live, dhid, Nt, 1.0, T0p, 0, tpr.d_live, dL, 0, dhid, gwexp, plan.Wro.off, dhid, ta=True)
            _mm(Nt, tpr.d_live, dhid, 1.0, dL, 0, dhid, wexp, plan.Wro.off, dhid, dT0, 0, tpr.d_live, tb=True)
        else:
            _mm(tpr.d_live, dhid, B, 1.0, T0p, 0, tpr.d_live, dz, 0, dhid, gwexp, plan.Wro.off, dhid, ta=True)
            dT0p = torch.empty((B, tpr.d_live), **f32)
            _mm(B, tpr.d_live, dhid, 1.0, dz, 0, dhid, wexp, plan.Wro.off, dhid, dT0p, 0, tpr.d_live, tb=True)
            ops.pool_bwd(dT0p, B, N, tpr.d_live, 1 if model.readout_agg == "mean" else 0, dT0)
        dhr = torch.empty((Nt, tpr.dxp), **f32)
        ops.node_tp_bwd(ro_id, saved["hr"], saved["attrs"], dT0, Nt, dhr)
        dh = torch.empty((Nt, D), **f32)
        _mm(Nt, D, tpr.dxp, 1.0, dhr, 0, tpr.dxp, wexp, plan.P.off, tpr.dxp, dh, 0, D, tb=True)

    dw_t = torch.empty((n_live, Et), **f32)
    dxe = torch.empty((Et, dxp), **f32)
    shared = pg.pairs is not None and _fused_radial_ok(model, plan)
    dw_rows = torch.empty((Et, (n_live + 3) // 4 * 4), **f32) if shared else None
    for st, sv in zip(reversed(plan.steps), reversed(saved["steps"])):
        h, xt, Zs, w_t, A, z, g = sv["h"], sv["xt"], sv["Zs"], sv["w_t"], sv["A"], sv["z"], sv["g"]
        # h' = g @ W3
        dA = torch.empty((Nt, d_live), **f32)
        dh_prev = torch.empty((Nt, D), **f32)
        if model.fuse_node_update:
            ops.node_update_bwd(A, d_live, d_live, h, D, wexp, st["W1"].off, inv_sqrt_k,
                                st["W2"].off if model.residual else None, plan.n_extra, plan.n_gates, plan.gate_mul,
                                plan.gate_l, plan.inv_c_act, plan.inv_c_gate, st["W3"].off, z, dh, Nt, dA, dh_prev,
                                gwexp)
        else:
            # h' = g @ W3
            dg = torch.empty((Nt, plan.d_g), **f32)
            _mm(Nt, plan.d_g, D, 1.0, dh, 0, D, wexp, st["W3"].off, D, dg, 0, plan.d_g, tb=True)
            _mm(plan.d_g, D, Nt, 1.0, g, 0, plan.d_g, dh, 0, D, gwexp, st["W3"].off, D, ta=True)
            dz = torch.empty((Nt, plan.d_z), **f32)
            ops.gate_bwd(z, dg, Nt, plan.n_extra, plan.n_gates, plan.gate_mul, plan.gate_l, plan.inv_c_act,
                         plan.inv_c_gate, dz)
            # z = A @ W1 / sqrt(k) + h @ W2
            _mm(Nt, d_live, plan.d_z, inv_sqrt_k, dz, 0, plan.d_z, wexp, st["W1"].off, plan.d_z, dA, 0, d_live, tb=True)
            _mm(d_live, plan.d_z, Nt, inv_sqrt_k, A, 0, d_live, dz, 0, plan.d_z, gwexp, st["W1"].off, plan.d_z, ta=True)
            if model.residual:
                _mm(Nt, D, plan.d_z, 1.0, dz, 0, plan.d_z, wexp, st["W2"].off, plan.d_z, dh_prev, 0, D, tb=True)
                _mm(D, plan.d_z, Nt, 1.0, h, 0, D, dz, 0, plan.d_z, gwexp, st["W2"].off, plan.d_z, ta=True)
            else:
                dh_prev.zero_()
        # interaction block
        if shared:
            ops.tpconv_bwd_shared(tp_id, pg.rowptr, pg.src, pg.perm, pg.geom, pg.pairs, w_t, xt, Nt, agg, dA, d_live,
                                  dw_rows, dxe, sv["amax"])
        else:
            ops.tpconv_bwd(tp_id, pg.rowptr, pg.src, pg.perm, pg.geom, w_t, xt, Nt, agg, dA, d_live, dw_t, dxe, sv["amax"])
        # sender side: dxt = sum over each sender's edges -- right behind the kernel that wrote dxe, while the tail of
        # it is still in L2 (the radial-MLP backward in between streams gigabytes)
        dxt = torch.empty((Nt, dxp), **f32)
        ops.segment_gather_sum(dxe, pg.rowptr_s, pg.perm_s, Nt, dxp, dxt)
        # radial MLP backward
        mlp = st["mlp"]
        nh = len(mlp) - 1
        _, fan, _ = mlp[-1]
        al = 1.0 / math.sqrt(fan)
        fused = _fused_radial_ok(model, plan)
        if shared:
            # cotangents of the two directions of a pair are summed, then one backward pass per unique edge length
            ops.pair_combine(dw_rows, n_live, pg.pairs, dw_t)
            ops.radial_mlp_bwd_rows(pg.pairs, model.n_radial_basis, model.r_cutoff, _radial_weights(theta, wexp, st, n_live),
                                    n_live, plan.inv_c_mlp, dw_t, _radial_weights(gtheta, gwexp, st, n_live),
                                    z_stash=sv["stash"], activation=model.activation)
        elif fused:
            ops.radial_mlp_bwd(pg.geom, model.n_radial_basis, model.r_cutoff, _radial_weights(theta, wexp, st, n_live),
                               n_live, plan.inv_c_mlp, dw_t, _radial_weights(gtheta, gwexp, st, n_live),
                               z_stash=sv["stash"], radial=pg.radial if sv["stash"] is not None else None,
                               activation=model.activation)
        a_last, pro_last = (Zs[-1], 1) if nh > 0 and not fused else (pg.radial, 0)
        # dW4L = al * act(Z_last)^T @ dw_t^T
        if not fused:
            _mm(fan, n_live, Et, al, a_last, 0, fan, dw_t, 0, Et, gwexp, st["W4L"].off, n_live, ta=True, tb=True,
                pro=pro_last, pro_scale=plan.inv_c_act, tag="radial_mlp_bwd", tf32x3=model.tensor_cores)
        if nh > 0 and not fused:
            dZ = torch.empty((Et, fan), **f32)
            _mm(Et, fan, n_live, al, dw_t, 0, Et, wexp, st["W4L"].off, n_live, dZ, 0, fan, ta=True, tb=True,
                epi=2, aux=Zs[-1], ld_aux=fan, epi_scale=plan.inv_c_act, tag="radial_mlp_bwd", tf32x3=model.tensor_cores)
            for l in range(nh - 1, -1, -1):
                koff, fan_l, hdim = mlp[l]
                al = 1.0 / math.sqrt(fan_l)
                a_prev, pro_prev = (Zs[l - 1], 1) if l > 0 else (pg.radial, 0)
                _mm(fan_l, hdim, Et, al, a_prev, 0, fan_l, dZ, 0, hdim, gtheta, koff, hdim, ta=True,
                    pro=pro_prev, pro_scale=plan.inv_c_act, tag="radial_mlp_bwd", tf32x3=model.tensor_cores)
                if l > 0:
                    dZp = torch.empty((Et, fan_l), **f32)
                    _mm(Et, fan_l, hdim, al, dZ, 0, hdim, theta, koff, hdim, dZp, 0, fan_l, tb=True,
                        epi=2, aux=Zs[l - 1], ld_aux=fan_l, epi_scale=plan.inv_c_act, tag="radial_mlp_bwd", tf32x3=model.tensor_cores)
                    dZ = dZp
        # sender side, through W0x
        _mm(Nt, D, dxp, 1.0, dxt, 0, dxp, wexp, st["W0x"].off, dxp, dh_prev, 0, D, tb=True, accumulate=True)
        _mm(D, dxp, Nt, 1.0, h, 0, D, dxt, 0, dxp, gwexp, st["W0x"].off, dxp, ta=True)
        dh = dh_prev
    ops.weights_contract(gwexp, tab["con_pidx"], tab["con_ptr"], tab["con_tidx"], tab["con_scale"], gtheta)
    return dh
