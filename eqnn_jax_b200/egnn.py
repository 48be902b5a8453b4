"""EGNN on B200 kernels, behind the reference's module interface.

Drop-in for ``models/egnn.py`` (reference).  The configuration the cosmology benchmark and the example notebook
use is ``positions_only=True`` on 3-d nodes (benchmarks/galaxies/train_cosmology.py:67-81 with feats='pos';
notebooks/examples.ipynb cell 14); the other node layouts are described below.  Same dataclass fields and defaults (:240-255), ``GraphsTuple`` in ->
``GraphsTuple`` (task="node") or ``(B, n_outputs)`` (task="graph") out, Flax parameter names (MLP_i / Dense_s).

Per message-passing step, edges in receiver-sorted order:
    r_ij, |r_ij|, bessel      eqnn_egnn_geom_fwd                       egnn.py:82-90
    m   = phi_e(bessel)       tcgen05 3xTF32 GEMMs (+bias, gelu)        egnn.py:63-67,98
    t   = Dense_1(phi_x(m))   tcgen05 3xTF32 GEMMs                      egnn.py:70-79,107
    x'  = wrap(x + agg clip(r_ij t))   eqnn_egnn_coord_fwd              egnn.py:110-114,146-150 + jraph segment reduce
With 3-d nodes the aggregated messages m_i are never read (egnn.py:143-150), so they are not aggregated
(XLA removes that dead reduction under jit as well).  soft_edges (egnn.py:101-105) is supported.

The (x, v, h) layout -- ``positions_only=False`` with >= 7 node features, the configuration of the reference's own
equivariance test (tests/test_equivariance.py:224-250) -- is supported as well: the edge MLP reads
[bessel | h_sender | h_receiver] (egnn.py:47-60), the node update is v' = phi_v(h) v + agg, x' = wrap(phi_xx(h) x + v')
(or x + v'), h' = h + phi_h(h), and (x', v, h') is returned (egnn.py:198-229); m_i is unused there too.
The (x, h) layout (``positions_only=True`` with scalar attributes: train_cosmology.py --feats all) reads the
aggregated messages: h' = h + phi_h([h, m_i]) (egnn.py:152-166), m_i = segment reduce of m_ij.  The (x, v) layout
(6 features) feeds m_i to phi_v / phi_xx and returns (x', v, m_i): the node width grows to 6 + d_hidden and the
following steps run as (x, v, h) with d_hidden scalars (egnn.py:171-196) -- all four layouts of the reference.
Graph globals (broadcast to every edge and node by jraph, concatenated to the pooled features and LayerNorm-ed
before the read-out, egnn.py:41-60,163-164,211-212,332-336) are supported for the (x), (x, h) and (x, v, h) layouts.
Aggregations: sum, mean and max (jraph.segment_max with the gradient routed to the winning edge; jnp.max read-out).
Not built yet: globals with the (x, v) layout -> NotImplementedError, never a silent fallback.
"""
from __future__ import annotations

import dataclasses
import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import ops
from .graph_utils import ApplyPBC, GraphsTuple
from .nequip import NequIPParams as FlatParams
from .nequip import _mm


@dataclasses.dataclass
class EGNN:
    """Field names and defaults of the reference module (models/egnn.py:240-255)."""
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
    apply_pbc: Optional[ApplyPBC] = None
    tensor_cores: bool = True     # not a reference field: tcgen05 3xTF32 GEMMs (False: fp32 CUDA-core GEMMs)

    # ---- static layout ------------------------------------------------------------------
    def _variant(self, n_feat: int) -> Tuple[bool, int]:
        """(has velocities, number of scalar attributes) for a node-feature width."""
        if self.positions_only:
            if n_feat < 3:
                raise ValueError("positions_only=True needs >= 3 node features")
            return False, n_feat - 3
        if n_feat < 6:
            raise ValueError("positions_only=False needs >= 6 node features (x, v)")
        return True, n_feat - 6

    def _check(self, n_feat: int) -> None:
        self._variant(n_feat)
        if self.activation != "gelu":
            raise NotImplementedError("only activation='gelu' has a B200 kernel on this path")
        if self.message_passing_agg not in ("sum", "mean", "max") or self.readout_agg not in ("sum", "mean", "max"):
            raise ValueError("aggregations must be 'sum', 'mean' or 'max' (models/egnn.py:282-284,326)")
        if self.n_layers < 2:
            raise NotImplementedError("n_layers >= 2 expected")

    def _layout(self, n_feat: int = 3, n_glob: int = 0):
        d, L = self.d_hidden, self.n_layers
        nf = self.n_radial_basis if self.n_radial_basis > 0 else 1
        layout: List[Tuple[Tuple[str, ...], Tuple[int, ...], int]] = []
        off = 0

        def add(path, shape):
            nonlocal off
            layout.append((path, tuple(shape), off))
            o = off
            off += int(np.prod(shape))
            return o

        def add_mlp(mi, fan, widths):
            out = []
            for l, w in enumerate(widths):
                out.append((add((f"MLP_{mi}", f"Dense_{l}", "kernel"), (fan, w)),
                            add((f"MLP_{mi}", f"Dense_{l}", "bias"), (w,)), fan, w))
                fan = w
            return out

        steps, mi, Fs = [], 0, n_feat
        for s in range(self.message_passing_steps):
            vel, dh = self._variant(Fs)
            st = {"phi_e": [], "phi_x": [], "F": Fs, "vel": vel, "dh": dh}
            if vel and dh == 0 and n_glob:
                raise NotImplementedError("B200 EGNN: graph globals with the (x, v) layout are not built")
            fan = nf + 2 * dh + n_glob                        # [bessel | h_s | h_r | globals], egnn.py:41-60,93-96
            for l in range(L):
                st["phi_e"].append((add((f"MLP_{mi}", f"Dense_{l}", "kernel"), (fan, d)),
                                    add((f"MLP_{mi}", f"Dense_{l}", "bias"), (d,)), fan))
                fan = d
            mi += 1
            for l in range(L - 1):
                st["phi_x"].append((add((f"MLP_{mi}", f"Dense_{l}", "kernel"), (d, d)),
                                    add((f"MLP_{mi}", f"Dense_{l}", "bias"), (d,)), d))
            mi += 1
            st["k_last"] = add((f"Dense_{s}", "kernel"), (d, 1))
            # phi_a is constructed in every step (takes the name) but owns parameters only with soft_edges
            st["phi_a"] = (add((f"MLP_{mi}", "Dense_0", "kernel"), (d, d)),
                           add((f"MLP_{mi}", "Dense_0", "bias"), (d,))) if self.soft_edges else None
            mi += 1
            if not vel and dh > 0:                           # node function, egnn.py:159-166: phi_h([h_i, m_i])
                st["phi_h"] = add_mlp(mi, dh + d + n_glob, [d] * (L - 1) + [dh]); mi += 1
            if vel:                                          # node function, egnn.py:171-229
                n_in = (dh if dh > 0 else d) + n_glob        # concats = h_i, or m_i when there are no scalars (+ globals)
                st["phi_v"] = add_mlp(mi, n_in, [d] * (L - 1) + [1]); mi += 1
                st["phi_h"] = None
                if dh > 0:
                    st["phi_h"] = add_mlp(mi, n_in, [d] * (L - 1) + [dh]); mi += 1
                st["phi_xx"] = None
                if self.decouple_pos_vel_updates:
                    st["phi_xx"] = add_mlp(mi, n_in, [d] * (L - 1) + [1]); mi += 1
                if dh == 0:
                    Fs = 6 + d                               # (x', v, m_i) is returned, egnn.py:196
            st["F_out"] = Fs
            steps.append(st)
        readout = []
        if self.task == "graph":
            n_pool = Fs + n_glob                              # globals + LayerNorm before the read-out, egnn.py:332-336
            self._ln = None
            if n_glob:
                self._ln = (add(("LayerNorm_0", "scale"), (n_pool,)), add(("LayerNorm_0", "bias"), (n_pool,)))
            ws = [self.mlp_readout_widths[0] * n_pool] + [w * d for w in self.mlp_readout_widths[1:]] + [self.n_outputs]
            readout = add_mlp(mi, n_pool, ws)
        return layout, steps, readout, off

    def init(self, rng: Union[int, np.random.Generator], graphs: GraphsTuple) -> FlatParams:
        self._check(graphs.nodes.shape[-1])
        n_glob = 0 if graphs.globals is None else int(graphs.globals.shape[-1])
        layout, steps, readout, n = self._layout(graphs.nodes.shape[-1], n_glob)
        rng = np.random.default_rng(rng) if not isinstance(rng, np.random.Generator) else rng
        host = np.zeros((n,), dtype=np.float32)
        for path, shape, off in layout:
            if path[0] == "LayerNorm_0":
                host[off:off + int(np.prod(shape))] = 1.0 if path[-1] == "scale" else 0.0
                continue
            if path[-1] == "bias":
                continue
            if path[0].startswith("Dense_"):       # variance_scaling(1e-2, fan_in, uniform), egnn.py:73-78
                lim = math.sqrt(3.0 * 1e-2 / shape[0])
            else:                                  # xavier_uniform, models/mlp.py:19
                lim = math.sqrt(6.0 / (shape[0] + shape[1]))
            host[off:off + int(np.prod(shape))] = rng.uniform(-lim, lim, size=shape).reshape(-1)
        return FlatParams(layout, torch.from_numpy(host).to(graphs.nodes.device))

    # ---- forward / backward -----------------------------------------------------------------
    def _prepare(self, graphs: GraphsTuple):
        x = graphs.nodes
        if x.dim() == 2:
            x = x.unsqueeze(0)
        B, N, F = x.shape
        self._check(F)
        glob = None
        if graphs.globals is not None:
            if not isinstance(graphs.globals, torch.Tensor) or not graphs.globals.is_cuda:
                raise RuntimeError("graphs.globals must be a CUDA tensor (this path has no CPU implementation)")
            glob = graphs.globals.reshape(B, -1).to(torch.float32).contiguous()
        senders = graphs.senders.reshape(B, -1).contiguous()
        receivers = graphs.receivers.reshape(B, -1).contiguous()
        rowptr, perm, src = ops.csr_build(senders, receivers, N)
        rowptr_s, perm_s, _ = ops.csr_build(receivers, senders, N)
        return x.contiguous(), B, N, senders.shape[1], (rowptr, perm, src, rowptr_s, perm_s, glob)

    def forward_backward(self, params: FlatParams, graphs: GraphsTuple, grad_fn=None):
        """Forward (and, if ``grad_fn(out) -> dL/dout`` is given, backward into ``params.grad``)."""
        x, B, N, E, csr = self._prepare(graphs)
        out, saved = self._forward(params.flat, x, B, N, E, csr)
        if grad_fn is not None:
            self._backward(params.flat, params.grad, saved, grad_fn(out).contiguous(), B, N, E, csr)
        return out

    def apply(self, params: FlatParams, graphs: GraphsTuple):
        out = _EGNNFunction.apply(params.flat, self, graphs)
        if self.task == "graph":
            return out
        return graphs._replace(nodes=out)

    __call__ = apply

    # ---- N-row MLPs of the node function (models/mlp.py: Dense + bias, gelu between layers) ----------------
    def _mlp_fwd(self, theta, layers, a, a_off, lda, M):
        zs, pro = [], 0
        for (ko, bo, fan, w) in layers:
            Z = torch.empty((M, w), dtype=torch.float32, device=theta.device)
            _mm(M, w, fan, 1.0, a, a_off, lda, theta, ko, w, Z, 0, w, pro=pro, pro_scale=1.0, epi=1, aux=theta, aux_off=bo,
                tag="egnn_node_mlp", tf32x3=self.tensor_cores)
            zs.append(Z)
            a, a_off, lda, pro = Z, 0, w, 1
        return zs

    def _mlp_bwd(self, theta, gtheta, layers, zs, a0, a0_off, lda0, M, dz, din=None):
        """dz: dense (M, width_last).  din = (tensor, offset, ld): the input gradient is ACCUMULATED there."""
        for l in range(len(layers) - 1, -1, -1):
            ko, bo, fan, w = layers[l]
            a_prev, ap_off, ap_ld, pro = (zs[l - 1], 0, fan, 1) if l > 0 else (a0, a0_off, lda0, 0)
            _mm(fan, w, M, 1.0, a_prev, ap_off, ap_ld, dz, 0, w, gtheta, ko, w, ta=True, pro=pro, pro_scale=1.0,
                tag="egnn_node_mlp", tf32x3=self.tensor_cores)
            ops.colsum_tall(dz, M, w, gtheta, y_off=bo)
            if l > 0:
                dprev = torch.empty((M, fan), dtype=torch.float32, device=theta.device)
                _mm(M, fan, w, 1.0, dz, 0, w, theta, ko, w, dprev, 0, fan, tb=True, epi=2, aux=zs[l - 1], ld_aux=fan,
                    epi_scale=1.0, tag="egnn_node_mlp", tf32x3=self.tensor_cores)
                dz = dprev
            elif din is not None:
                t, off, ld = din
                _mm(M, fan, w, 1.0, dz, 0, w, theta, ko, w, t, off, ld, tb=True, accumulate=True, tag="egnn_node_mlp")

    def _forward(self, theta, x, B, N, E, csr):
        rowptr, perm, src, rowptr_s, perm_s, glob = csr
        G_ = 0 if glob is None else glob.shape[1]
        dev = x.device
        f32 = dict(dtype=torch.float32, device=dev)
        layout, steps, readout, _ = self._layout(x.shape[-1], G_)
        Nt, Et = B * N, B * E
        d = self.d_hidden
        nf = self.n_radial_basis if self.n_radial_basis > 0 else 1
        cell = None if self.apply_pbc is None else self.apply_pbc.cell
        cinv = None if self.apply_pbc is None else self.apply_pbc.cell_inv
        agg = ops.AGG[self.message_passing_agg]
        tc = self.tensor_cores
        nodes = x.reshape(Nt, x.shape[-1])
        saved = {"steps": []}
        for st in steps:
            F, vel, dh, F_out = st["F"], st["vel"], st["dh"], st["F_out"]
            K0 = nf + 2 * dh + G_
            ho = 6 if vel else 3                             # first scalar-attribute column
            rij4 = torch.empty((Et, 4), **f32)
            feats = torch.empty((Et, nf), **f32)
            ops.egnn_geom_fwd(nodes, F, Nt, rowptr, src, cell, cinv, self.n_radial_basis, self.r_max, rij4, feats)
            U = feats
            if K0 != nf:                                     # [bessel | h_sender | h_receiver | globals], egnn.py:41-60,93-96
                U = torch.empty((Et, K0), **f32)
                ops.egnn_edge_input_fwd(feats, nf, nodes, F, ho, dh, rowptr, src, Nt, U, glob=glob, nodes_per_graph=N)
            acts, a, pro = [], U, 0
            Za = Mg = None
            nE = len(st["phi_e"])
            for li, (ko, bo, fan) in enumerate(st["phi_e"] + st["phi_x"]):
                if li == nE and st["phi_a"] is not None:
                    # soft edges (egnn.py:101-105): m <- m * sigmoid(phi_a(m)), m = gelu(Z_L)
                    ka, ba = st["phi_a"]
                    Za = torch.empty((Et, d), **f32)
                    _mm(Et, d, d, 1.0, a, 0, d, theta, ka, d, Za, 0, d, pro=1, pro_scale=1.0, epi=1, aux=theta,
                        aux_off=ba, tag="egnn_mlp_fwd", tf32x3=tc)
                    Mg = torch.empty((Et, d), **f32)
                    ops.softgate_fwd(a, Za, Mg)
                    a, pro = Mg, 0
                Z = torch.empty((Et, d), **f32)
                _mm(Et, d, fan, 1.0, a, 0, fan, theta, ko, d, Z, 0, d, pro=pro, pro_scale=1.0, epi=1, aux=theta,
                    aux_off=bo, tag="egnn_mlp_fwd", tf32x3=tc)
                acts.append(Z)
                a, pro = Z, 1
            t = torch.empty((Et,), **f32)
            _mm(Et, 1, d, 1.0, a, 0, d, theta, st["k_last"], 1, t, 0, Et, tc=True, pro=1, pro_scale=1.0,
                tag="egnn_mlp_fwd", tf32x3=tc)
            nodes_next = torch.empty((Nt, F_out), **f32)
            # jraph.segment_max (egnn.py:282-284): winners of the translations (3 components) and of the messages
            ax3 = torch.empty((Nt, 3), dtype=torch.int32, device=dev) if agg == 2 else None
            am = torch.empty((Nt, d), dtype=torch.int32, device=dev) if agg == 2 else None
            sv = dict(rij4=rij4, U=U, acts=acts, t=t, Za=Za, Mg=Mg, nodes=nodes, ax3=ax3, am=am)
            gated = st["phi_a"] is not None
            if not vel:
                ops.egnn_coord_fwd(rij4, t, self.tanh_out, agg, rowptr, Nt, nodes, F, cell, cinv, nodes_next, F, argmax=ax3)
                if dh > 0:                                   # (x, h): h' = h + phi_h([h, m_i]), egnn.py:152-166
                    wc = dh + d + G_                         # [h | m_i | globals]
                    Cc = torch.empty((Nt, wc), **f32)
                    ops.copy_cols(nodes, F, 3, dh, Nt, Cc, wc, 0)
                    ops.egnn_segment_rows_fwd(Mg if gated else acts[nE - 1], d, not gated, rowptr, Nt, agg, Cc, wc,
                                              col_off=dh, argmax=am)
                    if G_:
                        ops.copy_cols(glob, G_, 0, G_, Nt, Cc, wc, dh + d, group=N)
                    zh = self._mlp_fwd(theta, st["phi_h"], Cc, 0, wc, Nt)
                    ops.egnn_node_assemble(nodes, F, 3, zh[-1], Nt, nodes_next)
                    sv.update(Cc=Cc, zh=zh)
            else:
                if dh > 0 and G_:                            # concats = [h_i | globals] (egnn.py:207-212)
                    cin, c_off, c_ld = torch.empty((Nt, dh + G_), **f32), 0, dh + G_
                    ops.copy_cols(nodes, F, 6, dh, Nt, cin, dh + G_, 0)
                    ops.copy_cols(glob, G_, 0, G_, Nt, cin, dh + G_, dh, group=N)
                elif dh > 0:                                 # concats = h_i (egnn.py:207)
                    cin, c_off, c_ld = nodes, 6, F
                else:                                        # concats = m_i (egnn.py:180), returned as new scalars
                    cin, c_off, c_ld = torch.empty((Nt, d), **f32), 0, d
                    ops.egnn_segment_rows_fwd(Mg if gated else acts[nE - 1], d, not gated, rowptr, Nt, agg, cin, d, argmax=am)
                zv = self._mlp_fwd(theta, st["phi_v"], cin, c_off, c_ld, Nt)
                zh = self._mlp_fwd(theta, st["phi_h"], cin, c_off, c_ld, Nt) if dh > 0 else None
                zx = self._mlp_fwd(theta, st["phi_xx"], cin, c_off, c_ld, Nt) if st["phi_xx"] is not None else None
                base = torch.empty((Nt, 3), **f32)
                ops.egnn_xv_base(nodes, F, None if zx is None else zx[-1], zv[-1], Nt, base)
                ops.egnn_coord_fwd(rij4, t, self.tanh_out, agg, rowptr, Nt, base, 3, cell, cinv, nodes_next, F_out, argmax=ax3)
                if dh > 0:
                    ops.egnn_node_assemble(nodes, F, 6, zh[-1], Nt, nodes_next)
                else:
                    ops.copy_cols(nodes, F, 3, 3, Nt, nodes_next, F_out, 3)
                    ops.copy_cols(cin, d, 0, d, Nt, nodes_next, F_out, 6)
                sv.update(zv=zv, zh=zh, zx=zx, cin=cin, c_off=c_off, c_ld=c_ld)
            saved["steps"].append(sv)
            nodes = nodes_next
        F = nodes.shape[-1]
        if self.task == "node":
            return nodes.view(B, N, F), saved
        pooled = torch.empty((B, F), **f32)
        ro_arg = None
        if self.readout_agg == "max":                        # jnp.max over the nodes (egnn.py:326): value + arg-max
            ro_arg = torch.empty((B, F), dtype=torch.int32, device=dev)
            ops.pool_max_fwd(nodes, B, N, F, pooled, ro_arg)
        else:
            ops.pool_fwd(nodes, B, N, F, 1 if self.readout_agg == "mean" else 0, pooled)
        a, ln_stats = pooled, None
        if G_:
            a = torch.empty((B, F + G_), **f32)
            ln_stats = torch.empty((B, 2), **f32)
            ops.layernorm_fwd(pooled, glob, theta, self._ln[0], self._ln[1], 1e-6, a, ln_stats)
        ro_in = a
        zs, pro = [], 0
        for (ko, bo, fan, h) in readout:
            zl = torch.empty((B, h), **f32)
            _mm(B, h, fan, 1.0, a, 0, fan, theta, ko, h, zl, 0, h, pro=pro, pro_scale=1.0, epi=1, aux=theta, aux_off=bo)
            zs.append(zl)
            a, pro = zl, 1
        saved.update(pooled=pooled, ro_z=zs, ro_in=ro_in, ln_stats=ln_stats, ro_arg=ro_arg)
        return zs[-1], saved

    def _backward(self, theta, gtheta, saved, gout, B, N, E, csr):
        rowptr, perm, src, rowptr_s, perm_s, glob = csr
        G_ = 0 if glob is None else glob.shape[1]
        dev = theta.device
        f32 = dict(dtype=torch.float32, device=dev)
        layout, steps, readout, _ = self._layout(saved["steps"][0]["nodes"].shape[-1], G_)
        Nt, Et = B * N, B * E
        d = self.d_hidden
        nf = self.n_radial_basis if self.n_radial_basis > 0 else 1
        F = steps[-1]["F_out"]                               # width of the final node features
        agg = ops.AGG[self.message_passing_agg]
        tc = self.tensor_cores
        gtheta.zero_()
        if self.task == "node":
            dN = gout.reshape(Nt, F).clone()
        else:
            zs = saved["ro_z"]
            dz = gout.reshape(B, -1).contiguous()
            for l in range(len(readout) - 1, -1, -1):
                ko, bo, fan, h = readout[l]
                a_prev = zs[l - 1] if l > 0 else saved["ro_in"]
                _mm(fan, h, B, 1.0, a_prev, 0, fan, dz, 0, h, gtheta, ko, h, ta=True, pro=1 if l > 0 else 0, pro_scale=1.0)
                ops.colsum(dz, B, h, gtheta, y_off=bo)
                dprev = torch.empty((B, fan), **f32)
                if l > 0:
                    _mm(B, fan, h, 1.0, dz, 0, h, theta, ko, h, dprev, 0, fan, tb=True, epi=2, aux=zs[l - 1], ld_aux=fan,
                        epi_scale=1.0)
                else:
                    _mm(B, fan, h, 1.0, dz, 0, h, theta, ko, h, dprev, 0, fan, tb=True)
                dz = dprev
            if G_:
                d_pool = torch.empty((B, F), **f32)
                ops.layernorm_bwd(saved["pooled"], glob, theta, self._ln[0], saved["ln_stats"], dz, d_pool, gtheta,
                                  self._ln[0], self._ln[1])
                dz = d_pool
            dN = torch.empty((Nt, F), **f32)
            if self.readout_agg == "max":
                ops.pool_max_bwd(dz, saved["ro_arg"], B, N, F, dN)
            else:
                ops.pool_bwd(dz, B, N, F, 1 if self.readout_agg == "mean" else 0, dN)
        dt = torch.empty((Et,), **f32)
        dre = torch.empty((Et, 4), **f32)
        for si in range(len(steps) - 1, -1, -1):
            st, sv = steps[si], saved["steps"][si]
            rij4, U, acts, t, nodes = sv["rij4"], sv["U"], sv["acts"], sv["t"], sv["nodes"]
            F, vel, dh, F_out = st["F"], st["vel"], st["dh"], st["F_out"]      # dN is (Nt, F_out) here
            K0 = nf + 2 * dh + G_
            nE = len(st["phi_e"])
            g3 = dN
            if F_out != 3:                                   # dL/dx' as a dense (Nt, 3) array
                g3 = torch.empty((Nt, 3), **f32)
                ops.copy_cols(dN, F_out, 0, 3, Nt, g3, 3, 0)
            ops.egnn_coord_bwd(rij4, t, self.tanh_out, agg, rowptr, perm, Nt, g3, self.n_radial_basis, self.r_max,
                               None, dt, None, None, argmax=sv["ax3"])
            dC, dC_ld, dC_off = None, 0, 0                    # dL/dm_i lives in dC[:, dC_off:], row stride dC_ld
            if vel and dh == 0:
                # (x, v) node function first: m_i feeds phi_v / phi_xx and is returned as the new scalar attributes
                zv, zx = sv["zv"], sv["zx"]
                dsv = torch.empty((Nt, 1), **f32)
                dsx = torch.empty((Nt, 1), **f32) if zx is not None else None
                ops.egnn_xv_bwd(nodes, F, None if zx is None else zx[-1], zv[-1], dN, F_out, None, Nt, None, dsx, dsv)
                dC, dC_ld, dC_off = torch.empty((Nt, d), **f32), d, 0
                ops.copy_cols(dN, F_out, 6, d, Nt, dC, d, 0)
                self._mlp_bwd(theta, gtheta, st["phi_v"], zv, sv["cin"], 0, d, Nt, dsv, (dC, 0, d))
                if zx is not None:
                    self._mlp_bwd(theta, gtheta, st["phi_xx"], zx, sv["cin"], 0, d, Nt, dsx, (dC, 0, d))
            if not vel and dh > 0:
                dC_ld, dC_off = dh + d + G_, dh
                # (x, h) node function first: it yields dL/dm_i, which joins the message gradient below.
                # dC starts as [dL/dh' | 0]; the MLP input gradient is accumulated on top: [dL/dh (so far) | dL/dm_i]
                dph = torch.empty((Nt, dh), **f32)
                ops.copy_cols(dN, F, 3, dh, Nt, dph, dh, 0)
                dC = torch.zeros((Nt, dh + d + G_), **f32)
                ops.copy_cols(dN, F, 3, dh, Nt, dC, dh + d + G_, 0)
                self._mlp_bwd(theta, gtheta, st["phi_h"], sv["zh"], sv["Cc"], 0, dh + d + G_, Nt, dph,
                              (dC, 0, dh + d + G_))
            layers = st["phi_e"] + st["phi_x"]
            last = acts[-1]
            # t = gelu(U_last) @ k_last
            _mm(d, 1, Et, 1.0, last, 0, d, dt, 0, Et, gtheta, st["k_last"], 1, ta=True, tb=True, pro=1, pro_scale=1.0,
                tag="egnn_mlp_bwd", tf32x3=tc)
            dZ = torch.empty((Et, d), **f32)
            _mm(Et, d, 1, 1.0, dt, 0, Et, theta, st["k_last"], 1, dZ, 0, d, ta=True, tb=True, epi=2, aux=last, ld_aux=d,
                epi_scale=1.0, tag="egnn_mlp_bwd", tf32x3=tc)
            dU = None
            for l in range(len(layers) - 1, -1, -1):
                ko, bo, fan = layers[l]
                a_prev, pro = (acts[l - 1], 1) if l > 0 else (U, 0)
                gated = (l == nE and st["phi_a"] is not None)      # this layer's input is the soft-gated message
                if gated:
                    a_prev, pro = sv["Mg"], 0
                ops.dense_wgrad(fan, d, Et, a_prev, 0, fan, dZ, d, gtheta, ko, d, gtheta, bo, pro=pro, pro_scale=1.0,
                                tf32x3=tc, tag="egnn_mlp_bwd")            # kernel and bias gradients, one pass over dZ
                if gated:
                    ka, ba = st["phi_a"]
                    ZL, Za = acts[l - 1], sv["Za"]
                    dM = torch.empty((Et, d), **f32)
                    _mm(Et, d, d, 1.0, dZ, 0, d, theta, ko, d, dM, 0, d, tb=True, tag="egnn_mlp_bwd", tf32x3=tc)
                    if dC is not None:                              # + dL/dm_i / deg on the gated message
                        ops.egnn_segment_rows_bwd(dC, dC_ld, dC_off, d, None, 0, rowptr, Nt, agg, dM, argmax=sv["am"])
                    dZa = torch.empty((Et, d), **f32)
                    ops.softgate_bwd(ZL, Za, dM, dZa, dM)           # dM <- dM * sigmoid(Za)  (first branch into Z_L)
                    ops.dense_wgrad(d, d, Et, ZL, 0, d, dZa, d, gtheta, ka, d, gtheta, ba, pro=1, pro_scale=1.0, tf32x3=tc,
                                    tag="egnn_mlp_bwd")
                    # dZ_L = (dZa @ Ka^T + dM) * gelu'(Z_L): second branch accumulated in the epilogue (EPI 3)
                    _mm(Et, d, d, 1.0, dZa, 0, d, theta, ka, d, dM, 0, d, tb=True, epi=3, aux=ZL, ld_aux=d, epi_scale=1.0,
                        tag="egnn_mlp_bwd", tf32x3=tc)
                    dZ = dM
                    continue
                if l > 0:
                    dZp = torch.empty((Et, fan), **f32)
                    _mm(Et, fan, d, 1.0, dZ, 0, d, theta, ko, d, dZp, 0, fan, tb=True, epi=2, aux=acts[l - 1], ld_aux=fan,
                        epi_scale=1.0, tag="egnn_mlp_bwd", tf32x3=tc)
                    if dC is not None and l == nE:                  # message = gelu(Z_L): + dL/dm_i / deg * gelu'(Z_L)
                        ops.egnn_segment_rows_bwd(dC, dC_ld, dC_off, d, acts[l - 1], 1, rowptr, Nt, agg, dZp, argmax=sv["am"])
                    dZ = dZp
                elif si > 0:
                    dU = torch.empty((Et, fan), **f32)
                    _mm(Et, fan, d, 1.0, dZ, 0, d, theta, ko, d, dU, 0, fan, tb=True, tag="egnn_mlp_bwd", tf32x3=tc)
            mix = None
            if si > 0:
                dfeats = dU
                if K0 != nf:
                    dfeats = torch.empty((Et, nf), **f32)
                    ops.copy_cols(dU, K0, 0, nf, Et, dfeats, nf, 0)
                mix = torch.empty((Nt, 3), **f32)
                ops.egnn_coord_bwd(rij4, t, self.tanh_out, agg, rowptr, perm, Nt, g3, self.n_radial_basis, self.r_max,
                                   dfeats, None, dre, mix, argmax=sv["ax3"])
                ops.egnn_sender_acc(dre, rowptr_s, perm_s, Nt, mix)
            if not vel:
                if dh == 0 or si == 0:
                    dN = mix
                    continue
                dN_in = torch.empty((Nt, F), **f32)
                ops.copy_cols(mix, 3, 0, 3, Nt, dN_in, F, 0)
                ops.copy_cols(dC, dh + d + G_, 0, dh, Nt, dN_in, F, 3)
                dhs_e = torch.empty((Et, dh), **f32)
                ops.egnn_edge_input_bwd(dU, nf, dh, rowptr, perm, rowptr_s, perm_s, Nt, None, dhs_e, dN_in, F, 3, n_glob=G_)
                dN = dN_in
                continue
            # ---- node function backward (egnn.py:171-229)
            zv, zh, zx = sv["zv"], sv["zh"], sv["zx"]
            dN_in = torch.empty((Nt, F), **f32) if si > 0 else None
            dsv = torch.empty((Nt, 1), **f32)
            dsx = torch.empty((Nt, 1), **f32) if zx is not None else None
            ops.egnn_xv_bwd(nodes, F, None if zx is None else zx[-1], zv[-1], dN, F_out, mix, Nt, dN_in, dsx, dsv)
            if dh == 0:                                      # (x, v): the MLPs were differentiated above; 6 columns only
                dN = dN_in
                continue
            cin, c_off, c_ld = sv["cin"], sv["c_off"], sv["c_ld"]
            din, dCv = None, None
            if si > 0:
                dhs_e = torch.empty((Et, dh), **f32)
                ops.egnn_edge_input_bwd(dU, nf, dh, rowptr, perm, rowptr_s, perm_s, Nt, None, dhs_e, dN_in, F, 6, n_glob=G_)
                if G_:                                       # the MLPs read [h | globals]: gradient through a side buffer
                    dCv = torch.zeros((Nt, dh + G_), **f32)
                    din = (dCv, 0, dh + G_)
                else:
                    din = (dN_in, 6, F)
            dph = torch.empty((Nt, dh), **f32)
            ops.copy_cols(dN, F, 6, dh, Nt, dph, dh, 0)
            self._mlp_bwd(theta, gtheta, st["phi_h"], zh, cin, c_off, c_ld, Nt, dph, din)
            self._mlp_bwd(theta, gtheta, st["phi_v"], zv, cin, c_off, c_ld, Nt, dsv, din)
            if zx is not None:
                self._mlp_bwd(theta, gtheta, st["phi_xx"], zx, cin, c_off, c_ld, Nt, dsx, din)
            if dCv is not None:
                ops.copy_cols(dCv, dh + G_, 0, dh, Nt, dN_in, F, 6, accumulate=True)
            dN = dN_in
        return None


class _EGNNFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, theta, model, graphs):
        x, B, N, E, csr = model._prepare(graphs)
        out, saved = model._forward(theta.detach(), x, B, N, E, csr)
        ctx.stuff = (model, saved, B, N, E, csr, theta.detach())
        return out

    @staticmethod
    def backward(ctx, gout):
        model, saved, B, N, E, csr, theta = ctx.stuff
        g = torch.zeros_like(theta)
        model._backward(theta, g, saved, gout.contiguous(), B, N, E, csr)
        return g, None, None
