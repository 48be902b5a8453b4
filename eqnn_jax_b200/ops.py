"""Thin torch-tensor wrappers over the C ABI (include/eqnn_b200.h).

torch supplies device memory and the CUDA stream only; every computation below
is a hand-written sm_100a kernel reached through ctypes.  All tensors must be
CUDA tensors (fp32 / int32, contiguous unless strides are passed explicitly).
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from ._lib import check, lib

_F32, _I32 = torch.float32, torch.int32


class KernelTimer:
    """Optional CUDA-event timing of individual launches (bench.py's instrumented pass).
    Events are recorded on torch's current stream, the one every kernel is launched on."""

    def __init__(self):
        self.records = []   # (tag, start_event, end_event)

    def begin(self):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        return e

    def end(self, tag, start):
        e = torch.cuda.Event(enable_timing=True)
        e.record()
        self.records.append((tag, start, e))

    def summary(self):
        torch.cuda.synchronize()
        out = {}
        for tag, a, b in self.records:
            ms = a.elapsed_time(b)
            n, t = out.get(tag, (0, 0.0))
            out[tag] = (n + 1, t + ms)
        return out


TIMER: Optional[KernelTimer] = None


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def _req(t: torch.Tensor, dtype, name: str, contiguous: bool = True) -> torch.Tensor:
    if not isinstance(t, torch.Tensor) or not t.is_cuda:
        raise RuntimeError(f"{name}: expected a CUDA tensor (this path has no CPU implementation)")
    if t.dtype != dtype:
        raise RuntimeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if contiguous and not t.is_contiguous():
        raise RuntimeError(f"{name}: expected a contiguous tensor")
    return t


# ---------------------------------------------------------------- signatures
def tp_lookup(signature: str, spec=None) -> int:
    """Id of a compiled tensor-product signature.  spec = (x irreps, lmax, live irreps) lets an unknown signature be
    generated, compiled and loaded on first use (eqnn_jax_b200.build.add_tp_signature)."""
    i = lib().eqnn_tp_lookup(signature.encode())
    if i < 0 and spec is not None:
        from . import _lib as L, build
        build.add_tp_signature(*spec)
        L.reload()
        i = lib().eqnn_tp_lookup(signature.encode())
    if i < 0:
        have = [lib().eqnn_tp_signature(j).decode() for j in range(lib().eqnn_tp_signature_count())]
        raise RuntimeError(
            f"tensor-product signature {signature!r} is not compiled into libeqnn_b200.so; add it to "
            f"eqnn_jax_b200.codegen.PREBUILT and rebuild. Available: {have}")
    return i


def tp_dims(tp_id: int) -> Tuple[int, int, int, int]:
    v = [C.c_int() for _ in range(4)]
    check(lib().eqnn_tp_dims(tp_id, *[C.byref(x) for x in v]), "tp_dims")
    return tuple(x.value for x in v)


# ---------------------------------------------------------------- graph build
_KNN_WS = None


def knn_pbc(x: torch.Tensor, k: int, cell=None, cell_inv=None, fma: bool = False, want_dr: bool = True,
            dense: bool = False):
    """x [B, N, F>=3] -> senders [B,E], receivers [B,E] (int32), dr [B,E,3].
    ``dense=True`` forces the N x N sweep; otherwise large periodic clouds use the cell list (same output)."""
    global _KNN_WS
    _req(x, _F32, "x")
    B, N, F = x.shape
    E = N * k
    senders = torch.empty((B, E), dtype=_I32, device=x.device)
    receivers = torch.empty((B, E), dtype=_I32, device=x.device)
    dr = torch.empty((B, E, 3), dtype=_F32, device=x.device) if want_dr else None
    flags = (1 if cell is not None else 0) | (2 if fma else 0) | (4 if dense else 0)
    c9 = (C.c_float * 9)(*[float(v) for v in cell.reshape(-1)]) if cell is not None else None
    ci9 = (C.c_float * 9)(*[float(v) for v in cell_inv.reshape(-1)]) if cell is not None else None
    need = 0 if (dense or cell is None) else int(lib().eqnn_knn_workspace_bytes(B, N, k))
    ws = None
    if need:
        if _KNN_WS is None:
            _KNN_WS = GemmWorkspace()
        ws, need = _KNN_WS.get(need, x.device)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_knn_pbc(_ptr(x), F, c9, ci9, B, N, k, flags, _ptr(senders), _ptr(receivers), _ptr(dr),
                             ws, need, _stream()), "knn_pbc")
    if TIMER:
        TIMER.end("knn_pbc", t0)
    return senders, receivers, dr


def radius_graph(x: torch.Tensor, radius: float, cell=None, cell_inv=None):
    """x [B, N, F>=3] -> (senders, receivers [n_total] int32 local ids, dist [n_total], n_edge [B] int64 on the host):
    every pair with 0 < |apply_pbc(x_i - x_j)| < radius, in np.where order (graph, i, j).  The edge list is ragged,
    so its length is read back once (one host synchronisation) before the fill pass."""
    _req(x, _F32, "x")
    B, N, F = x.shape
    flags = 1 if cell is not None else 0
    c9 = (C.c_float * 9)(*[float(v) for v in cell.reshape(-1)]) if cell is not None else None
    ci9 = (C.c_float * 9)(*[float(v) for v in cell_inv.reshape(-1)]) if cell is not None else None
    count = torch.empty((B * N,), dtype=_I32, device=x.device)
    rowptr = torch.empty((B * N + 1,), dtype=_I32, device=x.device)
    check(lib().eqnn_radius_graph(_ptr(x), F, c9, ci9, B, N, float(radius), flags, _ptr(count), None, None, None, None,
                                  _stream()), "radius_graph (count)")
    check(lib().eqnn_exclusive_scan_i32(_ptr(count), B * N, _ptr(rowptr), _stream()), "exclusive_scan")
    offs = rowptr[::N].cpu().to(torch.int64)                       # B + 1 graph offsets (the host sync)
    total = int(offs[-1])
    senders = torch.empty((total,), dtype=_I32, device=x.device)
    receivers = torch.empty((total,), dtype=_I32, device=x.device)
    dist = torch.empty((total,), dtype=_F32, device=x.device)
    if total:
        check(lib().eqnn_radius_graph(_ptr(x), F, c9, ci9, B, N, float(radius), flags, None, _ptr(rowptr), _ptr(senders),
                                      _ptr(receivers), _ptr(dist), _stream()), "radius_graph (fill)")
    return senders, receivers, dist, offs[1:] - offs[:-1]


def edge_features(dr: torch.Tensor, n_rb: int, r_max: float) -> torch.Tensor:
    """dr [..., 3] -> |dr| [..., 1] (n_rb == 0) or bessel basis [..., n_rb]."""
    _req(dr, _F32, "dr")
    n = dr.numel() // 3
    out = torch.empty(dr.shape[:-1] + (max(n_rb, 1),), dtype=_F32, device=dr.device)
    check(lib().eqnn_edge_features(_ptr(dr), n, n_rb, float(r_max), _ptr(out), _stream()), "edge_features")
    return out


def csr_build(senders: torch.Tensor, receivers: torch.Tensor, n_nodes_per_graph: int):
    """[B,E] local ids -> rowptr [B*N+1], perm [B*E], src [B*E] (all int32, global ids)."""
    _req(senders, _I32, "senders")
    _req(receivers, _I32, "receivers")
    B, E = receivers.shape
    N = n_nodes_per_graph
    dev = receivers.device
    rowptr = torch.empty((B * N + 1,), dtype=_I32, device=dev)
    perm = torch.empty((B * E,), dtype=_I32, device=dev)
    src = torch.empty((B * E,), dtype=_I32, device=dev)
    nbytes = lib().eqnn_csr_workspace_bytes(B, N, E)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    check(lib().eqnn_csr_build(_ptr(senders), _ptr(receivers), B, N, E, _ptr(rowptr), _ptr(perm), _ptr(src),
                               _ptr(ws), nbytes, _stream()), "csr_build")
    return rowptr, perm, src


def edge_geom(pos: torch.Tensor, ld_pos: int, n_nodes: int, rowptr, src, n_rb: int, r_cut: float, want_radial: bool = True):
    """-> geom [Etot,4] (r_hat, d), radial [Etot, max(n_rb,1)] (None with want_radial=False)."""
    n_edges = src.numel()
    geom = torch.empty((n_edges, 4), dtype=_F32, device=src.device)
    radial = torch.empty((n_edges, max(n_rb, 1)), dtype=_F32, device=src.device) if want_radial else None
    check(lib().eqnn_edge_geom(_ptr(pos), ld_pos, n_nodes, _ptr(rowptr), _ptr(src), n_rb, float(r_cut),
                               _ptr(geom), _ptr(radial), _stream()), "edge_geom")
    return geom, radial


class EdgePairs:
    """Unique radial evaluations of a prepared graph (eqnn_edge_pairs): every CSR slot p reads row uidx[p]; row u is
    represented by slot rep_slot[u] (and its reverse edge rep_partner[u], -1 none, -2 the zero class = all self-loops of
    length 0, one slot per node in zero_slot, row zero_row); d_u / radial_u are the rows' lengths and Bessel bases; n_unique
    is a DEVICE int32 (the live row count), cap = the edge count sizes every per-row buffer."""
    __slots__ = ("uidx", "rep_slot", "rep_partner", "d_u", "zero_slot", "n_unique", "zero_row", "cap", "n_nodes", "radial_u")

    def __init__(self, uidx, rep_slot, rep_partner, d_u, zero_slot, n_unique, zero_row, cap, n_nodes, radial_u):
        self.uidx, self.rep_slot, self.rep_partner, self.d_u = uidx, rep_slot, rep_partner, d_u
        self.zero_slot, self.n_unique, self.zero_row = zero_slot, n_unique, zero_row
        self.cap, self.n_nodes, self.radial_u = cap, n_nodes, radial_u


def edge_pairs(rowptr, src, geom, n_nodes, n_rb: int, r_cut: float) -> EdgePairs:
    E = src.numel()
    dev = src.device
    uidx = torch.empty((E,), dtype=_I32, device=dev)
    rep_slot = torch.empty((E,), dtype=_I32, device=dev)
    rep_partner = torch.empty((E,), dtype=_I32, device=dev)
    d_u = torch.empty((E,), dtype=_F32, device=dev)
    zero_slot = torch.empty((max(n_nodes, 1),), dtype=_I32, device=dev)
    counts = torch.empty((2,), dtype=_I32, device=dev)
    n_unique, zero_row = counts[0:1], counts[1:2]
    nbytes = lib().eqnn_edge_pairs_workspace_bytes(n_nodes, E)
    ws = torch.empty((nbytes,), dtype=torch.uint8, device=dev)
    check(lib().eqnn_edge_pairs(_ptr(rowptr), _ptr(src), _ptr(geom), n_nodes, E, _ptr(uidx), _ptr(rep_slot),
                                _ptr(rep_partner), _ptr(d_u), _ptr(zero_slot), _ptr(n_unique), _ptr(zero_row), _ptr(ws), nbytes,
                                _stream()), "edge_pairs")
    radial_u = None
    if n_rb > 0:
        radial_u = torch.empty((E, n_rb), dtype=_F32, device=dev)
        check(lib().eqnn_bessel_rows(_ptr(d_u), E, _ptr(n_unique), n_rb, float(r_cut), _ptr(radial_u), _stream()),
              "bessel_rows")
    return EdgePairs(uidx, rep_slot, rep_partner, d_u, zero_slot, n_unique, zero_row, E, n_nodes, radial_u)


_PAIR_WS = None


def pair_combine(dw_rows, n_out, pairs: EdgePairs, dw_t):
    global _PAIR_WS
    if _PAIR_WS is None:
        _PAIR_WS = GemmWorkspace()
    ws, ws_bytes = _PAIR_WS.get(lib().eqnn_pair_combine_workspace_bytes(), dw_t.device)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_pair_combine(_ptr(dw_rows), dw_rows.shape[1], n_out, _ptr(pairs.rep_slot), _ptr(pairs.rep_partner),
                                  _ptr(pairs.zero_slot), _ptr(pairs.zero_row), pairs.n_nodes, pairs.cap, _ptr(pairs.n_unique),
                                  _ptr(dw_t), dw_t.shape[1], ws, ws_bytes, _stream()), "pair_combine")
    if TIMER:
        TIMER.end("pair_combine", t0)


# ---------------------------------------------------------------- interaction block
AGG = {"sum": 0, "mean": 1, "max": 2}


def tpconv_fwd(tp_id, rowptr, src, geom, w_t, xt, n_nodes, agg, out, ld_out, argmax=None):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_tpconv_fwd(tp_id, _ptr(rowptr), _ptr(src), _ptr(geom), _ptr(w_t), src.numel(), _ptr(xt),
                                n_nodes, agg, _ptr(out), ld_out, _ptr(argmax), _stream()), "tpconv_fwd")
    if TIMER:
        TIMER.end("tpconv_fwd", t0)


def tpconv_bwd(tp_id, rowptr, src, perm, geom, w_t, xt, n_nodes, agg, gA, ld_ga, dw_t, dxe, argmax=None):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_tpconv_bwd(tp_id, _ptr(rowptr), _ptr(src), _ptr(perm), _ptr(geom), _ptr(w_t), src.numel(),
                                _ptr(xt), n_nodes, agg, _ptr(gA), ld_ga, _ptr(dw_t), _ptr(dxe), _ptr(argmax), _stream()),
          "tpconv_bwd")
    if TIMER:
        TIMER.end("tpconv_bwd", t0)


def tpconv_fwd_shared(tp_id, rowptr, src, geom, pairs: EdgePairs, w_rows, xt, n_nodes, agg, out, ld_out, argmax=None):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_tpconv_fwd_shared(tp_id, _ptr(rowptr), _ptr(src), _ptr(geom), _ptr(pairs.uidx), _ptr(w_rows),
                                       w_rows.shape[1], _ptr(xt), n_nodes, agg, _ptr(out), ld_out, _ptr(argmax), _stream()),
          "tpconv_fwd_shared")
    if TIMER:
        TIMER.end("tpconv_fwd", t0)


def tpconv_bwd_shared(tp_id, rowptr, src, perm, geom, pairs: EdgePairs, w_rows, xt, n_nodes, agg, gA, ld_ga, dw_rows, dxe,
                      argmax=None):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_tpconv_bwd_shared(tp_id, _ptr(rowptr), _ptr(src), _ptr(perm), _ptr(geom), _ptr(pairs.uidx),
                                       _ptr(w_rows), w_rows.shape[1], _ptr(xt), n_nodes, agg, _ptr(gA), ld_ga,
                                       _ptr(dw_rows), _ptr(dxe), _ptr(argmax), _stream()), "tpconv_bwd_shared")
    if TIMER:
        TIMER.end("tpconv_bwd", t0)


def tp_messages(tp_id, src, perm, geom, w_rows, cols, scales, xt, out):
    """Un-reduced per-edge messages in the reference's edge order (node-task `edges`, models/nequip.py:154-171)."""
    fa = (C.c_float * len(scales))(*[float(v) for v in scales])
    check(lib().eqnn_tp_messages(tp_id, _ptr(src), _ptr(perm), _ptr(geom), _ptr(w_rows), w_rows.shape[1], _int_arr(cols),
                                 fa, _ptr(xt), src.numel(), _ptr(out), out.shape[1], _stream()), "tp_messages")


def sender_reduce(dxe, n_nodes, k, dxp, dxt):
    check(lib().eqnn_sender_reduce(_ptr(dxe), n_nodes, k, dxp, _ptr(dxt), _stream()), "sender_reduce")


def segment_gather_sum(dxe, rowptr_s, perm_s, n_nodes, dxp, dxt):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segment_gather_sum(_ptr(dxe), _ptr(rowptr_s), _ptr(perm_s), n_nodes, dxp, _ptr(dxt), _stream()),
          "segment_gather_sum")
    if TIMER:
        TIMER.end("sender_reduce", t0)


def sh_scatter(lmax, rowptr, geom, n_nodes, scale, attrs):
    check(lib().eqnn_sh_scatter(lmax, _ptr(rowptr), _ptr(geom), n_nodes, float(scale), _ptr(attrs), _stream()),
          "sh_scatter")


def node_tp_fwd(tp_id, x, y, n, T):
    check(lib().eqnn_node_tp_fwd(tp_id, _ptr(x), _ptr(y), n, _ptr(T), _stream()), "node_tp_fwd")


def node_tp_bwd(tp_id, x, y, gT, n, dx):
    check(lib().eqnn_node_tp_bwd(tp_id, _ptr(x), _ptr(y), _ptr(gT), n, _ptr(dx), _stream()), "node_tp_bwd")


# ---------------------------------------------------------------- dense
class GemmWorkspace:
    """Grow-only split-K scratch buffer (torch owns the memory)."""

    def __init__(self):
        self.buf: Optional[torch.Tensor] = None

    def get(self, nbytes: int, device) -> Tuple[C.c_void_p, int]:
        if nbytes == 0:
            return C.c_void_p(0), 0
        if self.buf is None or self.buf.numel() < nbytes or self.buf.device != device:
            self.buf = torch.empty((nbytes,), dtype=torch.uint8, device=device)
        return C.c_void_p(self.buf.data_ptr()), self.buf.numel()


_GEMM_WS = GemmWorkspace()


def gemm(M, N, K, alpha, A, sam, sak, B, sbk, sbn, Cm, scm, scn, accumulate=False, pro=0, pro_scale=1.0, epi=0,
         aux=None, ld_aux=0, epi_scale=1.0, a_off=0, b_off=0, c_off=0, aux_off=0, tag="gemm", tf32x3=False, tc_any=False):
    """C[m,n] (+)= epi(alpha * sum_k pro(A[m,k]) B[k,n]); element strides; *_off = element offsets."""
    L = lib()
    need = L.eqnn_gemm_workspace_bytes(M, N, K)
    ws, ws_bytes = _GEMM_WS.get(need, Cm.device)
    pa = C.c_void_p(A.data_ptr() + 4 * a_off)
    pb = C.c_void_p(B.data_ptr() + 4 * b_off)
    pc = C.c_void_p(Cm.data_ptr() + 4 * c_off)
    px = C.c_void_p(0 if aux is None else aux.data_ptr() + 4 * aux_off)
    t0 = TIMER.begin() if TIMER else None
    check(L.eqnn_gemm_f32(M, N, K, float(alpha), pa, sam, sak, pb, sbk, sbn, pc, scm, scn,
                          (1 if accumulate else 0) | (2 if (tf32x3 or tc_any) else 0) | (4 if tc_any else 0), pro, float(pro_scale), epi, px, ld_aux, float(epi_scale), ws, ws_bytes, _stream()), "gemm")
    if TIMER:
        TIMER.end(tag, t0)


def radial_mlp_supported(n_rb: int, d_hidden: int, n_hidden: int, n_out: int) -> bool:
    return bool(lib().eqnn_radial_mlp_supported(n_rb, d_hidden, n_hidden, n_out))


def _weight_ptrs(weights):
    """[(tensor, element offset, leading dimension)] -> host arrays of device pointers / leading dimensions."""
    n = len(weights)
    ptrs = (C.c_void_p * n)(*[t.data_ptr() + 4 * off for t, off, _ in weights])
    lds = (C.c_int64 * n)(*[ld for _, _, ld in weights])
    return ptrs, lds


def radial_mlp_stash(n_edges, n_hidden, device):
    """Buffer for the pre-activations the forward kernel keeps for the backward pass."""
    return torch.empty((lib().eqnn_radial_mlp_stash_bytes(n_edges, n_hidden) // 4,), dtype=_F32, device=device)


ACT = {"gelu": 0, "silu": 1}


def radial_mlp_fwd(geom, n_rb, r_cut, weights, n_out, act_scale, w_t, z_stash=None, activation="gelu"):
    """geom [E, 4] (r_hat, d) -> w_t [n_out, E]; weights = [(tensor, offset, ld)] for the hidden layers + the output layer."""
    E = geom.shape[0]
    ptrs, lds = _weight_ptrs(weights)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_radial_mlp_fwd(C.c_void_p(geom.data_ptr() + 12), 4, E, n_rb, float(r_cut), len(weights) - 1, 128,
                                    ptrs, lds, n_out, ACT[activation], float(act_scale), _ptr(w_t), w_t.shape[1],
                                    _ptr(z_stash), _stream()), "radial_mlp_fwd")
    if TIMER:
        TIMER.end("radial_mlp_fwd", t0)


def radial_mlp_fwd_rows(pairs: EdgePairs, n_rb, r_cut, weights, n_out, act_scale, w_rows, z_stash=None, activation="gelu"):
    """The radial MLP once per unique edge length: pairs.d_u [cap] -> w_rows [cap, n_out rounded up to 4] (row-major)."""
    ptrs, lds = _weight_ptrs(weights)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_radial_mlp_fwd_rows(_ptr(pairs.d_u), 1, pairs.cap, _ptr(pairs.n_unique), n_rb, float(r_cut),
                                         len(weights) - 1, 128, ptrs, lds, n_out, ACT[activation], float(act_scale), None, 0,
                                         _ptr(w_rows), w_rows.shape[1], _ptr(z_stash), _stream()), "radial_mlp_fwd_rows")
    if TIMER:
        TIMER.end("radial_mlp_fwd", t0)


_RMLP_WS = GemmWorkspace()


def radial_mlp_bwd(geom, n_rb, r_cut, weights, n_out, act_scale, dw_t, grad_weights, z_stash=None, radial=None,
                   activation="gelu"):
    """Weight gradients of the radial MLP from dw_t [n_out, E]; grad_weights = [(tensor, offset, ld)] like weights.
    With z_stash (from radial_mlp_fwd) and radial (edge_geom) the activations are not recomputed."""
    E = geom.shape[0]
    ptrs, lds = _weight_ptrs(weights)
    gptrs, glds = _weight_ptrs(grad_weights)
    assert list(lds) == list(glds)
    need = lib().eqnn_radial_mlp_bwd_workspace_bytes(E, n_rb, len(weights) - 1)
    ws, ws_bytes = _RMLP_WS.get(need, geom.device)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_radial_mlp_bwd(C.c_void_p(geom.data_ptr() + 12), 4, E, n_rb, float(r_cut), len(weights) - 1, 128,
                                    ptrs, lds, n_out, ACT[activation], float(act_scale), _ptr(dw_t), dw_t.shape[1],
                                    _ptr(z_stash), _ptr(radial), gptrs, ws, ws_bytes, _stream()), "radial_mlp_bwd")
    if TIMER:
        TIMER.end("radial_mlp_bwd", t0)


def radial_mlp_bwd_rows(pairs: EdgePairs, n_rb, r_cut, weights, n_out, act_scale, dw_t, grad_weights, z_stash,
                        activation="gelu"):
    """Weight gradients from the per-row cotangents dw_t [n_out, cap] (eqnn_pair_combine) and the forward stash."""
    ptrs, lds = _weight_ptrs(weights)
    gptrs, glds = _weight_ptrs(grad_weights)
    assert list(lds) == list(glds)
    need = lib().eqnn_radial_mlp_bwd_workspace_bytes(pairs.cap, n_rb, len(weights) - 1)
    ws, ws_bytes = _RMLP_WS.get(need, dw_t.device)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_radial_mlp_bwd_rows(_ptr(pairs.d_u), 1, pairs.cap, _ptr(pairs.n_unique), n_rb, float(r_cut),
                                         len(weights) - 1, 128, ptrs, lds, n_out, ACT[activation], float(act_scale),
                                         _ptr(dw_t), dw_t.shape[1], _ptr(z_stash), _ptr(pairs.radial_u), gptrs, ws, ws_bytes,
                                         _stream()), "radial_mlp_bwd_rows")
    if TIMER:
        TIMER.end("radial_mlp_bwd", t0)


def _int_arr(v: Sequence[int]):
    return (C.c_int * max(len(v), 1))(*v)


ACT = {"gelu": 0, "silu": 1}     # EQNN_ACT_GELU / EQNN_ACT_SILU


def gate_fwd(z, n, n_extra, n_gates, mul, l, inv_c_act, inv_c_gate, out, activation="gelu"):
    check(lib().eqnn_gate_act_fwd(_ptr(z), n, n_extra, n_gates, len(mul), _int_arr(mul), _int_arr(l), ACT[activation],
                                  inv_c_act, inv_c_gate, _ptr(out), _stream()), "gate_fwd")


def gate_bwd(z, gout, n, n_extra, n_gates, mul, l, inv_c_act, inv_c_gate, gz, activation="gelu"):
    check(lib().eqnn_gate_act_bwd(_ptr(z), _ptr(gout), n, n_extra, n_gates, len(mul), _int_arr(mul), _int_arr(l),
                                  ACT[activation], inv_c_act, inv_c_gate, _ptr(gz), _stream()), "gate_bwd")


_NODE_WS = GemmWorkspace()


def _off(t, off):
    return C.c_void_p(0) if off is None else C.c_void_p(t.data_ptr() + 4 * off)


def node_update_fwd(A, ld_a, d_a, h, d_h, wexp, w1_off, alpha1, w2_off, n_extra, n_gates, mul, l, inv_c_act,
                    inv_c_gate, w3_off, n, z, h_out, w0n_off=None, dxp=0, xt_out=None):
    """Fused z = alpha1*A@W1 + h@W2; h' = gate(z)@W3 (weights are slices of the expanded-weight buffer)."""
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_node_update_fwd(_ptr(A), ld_a, d_a, _ptr(h), d_h, _off(wexp, w1_off), float(alpha1),
                                     _off(wexp, w2_off), n_extra, n_gates, len(mul), _int_arr(mul), _int_arr(l),
                                     inv_c_act, inv_c_gate, _off(wexp, w3_off), _off(wexp, w0n_off), dxp, n, _ptr(z),
                                     _ptr(h_out), _ptr(xt_out), _stream()), "node_update_fwd")
    if TIMER:
        TIMER.end("node_update_fwd", t0)


def node_update_bwd(A, ld_a, d_a, h, d_h, wexp, w1_off, alpha1, w2_off, n_extra, n_gates, mul, l, inv_c_act,
                    inv_c_gate, w3_off, z, dh_out, n, dA, dh_in, gwexp):
    d_gated = sum(m * (2 * ll + 1) for m, ll in zip(mul, l))
    need = lib().eqnn_node_update_workspace_bytes(d_a, d_h, n_extra + n_gates + d_gated, n_extra + d_gated)
    ws, ws_bytes = _NODE_WS.get(need, z.device)
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_node_update_bwd(_ptr(A), ld_a, d_a, _ptr(h), d_h, _off(wexp, w1_off), float(alpha1),
                                     _off(wexp, w2_off), n_extra, n_gates, len(mul), _int_arr(mul), _int_arr(l),
                                     inv_c_act, inv_c_gate, _off(wexp, w3_off), _ptr(z), _ptr(dh_out), n, _ptr(dA),
                                     _ptr(dh_in), _off(gwexp, w1_off), _off(gwexp, w2_off), _off(gwexp, w3_off), ws,
                                     ws_bytes, _stream()), "node_update_bwd")
    if TIMER:
        TIMER.end("node_update_bwd", t0)


def _c9(a):
    return None if a is None else (C.c_float * 9)(*[float(v) for v in a.reshape(-1)])


def egnn_geom_fwd(pos, ld_pos, n_nodes, rowptr, src, cell, cell_inv, n_rb, r_max, rij4, feats):
    check(lib().eqnn_egnn_geom_fwd(_ptr(pos), ld_pos, n_nodes, _ptr(rowptr), _ptr(src), _c9(cell), _c9(cell_inv), n_rb,
                                   float(r_max), _ptr(rij4), _ptr(feats), _stream()), "egnn_geom_fwd")


def egnn_coord_fwd(rij4, t, use_tanh, agg, rowptr, n_nodes, pos, ld_pos, cell, cell_inv, pos_out, ld_out, argmax=None):
    check(lib().eqnn_egnn_coord_fwd(_ptr(rij4), _ptr(t), int(use_tanh), agg, _ptr(rowptr), n_nodes, _ptr(pos), ld_pos,
                                    _c9(cell), _c9(cell_inv), _ptr(pos_out), ld_out, _ptr(argmax), _stream()),
          "egnn_coord_fwd")


def egnn_coord_bwd(rij4, t, use_tanh, agg, rowptr, perm, n_nodes, dpos_out, n_rb, r_max, dfeats, dt, dre, dpos_in,
                   argmax=None):
    check(lib().eqnn_egnn_coord_bwd(_ptr(rij4), _ptr(t), int(use_tanh), agg, _ptr(rowptr), _ptr(perm), n_nodes,
                                    _ptr(dpos_out), n_rb, float(r_max), _ptr(dfeats), _ptr(dt), _ptr(dre),
                                    _ptr(dpos_in), _ptr(argmax), _stream()), "egnn_coord_bwd")


def egnn_sender_acc(dre, rowptr_s, perm_s, n_nodes, dpos):
    check(lib().eqnn_egnn_sender_acc(_ptr(dre), _ptr(rowptr_s), _ptr(perm_s), n_nodes, _ptr(dpos), _stream()),
          "egnn_sender_acc")


def segnn_fwd_fused_ok(R, Dx, Dy, Do):
    """Shapes eqnn_segnn_tplg_fwd accepts (mirrors its argument checks)."""
    N = Dy * Do
    return Dy in (1, 2, 4, 8) and R >= 32 and N >= 32 and Dx >= 16 and R * N >= (1 << 14) and R * N * Dx >= (1 << 24) \
        and R <= 128 * 65535


def segnn_tplg_fwd(x, ld_x, Dx, mcat, m_off, y, Dy, Do, bias_buf, bias_off, bias_col, n_bias, R, z, x_off=0, y_off=0, z_off=0,
                   tag="segnn_gemm"):
    """z[r, o] = bias + sum_j y[r, j] (x @ M_cat)[r, o*Dy + j], contraction fused into the tensor-core GEMM epilogue."""
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segnn_tplg_fwd(_off(x, x_off), ld_x, Dx, _off(mcat, m_off), _off(y, y_off), Dy, Dy, Do,
                                    None if bias_buf is None else _off(bias_buf, bias_off), bias_col, n_bias, R,
                                    _off(z, z_off), Do, _stream()), "segnn_tplg_fwd")
    if TIMER:
        TIMER.end(tag, t0)


def segnn_ycontract(Xbig, y, ld_y, bias_buf, bias_off, bias_col, n_bias, R, Dy, Do, z, x_off=0, y_off=0, z_off=0):
    check(lib().eqnn_segnn_ycontract(_off(Xbig, x_off), None if y is None else _off(y, y_off), ld_y,
                                     None if bias_buf is None else _off(bias_buf, bias_off), bias_col, n_bias, R, Dy, Do,
                                     _off(z, z_off), _stream()), "segnn_ycontract")


def segnn_yexpand(dz, y, ld_y, R, Dy, Do, dXbig, dz_off=0, y_off=0):
    check(lib().eqnn_segnn_yexpand(_off(dz, dz_off), None if y is None else _off(y, y_off), ld_y, R, Dy, Do, _ptr(dXbig),
                                   _stream()), "segnn_yexpand")


def segnn_cg_apply(inp, ld_in, n_in, y, ld_y, Dy, ptr, idx, coef, bias_buf, bias_off, bias_col, n_bias, R, n_out, out,
                   ld_out, in_off=0, y_off=0, out_off=0, ptr_off=0):
    """out[r, o] = bias + sum_t coef[t] y[r, idx[t] >> 24] in[r, idx[t] & 0xFFFFFF] over ptr[o] .. ptr[o+1]."""
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segnn_cg_apply(_off(inp, in_off), ld_in, n_in, None if y is None else _off(y, y_off), ld_y, Dy,
                                    _off(ptr, ptr_off), _ptr(idx), _ptr(coef),
                                    None if bias_buf is None else _off(bias_buf, bias_off), bias_col, n_bias, R, n_out,
                                    _off(out, out_off), ld_out, _stream()), "segnn_cg_apply")
    if TIMER:
        TIMER.end("segnn_cg", t0)


def segnn_edge0_fwd(Ps, Pr, add, n_add, wexp, ma_off, y, Dy, Do, bias_buf, bias_off, bias_col, n_bias, rowptr, src, n_nodes, z):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segnn_edge0_fwd(_ptr(Ps), _ptr(Pr), _ptr(add), n_add, _off(wexp, ma_off), _ptr(y), Dy, Do,
                                     None if bias_buf is None else _off(bias_buf, bias_off), bias_col, n_bias,
                                     _ptr(rowptr), _ptr(src), n_nodes, _ptr(z), _stream()), "segnn_edge0_fwd")
    if TIMER:
        TIMER.end("segnn_edge0", t0)


def segnn_edge0_bwd(dz, y, Dy, Do, rowptr, rowptr_s, spos, n_nodes, dPs, dPr):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segnn_edge0_bwd(_ptr(dz), _ptr(y), Dy, Do, _ptr(rowptr), _ptr(rowptr_s), _ptr(spos), n_nodes,
                                     _ptr(dPs), _ptr(dPr), _stream()), "segnn_edge0_bwd")
    if TIMER:
        TIMER.end("segnn_edge0", t0)


def segnn_edge0_wadd(add, n_add, y, Dy, dz, Do, E, gwexp, g_off):
    L = lib()
    ws, ws_bytes = _GEMM_WS.get(L.eqnn_segnn_edge0_wadd_workspace_bytes(n_add, Dy, Do), dz.device)
    t0 = TIMER.begin() if TIMER else None
    check(L.eqnn_segnn_edge0_wadd(_ptr(add), n_add, _ptr(y), Dy, _ptr(dz), Do, E, _off(gwexp, g_off), ws, ws_bytes,
                                  _stream()), "segnn_edge0_wadd")
    if TIMER:
        TIMER.end("segnn_edge0", t0)


def segnn_yagg_fwd(m, Dx, y, Dy, rowptr, n_nodes, agg, S):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segnn_yagg_fwd(_ptr(m), Dx, _ptr(y), Dy, _ptr(rowptr), n_nodes, agg, _ptr(S), _stream()), "segnn_yagg_fwd")
    if TIMER:
        TIMER.end("segnn_edgeL", t0)


def segnn_yagg_bwd(dS, Dx, y, Dy, rowptr, n_nodes, agg, dm):
    t0 = TIMER.begin() if TIMER else None
    check(lib().eqnn_segnn_yagg_bwd(_ptr(dS), Dx, _ptr(y), Dy, _ptr(rowptr), n_nodes, agg, _ptr(dm), _stream()), "segnn_yagg_bwd")
    if TIMER:
        TIMER.end("segnn_edgeL", t0)


def segnn_bias_deg_fwd(out, out_off, ld, bias_buf, bias_off, n_bias, rowptr, n_nodes, agg):
    check(lib().eqnn_segnn_bias_deg_fwd(_off(out, out_off), ld, _off(bias_buf, bias_off), n_bias, _ptr(rowptr), n_nodes, agg,
                                        _stream()), "segnn_bias_deg_fwd")


def segnn_bias_deg_bwd(d, d_off, ld, rowptr, n_nodes, agg, n_bias, gb, gb_off):
    check(lib().eqnn_segnn_bias_deg_bwd(_off(d, d_off), ld, _ptr(rowptr), n_nodes, agg, n_bias, _off(gb, gb_off), _stream()),
          "segnn_bias_deg_bwd")


def perm_compose(perm, perm_s):
    """spos[q] = receiver-sorted position of the q-th sender-sorted edge."""
    n = perm.numel()
    inv = torch.empty_like(perm)
    spos = torch.empty_like(perm)
    check(lib().eqnn_perm_compose(_ptr(perm), _ptr(perm_s), n, _ptr(inv), _ptr(spos), _stream()), "perm_compose")
    return spos


def gather_rows(src, perm, width):
    n = perm.numel()
    dst = torch.empty((n, width), dtype=_F32, device=src.device)
    check(lib().eqnn_gather_rows(_ptr(src), _ptr(perm), n, width, _ptr(dst), _stream()), "gather_rows")
    return dst


def axpy(a, x, y):
    check(lib().eqnn_axpy(x.numel(), float(a), _ptr(x), _ptr(y), _stream()), "axpy")


def steerable_attrs(pos, ld_pos, vel, ld_vel, B, N, E, senders, receivers, rowptr, perm, cell, cell_inv, lmax, sh_norm):
    """-> r_ij [B*E, 3], edge_attrs [B*E, (lmax+1)^2], node_attrs [B*N, (lmax+1)^2] (equivariant_graph_utils.py:27-83)."""
    ny = (lmax + 1) ** 2
    dev = pos.device
    r_ij = torch.empty((B * E, 3), dtype=_F32, device=dev)
    e_attr = torch.empty((B * E, ny), dtype=_F32, device=dev)
    n_attr = torch.empty((B * N, ny), dtype=_F32, device=dev)
    check(lib().eqnn_steerable_attrs(_ptr(pos), ld_pos, _ptr(vel), ld_vel, B, N, E, _ptr(senders), _ptr(receivers),
                                     _ptr(rowptr), _ptr(perm), _c9(cell), _c9(cell_inv), lmax, sh_norm, _ptr(r_ij),
                                     _ptr(e_attr), _ptr(n_attr), _stream()), "steerable_attrs")
    return r_ij, e_attr, n_attr


def egnn_edge_input_fwd(feats, nf, nodes, F, h_off, dh, rowptr, src, n_nodes, U, glob=None, nodes_per_graph=1):
    n_glob = 0 if glob is None else glob.shape[1]
    check(lib().eqnn_egnn_edge_input_fwd(_ptr(feats), nf, _off(nodes, h_off) if dh else None, F, dh, _ptr(glob), n_glob,
                                         nodes_per_graph, _ptr(rowptr), _ptr(src), n_nodes, _ptr(U), _stream()),
          "egnn_edge_input_fwd")


def egnn_edge_input_bwd(dU, nf, dh, rowptr, perm, rowptr_s, perm_s, n_nodes, dfeats, dhs_e, dnodes, F, h_off, n_glob=0):
    check(lib().eqnn_egnn_edge_input_bwd(_ptr(dU), nf, dh, n_glob, _ptr(rowptr), _ptr(perm), _ptr(rowptr_s), _ptr(perm_s), n_nodes,
                                         _ptr(dfeats), _ptr(dhs_e), _off(dnodes, h_off), F, _stream()),
          "egnn_edge_input_bwd")


def egnn_xv_base(nodes, F, sx, sv, n, base):
    check(lib().eqnn_egnn_xv_base(_ptr(nodes), F, _ptr(sx), _ptr(sv), n, _ptr(base), _stream()), "egnn_xv_base")


def egnn_node_assemble(nodes, F, h_off, ph, n, nodes_next):
    check(lib().eqnn_egnn_node_assemble(_ptr(nodes), F, h_off, _ptr(ph), n, _ptr(nodes_next), _stream()),
          "egnn_node_assemble")


def egnn_segment_rows_fwd(rows, d, act, rowptr, n_nodes, agg, out, ld_out, col_off=0, argmax=None):
    """agg 2 = jraph.segment_max: argmax [n_nodes, d] int32 receives the winning CSR positions."""
    check(lib().eqnn_egnn_segment_rows_fwd(_ptr(rows), d, int(act), _ptr(rowptr), n_nodes, agg, _off(out, col_off), ld_out,
                                           _ptr(argmax), _stream()), "egnn_segment_rows_fwd")


def egnn_segment_rows_bwd(dmi, ld_dmi, col_off, d, z, act, rowptr, n_nodes, agg, target, argmax=None):
    check(lib().eqnn_egnn_segment_rows_bwd(_off(dmi, col_off), ld_dmi, d, _ptr(z), int(act), _ptr(rowptr), n_nodes, agg,
                                           _ptr(target), _ptr(argmax), _stream()), "egnn_segment_rows_bwd")


def egnn_xv_bwd(nodes, F, sx, sv, dnext, ld_dnext, mix, n, din, dsx, dsv):
    check(lib().eqnn_egnn_xv_bwd(_ptr(nodes), F, _ptr(sx), _ptr(sv), _ptr(dnext), ld_dnext, _ptr(mix), n, _ptr(din),
                                 _ptr(dsx), _ptr(dsv), _stream()), "egnn_xv_bwd")


def copy_cols(src, ld_src, col_src, width, n, dst, ld_dst, col_dst, group=1, accumulate=False):
    """dst[r, col_dst:col_dst+width] (+)= src[r // group, col_src:col_src+width] for r < n."""
    check(lib().eqnn_copy_cols(_ptr(src), ld_src, col_src, width, n, group, int(accumulate), _ptr(dst), ld_dst, col_dst,
                               _stream()), "copy_cols")


def softgate_fwd(z, za, m):
    check(lib().eqnn_softgate_fwd(_ptr(z), _ptr(za), z.numel(), _ptr(m), _stream()), "softgate_fwd")


def softgate_bwd(z, za, dm, dza, g1):
    check(lib().eqnn_softgate_bwd(_ptr(z), _ptr(za), _ptr(dm), z.numel(), _ptr(dza), _ptr(g1), _stream()), "softgate_bwd")


_COLSUM_WS = GemmWorkspace()


def dense_wgrad(M, N, K, A, a_off, lda, G, ldg, gW, w_off, ldw, gb, b_off, pro=0, pro_scale=1.0, tf32x3=False,
                tag="dense_wgrad"):
    """gW[M x N] = pro(A)^T G and gb[n] = sum_k G[k, n] in one pass over G (A: [K x M], G: [K x N], row-major)."""
    L = lib()
    ws, ws_bytes = _GEMM_WS.get(L.eqnn_dense_wgrad_workspace_bytes(M, N, K), gW.device)
    t0 = TIMER.begin() if TIMER else None
    check(L.eqnn_dense_wgrad_f32(M, N, K, 1.0, _off(A, a_off), lda, pro, float(pro_scale), _ptr(G), ldg, _off(gW, w_off),
                                 ldw, _off(gb, b_off), 2 if tf32x3 else 0, ws, ws_bytes, _stream()), "dense_wgrad")
    if TIMER:
        TIMER.end(tag, t0)


def colsum_tall(x, M, N, y, y_off=0):
    need = lib().eqnn_colsum_workspace_bytes(N)
    ws, ws_bytes = _COLSUM_WS.get(need, x.device)
    check(lib().eqnn_colsum_tall(_ptr(x), M, N, C.c_void_p(y.data_ptr() + 4 * y_off), ws, ws_bytes, _stream()),
          "colsum_tall")


def pool_fwd(x, B, N, D, mode, out):
    check(lib().eqnn_pool_fwd(_ptr(x), B, N, D, mode, _ptr(out), _stream()), "pool_fwd")


def pool_bwd(gout, B, N, D, mode, gx):
    check(lib().eqnn_pool_bwd(_ptr(gout), B, N, D, mode, _ptr(gx), _stream()), "pool_bwd")


def pool_max_fwd(x, B, N, D, out, argmax):
    check(lib().eqnn_pool_max_fwd(_ptr(x), B, N, D, _ptr(out), _ptr(argmax), _stream()), "pool_max_fwd")


def pool_max_bwd(gout, argmax, B, N, D, gx):
    check(lib().eqnn_pool_max_bwd(_ptr(gout), _ptr(argmax), B, N, D, _ptr(gx), _stream()), "pool_max_bwd")


def layernorm_fwd(a, g, scale_buf, scale_off, bias_off, eps, y, stats):
    """y = LayerNorm(concat(a, g)) (nequip.py:201-205); scale/bias live in the flat parameter buffer."""
    B, da = a.shape
    dg = 0 if g is None else g.shape[1]
    check(lib().eqnn_layernorm_fwd(_ptr(a), da, _ptr(g) if g is not None else None, dg, B,
                                   C.c_void_p(scale_buf.data_ptr() + 4 * scale_off),
                                   C.c_void_p(scale_buf.data_ptr() + 4 * bias_off), eps, _ptr(y), _ptr(stats), _stream()),
          "layernorm_fwd")


def layernorm_bwd(a, g, scale_buf, scale_off, stats, dy, d_a, grad_buf, dscale_off, dbias_off):
    B, da = a.shape
    dg = 0 if g is None else g.shape[1]
    check(lib().eqnn_layernorm_bwd(_ptr(a), da, _ptr(g) if g is not None else None, dg, B,
                                   C.c_void_p(scale_buf.data_ptr() + 4 * scale_off), _ptr(stats), _ptr(dy), _ptr(d_a),
                                   C.c_void_p(grad_buf.data_ptr() + 4 * dscale_off),
                                   C.c_void_p(grad_buf.data_ptr() + 4 * dbias_off), _stream()), "layernorm_bwd")


def colsum(x, M, N, y, x_off=0, y_off=0):
    check(lib().eqnn_colsum(C.c_void_p(x.data_ptr() + 4 * x_off), M, N, C.c_void_p(y.data_ptr() + 4 * y_off),
                            _stream()), "colsum")


def weights_expand(params, idx, scale, dst):
    check(lib().eqnn_weights_expand(_ptr(params), _ptr(idx), _ptr(scale), dst.numel(), _ptr(dst), _stream()),
          "weights_expand")


def weights_contract(ddst, pidx, ptr, tidx, scale_t, grads):
    check(lib().eqnn_weights_contract(_ptr(ddst), _ptr(pidx), _ptr(ptr), _ptr(tidx), _ptr(scale_t), pidx.numel(),
                                      _ptr(grads), _stream()), "weights_contract")


def mse_loss(pred, target, gscale, loss, gpred):
    B, O = pred.shape
    check(lib().eqnn_mse_loss(_ptr(pred), _ptr(target), B, O, float(gscale), _ptr(loss), _ptr(gpred), _stream()),
          "mse_loss")


def adamw(params, grads, m, v, lr, b1, b2, eps, wd, step):
    check(lib().eqnn_adamw(_ptr(params), _ptr(grads), _ptr(m), _ptr(v), params.numel(), lr, b1, b2, eps, wd, step,
                           _stream()), "adamw")
