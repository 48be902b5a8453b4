"""jax.ffi registration + custom_vjp rules over libeqnn_xla.so (xla_ffi_shim.cc): what a maintainer of the reference
imports in models/nequip.py and models/utils/graph_utils.py to route the hot path to the sm_100a kernels.

NOT importable in this image (no jax): the module is exercised only where jax / jaxlib exist.  The torch binding
(eqnn_jax_b200.ops / nequip) is the one the tests run; both call the same C symbols with the same arguments.

    from eqnn_jax_b200.jax_binding import ffi as b200
    senders, receivers, dr = b200.nearest_neighbors(x, k, cell)            # graph_utils.py:40-72, batched
    rowptr, perm, src = b200.csr_build(senders, receivers, n_nodes)
    geom, radial = b200.edge_geom(pos, rowptr, src, n_rb, r_cut)
    w_t = b200.radial_mlp(geom, radial, kernels, n_rb, r_cut, act_scale)     # nequip.py:55-68, differentiable in `kernels`
    A = b200.tpconv(xt, w_t, rowptr, src, perm, geom, tp_id, agg, d_live)    # nequip.py:46-52,68,118-120
"""
import ctypes
import functools
import os

import jax
import jax.numpy as jnp
import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_core = ctypes.CDLL(os.path.join(_HERE, "..", "libeqnn_b200.so"), mode=ctypes.RTLD_GLOBAL)
_shim = ctypes.CDLL(os.path.join(_HERE, "libeqnn_xla.so"))
for _name in ("knn_pbc", "csr_build", "edge_geom", "radial_mlp_fwd", "radial_mlp_bwd", "tpconv_fwd", "tpconv_bwd",
              "node_update_fwd", "node_update_bwd", "edge_pairs", "radial_mlp_fwd_rows", "radial_mlp_bwd_rows",
              "tpconv_fwd_shared", "tpconv_bwd_shared"):
    jax.ffi.register_ffi_target(f"eqnn_{_name}", jax.ffi.pycapsule(getattr(_shim, f"eqnn_{_name}_ffi")), platform="CUDA")
for _fn, _res, _args in (("eqnn_knn_workspace_bytes", ctypes.c_size_t, [ctypes.c_int] * 3),
                         ("eqnn_csr_workspace_bytes", ctypes.c_size_t, [ctypes.c_int] * 3),
                         ("eqnn_radial_mlp_bwd_workspace_bytes", ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int, ctypes.c_int]),
                         ("eqnn_radial_mlp_stash_bytes", ctypes.c_size_t, [ctypes.c_int64, ctypes.c_int]),
                         ("eqnn_edge_pairs_workspace_bytes", ctypes.c_size_t, [ctypes.c_int, ctypes.c_int64]),
                         ("eqnn_pair_combine_workspace_bytes", ctypes.c_size_t, [])):
    getattr(_core, _fn).restype, getattr(_core, _fn).argtypes = _res, _args

_f32, _i32, _u8 = jnp.float32, jnp.int32, jnp.uint8
_S = jax.ShapeDtypeStruct


def nearest_neighbors(x, k, cell=None):
    """x [B, N, F] -> senders, receivers [B, N k] (int32, bit-exact with the reference), dr [B, N k, 3]."""
    B, N, _ = x.shape
    c = np.asarray(cell, np.float32).reshape(-1) if cell is not None else np.zeros((0,), np.float32)
    ci = np.linalg.inv(np.asarray(cell, np.float64)).astype(np.float32).reshape(-1) if cell is not None else c
    ws = max(int(_core.eqnn_knn_workspace_bytes(B, N, k)), 1)
    out = (_S((B, N * k), _i32), _S((B, N * k), _i32), _S((B, N * k, 3), _f32), _S((ws,), _u8))
    s, r, dr, _ = jax.ffi.ffi_call("eqnn_knn_pbc", out)(x, cell=c, cell_inv=ci, k=np.int32(k), flags=np.int32(0))
    return s, r, dr


def csr_build(senders, receivers, n_nodes):
    B, E = receivers.shape
    ws = int(_core.eqnn_csr_workspace_bytes(B, n_nodes, E))
    out = (_S((B * n_nodes + 1,), _i32), _S((B * E,), _i32), _S((B * E,), _i32), _S((ws,), _u8))
    rowptr, perm, src, _ = jax.ffi.ffi_call("eqnn_csr_build", out)(senders, receivers, n_nodes=np.int32(n_nodes))
    return rowptr, perm, src


def edge_geom(pos, rowptr, src, n_rb, r_cut):
    E = src.shape[0]
    out = (_S((E, 4), _f32), _S((E, max(n_rb, 1)), _f32))
    return jax.ffi.ffi_call("eqnn_edge_geom", out)(pos, rowptr, src, n_rb=np.int32(n_rb), r_cut=float(r_cut))


def _pad3(kernels):
    hidden = list(kernels[:-1]) + [jnp.zeros((0, 128), _f32)] * (3 - (len(kernels) - 1))
    return hidden, kernels[-1]


@functools.partial(jax.custom_vjp, nondiff_argnums=(3, 4, 5))
def radial_mlp(geom, radial, kernels, n_rb, r_cut, act_scale):
    """kernels: list of the MLP's Dense kernels, hidden ones [fan_in, 128] and the (live-column) output [128, n_out]."""
    return _radial_fwd(geom, radial, kernels, n_rb, r_cut, act_scale)[0]


def _radial_fwd(geom, radial, kernels, n_rb, r_cut, act_scale):
    hidden, wout = _pad3(kernels)
    E, H = geom.shape[0], len(kernels) - 1
    stash = int(_core.eqnn_radial_mlp_stash_bytes(E, H)) // 4
    out = (_S((wout.shape[1], E), _f32), _S((stash,), _f32))
    w_t, z = jax.ffi.ffi_call("eqnn_radial_mlp_fwd", out)(geom, *hidden, wout, n_hidden=np.int32(H), n_rb=np.int32(n_rb),
                                                           r_cut=float(r_cut), activation=np.int32(0),
                                                           act_scale=np.float32(act_scale))
    return w_t, (geom, radial, z, kernels)


def _radial_bwd(n_rb, r_cut, act_scale, res, dw_t):
    geom, radial, z, kernels = res
    hidden, wout = _pad3(kernels)
    E, H = geom.shape[0], len(kernels) - 1
    ws = int(_core.eqnn_radial_mlp_bwd_workspace_bytes(E, n_rb, H))
    out = tuple(_S(h.shape, _f32) for h in hidden) + (_S(wout.shape, _f32), _S((ws,), _u8))
    *g, _ = jax.ffi.ffi_call("eqnn_radial_mlp_bwd", out)(geom, radial, z, *hidden, wout, dw_t, n_hidden=np.int32(H),
                                                          n_rb=np.int32(n_rb), r_cut=float(r_cut), activation=np.int32(0),
                                                          act_scale=np.float32(act_scale))
    return None, None, list(g[:H]) + [g[3]]      # edge lengths carry no gradient (params only, train_cosmology.py:245)


radial_mlp.defvjp(_radial_fwd, _radial_bwd)


@functools.partial(jax.custom_vjp, nondiff_argnums=(6, 7, 8))
def tpconv(xt, w_t, rowptr, src, perm, geom, tp_id, agg, d_live):
    return _tp_fwd(xt, w_t, rowptr, src, perm, geom, tp_id, agg, d_live)[0]


def _tp_fwd(xt, w_t, rowptr, src, perm, geom, tp_id, agg, d_live):
    """agg: 0 sum, 1 mean, 2 max (the arg-max of every component is kept for the backward pass)."""
    out = (_S((xt.shape[0], d_live), _f32), _S((xt.shape[0], d_live) if agg == 2 else (0,), _i32))
    A, amax = jax.ffi.ffi_call("eqnn_tpconv_fwd", out)(rowptr, src, geom, w_t, xt, tp_id=np.int32(tp_id), agg=np.int32(agg))
    return A, (xt, w_t, rowptr, src, perm, geom, amax)


def _tp_bwd(tp_id, agg, d_live, res, gA):
    xt, w_t, rowptr, src, perm, geom, amax = res
    out = (_S(w_t.shape, _f32), _S((src.shape[0], xt.shape[1]), _f32))
    dw_t, dxe = jax.ffi.ffi_call("eqnn_tpconv_bwd", out)(rowptr, src, perm, geom, w_t, xt, gA, amax, tp_id=np.int32(tp_id),
                                                          agg=np.int32(agg))
    dxt = dxe.reshape(xt.shape[0], -1, xt.shape[1]).sum(1)      # kNN graphs: k contiguous edges per sender
    return dxt, dw_t, None, None, None, None


tpconv.defvjp(_tp_fwd, _tp_bwd)


# ---- shared radial evaluations: the radial MLP once per unique edge length ---------------------
def edge_pairs(rowptr, src, geom, n_rb, r_cut):
    """-> dict(uidx [E], rep_slot, rep_partner, d_u, radial_u [E, n_rb] (capacity E rows), zero_slot [n_nodes],
    counts [2] = (n_unique, zero_row) ON THE DEVICE).  Every shape is static; the kernels read the live count themselves."""
    E, n_nodes = src.shape[0], rowptr.shape[0] - 1
    ws = int(_core.eqnn_edge_pairs_workspace_bytes(n_nodes, E))
    out = (_S((E,), _i32), _S((E,), _i32), _S((E,), _i32), _S((E,), _f32), _S((n_nodes,), _i32), _S((2,), _i32),
           _S((E, max(n_rb, 1)), _f32), _S((ws,), _u8))
    uidx, rep_slot, rep_partner, d_u, zero_slot, counts, radial_u, _ = jax.ffi.ffi_call("eqnn_edge_pairs", out)(
        rowptr, src, geom, n_rb=np.int32(n_rb), r_cut=float(r_cut))
    return dict(uidx=uidx, rep_slot=rep_slot, rep_partner=rep_partner, d_u=d_u, zero_slot=zero_slot, counts=counts,
                radial_u=radial_u)


@functools.partial(jax.custom_vjp, nondiff_argnums=(2, 3, 4))
def radial_mlp_rows(prs, kernels, n_rb, r_cut, act_scale):
    """w_rows [E, n_out rounded up to 4]: row u = the radial weights of unique edge length prs["d_u"][u]."""
    return _radial_rows_fwd(prs, kernels, n_rb, r_cut, act_scale)[0]


def _radial_rows_fwd(prs, kernels, n_rb, r_cut, act_scale):
    hidden, wout = _pad3(kernels)
    E, H = prs["d_u"].shape[0], len(kernels) - 1
    stash = int(_core.eqnn_radial_mlp_stash_bytes(E, H)) // 4
    out = (_S((E, (wout.shape[1] + 3) // 4 * 4), _f32), _S((stash,), _f32))
    w_rows, z = jax.ffi.ffi_call("eqnn_radial_mlp_fwd_rows", out)(prs["d_u"], prs["counts"], *hidden, wout, n_hidden=np.int32(H),
                                                                   n_rb=np.int32(n_rb), r_cut=float(r_cut),
                                                                   activation=np.int32(0), act_scale=np.float32(act_scale))
    return w_rows, (prs, z, kernels)


def _radial_rows_bwd(n_rb, r_cut, act_scale, res, dw_rows):
    """dw_rows: dL/dw per CSR SLOT (the cotangent tpconv_shared hands back); pairs are summed inside the handler."""
    prs, z, kernels = res
    hidden, wout = _pad3(kernels)
    E, H = prs["d_u"].shape[0], len(kernels) - 1
    ws = int(_core.eqnn_radial_mlp_bwd_workspace_bytes(E, n_rb, H))
    wsp = int(_core.eqnn_pair_combine_workspace_bytes())
    out = tuple(_S(h.shape, _f32) for h in hidden) + (_S(wout.shape, _f32), _S((wout.shape[1], E), _f32), _S((wsp,), _u8),
                                                       _S((ws,), _u8))
    *g, _, _, _ = jax.ffi.ffi_call("eqnn_radial_mlp_bwd_rows", out)(
        prs["d_u"], prs["counts"], prs["rep_slot"], prs["rep_partner"], prs["zero_slot"], prs["radial_u"], z, *hidden, wout,
        dw_rows, n_hidden=np.int32(H), n_rb=np.int32(n_rb), r_cut=float(r_cut), activation=np.int32(0),
        act_scale=np.float32(act_scale))
    return None, list(g[:H]) + [g[3]]


radial_mlp_rows.defvjp(_radial_rows_fwd, _radial_rows_bwd)


@functools.partial(jax.custom_vjp, nondiff_argnums=(7, 8, 9))
def tpconv_shared(xt, w_rows, uidx, rowptr, src, perm, geom, tp_id, agg, d_live):
    return _tps_fwd(xt, w_rows, uidx, rowptr, src, perm, geom, tp_id, agg, d_live)[0]


def _tps_fwd(xt, w_rows, uidx, rowptr, src, perm, geom, tp_id, agg, d_live):
    out = (_S((xt.shape[0], d_live), _f32), _S((xt.shape[0], d_live) if agg == 2 else (0,), _i32))
    A, amax = jax.ffi.ffi_call("eqnn_tpconv_fwd_shared", out)(rowptr, src, geom, uidx, w_rows, xt, tp_id=np.int32(tp_id),
                                                              agg=np.int32(agg))
    return A, (xt, w_rows, uidx, rowptr, src, perm, geom, amax)


def _tps_bwd(tp_id, agg, d_live, res, gA):
    xt, w_rows, uidx, rowptr, src, perm, geom, amax = res
    # the cotangent of w_rows is returned PER SLOT ([E, ld_rows], same shape as w_rows: capacity E rows); radial_mlp_rows'
    # backward rule sums the members of every unique row
    out = (_S(w_rows.shape, _f32), _S((src.shape[0], xt.shape[1]), _f32))
    dw_rows, dxe = jax.ffi.ffi_call("eqnn_tpconv_bwd_shared", out)(rowptr, src, perm, geom, uidx, w_rows, xt, gA, amax,
                                                                   tp_id=np.int32(tp_id), agg=np.int32(agg))
    dxt = dxe.reshape(xt.shape[0], -1, xt.shape[1]).sum(1)
    return dxt, dw_rows, None, None, None, None, None


tpconv_shared.defvjp(_tps_fwd, _tps_bwd)
