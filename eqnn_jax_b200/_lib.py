"""ctypes binding of libeqnn_b200.so (the C ABI declared in include/eqnn_b200.h).

There is deliberately no fallback: if the shared library is missing or a call
returns a non-zero status, a RuntimeError is raised.  Nothing here touches the
CPU oracle.
"""
from __future__ import annotations

import ctypes as C
import os
import re
from typing import Dict, List

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libeqnn_b200.so")
HEADER_PATH = os.path.join(HERE, "..", "include", "eqnn_b200.h")

_p, _i, _i64, _f, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_float, C.c_double, C.c_size_t

# name -> (restype, argtypes); must list every function of include/eqnn_b200.h
SIGNATURES: Dict[str, tuple] = {
    "eqnn_last_error_string": (C.c_char_p, []),
    "eqnn_version": (_i, []),
    "eqnn_launch_count": (C.c_longlong, []),
    "eqnn_tp_signature_count": (_i, []),
    "eqnn_tp_signature": (C.c_char_p, [_i]),
    "eqnn_tp_lookup": (_i, [C.c_char_p]),
    "eqnn_tp_dims": (_i, [_i, _p, _p, _p, _p]),
    "eqnn_knn_workspace_bytes": (_sz, [_i, _i, _i]),
    "eqnn_knn_pbc": (_i, [_p, _i, _p, _p, _i, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "eqnn_edge_features": (_i, [_p, _i64, _i, _d, _p, _p]),
    "eqnn_csr_workspace_bytes": (_sz, [_i, _i, _i]),
    "eqnn_csr_build": (_i, [_p, _p, _i, _i, _i, _p, _p, _p, _p, _sz, _p]),
    "eqnn_edge_geom": (_i, [_p, _i, _i, _p, _p, _i, _d, _p, _p, _p]),
    "eqnn_tpconv_fwd": (_i, [_i, _p, _p, _p, _p, _i64, _p, _i, _i, _p, _i, _p, _p]),
    "eqnn_tpconv_bwd": (_i, [_i, _p, _p, _p, _p, _p, _i64, _p, _i, _i, _p, _i, _p, _p, _p, _p]),
    "eqnn_sender_reduce": (_i, [_p, _i, _i, _i, _p, _p]),
    "eqnn_segment_gather_sum": (_i, [_p, _p, _p, _i, _i, _p, _p]),
    "eqnn_sh_scatter": (_i, [_i, _p, _p, _i, _f, _p, _p]),
    "eqnn_node_tp_fwd": (_i, [_i, _p, _p, _i64, _p, _p]),
    "eqnn_node_tp_bwd": (_i, [_i, _p, _p, _p, _i64, _p, _p]),
    "eqnn_gemm_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "eqnn_gemm_f32": (_i, [_i64, _i64, _i64, _f, _p, _i64, _i64, _p, _i64, _i64, _p, _i64, _i64, _i, _i, _f, _i,
                           _p, _i64, _f, _p, _sz, _p]),
    "eqnn_radial_mlp_supported": (_i, [_i, _i, _i, _i]),
    "eqnn_radial_mlp_fwd": (_i, [_p, _i64, _i64, _i, _d, _i, _i, _p, _p, _i, _i, _f, _p, _i64, _p, _p]),
    "eqnn_radial_mlp_stash_bytes": (_sz, [_i64, _i]),
    "eqnn_radial_mlp_bwd_workspace_bytes": (_sz, [_i64, _i, _i]),
    "eqnn_radial_mlp_bwd": (_i, [_p, _i64, _i64, _i, _d, _i, _i, _p, _p, _i, _i, _f, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "eqnn_edge_pairs_workspace_bytes": (_sz, [_i, _i64]),
    "eqnn_edge_pairs": (_i, [_p, _p, _p, _i, _i64, _p, _p, _p, _p, _p, _p, _p, _p, _sz, _p]),
    "eqnn_pair_combine_workspace_bytes": (_sz, []),
    "eqnn_pair_combine": (_i, [_p, _i, _i, _p, _p, _p, _p, _i, _i64, _p, _p, _i64, _p, _sz, _p]),
    "eqnn_bessel_rows": (_i, [_p, _i64, _p, _i, _d, _p, _p]),
    "eqnn_radial_mlp_fwd_rows": (_i, [_p, _i64, _i64, _p, _i, _d, _i, _i, _p, _p, _i, _i, _f, _p, _i64, _p, _i, _p, _p]),
    "eqnn_radial_mlp_bwd_rows": (_i, [_p, _i64, _i64, _p, _i, _d, _i, _i, _p, _p, _i, _i, _f, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "eqnn_tpconv_fwd_shared": (_i, [_i, _p, _p, _p, _p, _p, _i, _p, _i, _i, _p, _i, _p, _p]),
    "eqnn_tpconv_bwd_shared": (_i, [_i, _p, _p, _p, _p, _p, _p, _i, _p, _i, _i, _p, _i, _p, _p, _p, _p]),
    "eqnn_gate_fwd": (_i, [_p, _i64, _i, _i, _i, _p, _p, _f, _f, _p, _p]),
    "eqnn_gate_bwd": (_i, [_p, _p, _i64, _i, _i, _i, _p, _p, _f, _f, _p, _p]),
    "eqnn_gate_act_fwd": (_i, [_p, _i64, _i, _i, _i, _p, _p, _i, _f, _f, _p, _p]),
    "eqnn_gate_act_bwd": (_i, [_p, _p, _i64, _i, _i, _i, _p, _p, _i, _f, _f, _p, _p]),
    "eqnn_node_update_workspace_bytes": (_sz, [_i, _i, _i, _i]),
    "eqnn_node_update_fwd": (_i, [_p, _i, _i, _p, _i, _p, _f, _p, _i, _i, _i, _p, _p, _f, _f, _p, _p, _i, _i64, _p, _p,
                                  _p, _p]),
    "eqnn_node_update_bwd": (_i, [_p, _i, _i, _p, _i, _p, _f, _p, _i, _i, _i, _p, _p, _f, _f, _p, _p, _p, _i64, _p, _p,
                                  _p, _p, _p, _p, _sz, _p]),
    "eqnn_egnn_geom_fwd": (_i, [_p, _i, _i, _p, _p, _p, _p, _i, _d, _p, _p, _p]),
    "eqnn_egnn_coord_fwd": (_i, [_p, _p, _i, _i, _p, _i, _p, _i, _p, _p, _p, _i, _p, _p]),
    "eqnn_egnn_coord_bwd": (_i, [_p, _p, _i, _i, _p, _p, _i, _p, _i, _d, _p, _p, _p, _p, _p, _p]),
    "eqnn_egnn_sender_acc": (_i, [_p, _p, _p, _i, _p, _p]),
    "eqnn_segnn_tplg_fwd": (_i, [_p, _i, _i, _p, _p, _i, _i, _i, _p, _i, _i, _i64, _p, _i, _p]),
    "eqnn_segnn_ycontract": (_i, [_p, _p, _i, _p, _i, _i, _i64, _i, _i, _p, _p]),
    "eqnn_segnn_yexpand": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p]),
    "eqnn_segnn_cg_apply": (_i, [_p, _i, _i, _p, _i, _i, _p, _p, _p, _p, _i, _i, _i64, _i, _p, _i, _p]),
    "eqnn_segnn_edge0_fwd": (_i, [_p, _p, _p, _i, _p, _p, _i, _i, _p, _i, _i, _p, _p, _i, _p, _p]),
    "eqnn_segnn_edge0_bwd": (_i, [_p, _p, _i, _i, _p, _p, _p, _i, _p, _p, _p]),
    "eqnn_segnn_edge0_wadd_workspace_bytes": (_sz, [_i, _i, _i]),
    "eqnn_segnn_edge0_wadd": (_i, [_p, _i, _p, _i, _p, _i, _i64, _p, _p, _sz, _p]),
    "eqnn_perm_compose": (_i, [_p, _p, _i64, _p, _p]),
    "eqnn_segnn_yagg_fwd": (_i, [_p, _i, _p, _i, _p, _i, _i, _p, _p]),
    "eqnn_segnn_yagg_bwd": (_i, [_p, _i, _p, _i, _p, _i, _i, _p, _p]),
    "eqnn_segnn_bias_deg_fwd": (_i, [_p, _i, _p, _i, _p, _i, _i, _p]),
    "eqnn_segnn_bias_deg_bwd": (_i, [_p, _i, _p, _i, _i, _i, _p, _p]),
    "eqnn_gather_rows": (_i, [_p, _p, _i64, _i, _p, _p]),
    "eqnn_axpy": (_i, [_i64, _f, _p, _p, _p]),
    "eqnn_steerable_attrs": (_i, [_p, _i, _p, _i, _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p, _p, _p, _p]),
    "eqnn_egnn_edge_input_fwd": (_i, [_p, _i, _p, _i, _i, _p, _i, _i, _p, _p, _i, _p, _p]),
    "eqnn_egnn_edge_input_bwd": (_i, [_p, _i, _i, _i, _p, _p, _p, _p, _i, _p, _p, _p, _i, _p]),
    "eqnn_egnn_xv_base": (_i, [_p, _i, _p, _p, _i64, _p, _p]),
    "eqnn_egnn_node_assemble": (_i, [_p, _i, _i, _p, _i64, _p, _p]),
    "eqnn_egnn_segment_rows_fwd": (_i, [_p, _i, _i, _p, _i, _i, _p, _i, _p, _p]),
    "eqnn_egnn_segment_rows_bwd": (_i, [_p, _i, _i, _p, _i, _p, _i, _i, _p, _p, _p]),
    "eqnn_egnn_xv_bwd": (_i, [_p, _i, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p]),
    "eqnn_copy_cols": (_i, [_p, _i, _i, _i, _i64, _i, _i, _p, _i, _i, _p]),
    "eqnn_softgate_fwd": (_i, [_p, _p, _i64, _p, _p]),
    "eqnn_softgate_bwd": (_i, [_p, _p, _p, _i64, _p, _p, _p]),
    "eqnn_dense_wgrad_workspace_bytes": (_sz, [_i64, _i64, _i64]),
    "eqnn_dense_wgrad_f32": (_i, [_i64, _i64, _i64, _f, _p, _i64, _i, _f, _p, _i64, _p, _i64, _p, _i, _p, _sz, _p]),
    "eqnn_colsum_workspace_bytes": (_sz, [_i]),
    "eqnn_colsum_tall": (_i, [_p, _i64, _i, _p, _p, _sz, _p]),
    "eqnn_tp_messages": (_i, [_i, _p, _p, _p, _p, _i, _p, _p, _p, _i64, _p, _i, _p]),
    "eqnn_radius_graph": (_i, [_p, _i, _p, _p, _i, _i, _f, _i, _p, _p, _p, _p, _p, _p]),
    "eqnn_exclusive_scan_i32": (_i, [_p, _i64, _p, _p]),
    "eqnn_pool_fwd": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "eqnn_pool_bwd": (_i, [_p, _i, _i, _i, _i, _p, _p]),
    "eqnn_pool_max_fwd": (_i, [_p, _i, _i, _i, _p, _p, _p]),
    "eqnn_pool_max_bwd": (_i, [_p, _p, _i, _i, _i, _p, _p]),
    "eqnn_colsum": (_i, [_p, _i64, _i, _p, _p]),
    "eqnn_layernorm_fwd": (_i, [_p, _i, _p, _i, _i, _p, _p, _f, _p, _p, _p]),
    "eqnn_layernorm_bwd": (_i, [_p, _i, _p, _i, _i, _p, _p, _p, _p, _p, _p, _p]),
    "eqnn_weights_expand": (_i, [_p, _p, _p, _i64, _p, _p]),
    "eqnn_weights_contract": (_i, [_p, _p, _p, _p, _p, _i64, _p, _p]),
    "eqnn_mse_loss": (_i, [_p, _p, _i, _i, _f, _p, _p, _p]),
    "eqnn_adamw": (_i, [_p, _p, _p, _p, _i64, _f, _f, _f, _f, _f, _i, _p]),
}

_lib = None


def declared_symbols() -> List[str]:
    """Function names declared in include/eqnn_b200.h."""
    text = open(HEADER_PATH).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(eqnn_[a-z0-9_]+)\s*\(", text)))


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m eqnn_jax_b200.build` "
                "(there is no CPU or PyTorch fallback for this path)")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def reload() -> C.CDLL:
    """Loads the library again after a rebuild (new tensor-product signature).  dlopen caches by path, so the rebuilt
    file is opened through a fresh hard link; the old image stays mapped but is no longer used."""
    global _lib, _generation
    _generation += 1
    link = f"{LIB_PATH}.{os.getpid()}.{_generation}.so"
    if os.path.exists(link):
        os.remove(link)
    try:
        os.link(LIB_PATH, link)
    except OSError:
        import shutil
        shutil.copy(LIB_PATH, link)
    L = C.CDLL(link)
    os.remove(link)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return _lib


_generation = 0


def check(status: int, what: str = "") -> None:
    if status != 0:
        msg = lib().eqnn_last_error_string().decode()
        raise RuntimeError(f"eqnn_b200 {what} failed (status {status}): {msg}")
