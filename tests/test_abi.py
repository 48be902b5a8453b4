"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/eqnn_b200.h declares
(no compute calls: this runs without a GPU)."""
import ctypes as C
import os

import pytest


def test_library_exports_every_declared_symbol(built_lib):
    from eqnn_jax_b200 import _lib
    declared = _lib.declared_symbols()
    assert len(declared) >= 25
    for name in declared:
        assert hasattr(built_lib, name), f"{name} declared in include/eqnn_b200.h but not exported"
    assert set(declared) == set(_lib.SIGNATURES), "ctypes table out of sync with the header"


def test_signature_registry(built_lib):
    from eqnn_jax_b200 import codegen
    from eqnn_jax_b200.plan import make_tp_plan
    n = built_lib.eqnn_tp_signature_count()
    sigs = [built_lib.eqnn_tp_signature(i).decode() for i in range(n)]
    for x, l, live in codegen.PREBUILT:
        p = make_tp_plan(x, l, live)
        i = built_lib.eqnn_tp_lookup(p.signature.encode())
        assert i >= 0 and sigs[i] == p.signature
        v = [C.c_int() for _ in range(4)]
        assert built_lib.eqnn_tp_dims(i, *[C.byref(t) for t in v]) == 0
        assert [t.value for t in v] == [p.dxp, p.ny, p.n_live, p.d_live]
    assert built_lib.eqnn_tp_lookup(b"x=7x9e;lmax=1;live=0e") == -1


def test_argument_errors_are_reported_not_raised(built_lib):
    # invalid arguments are rejected on the host before any launch: safe without a GPU
    rc = built_lib.eqnn_knn_pbc(None, 3, None, None, 1, 5, 3, 0, None, None, None, None, 0, None)
    assert rc == 1 and b"null" in built_lib.eqnn_last_error_string()
    rc = built_lib.eqnn_tp_dims(999, None, None, None, None)
    assert rc == 3


def test_no_oracle_import_in_product():
    root = os.path.join(os.path.dirname(__file__), "..", "eqnn_jax_b200")
    for dp, _, fs in os.walk(root):
        for f in fs:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
