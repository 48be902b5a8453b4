"""Host logic (no GPU): plan tables, Linear->dense lowering, and the GENERATED tensor-product
code itself, compiled for the host with g++ and diffed against the oracle."""
import ctypes as C
import os
import subprocess
import tempfile

import numpy as np
import pytest
import torch

from eqnn_jax_b200 import codegen, so3 as pso3
from eqnn_jax_b200.irreps import Irreps, balanced_irreps
from eqnn_jax_b200.nequip import NequIP
from eqnn_jax_b200.plan import make_linear_plan, make_tp_plan
from oracle import e3nn_ref as E
from oracle import nequip_ref as NQ
from oracle import so3

HERE = os.path.dirname(os.path.abspath(__file__))
GEN = os.path.join(HERE, "..", "eqnn_jax_b200", "csrc", "gen", "tp_gen.cuh")

HARNESS = r'''
#include <cmath>
#include <cstring>
#define __device__
#define __forceinline__ inline
#define __restrict__
static int eqnn_fail(int c, const char*) { return c; }
enum { EQNN_ERR_SIGNATURE = 3 };
#include "GEN_PATH"
extern "C" int tp_count() { return EQNN_TP_COUNT; }
extern "C" const char* tp_sig(int i) { return EQNN_TP_SIGNATURES[i]; }
extern "C" int tp_fwd(int id, const float* u, const float* x, const float* w, float* m) {
  EQNN_TP_DISPATCH(id, { float Y[TP::NY]; TP::sh(u[0], u[1], u[2], Y); TP::fwd(x, Y, w, m); });
  return 0;
}
extern "C" int tp_bwd(int id, const float* u, const float* x, const float* w, const float* g, float* dw, float* dx) {
  EQNN_TP_DISPATCH(id, { float Y[TP::NY]; TP::sh(u[0], u[1], u[2], Y); TP::bwd(x, Y, w, g, dw, dx); });
  return 0;
}
extern "C" int tp_sh(int id, const float* u, float* Y) {
  EQNN_TP_DISPATCH(id, { TP::sh(u[0], u[1], u[2], Y); });
  return 0;
}
'''


@pytest.fixture(scope="module")
def host_tp():
    codegen.generate(GEN)
    d = tempfile.mkdtemp()
    src = os.path.join(d, "h.cpp")
    open(src, "w").write(HARNESS.replace("GEN_PATH", os.path.abspath(GEN)))
    so = os.path.join(d, "h.so")
    subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", src, "-o", so])
    return C.CDLL(so)


def _fa(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


@pytest.mark.parametrize("x_irreps,lmax,live", codegen.PREBUILT)
def test_generated_tp_matches_oracle(host_tp, x_irreps, lmax, live):
    plan = make_tp_plan(x_irreps, lmax, live)
    host_tp.tp_sig.restype = C.c_char_p
    sigs = [host_tp.tp_sig(i).decode() for i in range(host_tp.tp_count())]
    tid = sigs.index(plan.signature)
    rng = np.random.default_rng(lmax)
    xr_irreps = so3.Irreps(str(plan.irreps_x))
    for trial in range(3):
        v = rng.normal(size=3)
        u = v / np.linalg.norm(v)
        x = rng.normal(size=plan.dx)
        wfull = rng.normal(size=plan.n_irr)
        # oracle: tensor_product(x, Y_norm) * w, full message
        xt = E.IrrepsArray(xr_irreps, torch.tensor(x)[None])
        Y = E.IrrepsArray(so3.Irreps.spherical_harmonics(lmax),
                          torch.tensor(so3.spherical_harmonics(lmax, u[None], True, "norm")))
        tp = E.tensor_product(xt, Y)
        assert str(tp.irreps) == str(plan.irreps_msg)
        msg = tp.mul_by_scalars(torch.tensor(wfull)[None]).array[0].numpy()
        want = np.zeros(plan.d_live)
        for c in plan.columns:
            if c.live >= 0:
                want[c.live_off:c.live_off + 2 * c.l3 + 1] = msg[c.msg_off:c.msg_off + 2 * c.l3 + 1]
        xp = np.zeros(plan.dxp); xp[:plan.dx] = x
        wl = wfull[plan.live_cols]
        (ua, up), (xa, xpp), (wa, wp) = _fa(u), _fa(xp), _fa(wl)
        m = np.zeros(plan.d_live, dtype=np.float32)
        assert host_tp.tp_fwd(tid, up, xpp, wp, m.ctypes.data_as(C.POINTER(C.c_float))) == 0
        np.testing.assert_allclose(m, want, atol=3e-6 * max(1.0, np.abs(want).max()))
        # backward against finite structure: dw[c] = <g_c, T_c>, dx = J^T g
        g = rng.normal(size=plan.d_live)
        ga, gp = _fa(g)
        dw = np.zeros(plan.n_live, dtype=np.float32)
        dx = np.zeros(plan.dxp, dtype=np.float32)
        assert host_tp.tp_bwd(tid, up, xpp, wp, gp, dw.ctypes.data_as(C.POINTER(C.c_float)),
                              dx.ctypes.data_as(C.POINTER(C.c_float))) == 0
        xt2 = torch.tensor(x, requires_grad=True)
        wt2 = torch.tensor(wfull, requires_grad=True)
        tp2 = E.tensor_product(E.IrrepsArray(xr_irreps, xt2[None]), Y).mul_by_scalars(wt2[None]).array[0]
        gfull = torch.zeros(plan.irreps_msg.dim, dtype=torch.float64)
        for c in plan.columns:
            if c.live >= 0:
                gfull[c.msg_off:c.msg_off + 2 * c.l3 + 1] = torch.tensor(g[c.live_off:c.live_off + 2 * c.l3 + 1])
        (tp2 * gfull).sum().backward()
        np.testing.assert_allclose(dx[:plan.dx], xt2.grad.numpy(), atol=1e-5 * max(1.0, float(xt2.grad.abs().max())))
        np.testing.assert_allclose(dw, wt2.grad.numpy()[plan.live_cols],
                                   atol=1e-5 * max(1.0, float(wt2.grad.abs().max())))


def _dense_from_plan(lp, params):
    W = np.zeros((lp.rows, lp.cols))
    for r, c, name, flat, sc in lp.entries:
        W[r, c] += sc * np.asarray(params[name]).reshape(-1)[flat]
    return W


@pytest.mark.parametrize("ir_in,ir_out,force", [
    ("1o+1o+1x0e", "1o+1o+1x0e", False),
    ("3x0e+5x1o+2x1e+3x2e+2x2o+2x3o", "110x0e+14x1o+8x2e", False),
    ("1o+1o+1x0e", "110x0e+14x1o+8x2e", True),
    ("88x0e+14x1o+8x2e", "1o+1o+1x0e", False),
    ("1o+1o+1x0e", "1x1o", False),
])
def test_linear_lowering_matches_oracle(ir_in, ir_out, force):
    rng = np.random.default_rng(3)
    params = E.linear_init(rng, so3.Irreps(ir_in), so3.Irreps(ir_out), force_irreps_out=force)
    lp = make_linear_plan(ir_in, ir_out, force_irreps_out=force)
    assert {k: tuple(v.shape) for k, v in params.items()} == lp.param_shapes
    x = torch.tensor(rng.normal(size=(5, so3.Irreps(ir_in).dim)))
    want = E.linear_apply(NQ.to_torch(params), E.IrrepsArray(ir_in, x), so3.Irreps(ir_out), force).array.numpy()
    got = x.numpy() @ _dense_from_plan(lp, params)
    np.testing.assert_allclose(got, want, atol=1e-12)


def test_linear_lowering_regrouped_output():
    rng = np.random.default_rng(4)
    params = E.linear_init(rng, so3.Irreps("1o+1o+1x0e"), so3.Irreps("1o+1o+1x0e"))
    lp = make_linear_plan("1o+1o+1x0e", "1o+1o+1x0e", out_layout="regrouped", cols_pad=8)
    x = torch.tensor(rng.normal(size=(3, 7)))
    want = E.linear_apply(NQ.to_torch(params), E.IrrepsArray("1o+1o+1x0e", x), so3.Irreps("1o+1o+1x0e")).regroup().array
    got = x.numpy() @ _dense_from_plan(lp, params)
    np.testing.assert_allclose(got[:, :7], want.numpy(), atol=1e-12)
    assert np.all(got[:, 7] == 0)


@pytest.mark.parametrize("lmax,task", [(1, "node"), (2, "graph"), (3, "graph")])
def test_model_plan_param_tree_matches_oracle(lmax, task):
    kw = dict(d_hidden=128, l_max=lmax, sphharm_norm="integral", message_passing_steps=4, n_layers=3, n_outputs=2,
              n_radial_basis=64, r_cutoff=0.6, task=task)
    plan = NequIP(**kw).plan("1o+1o+1x0e")
    tree = NQ.init_params(NQ.NequIPConfig(**kw), "1o+1o+1x0e")
    flat = {}

    def walk(d, pre=()):
        for k, v in d.items():
            if isinstance(v, dict):
                walk(v, pre + (k,))
            else:
                flat[pre + (k,)] = tuple(np.shape(v))
    walk(tree)
    assert {p: s for p, s, _ in plan.layout} == flat
    assert plan._n_params == NQ.count_params(tree)
    assert balanced_irreps(lmax, 128) == str(so3.balanced_irreps(lmax, 128))


def test_sh_norm_factors():
    for l in range(4):
        for n in ("norm", "component", "integral"):
            assert pso3.sh_norm_factor(l, n) == pytest.approx(so3.sh_normalization_factor(l, n))
