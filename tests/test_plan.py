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


# --------------------------------------------------------------------------------------------------
# SEGNN: parameter layout and dense lowering of TensorProductLinearGate (host logic, no GPU)
# --------------------------------------------------------------------------------------------------
def _leaves_(tree, pre=()):
    for k, v in tree.items():
        if isinstance(v, dict):
            yield from _leaves_(v, pre + (k,))
        else:
            yield pre + (k,), v


@pytest.mark.parametrize("kw,irr,lat,nadd", [
    (dict(d_hidden=128, l_max_hidden=1, n_layers=3, message_passing_steps=3, task="graph", output_irreps="1x0e",
          message_passing_agg="mean", n_outputs=2, residual=False), "1o", 1, 64),          # notebook cell 8: 573 253 parameters
    (dict(d_hidden=32, message_passing_steps=2, task="node", output_irreps="1o + 1o + 1x0e"), "1o+1o+1x0e", 1, 1),
    (dict(d_hidden=128, l_max_hidden=2, n_layers=3, message_passing_steps=3, task="node", output_irreps="1x1o"), "1o", 2, 1),
])
def test_segnn_plan_parameters_match_oracle(kw, irr, lat, nadd):
    from eqnn_jax_b200.segnn import SEGNN, _Plan
    from eqnn_jax_b200.irreps import Irreps
    from oracle import segnn_ref as S
    from oracle.so3 import Irreps as OI
    plan = _Plan(SEGNN(**kw), Irreps(irr), lat, nadd)
    tree = S.init_params(S.SEGNNConfig(**kw), irr, OI.spherical_harmonics(lat), f"{nadd}x0e")
    want = {p: tuple(v.shape) for p, v in _leaves_(tree)}
    got = {p: tuple(s) for p, s, _ in plan.layout}
    assert want == got
    if nadd == 64:
        assert plan._n_params == 573253


def _cg_rows(tab, ptr_off, n_rows, src, y):
    """Host statement of eqnn_segnn_cg_apply: out[r, o] = sum_t coef[t] y[r, idx[t] >> 24] src[r, idx[t] & 0xFFFFFF]."""
    out = np.zeros((src.shape[0], n_rows))
    ptr = tab["cg_ptr"][ptr_off:ptr_off + n_rows + 1]
    for o in range(n_rows):
        for t in range(ptr[o], ptr[o + 1]):
            q = int(tab["cg_idx"][t])
            out[:, o] += float(tab["cg_coef"][t]) * y[:, q >> 24] * src[:, q & 0xFFFFFF]
    return out


@pytest.mark.parametrize("lowering", ["blocked", "dense"])
def test_segnn_dense_lowering_equals_oracle_tplg(lowering):
    """Linear(TP(x, y)) + b evaluated on the host from the plan's tables -- dense: (x (x) y) @ M_cat; blocked: per-run
    multiplicity GEMMs into T, then the Clebsch-Gordan CSR table (and its transpose against the forward one) --
    equals the oracle's gate-less TensorProductLinearGate (segnn.py:48-54): the arithmetic the GPU path runs."""
    import torch
    from eqnn_jax_b200.segnn import SEGNN, _Plan
    from eqnn_jax_b200.irreps import Irreps
    from oracle import segnn_ref as S
    from oracle.e3nn_ref import IrrepsArray as OIA
    from oracle.so3 import Irreps as OI, spherical_harmonics
    rng = np.random.default_rng(0)
    kw = dict(d_hidden=32, l_max_hidden=2, n_layers=2, message_passing_steps=1, task="node", output_irreps="1o+1x0e")
    irr, lat, nadd = "1o+2x0e", 2, 3
    plan = _Plan(SEGNN(lowering=lowering, **kw), Irreps(irr), lat, nadd)
    tree = S.init_params(S.SEGNNConfig(**kw), irr, OI.spherical_harmonics(lat), f"{nadd}x0e", seed=3)
    theta = np.zeros((plan._n_params,))
    for path, shape, off in plan.layout:
        node = tree
        for q in path:
            node = node[q]
        theta[off:off + node.size] = (node + (0.3 if path[-1].startswith("b[") else 0.0)).reshape(-1)
    tab = plan._np
    wexp = np.where(tab["exp_idx"] >= 0, tab["exp_scale"].astype(np.float64) * theta[np.maximum(tab["exp_idx"], 0)], 0.0)
    checked = 0
    for t, x_irreps in [(plan.embed, irr), (plan.steps[0]["edge"][0], None), (plan.steps[0]["node"], None),
                        (plan.decode[0], None), (plan.decode[-1], None)]:
        R = 5
        x = rng.normal(size=(R, t.Dx))
        y = spherical_harmonics(lat, rng.normal(size=(R, 3)), True, "integral") if t.has_y else np.ones((R, 1))
        if lowering == "dense":
            M = wexp[t.M_off:t.M_off + t.Dx * t.Dy * t.Do].reshape(t.Dx, t.Do, t.Dy)
            z = np.einsum("ri,rj,ioj->ro", x, y, M)
        else:
            T = np.zeros((R, t.DT))
            first = set()
            for (c0, K, tb, Nr, W_off, acc) in t.runs:
                assert acc == (tb in first)
                first.add(tb)
                T[:, tb:tb + Nr] += x[:, c0:c0 + K] @ wexp[W_off:W_off + K * Nr].reshape(K, Nr)
            z = _cg_rows(tab, t.cg_f, t.Do, T, y)
            # the backward table is the transpose of the forward one: <dz, cg_f(T)> == <cg_b(dz), T>
            dz = rng.normal(size=z.shape)
            dT = _cg_rows(tab, t.cg_b, t.DT, dz, y)
            np.testing.assert_allclose((dz * z).sum(), (dT * T).sum(), rtol=1e-10)
            if not t.dx_full:
                cov = np.zeros(t.Dx, bool)
                for (c0, K, *_r) in t.runs:
                    cov[c0:c0 + K] = True
                assert not cov.all()
        if t.n_bias:
            z[:, t.b_col:t.b_col + t.n_bias] += theta[t.b_off:t.b_off + t.n_bias]
        # oracle: same parameters, x in its stored irreps layout
        p = {k: torch.tensor(v + (0.3 if k.startswith("b[") else 0.0)) for k, v in tree[t.name]["Linear_0"].items()}
        xi = OI(x_irreps) if x_irreps is not None else None
        if xi is None:                                   # recover the stored input irreps from the plan
            xi = {plan.steps[0]["edge"][0]: OI(f"{nadd}x0e") + OI(str(plan.embed.out_irreps)) + OI(str(plan.embed.out_irreps)),
                  plan.steps[0]["node"]: OI(str(plan.embed.out_irreps)) + OI(str(plan.steps[0]["edge"][-1].out_irreps)),
                  plan.decode[0]: OI(str(plan.h_irreps)), plan.decode[-1]: OI(str(plan.decode[-2].out_irreps))}[t]
        yo = OIA(OI.spherical_harmonics(lat), torch.tensor(y)) if t.has_y else None
        lin_out = S.tplg_linear_irreps(str(t.out_irreps) if not t.activation else str(plan.hidden), t.activation)
        from oracle import e3nn_ref as E
        yy = yo if yo is not None else OIA(S.ONE, torch.ones(R, 1, dtype=torch.float64))
        want = E.linear_apply(p, E.tensor_product(OIA(xi, torch.tensor(x)), yy), lin_out, gradient_normalization=0.0).array
        # the expand table stores the CG coefficients in fp32
        np.testing.assert_allclose(z, want.numpy(), atol=1e-6 * max(1.0, np.abs(want.numpy()).max()))
        checked += 1
    assert checked == 5


def test_segnn_node_level_edge_layer_identities():
    """The algebra behind segnn._edge0_fwd / _edgeL_fwd on the host, from the plan's own tables: (1) the first edge layer
    on concat(add, h_s, h_r) equals add @ M_a + (h @ M_s)[sender] + (h @ M_r)[receiver] contracted with y; (2) the last
    edge layer followed by the receiver aggregation equals sum_j S_j @ M_j + b w(n) with S_j[n] = f sum_p y[p, j] m[p]."""
    from eqnn_jax_b200.segnn import SEGNN, _Plan
    from eqnn_jax_b200.irreps import Irreps
    rng = np.random.default_rng(1)
    kw = dict(d_hidden=16, l_max_hidden=1, n_layers=3, message_passing_steps=1, task="node", output_irreps="1o")
    nadd, lat = 2, 1
    plan = _Plan(SEGNN(**kw), Irreps("1o"), lat, nadd)
    theta = rng.normal(size=(plan._n_params,))
    tab = plan._np
    wexp = np.where(tab["exp_idx"] >= 0, tab["exp_scale"].astype(np.float64) * theta[np.maximum(tab["exp_idx"], 0)], 0.0)
    N, k = 7, 3
    snd = rng.integers(0, N, size=(N * k,))
    rcv = np.repeat(np.arange(N), k)
    Dy = (lat + 1) ** 2
    y = rng.normal(size=(N * k, Dy))
    # (1) first edge layer
    t = plan.steps[0]["edge"][0]
    Dh = (t.Dx - nadd) // 2
    h, add = rng.normal(size=(N, Dh)), rng.normal(size=(N * k, nadd))
    M = wexp[t.M_off:t.M_off + t.Dx * t.Dy * t.Do].reshape(t.Dx, t.Do, t.Dy)
    U = np.concatenate([add, h[snd], h[rcv]], axis=1)
    want = np.einsum("ek,koj,ej->eo", U, M, y)
    Ps, Pr = np.einsum("nk,koj->noj", h, M[nadd:nadd + Dh]), np.einsum("nk,koj->noj", h, M[nadd + Dh:])
    got = np.einsum("eoj,ej->eo", np.einsum("ea,aoj->eoj", add, M[:nadd]) + Ps[snd] + Pr[rcv], y)
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
    # (2) last edge layer + aggregation (sum and mean), bias on the 0e block
    t = plan.steps[0]["edge"][-1]
    assert not t.activation
    M = wexp[t.M_off:t.M_off + t.Dx * t.Dy * t.Do].reshape(t.Dx, t.Do, t.Dy)
    b = np.zeros(t.Do)
    b[t.b_col:t.b_col + t.n_bias] = theta[t.b_off:t.b_off + t.n_bias]
    m = rng.normal(size=(N * k, t.Dx))
    z = np.einsum("ek,koj,ej->eo", m, M, y) + b
    for mean in (False, True):
        f = 1.0 / k if mean else 1.0
        want = f * np.add.reduceat(z, np.arange(0, N * k, k), axis=0)
        S = f * np.add.reduceat(y[:, :, None] * m[:, None, :], np.arange(0, N * k, k), axis=0)      # [n, j, k]
        got = np.einsum("njk,koj->no", S, M) + b * (f * k)
        np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12)
