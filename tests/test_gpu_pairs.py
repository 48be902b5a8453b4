"""GPU parity of the shared radial evaluations (csrc/pairs.cu): the pairing itself is integer work and must match the
NumPy restatement exactly; the model-level results must not depend on whether the radial MLP (models/nequip.py:55-68)
is evaluated once per edge or once per unique edge length."""
import numpy as np
import pytest
import torch

from oracle import pairs_ref as PR

pytestmark = pytest.mark.gpu


def _dev(a, dtype=torch.float32):
    return torch.as_tensor(np.asarray(a), dtype=dtype).cuda()


def _pairs_of(senders, receivers, x, N, n_rb=0):
    from eqnn_jax_b200 import ops
    B = x.shape[0]
    rowptr, perm, src = ops.csr_build(_dev(senders, torch.int32), _dev(receivers, torch.int32), N)
    geom, _ = ops.edge_geom(_dev(x), x.shape[-1], B * N, rowptr, src, 0, 1.0, want_radial=False)
    pr = ops.edge_pairs(rowptr, src, geom, B * N, n_rb, 0.6)
    torch.cuda.synchronize()
    return rowptr.cpu().numpy(), src.cpu().numpy(), geom.cpu().numpy(), pr


def _check_against_restatement(rowptr, src, geom, pr):
    d = geom[:, 3]
    uidx, rep_slot, rep_partner, zero_slot, zero_row = PR.edge_pairs(rowptr, src, d)
    U = int(pr.n_unique.item())
    assert U == len(rep_slot)
    np.testing.assert_array_equal(pr.uidx.cpu().numpy(), uidx)
    np.testing.assert_array_equal(pr.rep_slot.cpu().numpy()[:U], rep_slot)
    np.testing.assert_array_equal(pr.rep_partner.cpu().numpy()[:U], rep_partner)
    np.testing.assert_array_equal(pr.zero_slot.cpu().numpy(), zero_slot)
    assert int(pr.zero_row.item()) == zero_row
    du = pr.d_u.cpu().numpy()[:U]
    # what makes the sharing exact: every slot reads a row whose length is its own length, bit for bit
    assert np.array_equal(du[uidx].view(np.uint32), d.view(np.uint32))
    return U


@pytest.mark.parametrize("B,N,k", [(1, 5, 5), (2, 64, 8), (3, 300, 20), (1, 1000, 32)])
def test_pairs_on_knn_graphs_match_restatement(B, N, k):
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(N + k)
    x = rng.uniform(0.0, 1.0, size=(B, N, 3)).astype(np.float32)
    g = GU.build_graph(_dev(x), None, k, apply_pbc=None)
    s = g.senders.reshape(B, -1).cpu().numpy()
    r = g.receivers.reshape(B, -1).cpu().numpy()
    rowptr, src, geom, pr = _pairs_of(s, r, x, N)
    U = _check_against_restatement(rowptr, src, geom, pr)
    E = B * N * k
    assert int(pr.zero_row.item()) >= 0 and (pr.zero_slot >= 0).all()      # every node's self-loop is in the zero class
    assert U <= E - (B * N - 1)       # ... which is ONE row
    if N >= 300:
        assert U < 0.65 * E           # most edges of a kNN graph are reciprocated


def test_pairs_on_ragged_graph_with_duplicates_and_one_way_edges():
    # duplicated edges pair at most once, one-directional edges and self-loops stay alone, an isolated node has an empty row
    N = 7
    senders = np.array([[0, 1, 0, 1, 2, 3, 3, 4, 4, 2, 5, 5, 0]], dtype=np.int32)
    receivers = np.array([[1, 0, 1, 0, 3, 2, 2, 4, 0, 2, 0, 0, 5]], dtype=np.int32)
    rng = np.random.default_rng(0)
    x = rng.normal(size=(1, N, 3)).astype(np.float32)
    rowptr, src, geom, pr = _pairs_of(senders, receivers, x, N)
    U = _check_against_restatement(rowptr, src, geom, pr)
    members = np.bincount(pr.uidx.cpu().numpy(), minlength=U)
    zr = int(pr.zero_row.item())
    assert zr >= 0 and members[zr] == 2                      # the self-loops of nodes 2 and 4 (first of each row only)
    assert members.min() >= 1 and np.delete(members, zr).max() <= 2


def test_bessel_rows_matches_edge_geom_basis():
    from eqnn_jax_b200 import ops
    from eqnn_jax_b200 import graph_utils as GU
    rng = np.random.default_rng(3)
    B, N, k = 2, 200, 12
    x = rng.uniform(0.0, 1.0, size=(B, N, 3)).astype(np.float32)
    g = GU.build_graph(_dev(x), None, k, apply_pbc=None)
    rowptr, perm, src = ops.csr_build(g.senders.reshape(B, -1).contiguous(), g.receivers.reshape(B, -1).contiguous(), N)
    geom, radial = ops.edge_geom(_dev(x), 3, B * N, rowptr, src, 64, 0.6)
    pr = ops.edge_pairs(rowptr, src, geom, B * N, 64, 0.6)
    U = int(pr.n_unique.item())
    got = pr.radial_u[:U].cpu().numpy()
    want = radial.cpu().numpy()[pr.rep_slot[:U].cpu().numpy().astype(np.int64)]
    # n_rb 16 / 32 / 64: the forward kernel's packed sine (1e-7 relative to libm's); other widths: edge_geom's arithmetic
    assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    from eqnn_jax_b200 import _lib
    import ctypes as C
    out = torch.empty((pr.cap, 24), dtype=torch.float32, device="cuda")
    _, radial24 = ops.edge_geom(_dev(x), 3, B * N, rowptr, src, 24, 0.6)
    _lib.check(_lib.lib().eqnn_bessel_rows(C.c_void_p(pr.d_u.data_ptr()), pr.cap, C.c_void_p(pr.n_unique.data_ptr()), 24, 0.6,
                                           C.c_void_p(out.data_ptr()), C.c_void_p(torch.cuda.current_stream().cuda_stream)),
               "bessel_rows")
    want24 = radial24.cpu().numpy()[pr.rep_slot[:U].cpu().numpy().astype(np.int64)]
    assert np.array_equal(out[:U].cpu().numpy().view(np.uint32), want24.view(np.uint32))


@pytest.mark.parametrize("agg,n_layers,n_rb", [("mean", 3, 64), ("max", 3, 64), ("sum", 2, 32)])
def test_shared_radial_evaluation_does_not_change_the_model(agg, n_layers, n_rb):
    """Same inputs, same parameters: forward output bit-identical (the same MLP arithmetic on the same lengths), every
    gradient equal up to the summation order of the weight gradients (1e-5 of the module's gradient scale)."""
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200.nequip import IrrepsArray, NequIP
    rng = np.random.default_rng(11)
    B, N, k = 2, 256, 16
    x = rng.normal(size=(B, N, 7)).astype(np.float32)
    x[..., :3] = rng.uniform(0.0, 3.0, size=(B, N, 3))
    kw = dict(d_hidden=128, l_max=2, sphharm_norm="integral", message_passing_steps=2, n_layers=n_layers, n_outputs=2,
              n_radial_basis=n_rb, r_cutoff=0.6, message_passing_agg=agg, task="graph")
    xd = _dev(x)
    g = GU.build_graph(xd, None, k, apply_pbc=GU.get_apply_pbc(np.ones(3, dtype=np.float32), cell=np.eye(3) * 3))
    gg = g._replace(nodes=IrrepsArray("1o+1o+1x0e", xd), edges=None)
    cot = _dev(rng.normal(size=(B, 2)))
    res = {}
    for share in (False, True):
        model = NequIP(**kw, share_radial=share)
        params = model.init(0, gg)
        pg = model.prepare(gg)
        assert (pg.pairs is not None) == share
        out = model.forward_backward(params, gg, lambda o: cot, prepared=pg)
        res[share] = (out.clone(), params.grad.clone(), params)
    assert torch.equal(res[False][0], res[True][0]), "forward output must be bit-identical"
    ga, gb = res[False][1], res[True][1]
    params = res[True][2]
    for path, shape, off in params.layout:
        n = int(np.prod(shape))
        a, b = ga[off:off + n], gb[off:off + n]
        # module scale: largest gradient among the tensors that share the first path element
        mod = max(float(ga[o:o + int(np.prod(sh))].abs().max()) for pth, sh, o in params.layout if pth[0] == path[0])
        assert float((a - b).abs().max()) <= 1e-5 * max(mod, 1e-30), path


@pytest.mark.parametrize("lmax,live,agg,shared", [(2, "0e,1o", "mean", True), (2, "0e,1o", "sum", False), (2, "0e,1o,2e", "max", False),
                                                  (3, "0e,1o", "max", True), (1, "0e", "mean", False)])
def test_warp_autonomous_interaction_kernels_equal_the_cta_kernels(lmax, live, agg, shared, monkeypatch):
    """eqnn_tpconv_fwd / _bwd (and the _shared forms): the warp-per-rows kernels (default) against the CTA-cooperative ones
    (EQNN_TPCONV_CTA=1) -- same per-row summation order, so every output is bit-identical."""
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200 import ops
    rng = np.random.default_rng(5 + lmax)
    B, N, k = 2, 700, 20
    x = rng.uniform(0.0, 1.0, size=(B, N, 3)).astype(np.float32)
    g = GU.build_graph(_dev(x), None, k, apply_pbc=GU.get_apply_pbc(None, cell=np.eye(3)))
    rowptr, perm, src = ops.csr_build(g.senders.reshape(B, -1).contiguous(), g.receivers.reshape(B, -1).contiguous(), N)
    Nt, E = B * N, B * N * k
    geom, _ = ops.edge_geom(_dev(x), 3, Nt, rowptr, src, 0, 1.0, want_radial=False)
    tp_id = ops.tp_lookup(f"x=1x0e+2x1o;lmax={lmax};live={live}")
    dxp, ny, n_live, d_live = ops.tp_dims(tp_id)
    gen = torch.Generator(device="cuda").manual_seed(0)
    xt = torch.randn((Nt, dxp), device="cuda", generator=gen)
    xt[:, 7:] = 0
    dA = torch.randn((Nt, d_live), device="cuda", generator=gen)
    a = ops.AGG[agg]
    ldr = (n_live + 3) // 4 * 4
    pr = ops.edge_pairs(rowptr, src, geom, Nt, 0, 1.0) if shared else None
    w_t = torch.randn((n_live, E), device="cuda", generator=gen)
    w_rows = torch.randn((E, ldr), device="cuda", generator=gen)
    res = {}
    for mode in ("warp", "cta"):
        if mode == "cta":
            monkeypatch.setenv("EQNN_TPCONV_CTA", "1")
        else:
            monkeypatch.delenv("EQNN_TPCONV_CTA", raising=False)
        A = torch.zeros((Nt, d_live), device="cuda")
        amax = torch.zeros((Nt, d_live), dtype=torch.int32, device="cuda") if a == 2 else None
        dxe = torch.zeros((E, dxp), device="cuda")
        if shared:
            dw = torch.zeros((E, ldr), device="cuda")
            ops.tpconv_fwd_shared(tp_id, rowptr, src, geom, pr, w_rows, xt, Nt, a, A, d_live, amax)
            ops.tpconv_bwd_shared(tp_id, rowptr, src, perm, geom, pr, w_rows, xt, Nt, a, dA, d_live, dw, dxe, amax)
        else:
            dw = torch.zeros((n_live, E), device="cuda")
            ops.tpconv_fwd(tp_id, rowptr, src, geom, w_t, xt, Nt, a, A, d_live, amax)
            ops.tpconv_bwd(tp_id, rowptr, src, perm, geom, w_t, xt, Nt, a, dA, d_live, dw, dxe, amax)
        torch.cuda.synchronize()
        res[mode] = (A, dw, dxe, amax)
    for i, name in enumerate(("A", "dw", "dxe")):
        assert torch.equal(res["warp"][i], res["cta"][i]), name
    if a == 2:
        assert torch.equal(res["warp"][3], res["cta"][3]), "argmax"


def test_shared_radial_at_cloud_scale():
    """bench.py's model on 4 x 5000-node kNN-20 clouds (the CPU oracle cannot reach this size): the zero class sums 20 000
    self-loop cotangents into one row, which also sets the power-of-two scale of the backward kernels' fp16 operand split.
    Forward bit-identical to the per-edge evaluation, every gradient outside the radial MLPs bit-identical, the radial MLPs'
    weight gradients within 1e-5 of their module's gradient scale (measured 1.4e-6 at 8 clouds, scripts/dev/diag_shared_scale.py)."""
    from eqnn_jax_b200 import graph_utils as GU
    from eqnn_jax_b200 import ops
    from eqnn_jax_b200.nequip import NequIP
    from eqnn_jax_b200.train import wrap_graph
    rng = np.random.default_rng(2)
    B, N, k = 4, 5000, 20
    h = rng.normal(size=(B, N, 7)).astype(np.float32)
    h[..., :3] = rng.uniform(0.0, 1.0, size=(B, N, 3)) / 0.2887
    hd = _dev(h)
    pbc = GU.get_apply_pbc(std=np.full(3, 0.2887, dtype=np.float32))
    s_, r_, _ = ops.knn_pbc(hd, k, pbc.cell, pbc.cell_inv, want_dr=False)
    kw = dict(d_hidden=128, l_max=2, sphharm_norm="integral", message_passing_steps=4, n_layers=3, message_passing_agg="mean",
              readout_agg="mean", mlp_readout_widths=(4, 2, 2), task="graph", n_outputs=2, n_radial_basis=64, r_cutoff=0.6)
    y = _dev(rng.normal(size=(B, 2)))
    res = {}
    for share in (False, True):
        model = NequIP(**kw, share_radial=share)
        g = wrap_graph(model, hd, s_, r_, k, pbc)
        params = model.init(0, g)
        out = model.forward_backward(params, g, lambda o: 2.0 * (o - y) / o.numel())
        res[share] = (out.clone(), params.grad.clone(), params)
    assert torch.equal(res[False][0], res[True][0])
    ga, gb, params = res[False][1], res[True][1], res[True][2]
    scale = {}
    for path, shape, off in params.layout:
        n = int(np.prod(shape))
        scale[path[0]] = max(scale.get(path[0], 0.0), float(ga[off:off + n].abs().max()))
    for path, shape, off in params.layout:
        n = int(np.prod(shape))
        a, b = ga[off:off + n], gb[off:off + n]
        if path[0].startswith("MultiLayerPerceptron"):
            assert float((a - b).abs().max()) <= 1e-5 * scale[path[0]], path
        else:
            assert torch.equal(a, b), path
