"""GPU parity of the dense building block, gate, pooling and parameter plumbing against plain torch fp64."""
import math

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def gelu(x):
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def gelu_grad(x):
    x = x.clone().requires_grad_(True)
    gelu(x).sum().backward()
    return x.grad


def _rel(got, want):
    return float((got.double().cpu() - want).abs().max() / want.abs().max().clamp_min(1e-30))


@pytest.mark.parametrize("M,N,K", [(1, 1, 1), (8, 2, 256), (40000, 192, 33), (40000, 7, 170), (5000, 8, 7),
                                   (70001, 128, 64), (30000, 128, 128), (30000, 11, 128), (300, 130, 50),
                                   (40001, 18, 30), (7, 8, 50000), (18, 30, 20001), (8, 7, 100001)])   # skinny kernels
@pytest.mark.parametrize("ta,tb,tc", [(False, False, False), (True, False, False), (False, True, False),
                                      (True, True, True)])
def test_gemm_layouts(M, N, K, ta, tb, tc):
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(M + N + K)
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(K, N, generator=g, dtype=torch.float64)
    want = 0.37 * (A @ B)
    Ad = (A.t().contiguous() if ta else A).float().cuda()
    Bd = (B.t().contiguous() if tb else B).float().cuda()
    Cd = torch.full((N, M) if tc else (M, N), float("nan"), dtype=torch.float32).cuda()
    _mm(M, N, K, 0.37, Ad, 0, M if ta else K, Bd, 0, K if tb else N, Cd, 0, M if tc else N, ta=ta, tb=tb, tc=tc)
    got = Cd.t() if tc else Cd
    assert _rel(got, want) < 2e-6


def test_gemm_split_k_weight_gradient_shape_is_deterministic():
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(0)
    E, K, N = 200_000, 128, 128
    Z = torch.randn(E, K, generator=g).cuda()
    dZ = torch.randn(E, N, generator=g).cuda()
    outs = []
    for _ in range(2):
        C = torch.empty(K, N).cuda()
        _mm(K, N, E, 0.1, Z, 0, K, dZ, 0, N, C, 0, N, ta=True, pro=1, pro_scale=1.5)
        outs.append(C.clone())
    assert torch.equal(outs[0], outs[1])
    want = 0.1 * ((gelu(Z.double().cpu()) * 1.5).t() @ dZ.double().cpu())
    assert _rel(outs[0], want) < 1e-5


def test_gemm_prologue_epilogue_accumulate_offsets():
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(1)
    M, N, K = 999, 70, 45
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(K, N, generator=g, dtype=torch.float64)
    bias = torch.randn(N, generator=g, dtype=torch.float64)
    aux = torch.randn(M, N, generator=g, dtype=torch.float64)
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64)
    buf = torch.zeros(5 + K * N).cuda()
    buf[5:] = B.float().reshape(-1).cuda()
    # bias epilogue + gelu prologue, B read at an element offset
    C = torch.empty(M, N).cuda()
    _mm(M, N, K, 1.0, A.float().cuda(), 0, K, buf, 5, N, C, 0, N, pro=1, pro_scale=0.7, epi=1, aux=bias.float().cuda())
    assert _rel(C, (gelu(A) * 0.7) @ B + bias) < 3e-6
    # derivative epilogue + accumulate
    C = C0.float().cuda()
    _mm(M, N, K, 2.0, A.float().cuda(), 0, K, buf, 5, N, C, 0, N, epi=2, aux=aux.float().cuda(), ld_aux=N,
        epi_scale=1.3, accumulate=True)
    assert _rel(C, C0 + 2.0 * (A @ B) * gelu_grad(aux) * 1.3) < 3e-6


def test_gemm_skinny_kernels_accumulate_and_determinism():
    """Node-level Linear shapes: N-row GEMM with a few columns, and its weight gradient (K = #rows)."""
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(3)
    M, N, K = 60000, 7, 8
    A = torch.randn(M, K, generator=g, dtype=torch.float64)
    B = torch.randn(K, N, generator=g, dtype=torch.float64)
    C0 = torch.randn(M, N, generator=g, dtype=torch.float64)
    C = C0.float().cuda()
    _mm(M, N, K, 1.5, A.float().cuda(), 0, K, B.t().contiguous().float().cuda(), 0, K, C, 0, N, tb=True, accumulate=True)
    assert _rel(C, C0 + 1.5 * (A @ B)) < 2e-6
    G = torch.randn(M, N, generator=g, dtype=torch.float64)
    outs = []
    for _ in range(2):
        W = torch.full((K, N), 2.0).cuda()
        _mm(K, N, M, 0.5, A.float().cuda(), 0, K, G.float().cuda(), 0, N, W, 0, N, ta=True, accumulate=True)
        outs.append(W.clone())
    assert torch.equal(outs[0], outs[1])
    assert _rel(outs[0], 2.0 + 0.5 * (A.t() @ G)) < 2e-6


@pytest.mark.parametrize("n", [77, 5003])           # 5003: the tall-input kernels (thread = column, rows walked)
def test_gate_forward_backward(n):
    from eqnn_jax_b200 import ops
    g = torch.Generator().manual_seed(2)
    n_extra, mul, l = 9, [3, 2], [1, 2]
    n_gates = sum(mul)
    dg = sum(m * (2 * ll + 1) for m, ll in zip(mul, l))
    z = torch.randn(n, n_extra + n_gates + dg, generator=g, dtype=torch.float64, requires_grad=True)
    ca, cg = 1 / 0.652, 1 / 0.5416
    ext = gelu(z[:, :n_extra]) * ca
    gates = torch.sigmoid(z[:, n_extra:n_extra + n_gates]) * cg
    rep = torch.tensor([2 * ll + 1 for m, ll in zip(mul, l) for _ in range(m)])
    want = torch.cat([ext, torch.repeat_interleave(gates, rep, dim=1) * z[:, n_extra + n_gates:]], 1)
    cot = torch.randn(want.shape, generator=g, dtype=torch.float64)
    (want * cot).sum().backward()
    zd = z.detach().float().cuda()
    out = torch.empty(n, n_extra + dg).cuda()
    ops.gate_fwd(zd, n, n_extra, n_gates, mul, l, ca, cg, out)
    assert _rel(out, want.detach()) < 2e-6
    gz = torch.empty_like(zd)
    ops.gate_bwd(zd, cot.float().cuda(), n, n_extra, n_gates, mul, l, ca, cg, gz)
    assert _rel(gz, z.grad) < 2e-6


def test_pool_colsum_mse_adamw():
    from eqnn_jax_b200 import ops
    g = torch.Generator().manual_seed(3)
    B, N, D = 3, 211, 70
    x = torch.randn(B, N, D, generator=g, dtype=torch.float64)
    out = torch.empty(B, D).cuda()
    ops.pool_fwd(x.float().cuda(), B, N, D, 1, out)
    assert _rel(out, x.mean(1)) < 2e-6
    gx = torch.empty(B, N, D).cuda()
    ops.pool_bwd(out, B, N, D, 0, gx)
    assert torch.equal(gx[:, 5], out)
    y = torch.empty(D).cuda()
    ops.colsum(x.float().cuda().reshape(B * N, D), B * N, D, y)
    assert _rel(y, x.reshape(-1, D).sum(0)) < 5e-6
    pred, tgt = torch.randn(4, 2, generator=g), torch.randn(4, 2, generator=g)
    loss, gp = torch.empty(1).cuda(), torch.empty(4, 2).cuda()
    ops.mse_loss(pred.cuda(), tgt.cuda(), 1.0, loss, gp)
    assert _rel(loss, ((pred - tgt).double() ** 2).mean().reshape(1)) < 1e-6
    assert _rel(gp, 2 * (pred - tgt).double() / 8) < 1e-6
    p = torch.randn(1000, generator=g)
    gr = torch.randn(1000, generator=g)
    ref = torch.nn.Parameter(p.clone().double())
    opt = torch.optim.AdamW([ref], lr=3e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-5)
    pd, m, v = p.clone().cuda(), torch.zeros(1000).cuda(), torch.zeros(1000).cuda()
    for step in (1, 2, 3):
        ref.grad = gr.double()
        opt.step()
        ops.adamw(pd, gr.cuda(), m, v, 3e-4, 0.9, 0.999, 1e-8, 1e-5, step)
    assert _rel(pd, ref.detach()) < 1e-6


@pytest.mark.parametrize("case", ["w128", "w64", "w128x11", "w64x11"])
@pytest.mark.parametrize("E", [4096, 70001, 1_000_003])
def test_tensor_core_3xtf32_weight_gradient(case, E):
    """tcgen05 weight-gradient kernel: dW = alpha * act(Z)^T @ dZ with the edge axis as K."""
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(E)
    inv_c = 1.0 / 0.652
    Mi = 128 if case.startswith("w128") else 64
    pro = 1 if Mi == 128 else 0
    Z = (torch.randn(E, Mi, generator=g) * 1.5).cuda()
    Zd = Z.double().cpu()
    act = gelu(Zd) * inv_c if pro else Zd
    outs = []
    if case.endswith("x11"):
        dWt = torch.randn(11, E, generator=g).cuda()
        want = 0.1 * act.t() @ dWt.double().cpu().t()
        for _ in range(2):
            C = torch.full((Mi, 11), float("nan")).cuda()
            _mm(Mi, 11, E, 0.1, Z, 0, Mi, dWt, 0, E, C, 0, 11, ta=True, tb=True, pro=pro, pro_scale=inv_c, tf32x3=True)
            outs.append(C)
    else:
        dZ = torch.randn(E, 128, generator=g).cuda()
        want = 0.1 * act.t() @ dZ.double().cpu()
        for _ in range(2):
            C = torch.full((Mi, 128), float("nan")).cuda()
            _mm(Mi, 128, E, 0.1, Z, 0, Mi, dZ, 0, 128, C, 0, 128, ta=True, pro=pro, pro_scale=inv_c, tf32x3=True)
            outs.append(C)
    torch.cuda.synchronize()
    assert torch.equal(outs[0], outs[1]), "weight gradients must be bit-reproducible"
    assert _rel(outs[0], want) < 3e-6


def test_tensor_core_egnn_shapes():
    """K = 32 with bias epilogue, N = 32 data-gradient, M = 32 weight-gradient, N = 1 last layer (EGNN edge MLPs)."""
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(9)
    E = 70016
    F = torch.randn(E, 32, generator=g, dtype=torch.float64)
    K0 = torch.randn(32, 128, generator=g, dtype=torch.float64)
    b0 = torch.randn(128, generator=g, dtype=torch.float64)
    dZ = torch.randn(E, 128, generator=g, dtype=torch.float64)
    Z = torch.empty(E, 128).cuda()
    _mm(E, 128, 32, 1.0, F.float().cuda(), 0, 32, K0.float().cuda(), 0, 128, Z, 0, 128, epi=1, aux=b0.float().cuda(),
        tf32x3=True)
    assert _rel(Z, F @ K0 + b0) < 3e-6
    Z1 = (torch.randn(E, 128, generator=g, dtype=torch.float64) * 2)
    Z2 = torch.empty(E, 128).cuda()
    _mm(E, 128, 128, 1.0, Z1.float().cuda(), 0, 128, K0.float().cuda().new_tensor(torch.randn(128, 128, generator=g)), 0,
        128, Z2, 0, 128, pro=1, pro_scale=1.0, epi=1, aux=b0.float().cuda(), tf32x3=True)
    dF = torch.empty(E, 32).cuda()
    _mm(E, 32, 128, 1.0, dZ.float().cuda(), 0, 128, K0.float().cuda(), 0, 128, dF, 0, 32, tb=True, tf32x3=True)
    assert _rel(dF, dZ @ K0.t()) < 3e-6
    dK = torch.empty(32, 128).cuda()
    _mm(32, 128, E, 1.0, F.float().cuda(), 0, 32, dZ.float().cuda(), 0, 128, dK, 0, 128, ta=True, tf32x3=True)
    assert _rel(dK, F.t() @ dZ) < 3e-6
    kl = torch.randn(128, 1, generator=g, dtype=torch.float64)
    t = torch.empty(E).cuda()
    _mm(E, 1, 128, 1.0, Z1.float().cuda(), 0, 128, kl.float().cuda(), 0, 1, t, 0, E, tc=True, pro=1, pro_scale=1.0,
        tf32x3=True)
    assert _rel(t, (gelu(Z1) @ kl)[:, 0]) < 3e-6
    dt = torch.randn(E, generator=g, dtype=torch.float64)
    dkl = torch.empty(128, 1).cuda()
    _mm(128, 1, E, 1.0, Z1.float().cuda(), 0, 128, dt.float().cuda(), 0, E, dkl, 0, 1, ta=True, tb=True, pro=1,
        pro_scale=1.0, tf32x3=True)
    assert _rel(dkl, gelu(Z1).t() @ dt[:, None]) < 3e-6
    dU = torch.empty(E, 128).cuda()
    _mm(E, 128, 1, 1.0, dt.float().cuda(), 0, E, kl.float().cuda(), 0, 1, dU, 0, 128, ta=True, tb=True, epi=2,
        aux=Z1.float().cuda(), ld_aux=128, epi_scale=1.0, tf32x3=True)
    assert _rel(dU, dt[:, None] * kl.t() * gelu_grad(Z1)) < 3e-6
    y = torch.empty(128).cuda()
    from eqnn_jax_b200 import ops
    ops.colsum_tall(dZ.float().cuda(), E, 128, y)
    assert _rel(y, dZ.sum(0)) < 1e-5


@pytest.mark.parametrize("case", ["fwd64", "fwd128", "dgrad128", "last16", "dgrad_last"])
@pytest.mark.parametrize("M", [4096, 70001])
def test_tensor_core_3xtf32_gemm_matches_fp64(case, M):
    """tcgen05 radial-MLP GEMMs (operands split into two TF32 halves, 3 MMAs per product)."""
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(M)
    inv_c = 1.0 / 0.652
    if case in ("fwd64", "fwd128"):
        K, N = (64, 128) if case == "fwd64" else (128, 128)
        pro = 0 if case == "fwd64" else 1
        A = torch.randn(M, K, generator=g, dtype=torch.float64) * (3.0 if pro else 1.0)
        B = torch.randn(K, N, generator=g, dtype=torch.float64)
        want = 0.125 * ((gelu(A) * inv_c if pro else A) @ B)
        C = torch.full((M, N), float("nan")).cuda()
        _mm(M, N, K, 0.125, A.float().cuda(), 0, K, B.float().cuda(), 0, N, C, 0, N, pro=pro, pro_scale=inv_c,
            tf32x3=True)
        got = C
    elif case == "dgrad128":
        K, N = 128, 128
        dZ = torch.randn(M, K, generator=g, dtype=torch.float64)
        W = torch.randn(N, K, generator=g, dtype=torch.float64)      # stored [in][out]; B = W^T
        Zp = torch.randn(M, N, generator=g, dtype=torch.float64) * 2
        want = 0.1 * (dZ @ W.t()) * gelu_grad(Zp) * inv_c
        C = torch.full((M, N), float("nan")).cuda()
        _mm(M, N, K, 0.1, dZ.float().cuda(), 0, K, W.float().cuda(), 0, K, C, 0, N, tb=True, epi=2,
            aux=Zp.float().cuda(), ld_aux=N, epi_scale=inv_c, tf32x3=True)
        got = C
    elif case == "last16":
        K, N = 128, 11
        A = torch.randn(M, K, generator=g, dtype=torch.float64) * 2
        B = torch.randn(K, N, generator=g, dtype=torch.float64)
        want = (0.09 * (gelu(A) * inv_c) @ B).t()
        C = torch.full((N, M), float("nan")).cuda()
        _mm(M, N, K, 0.09, A.float().cuda(), 0, K, B.float().cuda(), 0, N, C, 0, M, tc=True, pro=1, pro_scale=inv_c,
            tf32x3=True)
        got = C
    else:
        K, N = 11, 128
        dWt = torch.randn(K, M, generator=g, dtype=torch.float64)     # column-major A
        W4 = torch.randn(N, K, generator=g, dtype=torch.float64)      # [in=128][live=11]; B = W4^T
        Zp = torch.randn(M, N, generator=g, dtype=torch.float64) * 2
        want = 0.09 * (dWt.t() @ W4.t()) * gelu_grad(Zp) * inv_c
        C = torch.full((M, N), float("nan")).cuda()
        _mm(M, N, K, 0.09, dWt.float().cuda(), 0, M, W4.float().cuda(), 0, K, C, 0, N, ta=True, tb=True, epi=2,
            aux=Zp.float().cuda(), ld_aux=N, epi_scale=inv_c, tf32x3=True)
        got = C
    torch.cuda.synchronize()
    assert not torch.isnan(got).any()
    assert _rel(got, want) < 3e-6


@pytest.mark.parametrize("case", ["fwd", "fwd_acc", "dgrad", "wgrad", "wgrad_acc", "wgrad_long", "small_k", "offsets"])
def test_general_tensor_core_gemm_matches_fp64(case):
    """General 3xTF32 tcgen05 GEMM (EQNN_GEMM_TF32X3_ANY; the SEGNN dense-lowered layers): unaligned leading
    dimensions, row / column tails, transposed operands, K split over CTAs, accumulation -- against fp64."""
    from eqnn_jax_b200.nequip import _mm
    g = torch.Generator().manual_seed(len(case))
    r64 = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    dev = lambda t: t.float().cuda()
    if case in ("fwd", "fwd_acc"):                       # x @ M_cat: lda 341 (not a multiple of 4), row tail
        M, N, K = 3001, 768, 341
        A, B, C0 = r64(M, K), r64(K, N), r64(M, N)
        C = dev(C0)
        _mm(M, N, K, 0.5, dev(A), 0, K, dev(B), 0, N, C, 0, N, accumulate=case == "fwd_acc", tc_any=True)
        want = 0.5 * (A @ B) + (C0 if case == "fwd_acc" else 0.0)
    elif case == "dgrad":                                # dXbig @ M_cat^T: B transposed, column tail (341)
        M, N, K = 2500, 341, 768
        A, W = r64(M, K), r64(N, K)
        C = torch.full((M, N), float("nan")).cuda()
        _mm(M, N, K, 1.0, dev(A), 0, K, dev(W), 0, K, C, 0, N, tb=True, tc_any=True)
        want = A @ W.t()
    elif case in ("wgrad", "wgrad_acc", "wgrad_long"):   # x^T @ dXbig: A transposed, K = rows, split over CTAs
        M, N, K = 170, 768, (70001 if case == "wgrad_long" else 20011)   # long: accumulation stays <= 1024 per split
        X, D, C0 = r64(K, M), r64(K, N), r64(M, N)
        C = dev(C0)
        _mm(M, N, K, 1.0, dev(X), 0, M, dev(D), 0, N, C, 0, N, ta=True, accumulate=case == "wgrad_acc", tc_any=True)
        want = X.t() @ D + (C0 if case == "wgrad_acc" else 0.0)
    elif case == "small_k":                              # K below one chunk, N below one tile
        M, N, K = 70001, 40, 20
        A, B = r64(M, K), r64(K, N)
        C = torch.full((M, N), float("nan")).cuda()
        _mm(M, N, K, 2.0, dev(A), 0, K, dev(B), 0, N, C, 0, N, tc_any=True)
        want = 2.0 * (A @ B)
    else:                                                # sub-blocks of wider buffers at odd element offsets
        M, N, K = 1000, 130, 89
        Ab, Bb, Cb = r64(M, 341), r64(K * 200 + 7), r64(M, 633)
        C = dev(Cb)
        _mm(M, N, K, 1.0, dev(Ab), 89, 341, dev(Bb), 3, 200, C, 125, 633, tc_any=True)
        want = Cb.clone()
        want[:, 125:125 + N] = Ab[:, 89:89 + K] @ Bb[3:3 + K * 200].reshape(K, 200)[:, :N]
    torch.cuda.synchronize()
    assert torch.isfinite(C).all()
    err = (C.double().cpu() - want).abs().max().item()
    assert err <= 1e-5 * want.abs().max().item(), (case, err, want.abs().max().item())


@pytest.mark.parametrize("M,N,K,pro,tc", [(128, 128, 70001, 1, True), (32, 128, 20000, 0, True), (64, 128, 5000, 1, True),
                                           (128, 128, 3000, 0, True), (48, 128, 9000, 1, True), (128, 128, 9000, 1, False),
                                           (128, 128, 1, 0, True), (64, 128, 0, 0, True)])
def test_dense_wgrad_kernel_and_bias_in_one_pass(M, N, K, pro, tc):
    """eqnn_dense_wgrad_f32: gW = pro(A)^T G and gb = colsum(G); on the tcgen05 weight-gradient kernel the column sums
    ride along with the operand transform (first three cases), other shapes fall back to GEMM + column sum."""
    from eqnn_jax_b200 import ops
    g = torch.Generator().manual_seed(M + K)
    A = torch.randn(K, M, generator=g, dtype=torch.float64) * (2.0 if pro else 1.0)
    G = torch.randn(K, N, generator=g, dtype=torch.float64)
    inv_c = 1.0 / 0.652
    want_w = (gelu(A) * inv_c if pro else A).t() @ G
    want_b = G.sum(0)
    buf = torch.full((M * N + N + 5,), float("nan")).cuda()
    ops.dense_wgrad(M, N, K, A.float().cuda(), 0, M, G.float().cuda(), N, buf, 3, N, buf, 3 + M * N, pro=pro, pro_scale=inv_c,
                    tf32x3=tc)
    got_w = buf[3:3 + M * N].view(M, N).double().cpu()
    got_b = buf[3 + M * N:3 + M * N + N].double().cpu()
    assert torch.isnan(buf[:3]).all() and torch.isnan(buf[3 + M * N + N:]).all()
    assert (got_w - want_w).abs().max().item() <= 1e-5 * max(want_w.abs().max().item(), 1e-30)
    assert (got_b - want_b).abs().max().item() <= 1e-5 * max(want_b.abs().max().item(), K ** 0.5)


@pytest.mark.parametrize("R,Dx,Dy,Do,nb", [(3001, 341, 4, 192, 110), (70001, 170, 4, 170, 88), (4000, 170, 8, 30, 0),
                                           (3000, 64, 1, 96, 96), (5000, 40, 2, 50, 7)])
def test_segnn_fused_forward_matches_fp64(R, Dx, Dy, Do, nb):
    """eqnn_segnn_tplg_fwd: z = bias + sum_j y_j (x @ M_cat)[:, o*Dy + j] with the contraction in the GEMM epilogue --
    against fp64 (row tails, partial column tiles, unaligned z rows, every supported Dy)."""
    from eqnn_jax_b200 import ops
    g = torch.Generator().manual_seed(R + Do)
    r64 = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x, y, M, b = r64(R, Dx + 1), r64(R, Dy), r64(Dx, Do * Dy), r64(max(nb, 1))
    assert ops.segnn_fwd_fused_ok(R, Dx, Dy, Do)
    dev = lambda t: t.float().cuda().contiguous()
    z = torch.full((R, Do), float("nan")).cuda()
    bcol = 3 if nb and nb + 3 <= Do else 0
    ops.segnn_tplg_fwd(dev(x), Dx + 1, Dx, dev(M), 0, dev(y), Dy, Do, dev(b) if nb else None, 0, bcol, nb, R, z)
    want = torch.einsum("rk,koj,rj->ro", x[:, :Dx], M.reshape(Dx, Do, Dy), y)
    if nb:
        want[:, bcol:bcol + nb] += b[:nb]
    assert torch.isfinite(z).all(), int((~torch.isfinite(z)).sum())
    err = (z.double().cpu() - want).abs().max().item()
    assert err <= 1e-5 * want.abs().max().item(), (err, want.abs().max().item())
