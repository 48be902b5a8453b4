import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eqnn_jax_b200.nequip import _mm
torch.manual_seed(0)
for E, N, Mi in [(4096, 11, 128), (4096 * 148, 11, 128), (3200000, 11, 128), (70000, 128, 128), (3200000, 128, 128), (3200000, 128, 64)]:
    Z = torch.randn(E, Mi).cuda()
    if N == 11:
        G = torch.randn(11, E).cuda(); want = Z.double().t() @ G.double().t()
        C = torch.full((Mi, 11), float("nan")).cuda()
        _mm(Mi, 11, E, 1.0, Z, 0, Mi, G, 0, E, C, 0, 11, ta=True, tb=True, tf32x3=True)
    else:
        G = torch.randn(E, 128).cuda(); want = Z.double().t() @ G.double()
        C = torch.full((Mi, 128), float("nan")).cuda()
        _mm(Mi, 128, E, 1.0, Z, 0, Mi, G, 0, 128, C, 0, 128, ta=True, tf32x3=True)
    torch.cuda.synchronize()
    d = (C.double() - want).abs()
    print(E, N, Mi, "max|got|", float(C.abs().max()), "max|want|", float(want.abs().max()), "maxerr", float(d.max()),
          "nan", int(torch.isnan(C).sum()), "zero frac", float((C == 0).float().mean()))
    if N == 128:
        # which (i, j) blocks are right?
        ok = (d < 1e-3 * want.abs().max())
        print("  ok fraction by 32x32 block:\n", ok.float().reshape(Mi // 32, 32, 4, 32).mean((1, 3)).cpu().numpy().round(2))
