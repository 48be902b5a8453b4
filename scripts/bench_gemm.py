"""Micro-benchmark of the radial-MLP GEMM shapes (CUDA events, inputs > L2)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eqnn_jax_b200.nequip import _mm

E = int(sys.argv[1]) if len(sys.argv) > 1 else 3_200_000
dev = "cuda"
def t(fn, n=int(os.environ.get('REPS', '5'))):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
R = torch.randn(E, 64, device=dev); Z = torch.randn(E, 128, device=dev); Z2 = torch.empty(E, 128, device=dev)
W1 = torch.randn(64, 128, device=dev); W2 = torch.randn(128, 128, device=dev); W4 = torch.randn(128, 11, device=dev)
wt = torch.empty(11, E, device=dev); dW = torch.empty(128, 128, device=dev); dW4 = torch.empty(128, 11, device=dev)
for tc in ((True,) if os.environ.get('TC_ONLY') else (False, True)):
    r = {}
    r["fwd 64->128"] = t(lambda: _mm(E, 128, 64, .1, R, 0, 64, W1, 0, 128, Z2, 0, 128, tf32x3=tc))
    r["fwd 128->128 (gelu pro)"] = t(lambda: _mm(E, 128, 128, .1, Z, 0, 128, W2, 0, 128, Z2, 0, 128, pro=1, pro_scale=1.5, tf32x3=tc))
    r["fwd 128->11 colmajor"] = t(lambda: _mm(E, 11, 128, .1, Z, 0, 128, W4, 0, 11, wt, 0, E, tc=True, pro=1, pro_scale=1.5, tf32x3=tc))
    r["dgrad 128<-128 (gelu' epi)"] = t(lambda: _mm(E, 128, 128, .1, Z, 0, 128, W2, 0, 128, Z2, 0, 128, tb=True, epi=2, aux=Z, ld_aux=128, epi_scale=1.5, tf32x3=tc))
    r["dgrad 128<-11"] = t(lambda: _mm(E, 128, 11, .1, wt, 0, E, W4, 0, 11, Z2, 0, 128, ta=True, tb=True, epi=2, aux=Z, ld_aux=128, epi_scale=1.5, tf32x3=tc))
    r["wgrad 128x128"] = t(lambda: _mm(128, 128, E, .1, Z, 0, 128, Z2, 0, 128, dW, 0, 128, ta=True, pro=1, pro_scale=1.5, tf32x3=tc))
    r["wgrad 128x11"] = t(lambda: _mm(128, 11, E, .1, Z, 0, 128, wt, 0, E, dW4, 0, 11, ta=True, tb=True, pro=1, pro_scale=1.5, tf32x3=tc))
    r["wgrad 64x128"] = t(lambda: _mm(64, 128, E, .1, R, 0, 64, Z2, 0, 128, dW, 0, 128, ta=True, tf32x3=tc))
    print("tensor cores" if tc else "cuda cores", {k: round(v, 3) for k, v in r.items()})
a = torch.empty(E * 128, device=dev); b = torch.empty_like(a)
print("copy 2x%.1f GB: %.3f ms" % (a.numel() * 4 / 1e9, t(lambda: b.copy_(a))))
