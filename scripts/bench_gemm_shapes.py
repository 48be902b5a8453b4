"""Times eqnn_gemm_f32 on the SEGNN irrep-blocked GEMM shapes (forward, dgrad, wgrad), aligned and unaligned strides.
   python scripts/bench_gemm_shapes.py  -> one line per shape: ms, TFLOP/s."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from eqnn_jax_b200.nequip import _mm

n = 65536
TC = bool(int(os.environ.get("TC", "0")))
dev = "cuda"
buf = lambda *s: torch.randn(*s, device=dev)


def run(tag, M, N, K, A, a_off, lda, B, b_off, ldb, C, c_off, ldc, **kw):
    for _ in range(3):
        _mm(M, N, K, 1.0, A, a_off, lda, B, b_off, ldb, C, c_off, ldc, tc_any=TC, **kw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        _mm(M, N, K, 1.0, A, a_off, lda, B, b_off, ldb, C, c_off, ldc, tc_any=TC, **kw)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag:44s} M {M:6d} N {N:4d} K {K:6d}  {ms:8.3f} ms  {2.0 * M * N * K / ms * 1e-9:7.2f} TFLOP/s", flush=True)


ONLY_DENSE = bool(int(os.environ.get("ONLY_DENSE", "0")))
for (ldx, DT, c0, tb, woff, label) in ([] if ONLY_DENSE else [(341, 630, 89, 124, 3, "unaligned"), (344, 632, 88, 124, 0, "aligned")]):
    x, T, W = buf(n, ldx), buf(n, DT), buf(200000)
    dx, gW = buf(n, ldx), buf(200000)
    for (K, Nr) in [(89 if label == "unaligned" else 88, 124), (42, 396), (40, 110)]:
        run(f"fwd   {label} K{K} N{Nr}", n, Nr, K, x, c0, ldx, W, woff, Nr, T, tb, DT)
        run(f"fwd+  {label} K{K} N{Nr}", n, Nr, K, x, c0, ldx, W, woff, Nr, T, tb, DT, accumulate=True)
        run(f"dgrad {label} K{K} N{Nr}", n, K, Nr, T, tb, DT, W, woff, Nr, dx, c0, ldx, tb=True)
        run(f"wgrad {label} K{K} N{Nr}", K, Nr, n, x, c0, ldx, T, tb, DT, gW, woff, Nr, ta=True)
# dense lowering of the same layer (first edge layer K = 341, hidden layers K = 170)
x, Xb, M = buf(n, 341), buf(n, 768), buf(341 * 768)
run("dense fwd", n, 768, 341, x, 0, 341, M, 0, 768, Xb, 0, 768)
run("dense dgrad", n, 341, 768, Xb, 0, 768, M, 0, 768, x, 0, 341, tb=True)
run("dense wgrad", 341, 768, n, x, 0, 341, Xb, 0, 768, M, 0, 768, ta=True)
x2 = buf(n, 170)
run("dense170 fwd", n, 768, 170, x2, 0, 170, M, 0, 768, Xb, 0, 768)
run("dense170 dgrad", n, 170, 768, Xb, 0, 768, M, 0, 768, x2, 0, 170, tb=True)
run("dense170 wgrad", 170, 768, n, x2, 0, 170, Xb, 0, 768, M, 0, 768, ta=True)
