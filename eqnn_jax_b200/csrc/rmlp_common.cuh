// Device helpers shared by the fused radial-MLP kernels (rmlp.cu): fp16 (hi, lo) operand splitting, tensor-memory
// stores, tcgen05.mma.kind::f16 with the A operand in tensor memory, the in-kernel Bessel basis.
//
// Precision model.  The radial MLP (models/nequip.py:55-68) is fp32 in the reference and the parity contract is
// 1e-5, so one fp16/bf16/tf32 pass (~5e-4) is not admissible.  Every operand x is split into two fp16 numbers,
// hi = rn16(x), lo = rn16(x - hi) (x - hi is exact in fp32), so hi + lo carries 22 significand bits, and each
// product is issued as three MMAs, hi*hi + lo*hi + hi*lo, accumulated in fp32 in tensor memory: the same 2^-21
// error as the 3xTF32 scheme of tc_mlp.cu at half the shared-memory / tensor-memory footprint (4 instead of 8 bytes
// per element) and twice the tensor-pipe rate -- which is what lets all four weight matrices stay resident in one
// SM's shared memory and the activations stay in tensor memory between layers.  fp16 has 5 exponent bits: the
// operands of this MLP are O(1) by construction (e3nn normalises activations and weights), gradients are brought
// to O(1) by one power-of-two scale per launch, and conversions saturate instead of producing infinities.
#pragma once
#include <cuda_fp16.h>

#include "tc_common.cuh"

namespace rmlp {

using namespace tc;

// ---- fp16 (hi, lo) split of a pair; low half-word = first element --------------------------------------------
__device__ __forceinline__ uint32_t cvt_h2(float a0, float a1) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(a1), "f"(a0));
  return r;
}
__device__ __forceinline__ void split_h2(float a0, float a1, uint32_t& hi, uint32_t& lo) {
  hi = cvt_h2(a0, a1);
  const float2 hf = __half22float2(*reinterpret_cast<const __half2*>(&hi));
  float l0, l1;
  upk2(fma2(pk2(hf.x, hf.y), pk2(-1.f, -1.f), pk2(a0, a1)), l0, l1);   // a - hi: exact in fp32
  lo = cvt_h2(l0, l1);
}

// ---- activation: x * sigma(x) with sigma = 1 / (1 + 2^(x (C0 + C1 x^2))) -------------------------------------------
// gelu_tanh (flax.linen.gelu default): C0 = -2 sqrt(2/pi) log2(e), C1 = 0.044715 C0;  silu: C0 = -log2(e), C1 = 0.
// The derivative is s + x s (1 - s) (D0 + D1 x^2) with D0 = -ln(2) C0, D1 = -3 ln(2) C1.
struct ActC {
  float c0, c1, d0, d1;
};
__host__ __device__ inline ActC act_constants(int kind) {
  const float c0 = kind == 0 ? -2.0f * 0.7978845608028654f * 1.4426950408889634f : -1.4426950408889634f;
  const float c1 = kind == 0 ? c0 * 0.044715f : 0.f;
  return ActC{c0, c1, -0.6931471805599453f * c0, -3.0f * 0.6931471805599453f * c1};
}
__device__ __forceinline__ f32x2 act_sigmoid2(f32x2 x, f32x2 xx, const ActC& a) {
  const f32x2 t = mul2(fma2(xx, pk2(a.c1, a.c1), pk2(a.c0, a.c0)), x);
  float t0, t1;
  upk2(t, t0, t1);
  const f32x2 d = add2(pk2(ex2_ftz(t0), ex2_ftz(t1)), pk2(1.f, 1.f));
  float d0, d1;
  upk2(d, d0, d1);
  return pk2(rcp_ftz(d0), rcp_ftz(d1));
}
// x pair -> act(x) and act'(x) from one sigmoid
__device__ __forceinline__ void act_both2(f32x2 x, const ActC& a, f32x2& g, f32x2& gp) {
  const f32x2 xx = mul2(x, x);
  const f32x2 sg = act_sigmoid2(x, xx, a);
  g = mul2(x, sg);
  const f32x2 w = mul2(x, fma2(xx, pk2(a.d1, a.d1), pk2(a.d0, a.d0)));
  const f32x2 q = mul2(sg, fma2(sg, pk2(-1.f, -1.f), pk2(1.f, 1.f)));
  gp = fma2(q, w, sg);
}
__device__ __forceinline__ f32x2 act_grad2(f32x2 x, const ActC& a) {
  f32x2 g, gp;
  act_both2(x, a, g, gp);
  return gp;
}

// ---- tensor memory stores: thread i of the warp writes lane (warp % 4) * 32 + i, N consecutive columns -------
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, "
      "%15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- tcgen05.mma.kind::f16, fp32 accumulate ------------------------------------------------------------------
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (static_cast<uint32_t>(a_mn_major) << 15) | (static_cast<uint32_t>(b_mn_major) << 16) |
         (static_cast<uint32_t>(N >> 3) << 17) | (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[tmem] * B[smem]: A is 128 lanes x (K = 16 halves = 8 columns), lane = row
__device__ __forceinline__ void mma_f16_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(tmem_d), "r"(tmem_a), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}

// ---- weight images -------------------------------------------------------------------------------------------
// A K x N weight matrix (Flax kernel: element (k, n) at W[k*ld + n]) lives in shared memory as fp16 hi and lo
// images of ceil(K/64) slabs; a slab holds 64 k-values: N rows (n) of 128 bytes, 128-byte swizzled.  Read with a
// K-major descriptor it is the B operand of the forward product a @ W (N = n, K = k); read with an MN-major
// descriptor it is the B operand of the data-gradient product g @ W^T (N = k, K = n).
__host__ __device__ constexpr uint32_t w_half_bytes(int K, int N) { return static_cast<uint32_t>((K + 63) / 64) * N * 128u; }

__device__ __forceinline__ void fill_weight_image(uint8_t* img_hi, uint8_t* img_lo, const float* __restrict__ W, int64_t ld,
                                                  int K, int N, int n_valid, float scale, int tid, int nthreads) {
  const int Kp = ((K + 63) / 64) * 64;
  for (int idx = tid; idx < Kp * N; idx += nthreads) {
    const int n = idx % N, k = idx / N;
    float v = 0.f;
    if (k < K && n < n_valid) v = __ldg(W + (int64_t)k * ld + n) * scale;
    const __half h = __float2half_rn(v);
    const __half l = __float2half_rn(v - __half2float(h));
    const uint32_t off = (uint32_t)(k >> 6) * (uint32_t)N * 128u + sw128((uint32_t)n, (uint32_t)(k & 63) >> 3) + (k & 7) * 2;
    *reinterpret_cast<__half*>(img_hi + off) = h;
    *reinterpret_cast<__half*>(img_lo + off) = l;
  }
}

// ---- Bessel basis (e3nn.bessel, models/nequip.py:57) -----------------------------------------------------------
// sin(x) for |x| < 1e5: three-constant Cody-Waite reduction to [-pi/4, pi/4] and degree-7 / degree-8 polynomials
// (absolute error ~1e-7).  The argument is the fp32-rounded phase fl(w_j * d), the reference's rounding point.
__device__ __forceinline__ float sin_reduced(float x) {
  const float t = fmaf(x, 0.636619747f, 12582912.0f);
  const int q = __float_as_int(t);
  const float n = t - 12582912.0f;
  float r = fmaf(n, -1.57079601e+00f, x);
  r = fmaf(n, -3.13916473e-07f, r);
  r = fmaf(n, -5.39030253e-15f, r);
  const float z = r * r;
  float s = fmaf(-1.95152959e-4f, z, 8.33216087e-3f);
  s = fmaf(s, z, -1.66666546e-1f);
  s = fmaf(s * z, r, r);
  float c = fmaf(2.44331571e-5f, z, -1.38873163e-3f);
  c = fmaf(c, z, 4.16666457e-2f);
  c = fmaf(c, z, -0.5f);
  c = fmaf(c, z, 1.0f);
  float v = (q & 1) ? c : s;
  return (q & 2) ? -v : v;
}

// the same for a pair of arguments, on packed fp32x2 arithmetic (the reduction and both polynomials are FMA chains)
__device__ __forceinline__ void sin_reduced2(f32x2 x, float& s0, float& s1) {
  const f32x2 magic = pk2(12582912.0f, 12582912.0f);
  const f32x2 t = fma2(x, pk2(0.636619747f, 0.636619747f), magic);
  float tq0, tq1;
  upk2(t, tq0, tq1);
  const f32x2 n = add2(t, pk2(-12582912.0f, -12582912.0f));
  f32x2 r = fma2(n, pk2(-1.57079601e+00f, -1.57079601e+00f), x);
  r = fma2(n, pk2(-3.13916473e-07f, -3.13916473e-07f), r);
  r = fma2(n, pk2(-5.39030253e-15f, -5.39030253e-15f), r);
  const f32x2 z = mul2(r, r);
  f32x2 s = fma2(pk2(-1.95152959e-4f, -1.95152959e-4f), z, pk2(8.33216087e-3f, 8.33216087e-3f));
  s = fma2(s, z, pk2(-1.66666546e-1f, -1.66666546e-1f));
  s = fma2(mul2(s, z), r, r);
  f32x2 c = fma2(pk2(2.44331571e-5f, 2.44331571e-5f), z, pk2(-1.38873163e-3f, -1.38873163e-3f));
  c = fma2(c, z, pk2(4.16666457e-2f, 4.16666457e-2f));
  c = fma2(c, z, pk2(-0.5f, -0.5f));
  c = fma2(c, z, pk2(1.0f, 1.0f));
  float sa, sb, ca, cb;
  upk2(s, sa, sb);
  upk2(c, ca, cb);
  const int q0 = __float_as_int(tq0), q1 = __float_as_int(tq1);
  const float v0 = (q0 & 1) ? ca : sa, v1 = (q1 & 1) ? cb : sb;
  s0 = (q0 & 2) ? -v0 : v0;
  s1 = (q1 & 2) ? -v1 : v1;
}

// 16 basis functions j0 + 1 .. j0 + 16 of one edge: pref * sin(fl(w_j d)) / d (w_j at d = 0), as fp32 values
__device__ __forceinline__ void bessel16(const float* __restrict__ wfreq, int j0, float d, float pref, float* v) {
  const float inv_d = (d == 0.f) ? 0.f : __fdiv_rn(1.0f, d);
  const f32x2 d2 = pk2(d, d), sc = pk2(pref * inv_d, pref * inv_d);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float w0 = wfreq[j0 + 2 * j], w1 = wfreq[j0 + 2 * j + 1];
    float s0, s1;
    sin_reduced2(mul2(pk2(w0, w1), d2), s0, s1);            // mul.rn: the reference's fp32-rounded phase
    upk2(mul2(pk2(s0, s1), sc), v[2 * j], v[2 * j + 1]);
    if (d == 0.f) {
      v[2 * j] = __fmul_rn(pref, w0);
      v[2 * j + 1] = __fmul_rn(pref, w1);
    }
  }
}

}  // namespace rmlp
