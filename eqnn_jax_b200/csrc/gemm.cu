// Dense fp32 building block: C (+)= epi(alpha * pro(A) @ B) with arbitrary element
// strides, fused activation prologue / epilogue and deterministic split-K (sm_100a).
//
// Serves (a) the radial MLP of models/nequip.py:62-66 (e3nn MultiLayerPerceptron:
// x @ W / sqrt(fan_in), normalised gelu) forward, data-gradient and weight-gradient,
// (b) every irreps Linear of the interaction block lowered to a dense matrix
// (nequip.py:43,85,162,197) and (c) the read-out MLP (models/mlp.py).  fp32 FMA on the
// CUDA cores: this is the exact-arithmetic path the 1e-5 parity contract is checked on.
#include "common.cuh"

namespace {

struct GemmP {
  int64_t M, N, K;
  float alpha;
  const float* A;
  int64_t sam, sak;
  const float* B;
  int64_t sbk, sbn;
  float* C;
  int64_t scm, scn;
  int accumulate, pro, epi;
  float pro_scale, epi_scale;
  const float* aux;
  int64_t ld_aux;
  int ksplit;
  int64_t k_per_split;
  float* partial;
};

__device__ __forceinline__ float apply_pro(float v, int pro, float s) {
  return pro == 1 ? gelu_tanh(v) * s : v;
}

__device__ __forceinline__ void store_out(const GemmP& p, int64_t m, int64_t n, float v) {
  if (p.epi == 1) v += p.aux[n];
  else if (p.epi == 2) v *= gelu_tanh_grad(p.aux[m * p.ld_aux + n]) * p.epi_scale;
  float* c = p.C + m * p.scm + n * p.scn;
  if (p.epi == 3) v = (v + *c) * gelu_tanh_grad(p.aux[m * p.ld_aux + n]) * p.epi_scale;
  else if (p.accumulate) v += *c;
  *c = v;
}

constexpr int BK = 16;

// AMODE: 0 scalar generic | 1 float4 along K (sak == 1) | 2 float4 along M (sam == 1)
// BMODE: 0 scalar generic | 1 float4 along N (sbn == 1)
template <int BM, int BN, int TM, int TN, int AMODE, int BMODE>
__global__ void __launch_bounds__((BM / TM) * (BN / TN))
gemm_kernel(const GemmP p) {
  constexpr int NT = (BM / TM) * (BN / TN);
  constexpr int LDA_S = BM + 4, LDB_S = BN + 4;
  constexpr int EA = BM * BK / NT, EB = (BN * BK + NT - 1) / NT;
  constexpr int EBV = (BN * BK / 4 + NT - 1) / NT;   // float4s of the B tile per thread
  constexpr int RB = EB > 4 * EBV ? EB : 4 * EBV;
  static_assert(BM * BK % NT == 0 && EA % 4 == 0, "A tile must split evenly into float4s");
  __shared__ __align__(16) float As[2][BK][LDA_S];
  __shared__ __align__(16) float Bs[2][BK][LDB_S];

  const int tid = threadIdx.x;
  const int tx = tid % (BN / TN), ty = tid / (BN / TN);
  const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;
  const int64_t kbeg = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t kend = min(p.K, kbeg + p.k_per_split);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[EA], rb[RB];
  // column j of this thread's TN-wide fragment (two interleaved halves when TN == 8: conflict-free LDS.128)
  auto col_of = [&](int j) { return TN == 8 ? (j < 4 ? tx * 4 + j : BN / 2 + tx * 4 + (j - 4)) : tx * TN + j; };

  auto load_tile = [&](int64_t k0) {
    if (AMODE == 1) {
#pragma unroll
      for (int v = 0; v < EA / 4; ++v) {
        const int idx = tid + v * NT;           // float4 index: (mm, k4)
        const int mm = idx / (BK / 4), k4 = idx % (BK / 4);
        const int64_t m = m0 + mm, k = k0 + 4 * k4;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (m < p.M && k + 3 < kend) t = *reinterpret_cast<const float4*>(p.A + m * p.sam + k);
        else if (m < p.M) {
          float* tp = reinterpret_cast<float*>(&t);
          for (int q = 0; q < 4; ++q) if (k + q < kend) tp[q] = p.A[m * p.sam + k + q];
        }
        ra[4 * v + 0] = apply_pro(t.x, p.pro, p.pro_scale);
        ra[4 * v + 1] = apply_pro(t.y, p.pro, p.pro_scale);
        ra[4 * v + 2] = apply_pro(t.z, p.pro, p.pro_scale);
        ra[4 * v + 3] = apply_pro(t.w, p.pro, p.pro_scale);
      }
    } else if (AMODE == 2) {
#pragma unroll
      for (int v = 0; v < EA / 4; ++v) {
        const int idx = tid + v * NT;           // float4 index: (kk, m4)
        const int kk = idx / (BM / 4), m4 = idx % (BM / 4);
        const int64_t m = m0 + 4 * m4, k = k0 + kk;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (k < kend && m + 3 < p.M) t = *reinterpret_cast<const float4*>(p.A + k * p.sak + m);
        else if (k < kend) {
          float* tp = reinterpret_cast<float*>(&t);
          for (int q = 0; q < 4; ++q) if (m + q < p.M) tp[q] = p.A[k * p.sak + m + q];
        }
        ra[4 * v + 0] = apply_pro(t.x, p.pro, p.pro_scale);
        ra[4 * v + 1] = apply_pro(t.y, p.pro, p.pro_scale);
        ra[4 * v + 2] = apply_pro(t.z, p.pro, p.pro_scale);
        ra[4 * v + 3] = apply_pro(t.w, p.pro, p.pro_scale);
      }
    } else {
      const bool kfast = p.sak <= p.sam;
#pragma unroll
      for (int v = 0; v < EA; ++v) {
        const int idx = tid + v * NT;
        const int mm = kfast ? idx / BK : idx % BM, kk = kfast ? idx % BK : idx / BM;
        const int64_t m = m0 + mm, k = k0 + kk;
        float t = 0.f;
        if (m < p.M && k < kend) t = apply_pro(p.A[m * p.sam + k * p.sak], p.pro, p.pro_scale);
        ra[v] = t;
      }
    }
    if (BMODE == 1) {
#pragma unroll
      for (int v = 0; v < EBV; ++v) {
        const int idx = tid + v * NT;           // float4 index: (kk, n4)
        const int kk = idx / (BN / 4), n4 = idx % (BN / 4);
        const int64_t n = n0 + 4 * n4, k = k0 + kk;
        float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < BN * BK / 4) {
          if (k < kend && n + 3 < p.N) t = *reinterpret_cast<const float4*>(p.B + k * p.sbk + n);
          else if (k < kend) {
            float* tp = reinterpret_cast<float*>(&t);
            for (int q = 0; q < 4; ++q) if (n + q < p.N) tp[q] = p.B[k * p.sbk + n + q];
          }
        }
        rb[4 * v + 0] = t.x;
        rb[4 * v + 1] = t.y;
        rb[4 * v + 2] = t.z;
        rb[4 * v + 3] = t.w;
      }
    } else {
      const bool nfast = p.sbn <= p.sbk;
#pragma unroll
      for (int v = 0; v < EB; ++v) {
        const int idx = tid + v * NT;
        const int kk = nfast ? idx / BN : idx % BK, nn = nfast ? idx % BN : idx / BK;
        const int64_t n = n0 + nn, k = k0 + kk;
        float t = 0.f;
        if (idx < BN * BK && n < p.N && k < kend) t = p.B[k * p.sbk + n * p.sbn];
        rb[v] = t;
      }
    }
  };

  auto store_tile = [&](int buf) {
    if (AMODE == 1) {
#pragma unroll
      for (int v = 0; v < EA / 4; ++v) {
        const int idx = tid + v * NT;
        const int mm = idx / (BK / 4), k4 = idx % (BK / 4);
#pragma unroll
        for (int q = 0; q < 4; ++q) As[buf][4 * k4 + q][mm] = ra[4 * v + q];
      }
    } else if (AMODE == 2) {
#pragma unroll
      for (int v = 0; v < EA / 4; ++v) {
        const int idx = tid + v * NT;
        const int kk = idx / (BM / 4), m4 = idx % (BM / 4);
        *reinterpret_cast<float4*>(&As[buf][kk][4 * m4]) =
            make_float4(ra[4 * v], ra[4 * v + 1], ra[4 * v + 2], ra[4 * v + 3]);
      }
    } else {
      const bool kfast = p.sak <= p.sam;
#pragma unroll
      for (int v = 0; v < EA; ++v) {
        const int idx = tid + v * NT;
        const int mm = kfast ? idx / BK : idx % BM, kk = kfast ? idx % BK : idx / BM;
        As[buf][kk][mm] = ra[v];
      }
    }
    if (BMODE == 1) {
#pragma unroll
      for (int v = 0; v < EBV; ++v) {
        const int idx = tid + v * NT;
        const int kk = idx / (BN / 4), n4 = idx % (BN / 4);
        if (idx < BN * BK / 4)
          *reinterpret_cast<float4*>(&Bs[buf][kk][4 * n4]) =
              make_float4(rb[4 * v], rb[4 * v + 1], rb[4 * v + 2], rb[4 * v + 3]);
      }
    } else {
      const bool nfast = p.sbn <= p.sbk;
#pragma unroll
      for (int v = 0; v < EB; ++v) {
        const int idx = tid + v * NT;
        const int kk = nfast ? idx / BN : idx % BK, nn = nfast ? idx % BN : idx / BK;
        if (idx < BN * BK) Bs[buf][kk][nn] = rb[v];
      }
    }
  };

  int buf = 0;
  if (kbeg < kend) {
    load_tile(kbeg);
    store_tile(0);
  }
  __syncthreads();
  for (int64_t k0 = kbeg; k0 < kend; k0 += BK) {
    const bool more = k0 + BK < kend;
    if (more) load_tile(k0 + BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      float a[TM], b[TN];
#pragma unroll
      for (int i = 0; i < TM; i += 4) {
        const float4 t = *reinterpret_cast<const float4*>(&As[buf][kk][ty * TM + i]);
        a[i] = t.x; a[i + 1] = t.y; a[i + 2] = t.z; a[i + 3] = t.w;
      }
      if (TN % 4 == 0) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const float4 t = *reinterpret_cast<const float4*>(&Bs[buf][kk][col_of(j)]);
          b[j] = t.x; b[j + 1] = t.y; b[j + 2] = t.z; b[j + 3] = t.w;
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) b[j] = Bs[buf][kk][tx * TN + j];
      }
#pragma unroll
      for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) store_tile(buf ^ 1);
    __syncthreads();
    buf ^= 1;
  }

  if (p.ksplit > 1) {
    float* part = p.partial + (size_t)blockIdx.z * p.M * p.N;
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t m = m0 + ty * TM + i;
      if (m >= p.M) continue;
#pragma unroll
      for (int j = 0; j < TN; ++j) {
        const int64_t n = n0 + col_of(j);
        if (n < p.N) part[m * p.N + n] = acc[i][j];
      }
    }
  } else {
    // fast path: row-major C, 16-byte aligned rows, plain or accumulate epilogue -> 128-bit stores
    const bool vec = (TN % 4 == 0) && p.scn == 1 && (p.scm % 4 == 0) && p.epi == 0 &&
                     ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll
    for (int i = 0; i < TM; ++i) {
      const int64_t m = m0 + ty * TM + i;
      if (m >= p.M) continue;
      if (TN % 4 == 0 && vec) {
#pragma unroll
        for (int j = 0; j < TN; j += 4) {
          const int64_t n = n0 + col_of(j);
          if (n + 3 < p.N) {
            float4* c = reinterpret_cast<float4*>(p.C + m * p.scm + n);
            float4 o = make_float4(p.alpha * acc[i][j], p.alpha * acc[i][j + 1], p.alpha * acc[i][j + 2],
                                   p.alpha * acc[i][j + 3]);
            if (p.accumulate) {
              const float4 q = *c;
              o.x += q.x; o.y += q.y; o.z += q.z; o.w += q.w;
            }
            *c = o;
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              if (n + q < p.N) store_out(p, m, n + q, p.alpha * acc[i][j + q]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < TN; ++j) {
          const int64_t n = n0 + col_of(j);
          if (n < p.N) store_out(p, m, n, p.alpha * acc[i][j]);
        }
      }
    }
  }
}

__global__ void splitk_reduce_kernel(const GemmP p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.M * p.N) return;
  float s = 0.f;
  for (int z = 0; z < p.ksplit; ++z) s += p.partial[(size_t)z * p.M * p.N + i];
  store_out(p, i / p.N, i % p.N, p.alpha * s);
}

struct SplitPlan {
  int ksplit;
  int64_t k_per_split;
};


// ------------------------------------------------------------------------------------------------
// Skinny shapes: the node-level Linear layers have a handful of columns (7, 8, 18) and hundreds of thousands of
// rows.  A 128-wide tile wastes >90% of its lanes on them, so they get two dedicated kernels.
//   rows:   C[M x N] (+)= alpha * A[M x K] @ B[K x N],  K, N <= 32: one thread per row, B in shared memory
//   wgrad:  C[Mi x N]  = alpha * sum_k A(m, k) B(k, n), Mi, N <= 32, K = #rows: per-CTA partials over row
//           chunks (operands staged through shared memory), fixed-order reduction (deterministic)
// ------------------------------------------------------------------------------------------------
template <int NB>
__global__ void __launch_bounds__(256) skinny_rows_kernel(const GemmP p) {
  __shared__ float sB[32 * NB];
  const int K = (int)p.K, N = (int)p.N;
  for (int i = threadIdx.x; i < K * NB; i += 256) {
    const int k = i / NB, n = i - k * NB;
    sB[i] = n < N ? p.B[(int64_t)k * p.sbk + (int64_t)n * p.sbn] : 0.f;
  }
  __syncthreads();
  const int64_t r = (int64_t)blockIdx.x * 256 + threadIdx.x;
  if (r >= p.M) return;
  float acc[NB];
#pragma unroll
  for (int n = 0; n < NB; ++n) acc[n] = 0.f;
  const float* a = p.A + r * p.sam;
  for (int k = 0; k < K; ++k) {
    const float v = a[(int64_t)k * p.sak];
#pragma unroll
    for (int n = 0; n < NB; ++n) acc[n] = fmaf(v, sB[k * NB + n], acc[n]);
  }
  float* c = p.C + r * p.scm;
#pragma unroll
  for (int n = 0; n < NB; ++n)
    if (n < N) {
      const float v = p.alpha * acc[n];
      c[(int64_t)n * p.scn] = p.accumulate ? c[(int64_t)n * p.scn] + v : v;
    }
}

constexpr int SKW_ROWS = 64;      // rows staged per step
template <int MB>                  // Mi, N <= MB; thread = (m, n) pair
__global__ void __launch_bounds__(MB * MB) skinny_wgrad_kernel(const GemmP p, int64_t rows_per_cta, float* __restrict__ partial) {
  __shared__ float sA[SKW_ROWS][MB + 1];
  __shared__ float sBt[SKW_ROWS][MB + 1];
  const int Mi = (int)p.M, N = (int)p.N;
  const int tid = threadIdx.x, m = tid / MB, n = tid - m * MB;
  const int64_t k0 = (int64_t)blockIdx.x * rows_per_cta, k1 = min(p.K, k0 + rows_per_cta);
  float acc = 0.f;
  for (int64_t kb = k0; kb < k1; kb += SKW_ROWS) {
    const int nr = (int)min((int64_t)SKW_ROWS, k1 - kb);
    __syncthreads();
    for (int i = tid; i < SKW_ROWS * MB; i += MB * MB) {
      const int r = i / MB, c = i - r * MB;
      sA[r][c] = (r < nr && c < Mi) ? p.A[(int64_t)c * p.sam + (kb + r) * p.sak] : 0.f;
      sBt[r][c] = (r < nr && c < N) ? p.B[(kb + r) * p.sbk + (int64_t)c * p.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll 8
    for (int r = 0; r < SKW_ROWS; ++r) acc = fmaf(sA[r][m], sBt[r][n], acc);
  }
  if (m < Mi && n < N) partial[((size_t)blockIdx.x * Mi + m) * N + n] = acc;
}

// one WARP per output: the lanes split the per-CTA partials (a thread per output walked all of them alone: 36 us for 56
// outputs x 296 partials), fixed lane partition + fixed shuffle tree = deterministic
__global__ void skinny_wgrad_reduce_kernel(const GemmP p, int n_parts, const float* __restrict__ partial) {
  const int lane = threadIdx.x & 31;
  const int i = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int Mi = (int)p.M, N = (int)p.N;
  if (i >= Mi * N) return;
  float s = 0.f;
  for (int q = lane; q < n_parts; q += 32) s += partial[(size_t)q * Mi * N + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane != 0) return;
  const int m = i / N, n = i - m * N;
  float* c = p.C + (int64_t)m * p.scm + (int64_t)n * p.scn;
  *c = p.accumulate ? *c + p.alpha * s : p.alpha * s;
}


//   tiny:   problems with a few rows or a short reduction (the B-row read-out MLPs and their weight gradients):
//           one thread per output element, threads along n so that B(k, n) is read coalesced; prologue and
//           epilogue as in the tiled kernel
__global__ void __launch_bounds__(256) tiny_gemm_kernel(const GemmP p) {
  // one warp per output element: the 32 lanes split the reduction (these problems have few outputs and a
  // reduction of up to a few hundred terms, so a thread per output would be a chain of dependent loads)
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * 256 + threadIdx.x) >> 5;
  if (i >= p.M * p.N) return;
  const int64_t m = i / p.N, n = i - m * p.N;
  const float* a = p.A + m * p.sam;
  const float* b = p.B + n * p.sbn;
  float acc = 0.f;
  for (int64_t k = lane; k < p.K; k += 32)
    acc = fmaf(apply_pro(a[k * p.sak], p.pro, p.pro_scale), b[k * p.sbk], acc);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) store_out(p, m, n, p.alpha * acc);
}

inline bool tiny_ok(int64_t M, int64_t N, int64_t K) {
  return K <= 2048 && (M <= 64 || K <= 64) && M * N <= (int64_t)1 << 17 && M * N * K <= (int64_t)1 << 25;
}

inline bool skinny_rows_ok(int64_t M, int64_t N, int64_t K, int pro, int epi) {
  return pro == 0 && epi == 0 && M >= 4096 && K >= 1 && K <= 32 && N <= 32;
}
inline bool skinny_wgrad_ok(int64_t M, int64_t N, int64_t K, int pro, int epi) {
  return pro == 0 && epi == 0 && K >= 4096 && M <= 32 && N <= 32;
}
inline int skinny_wgrad_parts(int64_t K) {
  const int64_t want = ceil_div64(K, 4 * SKW_ROWS);
  return (int)(want < 1184 ? want : 1184);             // up to 8 small CTAs per SM (296 left most of the machine idle)
}

SplitPlan plan_split(int64_t M, int64_t N, int64_t K, int bm, int bn) {
  const int64_t tiles = ceil_div64(M, bm) * ceil_div64(N, bn);
  SplitPlan s{1, K};
  if (tiles < 148 && K >= 4096) {
    int64_t want = (2 * 148 + tiles - 1) / tiles;
    int64_t maxs = K / 1024;
    int64_t ks = want < maxs ? want : maxs;
    if (ks > 1) {
      int64_t per = ceil_div64(ceil_div64(K, ks), BK) * BK;
      s.k_per_split = per;
      s.ksplit = (int)ceil_div64(K, per);
    }
  }
  return s;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

template <int BM, int BN, int TM, int TN>
void launch(const GemmP& p, int amode, int bmode, cudaStream_t st) {
  dim3 grid((unsigned)ceil_div64(p.M, BM), (unsigned)ceil_div64(p.N, BN), (unsigned)p.ksplit);
  constexpr int NT = (BM / TM) * (BN / TN);
#define EQNN_GEMM_CASE(AM, BMD) \
  if (amode == AM && bmode == BMD) { gemm_kernel<BM, BN, TM, TN, AM, BMD><<<grid, NT, 0, st>>>(p); return; }
  EQNN_GEMM_CASE(0, 0) EQNN_GEMM_CASE(0, 1) EQNN_GEMM_CASE(1, 0) EQNN_GEMM_CASE(1, 1) EQNN_GEMM_CASE(2, 0)
  EQNN_GEMM_CASE(2, 1)
#undef EQNN_GEMM_CASE
}

}  // namespace

int eqnn_tc_rowgemm_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                        const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                        int pro, float pro_scale, int epi, const float* aux, int64_t ld_aux, float epi_scale,
                        cudaStream_t st);

size_t eqnn_tc_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K);
int eqnn_tc_wgrad_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                      const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                      int pro, float pro_scale, int epi, void* ws, size_t ws_bytes, cudaStream_t st,
                      float* colsum_out = nullptr, int blocked = 0, const int32_t* n_dev = nullptr);

size_t eqnn_tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K);
int eqnn_tc_gemm_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                     const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                     int pro, int epi, void* ws, size_t ws_bytes, cudaStream_t st);

extern "C" size_t eqnn_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const int bn = N <= 16 ? 16 : (N <= 32 ? 32 : 128);
  SplitPlan s = plan_split(M, N, K, 128, bn);
  const size_t simt = s.ksplit > 1 ? (size_t)s.ksplit * M * N * sizeof(float) : 0;
  const size_t tcw = eqnn_tc_wgrad_workspace_bytes(M, N, K);
  const size_t skw = skinny_wgrad_ok(M, N, K, 0, 0) ? (size_t)skinny_wgrad_parts(K) * M * N * sizeof(float) : 0;
  const size_t tcg = eqnn_tc_gemm_workspace_bytes(M, N, K);
  size_t need = simt > tcw ? simt : tcw;
  need = need > tcg ? need : tcg;
  return need > skw ? need : skw;
}

extern "C" int eqnn_gemm_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam,
                             int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm,
                             int64_t scn, int flags, int pro, float pro_scale, int epi, const float* aux,
                             int64_t ld_aux, float epi_scale, void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(M >= 0 && N >= 0 && K >= 0, "gemm: negative dimension");
  if (M == 0 || N == 0) return EQNN_OK;
  EQNN_CHECK_ARG(A && B && C, "gemm: null pointer");
  EQNN_CHECK_ARG(pro == 0 || pro == 1, "gemm: unknown prologue");
  EQNN_CHECK_ARG(epi >= 0 && epi <= 3, "gemm: unknown epilogue");
  EQNN_CHECK_ARG(epi == 0 || aux, "gemm: epilogue needs aux");
  EQNN_CHECK_ARG(ceil_div64(N, 16) <= 65535, "gemm: N too large for grid.y");
  if (flags & EQNN_GEMM_TF32X3) {
    const int r = eqnn_tc_rowgemm_try(M, N, K, alpha, A, sam, sak, B, sbk, sbn, C, scm, scn,
                                      (flags & EQNN_GEMM_ACCUMULATE) ? 1 : 0, pro, pro_scale, epi, aux, ld_aux,
                                      epi_scale, as_stream(stream));
    if (r == 1) return EQNN_OK;
    if (r < 0) return -r;
    const int w = eqnn_tc_wgrad_try(M, N, K, alpha, A, sam, sak, B, sbk, sbn, C, scm, scn,
                                    (flags & EQNN_GEMM_ACCUMULATE) ? 1 : 0, pro, pro_scale, epi, ws, ws_bytes,
                                    as_stream(stream));
    if (w == 1) return EQNN_OK;
    if (w < 0) return -w;
    if (flags & EQNN_GEMM_TF32X3_ANY) {
      const int g = eqnn_tc_gemm_try(M, N, K, alpha, A, sam, sak, B, sbk, sbn, C, scm, scn,
                                     (flags & EQNN_GEMM_ACCUMULATE) ? 1 : 0, pro, epi, ws, ws_bytes, as_stream(stream));
      if (g == 1) return EQNN_OK;
      if (g < 0) return -g;
    }
  }
  GemmP p;
  p.M = M; p.N = N; p.K = K; p.alpha = alpha;
  p.A = A; p.sam = sam; p.sak = sak;
  p.B = B; p.sbk = sbk; p.sbn = sbn;
  p.C = C; p.scm = scm; p.scn = scn;
  p.accumulate = (flags & EQNN_GEMM_ACCUMULATE) ? 1 : 0;
  p.pro = pro; p.pro_scale = pro_scale;
  p.epi = epi; p.aux = aux; p.ld_aux = ld_aux; p.epi_scale = epi_scale;
  cudaStream_t st = as_stream(stream);
  if (tiny_ok(M, N, K)) {
    tiny_gemm_kernel<<<(unsigned)ceil_div64(M * N * 32, 256), 256, 0, st>>>(p);
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  if (skinny_rows_ok(M, N, K, pro, epi)) {
    const unsigned grid = (unsigned)ceil_div64(M, 256);
    if (N <= 8) skinny_rows_kernel<8><<<grid, 256, 0, st>>>(p);
    else if (N <= 16) skinny_rows_kernel<16><<<grid, 256, 0, st>>>(p);
    else skinny_rows_kernel<32><<<grid, 256, 0, st>>>(p);
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  if (skinny_wgrad_ok(M, N, K, pro, epi) && ws && ws_bytes >= (size_t)skinny_wgrad_parts(K) * M * N * sizeof(float)) {
    const int parts = skinny_wgrad_parts(K);
    const int64_t rows_per_cta = ceil_div64(ceil_div64(K, parts), SKW_ROWS) * SKW_ROWS;
    const int used = (int)ceil_div64(K, rows_per_cta);
    float* partial = reinterpret_cast<float*>(ws);
    const int mx = (int)(M > N ? M : N);
    if (mx <= 8) skinny_wgrad_kernel<8><<<used, 64, 0, st>>>(p, rows_per_cta, partial);
    else if (mx <= 16) skinny_wgrad_kernel<16><<<used, 256, 0, st>>>(p, rows_per_cta, partial);
    else skinny_wgrad_kernel<32><<<used, 1024, 0, st>>>(p, rows_per_cta, partial);
    skinny_wgrad_reduce_kernel<<<(unsigned)ceil_div64(M * N * 32, 128), 128, 0, st>>>(p, used, partial);
    EQNN_CHECK_LAUNCH_N(2);
    return EQNN_OK;
  }
  const int bn = N <= 16 ? 16 : (N <= 32 ? 32 : 128);
  SplitPlan s = plan_split(M, N, K, 128, bn);
  p.ksplit = s.ksplit; p.k_per_split = s.k_per_split;
  p.partial = reinterpret_cast<float*>(ws);
  if (s.ksplit > 1)
    EQNN_CHECK_ARG(ws && ws_bytes >= (size_t)s.ksplit * M * N * sizeof(float) ? 1 : 0, "gemm: split-K workspace too small");
  int amode = 0, bmode = 0;
  if (sak == 1 && sam % 4 == 0 && aligned16(A)) amode = 1;
  else if (sam == 1 && sak % 4 == 0 && aligned16(A)) amode = 2;
  if (sbn == 1 && sbk % 4 == 0 && aligned16(B)) bmode = 1;
  // float4 loads along K need every split to start on a multiple of 4 (k_per_split is a multiple of BK)
  if (bn == 16) launch<128, 16, 8, 1>(p, amode, bmode, st);
  else if (bn == 32) launch<128, 32, 8, 2>(p, amode, bmode, st);
  else launch<128, 128, 8, 8>(p, amode, bmode, st);
  if (s.ksplit > 1) splitk_reduce_kernel<<<(unsigned)ceil_div64(M * N, 256), 256, 0, st>>>(p);
  EQNN_CHECK_LAUNCH_N(s.ksplit > 1 ? 2 : 1);
  return EQNN_OK;
}

// ---- parameter gradients of a Dense layer over a tall batch: kernel and bias in one pass over G ----------------
extern "C" size_t eqnn_colsum_workspace_bytes(int N);
extern "C" int eqnn_colsum_tall(const float* x, int64_t M, int N, float* y, void* ws, size_t ws_bytes,
                                eqnn_stream_t stream);

extern "C" size_t eqnn_dense_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  const size_t a = eqnn_gemm_workspace_bytes(M, N, K), b = eqnn_colsum_workspace_bytes((int)N);
  return a > b ? a : b;
}

extern "C" int eqnn_dense_wgrad_f32(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t lda, int pro,
                                    float pro_scale, const float* G, int64_t ldg, float* gW, int64_t ldw, float* gb,
                                    int flags, void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(M >= 1 && N >= 1 && K >= 0 && gW && gb && ldw >= N, "dense_wgrad: bad argument");
  if (K == 0) {                                        // empty batch: both gradients are zero
    EQNN_CUDA(cudaMemset2DAsync(gW, (size_t)ldw * sizeof(float), 0, (size_t)N * sizeof(float), (size_t)M, as_stream(stream)));
    EQNN_CUDA(cudaMemsetAsync(gb, 0, (size_t)N * sizeof(float), as_stream(stream)));
    return EQNN_OK;
  }
  EQNN_CHECK_ARG(A && G && lda >= M && ldg >= N, "dense_wgrad: bad argument");
  EQNN_CHECK_ARG(!(flags & EQNN_GEMM_ACCUMULATE), "dense_wgrad: accumulation is not supported");
  if (flags & EQNN_GEMM_TF32X3) {
    const int w = eqnn_tc_wgrad_try(M, N, K, alpha, A, 1, lda, G, ldg, 1, gW, ldw, 1, 0, pro, pro_scale, 0, ws, ws_bytes,
                                    as_stream(stream), gb);
    if (w == 1) return EQNN_OK;
    if (w < 0) return -w;
  }
  int rc = eqnn_gemm_f32(M, N, K, alpha, A, 1, lda, G, ldg, 1, gW, ldw, 1, flags, pro, pro_scale, 0, nullptr, 0, 1.f, ws,
                         ws_bytes, stream);
  if (rc) return rc;
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(G) & 15) == 0 && ldg == N, "dense_wgrad: G must be dense and 16-byte aligned");
  return eqnn_colsum_tall(G, K, (int)N, gb, ws, ws_bytes, stream);
}
