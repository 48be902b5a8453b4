// Radial-MLP GEMMs on the 5th-generation tensor cores (tcgen05 + TMEM), fp32-class accuracy.
//
// The radial MLP of models/nequip.py:62-66 is >95% of the interaction block's FLOPs, with
// M = #edges (millions) and K, N <= 128.  The reference computes it in true fp32, and the
// parity contract is 1e-5, so a single TF32/BF16 pass (~1e-3) is not admissible.  Every
// operand is therefore split on the fly into (hi, lo) TF32 halves and each product is issued
// as three tcgen05.mma.kind::tf32 instructions (hi*hi + lo*hi + hi*lo, error ~2^-21),
// accumulating in fp32 in tensor memory.
//
// tc_rowgemm:  C[M x N] = epi( alpha * pro(A)[M x K] @ B[K x N] ),  M tiled by 128 rows.
//   one persistent CTA per SM, warp-specialised:
//     warps 5..12 (producers)  stream 128-row A chunks (32 k-columns) from HBM with 128-bit loads,
//                              apply the activation prologue, split hi/lo, store into a ring of
//                              128B-swizzled K-major shared-memory stages;
//     warp  4     (MMA)        one elected thread issues 3 MMAs per 8-wide k-step, commits to mbarriers;
//     warps 0..3  (epilogue)   tcgen05.ld the fp32 accumulator (double-buffered in TMEM),
//                              apply alpha / activation-derivative epilogue, 128-bit stores.
//   B (the <=128x128 weight matrix) is split once per CTA and stays resident in shared memory.
#include "common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int EPI_WARPS = 4, PROD_WARPS = 8;
constexpr int MMA_WARP = EPI_WARPS;                          // warp 4
constexpr int TC_THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32; // 416
constexpr int PT = PROD_WARPS * 32;                          // 256 producer threads
constexpr int TILE_M = 128;
constexpr uint32_t A_HALF = TILE_M * 128;                    // 16 KB: 128 rows x 128 B (32 tf32)
constexpr uint32_t STAGE_BYTES = 2 * A_HALF;                 // hi + lo
constexpr uint32_t SMEM_LIMIT = 227 * 1024;

struct RowGemmP {
  const float* A;
  int64_t lda;
  const float* B;
  int64_t sbk, sbn;
  int k_valid, n_valid;
  float* C;
  int64_t ldc;
  const float* aux;
  int64_t ld_aux;
  int64_t M;
  float alpha, pro_scale, epi_scale;
};

template <int K, int N>
struct RowCfg {
  static constexpr int NCH = (K + 31) / 32;
  static constexpr int KSTEPS = (K < 32 ? K : 32) / 8;
  static constexpr uint32_t B_ATOM = N * 128;                 // one 32-wide k slab of B: N rows x 128 B
  static constexpr uint32_t B_HALF = NCH * B_ATOM;
  static constexpr uint32_t B_BYTES = 2 * B_HALF;
  static constexpr int NSTAGE = (B_BYTES + 3 * STAGE_BYTES + 2048 <= SMEM_LIMIT) ? 3 : 2;
  static constexpr uint32_t SMEM = B_BYTES + NSTAGE * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr uint32_t ACC_STRIDE = N < 32 ? 32 : N;     // TMEM columns between the two accumulators
  static constexpr uint32_t TMEM_COLS = (2 * ACC_STRIDE <= 64) ? 64 : (2 * ACC_STRIDE <= 128 ? 128 : 256);
};

// AMODE 0: A row-major (k contiguous), K % 32 == 0.  AMODE 1: A column-major (m contiguous), K <= 32.
// PRO 1: A := gelu_tanh(A) * pro_scale.  EPI 2: C *= gelu_tanh'(aux) * epi_scale.
// CMODE 0: C row-major [M][ldc]; CMODE 1: C column-major [n][ldc] (only the first n_valid columns).
template <int K, int N, int AMODE, int PRO, int EPI, int CMODE>
__global__ void __launch_bounds__(TC_THREADS, 1) tc_rowgemm_kernel(const RowGemmP p) {
  using Cfg = RowCfg<K, N>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB_hi = base, sB_lo = base + Cfg::B_HALF;
  const uint32_t sA = base + Cfg::B_BYTES;
  const uint32_t bars = sA + Cfg::NSTAGE * STAGE_BYTES;
  // barrier slots (8 bytes each)
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (Cfg::NSTAGE + s); };
  auto t_full = [&](int b) { return bars + 8u * (2 * Cfg::NSTAGE + b); };
  auto t_empty = [&](int b) { return bars + 8u * (2 * Cfg::NSTAGE + 2 + b); };
  const uint32_t tmem_slot = bars + 8u * (2 * Cfg::NSTAGE + 4);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;

  // ---- one-time setup ----
  if (tid == 0) {
    for (int s = 0; s < Cfg::NSTAGE; ++s) {
      mbar_init(a_full(s), PT);
      mbar_init(a_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full(b), 1);
      mbar_init(t_empty(b), EPI_WARPS * 32);
    }
    mbar_fence_init();
  }
  if (warp == MMA_WARP) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  // resident B: element (n, k) <- B[k*sbk + n*sbn], split hi/lo, K-major 128B-swizzled slabs
  for (int idx = tid; idx < Cfg::NCH * 32 * N; idx += TC_THREADS) {
    const int k = idx % (Cfg::NCH * 32), n = idx / (Cfg::NCH * 32);
    float v = 0.f;
    if (k < p.k_valid && n < p.n_valid) v = p.B[(int64_t)k * p.sbk + (int64_t)n * p.sbn];
    float hi, lo;
    split_tf32(v, hi, lo);
    const uint32_t off = (k >> 5) * Cfg::B_ATOM + sw128(n, (k & 31) >> 2) + (k & 3) * 4;
    *reinterpret_cast<float*>(gen_base + off) = hi;
    *reinterpret_cast<float*>(gen_base + Cfg::B_HALF + off) = lo;
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp >= MMA_WARP + 1) {
    // =========================== producers ===========================
    const int pt = tid - (MMA_WARP + 1) * 32;
    uint32_t it = 0;   // global chunk counter -> ring stage / parity
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t m0 = tile * TILE_M;
      for (int c = 0; c < Cfg::NCH; ++c, ++it) {
        const int s = it % Cfg::NSTAGE;
        const uint32_t par = (it / Cfg::NSTAGE) & 1u;
        float4 v[4];
        if (AMODE == 0) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = pt + i * PT, row = f >> 3, c4 = f & 7;
            const int64_t m = m0 + row;
            v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (m < p.M) v[i] = __ldg(reinterpret_cast<const float4*>(p.A + m * p.lda + 32 * c + 4 * c4));
          }
        } else {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int f = pt + i * PT, row = f & 127, c4 = f >> 7;   // c4 in 0..7
            const int64_t m = m0 + row;
            float t[4] = {0.f, 0.f, 0.f, 0.f};
            if (m < p.M) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                if (4 * c4 + q < p.k_valid) t[q] = __ldg(p.A + (int64_t)(4 * c4 + q) * p.lda + m);
            }
            v[i] = make_float4(t[0], t[1], t[2], t[3]);
          }
        }
        mbar_wait(a_empty(s), par ^ 1u);
        const uint32_t st_hi = sA + s * STAGE_BYTES - base, st_lo = st_hi + A_HALF;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int f = pt + i * PT;
          const int row = (AMODE == 0) ? (f >> 3) : (f & 127), c4 = (AMODE == 0) ? (f & 7) : (f >> 7);
          float x[4] = {v[i].x, v[i].y, v[i].z, v[i].w}, hi[4], lo[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            if (PRO == 1) x[q] = gelu_tanh(x[q]) * p.pro_scale;
            split_tf32(x[q], hi[q], lo[q]);
          }
          const uint32_t off = sw128(row, c4);
          *reinterpret_cast<float4*>(gen_base + st_hi + off) = make_float4(hi[0], hi[1], hi[2], hi[3]);
          *reinterpret_cast<float4*>(gen_base + st_lo + off) = make_float4(lo[0], lo[1], lo[2], lo[3]);
        }
        fence_proxy_async_smem();
        mbar_arrive(a_full(s));
      }
    }
  } else if (warp == MMA_WARP) {
    // =========================== MMA issuer ===========================
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(TILE_M, N, 0, 0);
      uint32_t it = 0, tcount = 0;
      for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
        const uint32_t ab = tcount & 1u;
        mbar_wait(t_empty(ab), ((tcount >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * Cfg::ACC_STRIDE;
        for (int c = 0; c < Cfg::NCH; ++c, ++it) {
          const int s = it % Cfg::NSTAGE;
          mbar_wait(a_full(s), (it / Cfg::NSTAGE) & 1u);
          tc_fence_after();
          const uint32_t a_hi = sA + s * STAGE_BYTES, a_lo = a_hi + A_HALF;
          const uint32_t b_hi = sB_hi + c * Cfg::B_ATOM, b_lo = sB_lo + c * Cfg::B_ATOM;
#pragma unroll
          for (int kk = 0; kk < Cfg::KSTEPS; ++kk) {
            const uint64_t dah = make_desc(a_hi + kk * 32, 16, 1024), dal = make_desc(a_lo + kk * 32, 16, 1024);
            const uint64_t dbh = make_desc(b_hi + kk * 32, 16, 1024), dbl = make_desc(b_lo + kk * 32, 16, 1024);
            mma_tf32(d_tmem, dah, dbh, idesc, (c | kk) ? 1u : 0u);
            mma_tf32(d_tmem, dal, dbh, idesc, 1u);
            mma_tf32(d_tmem, dah, dbl, idesc, 1u);
          }
          mma_commit(a_empty(s));       // stage reusable once these MMAs have read it
        }
        mma_commit(t_full(ab));         // accumulator complete
      }
    }
    __syncwarp();
  } else {
    // =========================== epilogue ===========================
    const int row = warp * 32 + lane;
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t ab = tcount & 1u;
      const int64_t m = tile * TILE_M + row;
      mbar_wait(t_full(ab), (tcount >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + ab * Cfg::ACC_STRIDE;
      if (N >= 32) {
#pragma unroll 1
        for (int cb = 0; cb < N / 32; ++cb) {
          float v[32];
          tmem_ld32(taddr + cb * 32, v);
          tmem_ld_wait();
          if (cb == N / 32 - 1) {
            tc_fence_before();
            mbar_arrive(t_empty(ab));
          }
          if (m < p.M) {
            if (CMODE == 0) {
              float* crow = p.C + m * p.ldc + cb * 32;
              const float* arow = (EPI == 2) ? p.aux + m * p.ld_aux + cb * 32 : nullptr;
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                float o[4] = {p.alpha * v[j], p.alpha * v[j + 1], p.alpha * v[j + 2], p.alpha * v[j + 3]};
                if (EPI == 2) {
                  const float4 a = __ldg(reinterpret_cast<const float4*>(arow + j));
                  o[0] *= gelu_tanh_grad(a.x) * p.epi_scale;
                  o[1] *= gelu_tanh_grad(a.y) * p.epi_scale;
                  o[2] *= gelu_tanh_grad(a.z) * p.epi_scale;
                  o[3] *= gelu_tanh_grad(a.w) * p.epi_scale;
                }
                *reinterpret_cast<float4*>(crow + j) = make_float4(o[0], o[1], o[2], o[3]);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (cb * 32 + j < p.n_valid) p.C[(int64_t)(cb * 32 + j) * p.ldc + m] = p.alpha * v[j];
            }
          }
        }
      } else {
        float v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        tc_fence_before();
        mbar_arrive(t_empty(ab));
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            if (j < p.n_valid) {
              if (CMODE == 1) p.C[(int64_t)j * p.ldc + m] = p.alpha * v[j];
              else p.C[m * p.ldc + j] = p.alpha * v[j];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int K, int N, int AMODE, int PRO, int EPI, int CMODE>
int launch_rowgemm(const RowGemmP& p, cudaStream_t st) {
  using Cfg = RowCfg<K, N>;
  auto kern = tc_rowgemm_kernel<K, N, AMODE, PRO, EPI, CMODE>;
  static bool configured = false;   // attribute is per function; setting it again is harmless
  if (!configured) {
    EQNN_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM));
    configured = true;
  }
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  kern<<<grid, TC_THREADS, Cfg::SMEM, st>>>(p);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

}  // namespace

// Returns 1 if the problem was handled on the tensor cores, 0 if the caller must use the SIMT
// path, negative EQNN_ERR_* on a launch error.
int eqnn_tc_rowgemm_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                        const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                        int pro, float pro_scale, int epi, const float* aux, int64_t ld_aux, float epi_scale,
                        cudaStream_t st) {
  if (accumulate || M < 4096 || epi == 1) return 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  RowGemmP p;
  p.A = A; p.B = B; p.sbk = sbk; p.sbn = sbn; p.k_valid = (int)K; p.n_valid = (int)N;
  p.C = C; p.aux = aux; p.ld_aux = ld_aux; p.M = M;
  p.alpha = alpha; p.pro_scale = pro_scale; p.epi_scale = epi_scale;
  int rc = 0;
  const bool a_row = (sak == 1 && sam % 4 == 0 && al16(A));
  const bool c_row = (scn == 1 && scm % 4 == 0 && al16(C));
  const bool aux_ok = (epi == 0) || (ld_aux % 4 == 0 && al16(aux));
  if (N == 128 && a_row && c_row && aux_ok && (K == 128 || K == 64)) {
    p.lda = sam; p.ldc = scm;
#define EQNN_TC_CASE(KK, PR, EP)                                                              \
  if (K == KK && pro == PR && epi == EP) { rc = launch_rowgemm<KK, 128, 0, PR, EP, 0>(p, st); return rc ? -rc : 1; }
    EQNN_TC_CASE(64, 0, 0) EQNN_TC_CASE(128, 1, 0) EQNN_TC_CASE(128, 0, 2) EQNN_TC_CASE(128, 0, 0)
    EQNN_TC_CASE(64, 1, 0) EQNN_TC_CASE(128, 1, 2) EQNN_TC_CASE(64, 0, 2)
#undef EQNN_TC_CASE
    return 0;
  }
  if (N <= 16 && K == 128 && a_row && scm == 1 && epi == 0) {      // last radial layer -> column-major w_t
    p.lda = sam; p.ldc = scn;
    rc = (pro == 1) ? launch_rowgemm<128, 16, 0, 1, 0, 1>(p, st) : launch_rowgemm<128, 16, 0, 0, 0, 1>(p, st);
    return rc ? -rc : 1;
  }
  if (N == 128 && K <= 16 && sam == 1 && c_row && aux_ok && pro == 0) {   // data-gradient of the last layer
    p.lda = sak; p.ldc = scm;
    rc = (epi == 2) ? launch_rowgemm<16, 128, 1, 0, 2, 0>(p, st) : launch_rowgemm<16, 128, 1, 0, 0, 0>(p, st);
    return rc ? -rc : 1;
  }
  return 0;
}
