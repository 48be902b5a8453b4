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
//   one persistent CTA per SM, 800 threads, warp-specialised:
//     warps 0..7   (epilogue)  tcgen05.ld the fp32 accumulator (double-buffered in TMEM), apply alpha / bias /
//                              activation-derivative epilogue in packed fp32x2 math, transpose through XOR-swizzled
//                              shared memory so that every global store covers whole 128-byte lines;
//     warp  8      (MMA)       one elected thread issues 3 MMAs per 8-wide k-step, commits to mbarriers;
//     warps 9..24  (producers) stream 128-row A chunks (32 k-columns) from HBM with 128-bit loads three chunks
//                              ahead (register ring) plus an L2 bulk prefetch of the next tile, apply the activation
//                              prologue and the hi/lo split (packed fp32x2), store into a ring of 128B-swizzled
//                              K-major shared-memory stages, one elected mbarrier arrival per warp.
//   B (the <=128x128 weight matrix) is split once per CTA and stays resident in shared memory.
// tc_wgrad:  dW = alpha * pro(X)^T @ G with the edge axis as K: a loader warp issues TMA bulk copies
//   (cp.async.bulk + mbarrier complete_tx) of whole 32-edge chunks into a raw ring; transform warps split them into
//   MN-major SWIZZLE_128B_BASE32B operands; registers are re-balanced between roles with setmaxnreg.
#include "common.cuh"
#include "tc_common.cuh"
#include "rmlp_common.cuh"
#include <type_traits>

namespace {

using namespace tc;

constexpr int EPI_WARPS = 8, PROD_WARPS = 8;   // epilogue warp w: TMEM lane quarter w%4, column half w/4
constexpr int MMA_WARP = EPI_WARPS;                          // warp 8
constexpr int TC_THREADS = (EPI_WARPS + 1 + PROD_WARPS) * 32; // 544
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

// ---- row-GEMM thread layout: 8 epilogue warps, 1 MMA warp, 16 producer warps ----
constexpr int RG_PROD_WARPS = 16;
constexpr int RG_PT = RG_PROD_WARPS * 32;                          // 512 producer threads
constexpr int RG_THREADS = (EPI_WARPS + 1 + RG_PROD_WARPS) * 32;   // 800
#ifndef RG_L2_PREFETCH
#define RG_L2_PREFETCH 1
#endif
constexpr uint32_t STG_WARP = 32 * 128;                            // epilogue staging: 32 rows x 128 B per warp

template <int K, int N, bool STAGED = true>
struct RowCfg {
  static constexpr int NCH = (K + 31) / 32;
  static constexpr int KSTEPS = (K < 32 ? K : 32) / 8;
  static constexpr uint32_t B_ATOM = N * 128;                 // one 32-wide k slab of B: N rows x 128 B
  static constexpr uint32_t B_HALF = NCH * B_ATOM;
  static constexpr uint32_t B_BYTES = 2 * B_HALF;
  // wide outputs are transposed through shared memory so that global stores cover whole 128-byte lines
  static constexpr uint32_t STG_BYTES = (N >= 64 && STAGED) ? EPI_WARPS * STG_WARP : 0;
  static constexpr uint32_t FIXED = B_BYTES + STG_BYTES + 1024 /*align*/ + 256 /*barriers*/;
  static constexpr int NSTAGE = (FIXED + 3 * STAGE_BYTES <= SMEM_LIMIT) ? 3 : 2;   // A-operand ring depth
  static constexpr uint32_t SMEM = FIXED + NSTAGE * STAGE_BYTES;
  static constexpr uint32_t ACC_STRIDE = N < 32 ? 32 : N;     // TMEM columns between the two accumulators
  static constexpr uint32_t TMEM_COLS = (2 * ACC_STRIDE <= 64) ? 64 : (2 * ACC_STRIDE <= 128 ? 128 : 256);
  static_assert(SMEM <= SMEM_LIMIT, "shared memory budget exceeded");
};

// AMODE 0: A row-major (k contiguous), K % 32 == 0.  AMODE 1: A column-major (m contiguous), K <= 32.
// PRO 1: A := gelu_tanh(A) * pro_scale.  EPI 1: C += bias (aux[n]).  EPI 2: C *= gelu_tanh'(aux) * epi_scale.
// EPI 3: C = (alpha*acc + C_old) * gelu_tanh'(aux) * epi_scale  (two gradient branches meeting at one activation).
// CMODE 0: C row-major [M][ldc]; CMODE 1: C column-major [n][ldc] (only the first n_valid columns).
template <int K, int N, int AMODE, int PRO, int EPI, int CMODE>
__global__ void __launch_bounds__(RG_THREADS, 1) tc_rowgemm_kernel(const RowGemmP p) {
  using Cfg = RowCfg<K, N, AMODE == 0>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t sB_hi = base, sB_lo = base + Cfg::B_HALF;
  const uint32_t sA = base + Cfg::B_BYTES;
  const uint32_t sStg = sA + Cfg::NSTAGE * STAGE_BYTES;
  const uint32_t bars = sStg + Cfg::STG_BYTES;
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
      mbar_init(a_full(s), RG_PROD_WARPS);
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
  for (int idx = tid; idx < Cfg::NCH * 32 * N; idx += RG_THREADS) {
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
    // three chunks of 128-bit loads are in flight per thread (register ring): the stream is bound by
    // HBM latency x bytes in flight (512 threads x 3 x 32 B = 48 KB per SM), not by the split arithmetic.
    const int pt = tid - (MMA_WARP + 1) * 32;
    const uint32_t my_tiles = (blockIdx.x < n_tiles) ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;
    const uint32_t total = my_tiles * Cfg::NCH;
    // this thread's two 16-byte pieces of every chunk: fixed (row, 16-byte column) and swizzled smem offset
    int prow[2], pc4[2];
    uint32_t poff[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int f = pt + i * RG_PT;
      prow[i] = (AMODE == 0) ? (f >> 3) : (f & 127);
      pc4[i] = (AMODE == 0) ? (f & 7) : (f >> 7);
      poff[i] = sw128(prow[i], pc4[i]);
    }
    const int64_t tile_rows = (int64_t)gridDim.x * TILE_M;
    const int64_t m_last = p.M - 1;
    auto issue = [&](uint32_t g, float4 (&v)[2]) {
      if (g >= total) return;
      const int64_t m0 = (int64_t)blockIdx.x * TILE_M + (int64_t)(g / Cfg::NCH) * tile_rows;
      const int c = g % Cfg::NCH;
      if (RG_L2_PREFETCH && AMODE == 0 && c == 0 && pt == 0) {
        // The k-chunks of a tile visit every 512-byte row four times, microseconds apart: to DRAM that is one
        // row activation per 128 bytes.  Pull the NEXT tile (and its aux rows) into L2 as whole contiguous blocks.
        const int64_t mn = m0 + tile_rows;
        if (mn < p.M) {
          const int64_t rows = (p.M - mn < TILE_M) ? (p.M - mn) : TILE_M;
          if (p.lda == K) l2_prefetch_bulk(p.A + mn * p.lda, (uint32_t)(rows * K * 4));
          if ((EPI == 2 || EPI == 3) && p.ld_aux == N) l2_prefetch_bulk(p.aux + mn * p.ld_aux, (uint32_t)(rows * N * 4));
        }
      }
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        // rows past M (last tile only) re-read row M-1: their accumulator rows are never stored
        const int64_t m = (m0 + prow[i] < m_last) ? (m0 + prow[i]) : m_last;
        if (AMODE == 0) {
          v[i] = __ldg(reinterpret_cast<const float4*>(p.A + m * p.lda + 32 * c + 4 * pc4[i]));
        } else {
          float t[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (4 * pc4[i] + q < p.k_valid) t[q] = __ldg(p.A + (int64_t)(4 * pc4[i] + q) * p.lda + m);
          v[i] = make_float4(t[0], t[1], t[2], t[3]);
        }
      }
    };
    uint32_t cs = 0, cpar = 1;                       // ring position of the next chunk to publish
    auto consume = [&](float4 (&v)[2]) {
      mbar_wait(a_empty(cs), cpar);
      uint8_t* st_hi = gen_base + (sA - base) + cs * STAGE_BYTES;
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        float4 hi, lo;
        act_split4<PRO == 1>(v[i], p.pro_scale, hi, lo);
        *reinterpret_cast<float4*>(st_hi + poff[i]) = hi;
        *reinterpret_cast<float4*>(st_hi + A_HALF + poff[i]) = lo;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(a_full(cs));        // one arrival per producer warp
      if (++cs == Cfg::NSTAGE) {
        cs = 0;
        cpar ^= 1u;
      }
    };
    float4 v0[2], v1[2], v2[2], v3[2];
    issue(0, v0);
    issue(1, v1);
    issue(2, v2);
    for (uint32_t g = 0; g < total; g += 4) {
      issue(g + 3, v3);
      consume(v0);
      if (g + 1 < total) { issue(g + 4, v0); consume(v1); }
      if (g + 2 < total) { issue(g + 5, v1); consume(v2); }
      if (g + 3 < total) { issue(g + 6, v2); consume(v3); }
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
    // warp w reads TMEM lanes (w%4)*32.. (hardware rule) and the column half w/4 of the accumulator.
    // tcgen05.ld hands every thread one ROW; rows are read (aux) and written (C) with 256-bit accesses so
    // that each lane moves whole 32-byte sectors.
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;
    uint32_t tcount = 0;
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++tcount) {
      const uint32_t ab = tcount & 1u;
      const int64_t m = tile * TILE_M + row;
      const bool live = m < p.M;
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ab * Cfg::ACC_STRIDE;
      if (N >= 64) {
        constexpr int NB16 = (N / 2) / 16;               // 16-column batches per warp
        const int col0 = half * (N / 2);
        uint8_t* stg = gen_base + (sStg - base) + warp * STG_WARP;
        float ax[16], an[16];
        const float* arow = ((EPI == 2 || EPI == 3) && live) ? p.aux + m * p.ld_aux + col0 : nullptr;
        if ((EPI == 2 || EPI == 3) && live) {
          ldg256(arow, ax);
          ldg256(arow + 8, ax + 8);
        }
        mbar_wait(t_full(ab), (tcount >> 1) & 1u);
        tc_fence_after();
#pragma unroll
        for (int cb = 0; cb < NB16; ++cb) {
          float v[16];
          tmem_ld16(taddr + col0 + cb * 16, v);
          if ((EPI == 2 || EPI == 3) && live && cb + 1 < NB16) {
            ldg256(arow + (cb + 1) * 16, an);
            ldg256(arow + (cb + 1) * 16 + 8, an + 8);
          }
          tmem_ld_wait();
          if (cb == NB16 - 1) {
            tc_fence_before();
            mbar_arrive(t_empty(ab));
          }
          if (live) {
            float o[16], cold[16];
            if (EPI == 3) {
              ldg256(p.C + m * p.ldc + col0 + cb * 16, cold);
              ldg256(p.C + m * p.ldc + col0 + cb * 16 + 8, cold + 8);
            }
            const f32x2 al2 = pk2(p.alpha, p.alpha), es2 = pk2(p.epi_scale, p.epi_scale);
#pragma unroll
            for (int j = 0; j < 16; j += 2) {
              f32x2 r = mul2(pk2(v[j], v[j + 1]), al2);
              if (EPI == 3) r = add2(r, pk2(cold[j], cold[j + 1]));            // second gradient branch already in C
              if (EPI == 1)                                                      // bias[n] (models/mlp.py Dense)
                r = add2(r, pk2(__ldg(p.aux + col0 + cb * 16 + j), __ldg(p.aux + col0 + cb * 16 + j + 1)));
              if (EPI == 2 || EPI == 3) r = mul2(r, mul2(gelu_grad2(pk2(ax[j], ax[j + 1])), es2));
              upk2(r, o[j], o[j + 1]);
            }
            if (AMODE != 0) {
              float* crow = p.C + m * p.ldc + col0 + cb * 16;
              stg256(crow, o);
              stg256(crow + 8, o + 8);
            }
            // stage: lane = row, 16-byte chunk index XOR-swizzled by the row (conflict-free both ways)
#pragma unroll
            for (int j = 0; AMODE == 0 && j < 4; ++j)
              *reinterpret_cast<float4*>(stg + lane * 128 + ((((cb & 1) * 4 + j) ^ (lane & 7)) << 4)) =
                  make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
          }
          if (AMODE == 0 && (cb & 1)) {
            // 32 staged columns: each store instruction writes four whole 128-byte row segments
            __syncwarp();
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int r = i * 4 + (lane >> 3), ch = lane & 7;
              const float4 t = *reinterpret_cast<const float4*>(stg + r * 128 + ((ch ^ (r & 7)) << 4));
              const int64_t mr = tile * TILE_M + quarter * 32 + r;
              if (mr < p.M) *reinterpret_cast<float4*>(p.C + mr * p.ldc + col0 + (cb >> 1) * 32 + ch * 4) = t;
            }
            __syncwarp();
          }
          if ((EPI == 2 || EPI == 3) && cb + 1 < NB16) {
#pragma unroll
            for (int j = 0; j < 16; ++j) ax[j] = an[j];
          }
        }
      } else {
        // narrow output (N = 16 or 32): the lower column half stores; the upper half only keeps the protocol
        mbar_wait(t_full(ab), (tcount >> 1) & 1u);
        tc_fence_after();
        float v[N];
        if (half == 0) {
          if (N == 32) tmem_ld32(taddr, v); else tmem_ld16(taddr, v);
          tmem_ld_wait();
        }
        tc_fence_before();
        mbar_arrive(t_empty(ab));
        if (half == 0 && live) {
          if (CMODE == 0 && N == 32) {
            float o[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) o[j] = p.alpha * v[j];
            float* crow = p.C + m * p.ldc;
#pragma unroll
            for (int j = 0; j < 32; j += 8) stg256(crow + j, o + j);
          } else {
#pragma unroll
            for (int j = 0; j < N; ++j) {
              if (j < p.n_valid) {
                if (CMODE == 1) p.C[(int64_t)j * p.ldc + m] = p.alpha * v[j];
                else p.C[m * p.ldc + j] = p.alpha * v[j];
              }
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
  using Cfg = RowCfg<K, N, AMODE == 0>;
  auto kern = tc_rowgemm_kernel<K, N, AMODE, PRO, EPI, CMODE>;
  static DeviceOnce once;           // one per kernel instantiation; the attribute is per device
  EQNN_CUDA(configure_smem(once, kern, (int)Cfg::SMEM));
  const int sms = sm_count();
  const int64_t n_tiles = (p.M + TILE_M - 1) / TILE_M;
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  kern<<<grid, RG_THREADS, Cfg::SMEM, st>>>(p);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

// ------------------------------------------------------------------------------------------------
// tc_wgrad:  dW[Mi x N] = alpha * sum_e pro(X[e, :Mi])^T (x) G[e, :N]      (weight gradients, K = #edges)
//   X = saved pre-activations (row-major, Mi in {64,128} contiguous), G = upstream gradient, either
//   row-major [E][128] (MN-major B operand) or column-major [N<=16][E] (K-major B operand).
//   Each CTA reduces a contiguous slice of edges into one TMEM accumulator and writes a partial
//   [Mi x N]; a second kernel sums the per-CTA partials in a fixed order (deterministic).
// ------------------------------------------------------------------------------------------------
struct WgradP {
  const float* X;
  int64_t ldx;
  const float* G;
  int64_t ldg;
  int n_valid;
  int64_t E;
  float pro_scale;
  float* partial;   // [gridDim.x][MI][NPAD]
  float* colsum;    // [gridDim.x][N] per-CTA column sums of G (the bias gradient of a Dense layer), or NULL; GMODE 0
  // "blocked" operands (written by the fused radial-MLP chain kernel, rmlp.cu): every 32-edge chunk is one contiguous
  // block laid out [16-byte column group c4][row][4 floats] instead of [row][column], so that the producer's lanes
  // (one row each) write 512 contiguous bytes per store; the buffer is padded to whole chunks.
  int x_blocked, g_blocked;
  // PRO == 2: X is not read at all -- it is the Bessel basis of the edge length (e3nn.bessel, the first radial layer's
  // input, models/nequip.py:57), generated by the transform warps from the 32 edge records (r_hat, d) of a chunk:
  // 512 bytes per chunk cross HBM instead of 32 * MI * 4.  geom = [E][4] floats, d in component 3.
  const float* geom;
  float r_cut, pref;
  const int32_t* n_dev;   // live edge count on the device (<= E), or NULL: shared radial evaluations (pairs.cu)
};

// wgrad thread layout (warpgroup-aligned so that registers can be re-balanced with setmaxnreg):
//   warps 0..7 drain (TMEM -> fp32 register accumulators, 64 columns each: register-hungry),
//   warps 8..19 transform (raw ring -> split operands: latency-bound, few registers), warp 20 MMA, warp 21 loader
constexpr int WG_TW = 12;                                     // transform warps
constexpr int WG_PT = WG_TW * 32;                             // 384 transform threads
constexpr int WG_TW0 = EPI_WARPS;                             // first transform warp (8)
constexpr int WG_MMA_WARP = WG_TW0 + WG_TW;                   // warp 20
constexpr int WG_LOADER_WARP = WG_MMA_WARP + 1;               // warp 21
constexpr int WG_THREADS = (WG_LOADER_WARP + 1) * 32;         // 704 threads, 80 registers each at launch
constexpr int WG_REGS_DRAIN = 112, WG_REGS_TRANSFORM = 56;    // launch: 80/thread; 8*(112-80) <= 12*(80-56)

template <int MI, int N, int GMODE>
struct WgradCfg {
  static constexpr uint32_t A_HALF_W = 4 * 4096;              // 4 m-blocks x (4 k-groups x 1 KB): M padded to 128
  static constexpr uint32_t B_HALF_W = (GMODE == 0) ? (N / 32) * 4096 : N * 128;
  static constexpr uint32_t STAGE_W = 2 * A_HALF_W + 2 * B_HALF_W;
  static constexpr int NSTAGE = 2;                            // split (hi, lo) operand stages
  static constexpr uint32_t XRAW = 32 * MI * 4;               // one 32-edge chunk of X, as it lies in HBM
  static constexpr uint32_t GRAW = (GMODE == 0) ? 32 * N * 4 : 16 * 128;
  static constexpr uint32_t RAW_STAGE = XRAW + GRAW;
  static constexpr int RAW_NS = (NSTAGE * STAGE_W + 4 * RAW_STAGE + 2560 <= SMEM_LIMIT) ? 4 : 3;   // raw ring depth
  static constexpr uint32_t SMEM = NSTAGE * STAGE_W + RAW_NS * RAW_STAGE + 1024 + 256 + 512 /* bessel frequencies */;
  static_assert(SMEM <= SMEM_LIMIT, "shared memory budget exceeded");
};

template <int MI, int N, int GMODE, int PRO>   // GMODE 0: G row-major [E][N]; 1: G column-major [n][E], N = 16; PRO 0 none, 1 gelu, 2 bessel(d)
__global__ void __launch_bounds__(WG_THREADS, 1) tc_wgrad_kernel(const WgradP p) {
  using Cfg = WgradCfg<MI, N, GMODE>;
  constexpr uint32_t A_HALF_W = Cfg::A_HALF_W, B_HALF_W = Cfg::B_HALF_W, STAGE_W = Cfg::STAGE_W;
  constexpr int NSTAGE = Cfg::NSTAGE, RAW_NS = Cfg::RAW_NS;
  // The tensor core's fp32 accumulation is not round-to-nearest: its error grows linearly with the number of
  // accumulated MMAs.  The accumulator is therefore drained every SEG chunks (128 edges) into fp32 registers of
  // the epilogue warps (IEEE adds) while the MMAs continue into the second TMEM accumulator.
  constexpr uint32_t SEG = 4;
  constexpr uint32_t ACC_STRIDE = N < 32 ? 32 : N;
  constexpr uint32_t TMEM_COLS = 2 * ACC_STRIDE;
  constexpr int EPI_ACTIVE = (N >= 64) ? EPI_WARPS : 4;     // warps that drain (column halves only when N >= 64)
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t raw0 = base + NSTAGE * STAGE_W;
  const uint32_t bars = raw0 + RAW_NS * Cfg::RAW_STAGE;
  auto a_full = [&](int s) { return bars + 8u * s; };
  auto a_empty = [&](int s) { return bars + 8u * (NSTAGE + s); };
  auto t_full = [&](int b) { return bars + 8u * (2 * NSTAGE + b); };
  auto t_empty = [&](int b) { return bars + 8u * (2 * NSTAGE + 2 + b); };
  auto r_full = [&](int s) { return bars + 8u * (2 * NSTAGE + 4 + s); };
  auto r_empty = [&](int s) { return bars + 8u * (2 * NSTAGE + 4 + RAW_NS + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * NSTAGE + 4 + 2 * RAW_NS);
  uint8_t* gen_base = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  float* wfreq = reinterpret_cast<float*>(gen_base + (bars - base) + 256);   // PRO == 2: j * pi / r_cut, j = 1..MI
  if (PRO == 2 && tid < MI) wfreq[tid] = __fdiv_rn(__fmul_rn((float)(tid + 1), 3.14159265358979323846f), p.r_cut);

  const int64_t E = p.n_dev ? min((int64_t)__ldg(p.n_dev), p.E) : p.E;   // live edges
  const int64_t n_chunks = (E + 31) / 32;
  const int64_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const int64_t c_beg = (int64_t)blockIdx.x * per;
  const int64_t c_end = (c_beg + per < n_chunks) ? c_beg + per : n_chunks;
  const uint32_t total = c_end > c_beg ? (uint32_t)(c_end - c_beg) : 0u;

  if (tid == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(a_full(s), WG_TW);
      mbar_init(a_empty(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(t_full(b), 1);
      mbar_init(t_empty(b), EPI_ACTIVE * 32);
    }
    for (int s = 0; s < RAW_NS; ++s) {
      mbar_init(r_full(s), 1);
      mbar_init(r_empty(s), WG_TW);
    }
    mbar_fence_init();
  }
  if (warp == WG_MMA_WARP) tmem_alloc(tmem_slot, TMEM_COLS);
  if (MI < 128) {   // rows MI..127 of the (padded) A operand are never written by the producers: zero them once
    for (uint32_t o = tid * 16; o < NSTAGE * STAGE_W; o += WG_THREADS * 16)
      *reinterpret_cast<float4*>(gen_base + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen_base + (tmem_slot - base));

  if (warp == WG_LOADER_WARP) {
    // =========================== loader: bulk copies HBM -> raw ring ===========================
    // A 32-edge chunk of X (and of a row-major G) is one contiguous block: one bulk copy each, up to RAW_NS chunks
    // (128 KB) in flight per SM with no registers involved.
    const bool x_dense = (p.ldx == MI), g_dense = (GMODE == 0 && p.ldg == N);
    for (uint32_t g = 0; g < total; ++g) {
      const int rs = g % RAW_NS;
      if (lane == 0) mbar_wait(r_empty(rs), ((g / RAW_NS) & 1u) ^ 1u);
      __syncwarp();
      const int64_t e0 = (c_beg + g) * 32;
      const uint32_t rows = (uint32_t)((E - e0 < 32) ? (E - e0) : 32);
      const uint32_t xdst = raw0 + rs * Cfg::RAW_STAGE, gdst = xdst + Cfg::XRAW;
      // column-major G: 32 edges of a column are one copy, whose size must be a multiple of 16 bytes -- a live count that
      // is not a multiple of 4 (device-side counts are arbitrary) copies up to 3 floats past it, inside the padded leading
      // dimension (ldg % 4 == 0); the transform warps mask them
      const uint32_t rows4 = (rows + 3u) & ~3u;
      const uint32_t gbytes = (GMODE == 0) ? rows * N * 4u : (uint32_t)p.n_valid * rows4 * 4u;
      const uint32_t xbytes = PRO == 2 ? rows * 16u : (p.x_blocked ? 32u * MI * 4u : rows * MI * 4u);
      const uint32_t gbytes_c = (GMODE == 0 && p.g_blocked) ? 32u * N * 4u : gbytes;
      if (lane == 0) mbar_expect_tx(r_full(rs), xbytes + gbytes_c);
      __syncwarp();
      if (PRO == 2) {
        if (lane == 0) bulk_g2s(xdst, p.geom + e0 * 4, xbytes, r_full(rs));     // 32 edge records (r_hat, d)
      } else if (x_dense) {
        if (lane == 0) bulk_g2s(xdst, p.X + e0 * p.ldx, xbytes, r_full(rs));
      } else if ((uint32_t)lane < rows) {
        bulk_g2s(xdst + lane * MI * 4u, p.X + (e0 + lane) * p.ldx, MI * 4u, r_full(rs));
      }
      if (GMODE == 0) {
        if (g_dense) {
          if (lane == 0) bulk_g2s(gdst, p.G + e0 * p.ldg, gbytes_c, r_full(rs));
        } else if ((uint32_t)lane < rows) {
          bulk_g2s(gdst + lane * N * 4u, p.G + (e0 + lane) * p.ldg, N * 4u, r_full(rs));
        }
      } else if (lane < p.n_valid) {
        bulk_g2s(gdst + lane * 128u, p.G + (int64_t)lane * p.ldg + e0, rows4 * 4u, r_full(rs));   // 32 edges of column n
      }
    }
  } else if (warp >= WG_TW0 && warp < WG_MMA_WARP) {
    // =========================== transform: raw ring -> split (hi, lo) operand stages ===========================
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(WG_REGS_TRANSFORM));
    const int pt = tid - WG_TW0 * 32;
    constexpr int XI = (8 * MI + WG_PT - 1) / WG_PT;                       // float4 of X per thread per chunk
    constexpr int GI = (GMODE == 0) ? (8 * N + WG_PT - 1) / WG_PT : 1;     // float4 of G per thread per chunk
    auto put = [&](uint32_t off_hi, uint32_t off_lo, const float4& v, auto act) {
      float4 hi, lo;
      act_split4<decltype(act)::value>(v, p.pro_scale, hi, lo);
      *reinterpret_cast<float4*>(gen_base + off_hi) = hi;
      *reinterpret_cast<float4*>(gen_base + off_lo) = lo;
    };
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    // Column sums of G ride along: thread pt always meets the same four columns (WG_PT is a multiple of N / 4), so
    // they accumulate in one register quadruple -- the Dense bias gradient costs no second pass over G.
    static_assert(GMODE != 0 || WG_PT % (N / 4) == 0, "a transform thread must keep its G columns across chunks");
    const bool want_cs = (GMODE == 0) && p.colsum != nullptr;
    float4 cs = zero4;
    for (uint32_t g = 0; g < total; ++g) {
      const int s = g % NSTAGE, rs = g % RAW_NS;
      const int64_t e0 = (c_beg + g) * 32;
      const int rows = (int)((E - e0 < 32) ? (E - e0) : 32);
      mbar_wait(r_full(rs), (g / RAW_NS) & 1u);
      mbar_wait(a_empty(s), ((g / NSTAGE) & 1u) ^ 1u);
      const uint8_t* xraw = gen_base + (raw0 - base) + rs * Cfg::RAW_STAGE;
      const uint8_t* graw = xraw + Cfg::XRAW;
      const uint32_t sa = s * STAGE_W, sb = sa + 2 * A_HALF_W;
#pragma unroll
      for (int i = 0; i < XI; ++i) {
        const int f = pt + i * WG_PT;
        if (f < 8 * MI) {
          const int row = (p.x_blocked || PRO == 2) ? (f & 31) : f / (MI / 4), c4 = (p.x_blocked || PRO == 2) ? (f >> 5) : f % (MI / 4);
          float4 v = zero4;
          if (PRO == 2) {
            // basis functions 4 c4 + 1 .. 4 c4 + 4 of edge `row`: the forward kernel's arithmetic (rmlp::bessel16)
            if (row < rows) {
              const float d = *reinterpret_cast<const float*>(xraw + (uint32_t)row * 16u + 12u);
              const float inv_d = (d == 0.f) ? 0.f : __fdiv_rn(1.0f, d);
              const float w0 = wfreq[4 * c4], w1 = wfreq[4 * c4 + 1], w2 = wfreq[4 * c4 + 2], w3 = wfreq[4 * c4 + 3];
              const f32x2 d2 = pk2(d, d), sc = pk2(p.pref * inv_d, p.pref * inv_d);
              float s0, s1, s2, s3;
              rmlp::sin_reduced2(mul2(pk2(w0, w1), d2), s0, s1);
              rmlp::sin_reduced2(mul2(pk2(w2, w3), d2), s2, s3);
              upk2(mul2(pk2(s0, s1), sc), v.x, v.y);
              upk2(mul2(pk2(s2, s3), sc), v.z, v.w);
              if (d == 0.f) v = make_float4(__fmul_rn(p.pref, w0), __fmul_rn(p.pref, w1), __fmul_rn(p.pref, w2), __fmul_rn(p.pref, w3));
            }
          } else if (row < rows) {
            v = *reinterpret_cast<const float4*>(xraw + (uint32_t)f * 16u);
          }
          // MN-major slab [m-block of 32][32 k-rows]: 4-row atoms, 32-byte-chunk swizzle
          const uint32_t off = (c4 >> 3) * 4096 + sw128_base32(row, c4 & 7);
          put(sa + off, sa + A_HALF_W + off, v, std::integral_constant<bool, PRO == 1>{});
        }
      }
      if (GMODE == 0) {
#pragma unroll
        for (int i = 0; i < GI; ++i) {
          const int f = pt + i * WG_PT;
          if (f < 8 * N) {
            const int row = p.g_blocked ? (f & 31) : f / (N / 4), c4 = p.g_blocked ? (f >> 5) : f % (N / 4);
            const float4 v = (row < rows) ? *reinterpret_cast<const float4*>(graw + (uint32_t)f * 16u) : zero4;
            const uint32_t off = (c4 >> 3) * 4096 + sw128_base32(row, c4 & 7);
            put(sb + off, sb + B_HALF_W + off, v, std::false_type{});
            if (want_cs) {
              cs.x += v.x; cs.y += v.y; cs.z += v.z; cs.w += v.w;
            }
          }
        }
      } else if (pt < 128) {
        const int n = pt >> 3, c4 = pt & 7;             // K-major: row n, 32 edges per 128-byte row
        float4 v = zero4;
        if (n < p.n_valid) {
          v = *reinterpret_cast<const float4*>(graw + n * 128 + c4 * 16);
          if (4 * c4 + 0 >= rows) v.x = 0.f;
          if (4 * c4 + 1 >= rows) v.y = 0.f;
          if (4 * c4 + 2 >= rows) v.z = 0.f;
          if (4 * c4 + 3 >= rows) v.w = 0.f;
        }
        const uint32_t off = sw128(n, c4);
        put(sb + off, sb + B_HALF_W + off, v, std::false_type{});
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) {
        mbar_arrive(a_full(s));       // operands published to the MMA warp
        mbar_arrive(r_empty(rs));     // raw slot may be refilled
      }
    }
    if (GMODE == 0 && want_cs) {
      // the raw ring is dead once every transform warp has left the loop: reduce the WG_PT / (N / 4) partial sums
      // of every column through it, in a fixed order
      constexpr int NQ = N / 4, NG = WG_PT / NQ;
      float* scratch = reinterpret_cast<float*>(gen_base + (raw0 - base));
      asm volatile("bar.sync 1, %0;" ::"n"(WG_PT) : "memory");
      *reinterpret_cast<float4*>(scratch + (pt / NQ) * N + 4 * (pt % NQ)) = cs;
      asm volatile("bar.sync 1, %0;" ::"n"(WG_PT) : "memory");
      if (pt < N) {
        float t = 0.f;
#pragma unroll 4
        for (int q = 0; q < NG; ++q) t += scratch[q * N + pt];
        p.colsum[(size_t)blockIdx.x * N + pt] = t;
      }
    }
  } else if (warp == WG_MMA_WARP) {
    if (lane == 0 && total > 0) {
      constexpr uint32_t idesc = make_idesc_tf32(128, N, 1, GMODE == 0 ? 1 : 0);
      for (uint32_t g = 0; g < total; ++g) {
        const int s = g % NSTAGE;
        const uint32_t sgm = g / SEG, ab = sgm & 1u;
        if (g % SEG == 0) {
          mbar_wait(t_empty(ab), ((sgm >> 1) & 1u) ^ 1u);
          tc_fence_after();
        }
        mbar_wait(a_full(s), (g / NSTAGE) & 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + ab * ACC_STRIDE;
        const uint32_t a_hi = base + s * STAGE_W, a_lo = a_hi + A_HALF_W;
        const uint32_t b_hi = a_hi + 2 * A_HALF_W, b_lo = b_hi + B_HALF_W;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {               // 8 edges per MMA
          // MN-major: LBO = bytes between 32-wide M/N blocks, SBO = bytes between 4-row K atoms
          const uint64_t dah = make_desc(a_hi + kk * 1024, 4096, 512, 1), dal = make_desc(a_lo + kk * 1024, 4096, 512, 1);
          uint64_t dbh, dbl;
          if (GMODE == 0) {
            dbh = make_desc(b_hi + kk * 1024, 4096, 512, 1);
            dbl = make_desc(b_lo + kk * 1024, 4096, 512, 1);
          } else {
            dbh = make_desc(b_hi + kk * 32, 16, 1024);
            dbl = make_desc(b_lo + kk * 32, 16, 1024);
          }
          mma_tf32(d_tmem, dah, dbh, idesc, ((g % SEG) | kk) ? 1u : 0u);
          mma_tf32(d_tmem, dal, dbh, idesc, 1u);
          mma_tf32(d_tmem, dah, dbl, idesc, 1u);
        }
        mma_commit(a_empty(s));
        if (g % SEG == SEG - 1 || g == total - 1) mma_commit(t_full(ab));
      }
    }
    __syncwarp();
  } else if (warp < EPI_WARPS) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(WG_REGS_DRAIN));
    if (warp < EPI_ACTIVE) {
    // ============== epilogue: drain TMEM segments into fp32 registers, then partial[cta] ==============
    const int quarter = warp & 3, half = warp >> 2;
    const int row = quarter * 32 + lane;               // = input-feature index i
    constexpr int CW = (N >= 64) ? N / 2 : N;          // columns owned by this warp
    const int col0 = half * CW;
    float acc[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) acc[j] = 0.f;
    const uint32_t n_seg = (total + SEG - 1) / SEG;
    for (uint32_t sgm = 0; sgm < n_seg; ++sgm) {
      const uint32_t ab = sgm & 1u;
      mbar_wait(t_full(ab), (sgm >> 1) & 1u);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + ab * ACC_STRIDE + col0;
#pragma unroll
      for (int cb = 0; cb < CW / 16; ++cb) {
        float v[16];
        tmem_ld16(taddr + cb * 16, v);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[cb * 16 + j] += v[j];
      }
      tc_fence_before();
      mbar_arrive(t_empty(ab));
    }
    if (row < MI) {
      float* out = p.partial + ((size_t)blockIdx.x * MI + row) * N + col0;
#pragma unroll
      for (int j = 0; j < CW; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == WG_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// out = alpha * sum over the per-CTA partials.  32 outputs x 8 part-groups per block: every group walks its share of
// the partials (coalesced across the 32 outputs), the 8 group sums are added in a fixed order (deterministic).
__global__ void __launch_bounds__(256) wgrad_reduce_kernel(const float* __restrict__ partial, int n_parts, int MI,
                                                           int NPAD, int n_valid, float alpha, float* __restrict__ out,
                                                           int64_t ldo) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  const bool ok = i < MI * n_valid;
  const int r = ok ? i / n_valid : 0, c = ok ? i - r * n_valid : 0;
  float s = 0.f;
  if (ok)
    for (int q = grp; q < n_parts; q += 8) s += partial[((size_t)q * MI + r) * NPAD + c];
  red[grp][lane] = s;
  __syncthreads();
  if (grp == 0 && ok) {
    float t = red[0][lane];
#pragma unroll
    for (int g = 1; g < 8; ++g) t += red[g][lane];
    out[(int64_t)r * ldo + c] = alpha * t;
  }
}

// y[n] = sum over the per-CTA column sums, CTA order (deterministic)
__global__ void colsum_parts_reduce_kernel(const float* __restrict__ parts, int n_parts, int N, float* __restrict__ y) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  float t = 0.f;
  for (int q = 0; q < n_parts; ++q) t += parts[(size_t)q * N + n];
  y[n] = t;
}

template <int MI, int N, int GMODE, int PRO>
int launch_wgrad(const WgradP& p, int grid, cudaStream_t st) {
  constexpr uint32_t SMEM = WgradCfg<MI, N, GMODE>::SMEM;
  auto kern = tc_wgrad_kernel<MI, N, GMODE, PRO>;
  static DeviceOnce once;
  EQNN_CUDA(configure_smem(once, kern, (int)SMEM));
  kern<<<grid, WG_THREADS, SMEM, st>>>(p);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

}  // namespace

size_t eqnn_tc_wgrad_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (K < 4096 || M > 128 || N > 128) return 0;
  const int sms = sm_count();
  return (size_t)sms * (128 * 128 + 128) * sizeof(float);      // weight partials + column-sum partials
}

// dW[M x N] = alpha * pro(X)^T @ G with the edge dimension as K.  Same return convention as below.
// colsum_out (optional, row-major G with N = 128 only): y[n] = sum_k G[k, n], accumulated by the same pass over G.
int eqnn_tc_wgrad_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                      const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                      int pro, float pro_scale, int epi, void* ws, size_t ws_bytes, cudaStream_t st,
                      float* colsum_out, int blocked, const int32_t* n_dev) {
  // blocked: bit 0 = A (X) in the blocked chunk layout, bit 1 = B (G); see WgradP
  if (accumulate || epi != 0 || K < 4096 || scn != 1) return 0;
  if (colsum_out && !(N == 128 && sbn == 1 && !(blocked & 2))) return 0;
  if (!(M == 32 || M == 64 || M == 128) || sam != 1 || sak % 4 != 0 || (reinterpret_cast<uintptr_t>(A) & 15)) return 0;
  if (reinterpret_cast<uintptr_t>(B) & 15) return 0;
  const int sms = sm_count();
  const int64_t n_chunks = (K + 31) / 32;
  const int grid = (int)(n_chunks < sms ? n_chunks : sms);
  WgradP p;
  p.X = A; p.ldx = sak; p.G = B; p.E = K; p.pro_scale = pro_scale; p.n_valid = (int)N;
  p.x_blocked = blocked & 1; p.g_blocked = (blocked >> 1) & 1;
  p.geom = nullptr; p.r_cut = 1.f; p.pref = 0.f; p.n_dev = n_dev;
  if (p.x_blocked && sak != M) return 0;
  p.partial = reinterpret_cast<float*>(ws);
  p.colsum = nullptr;
  int rc = 0, npad = 0;
  if (N == 128 && sbn == 1 && sbk % 4 == 0 && (!p.g_blocked || sbk == 128)) {
    npad = 128; p.ldg = sbk;
    if (!ws || ws_bytes < (size_t)grid * (M * npad + (colsum_out ? npad : 0)) * 4) return 0;
    if (colsum_out) p.colsum = p.partial + (size_t)grid * M * npad;
    if (M == 128) rc = pro ? launch_wgrad<128, 128, 0, 1>(p, grid, st) : launch_wgrad<128, 128, 0, 0>(p, grid, st);
    else if (M == 64) rc = pro ? launch_wgrad<64, 128, 0, 1>(p, grid, st) : launch_wgrad<64, 128, 0, 0>(p, grid, st);
    else rc = pro ? launch_wgrad<32, 128, 0, 1>(p, grid, st) : launch_wgrad<32, 128, 0, 0>(p, grid, st);
  } else if (N <= 16 && sbk == 1 && sbn % 4 == 0 && !p.g_blocked) {
    npad = 16; p.ldg = sbn;
    if (!ws || ws_bytes < (size_t)grid * M * npad * 4) return 0;
    if (M == 128) rc = pro ? launch_wgrad<128, 16, 1, 1>(p, grid, st) : launch_wgrad<128, 16, 1, 0>(p, grid, st);
    else if (M == 64) rc = pro ? launch_wgrad<64, 16, 1, 1>(p, grid, st) : launch_wgrad<64, 16, 1, 0>(p, grid, st);
    else return 0;
  } else {
    return 0;
  }
  if (rc) return -rc;
  const int tot = (int)(M * N);
  wgrad_reduce_kernel<<<(tot + 31) / 32, 256, 0, st>>>(p.partial, grid, (int)M, npad, (int)N, alpha, C, scm);
  eqnn_count_launches(1);
  if (p.colsum) {
    colsum_parts_reduce_kernel<<<1, 128, 0, st>>>(p.colsum, grid, (int)N, colsum_out);
    eqnn_count_launches(1);
  }
  return 1;
}

// First radial layer's weight gradient with the Bessel basis generated in the kernel (PRO = 2):
//   dW[n_rb x 128] = alpha * bessel(d)^T @ G,  d = geom[e][3], G = [E][128] row-major or blocked.
// Same return convention as eqnn_tc_wgrad_try.
int eqnn_tc_wgrad_bessel_try(int n_rb, int64_t E, float alpha, const float* geom, double r_cut, const float* G, int g_blocked,
                             float* C, int64_t scm, void* ws, size_t ws_bytes, cudaStream_t st, const int32_t* n_dev) {
  if (!(n_rb == 32 || n_rb == 64) || E < 4096 || !geom || (reinterpret_cast<uintptr_t>(geom) & 15) ||
      (reinterpret_cast<uintptr_t>(G) & 15))
    return 0;
  const int sms = sm_count();
  const int64_t n_chunks = (E + 31) / 32;
  const int grid = (int)(n_chunks < sms ? n_chunks : sms);
  if (!ws || ws_bytes < (size_t)grid * n_rb * 128 * 4) return 0;
  WgradP p;
  p.X = nullptr; p.ldx = n_rb; p.G = G; p.ldg = 128; p.E = E; p.pro_scale = 1.f; p.n_valid = 128;
  p.x_blocked = 0; p.g_blocked = g_blocked ? 1 : 0;
  p.partial = reinterpret_cast<float*>(ws);
  p.colsum = nullptr;
  p.geom = geom; p.r_cut = (float)r_cut; p.pref = (float)sqrt(2.0 / r_cut); p.n_dev = n_dev;
  const int rc = n_rb == 64 ? launch_wgrad<64, 128, 0, 2>(p, grid, st) : launch_wgrad<32, 128, 0, 2>(p, grid, st);
  if (rc) return -rc;
  wgrad_reduce_kernel<<<(n_rb * 128 + 31) / 32, 256, 0, st>>>(p.partial, grid, n_rb, 128, 128, alpha, C, scm);
  eqnn_count_launches(1);
  return 1;
}

// Returns 1 if the problem was handled on the tensor cores, 0 if the caller must use the SIMT
// path, negative EQNN_ERR_* on a launch error.
int eqnn_tc_rowgemm_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                        const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                        int pro, float pro_scale, int epi, const float* aux, int64_t ld_aux, float epi_scale,
                        cudaStream_t st) {
  if (accumulate || M < 4096) return 0;
  auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
  RowGemmP p;
  p.A = A; p.B = B; p.sbk = sbk; p.sbn = sbn; p.k_valid = (int)K; p.n_valid = (int)N;
  p.C = C; p.aux = aux; p.ld_aux = ld_aux; p.M = M;
  p.alpha = alpha; p.pro_scale = pro_scale; p.epi_scale = epi_scale;
  int rc = 0;
  const bool a_row = (sak == 1 && sam % 4 == 0 && al16(A));
  auto al32 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 31) == 0; };
  const bool c_row = (scn == 1 && scm % 8 == 0 && al32(C));          // 256-bit row stores
  const bool aux_ok = (epi != 2 && epi != 3) || (ld_aux % 8 == 0 && al32(aux));  // 256-bit row loads
  if (N == 128 && a_row && c_row && aux_ok && (K == 128 || K == 64 || K == 32)) {
    p.lda = sam; p.ldc = scm;
#define EQNN_TC_CASE(KK, PR, EP)                                                              \
  if (K == KK && pro == PR && epi == EP) { rc = launch_rowgemm<KK, 128, 0, PR, EP, 0>(p, st); return rc ? -rc : 1; }
    EQNN_TC_CASE(64, 0, 0) EQNN_TC_CASE(128, 1, 0) EQNN_TC_CASE(128, 0, 2) EQNN_TC_CASE(128, 0, 0)
    EQNN_TC_CASE(64, 1, 0) EQNN_TC_CASE(128, 1, 2) EQNN_TC_CASE(64, 0, 2)
    EQNN_TC_CASE(128, 0, 3) EQNN_TC_CASE(32, 0, 0) EQNN_TC_CASE(32, 0, 1) EQNN_TC_CASE(64, 0, 1) EQNN_TC_CASE(128, 0, 1) EQNN_TC_CASE(128, 1, 1)
#undef EQNN_TC_CASE
    return 0;
  }
  if (N <= 16 && K == 128 && a_row && scm == 1 && epi == 0) {      // narrow last layer -> column-major output
    p.lda = sam; p.ldc = scn;
    rc = (pro == 1) ? launch_rowgemm<128, 16, 0, 1, 0, 1>(p, st) : launch_rowgemm<128, 16, 0, 0, 0, 1>(p, st);
    return rc ? -rc : 1;
  }
  if (N == 32 && K == 128 && a_row && c_row && scm == 32 && pro == 0 && epi == 0) {   // data-gradient w.r.t. 32 radial features
    p.lda = sam; p.ldc = scm;
    rc = launch_rowgemm<128, 32, 0, 0, 0, 0>(p, st);
    return rc ? -rc : 1;
  }
  if (N == 128 && K <= 16 && sam == 1 && c_row && aux_ok && pro == 0 && epi != 1) {   // data-gradient of the last layer
    p.lda = sak; p.ldc = scm;
    rc = (epi == 2) ? launch_rowgemm<16, 128, 1, 0, 2, 0>(p, st) : launch_rowgemm<16, 128, 1, 0, 0, 0>(p, st);
    return rc ? -rc : 1;
  }
  return 0;
}
