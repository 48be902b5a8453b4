// General fp32 GEMM on the tcgen05 tensor cores with 3xTF32 operand splitting (fp32-class accuracy):
//   C[m, n] (+)= alpha * sum_k A[m, k] B[k, n],   any M, N, K, element strides and alignment.
// Used by the SEGNN dense-lowered TensorProductLinearGate (models/segnn.py:30-64): x @ M_cat (forward),
// dXbig @ M_cat^T (data gradient) and x^T @ dXbig (weight gradient, split over K with a fixed-order reduction).
//
// One CTA computes one 128 x 128 tile of one K split:
//   warp  0     MMA issuer first (one elected thread): per 32-wide k chunk 4 x 3 tcgen05.mma kind::tf32
//               (A_hi B_hi + A_lo B_hi + A_hi B_lo), accumulator in TMEM
//   warps 0-3   epilogue: TMEM -> registers -> (shared-memory transpose for row-major C) -> global
//   warps 4-19  producers (4-11: A tile, 12-19: B tile): strided, predicated loads (zero fill past M / N / K) held
//               in registers two chunks ahead, hi/lo split (round to nearest TF32), 16-byte stores into
//               128-byte-swizzled K-major operand tiles; threads run along whichever index is contiguous in
//               global memory
// Ring of NSTAGE operand stages with full/empty mbarriers; nothing is assumed about alignment, so the same kernel
// serves row-major, transposed and offset operands.
#include "common.cuh"
#include "tc_common.cuh"

namespace {
using namespace tc;

constexpr int GT_M = 128;
// 20 warps: registers are allocated in groups of four warps, so a 21st warp would cost every thread 16 registers
constexpr int GT_EPI_WARPS = 4, GT_MMA_WARP = 0, GT_PROD_WARPS = 16;
constexpr int GT_THREADS = (GT_EPI_WARPS + GT_PROD_WARPS) * 32;       // 640

// fp32 -> (hi, lo): hi = a rounded to the nearest TF32 (half away from zero, two integer instructions; cvt.rna.tf32
// is emulated in ~5), lo = a - hi exact, |lo| <= 2^-11 |a|.  Rounding instead of truncating halves |lo|, which
// quarters both the dropped lo*lo term and the tensor core's own truncation of lo.  (|a| within 2^-11 of FLT_MAX
// would round to infinity; the operands here are activations and weights.)
__device__ __forceinline__ float round_tf32(float a) {
  return __uint_as_float((__float_as_uint(a) + 0x1000u) & 0xffffe000u);
}
__device__ __forceinline__ void split4(const float4& a, float4& hi, float4& lo) {
  hi = make_float4(round_tf32(a.x), round_tf32(a.y), round_tf32(a.z), round_tf32(a.w));
  const f32x2 neg1 = pk2(-1.f, -1.f);
  upk2(fma2(pk2(hi.x, hi.y), neg1, pk2(a.x, a.y)), lo.x, lo.y);
  upk2(fma2(pk2(hi.z, hi.w), neg1, pk2(a.z, a.w)), lo.z, lo.w);
}

// waits that are expected to be long back off, so the waiting warps do not eat the producers' issue slots
__device__ __forceinline__ void mbar_wait_backoff(uint32_t bar, uint32_t parity, unsigned ns) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin) {
    __nanosleep(ns);
    if (spin > (1u << 24)) __trap();
  }
}

struct TcGemmP {
  const float* A; int64_t sam, sak;
  const float* B; int64_t sbk, sbn;
  float* C; int64_t scm, scn;
  int64_t M, N, K;
  float alpha;
  int accumulate;
  int ksplit;
  int64_t k_per_split;
  float* partial;      // [ksplit][M][N] when ksplit > 1
  // fused attribute contraction (SEGNN forward): when yc_Dy > 0 the accumulator columns are (output o, attribute j)
  // pairs, j fastest, and the epilogue stores z[m, o] = bias + sum_j yc_y[m, j] * acc[m, o * yc_Dy + j] into C
  // (row stride scm) instead of the N-wide product; yc_Dy divides 32, no K split, no accumulation
  const float* yc_y; const float* yc_bias;
  int yc_ldy, yc_Dy, yc_bias_col, yc_nbias;
};

template <int BN>
struct GtCfg {
  static constexpr uint32_t A_HALF = GT_M * 128, B_HALF = BN * 128;
  static constexpr uint32_t STAGE = 2 * A_HALF + 2 * B_HALF;
  // three stages: a stage is refilled only after its MMAs have completed, and the refill (wait, split, 16-byte
  // stores, proxy fence, arrive) takes longer than one chunk of MMAs -- it needs two chunks of slack
  static constexpr int NSTAGE = 3;
  static constexpr uint32_t SMEM = NSTAGE * STAGE + 1024 /*align*/ + 256 /*barriers*/;
  static_assert(SMEM <= 227 * 1024, "shared memory budget exceeded");
};

// zo[i] = sum_j yv[j] v[i * DY + j] for the 32 / DY outputs held in 32 accumulator columns (compile-time indices)
template <int DY>
__device__ __forceinline__ void contract_y(const float (&v)[32], const float (&yv)[8], float (&zo)[32]) {
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    float a = 0.f;
    if (i < 32 / DY) {
#pragma unroll
      for (int j = 0; j < DY; ++j) a = fmaf(yv[j], v[i * DY + j], a);
    }
    zo[i] = a;
  }
}

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, 1) tc_gemm_kernel(const TcGemmP p) {
  using Cfg = GtCfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const uint32_t bars = base + Cfg::NSTAGE * Cfg::STAGE;
  auto full = [&](int s) { return bars + 8u * s; };
  auto empty = [&](int s) { return bars + 8u * (Cfg::NSTAGE + s); };
  const uint32_t t_full = bars + 8u * (2 * Cfg::NSTAGE);
  const uint32_t tmem_slot = t_full + 8u;

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n0 = (int64_t)blockIdx.x * BN, m0 = (int64_t)blockIdx.y * GT_M;
  const int64_t kbeg = (int64_t)blockIdx.z * p.k_per_split;
  const int64_t kend = (kbeg + p.k_per_split < p.K) ? kbeg + p.k_per_split : p.K;
  const int nch = (int)((kend - kbeg + 31) / 32);

  if (tid == 0) {
    for (int s = 0; s < Cfg::NSTAGE; ++s) {
      mbar_init(full(s), GT_PROD_WARPS);
      mbar_init(empty(s), 1);
    }
    mbar_init(t_full, 1);
    mbar_fence_init();
  }
  if (warp == GT_MMA_WARP) tmem_alloc(tmem_slot, BN);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  if (warp >= GT_EPI_WARPS) {
    // =========================== producers ===========================
    // Warps 4-11 feed A, warps 12-19 feed B; both operands are 128-row x 32-k tiles stored K-major, so one code path
    // serves them.  Every thread owns four groups of 4 consecutive k per chunk (16 elements):
    //   k contiguous in global memory   : group q = (row t/8 + 32 q, k 4 (t%8) ..)   -- 8 threads read one 128-byte row
    //   rows contiguous in global memory: group q = (row t%128, k 16 (t/128) + 4 q ..) -- a warp reads 32 consecutive rows
    // Zero fill past M / N / K.  Each group becomes one hi and one lo 16-byte store into the swizzled tile
    // (conflict-free in both maps).
    static_assert(BN == GT_M, "the two producer teams assume equally sized operand tiles");
    const int pt = tid - GT_EPI_WARPS * 32;
    const int ot = pt & 255;
    const bool is_b = pt >= 256;
    const int64_t s_row = is_b ? p.sbn : p.sam, s_k = is_b ? p.sbk : p.sak;
    const float* obase = is_b ? p.B + n0 * p.sbn + kbeg * p.sbk : p.A + m0 * p.sam + kbeg * p.sak;
    const int64_t rows_avail = is_b ? p.N - n0 : p.M - m0;
    const uint32_t st_off = is_b ? 2 * Cfg::A_HALF : 0u;
    const bool kmode = s_k <= s_row;
    const bool vec = kmode && s_k == 1 && (s_row & 3) == 0 && (reinterpret_cast<uintptr_t>(obase) & 15) == 0;
    const float* gq[4];
    uint32_t offq[4];
    int kq[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const int row = kmode ? (ot >> 3) + 32 * q : (ot & 127);
      kq[q] = kmode ? (ot & 7) * 4 : (ot >> 7) * 16 + 4 * q;
      // rows past the end of A / B re-read the last one: they only feed accumulator rows / columns never stored
      const int64_t grow = row < rows_avail ? row : rows_avail - 1;
      gq[q] = obase + grow * s_row + kq[q] * s_k;
      offq[q] = st_off + sw128((uint32_t)row, (uint32_t)kq[q] >> 2);
    }
    auto publish = [&](int c, const float4 (&g)[4]) {
      const int s = c % Cfg::NSTAGE;
      mbar_wait_backoff(empty(s), ((c / Cfg::NSTAGE) & 1u) ^ 1u, 32);
      uint8_t* stage = gen + s * Cfg::STAGE;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        float4 hi, lo;
        split4(g[q], hi, lo);
        *reinterpret_cast<float4*>(stage + offq[q]) = hi;
        *reinterpret_cast<float4*>(stage + Cfg::A_HALF + offq[q]) = lo;
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(full(s));            // one arrival per producer warp
    };
    // loads go to registers, two chunks ahead of the one being published (16-byte loads when the operand allows it);
    // only the last chunk of a K range can be partial, and only there are loads predicated (zero fill past K)
    auto ldg_chunk = [&](int c, float4 (&g)[4]) {
      if (c >= nch) return;
      const int64_t krem = (kend - kbeg) - 32 * (int64_t)c;           // k entries of this chunk that exist
      const int64_t adv = 32 * (int64_t)c * s_k;
      if (krem >= 32) {
        if (vec) {
#pragma unroll
          for (int q = 0; q < 4; ++q) g[q] = __ldg(reinterpret_cast<const float4*>(gq[q] + adv));
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const float* src = gq[q] + adv;
            g[q] = make_float4(__ldg(src), __ldg(src + s_k), __ldg(src + 2 * s_k), __ldg(src + 3 * s_k));
          }
        }
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float* src = gq[q] + adv;
          g[q].x = (kq[q] + 0 < krem) ? __ldg(src) : 0.f;
          g[q].y = (kq[q] + 1 < krem) ? __ldg(src + s_k) : 0.f;
          g[q].z = (kq[q] + 2 < krem) ? __ldg(src + 2 * s_k) : 0.f;
          g[q].w = (kq[q] + 3 < krem) ? __ldg(src + 3 * s_k) : 0.f;
        }
      }
    };
    {
      float4 g0[4], g1[4], g2[4];
      ldg_chunk(0, g0);
      ldg_chunk(1, g1);
      for (int c = 0; c < nch; c += 3) {
        ldg_chunk(c + 2, g2);
        publish(c, g0);
        if (c + 1 < nch) {
          ldg_chunk(c + 3, g0);
          publish(c + 1, g1);
        }
        if (c + 2 < nch) {
          ldg_chunk(c + 4, g1);
          publish(c + 2, g2);
        }
      }
    }
  } else {
    // =========================== MMA issuer (warp 0, one thread), then epilogue ===========================
    const int64_t m = m0 + warp * 32 + lane;         // this thread's accumulator row (tcgen05.ld 32x32b)
    float yv[8];                                     // its attributes (fused contraction, Dy <= 8): loaded up front so
    if (p.yc_Dy > 0) {                               // that the latency hides behind the main loop
#pragma unroll
      for (int j = 0; j < 8; ++j) yv[j] = (j < p.yc_Dy && m < p.M) ? __ldg(p.yc_y + m * p.yc_ldy + j) : 0.f;
    }
    if (warp == GT_MMA_WARP && lane == 0) {
      constexpr uint32_t idesc = make_idesc_tf32(GT_M, BN, 0, 0);
      for (int c = 0; c < nch; ++c) {
        const int s = c % Cfg::NSTAGE;
        mbar_wait(full(s), (c / Cfg::NSTAGE) & 1u);
        tc_fence_after();
        const uint32_t a_hi = base + s * Cfg::STAGE, a_lo = a_hi + Cfg::A_HALF;
        const uint32_t b_hi = a_hi + 2 * Cfg::A_HALF, b_lo = b_hi + Cfg::B_HALF;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t dah = make_desc(a_hi + kk * 32, 16, 1024), dal = make_desc(a_lo + kk * 32, 16, 1024);
          const uint64_t dbh = make_desc(b_hi + kk * 32, 16, 1024), dbl = make_desc(b_lo + kk * 32, 16, 1024);
          mma_tf32(tmem_base, dah, dbh, idesc, (c | kk) ? 1u : 0u);
          mma_tf32(tmem_base, dal, dbh, idesc, 1u);
          mma_tf32(tmem_base, dah, dbl, idesc, 1u);
        }
        mma_commit(empty(s));
      }
      mma_commit(t_full);
    }
    __syncwarp();
    // thread = accumulator row (tcgen05.ld 32x32b); row-major destinations go through a 32 x 33 shared-memory
    // transpose (the operand stages are dead once t_full fires) so that lanes run along n.
    mbar_wait_backoff(t_full, 0, 128);
    tc_fence_after();
    const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);
    float* stg = reinterpret_cast<float*>(gen) + warp * (32 * 33);
    const bool split = p.ksplit > 1;
    float* dst = split ? p.partial + (size_t)blockIdx.z * p.M * p.N : p.C;
    const int64_t dm = split ? p.N : p.scm, dn = split ? 1 : p.scn;
    const bool acc = !split && p.accumulate;
    const bool transpose = dn <= dm;
    // row-major destination with 32-byte-aligned rows: every thread writes whole sectors of its own row with
    // 256-bit stores -- four instructions per 32 columns instead of a shared-memory transpose and 32 scalar stores
    const bool rows256 = dn == 1 && !acc && (dm & 7) == 0 && (reinterpret_cast<uintptr_t>(dst) & 31) == 0;
#pragma unroll 1
    for (int cb = 0; cb < BN / 32; ++cb) {
      const int64_t nb = n0 + cb * 32;
      if (nb >= p.N) break;                          // warp-uniform
      float v[32];
      tmem_ld32(taddr + cb * 32, v);
      tmem_ld_wait();
      if (p.yc_Dy > 0) {
        // 32 accumulator columns = 32 / Dy whole outputs: contract with y, add the bias of the 0e block, and write
        // the outputs of this row as one run
        const int Dy = p.yc_Dy, per = 32 / Dy;
        const int64_t o0 = nb / Dy, Do = p.N / Dy;
        const bool has_bias = p.yc_bias != nullptr && o0 < p.yc_bias_col + p.yc_nbias && o0 + per > p.yc_bias_col;
        if (Dy == 4 && o0 + 8 <= Do && (p.scm & 7) == 0 && (reinterpret_cast<uintptr_t>(p.C) & 31) == 0) {
          // attributes l <= 1: eight outputs per 32 columns, one 256-bit store per row
          float zo[8];
#pragma unroll
          for (int i = 0; i < 8; ++i)
            zo[i] = fmaf(yv[3], v[4 * i + 3], fmaf(yv[2], v[4 * i + 2], fmaf(yv[1], v[4 * i + 1], yv[0] * v[4 * i])));
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int64_t o = o0 + i;
              if (o >= p.yc_bias_col && o < p.yc_bias_col + p.yc_nbias) zo[i] += __ldg(p.yc_bias + (o - p.yc_bias_col));
            }
          }
          if (m < p.M) stg256(p.C + m * p.scm + o0, zo);
        } else {
          float zo[32];
          if (Dy == 4) contract_y<4>(v, yv, zo);
          else if (Dy == 8) contract_y<8>(v, yv, zo);
          else if (Dy == 2) contract_y<2>(v, yv, zo);
          else contract_y<1>(v, yv, zo);
          if (m < p.M) {
            float* q = p.C + m * p.scm + o0;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int64_t o = o0 + i;
              if (i < per && o < Do) {
                float t = zo[i];
                if (has_bias && o >= p.yc_bias_col && o < p.yc_bias_col + p.yc_nbias) t += __ldg(p.yc_bias + (o - p.yc_bias_col));
                q[i] = t;
              }
            }
          }
        }
      } else if (rows256 && nb + 32 <= p.N) {
        if (m < p.M) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
          float* q = dst + m * dm + nb;
#pragma unroll
          for (int j = 0; j < 32; j += 8) stg256(q + j, v + j);
        }
      } else if (transpose) {
#pragma unroll
        for (int j = 0; j < 32; ++j) stg[lane * 33 + j] = v[j];
        __syncwarp();
        const int64_t n = nb + lane;
        if (n < p.N) {
          float* q = dst + (m0 + warp * 32) * dm + n * dn;
          const int64_t left = p.M - (m0 + warp * 32);
          const int rows = left < 32 ? (int)(left < 0 ? 0 : left) : 32;
          if (acc) {                                   // all reads first: the loads are independent of the stores
#pragma unroll
            for (int r = 0; r < 32; ++r) v[r] = r < rows ? q[r * dm] : 0.f;
#pragma unroll
            for (int r = 0; r < 32; ++r)
              if (r < rows) q[r * dm] = fmaf(p.alpha, stg[r * 33 + lane], v[r]);
          } else {
#pragma unroll 8
            for (int r = 0; r < rows; ++r) q[r * dm] = p.alpha * stg[r * 33 + lane];
          }
        }
        __syncwarp();
      } else if (m < p.M) {
#pragma unroll
        for (int j = 0; j < 32; ++j) {
          if (nb + j < p.N) {
            float* q = dst + m * dm + (nb + j) * dn;
            const float o = p.alpha * v[j];
            *q = acc ? *q + o : o;
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == GT_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, BN);
  }
}

// C[m, n] (+)= sum_z partial[z][m][n], z ascending (deterministic)
__global__ void tc_gemm_reduce_kernel(const TcGemmP p) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= p.M * p.N) return;
  const int64_t m = i / p.N, n = i - m * p.N;
  float s = 0.f;
  for (int z = 0; z < p.ksplit; ++z) s += p.partial[(size_t)z * p.M * p.N + i];
  float* q = p.C + m * p.scm + n * p.scn;
  *q = p.accumulate ? *q + s : s;
}

struct GtPlan {
  int bn, ksplit;
  int64_t k_per_split;
};

GtPlan gt_plan(int64_t M, int64_t N, int64_t K) {
  GtPlan g;
  g.bn = 128;
  const int64_t tiles = ((M + GT_M - 1) / GT_M) * ((N + g.bn - 1) / g.bn);
  int64_t ks = 1;
  if (tiles < 148 && K >= 1024) {
    ks = (2 * 148) / tiles;
    const int64_t most = K / 256;                  // at least 8 k chunks per split
    if (ks > most) ks = most;
    if (ks < 1) ks = 1;
  }
  // The tensor core adds into the fp32 accumulator without round-to-nearest, an error that grows with the number of
  // products summed in TMEM (2e-5 of the result after 3000, measured): no split accumulates more than 1024 k
  // entries, the partials are summed in fp32 registers by the reduction kernel.  (Bounded by 256 MB of partials.)
  const int64_t for_accuracy = (K + 1023) / 1024;
  const int64_t fits = (int64_t)(256u << 20) / (M * N * (int64_t)sizeof(float));
  if (ks < for_accuracy) ks = for_accuracy < fits ? for_accuracy : (fits > ks ? fits : ks);
  int64_t per = (K + ks - 1) / ks;
  per = (per + 31) / 32 * 32;
  g.k_per_split = per;
  g.ksplit = (int)((K + per - 1) / per);
  return g;
}

bool gt_shape_ok(int64_t M, int64_t N, int64_t K) {
  return M >= 32 && N >= 32 && K >= 16 && M * N >= (1 << 14) && M * N * K >= (1 << 24) && N < (1ll << 31) &&
         (M + GT_M - 1) / GT_M <= 65535;
}

template <int BN>
int launch_tc_gemm(const TcGemmP& p, cudaStream_t st) {
  auto kern = tc_gemm_kernel<BN>;
  static tc::DeviceOnce once;   // one per kernel instantiation; the attribute is per device
  EQNN_CUDA(tc::configure_smem(once, kern, (int)GtCfg<BN>::SMEM));
  dim3 grid((unsigned)((p.N + BN - 1) / BN), (unsigned)((p.M + GT_M - 1) / GT_M), (unsigned)p.ksplit);
  kern<<<grid, GT_THREADS, GtCfg<BN>::SMEM, st>>>(p);
  if (p.ksplit > 1) tc_gemm_reduce_kernel<<<(unsigned)ceil_div64(p.M * p.N, 256), 256, 0, st>>>(p);
  EQNN_CHECK_LAUNCH_N(p.ksplit > 1 ? 2 : 1);
  return EQNN_OK;
}

}  // namespace

size_t eqnn_tc_gemm_workspace_bytes(int64_t M, int64_t N, int64_t K) {
  if (!gt_shape_ok(M, N, K)) return 0;
  const GtPlan g = gt_plan(M, N, K);
  return g.ksplit > 1 ? (size_t)g.ksplit * M * N * sizeof(float) : 0;
}

// returns 1 when the GEMM was launched here, 0 when the shape is declined, < 0 on error (-status)
int eqnn_tc_gemm_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                     const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                     int pro, int epi, void* ws, size_t ws_bytes, cudaStream_t st) {
  if (pro != 0 || epi != 0 || !gt_shape_ok(M, N, K)) return 0;
  const GtPlan g = gt_plan(M, N, K);
  if (g.ksplit > 1 && (!ws || ws_bytes < (size_t)g.ksplit * M * N * sizeof(float))) return 0;
  TcGemmP p;
  p.A = A; p.sam = sam; p.sak = sak;
  p.B = B; p.sbk = sbk; p.sbn = sbn;
  p.C = C; p.scm = scm; p.scn = scn;
  p.M = M; p.N = N; p.K = K;
  p.alpha = alpha; p.accumulate = accumulate;
  p.ksplit = g.ksplit; p.k_per_split = g.k_per_split;
  p.partial = reinterpret_cast<float*>(ws);
  p.yc_y = nullptr; p.yc_bias = nullptr; p.yc_ldy = p.yc_Dy = p.yc_bias_col = p.yc_nbias = 0;
  const int rc = launch_tc_gemm<128>(p, st);
  return rc ? -rc : 1;
}

// ---- SEGNN TensorProductLinearGate forward with the attribute contraction in the GEMM epilogue ----------------
// z[r, o] = bias + sum_j y[r, j] (x @ M_cat)[r, o * Dy + j]: the (R x Do*Dy) product never leaves tensor memory.
extern "C" int eqnn_segnn_tplg_fwd(const float* x, int ld_x, int Dx, const float* Mcat, const float* y, int ld_y, int Dy,
                                   int Do, const float* bias, int bias_col, int n_bias, int64_t R, float* z, int ld_z,
                                   eqnn_stream_t stream) {
  if (R <= 0) return EQNN_OK;
  const int64_t N = (int64_t)Dy * Do;
  EQNN_CHECK_ARG(x && Mcat && y && z && Dx >= 1 && ld_x >= Dx && Do >= 1 && ld_z >= Do && ld_y >= Dy &&
                     (Dy == 1 || Dy == 2 || Dy == 4 || Dy == 8) && n_bias >= 0 && bias_col >= 0 && bias_col + n_bias <= Do &&
                     (n_bias == 0 || bias),
                 "segnn_tplg_fwd: bad argument (Dy must be 1, 2, 4 or 8)");
  EQNN_CHECK_ARG(gt_shape_ok(R, N, Dx), "segnn_tplg_fwd: shape not covered by the tensor-core kernel");
  TcGemmP p;
  p.A = x; p.sam = ld_x; p.sak = 1;
  p.B = Mcat; p.sbk = N; p.sbn = 1;
  p.C = z; p.scm = ld_z; p.scn = 1;
  p.M = R; p.N = N; p.K = Dx;
  p.alpha = 1.f; p.accumulate = 0;
  p.ksplit = 1; p.k_per_split = ((int64_t)Dx + 31) / 32 * 32;
  p.partial = nullptr;
  p.yc_y = y; p.yc_bias = n_bias ? bias : nullptr; p.yc_ldy = ld_y; p.yc_Dy = Dy; p.yc_bias_col = bias_col; p.yc_nbias = n_bias;
  return launch_tc_gemm<128>(p, as_stream(stream));
}
