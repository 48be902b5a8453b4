// Periodic-boundary k-nearest-neighbour graph construction (sm_100a).
//
// Replaces models/utils/graph_utils.py:40-72 (dense (N,N,3) difference + full argsort)
// with a tiled sweep that never materialises the N x N matrix: candidates are staged in
// shared-memory tiles, each warp keeps the running k best of its queries spread over its
// lanes and updates them with ballot + shuffle (warp-level top-k, fixed tie-break).
// Distances follow SURVEY.md C.3 with every fp32 operation rounded separately, so the
// neighbour indices are bit-identical to the reference's stable argsort.
#include "common.cuh"

namespace {

constexpr int WARPS = 8;      // warps per CTA
constexpr int QPW = 4;        // queries per warp (register-blocked: one candidate load feeds 4 distances)
constexpr int TJ = 2048;      // candidates per shared-memory tile (32 KB)
constexpr float KNN_EPS = 1e-7f;

struct Cell {
  float c[9];   // cell, row-major
  float ci[9];  // inverse cell
};

template <bool PBC, bool DIAG, bool FMA>
__device__ __forceinline__ float pair_dist(float xi, float yi, float zi, float xj, float yj, float zj,
                                           const Cell& cl, float* out3 = nullptr) {
  float d0 = __fadd_rn(__fsub_rn(xi, xj), KNN_EPS);
  float d1 = __fadd_rn(__fsub_rn(yi, yj), KNN_EPS);
  float d2 = __fadd_rn(__fsub_rn(zi, zj), KNN_EPS);
  if (PBC) {
    if (DIAG) {
      float n0 = rintf(__fmul_rn(d0, cl.ci[0]));
      float n1 = rintf(__fmul_rn(d1, cl.ci[4]));
      float n2 = rintf(__fmul_rn(d2, cl.ci[8]));
      d0 = __fsub_rn(d0, __fmul_rn(n0, cl.c[0]));
      d1 = __fsub_rn(d1, __fmul_rn(n1, cl.c[4]));
      d2 = __fsub_rn(d2, __fmul_rn(n2, cl.c[8]));
    } else {
      float n[3], w[3];
#pragma unroll
      for (int c = 0; c < 3; ++c)
        n[c] = rintf(__fadd_rn(__fadd_rn(__fmul_rn(d0, cl.ci[c]), __fmul_rn(d1, cl.ci[3 + c])),
                               __fmul_rn(d2, cl.ci[6 + c])));
#pragma unroll
      for (int c = 0; c < 3; ++c)
        w[c] = __fadd_rn(__fadd_rn(__fmul_rn(n[0], cl.c[c]), __fmul_rn(n[1], cl.c[3 + c])),
                         __fmul_rn(n[2], cl.c[6 + c]));
      d0 = __fsub_rn(d0, w[0]);
      d1 = __fsub_rn(d1, w[1]);
      d2 = __fsub_rn(d2, w[2]);
    }
  }
  if (out3) {
    out3[0] = d0;
    out3[1] = d1;
    out3[2] = d2;
  }
  if (FMA) return __fmaf_rn(d2, d2, __fmaf_rn(d1, d1, __fmul_rn(d0, d0)));
  return __fadd_rn(__fadd_rn(__fmul_rn(d0, d0), __fmul_rn(d1, d1)), __fmul_rn(d2, d2));
}

// One warp owns QPW queries; the 32 lanes sweep 32 candidates per step (one conflict-free
// LDS.128 each).  The k best of a query live ONE ENTRY PER LANE (R entries per lane for
// k > 32) as a sorted list; candidates that beat the current k-th distance are found with a
// ballot and inserted one at a time in ascending index order by a shuffle-shift:
//     entry e takes its left neighbour if the new distance is smaller than the neighbour's,
//     the new candidate if it is smaller than its own, else keeps its value.
// Strict '<' everywhere + ascending arrival order = the stable argsort of the reference.
template <int R, bool PBC, bool DIAG, bool FMA>
__global__ void __launch_bounds__(WARPS * 32)
knn_kernel(const float* __restrict__ x, int ldx, int N, int k, Cell cell, int32_t* __restrict__ senders,
           int32_t* __restrict__ receivers, float* __restrict__ dr) {
  __shared__ float4 tile[TJ];
  const unsigned FULL = 0xffffffffu;
  const float INF = __int_as_float(0x7f800000);
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float* xb = x + (size_t)b * N * ldx;
  const int q0 = (blockIdx.x * WARPS + warp) * QPW;

  float qx[QPW], qy[QPW], qz[QPW], thr[QPW];
  float ed[QPW][R];
  int ej[QPW][R];
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    const int qi = min(q0 + q, N - 1);
    qx[q] = xb[(size_t)qi * ldx + 0];
    qy[q] = xb[(size_t)qi * ldx + 1];
    qz[q] = xb[(size_t)qi * ldx + 2];
    thr[q] = INF;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ed[q][r] = INF;
      ej[q][r] = 0x7fffffff;
    }
  }
  const int kr = (k - 1) >> 5, kl = (k - 1) & 31;   // register / lane holding the k-th entry

  for (int j0 = 0; j0 < N; j0 += TJ) {
    const int nt = min(TJ, N - j0);
    __syncthreads();
    for (int t = threadIdx.x; t < nt; t += WARPS * 32) {
      const float* p = xb + (size_t)(j0 + t) * ldx;
      tile[t] = make_float4(p[0], p[1], p[2], 0.f);
    }
    __syncthreads();
    if (q0 >= N) continue;
    for (int tb = 0; tb < nt; tb += 32) {
      const int t = tb + lane;
      const bool tv = t < nt;
      const float4 c = tile[tv ? t : 0];
      float D[QPW];
#pragma unroll
      for (int q = 0; q < QPW; ++q) D[q] = pair_dist<PBC, DIAG, FMA>(qx[q], qy[q], qz[q], c.x, c.y, c.z, cell);
#pragma unroll
      for (int q = 0; q < QPW; ++q) {
        unsigned m = __ballot_sync(FULL, tv && D[q] < thr[q]);
        while (m) {
          const int srcl = __ffs(m) - 1;
          m &= m - 1;
          const float Dn = __shfl_sync(FULL, D[q], srcl);
          if (!(Dn < thr[q])) continue;          // threshold moved since the ballot (warp-uniform)
          const int jn = j0 + tb + srcl;
#pragma unroll
          for (int r = R - 1; r >= 0; --r) {
            float ld = __shfl_up_sync(FULL, ed[q][r], 1);
            int lj = __shfl_up_sync(FULL, ej[q][r], 1);
            if (r > 0) {
              const float wd = __shfl_sync(FULL, ed[q][r - 1], 31);
              const int wj = __shfl_sync(FULL, ej[q][r - 1], 31);
              if (lane == 0) { ld = wd; lj = wj; }
            } else if (lane == 0) {
              ld = -INF;
            }
            if (r * 32 + lane < k) {
              if (Dn < ld) { ed[q][r] = ld; ej[q][r] = lj; }
              else if (Dn < ed[q][r]) { ed[q][r] = Dn; ej[q][r] = jn; }
            }
          }
          float tq = ed[q][0];
          if (R > 1 && kr == 1) tq = ed[q][R - 1];
          thr[q] = __shfl_sync(FULL, tq, kl);
        }
      }
    }
  }
#pragma unroll
  for (int q = 0; q < QPW; ++q) {
    const int qi = q0 + q;
    if (qi >= N) continue;
    const size_t ebase = ((size_t)b * N + qi) * k;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int e = r * 32 + lane;
      if (e < k) {
        const int j = ej[q][r];
        senders[ebase + e] = qi;
        receivers[ebase + e] = j;
        if (dr) {
          float o[3] = {0.f, 0.f, 0.f};
          if (j >= 0 && j < N) {
            const float* p = xb + (size_t)j * ldx;
            pair_dist<PBC, DIAG, false>(qx[q], qy[q], qz[q], p[0], p[1], p[2], cell, o);
          }
          dr[(ebase + e) * 3 + 0] = o[0];
          dr[(ebase + e) * 3 + 1] = o[1];
          dr[(ebase + e) * 3 + 2] = o[2];
        }
      }
    }
  }
}

template <int R>
int launch_knn(const float* x, int ldx, int N, int k, const Cell& cell, bool pbc, bool diag, bool fma, int B,
               int32_t* s, int32_t* r, float* dr, cudaStream_t st) {
  dim3 grid((N + WARPS * QPW - 1) / (WARPS * QPW), B);
#define EQNN_KNN_LAUNCH(P, D, F) knn_kernel<R, P, D, F><<<grid, WARPS * 32, 0, st>>>(x, ldx, N, k, cell, s, r, dr)
  if (!pbc) {
    if (fma) EQNN_KNN_LAUNCH(false, true, true); else EQNN_KNN_LAUNCH(false, true, false);
  } else if (diag) {
    if (fma) EQNN_KNN_LAUNCH(true, true, true); else EQNN_KNN_LAUNCH(true, true, false);
  } else {
    if (fma) EQNN_KNN_LAUNCH(true, false, true); else EQNN_KNN_LAUNCH(true, false, false);
  }
#undef EQNN_KNN_LAUNCH
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

}  // namespace

extern "C" size_t eqnn_knn_workspace_bytes(int, int, int) { return 0; }

extern "C" int eqnn_knn_pbc(const float* x, int ld_x, const float* cell, const float* cell_inv, int B, int N,
                            int k, int flags, int32_t* senders, int32_t* receivers, float* dr, void*, size_t,
                            eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && senders && receivers, "knn: null pointer");
  EQNN_CHECK_ARG(B >= 0 && N >= 0 && ld_x >= 3, "knn: bad shape");
  EQNN_CHECK_ARG(k >= 1 && k <= N, "knn: need 1 <= k <= N");
  EQNN_CHECK_ARG(k <= 64, "knn: k > 64 not supported");
  EQNN_CHECK_ARG((int64_t)B * N * k < (int64_t)1 << 31, "knn: more than 2^31 edges");
  if (B == 0 || N == 0) return EQNN_OK;
  const bool pbc = (flags & EQNN_KNN_PBC) != 0;
  Cell c;
  bool diag = true;
  for (int i = 0; i < 9; ++i) {
    c.c[i] = (i % 4 == 0) ? 1.f : 0.f;
    c.ci[i] = c.c[i];
  }
  if (pbc) {
    EQNN_CHECK_ARG(cell && cell_inv, "knn: PBC requested without cell / cell_inv");
    for (int i = 0; i < 9; ++i) {
      c.c[i] = cell[i];
      c.ci[i] = cell_inv[i];
      if (i % 4 != 0 && (cell[i] != 0.f || cell_inv[i] != 0.f)) diag = false;
    }
  }
  const bool fma = (flags & EQNN_KNN_FMA) != 0;
  cudaStream_t st = as_stream(stream);
  if (k <= 32) return launch_knn<1>(x, ld_x, N, k, c, pbc, diag, fma, B, senders, receivers, dr, st);
  return launch_knn<2>(x, ld_x, N, k, c, pbc, diag, fma, B, senders, receivers, dr, st);
}
