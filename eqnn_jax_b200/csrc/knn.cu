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

// EPS: nearest_neighbors adds 1e-7 to the difference (graph_utils.py:55); compute_distances (:223) does not
template <bool PBC, bool DIAG, bool FMA, bool EPS = true>
__device__ __forceinline__ float pair_dist(float xi, float yi, float zi, float xj, float yj, float zj,
                                           const Cell& cl, float* out3 = nullptr) {
  float d0 = __fsub_rn(xi, xj), d1 = __fsub_rn(yi, yj), d2 = __fsub_rn(zi, zj);
  if (EPS) {
    d0 = __fadd_rn(d0, KNN_EPS);
    d1 = __fadd_rn(d1, KNN_EPS);
    d2 = __fadd_rn(d2, KNN_EPS);
  }
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


// ------------------------------------------------------------------------------------------------
// Cell-list search (periodic, diagonal cell): same distances, same order, far fewer pairs.
//
// Points are binned into G^3 cells of the periodic box and stored cell by cell.  A warp owns one query and
// sweeps the (2R+1)^3 cells around it.  The result is accepted only if the k-th distance lies strictly inside
// the searched region -- sqrt(D_k) < R*h - margin, h = cell width, margin covers fp32 binning and the +1e-7 of
// the distance expression -- because then every point at distance <= D_k was a candidate; otherwise R grows
// (up to the whole box = brute force).  Distances come from the same pair_dist as the dense sweep and the list
// is ordered lexicographically by (distance, index), which is exactly the stable argsort of the reference, so the
// output is bit-identical to the dense path whatever the visiting order.
// ------------------------------------------------------------------------------------------------
struct Grid {
  int G;            // cells per axis
  float h[3];       // cell width per axis (box length / G)
  float margin[3];  // safety margin per axis
};

__device__ __forceinline__ int cell_coord(float x, float ci, int G) {
  float u = x * ci;
  u -= floorf(u);                       // [0, 1]
  int c = (int)(u * (float)G);
  return c >= G ? G - 1 : (c < 0 ? 0 : c);
}

// position of x inside its cell c (cell_coord's arithmetic), in [0, 1]
__device__ __forceinline__ float cell_frac(float x, float ci, int G, int c) {
  float u = x * ci;
  u -= floorf(u);
  return fminf(fmaxf(u * (float)G - (float)c, 0.f), 1.f);
}

__global__ void knn_bin_kernel(const float* __restrict__ x, int ldx, int N, Cell cl, Grid g, int32_t* __restrict__ cell_of,
                               int32_t* __restrict__ count) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const float* p = x + ((size_t)b * N + i) * ldx;
  const int cx = cell_coord(p[0], cl.ci[0], g.G), cy = cell_coord(p[1], cl.ci[4], g.G), cz = cell_coord(p[2], cl.ci[8], g.G);
  const int c = (cz * g.G + cy) * g.G + cx;
  cell_of[(size_t)b * N + i] = c;
  atomicAdd(count + (size_t)b * (g.G * g.G * g.G + 1) + c, 1);
}

// exclusive scan of the per-cell counts of one cloud (G^3 <= 4096 cells): one CTA per cloud
__global__ void __launch_bounds__(1024) knn_scan_kernel(int32_t* __restrict__ count, int n_cells) {
  __shared__ int sm[1024];
  int32_t* c = count + (size_t)blockIdx.x * (n_cells + 1);
  const int per = (n_cells + 1023) / 1024;
  const int t = threadIdx.x, lo = t * per, hi = min(lo + per, n_cells);
  int s = 0;
  for (int i = lo; i < hi; ++i) s += c[i];
  sm[t] = s;
  __syncthreads();
  for (int o = 1; o < 1024; o <<= 1) {
    const int v = (t >= o) ? sm[t - o] : 0;
    __syncthreads();
    sm[t] += v;
    __syncthreads();
  }
  int run = sm[t] - s;
  for (int i = lo; i < hi; ++i) {
    const int v = c[i];
    c[i] = run;
    run += v;
  }
  if (t == 1023) c[n_cells] = sm[1023];
}

// sorted[b][slot] = (x, y, z, original index); slot order inside a cell is arbitrary (the search does not depend on it)
__global__ void knn_scatter_kernel(const float* __restrict__ x, int ldx, int N, int n_cells, const int32_t* __restrict__ cell_of,
                                   const int32_t* __restrict__ start, int32_t* __restrict__ cursor,
                                   float4* __restrict__ sorted) {
  const int b = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int c = cell_of[(size_t)b * N + i];
  const int slot = start[(size_t)b * (n_cells + 1) + c] + atomicAdd(cursor + (size_t)b * n_cells + c, 1);
  const float* p = x + ((size_t)b * N + i) * ldx;
  sorted[(size_t)b * N + slot] = make_float4(p[0], p[1], p[2], __int_as_float(i));
}

__device__ __forceinline__ bool lex_lt(float d, int j, float d2, int j2) { return d < d2 || (d == d2 && j < j2); }

#ifndef KNN_MERGE_MIN
#define KNN_MERGE_MIN 8      // candidates per batch from which the bitonic merge beats serial insertion (measured: 5 and 8 both 0.49-0.50 ms per step)
#endif
template <int R, bool FMA>
__global__ void __launch_bounds__(WARPS * 32)
knn_cell_kernel(const float4* __restrict__ sorted, const int32_t* __restrict__ start, int N, int k, Cell cell, Grid g,
                int32_t* __restrict__ senders, int32_t* __restrict__ receivers, float* __restrict__ dr) {
  const unsigned FULL = 0xffffffffu;
  const float INF = __int_as_float(0x7f800000);
  const int b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = blockIdx.x * WARPS + warp;
  if (slot >= N) return;
  const int G = g.G, n_cells = G * G * G;
  const float4* pts = sorted + (size_t)b * N;
  const int32_t* cs = start + (size_t)b * (n_cells + 1);
  const float4 q = pts[slot];
  const int qi = __float_as_int(q.w);
  const int cx = cell_coord(q.x, cell.ci[0], G), cy = cell_coord(q.y, cell.ci[4], G), cz = cell_coord(q.z, cell.ci[8], G);
  const int kr = (k - 1) >> 5, kl = (k - 1) & 31;
  // Distance from the query to the faces of its own cell, per axis and side, less the safety margin: a cell at offset
  // o != 0 along an axis holds no point closer than (|o| - 1) h + that distance along the axis, so a row of cells (or
  // the cells at the ends of a row) whose bound already exceeds the current k-th distance cannot change the list.
  float gap_dn[3], gap_up[3];
  {
    const float f[3] = {cell_frac(q.x, cell.ci[0], G, cx), cell_frac(q.y, cell.ci[4], G, cy), cell_frac(q.z, cell.ci[8], G, cz)};
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      gap_dn[a] = f[a] * g.h[a] - g.margin[a];
      gap_up[a] = (1.f - f[a]) * g.h[a] - g.margin[a];
    }
  }
  auto axis_gap = [&](int a, int o) {       // lower bound of |displacement| along axis a to a cell at signed offset o
    if (o == 0) return 0.f;
    const float gp = (float)(abs(o) - 1) * g.h[a] + (o > 0 ? gap_up[a] : gap_dn[a]);
    return fmaxf(gp, 0.f);
  };
  float ed[R];
  int ej[R];
  for (int Rr = 1;; ++Rr) {
    float thr = INF;
    int thrj = 0x7fffffff;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      ed[r] = INF;
      ej[r] = 0x7fffffff;
    }
    const int span = min(2 * Rr + 1, G);                   // cells visited per axis (never the same cell twice)
    const int z0 = (span == G) ? 0 : cz - Rr, y0 = (span == G) ? 0 : cy - Rr, x0 = (span == G) ? 0 : cx - Rr;
    // Rows of cells are visited centre-out (offset 0, +1, -1, +2, ...): the query's own row first, so that the k-th
    // distance is close to its final value after the first few dozen candidates and far fewer of the later ones pass
    // the threshold test into the (serial) insertion.  The result does not depend on the order: the list is kept in
    // (distance, index) order and holds the k smallest of everything visited.
    for (int iz = 0; iz < span; ++iz) {
      const int dz = (span == G) ? iz : Rr + ((iz & 1) ? (iz + 1) / 2 : -(iz / 2));
      const int zz = ((z0 + dz) % G + G) % G;
      for (int iy = 0; iy < span; ++iy) {
        const int dy = (span == G) ? iy : Rr + ((iy & 1) ? (iy + 1) / 2 : -(iy / 2));
        const int yy = ((y0 + dy) % G + G) % G;
        const int rowc = (zz * G + yy) * G;
        // the x-run [x0, x0 + span) modulo G is one or two contiguous slot ranges
        int xa = ((x0 % G) + G) % G, len = span;
        if (span != G) {
          // prune: the whole row, or cells from both ends of it, by the lower bound of the squared distance
          const float thr_hi = thr * 1.00001f;                 // (inf stays inf: nothing is pruned before k candidates)
          const float gz = axis_gap(2, dz - Rr), gy = axis_gap(1, dy - Rr);
          const float rowb = gz * gz + gy * gy;
          if (rowb > thr_hi) continue;
          int o_lo = -Rr, o_hi = Rr;
          while (o_lo < 0) {
            const float gx = axis_gap(0, o_lo);
            if (rowb + gx * gx > thr_hi) ++o_lo; else break;
          }
          while (o_hi > 0) {
            const float gx = axis_gap(0, o_hi);
            if (rowb + gx * gx > thr_hi) --o_hi; else break;
          }
          xa = (((x0 + o_lo + Rr) % G) + G) % G;
          len = o_hi - o_lo + 1;
        }
        while (len > 0) {
          const int run = min(len, G - xa);
          const int p0 = cs[rowc + xa], p1 = cs[rowc + xa + run];
          for (int tb = p0; tb < p1; tb += 32) {
            const int t = tb + lane;
            const bool tv = t < p1;
            const float4 c = pts[tv ? t : p0];
            const int jc = __float_as_int(c.w);
            const float D = pair_dist<true, true, FMA>(q.x, q.y, q.z, c.x, c.y, c.z, cell);
            unsigned m = __ballot_sync(FULL, tv && lex_lt(D, jc, thr, thrj));
            if (R == 1 && __popc(m) >= KNN_MERGE_MIN) {
              // Many candidates beat the k-th distance (the first batches of a query, when the list is still filling): instead
              // of 25 instructions of serial insertion per candidate, sort the batch across the lanes (bitonic network over
              // the total order (distance, index); the others become +inf padding), merge it with the sorted list (reverse,
              // lane-wise minimum = the 32 smallest of the union as a bitonic sequence, five merge stages) and cut at k.
              // The list holds the k smallest of everything visited either way: same result, ~170 instructions per batch.
              float bd = (m >> lane) & 1u ? D : INF;
              int bj = (m >> lane) & 1u ? jc : 0x7fffffff;
#pragma unroll
              for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
                for (int j2 = k2 >> 1; j2 > 0; j2 >>= 1) {
                  const float pd = __shfl_xor_sync(FULL, bd, j2);
                  const int pj = __shfl_xor_sync(FULL, bj, j2);
                  const bool keep_min = (((lane & k2) == 0) == ((lane & j2) == 0));
                  const bool partner_less = lex_lt(pd, pj, bd, bj);
                  if (keep_min == partner_less) { bd = pd; bj = pj; }
                }
              }
              const float rd = __shfl_sync(FULL, bd, 31 - lane);
              const int rj = __shfl_sync(FULL, bj, 31 - lane);
              if (lex_lt(rd, rj, ed[0], ej[0])) { ed[0] = rd; ej[0] = rj; }
#pragma unroll
              for (int j2 = 16; j2 > 0; j2 >>= 1) {
                const float pd = __shfl_xor_sync(FULL, ed[0], j2);
                const int pj = __shfl_xor_sync(FULL, ej[0], j2);
                const bool keep_min = (lane & j2) == 0;
                const bool partner_less = lex_lt(pd, pj, ed[0], ej[0]);
                if (keep_min == partner_less) { ed[0] = pd; ej[0] = pj; }
              }
              if (lane >= k) { ed[0] = INF; ej[0] = 0x7fffffff; }
              thr = __shfl_sync(FULL, ed[0], kl);
              thrj = __shfl_sync(FULL, ej[0], kl);
              m = 0;
            }
            while (m) {
              const int srcl = __ffs(m) - 1;
              m &= m - 1;
              const float Dn = __shfl_sync(FULL, D, srcl);
              const int jn = __shfl_sync(FULL, jc, srcl);
              if (!lex_lt(Dn, jn, thr, thrj)) continue;    // threshold moved since the ballot (warp-uniform)
#pragma unroll
              for (int r = R - 1; r >= 0; --r) {
                float ld = __shfl_up_sync(FULL, ed[r], 1);
                int lj = __shfl_up_sync(FULL, ej[r], 1);
                if (r > 0) {
                  const float wd = __shfl_sync(FULL, ed[r - 1], 31);
                  const int wj = __shfl_sync(FULL, ej[r - 1], 31);
                  if (lane == 0) { ld = wd; lj = wj; }
                } else if (lane == 0) {
                  ld = -INF;
                  lj = -1;
                }
                if (r * 32 + lane < k) {
                  if (lex_lt(Dn, jn, ld, lj)) { ed[r] = ld; ej[r] = lj; }
                  else if (lex_lt(Dn, jn, ed[r], ej[r])) { ed[r] = Dn; ej[r] = jn; }
                }
              }
              float tq = ed[0];
              int tj = ej[0];
              if (R > 1 && kr == 1) { tq = ed[R - 1]; tj = ej[R - 1]; }
              thr = __shfl_sync(FULL, tq, kl);
              thrj = __shfl_sync(FULL, tj, kl);
            }
          }
          len -= run;
          xa = 0;
        }
      }
    }
    if (span == G) break;                                  // the whole box was searched
    // accept iff sqrt(D_k) < min over axes of (R h - margin)
    float reach = INF;
#pragma unroll
    for (int c = 0; c < 3; ++c) reach = fminf(reach, (float)Rr * g.h[c] - g.margin[c]);
    if (thr < INF && reach > 0.f && thr < reach * reach) break;
  }
  const size_t ebase = ((size_t)b * N + qi) * k;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const int e = r * 32 + lane;
    if (e < k) {
      const int j = ej[r];
      senders[ebase + e] = qi;
      receivers[ebase + e] = j;
    }
  }
  (void)dr;   // displacements are written by knn_dr_kernel from the caller's coordinate array
}

// dr[e] = wrapped displacement of the selected pair, same expression as the dense path
template <bool PBC, bool DIAG>
__global__ void knn_dr_kernel(const float* __restrict__ x, int ldx, int N, int k, Cell cell,
                              const int32_t* __restrict__ receivers, float* __restrict__ dr) {
  const int b = blockIdx.y;
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= (int64_t)N * k) return;
  const int qi = (int)(e / k);
  const float* xb = x + (size_t)b * N * ldx;
  const int j = receivers[(size_t)b * N * k + e];
  float o[3] = {0.f, 0.f, 0.f};
  if (j >= 0 && j < N) {
    const float* q = xb + (size_t)qi * ldx;
    const float* p = xb + (size_t)j * ldx;
    pair_dist<PBC, DIAG, false>(q[0], q[1], q[2], p[0], p[1], p[2], cell, o);
  }
  float* out = dr + ((size_t)b * N * k + e) * 3;
  out[0] = o[0]; out[1] = o[1]; out[2] = o[2];
}

// cells per axis for a uniform-density estimate of the k-th neighbour distance r_k = L (3k / (4 pi N))^(1/3):
// cell width >= 1.3 r_k so that R = 1 is accepted for almost every query.  0 = use the dense sweep.
int grid_cells(int N, int k) {
  if (N < 2048) return 0;
  const double rk = cbrt(3.0 * k / (4.0 * 3.14159265358979323846 * N));
  int G = (int)floor(1.0 / (1.3 * rk));
  if (G > 16) G = 16;
  return G >= 4 ? G : 0;
}

size_t cell_ws_bytes(int B, int N, int G) {
  const size_t n_cells = (size_t)G * G * G;
  return (size_t)B * N * 16 + (size_t)B * N * 4 + (size_t)B * (n_cells + 1) * 4 + (size_t)B * n_cells * 4 + 256;
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

// ---- radius graph (the radius branch of build_graph, graph_utils.py:266-277, as it evidently means to work) ----
// edges (i -> j) of one graph: 0 < |apply_pbc(x_i - x_j)| < radius, in np.where order (i ascending, then j ascending);
// the distance is compute_distances' sqrt(sum(dr^2)) (:209-226), every fp32 operation rounded separately.
// One warp per query point i; FILL = false counts the row, FILL = true writes it behind the row pointer.
template <bool PBC, bool DIAG, bool FILL>
__global__ void __launch_bounds__(256)
radius_kernel(const float* __restrict__ x, int ldx, int N, Cell cell, float radius, int32_t* __restrict__ count,
              const int32_t* __restrict__ rowptr, int32_t* __restrict__ senders, int32_t* __restrict__ receivers,
              float* __restrict__ dist) {
  const unsigned FULL = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const int b = blockIdx.y;
  const int i = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (i >= N) return;
  const float* xb = x + (size_t)b * N * ldx;
  const float xi = xb[(size_t)i * ldx], yi = xb[(size_t)i * ldx + 1], zi = xb[(size_t)i * ldx + 2];
  int n = 0;
  int64_t base = FILL ? rowptr[(size_t)b * N + i] : 0;
  for (int j0 = 0; j0 < N; j0 += 32) {
    const int j = j0 + lane;
    bool hit = false;
    float d = 0.f;
    if (j < N) {
      const float* pj = xb + (size_t)j * ldx;
      d = __fsqrt_rn(pair_dist<PBC, DIAG, false, false>(xi, yi, zi, pj[0], pj[1], pj[2], cell));
      hit = d < radius && d > 0.f;
    }
    const unsigned m = __ballot_sync(FULL, hit);
    if (FILL && hit) {
      const int64_t o = base + __popc(m & ((1u << lane) - 1u));
      senders[o] = i;
      receivers[o] = j;
      dist[o] = d;
    }
    n += __popc(m);
    base += __popc(m);
  }
  if (!FILL && lane == 0) count[(size_t)b * N + i] = n;
}

}  // namespace

static int radius_cell(const float* cell, const float* cell_inv, int flags, Cell& c, bool& pbc, bool& diag) {
  pbc = (flags & EQNN_KNN_PBC) != 0;
  diag = true;
  for (int i = 0; i < 9; ++i) {
    c.c[i] = (i % 4 == 0) ? 1.f : 0.f;
    c.ci[i] = c.c[i];
  }
  if (pbc) {
    EQNN_CHECK_ARG(cell && cell_inv, "radius_graph: PBC requested without cell / cell_inv");
    for (int i = 0; i < 9; ++i) {
      c.c[i] = cell[i];
      c.ci[i] = cell_inv[i];
      if (i % 4 != 0 && (cell[i] != 0.f || cell_inv[i] != 0.f)) diag = false;
    }
  }
  return EQNN_OK;
}

extern "C" int eqnn_radius_graph(const float* x, int ld_x, const float* cell, const float* cell_inv, int B, int N,
                                 float radius, int flags, int32_t* count, const int32_t* rowptr, int32_t* senders,
                                 int32_t* receivers, float* dist, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && B >= 0 && N >= 0 && ld_x >= 3, "radius_graph: bad argument");
  const bool fill = rowptr != nullptr;
  EQNN_CHECK_ARG(fill ? (senders && receivers && dist) : (count != nullptr), "radius_graph: null output");
  if (B == 0 || N == 0) return EQNN_OK;
  Cell c;
  bool pbc, diag;
  if (int rc = radius_cell(cell, cell_inv, flags, c, pbc, diag)) return rc;
  dim3 grid((unsigned)(((int64_t)N * 32 + 255) / 256), B);
  cudaStream_t st = as_stream(stream);
#define EQNN_RADIUS(P, D)                                                                                            \
  do {                                                                                                               \
    if (fill) radius_kernel<P, D, true><<<grid, 256, 0, st>>>(x, ld_x, N, c, radius, count, rowptr, senders, receivers, dist); \
    else radius_kernel<P, D, false><<<grid, 256, 0, st>>>(x, ld_x, N, c, radius, count, rowptr, senders, receivers, dist);     \
  } while (0)
  if (!pbc) EQNN_RADIUS(false, true);
  else if (diag) EQNN_RADIUS(true, true);
  else EQNN_RADIUS(true, false);
#undef EQNN_RADIUS
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" size_t eqnn_knn_workspace_bytes(int B, int N, int k) {
  const int G = grid_cells(N, k);
  return G ? cell_ws_bytes(B, N, G) : 0;
}

extern "C" int eqnn_knn_pbc(const float* x, int ld_x, const float* cell, const float* cell_inv, int B, int N,
                            int k, int flags, int32_t* senders, int32_t* receivers, float* dr, void* ws, size_t ws_bytes,
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
  // cell-list search: periodic diagonal cells, large clouds, and a caller-provided workspace; bit-identical output
  const int G = (pbc && diag && !(flags & EQNN_KNN_DENSE)) ? grid_cells(N, k) : 0;
  if (G && ws && ws_bytes >= cell_ws_bytes(B, N, G)) {
    const int n_cells = G * G * G;
    Grid g;
    g.G = G;
    for (int a = 0; a < 3; ++a) {
      const float L = fabsf(c.c[4 * a]);
      g.h[a] = L / (float)G;
      g.margin[a] = 1e-4f * L + 4e-7f;
    }
    uint8_t* w = reinterpret_cast<uint8_t*>(ws);
    float4* sorted = reinterpret_cast<float4*>(w);
    int32_t* cell_of = reinterpret_cast<int32_t*>(w + (size_t)B * N * 16);
    int32_t* start = cell_of + (size_t)B * N;
    int32_t* cursor = start + (size_t)B * (n_cells + 1);
    EQNN_CUDA(cudaMemsetAsync(start, 0, ((size_t)B * (n_cells + 1) + (size_t)B * n_cells) * 4, st));
    dim3 gp((N + 255) / 256, B);
    knn_bin_kernel<<<gp, 256, 0, st>>>(x, ld_x, N, c, g, cell_of, start);
    knn_scan_kernel<<<B, 1024, 0, st>>>(start, n_cells);
    knn_scatter_kernel<<<gp, 256, 0, st>>>(x, ld_x, N, n_cells, cell_of, start, cursor, sorted);
    dim3 gq((N + WARPS - 1) / WARPS, B);
    if (k <= 32) {
      if (fma) knn_cell_kernel<1, true><<<gq, WARPS * 32, 0, st>>>(sorted, start, N, k, c, g, senders, receivers, dr);
      else knn_cell_kernel<1, false><<<gq, WARPS * 32, 0, st>>>(sorted, start, N, k, c, g, senders, receivers, dr);
    } else {
      if (fma) knn_cell_kernel<2, true><<<gq, WARPS * 32, 0, st>>>(sorted, start, N, k, c, g, senders, receivers, dr);
      else knn_cell_kernel<2, false><<<gq, WARPS * 32, 0, st>>>(sorted, start, N, k, c, g, senders, receivers, dr);
    }
    int n_launch = 4;
    if (dr) {
      dim3 gd((unsigned)(((int64_t)N * k + 255) / 256), B);
      knn_dr_kernel<true, true><<<gd, 256, 0, st>>>(x, ld_x, N, k, c, receivers, dr);
      ++n_launch;
    }
    EQNN_CHECK_LAUNCH_N(n_launch);
    return EQNN_OK;
  }
  if (k <= 32) return launch_knn<1>(x, ld_x, N, k, c, pbc, diag, fma, B, senders, receivers, dr, st);
  return launch_knn<2>(x, ld_x, N, k, c, pbc, diag, fma, B, senders, receivers, dr, st);
}
