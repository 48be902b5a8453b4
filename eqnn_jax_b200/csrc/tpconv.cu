// NequIP interaction block: spherical harmonics + Clebsch-Gordan tensor product +
// per-irrep radial weights + receiver-side reduction, fused (sm_100a).
//
// Reference arithmetic: models/nequip.py:46-52 (Y_lm, tensor_product), :68
// (W_R * m), jraph segment_sum / segment_mean (:118-120).  One CTA owns ROWS
// consecutive receiver rows of the receiver-sorted CSR:
//   phase 1  one lane per incoming edge: coalesced stream of (src, r_hat, w_t columns),
//            32-byte gather of the sender record, Y_lm and the sparse CG contraction in
//            registers (generated straight-line code, csrc/gen/tp_gen.cuh), message
//            written to shared memory (row stride odd => conflict-free);
//   phase 2  one thread per (row, component): sequential sum over the row's edges in
//            ascending edge id -- no atomics, deterministic, and the same order as the
//            CPU scatter-add of the reference.
// The backward kernel mirrors it: dL/dA rows staged in shared memory, one lane per
// edge produces dL/dw_t (coalesced column-major store) and the 32-byte dL/dxt record
// stored at the original edge id (one full sector per edge).
#include "common.cuh"
#include <stdlib.h>
#include "tc_common.cuh"
#include "gen/tp_gen.cuh"

namespace {

#ifndef TPC_THREADS_N
#define TPC_THREADS_N 128
#endif
#ifndef TPC_ROWS_N
#define TPC_ROWS_N 6
#endif
constexpr int TPC_THREADS = TPC_THREADS_N;
constexpr int TPC_ROWS = TPC_ROWS_N;
#ifndef TPC_BWD_MIN_CTAS
#define TPC_BWD_MIN_CTAS 12
#endif
#ifndef TPC_FWD_MIN_CTAS
#define TPC_FWD_MIN_CTAS 12      // 40 registers, no spills (16 CTAs spill and measured 101 us against 78; without a bound: 48 registers, 10 CTAs / SM)
#endif

// 32-byte records (sender record xt, shared radial-weight record, per-slot dL/dw and dL/dxt records) move with ONE 256-bit
// access per lane: the interaction kernels are bound by the L1 data pipe (81 % of its wavefront peak in tpconv_fwd,
// profiles/r2h_tpconv_node_raw.csv), and two 128-bit accesses to the same 32-byte sector cost two wavefronts.
// (Buffers are 32-byte aligned: checked on the host.)
template <class TP>
__device__ __forceinline__ void load_x(const float* __restrict__ xt, int s, float* x) {
  if constexpr (TP::DXP == 8) {
    tc::ldg256(xt + (size_t)s * 8, x);
  } else {
    const float4* p = reinterpret_cast<const float4*>(xt + (size_t)s * TP::DXP);
#pragma unroll
    for (int i = 0; i < TP::DXP / 4; ++i) {
      float4 v = __ldg(p + i);
      x[4 * i + 0] = v.x;
      x[4 * i + 1] = v.y;
      x[4 * i + 2] = v.z;
      x[4 * i + 3] = v.w;
    }
  }
}
// NLIVE values of the row-major record `r` (row stride NLIVE rounded up to 4)
template <class TP>
__device__ __forceinline__ void load_w_record(const float* __restrict__ w_rows, int r, float* w) {
  constexpr int LDR = (TP::NLIVE + 3) / 4 * 4;
  if constexpr (LDR == 8) {
    float v[8];
    tc::ldg256(w_rows + (size_t)r * 8, v);
#pragma unroll
    for (int c = 0; c < TP::NLIVE; ++c) w[c] = v[c];
  } else {
    const float4* q = reinterpret_cast<const float4*>(w_rows + (size_t)r * LDR);
#pragma unroll
    for (int i = 0; i < LDR / 4; ++i) {
      const float4 v = __ldg(q + i);
      if (4 * i + 0 < TP::NLIVE) w[4 * i + 0] = v.x;
      if (4 * i + 1 < TP::NLIVE) w[4 * i + 1] = v.y;
      if (4 * i + 2 < TP::NLIVE) w[4 * i + 2] = v.z;
      if (4 * i + 3 < TP::NLIVE) w[4 * i + 3] = v.w;
    }
  }
}
template <class TP>
__device__ __forceinline__ void store_dw_record(float* __restrict__ dw_rows, int p, const float* dw) {
  constexpr int LDR = (TP::NLIVE + 3) / 4 * 4;
  float dwp[LDR];
#pragma unroll
  for (int c = 0; c < LDR; ++c) dwp[c] = c < TP::NLIVE ? dw[c < TP::NLIVE ? c : 0] : 0.f;
  if constexpr (LDR == 8) {
    tc::stg256(dw_rows + (size_t)p * 8, dwp);
  } else {
    float4* dr = reinterpret_cast<float4*>(dw_rows + (size_t)p * LDR);
#pragma unroll
    for (int i = 0; i < LDR / 4; ++i) dr[i] = make_float4(dwp[4 * i], dwp[4 * i + 1], dwp[4 * i + 2], dwp[4 * i + 3]);
  }
}
template <class TP>
__device__ __forceinline__ void store_dx_record(float* __restrict__ dxe, int e, const float* dx) {
  if constexpr (TP::DXP == 8) {
    tc::stg256(dxe + (size_t)e * 8, dx);
  } else {
    float4* o = reinterpret_cast<float4*>(dxe + (size_t)e * TP::DXP);
#pragma unroll
    for (int i = 0; i < TP::DXP / 4; ++i) o[i] = make_float4(dx[4 * i], dx[4 * i + 1], dx[4 * i + 2], dx[4 * i + 3]);
  }
}

// radial weights of CSR slot p: column-major w_t [NLIVE][n_edges], or -- shared radial evaluations (pairs.cu) -- the
// row-major record of the slot's unique evaluation, w_rows [n_unique][LDR], LDR = NLIVE rounded up to 4 (one or two
// 128-bit loads behind the uidx load, which sits beside the src -> xt gather in the dependency chain)
template <class TP>
__device__ __forceinline__ void load_w(const float* __restrict__ w_t, int64_t n_edges, const int32_t* __restrict__ uidx,
                                       const float* __restrict__ w_rows, int p, float* w) {
  if (uidx) {
    load_w_record<TP>(w_rows, __ldg(uidx + p), w);
  } else {
#pragma unroll
    for (int c = 0; c < TP::NLIVE; ++c) w[c] = w_t[(size_t)c * n_edges + p];
  }
}

// AMAX: jraph.segment_max (message_passing_agg = "max", models/nequip.py:118-120): the running value is the maximum
// over the row's edges and the CSR position of the FIRST edge that attains it is recorded, so that the backward
// pass can route the gradient to that edge alone (ties are measure-zero for real messages; the self-loop's exact
// zeros in the l > 0 components can tie only with another exact zero).
// message row stride in floats: a multiple of 4 with an odd number of 16-byte groups, so that the 128-bit stores of 8
// consecutive lanes (one per edge) fall into 8 distinct bank groups
__host__ __device__ constexpr int tpc_stride(int dl) {
  int s = (dl + 3) / 4 * 4;
  return ((s / 4) & 1) ? s : s + 4;
}

template <class TP, bool AMAX>
__global__ void __launch_bounds__(TPC_THREADS, TPC_FWD_MIN_CTAS)
tpconv_fwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                  const float4* __restrict__ geom, const float* __restrict__ w_t, int64_t n_edges,
                  const float* __restrict__ xt, int n_nodes, int agg, float* __restrict__ A, int ld_a,
                  int32_t* __restrict__ argmax, const int32_t* __restrict__ uidx, const float* __restrict__ w_rows) {
  constexpr int DL = TP::DLIVE;
  constexpr int S = tpc_stride(DL), S4 = S / 4;
  constexpr int NQ = (DL + 3) / 4;                      // component quads per row
  constexpr int ITEMS = (TPC_ROWS * NQ + TPC_THREADS - 1) / TPC_THREADS;
  __shared__ __align__(16) float sm[TPC_THREADS * S];
  __shared__ int s_rowptr[TPC_ROWS + 1];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TPC_ROWS;
  const int nr = min(TPC_ROWS, n_nodes - r0);
  if (tid <= nr) s_rowptr[tid] = rowptr[r0 + tid];
  __syncthreads();
  const int p0 = s_rowptr[0], p1 = s_rowptr[nr];
  float4 acc[ITEMS];
  int4 amax[ITEMS];
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const float v0 = AMAX ? -INFINITY : 0.f;
    acc[it] = make_float4(v0, v0, v0, v0);
    amax[it] = make_int4(-1, -1, -1, -1);
  }

  for (int base = p0; base < p1; base += TPC_THREADS) {
    const int p = base + tid;
    if (p < p1) {
      const int s = src[p];
      const float4 g = geom[p];
      float x[TP::DXP], w[TP::NLIVE], Y[TP::NY];
      __align__(16) float m[S];
      load_x<TP>(xt, s, x);
      load_w<TP>(w_t, n_edges, uidx, w_rows, p, w);
      TP::sh(g.x, g.y, g.z, Y);
#pragma unroll
      for (int i = DL; i < S; ++i) m[i] = 0.f;
      TP::fwd(x, Y, w, m);
      float4* dst = reinterpret_cast<float4*>(sm + tid * S);
#pragma unroll
      for (int i = 0; i < NQ; ++i) dst[i] = make_float4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]);
    }
    __syncthreads();
#pragma unroll
    for (int it = 0; it < ITEMS; ++it) {
      const int item = tid + it * TPC_THREADS;
      if (item < nr * NQ) {
        const int row = item / NQ, quad = item - row * NQ;
        const int lo = max(s_rowptr[row], base) - base;
        const int hi = min(s_rowptr[row + 1], base + TPC_THREADS) - base;
        float4 a = acc[it];
        const float4* col = reinterpret_cast<const float4*>(sm + 4 * quad);
        int q = lo;
        if (AMAX) {
          int4 am = amax[it];
          for (; q < hi; ++q) {
            const float4 v = col[q * S4];
            if (v.x > a.x) { a.x = v.x; am.x = base + q; }
            if (v.y > a.y) { a.y = v.y; am.y = base + q; }
            if (v.z > a.z) { a.z = v.z; am.z = base + q; }
            if (v.w > a.w) { a.w = v.w; am.w = base + q; }
          }
          amax[it] = am;
        }
        for (; q + 1 < hi; q += 2) {                   // ascending order: the order of a sequential scatter-add
          const float4 v0 = col[q * S4], v1 = col[(q + 1) * S4];
          a.x = (a.x + v0.x) + v1.x;
          a.y = (a.y + v0.y) + v1.y;
          a.z = (a.z + v0.z) + v1.z;
          a.w = (a.w + v0.w) + v1.w;
        }
        if (q < hi) {
          const float4 v = col[q * S4];
          a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
        }
        acc[it] = a;
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int it = 0; it < ITEMS; ++it) {
    const int item = tid + it * TPC_THREADS;
    if (item < nr * NQ) {
      const int row = item / NQ, quad = item - row * NQ;
      const float a[4] = {acc[it].x, acc[it].y, acc[it].z, acc[it].w};
      const int am[4] = {amax[it].x, amax[it].y, amax[it].z, amax[it].w};
      const float deg = (float)max(s_rowptr[row + 1] - s_rowptr[row], 1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int comp = 4 * quad + j;
        if (comp < DL) {
          float v = a[j];
          if (agg == 1) v = v / deg;
          if (AMAX) {
            if (am[j] < 0) v = 0.f;                    // receiver without incoming edges (cannot happen with self-loops)
            argmax[(size_t)(r0 + row) * DL + comp] = am[j];
          }
          A[(size_t)(r0 + row) * ld_a + comp] = v;
        }
      }
    }
  }
}

template <class TP, bool AMAX>
__global__ void __launch_bounds__(TPC_THREADS, (TP::DLIVE <= 24 ? TPC_BWD_MIN_CTAS : 8))
tpconv_bwd_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                  const int32_t* __restrict__ perm, const float4* __restrict__ geom,
                  const float* __restrict__ w_t, int64_t n_edges, const float* __restrict__ xt, int n_nodes,
                  int agg, const float* __restrict__ gA, int ld_ga, float* __restrict__ dw_t,
                  float* __restrict__ dxe, const int32_t* __restrict__ argmax, const int32_t* __restrict__ uidx,
                  const float* __restrict__ w_rows, float* __restrict__ dw_rows) {
  constexpr int DL = TP::DLIVE;
  constexpr int GS = DL | 1;
  __shared__ float sG[TPC_ROWS * GS];
  __shared__ int sArg[AMAX ? TPC_ROWS * GS : 1];
  __shared__ int s_rowptr[TPC_ROWS + 1];
  const int tid = threadIdx.x;
  const int r0 = blockIdx.x * TPC_ROWS;
  const int nr = min(TPC_ROWS, n_nodes - r0);
  if (tid <= nr) s_rowptr[tid] = rowptr[r0 + tid];
  __syncthreads();
  for (int item = tid; item < nr * DL; item += TPC_THREADS) {
    const int row = item / DL, comp = item - row * DL;
    float g = gA[(size_t)(r0 + row) * ld_ga + comp];
    if (agg == 1) {
      const int deg = s_rowptr[row + 1] - s_rowptr[row];
      g = g / (float)max(deg, 1);
    }
    sG[row * GS + comp] = g;
    if (AMAX) sArg[row * GS + comp] = argmax[(size_t)(r0 + row) * DL + comp];
  }
  __syncthreads();
  const int p0 = s_rowptr[0], p1 = s_rowptr[nr];
  for (int p = p0 + tid; p < p1; p += TPC_THREADS) {
    int row = 0;
#pragma unroll
    for (int r = 1; r < TPC_ROWS; ++r) row += (r < nr && p >= s_rowptr[r]) ? 1 : 0;
    const int s = src[p];
    const int e = perm[p];
    const float4 gm = geom[p];
    float x[TP::DXP], w[TP::NLIVE], Y[TP::NY], g[DL], dw[TP::NLIVE], dx[TP::DXP];
    load_x<TP>(xt, s, x);
    load_w<TP>(w_t, n_edges, uidx, w_rows, p, w);
#pragma unroll
    for (int i = 0; i < DL; ++i) g[i] = (!AMAX || sArg[row * GS + i] == p) ? sG[row * GS + i] : 0.f;
    TP::sh(gm.x, gm.y, gm.z, Y);
    TP::bwd(x, Y, w, g, dw, dx);
    if (dw_rows) {                           // shared radial evaluations: one row-major record per slot (pair_combine sums the pair)
      store_dw_record<TP>(dw_rows, p, dw);
    } else {
#pragma unroll
      for (int c = 0; c < TP::NLIVE; ++c) dw_t[(size_t)c * n_edges + p] = dw[c];
    }
    store_dx_record<TP>(dxe, e, dx);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Warp-autonomous variants (round 2, the default): a WARP owns RW consecutive receiver rows and walks their edges in passes
// of 32 -- lane per edge, message into the warp's own 32-row shared-memory tile, `__syncwarp`, (row, component quad) lanes
// add the pass's rows in ascending edge order, `__syncwarp`.  No block barrier anywhere: in the CTA-cooperative kernels above
// the three barriers of a 6-row CTA were the second largest stall (8.7 warps per issue, behind the load chain's 11.6;
// profiles/r2h_tpconv_node_raw.csv), because every warp waits for the slowest warp's dependent loads twice.  The indices of
// the NEXT pass (sender, record index, geometry) are loaded before the current pass's arithmetic.  Measured: 78 us forward
// against 80, the same 90 us backward -- the barrier stalls were a symptom of the load chain, not its cause; what did move
// the backward kernel (116 -> 90 us) was one 256-bit access per 32-byte record instead of two 128-bit ones.
// Same per-row summation order (ascending CSR position), so the results are bit-identical to the kernels above.
constexpr int TPW_WARPS = 4;
template <class TP>
struct TpwCfg {
  static constexpr int DL = TP::DLIVE;
  static constexpr int S = tpc_stride(DL), S4 = S / 4;
  static constexpr int NQ = (DL + 3) / 4;
  static constexpr int RW = (32 / NQ) < 6 ? (32 / NQ) : 6;          // rows per warp: RW * NQ reduce lanes <= 32
  static constexpr bool OK = NQ <= 32 && S * 32 * TPW_WARPS * 4 <= 40 * 1024;
};

template <class TP, bool AMAX>
__global__ void __launch_bounds__(TPW_WARPS * 32)
tpconv_fwd_warp_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const float4* __restrict__ geom,
                       const float* __restrict__ w_t, int64_t n_edges, const float* __restrict__ xt, int n_nodes, int agg,
                       float* __restrict__ A, int ld_a, int32_t* __restrict__ argmax, const int32_t* __restrict__ uidx,
                       const float* __restrict__ w_rows) {
  using C = TpwCfg<TP>;
  constexpr int DL = C::DL, S = C::S, S4 = C::S4, NQ = C::NQ, RW = C::RW;
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) float sm_all[TPW_WARPS * 32 * S];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sm = sm_all + warp * 32 * S;
  const int r0 = (blockIdx.x * TPW_WARPS + warp) * RW;
  const int nr = min(RW, n_nodes - r0);
  if (nr <= 0) return;
  const int rp = (lane <= nr) ? __ldg(rowptr + r0 + lane) : 0;
  const int p_beg = __shfl_sync(FULL, rp, 0), p_end = __shfl_sync(FULL, rp, nr);
  const int row_i = lane / NQ, quad = lane - row_i * NQ;
  const bool active = lane < nr * NQ;
  const int rlo = __shfl_sync(FULL, rp, active ? row_i : 0), rhi = __shfl_sync(FULL, rp, active ? row_i + 1 : 0);
  const float v0 = AMAX ? -INFINITY : 0.f;
  float4 acc = make_float4(v0, v0, v0, v0);
  int4 amax = make_int4(-1, -1, -1, -1);
  // pass 0 indices
  int p = p_beg + lane;
  int c_s = 0, c_u = 0;
  float4 c_g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p < p_end) {
    c_s = __ldg(src + p);
    if (uidx) c_u = __ldg(uidx + p);
    c_g = __ldg(geom + p);
  }
  for (int base = p_beg; base < p_end; base += 32) {
    const int pn = base + 32 + lane;
    int n_s = 0, n_u = 0;
    float4 n_g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pn < p_end) {                                      // next pass: in flight across this pass's arithmetic
      n_s = __ldg(src + pn);
      if (uidx) n_u = __ldg(uidx + pn);
      n_g = __ldg(geom + pn);
    }
    p = base + lane;
    if (p < p_end) {
      float x[TP::DXP], w[TP::NLIVE], Y[TP::NY];
      __align__(16) float m[S];
      load_x<TP>(xt, c_s, x);
      if (uidx) {
        load_w_record<TP>(w_rows, c_u, w);
      } else {
#pragma unroll
        for (int c = 0; c < TP::NLIVE; ++c) w[c] = w_t[(size_t)c * n_edges + p];
      }
      TP::sh(c_g.x, c_g.y, c_g.z, Y);
#pragma unroll
      for (int i = DL; i < S; ++i) m[i] = 0.f;
      TP::fwd(x, Y, w, m);
      float4* dst = reinterpret_cast<float4*>(sm + lane * S);
#pragma unroll
      for (int i = 0; i < NQ; ++i) dst[i] = make_float4(m[4 * i], m[4 * i + 1], m[4 * i + 2], m[4 * i + 3]);
    }
    __syncwarp();
    if (active) {
      const int lo = max(rlo, base) - base, hi = min(rhi, base + 32) - base;
      const float4* col = reinterpret_cast<const float4*>(sm + 4 * quad);
      int q = lo;
      if (AMAX) {
        for (; q < hi; ++q) {
          const float4 v = col[q * S4];
          if (v.x > acc.x) { acc.x = v.x; amax.x = base + q; }
          if (v.y > acc.y) { acc.y = v.y; amax.y = base + q; }
          if (v.z > acc.z) { acc.z = v.z; amax.z = base + q; }
          if (v.w > acc.w) { acc.w = v.w; amax.w = base + q; }
        }
      }
      for (; q + 3 < hi; q += 4) {                         // ascending order: the order of a sequential scatter-add
        const float4 a0 = col[q * S4], a1 = col[(q + 1) * S4], a2 = col[(q + 2) * S4], a3 = col[(q + 3) * S4];
        acc.x = (((acc.x + a0.x) + a1.x) + a2.x) + a3.x;   // four loads in flight, the adds stay sequential
        acc.y = (((acc.y + a0.y) + a1.y) + a2.y) + a3.y;
        acc.z = (((acc.z + a0.z) + a1.z) + a2.z) + a3.z;
        acc.w = (((acc.w + a0.w) + a1.w) + a2.w) + a3.w;
      }
      for (; q + 1 < hi; q += 2) {
        const float4 a0 = col[q * S4], a1 = col[(q + 1) * S4];
        acc.x = (acc.x + a0.x) + a1.x;
        acc.y = (acc.y + a0.y) + a1.y;
        acc.z = (acc.z + a0.z) + a1.z;
        acc.w = (acc.w + a0.w) + a1.w;
      }
      if (q < hi) {
        const float4 a0 = col[q * S4];
        acc.x += a0.x; acc.y += a0.y; acc.z += a0.z; acc.w += a0.w;
      }
    }
    __syncwarp();
    c_s = n_s; c_u = n_u; c_g = n_g;
  }
  if (active) {
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
    const int am[4] = {amax.x, amax.y, amax.z, amax.w};
    const float deg = (float)max(rhi - rlo, 1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int comp = 4 * quad + j;
      if (comp < DL) {
        float v = a[j];
        if (agg == 1) v = v / deg;
        if (AMAX) {
          if (am[j] < 0) v = 0.f;
          argmax[(size_t)(r0 + row_i) * DL + comp] = am[j];
        }
        A[(size_t)(r0 + row_i) * ld_a + comp] = v;
      }
    }
  }
}

template <class TP, bool AMAX>
__global__ void __launch_bounds__(TPW_WARPS * 32)
tpconv_bwd_warp_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src, const int32_t* __restrict__ perm,
                       const float4* __restrict__ geom, const float* __restrict__ w_t, int64_t n_edges,
                       const float* __restrict__ xt, int n_nodes, int agg, const float* __restrict__ gA, int ld_ga,
                       float* __restrict__ dw_t, float* __restrict__ dxe, const int32_t* __restrict__ argmax,
                       const int32_t* __restrict__ uidx, const float* __restrict__ w_rows, float* __restrict__ dw_rows) {
  using C = TpwCfg<TP>;
  constexpr int DL = C::DL, RW = C::RW;
  constexpr int GS = (DL + 3) / 4 * 4;                       // dL/dA rows padded to whole 128-bit groups
  constexpr unsigned FULL = 0xffffffffu;
  __shared__ __align__(16) float sG_all[TPW_WARPS * RW * GS];
  __shared__ int sArg_all[AMAX ? TPW_WARPS * RW * GS : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* sG = sG_all + warp * RW * GS;
  int* sArg = sArg_all + (AMAX ? warp * RW * GS : 0);
  const int r0 = (blockIdx.x * TPW_WARPS + warp) * RW;
  const int nr = min(RW, n_nodes - r0);
  if (nr <= 0) return;
  const int rp = (lane <= nr) ? __ldg(rowptr + r0 + lane) : 0x7fffffff;
  const int p_beg = __shfl_sync(FULL, rp, 0), p_end = __shfl_sync(FULL, rp, nr);
  int rps[RW];                                               // row starts 1 .. RW - 1 (0x7fffffff past the last row)
#pragma unroll
  for (int r = 1; r < RW; ++r) {
    const int v = __shfl_sync(FULL, rp, r);
    rps[r] = r < nr ? v : 0x7fffffff;
  }
  for (int item = lane; item < nr * GS; item += 32) {
    const int row = item / GS, comp = item - row * GS;
    float g = 0.f;
    if (comp < DL) {
      g = __ldg(gA + (size_t)(r0 + row) * ld_ga + comp);
      if (agg == 1) {
        const int deg = __ldg(rowptr + r0 + row + 1) - __ldg(rowptr + r0 + row);
        g = g / (float)max(deg, 1);
      }
      if (AMAX) sArg[item] = __ldg(argmax + (size_t)(r0 + row) * DL + comp);
    }
    sG[item] = g;
  }
  __syncwarp();
  int p = p_beg + lane;
  int c_s = 0, c_u = 0, c_e = 0;
  float4 c_g = make_float4(0.f, 0.f, 0.f, 0.f);
  if (p < p_end) {
    c_s = __ldg(src + p);
    c_e = __ldg(perm + p);
    if (uidx) c_u = __ldg(uidx + p);
    c_g = __ldg(geom + p);
  }
  for (int base = p_beg; base < p_end; base += 32) {
    const int pn = base + 32 + lane;
    int n_s = 0, n_u = 0, n_e = 0;
    float4 n_g = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pn < p_end) {
      n_s = __ldg(src + pn);
      n_e = __ldg(perm + pn);
      if (uidx) n_u = __ldg(uidx + pn);
      n_g = __ldg(geom + pn);
    }
    p = base + lane;
    if (p < p_end) {
      int row = 0;
#pragma unroll
      for (int r = 1; r < RW; ++r) row += (p >= rps[r]) ? 1 : 0;
      float x[TP::DXP], w[TP::NLIVE], Y[TP::NY], g[GS], dw[TP::NLIVE], dx[TP::DXP];
      load_x<TP>(xt, c_s, x);
      if (uidx) {
        load_w_record<TP>(w_rows, c_u, w);
      } else {
#pragma unroll
        for (int c = 0; c < TP::NLIVE; ++c) w[c] = w_t[(size_t)c * n_edges + p];
      }
      const float4* gr = reinterpret_cast<const float4*>(sG + row * GS);
#pragma unroll
      for (int i = 0; i < GS / 4; ++i) {
        const float4 v = gr[i];
        g[4 * i] = v.x; g[4 * i + 1] = v.y; g[4 * i + 2] = v.z; g[4 * i + 3] = v.w;
      }
      if (AMAX) {
#pragma unroll
        for (int i = 0; i < DL; ++i) g[i] = (sArg[row * GS + i] == p) ? g[i] : 0.f;
      }
      TP::sh(c_g.x, c_g.y, c_g.z, Y);
      TP::bwd(x, Y, w, g, dw, dx);
      if (dw_rows) {
        store_dw_record<TP>(dw_rows, p, dw);
      } else {
#pragma unroll
        for (int c = 0; c < TP::NLIVE; ++c) dw_t[(size_t)c * n_edges + p] = dw[c];
      }
      store_dx_record<TP>(dxe, c_e, dx);
    }
    c_s = n_s; c_u = n_u; c_e = n_e; c_g = n_g;
  }
}

__global__ void sender_reduce_kernel(const float* __restrict__ dxe, int n_nodes, int k, int dxp,
                                     float* __restrict__ dxt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * dxp) return;
  const int64_t s = i / dxp;
  const int c = (int)(i - s * dxp);
  const float* p = dxe + (size_t)s * k * dxp + c;
  float a = 0.f;
  for (int j = 0; j < k; ++j) a += p[(size_t)j * dxp];
  dxt[i] = a;
}

// general out-degree: dxt[s, :] = sum_{q in row s of the sender-sorted CSR} dxe[perm_s[q], :]
__global__ void segment_gather_sum_kernel(const float* __restrict__ dxe, const int32_t* __restrict__ rowptr_s,
                                          const int32_t* __restrict__ perm_s, int n_nodes, int dxp,
                                          float* __restrict__ dxt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * dxp) return;
  const int64_t s = i / dxp;
  const int c = (int)(i - s * dxp);
  float a = 0.f;
  for (int q = rowptr_s[s]; q < rowptr_s[s + 1]; ++q) a += dxe[(size_t)perm_s[q] * dxp + c];
  dxt[i] = a;
}

// node_attrs[r] = scale * sum_{e->r} Y(r_hat_e): one warp per row, lane per edge, shuffle reduce
template <class SH>
__global__ void __launch_bounds__(256) sh_scatter_kernel(const int32_t* __restrict__ rowptr,
                                                         const float4* __restrict__ geom, int n_nodes, float scale,
                                                         float* __restrict__ attrs) {
  constexpr int NY = SH::NY;
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  float acc[NY];
#pragma unroll
  for (int i = 0; i < NY; ++i) acc[i] = 0.f;
  for (int p = p0 + lane; p < p1; p += 32) {
    const float4 g = geom[p];
    float Y[NY];
    SH::sh(g.x, g.y, g.z, Y);
#pragma unroll
    for (int i = 0; i < NY; ++i) acc[i] += Y[i];
  }
#pragma unroll
  for (int i = 0; i < NY; ++i) {
    float v = acc[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    acc[i] = v;
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NY; ++i) attrs[(size_t)row * NY + i] = acc[i] * scale;
  }
}

// Per-edge messages W_e * TP(xt[src], Y(r_hat_e)) written at the ORIGINAL edge id: the `edges` field the reference's
// node task returns (the last step's update_edge_fn output, models/nequip.py:154-171).  w_rows holds one row of radial-MLP
// outputs per CSR position; column c of the signature reads MLP column col[c] scaled by scale[c] (SH normalisation).
struct TpMsgCols {
  int col[32];
  float scale[32];
};
template <class TP>
__global__ void __launch_bounds__(128)
tp_messages_kernel(const int32_t* __restrict__ src, const int32_t* __restrict__ perm, const float4* __restrict__ geom,
                   const float* __restrict__ w_rows, int ld_w, TpMsgCols cols, const float* __restrict__ xt,
                   int64_t n_edges, float* __restrict__ out, int ld_out) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_edges) return;
  const float4 g = geom[p];
  float x[TP::DXP], w[TP::NLIVE], Y[TP::NY], m[TP::DLIVE];
  load_x<TP>(xt, src[p], x);
#pragma unroll
  for (int c = 0; c < TP::NLIVE; ++c) w[c] = w_rows[p * ld_w + cols.col[c]] * cols.scale[c];
  TP::sh(g.x, g.y, g.z, Y);
  TP::fwd(x, Y, w, m);
  float* o = out + (size_t)perm[p] * ld_out;
#pragma unroll
  for (int i = 0; i < TP::DLIVE; ++i) o[i] = m[i];
}

template <class TP>
__global__ void node_tp_fwd_kernel(const float* __restrict__ x, const float* __restrict__ y, int64_t n,
                                   float* __restrict__ T) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xr[TP::DXP], Y[TP::NY], w[TP::NLIVE], m[TP::DLIVE];
  load_x<TP>(x, (int)i, xr);
#pragma unroll
  for (int j = 0; j < TP::NY; ++j) Y[j] = y[(size_t)i * TP::NY + j];
#pragma unroll
  for (int c = 0; c < TP::NLIVE; ++c) w[c] = 1.f;
  TP::fwd(xr, Y, w, m);
#pragma unroll
  for (int j = 0; j < TP::DLIVE; ++j) T[(size_t)i * TP::DLIVE + j] = m[j];
}

template <class TP>
__global__ void node_tp_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                   const float* __restrict__ gT, int64_t n, float* __restrict__ dxo) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float xr[TP::DXP], Y[TP::NY], w[TP::NLIVE], g[TP::DLIVE], dw[TP::NLIVE], dx[TP::DXP];
  load_x<TP>(x, (int)i, xr);
#pragma unroll
  for (int j = 0; j < TP::NY; ++j) Y[j] = y[(size_t)i * TP::NY + j];
#pragma unroll
  for (int c = 0; c < TP::NLIVE; ++c) w[c] = 1.f;
#pragma unroll
  for (int j = 0; j < TP::DLIVE; ++j) g[j] = gT[(size_t)i * TP::DLIVE + j];
  TP::bwd(xr, Y, w, g, dw, dx);
#pragma unroll
  for (int j = 0; j < TP::DXP; ++j) dxo[(size_t)i * TP::DXP + j] = dx[j];
}

}  // namespace

extern "C" int eqnn_tp_signature_count(void) { return EQNN_TP_COUNT; }
extern "C" const char* eqnn_tp_signature(int id) {
  return (id >= 0 && id < EQNN_TP_COUNT) ? EQNN_TP_SIGNATURES[id] : nullptr;
}
extern "C" int eqnn_tp_lookup(const char* sig) {
  if (!sig) return -1;
  for (int i = 0; i < EQNN_TP_COUNT; ++i)
    if (strcmp(sig, EQNN_TP_SIGNATURES[i]) == 0) return i;
  return -1;
}
extern "C" int eqnn_tp_dims(int id, int* dxp, int* ny, int* n_live, int* d_live) {
  EQNN_TP_DISPATCH(id, {
    if (dxp) *dxp = TP::DXP;
    if (ny) *ny = TP::NY;
    if (n_live) *n_live = TP::NLIVE;
    if (d_live) *d_live = TP::DLIVE;
  });
  return EQNN_OK;
}

static int tpconv_fwd_impl(int tp_id, const int32_t* rowptr, const int32_t* src, const float* geom, const float* w_t,
                           int64_t n_edges, const int32_t* uidx, const float* w_rows, int ld_rows, const float* xt,
                           int n_nodes, int agg, float* A, int ld_a, int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rowptr && src && geom && (w_t || (uidx && w_rows)) && xt && A, "tpconv_fwd: null pointer");
  EQNN_CHECK_ARG(agg >= 0 && agg <= 2, "tpconv_fwd: agg must be 0 (sum), 1 (mean) or 2 (max)");
  EQNN_CHECK_ARG(agg != 2 || argmax, "tpconv_fwd: max aggregation needs the argmax buffer");
  EQNN_CHECK_ARG(((uintptr_t)geom & 15) == 0 && ((uintptr_t)xt & 31) == 0, "tpconv_fwd: geom must be 16-byte, xt 32-byte aligned");
  EQNN_CHECK_ARG(!uidx || ((uintptr_t)w_rows & 31) == 0, "tpconv_fwd: w_rows must be 32-byte aligned");
  if (n_nodes <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)((n_nodes + TPC_ROWS - 1) / TPC_ROWS);
  EQNN_TP_DISPATCH(tp_id, {
    EQNN_CHECK_ARG(ld_a >= TP::DLIVE, "tpconv_fwd: ld_a smaller than live width");
    EQNN_CHECK_ARG(!uidx || ld_rows == (TP::NLIVE + 3) / 4 * 4, "tpconv_fwd: ld_rows must be NLIVE rounded up to 4");
    // the widest signatures (every message column at l_max 3: eqnn_tp_messages only) do not fit the message transpose
    if constexpr (TpwCfg<TP>::OK) {
      if (!getenv("EQNN_TPCONV_CTA")) {                       // default: the warp-autonomous kernel
        const unsigned wgrid = (unsigned)ceil_div64(n_nodes, TPW_WARPS * TpwCfg<TP>::RW);
        if (agg == 2)
          tpconv_fwd_warp_kernel<TP, true><<<wgrid, TPW_WARPS * 32, 0, as_stream(stream)>>>(
              rowptr, src, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, A, ld_a, argmax, uidx, w_rows);
        else
          tpconv_fwd_warp_kernel<TP, false><<<wgrid, TPW_WARPS * 32, 0, as_stream(stream)>>>(
              rowptr, src, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, A, ld_a, nullptr, uidx, w_rows);
        EQNN_CHECK_LAUNCH();
        return EQNN_OK;
      }
    }
    if constexpr (tpc_stride(TP::DLIVE) * TPC_THREADS * 4 <= 40 * 1024) {
      if (agg == 2)
        tpconv_fwd_kernel<TP, true><<<grid, TPC_THREADS, 0, as_stream(stream)>>>(
            rowptr, src, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, A, ld_a, argmax, uidx, w_rows);
      else
        tpconv_fwd_kernel<TP, false><<<grid, TPC_THREADS, 0, as_stream(stream)>>>(
            rowptr, src, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, A, ld_a, nullptr, uidx, w_rows);
    } else {
      return eqnn_fail(EQNN_ERR_ARG, "tpconv_fwd: this signature is wider than the reduction kernel covers");
    }
  });
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_tpconv_fwd(int tp_id, const int32_t* rowptr, const int32_t* src, const float* geom,
                               const float* w_t, int64_t n_edges, const float* xt, int n_nodes, int agg, float* A,
                               int ld_a, int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(w_t, "tpconv_fwd: null pointer");
  return tpconv_fwd_impl(tp_id, rowptr, src, geom, w_t, n_edges, nullptr, nullptr, 0, xt, n_nodes, agg, A, ld_a, argmax, stream);
}

extern "C" int eqnn_tpconv_fwd_shared(int tp_id, const int32_t* rowptr, const int32_t* src, const float* geom,
                                      const int32_t* uidx, const float* w_rows, int ld_rows, const float* xt, int n_nodes,
                                      int agg, float* A, int ld_a, int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(uidx && w_rows, "tpconv_fwd_shared: null pointer");
  return tpconv_fwd_impl(tp_id, rowptr, src, geom, nullptr, 0, uidx, w_rows, ld_rows, xt, n_nodes, agg, A, ld_a, argmax, stream);
}

static int tpconv_bwd_impl(int tp_id, const int32_t* rowptr, const int32_t* src, const int32_t* perm, const float* geom,
                           const float* w_t, int64_t n_edges, const int32_t* uidx, const float* w_rows, int ld_rows,
                           const float* xt, int n_nodes, int agg, const float* gA, int ld_ga, float* dw_t, float* dw_rows,
                           float* dxe, const int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rowptr && src && perm && geom && (w_t || (uidx && w_rows)) && xt && gA && (dw_t || dw_rows) && dxe,
                 "tpconv_bwd: null pointer");
  EQNN_CHECK_ARG(agg >= 0 && agg <= 2, "tpconv_bwd: agg must be 0 (sum), 1 (mean) or 2 (max)");
  EQNN_CHECK_ARG(agg != 2 || argmax, "tpconv_bwd: max aggregation needs the argmax buffer of the forward pass");
  EQNN_CHECK_ARG(((uintptr_t)geom & 15) == 0 && ((uintptr_t)xt & 31) == 0 && ((uintptr_t)dxe & 31) == 0,
                 "tpconv_bwd: geom must be 16-byte, xt/dxe 32-byte aligned");
  EQNN_CHECK_ARG(!uidx || (((uintptr_t)w_rows & 31) == 0 && ((uintptr_t)dw_rows & 31) == 0),
                 "tpconv_bwd: w_rows/dw_rows must be 32-byte aligned");
  if (n_nodes <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)((n_nodes + TPC_ROWS - 1) / TPC_ROWS);
  EQNN_TP_DISPATCH(tp_id, {
    EQNN_CHECK_ARG(ld_ga >= TP::DLIVE, "tpconv_bwd: ld_ga smaller than live width");
    EQNN_CHECK_ARG(!uidx || ld_rows == (TP::NLIVE + 3) / 4 * 4, "tpconv_bwd: ld_rows must be NLIVE rounded up to 4");
    if constexpr (TpwCfg<TP>::OK) {
      if (!getenv("EQNN_TPCONV_CTA")) {                       // default: the warp-autonomous kernel
        const unsigned wgrid = (unsigned)ceil_div64(n_nodes, TPW_WARPS * TpwCfg<TP>::RW);
        if (agg == 2)
          tpconv_bwd_warp_kernel<TP, true><<<wgrid, TPW_WARPS * 32, 0, as_stream(stream)>>>(
              rowptr, src, perm, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, gA, ld_ga, dw_t, dxe,
              argmax, uidx, w_rows, dw_rows);
        else
          tpconv_bwd_warp_kernel<TP, false><<<wgrid, TPW_WARPS * 32, 0, as_stream(stream)>>>(
              rowptr, src, perm, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, gA, ld_ga, dw_t, dxe,
              nullptr, uidx, w_rows, dw_rows);
        EQNN_CHECK_LAUNCH();
        return EQNN_OK;
      }
    }
    if (agg == 2)
      tpconv_bwd_kernel<TP, true><<<grid, TPC_THREADS, 0, as_stream(stream)>>>(
          rowptr, src, perm, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, gA, ld_ga, dw_t,
          dxe, argmax, uidx, w_rows, dw_rows);
    else
      tpconv_bwd_kernel<TP, false><<<grid, TPC_THREADS, 0, as_stream(stream)>>>(
          rowptr, src, perm, reinterpret_cast<const float4*>(geom), w_t, n_edges, xt, n_nodes, agg, gA, ld_ga, dw_t,
          dxe, nullptr, uidx, w_rows, dw_rows);
  });
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_tpconv_bwd(int tp_id, const int32_t* rowptr, const int32_t* src, const int32_t* perm,
                               const float* geom, const float* w_t, int64_t n_edges, const float* xt, int n_nodes,
                               int agg, const float* gA, int ld_ga, float* dw_t, float* dxe,
                               const int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(w_t && dw_t, "tpconv_bwd: null pointer");
  return tpconv_bwd_impl(tp_id, rowptr, src, perm, geom, w_t, n_edges, nullptr, nullptr, 0, xt, n_nodes, agg, gA, ld_ga, dw_t,
                         nullptr, dxe, argmax, stream);
}

extern "C" int eqnn_tpconv_bwd_shared(int tp_id, const int32_t* rowptr, const int32_t* src, const int32_t* perm,
                                      const float* geom, const int32_t* uidx, const float* w_rows, int ld_rows,
                                      const float* xt, int n_nodes, int agg, const float* gA, int ld_ga, float* dw_rows,
                                      float* dxe, const int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(uidx && w_rows && dw_rows, "tpconv_bwd_shared: null pointer");
  return tpconv_bwd_impl(tp_id, rowptr, src, perm, geom, nullptr, 0, uidx, w_rows, ld_rows, xt, n_nodes, agg, gA, ld_ga, nullptr,
                         dw_rows, dxe, argmax, stream);
}

// 128-bit variants (dxp a multiple of 4, 16-byte aligned buffers): one thread per (sender, component quad); the
// threads of a sender read whole 32-byte sectors of its records
__global__ void sender_reduce4_kernel(const float4* __restrict__ dxe, int n_nodes, int k, int q4, float4* __restrict__ dxt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * q4) return;
  const int64_t s = i / q4;
  const int c = (int)(i - s * q4);
  const float4* p = dxe + (size_t)s * k * q4 + c;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int j = 0;
  for (; j + 3 < k; j += 4) {                     // four independent loads in flight, summed in ascending edge order
    const float4 v0 = __ldg(p + (size_t)j * q4), v1 = __ldg(p + (size_t)(j + 1) * q4), v2 = __ldg(p + (size_t)(j + 2) * q4),
                 v3 = __ldg(p + (size_t)(j + 3) * q4);
    a.x = (((a.x + v0.x) + v1.x) + v2.x) + v3.x;
    a.y = (((a.y + v0.y) + v1.y) + v2.y) + v3.y;
    a.z = (((a.z + v0.z) + v1.z) + v2.z) + v3.z;
    a.w = (((a.w + v0.w) + v1.w) + v2.w) + v3.w;
  }
  for (; j < k; ++j) {
    const float4 v = __ldg(p + (size_t)j * q4);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  dxt[i] = a;
}
__global__ void segment_gather_sum4_kernel(const float4* __restrict__ dxe, const int32_t* __restrict__ rowptr_s,
                                           const int32_t* __restrict__ perm_s, int n_nodes, int q4,
                                           float4* __restrict__ dxt) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * q4) return;
  const int64_t s = i / q4;
  const int c = (int)(i - s * q4);
  const int q1 = rowptr_s[s + 1];
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  int q = rowptr_s[s];
  for (; q + 3 < q1; q += 4) {
    const int e0 = perm_s[q], e1 = perm_s[q + 1], e2 = perm_s[q + 2], e3 = perm_s[q + 3];
    const float4 v0 = __ldg(dxe + (size_t)e0 * q4 + c), v1 = __ldg(dxe + (size_t)e1 * q4 + c),
                 v2 = __ldg(dxe + (size_t)e2 * q4 + c), v3 = __ldg(dxe + (size_t)e3 * q4 + c);
    a.x = (((a.x + v0.x) + v1.x) + v2.x) + v3.x;
    a.y = (((a.y + v0.y) + v1.y) + v2.y) + v3.y;
    a.z = (((a.z + v0.z) + v1.z) + v2.z) + v3.z;
    a.w = (((a.w + v0.w) + v1.w) + v2.w) + v3.w;
  }
  for (; q < q1; ++q) {
    const float4 v = __ldg(dxe + (size_t)perm_s[q] * q4 + c);
    a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
  }
  dxt[i] = a;
}
// 256-bit variants (dxp == 8, 32-byte aligned buffers): one thread per sender, one access per 32-byte record
__global__ void sender_reduce8_kernel(const float* __restrict__ dxe, int n_nodes, int k, float* __restrict__ dxt) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_nodes) return;
  const float* p = dxe + (size_t)s * k * 8;
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int j = 0;
  for (; j + 1 < k; j += 2) {                      // two records in flight, summed in ascending edge order
    float v0[8], v1[8];
    tc::ldg256(p + (size_t)j * 8, v0);
    tc::ldg256(p + (size_t)(j + 1) * 8, v1);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = (a[c] + v0[c]) + v1[c];
  }
  if (j < k) {
    float v0[8];
    tc::ldg256(p + (size_t)j * 8, v0);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] += v0[c];
  }
  tc::stg256(dxt + (size_t)s * 8, a);
}
__global__ void segment_gather_sum8_kernel(const float* __restrict__ dxe, const int32_t* __restrict__ rowptr_s,
                                           const int32_t* __restrict__ perm_s, int n_nodes, float* __restrict__ dxt) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_nodes) return;
  const int q1 = rowptr_s[s + 1];
  float a[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
  int q = rowptr_s[s];
  for (; q + 3 < q1; q += 4) {                     // four records in flight, summed in ascending order
    const int e0 = perm_s[q], e1 = perm_s[q + 1], e2 = perm_s[q + 2], e3 = perm_s[q + 3];
    float v0[8], v1[8], v2[8], v3[8];
    tc::ldg256(dxe + (size_t)e0 * 8, v0);
    tc::ldg256(dxe + (size_t)e1 * 8, v1);
    tc::ldg256(dxe + (size_t)e2 * 8, v2);
    tc::ldg256(dxe + (size_t)e3 * 8, v3);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] = (((a[c] + v0[c]) + v1[c]) + v2[c]) + v3[c];
  }
  for (; q < q1; ++q) {
    float v0[8];
    tc::ldg256(dxe + (size_t)perm_s[q] * 8, v0);
#pragma unroll
    for (int c = 0; c < 8; ++c) a[c] += v0[c];
  }
  tc::stg256(dxt + (size_t)s * 8, a);
}
static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }
static inline bool aligned32(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 31) == 0; }

extern "C" int eqnn_tp_messages(int tp_id, const int32_t* src, const int32_t* perm, const float* geom, const float* w_rows,
                                int ld_w, const int* col, const float* scale, const float* xt, int64_t n_edges, float* out,
                                int ld_out, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(src && perm && geom && w_rows && col && scale && xt && out, "tp_messages: null pointer");
  EQNN_CHECK_ARG(((uintptr_t)geom & 15) == 0 && ((uintptr_t)xt & 31) == 0, "tp_messages: geom must be 16-byte, xt 32-byte aligned");
  if (n_edges <= 0) return EQNN_OK;
  EQNN_TP_DISPATCH(tp_id, {
    EQNN_CHECK_ARG(TP::NLIVE <= 32, "tp_messages: more than 32 message columns");
    EQNN_CHECK_ARG(ld_out >= TP::DLIVE, "tp_messages: ld_out smaller than the message width");
    TpMsgCols c;
    for (int i = 0; i < 32; ++i) {
      c.col[i] = i < TP::NLIVE ? col[i] : 0;
      c.scale[i] = i < TP::NLIVE ? scale[i] : 0.f;
      EQNN_CHECK_ARG(c.col[i] >= 0 && c.col[i] < ld_w, "tp_messages: column index outside w_rows");
    }
    tp_messages_kernel<TP><<<(unsigned)ceil_div64(n_edges, 128), 128, 0, as_stream(stream)>>>(
        src, perm, reinterpret_cast<const float4*>(geom), w_rows, ld_w, c, xt, n_edges, out, ld_out);
  });
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_sender_reduce(const float* dxe, int n_nodes, int k, int dxp, float* dxt,
                                  eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dxe && dxt && k >= 1 && dxp >= 1, "sender_reduce: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  if (dxp == 8 && aligned32(dxe) && aligned32(dxt)) {
    sender_reduce8_kernel<<<(unsigned)ceil_div64(n_nodes, 128), 128, 0, as_stream(stream)>>>(dxe, n_nodes, k, dxt);
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  if ((dxp & 3) == 0 && aligned16(dxe) && aligned16(dxt)) {
    const int64_t n4 = (int64_t)n_nodes * (dxp / 4);
    sender_reduce4_kernel<<<(unsigned)ceil_div64(n4, 128), 128, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(dxe), n_nodes, k, dxp / 4, reinterpret_cast<float4*>(dxt));
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  const int64_t n = (int64_t)n_nodes * dxp;
  sender_reduce_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(dxe, n_nodes, k, dxp, dxt);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segment_gather_sum(const float* dxe, const int32_t* rowptr_s, const int32_t* perm_s,
                                       int n_nodes, int dxp, float* dxt, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dxe && rowptr_s && perm_s && dxt && dxp >= 1, "segment_gather_sum: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  if (dxp == 8 && aligned32(dxe) && aligned32(dxt)) {
    segment_gather_sum8_kernel<<<(unsigned)ceil_div64(n_nodes, 128), 128, 0, as_stream(stream)>>>(dxe, rowptr_s, perm_s, n_nodes,
                                                                                              dxt);
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  if ((dxp & 3) == 0 && aligned16(dxe) && aligned16(dxt)) {
    const int64_t n4 = (int64_t)n_nodes * (dxp / 4);
    segment_gather_sum4_kernel<<<(unsigned)ceil_div64(n4, 128), 128, 0, as_stream(stream)>>>(
        reinterpret_cast<const float4*>(dxe), rowptr_s, perm_s, n_nodes, dxp / 4, reinterpret_cast<float4*>(dxt));
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  const int64_t n = (int64_t)n_nodes * dxp;
  segment_gather_sum_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(dxe, rowptr_s, perm_s,
                                                                                        n_nodes, dxp, dxt);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_sh_scatter(int lmax, const int32_t* rowptr, const float* geom, int n_nodes, float scale,
                               float* attrs, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rowptr && geom && attrs, "sh_scatter: null pointer");
  EQNN_CHECK_ARG(lmax >= 1 && lmax <= 3, "sh_scatter: lmax must be 1..3");
  if (n_nodes <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)ceil_div64((int64_t)n_nodes * 32, 256);
  const float4* g = reinterpret_cast<const float4*>(geom);
  cudaStream_t st = as_stream(stream);
  if (lmax == 1) sh_scatter_kernel<SH_1><<<grid, 256, 0, st>>>(rowptr, g, n_nodes, scale, attrs);
  else if (lmax == 2) sh_scatter_kernel<SH_2><<<grid, 256, 0, st>>>(rowptr, g, n_nodes, scale, attrs);
  else sh_scatter_kernel<SH_3><<<grid, 256, 0, st>>>(rowptr, g, n_nodes, scale, attrs);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_node_tp_fwd(int tp_id, const float* x, const float* y, int64_t n, float* T,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && y && T, "node_tp_fwd: null pointer");
  EQNN_CHECK_ARG(((uintptr_t)x & 31) == 0, "node_tp_fwd: x must be 32-byte aligned");
  if (n <= 0) return EQNN_OK;
  EQNN_TP_DISPATCH(tp_id, {
    node_tp_fwd_kernel<TP><<<(unsigned)ceil_div64(n, 128), 128, 0, as_stream(stream)>>>(x, y, n, T);
  });
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_node_tp_bwd(int tp_id, const float* x, const float* y, const float* gT, int64_t n, float* dx,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && y && gT && dx, "node_tp_bwd: null pointer");
  EQNN_CHECK_ARG(((uintptr_t)x & 31) == 0, "node_tp_bwd: x must be 32-byte aligned");
  if (n <= 0) return EQNN_OK;
  EQNN_TP_DISPATCH(tp_id, {
    node_tp_bwd_kernel<TP><<<(unsigned)ceil_div64(n, 128), 128, 0, as_stream(stream)>>>(x, y, gT, n, dx);
  });
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}
