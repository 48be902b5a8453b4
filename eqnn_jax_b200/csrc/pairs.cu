// Shared radial evaluations: the radial MLP of the interaction block (models/nequip.py:55-68) sees an edge only through
// its length d = |x_s - x_r|, and the two directions of a reciprocated edge (i -> j and j -> i; 85 % of the edges of a
// kNN-20 graph on a uniform cloud) have BIT-IDENTICAL lengths in fp32 (the subtraction is exactly antisymmetric, the
// squares and their sum are even).  Evaluating the MLP once per unordered pair is therefore the same arithmetic on the
// same number -- the forward result is identical per edge, and because the backward pass is linear in dL/dw for a fixed
// d, the weight gradient of the pair is the gradient of the summed cotangent.  Exact work elimination, like the live
// message columns: every edge still receives its radial weights and contributes its gradient.
//
// This file finds the pairs on the receiver-sorted CSR (no sort, no atomics, deterministic numbering):
//   pair_find      warp per receiver row r, lane per incoming slot p (sender s): q = first slot of row s whose sender is
//                  r; p and q are paired iff q != p, the first slot of row r whose sender is s is p itself (duplicate
//                  edges pair at most once) and the two lengths compare equal bit for bit.  The smaller slot is the
//                  pair's representative; every unpaired slot (self-loops, one-directional edges) represents itself.
//   scan           exclusive scan of the representatives per row (csr.cu's scan) -> numbering in CSR order
//   pair_number    u = position of the representative: uidx[p], rep_slot[u], rep_partner[u], d_u[u]; *n_unique
//   pair_follow    uidx of the non-representatives = uidx of their partner
//   pair_combine   dL/dw of a shared evaluation = sum over its (one or two) member edges, row-major records in,
//                  column-major [n_out][ld] out (what eqnn_radial_mlp_bwd reads)
//   bessel_rows    e3nn.bessel of the unique lengths (first radial layer's weight-gradient operand)
// The number of unique evaluations stays ON THE DEVICE (n_unique): every consumer takes the capacity as its launch
// bound and reads the live count from that word, so the step needs no host synchronisation and replays from a CUDA graph.
#include "common.cuh"

namespace {

__device__ __forceinline__ int first_slot_with_sender(const int32_t* __restrict__ src, int p0, int p1, int who) {
  for (int q = p0; q < p1; ++q)
    if (__ldg(src + q) == who) return q;
  return -1;
}

__global__ void __launch_bounds__(256) pair_find_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                                                        const float4* __restrict__ geom, int n_nodes,
                                                        int32_t* __restrict__ partner, int32_t* __restrict__ rowcount) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int r = (int)row;
  const int p0 = rowptr[r], p1 = rowptr[r + 1];
  int reps = 0;
  for (int base = p0; base < p1; base += 32) {
    const int p = base + lane;
    bool is_rep = false;
    if (p < p1) {
      const int s = __ldg(src + p);
      int q = -1;
      if (s != r && s >= 0 && s < n_nodes) {
        q = first_slot_with_sender(src, rowptr[s], rowptr[s + 1], r);
        if (q >= 0) {
          const bool mutual = first_slot_with_sender(src, p0, p1, s) == p;
          const bool same_d = __float_as_uint(geom[p].w) == __float_as_uint(geom[q].w);
          if (!(mutual && same_d)) q = -1;
        }
      }
      partner[p] = q;
      is_rep = q < 0 || p < q;
    }
    reps += __popc(__ballot_sync(0xffffffffu, is_rep));
  }
  if (lane == 0) rowcount[r] = reps;
}

__global__ void __launch_bounds__(256) pair_number_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ urow,
                                                          const int32_t* __restrict__ partner, const float4* __restrict__ geom,
                                                          int n_nodes, int32_t* __restrict__ uidx, int32_t* __restrict__ rep_slot,
                                                          int32_t* __restrict__ rep_partner, float* __restrict__ d_u,
                                                          int32_t* __restrict__ n_unique) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  if (row == 0 && lane == 0) *n_unique = urow[n_nodes];
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  int u0 = urow[row];
  for (int base = p0; base < p1; base += 32) {
    const int p = base + lane;
    const int q = p < p1 ? partner[p] : 0;
    const bool is_rep = p < p1 && (q < 0 || p < q);
    const unsigned m = __ballot_sync(0xffffffffu, is_rep);
    if (is_rep) {
      const int u = u0 + __popc(m & ((1u << lane) - 1u));
      uidx[p] = u;
      rep_slot[u] = p;
      rep_partner[u] = q;
      d_u[u] = geom[p].w;
    }
    u0 += __popc(m);
  }
}

__global__ void pair_follow_kernel(const int32_t* __restrict__ partner, int64_t n_edges, int32_t* __restrict__ uidx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_edges) return;
  const int q = partner[p];
  if (q >= 0 && q < p) uidx[p] = uidx[q];     // q is a representative: written by pair_number_kernel, never here
}

template <int NQ>
__global__ void __launch_bounds__(256) pair_combine_kernel(const float4* __restrict__ dw_rows, const int32_t* __restrict__ rep_slot,
                                                           const int32_t* __restrict__ rep_partner, int64_t cap,
                                                           const int32_t* __restrict__ n_unique, int n_out,
                                                           float* __restrict__ dw_t, int64_t ld) {
  const int64_t u = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = n_unique ? min((int64_t)*n_unique, cap) : cap;
  if (u >= n) return;
  const int p = rep_slot[u], q = rep_partner[u];
  float v[4 * NQ];
#pragma unroll
  for (int i = 0; i < NQ; ++i) {
    const float4 a = __ldg(dw_rows + (size_t)p * NQ + i);
    v[4 * i] = a.x; v[4 * i + 1] = a.y; v[4 * i + 2] = a.z; v[4 * i + 3] = a.w;
  }
  if (q >= 0) {
#pragma unroll
    for (int i = 0; i < NQ; ++i) {
      const float4 b = __ldg(dw_rows + (size_t)q * NQ + i);
      v[4 * i] += b.x; v[4 * i + 1] += b.y; v[4 * i + 2] += b.z; v[4 * i + 3] += b.w;
    }
  }
#pragma unroll
  for (int c = 0; c < 4 * NQ; ++c)
    if (c < n_out) dw_t[(int64_t)c * ld + u] = v[c];
}

// e3nn.bessel with the reference's fp32 rounding points -- the arithmetic of csr.cu's geom_kernel, one warp per row of 32
// lanes over the basis index
__global__ void __launch_bounds__(256) bessel_rows_kernel(const float* __restrict__ d_rows, int64_t cap,
                                                          const int32_t* __restrict__ n_rows, int n_rb, float r_cut,
                                                          float pref, float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t warp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n = n_rows ? min((int64_t)*n_rows, cap) : cap;
  const float PI_F = 3.14159265358979323846f;
  const int64_t e0 = warp * 8;
  if (e0 >= n) return;
  for (int j = 1 + lane; j <= n_rb; j += 32) {
    const float w = __fdiv_rn(__fmul_rn((float)j, PI_F), r_cut);
    for (int64_t e = e0; e < min(e0 + 8, n); ++e) {
      const float d = __ldg(d_rows + e);
      const float v = (d == 0.f) ? w : __fdiv_rn(sinf(__fmul_rn(w, d)), d);
      out[e * n_rb + j - 1] = __fmul_rn(pref, v);
    }
  }
}

}  // namespace

int eqnn_scan_rows_i32(const int32_t* count, int64_t n, int32_t* out, cudaStream_t st);   // csr.cu

extern "C" size_t eqnn_edge_pairs_workspace_bytes(int n_nodes, int64_t n_edges) {
  return align_up((size_t)n_edges * 4, 256) + 2 * align_up(((size_t)n_nodes + 1) * 4, 256);
}

extern "C" int eqnn_edge_pairs(const int32_t* rowptr, const int32_t* src, const float* geom, int n_nodes, int64_t n_edges,
                               int32_t* uidx, int32_t* rep_slot, int32_t* rep_partner, float* d_u, int32_t* n_unique,
                               void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rowptr && src && geom && uidx && rep_slot && rep_partner && d_u && n_unique, "edge_pairs: null pointer");
  EQNN_CHECK_ARG(n_nodes >= 0 && n_edges >= 0 && n_edges < ((int64_t)1 << 31), "edge_pairs: bad shape");
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(geom) & 15) == 0, "edge_pairs: geom must be 16-byte aligned");
  EQNN_CHECK_ARG(ws && ws_bytes >= eqnn_edge_pairs_workspace_bytes(n_nodes, n_edges), "edge_pairs: workspace too small");
  cudaStream_t st = as_stream(stream);
  if (n_nodes == 0 || n_edges == 0) {
    EQNN_CUDA(cudaMemsetAsync(n_unique, 0, 4, st));
    return EQNN_OK;
  }
  char* w = reinterpret_cast<char*>(ws);
  int32_t* partner = reinterpret_cast<int32_t*>(w);
  w += align_up((size_t)n_edges * 4, 256);
  int32_t* rowcount = reinterpret_cast<int32_t*>(w);
  w += align_up(((size_t)n_nodes + 1) * 4, 256);
  int32_t* urow = reinterpret_cast<int32_t*>(w);
  const float4* g4 = reinterpret_cast<const float4*>(geom);
  const unsigned wgrid = (unsigned)ceil_div64((int64_t)n_nodes * 32, 256);
  pair_find_kernel<<<wgrid, 256, 0, st>>>(rowptr, src, g4, n_nodes, partner, rowcount);
  EQNN_CHECK_LAUNCH();
  const int rc = eqnn_scan_rows_i32(rowcount, n_nodes, urow, st);
  if (rc) return rc;
  pair_number_kernel<<<wgrid, 256, 0, st>>>(rowptr, urow, partner, g4, n_nodes, uidx, rep_slot, rep_partner, d_u, n_unique);
  pair_follow_kernel<<<(unsigned)ceil_div64(n_edges, 256), 256, 0, st>>>(partner, n_edges, uidx);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" int eqnn_pair_combine(const float* dw_rows, int ld_rows, int n_out, const int32_t* rep_slot,
                                 const int32_t* rep_partner, int64_t cap, const int32_t* n_unique, float* dw_t, int64_t ld_dwt,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dw_rows && rep_slot && rep_partner && dw_t, "pair_combine: null pointer");
  EQNN_CHECK_ARG(n_out >= 1 && n_out <= 16 && ld_rows % 4 == 0 && ld_rows >= n_out && ld_rows <= 16, "pair_combine: bad row width");
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(dw_rows) & 15) == 0 && ld_dwt >= cap, "pair_combine: alignment / leading dimension");
  if (cap <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)ceil_div64(cap, 256);
  const float4* r4 = reinterpret_cast<const float4*>(dw_rows);
  cudaStream_t st = as_stream(stream);
  switch (ld_rows / 4) {
    case 1: pair_combine_kernel<1><<<grid, 256, 0, st>>>(r4, rep_slot, rep_partner, cap, n_unique, n_out, dw_t, ld_dwt); break;
    case 2: pair_combine_kernel<2><<<grid, 256, 0, st>>>(r4, rep_slot, rep_partner, cap, n_unique, n_out, dw_t, ld_dwt); break;
    case 3: pair_combine_kernel<3><<<grid, 256, 0, st>>>(r4, rep_slot, rep_partner, cap, n_unique, n_out, dw_t, ld_dwt); break;
    default: pair_combine_kernel<4><<<grid, 256, 0, st>>>(r4, rep_slot, rep_partner, cap, n_unique, n_out, dw_t, ld_dwt); break;
  }
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_bessel_rows(const float* d_rows, int64_t cap, const int32_t* n_rows, int n_rb, double r_cut, float* out,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(d_rows && out && n_rb >= 1 && r_cut > 0.0 && cap >= 0, "bessel_rows: bad argument");
  if (cap == 0) return EQNN_OK;
  const float pref = (float)sqrt(2.0 / r_cut);
  const int64_t warps = ceil_div64(cap, 8);
  bessel_rows_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, as_stream(stream)>>>(d_rows, cap, n_rows, n_rb,
                                                                                         (float)r_cut, pref, out);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}
