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
//   pair_follow    uidx of the non-representatives = uidx of their partner (zero class: of its first member)
//   pair_combine   dL/dw of a shared evaluation = sum over its (one or two) member edges, row-major records in,
//                  column-major [n_out][ld] out (what eqnn_radial_mlp_bwd reads)
//   bessel_rows    e3nn.bessel of the unique lengths (first radial layer's weight-gradient operand)
// The number of unique evaluations stays ON THE DEVICE (n_unique): every consumer takes the capacity as its launch
// bound and reads the live count from that word, so the step needs no host synchronisation and replays from a CUDA graph.
#include "common.cuh"
#include "rmlp_common.cuh"

namespace {

// first slot of [p0, p1) whose sender is `who`; four loads in flight per step (the rows are ~k long and the loop is a
// chain of L2 round trips otherwise)
__device__ __forceinline__ int first_slot_with_sender(const int32_t* __restrict__ src, int p0, int p1, int who) {
  for (int q = p0; q < p1; q += 4) {
    const int a0 = __ldg(src + q);
    const int a1 = q + 1 < p1 ? __ldg(src + q + 1) : -1;
    const int a2 = q + 2 < p1 ? __ldg(src + q + 2) : -1;
    const int a3 = q + 3 < p1 ? __ldg(src + q + 3) : -1;
    if (a0 == who) return q;
    if (a1 == who) return q + 1;
    if (a2 == who) return q + 2;
    if (a3 == who) return q + 3;
  }
  return -1;
}

// partner[p]: >= 0 the reverse edge's slot; -1 unpaired; -2 member of the ZERO CLASS -- a self-loop (s == r) of length
// +0.0: every node of a kNN graph has one (graph_utils.py:67, self is neighbour 0), all with the same length, so all of
// them share ONE evaluation (5 % of the edges at k = 20).  Its representative is the smallest such slot (atomicMin: the
// result does not depend on the order of the updates).
__global__ void __launch_bounds__(256) pair_find_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ src,
                                                        const float4* __restrict__ geom, int n_nodes,
                                                        int32_t* __restrict__ partner, int32_t* __restrict__ zero_slot,
                                                        int32_t* __restrict__ zero_first) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int r = (int)row;
  const int p0 = rowptr[r], p1 = rowptr[r + 1];
  int zs = -1;
  for (int base = p0; base < p1; base += 32) {
    const int p = base + lane;
    bool zero = false;
    if (p < p1) {
      const int s = __ldg(src + p);
      int q = -1;
      if (s == r) {
        if (__float_as_uint(geom[p].w) == 0u && first_slot_with_sender(src, p0, p1, r) == p) {
          q = -2;
          zero = true;
        }
      } else if (s >= 0 && s < n_nodes) {
        q = first_slot_with_sender(src, rowptr[s], rowptr[s + 1], r);
        if (q >= 0) {
          const bool mutual = first_slot_with_sender(src, p0, p1, s) == p;
          const bool same_d = __float_as_uint(geom[p].w) == __float_as_uint(geom[q].w);
          if (!(mutual && same_d)) q = -1;
        }
      }
      partner[p] = q;
    }
    const unsigned m = __ballot_sync(0xffffffffu, zero);
    if (m && zs < 0) zs = base + __ffs(m) - 1;
  }
  if (lane == 0) {
    zero_slot[r] = zs;
    if (zs >= 0) atomicMin(zero_first, zs);
  }
}

__device__ __forceinline__ bool pair_is_rep(int p, int q, int zero_first) {
  return q == -1 || (q >= 0 && p < q) || (q == -2 && p == zero_first);
}

__global__ void __launch_bounds__(256) pair_count_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ partner,
                                                         const int32_t* __restrict__ zero_first, int n_nodes,
                                                         int32_t* __restrict__ rowcount) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1], zf = *zero_first;
  int reps = 0;
  for (int base = p0; base < p1; base += 32) {
    const int p = base + lane;
    reps += __popc(__ballot_sync(0xffffffffu, p < p1 && pair_is_rep(p, partner[p < p1 ? p : p0], zf)));
  }
  if (lane == 0) rowcount[row] = reps;
}

__global__ void __launch_bounds__(256) pair_number_kernel(const int32_t* __restrict__ rowptr, const int32_t* __restrict__ urow,
                                                          const int32_t* __restrict__ partner, const float4* __restrict__ geom,
                                                          const int32_t* __restrict__ zero_first, int n_nodes,
                                                          int32_t* __restrict__ uidx, int32_t* __restrict__ rep_slot,
                                                          int32_t* __restrict__ rep_partner, float* __restrict__ d_u,
                                                          int32_t* __restrict__ n_unique, int32_t* __restrict__ zero_row) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  if (row == 0 && lane == 0) *n_unique = urow[n_nodes];
  const int p0 = rowptr[row], p1 = rowptr[row + 1], zf = *zero_first;
  int u0 = urow[row];
  for (int base = p0; base < p1; base += 32) {
    const int p = base + lane;
    const int q = p < p1 ? partner[p] : 0;
    const bool is_rep = p < p1 && pair_is_rep(p, q, zf);
    const unsigned m = __ballot_sync(0xffffffffu, is_rep);
    if (is_rep) {
      const int u = u0 + __popc(m & ((1u << lane) - 1u));
      uidx[p] = u;
      rep_slot[u] = p;
      rep_partner[u] = q;
      d_u[u] = geom[p].w;
      if (q == -2) *zero_row = u;
    }
    u0 += __popc(m);
  }
}

__global__ void pair_follow_kernel(const int32_t* __restrict__ partner, const int32_t* __restrict__ zero_first, int64_t n_edges,
                                   int32_t* __restrict__ uidx) {
  const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n_edges) return;
  const int q = partner[p];
  // the slots read here are representatives: written by pair_number_kernel, never by this kernel
  if (q >= 0 && q < p) uidx[p] = uidx[q];
  else if (q == -2 && p != *zero_first) uidx[p] = uidx[*zero_first];
}

// dL/dw of the zero class = sum over its member slots (one per node): per-CTA sums over a contiguous range of nodes in a
// fixed order, then the CTA sums in CTA order -- deterministic
constexpr int ZS_BLOCKS = 128;
template <int NQ>
__global__ void __launch_bounds__(256) zero_sum_kernel(const float4* __restrict__ dw_rows, const int32_t* __restrict__ zero_slot,
                                                       int n_nodes, float* __restrict__ partial) {
  __shared__ float red[256][4 * NQ + 1];
  const int tid = threadIdx.x;
  const int per = (n_nodes + gridDim.x - 1) / gridDim.x;
  const int i0 = blockIdx.x * per, i1 = min(i0 + per, n_nodes);
  float v[4 * NQ];
#pragma unroll
  for (int c = 0; c < 4 * NQ; ++c) v[c] = 0.f;
  for (int i = i0 + tid; i < i1; i += 256) {
    const int p = zero_slot[i];
    if (p >= 0) {
#pragma unroll
      for (int j = 0; j < NQ; ++j) {
        const float4 a = __ldg(dw_rows + (size_t)p * NQ + j);
        v[4 * j] += a.x; v[4 * j + 1] += a.y; v[4 * j + 2] += a.z; v[4 * j + 3] += a.w;
      }
    }
  }
#pragma unroll
  for (int c = 0; c < 4 * NQ; ++c) red[tid][c] = v[c];
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) {
#pragma unroll
      for (int c = 0; c < 4 * NQ; ++c) red[tid][c] += red[tid + o][c];
    }
    __syncthreads();
  }
  if (tid < 4 * NQ) partial[blockIdx.x * 16 + tid] = red[0][tid];
}
__global__ void zero_sum_final_kernel(const float* __restrict__ partial, int n_parts, int n_out, const int32_t* __restrict__ zero_row,
                                      float* __restrict__ dw_t, int64_t ld) {
  const int c = threadIdx.x;
  const int u = *zero_row;
  if (c >= n_out || u < 0) return;
  float t = 0.f;
  for (int b = 0; b < n_parts; ++b) t += partial[b * 16 + c];
  dw_t[(int64_t)c * ld + u] = t;
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

// the same basis with the FORWARD kernel's arithmetic (rmlp::bessel16: packed Cody-Waite sine on the fp32-rounded phase,
// 1e-7 relative to the libm form above): the first radial layer's weight gradient then differentiates exactly what the
// forward kernel fed to that layer, at a third of the instructions.  Thread = (row, group of 16 basis functions).
__global__ void __launch_bounds__(256) bessel_rows16_kernel(const float* __restrict__ d_rows, int64_t cap,
                                                            const int32_t* __restrict__ n_rows, int n_rb, float r_cut,
                                                            float pref, float* __restrict__ out) {
  __shared__ float wfreq[64];
  if (threadIdx.x < n_rb) wfreq[threadIdx.x] = __fdiv_rn(__fmul_rn((float)(threadIdx.x + 1), 3.14159265358979323846f), r_cut);
  __syncthreads();
  const int gpr = n_rb / 16;                                   // groups per row
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n = n_rows ? min((int64_t)*n_rows, cap) : cap;
  const int64_t e = i / gpr;
  const int g = (int)(i - e * gpr);
  if (e >= n) return;
  float v[16];
  rmlp::bessel16(wfreq, 16 * g, __ldg(d_rows + e), pref, v);
  float4* o = reinterpret_cast<float4*>(out + e * n_rb + 16 * g);
#pragma unroll
  for (int q = 0; q < 4; ++q) o[q] = make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]);
}

}  // namespace

size_t eqnn_scan_scratch_bytes(int64_t n);                                                             // csr.cu
int eqnn_scan_rows_i32(const int32_t* count, int64_t n, int32_t* out, int32_t* cursor, void* scratch, cudaStream_t st);

extern "C" size_t eqnn_edge_pairs_workspace_bytes(int n_nodes, int64_t n_edges) {
  return align_up((size_t)n_edges * 4, 256) + 2 * align_up(((size_t)n_nodes + 1) * 4, 256) + 256 + eqnn_scan_scratch_bytes(n_nodes);
}

extern "C" int eqnn_edge_pairs(const int32_t* rowptr, const int32_t* src, const float* geom, int n_nodes, int64_t n_edges,
                               int32_t* uidx, int32_t* rep_slot, int32_t* rep_partner, float* d_u, int32_t* zero_slot,
                               int32_t* n_unique, int32_t* zero_row, void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rowptr && src && geom && uidx && rep_slot && rep_partner && d_u && zero_slot && n_unique && zero_row,
                 "edge_pairs: null pointer");
  EQNN_CHECK_ARG(n_nodes >= 0 && n_edges >= 0 && n_edges < ((int64_t)1 << 31), "edge_pairs: bad shape");
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(geom) & 15) == 0, "edge_pairs: geom must be 16-byte aligned");
  EQNN_CHECK_ARG(ws && ws_bytes >= eqnn_edge_pairs_workspace_bytes(n_nodes, n_edges), "edge_pairs: workspace too small");
  cudaStream_t st = as_stream(stream);
  EQNN_CUDA(cudaMemsetAsync(zero_row, 0xFF, 4, st));          // -1: no zero class
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
  w += align_up(((size_t)n_nodes + 1) * 4, 256);
  int32_t* zero_first = reinterpret_cast<int32_t*>(w);
  w += 256;
  void* scan_scratch = w;
  const float4* g4 = reinterpret_cast<const float4*>(geom);
  const unsigned wgrid = (unsigned)ceil_div64((int64_t)n_nodes * 32, 256);
  EQNN_CUDA(cudaMemsetAsync(zero_first, 0x7F, 4, st));        // 0x7f7f7f7f: above every slot index
  pair_find_kernel<<<wgrid, 256, 0, st>>>(rowptr, src, g4, n_nodes, partner, zero_slot, zero_first);
  pair_count_kernel<<<wgrid, 256, 0, st>>>(rowptr, partner, zero_first, n_nodes, rowcount);
  EQNN_CHECK_LAUNCH_N(2);
  const int rc = eqnn_scan_rows_i32(rowcount, n_nodes, urow, nullptr, scan_scratch, st);
  if (rc) return rc;
  pair_number_kernel<<<wgrid, 256, 0, st>>>(rowptr, urow, partner, g4, zero_first, n_nodes, uidx, rep_slot, rep_partner, d_u,
                                            n_unique, zero_row);
  pair_follow_kernel<<<(unsigned)ceil_div64(n_edges, 256), 256, 0, st>>>(partner, zero_first, n_edges, uidx);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" size_t eqnn_pair_combine_workspace_bytes(void) { return (size_t)ZS_BLOCKS * 16 * sizeof(float); }

extern "C" int eqnn_pair_combine(const float* dw_rows, int ld_rows, int n_out, const int32_t* rep_slot,
                                 const int32_t* rep_partner, const int32_t* zero_slot, const int32_t* zero_row, int n_nodes,
                                 int64_t cap, const int32_t* n_unique, float* dw_t, int64_t ld_dwt, void* ws, size_t ws_bytes,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dw_rows && rep_slot && rep_partner && dw_t, "pair_combine: null pointer");
  EQNN_CHECK_ARG(n_out >= 1 && n_out <= 16 && ld_rows % 4 == 0 && ld_rows >= n_out && ld_rows <= 16, "pair_combine: bad row width");
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(dw_rows) & 15) == 0 && ld_dwt >= cap, "pair_combine: alignment / leading dimension");
  EQNN_CHECK_ARG(!zero_slot || (zero_row && ws && ws_bytes >= eqnn_pair_combine_workspace_bytes()),
                 "pair_combine: the zero class needs zero_row and the workspace");
  if (cap <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)ceil_div64(cap, 256);
  const float4* r4 = reinterpret_cast<const float4*>(dw_rows);
  cudaStream_t st = as_stream(stream);
  float* partial = reinterpret_cast<float*>(ws);
  const bool zc = zero_slot != nullptr && n_nodes > 0;
  const int zb = n_nodes < ZS_BLOCKS * 256 ? (n_nodes + 255) / 256 : ZS_BLOCKS;
#define EQNN_PAIR_COMBINE(NQ)                                                                                          \
  pair_combine_kernel<NQ><<<grid, 256, 0, st>>>(r4, rep_slot, rep_partner, cap, n_unique, n_out, dw_t, ld_dwt);        \
  if (zc) zero_sum_kernel<NQ><<<zb, 256, 0, st>>>(r4, zero_slot, n_nodes, partial)
  switch (ld_rows / 4) {
    case 1: EQNN_PAIR_COMBINE(1); break;
    case 2: EQNN_PAIR_COMBINE(2); break;
    case 3: EQNN_PAIR_COMBINE(3); break;
    default: EQNN_PAIR_COMBINE(4); break;
  }
#undef EQNN_PAIR_COMBINE
  if (zc) zero_sum_final_kernel<<<1, 32, 0, st>>>(partial, zb, n_out, zero_row, dw_t, ld_dwt);
  EQNN_CHECK_LAUNCH_N(zc ? 3 : 1);
  return EQNN_OK;
}

extern "C" int eqnn_bessel_rows(const float* d_rows, int64_t cap, const int32_t* n_rows, int n_rb, double r_cut, float* out,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(d_rows && out && n_rb >= 1 && r_cut > 0.0 && cap >= 0, "bessel_rows: bad argument");
  if (cap == 0) return EQNN_OK;
  const float pref = (float)sqrt(2.0 / r_cut);
  if ((n_rb == 16 || n_rb == 32 || n_rb == 64) && (reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const int64_t threads = cap * (n_rb / 16);
    bessel_rows16_kernel<<<(unsigned)ceil_div64(threads, 256), 256, 0, as_stream(stream)>>>(d_rows, cap, n_rows, n_rb,
                                                                                         (float)r_cut, pref, out);
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  const int64_t warps = ceil_div64(cap, 8);
  bessel_rows_kernel<<<(unsigned)ceil_div64(warps * 32, 256), 256, 0, as_stream(stream)>>>(d_rows, cap, n_rows, n_rb,
                                                                                         (float)r_cut, pref, out);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}
