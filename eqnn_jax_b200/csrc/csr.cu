// Receiver-sorted CSR of a batch of graphs + per-edge geometry (sm_100a).
//
// jraph aggregates messages with an unsorted scatter-add over `receivers`
// (models/nequip.py:118-120 via jraph.GraphNetwork).  Sorting the edge list by
// receiver once per batch lets the interaction kernel reduce each receiver's
// messages without atomics, and -- because edges inside a row are kept in
// ascending edge id -- in the same order a sequential CPU scatter-add would.
#include "common.cuh"

namespace {

__global__ void hist_kernel(const int32_t* __restrict__ recv, int64_t n_edges, int E, int N,
                            int32_t* __restrict__ count) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int b = (int)(e / E);
  int r = recv[e];
  if (r >= 0 && r < N) atomicAdd(&count[(int64_t)b * N + r], 1);
}

// single-CTA exclusive scan of n ints: 16 consecutive elements per thread and round (16 384 per round, four barriers
// each), so a batch of 160 000 rows takes ten rounds instead of 157
constexpr int SCAN_PER_THREAD = 16;
__global__ void __launch_bounds__(1024) scan_kernel(const int32_t* __restrict__ count, int64_t n,
                                                    int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor) {
  __shared__ int32_t warp_sum[32];
  __shared__ int32_t warp_excl[32];
  __shared__ int32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int64_t base = 0; base < n; base += 1024 * SCAN_PER_THREAD) {
    const int64_t i0 = base + (int64_t)tid * SCAN_PER_THREAD;
    int v[SCAN_PER_THREAD];
    int tot = 0;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
      v[j] = (i0 + j < n) ? count[i0 + j] : 0;
      tot += v[j];
    }
    int incl = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) warp_sum[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      const int w = warp_sum[lane];
      int winc = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        int t = __shfl_up_sync(0xffffffffu, winc, o);
        if (lane >= o) winc += t;
      }
      warp_excl[lane] = winc - w;
    }
    __syncthreads();
    int excl = carry_s + warp_excl[warp] + incl - tot;
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) {
      if (i0 + j < n) {
        rowptr[i0 + j] = excl;
        if (cursor) cursor[i0 + j] = excl;
      }
      excl += v[j];
    }
    __syncthreads();
    if (tid == 1023) carry_s = excl;
    __syncthreads();
  }
  if (tid == 0) rowptr[n] = carry_s;
}

// Multi-CTA scan for long arrays (the single-CTA kernel above takes 190 us for a batch of 160 000 rows: one SM walks the
// array in ten latency-bound rounds).  Three launches: per-chunk totals (4096 elements per CTA) -> scan of the totals (the
// kernel above, one round) -> per-chunk exclusive scan + offset.  Same results (integer sums).
constexpr int SCAN_CHUNK = 256 * SCAN_PER_THREAD;
__device__ __forceinline__ void scan_load16(const int32_t* __restrict__ count, int64_t n, int64_t i0, int* v) {
  if (i0 + SCAN_PER_THREAD <= n && (reinterpret_cast<uintptr_t>(count + i0) & 15) == 0) {
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD / 4; ++j) {
      const int4 t = __ldg(reinterpret_cast<const int4*>(count + i0) + j);
      v[4 * j] = t.x; v[4 * j + 1] = t.y; v[4 * j + 2] = t.z; v[4 * j + 3] = t.w;
    }
  } else {
#pragma unroll
    for (int j = 0; j < SCAN_PER_THREAD; ++j) v[j] = (i0 + j < n) ? count[i0 + j] : 0;
  }
}
__global__ void __launch_bounds__(256) scan_totals_kernel(const int32_t* __restrict__ count, int64_t n,
                                                          int32_t* __restrict__ totals) {
  __shared__ int32_t ws[8];
  const int tid = threadIdx.x;
  int v[SCAN_PER_THREAD];
  scan_load16(count, n, (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)tid * SCAN_PER_THREAD, v);
  int tot = 0;
#pragma unroll
  for (int j = 0; j < SCAN_PER_THREAD; ++j) tot += v[j];
  tot = __reduce_add_sync(0xffffffffu, tot);
  if ((tid & 31) == 0) ws[tid >> 5] = tot;
  __syncthreads();
  if (tid == 0) {
    int t = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += ws[w];
    totals[blockIdx.x] = t;
  }
}
__global__ void __launch_bounds__(256) scan_apply_kernel(const int32_t* __restrict__ count, int64_t n,
                                                         const int32_t* __restrict__ offs, int n_chunks,
                                                         int32_t* __restrict__ rowptr, int32_t* __restrict__ cursor) {
  __shared__ int32_t ws[8];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int64_t i0 = (int64_t)blockIdx.x * SCAN_CHUNK + (int64_t)tid * SCAN_PER_THREAD;
  int v[SCAN_PER_THREAD];
  scan_load16(count, n, i0, v);
  int tot = 0;
#pragma unroll
  for (int j = 0; j < SCAN_PER_THREAD; ++j) tot += v[j];
  int incl = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) ws[warp] = incl;
  __syncthreads();
  int excl = offs[blockIdx.x] + incl - tot;
  for (int w = 0; w < warp; ++w) excl += ws[w];
#pragma unroll
  for (int j = 0; j < SCAN_PER_THREAD; ++j) {
    if (i0 + j < n) {
      rowptr[i0 + j] = excl;
      if (cursor) cursor[i0 + j] = excl;
    }
    excl += v[j];
  }
  if (blockIdx.x == 0 && tid == 0) rowptr[n] = offs[n_chunks];
}

__global__ void fill_kernel(const int32_t* __restrict__ recv, int64_t n_edges, int E, int N,
                            int32_t* __restrict__ cursor, int32_t* __restrict__ perm) {
  int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  int b = (int)(e / E);
  int r = recv[e];
  if (r < 0 || r >= N) return;
  int p = atomicAdd(&cursor[(int64_t)b * N + r], 1);
  perm[p] = (int32_t)e;
}

// one warp per row: order the row's edge ids ascending (rank by counting), emit src
__global__ void __launch_bounds__(256) rowsort_kernel(const int32_t* __restrict__ rowptr, int64_t n_rows,
                                                      const int32_t* __restrict__ senders, int E, int N,
                                                      int32_t* __restrict__ perm, int32_t* __restrict__ perm_tmp,
                                                      int32_t* __restrict__ src) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_rows) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const int deg = p1 - p0;
  // ranks: every lane takes elements lane, lane+32, ...
  for (int a = lane; a < deg; a += 32) {
    const int ea = perm[p0 + a];
    int rank = 0;
    for (int c = 0; c < deg; ++c) rank += (perm[p0 + c] < ea) ? 1 : 0;
    perm_tmp[p0 + rank] = ea;
  }
  __syncwarp();
  for (int a = lane; a < deg; a += 32) {
    const int e = perm_tmp[p0 + a];
    perm[p0 + a] = e;
    const int b = e / E;
    src[p0 + a] = b * N + senders[e];
  }
}

__global__ void geom_kernel(const float* __restrict__ pos, int ld, int n_nodes, const int32_t* __restrict__ rowptr,
                            const int32_t* __restrict__ src, int n_rb, float r_cut, float pref,
                            float4* __restrict__ geom, float* __restrict__ radial) {
  // one warp per receiver row; lanes stride over the row's edges
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float xr = pos[(size_t)row * ld + 0], yr = pos[(size_t)row * ld + 1], zr = pos[(size_t)row * ld + 2];
  const float PI_F = 3.14159265358979323846f;
  for (int p = p0 + lane; p < p1; p += 32) {
    const int s = src[p];
    // r_ij = x_s - x_r  (models/nequip.py:134), every op rounded to fp32 like the reference
    const float rx = __fsub_rn(pos[(size_t)s * ld + 0], xr);
    const float ry = __fsub_rn(pos[(size_t)s * ld + 1], yr);
    const float rz = __fsub_rn(pos[(size_t)s * ld + 2], zr);
    const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(rx, rx), __fmul_rn(ry, ry)), __fmul_rn(rz, rz));
    const float d = __fsqrt_rn(d2);
    const float dn = (d2 == 0.f) ? 1.f : d;  // e3nn guard: divide by 1 at the origin
    geom[p] = make_float4(__fdiv_rn(rx, dn), __fdiv_rn(ry, dn), __fdiv_rn(rz, dn), d);
    if (radial && n_rb == 0) radial[p] = d;
  }
  __syncwarp();   // the Bessel loop reads geom[p].w written by other lanes of this warp
  if (radial && n_rb > 0) {
    // e3nn.bessel: sqrt(2/c) * sin(j*pi/c * x)/x with the fp32 rounding points of the reference.  Lanes own
    // basis indices j (coalesced stores of each edge's n_rb values); edge lengths are re-read (L1 hits).
    // a lane's frequencies fl(fl(j pi) / c) do not depend on the edge: formed once (the correctly rounded division is
    // a dozen instructions), for up to 128 basis functions; wider bases recompute them
    float wj[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) wj[t] = __fdiv_rn(__fmul_rn((float)(1 + lane + 32 * t), PI_F), r_cut);
    for (int p = p0; p < p1; ++p) {
      const float d = geom[p].w;
      float* out = radial + (size_t)p * n_rb;
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const int j = 1 + lane + 32 * t;
        if (j <= n_rb) {
          const float v = (d == 0.f) ? wj[t] : __fdiv_rn(sinf(__fmul_rn(wj[t], d)), d);
          out[j - 1] = __fmul_rn(pref, v);
        }
      }
      for (int j = 129 + lane; j <= n_rb; j += 32) {
        const float w = __fdiv_rn(__fmul_rn((float)j, PI_F), r_cut);
        const float v = (d == 0.f) ? w : __fdiv_rn(sinf(__fmul_rn(w, d)), d);
        out[j - 1] = __fmul_rn(pref, v);
      }
    }
  }
}

// edges = |dr| (graph_utils.py:298) or e3nn.bessel(|dr|, n_rb, r_max) (:301), reference rounding points
__global__ void edge_feat_kernel(const float* __restrict__ dr, int64_t n, int n_rb, float r_max, float pref,
                                 float* __restrict__ out) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  const float x = dr[3 * e], y = dr[3 * e + 1], z = dr[3 * e + 2];
  const float d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
  if (n_rb == 0) {
    out[e] = d;
    return;
  }
  const float PI_F = 3.14159265358979323846f;
  for (int j = 1; j <= n_rb; ++j) {
    const float w = __fdiv_rn(__fmul_rn((float)j, PI_F), r_max);
    const float v = (d == 0.f) ? w : __fdiv_rn(sinf(__fmul_rn(w, d)), d);
    out[e * n_rb + j - 1] = __fmul_rn(pref, v);
  }
}

}  // namespace

extern "C" int eqnn_edge_features(const float* dr, int64_t n_edges, int n_rb, double r_max, float* out,
                                  eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dr && out && n_edges >= 0 && n_rb >= 0, "edge_features: bad argument");
  EQNN_CHECK_ARG(n_rb == 0 || r_max > 0.0, "edge_features: r_max must be positive");
  if (n_edges == 0) return EQNN_OK;
  const float pref = n_rb > 0 ? (float)sqrt(2.0 / r_max) : 1.f;
  edge_feat_kernel<<<(unsigned)ceil_div64(n_edges, 256), 256, 0, as_stream(stream)>>>(dr, n_edges, n_rb, (float)r_max,
                                                                                     pref, out);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

// out[i] = sum_{j < i} count[j], i = 0..n (n + 1 outputs): the row pointers of a ragged edge list
extern "C" int eqnn_exclusive_scan_i32(const int32_t* count, int64_t n, int32_t* out, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(count && out && n >= 0, "exclusive_scan: bad argument");
  scan_kernel<<<1, 1024, 0, as_stream(stream)>>>(count, n, out, nullptr);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

// out[0..n] = exclusive scan of count[0..n) (+ a copy in cursor); scratch: eqnn_scan_scratch_bytes(n), 4-byte aligned.
// Long arrays take the multi-CTA route; also used by other translation units (pairs.cu).
size_t eqnn_scan_scratch_bytes(int64_t n) { return align_up((size_t)(ceil_div64(n, SCAN_CHUNK) + 1) * 2 * 4, 256); }
int eqnn_scan_rows_i32(const int32_t* count, int64_t n, int32_t* out, int32_t* cursor, void* scratch, cudaStream_t st) {
  const int64_t n_chunks = ceil_div64(n, SCAN_CHUNK);
  if (!scratch || n_chunks < 4 || n_chunks > 1024 * SCAN_PER_THREAD * 64) {
    scan_kernel<<<1, 1024, 0, st>>>(count, n, out, cursor);
    EQNN_CHECK_LAUNCH();
    return EQNN_OK;
  }
  int32_t* totals = reinterpret_cast<int32_t*>(scratch);
  int32_t* offs = totals + n_chunks;
  scan_totals_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(count, n, totals);
  scan_kernel<<<1, 1024, 0, st>>>(totals, n_chunks, offs, nullptr);
  scan_apply_kernel<<<(unsigned)n_chunks, 256, 0, st>>>(count, n, offs, (int)n_chunks, out, cursor);
  EQNN_CHECK_LAUNCH_N(3);
  return EQNN_OK;
}

extern "C" size_t eqnn_csr_workspace_bytes(int B, int N, int E) {
  return align_up(((size_t)B * N + 1) * 4, 256) * 2 + align_up((size_t)B * E * 4, 256) + eqnn_scan_scratch_bytes((int64_t)B * N);
}

extern "C" int eqnn_csr_build(const int32_t* senders, const int32_t* receivers, int B, int N, int E,
                              int32_t* rowptr, int32_t* perm, int32_t* src, void* ws, size_t ws_bytes,
                              eqnn_stream_t stream) {
  EQNN_CHECK_ARG(senders && receivers && rowptr && perm && src, "csr: null pointer");
  EQNN_CHECK_ARG(B >= 1 && N >= 1 && E >= 0, "csr: bad shape");
  EQNN_CHECK_ARG((int64_t)B * E < ((int64_t)1 << 31) && (int64_t)B * N < ((int64_t)1 << 31), "csr: index overflow");
  EQNN_CHECK_ARG(ws && ws_bytes >= eqnn_csr_workspace_bytes(B, N, E), "csr: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int64_t n_rows = (int64_t)B * N, n_edges = (int64_t)B * E;
  const size_t seg = align_up((n_rows + 1) * 4, 256);
  int32_t* count = reinterpret_cast<int32_t*>(ws);
  int32_t* cursor = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + seg);
  int32_t* perm_tmp = reinterpret_cast<int32_t*>(reinterpret_cast<char*>(ws) + 2 * seg);
  EQNN_CUDA(cudaMemsetAsync(count, 0, (n_rows + 1) * 4, st));
  if (n_edges > 0) hist_kernel<<<(unsigned)ceil_div64(n_edges, 256), 256, 0, st>>>(receivers, n_edges, E, N, count);
  void* scan_scratch = reinterpret_cast<char*>(ws) + 2 * seg + align_up((size_t)n_edges * 4, 256);
  {
    const int rc = eqnn_scan_rows_i32(count, n_rows, rowptr, cursor, scan_scratch, st);
    if (rc) return rc;
  }
  if (n_edges > 0) {
    fill_kernel<<<(unsigned)ceil_div64(n_edges, 256), 256, 0, st>>>(receivers, n_edges, E, N, cursor, perm);
    rowsort_kernel<<<(unsigned)ceil_div64(n_rows * 32, 256), 256, 0, st>>>(rowptr, n_rows, senders, E, N, perm,
                                                                         perm_tmp, src);
  }
  EQNN_CHECK_LAUNCH_N(n_edges > 0 ? 3 : 0);
  return EQNN_OK;
}

extern "C" int eqnn_edge_geom(const float* pos, int ld_pos, int n_nodes, const int32_t* rowptr,
                              const int32_t* src, int n_rb, double r_cut, float* geom, float* radial,
                              eqnn_stream_t stream) {
  EQNN_CHECK_ARG(pos && rowptr && src && geom, "edge_geom: null pointer");
  EQNN_CHECK_ARG(ld_pos >= 3 && n_nodes >= 0 && n_rb >= 0, "edge_geom: bad shape");
  EQNN_CHECK_ARG(n_rb == 0 || r_cut > 0.0, "edge_geom: r_cut must be positive");
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(geom) & 15) == 0, "edge_geom: geom must be 16-byte aligned");
  if (n_nodes == 0) return EQNN_OK;
  // jnp.sqrt(2.0 / x_max) is a Python double folded to fp32 (weak type)
  const float pref = n_rb > 0 ? (float)sqrt(2.0 / r_cut) : 1.f;
  geom_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      pos, ld_pos, n_nodes, rowptr, src, n_rb, (float)r_cut, pref, reinterpret_cast<float4*>(geom), radial);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}
