// EGNN edge geometry, coordinate update and their backward passes (sm_100a).
//
// Reference: models/egnn.py (positions-only layer): :82-90 r_ij = x_i - x_j + 1e-7 (+PBC), |r_ij|, e3nn.bessel;
// :107-114 x_ij = clip(r_ij * [tanh](t), +-100); jraph segment_{sum,mean} over receivers; :146-150
// x' = PBC(x + agg).  The two 128-wide edge MLPs in between (phi_e, phi_x) run on the tensor-core GEMMs of
// tc_mlp.cu.  All kernels walk the receiver-sorted CSR: one warp per receiver row, lanes over its edges,
// shuffle reductions -- no atomics, deterministic.
#include "common.cuh"
#include "gen/tp_gen.cuh"

namespace {

struct Cell3 {
  float c[9], ci[9];
  int pbc, diag;
};

__device__ __forceinline__ void wrap(float& x, float& y, float& z, const Cell3& cl) {
  if (!cl.pbc) return;
  if (cl.diag) {
    x -= rintf(x * cl.ci[0]) * cl.c[0];
    y -= rintf(y * cl.ci[4]) * cl.c[4];
    z -= rintf(z * cl.ci[8]) * cl.c[8];
  } else {
    const float n0 = rintf(x * cl.ci[0] + y * cl.ci[3] + z * cl.ci[6]);
    const float n1 = rintf(x * cl.ci[1] + y * cl.ci[4] + z * cl.ci[7]);
    const float n2 = rintf(x * cl.ci[2] + y * cl.ci[5] + z * cl.ci[8]);
    x -= n0 * cl.c[0] + n1 * cl.c[3] + n2 * cl.c[6];
    y -= n0 * cl.c[1] + n1 * cl.c[4] + n2 * cl.c[7];
    z -= n0 * cl.c[2] + n1 * cl.c[5] + n2 * cl.c[8];
  }
}

// rij4[p] = (r_ij, |r_ij|) in CSR order; feats[p, :] = bessel(|r_ij|) (or |r_ij| if n_rb == 0)
__global__ void __launch_bounds__(256) egnn_geom_fwd_kernel(const float* __restrict__ pos, int ld, int n_nodes,
                                                            const int32_t* __restrict__ rowptr,
                                                            const int32_t* __restrict__ src, Cell3 cl, int n_rb,
                                                            float r_max, float pref, float4* __restrict__ rij4,
                                                            float* __restrict__ feats) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float xr = pos[(size_t)row * ld], yr = pos[(size_t)row * ld + 1], zr = pos[(size_t)row * ld + 2];
  for (int p = p0 + lane; p < p1; p += 32) {
    const int s = src[p];
    float rx = (pos[(size_t)s * ld] - xr) + 1e-7f;
    float ry = (pos[(size_t)s * ld + 1] - yr) + 1e-7f;
    float rz = (pos[(size_t)s * ld + 2] - zr) + 1e-7f;
    wrap(rx, ry, rz, cl);
    const float d = sqrtf(rx * rx + ry * ry + rz * rz);
    rij4[p] = make_float4(rx, ry, rz, d);
    if (n_rb == 0) feats[p] = d;
  }
  if (n_rb > 0) {
    const float PI_F = 3.14159265358979323846f;
    for (int p = p0; p < p1; ++p) {
      const float d = rij4[p].w;
      for (int j = 1 + lane; j <= n_rb; j += 32) {
        const float w = (float)j * PI_F / r_max;
        feats[(size_t)p * n_rb + j - 1] = pref * ((d == 0.f) ? w : sinf(w * d) / d);
      }
    }
  }
}

// x_ij = clip(r_ij * t', +-100), t' = tanh(t) or t; agg over the row; pos_out = wrap(pos + agg)
__global__ void __launch_bounds__(256) egnn_coord_fwd_kernel(const float4* __restrict__ rij4,
                                                             const float* __restrict__ t, int use_tanh, int agg,
                                                             const int32_t* __restrict__ rowptr, int n_nodes,
                                                             const float* __restrict__ pos, int ld, Cell3 cl,
                                                             float* __restrict__ pos_out, int ld_out,
                                                             int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  float ax = 0.f, ay = 0.f, az = 0.f;
  if (agg == 2) {
    // jraph.segment_max of the clipped translations, component by component; the first edge attaining a maximum wins
    float m[3] = {-INFINITY, -INFINITY, -INFINITY};
    int am[3] = {0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int p = p0 + lane; p < p1; p += 32) {
      const float4 r = rij4[p];
      float tt = t[p];
      if (use_tanh) tt = tanhf(tt);
      const float v[3] = {fminf(fmaxf(r.x * tt, -100.f), 100.f), fminf(fmaxf(r.y * tt, -100.f), 100.f),
                          fminf(fmaxf(r.z * tt, -100.f), 100.f)};
#pragma unroll
      for (int c = 0; c < 3; ++c)
        if (v[c] > m[c] || am[c] == 0x7fffffff) {
          m[c] = v[c];
          am[c] = p;
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float mo = __shfl_xor_sync(0xffffffffu, m[c], o);
        const int ao = __shfl_xor_sync(0xffffffffu, am[c], o);
        if (ao != 0x7fffffff && (am[c] == 0x7fffffff || mo > m[c] || (mo == m[c] && ao < am[c]))) {
          m[c] = mo;
          am[c] = ao;
        }
      }
    if (lane == 0) {
#pragma unroll
      for (int c = 0; c < 3; ++c) argmax[(size_t)row * 3 + c] = am[c] == 0x7fffffff ? -1 : am[c];
    }
    ax = am[0] == 0x7fffffff ? 0.f : m[0];
    ay = am[1] == 0x7fffffff ? 0.f : m[1];
    az = am[2] == 0x7fffffff ? 0.f : m[2];
  } else {
    for (int p = p0 + lane; p < p1; p += 32) {
      const float4 r = rij4[p];
      float tt = t[p];
      if (use_tanh) tt = tanhf(tt);
      ax += fminf(fmaxf(r.x * tt, -100.f), 100.f);
      ay += fminf(fmaxf(r.y * tt, -100.f), 100.f);
      az += fminf(fmaxf(r.z * tt, -100.f), 100.f);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ax += __shfl_xor_sync(0xffffffffu, ax, o);
      ay += __shfl_xor_sync(0xffffffffu, ay, o);
      az += __shfl_xor_sync(0xffffffffu, az, o);
    }
  }
  if (lane == 0) {
    const float inv = (agg == 1) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
    float x = pos[(size_t)row * ld] + ax * inv, y = pos[(size_t)row * ld + 1] + ay * inv,
          z = pos[(size_t)row * ld + 2] + az * inv;
    wrap(x, y, z, cl);
    pos_out[(size_t)row * ld_out] = x;
    pos_out[(size_t)row * ld_out + 1] = y;
    pos_out[(size_t)row * ld_out + 2] = z;
  }
}

// Backward of the coordinate update and of the geometry, fused:
//   g = dL/dpos_out[row] (wrap has unit Jacobian almost everywhere), d x_ij = g / deg (mean) ,
//   dt[p]  = [1 - tanh^2] * sum_c r_c g_c mask_c ,   dr = t' g mask + (dL/d|r|) r/|r| ,
//   dL/d|r| = sum_j dfeats[p,j] * d bessel_j / d|r|        (dfeats may be NULL: first sweep computes dt only)
//   dpos_in[row] = g - sum_p dr[p]   (receiver end);  dre[perm[p]] = dr[p]   (sender end, gathered afterwards)
// mode 0: write dt only (phase A, before the MLP backward); mode 1: needs dfeats, writes dre and dpos_in.
__global__ void __launch_bounds__(256) egnn_coord_bwd_kernel(const float4* __restrict__ rij4,
                                                             const float* __restrict__ t, int use_tanh, int agg,
                                                             const int32_t* __restrict__ rowptr,
                                                             const int32_t* __restrict__ perm, int n_nodes,
                                                             const float* __restrict__ dpos_out, int n_rb, float r_max,
                                                             float pref, const float* __restrict__ dfeats, int mode,
                                                             float* __restrict__ dt, float4* __restrict__ dre,
                                                             float* __restrict__ dpos_in,
                                                             const int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float inv = (agg == 1) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
  const float gx = dpos_out[(size_t)row * 3] * inv, gy = dpos_out[(size_t)row * 3 + 1] * inv,
              gz = dpos_out[(size_t)row * 3 + 2] * inv;
  const float PI_F = 3.14159265358979323846f;
  float sx = 0.f, sy = 0.f, sz = 0.f;
  for (int pb = p0; pb < p1; pb += 32) {
    const int p = pb + lane;
    const bool on = p < p1;
    float4 r = make_float4(0.f, 0.f, 0.f, 1.f);
    float tt = 0.f, tp = 0.f;
    if (on) {
      r = rij4[p];
      tt = t[p];
      tp = use_tanh ? tanhf(tt) : tt;
    }
    float mx = (fabsf(r.x * tp) <= 100.f) ? 1.f : 0.f, my = (fabsf(r.y * tp) <= 100.f) ? 1.f : 0.f,
          mz = (fabsf(r.z * tp) <= 100.f) ? 1.f : 0.f;
    if (agg == 2) {                                   // segment_max: only the winning edge of a component sees its gradient
      if (argmax[(size_t)row * 3] != p) mx = 0.f;
      if (argmax[(size_t)row * 3 + 1] != p) my = 0.f;
      if (argmax[(size_t)row * 3 + 2] != p) mz = 0.f;
    }
    if (mode == 0) {
      if (on) {
        float d = r.x * gx * mx + r.y * gy * my + r.z * gz * mz;
        if (use_tanh) d *= (1.f - tp * tp);
        dt[p] = d;
      }
      continue;
    }
    // dL/d|r| from the radial features: lanes cooperate over the basis index for each edge of this batch
    float dd = 0.f;
    const int nb = min(32, p1 - pb);
    for (int q = 0; q < nb; ++q) {
      const float dq = __shfl_sync(0xffffffffu, r.w, q);
      float part = 0.f;
      if (n_rb == 0) {
        if (lane == 0) part = dfeats[pb + q];
      } else {
        for (int j = 1 + lane; j <= n_rb; j += 32) {
          const float w = (float)j * PI_F / r_max;
          float sn, cs;
          sincosf(w * dq, &sn, &cs);
          const float db = (dq == 0.f) ? 0.f : pref * (w * cs * dq - sn) / (dq * dq);
          part += dfeats[(size_t)(pb + q) * n_rb + j - 1] * db;
        }
      }
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
      if (lane == q) dd = part;
    }
    if (on) {
      const float invd = (r.w > 0.f) ? 1.f / r.w : 0.f;
      const float dx = tp * gx * mx + dd * r.x * invd, dy = tp * gy * my + dd * r.y * invd,
                  dz = tp * gz * mz + dd * r.z * invd;
      dre[perm[p]] = make_float4(dx, dy, dz, 0.f);
      sx += dx; sy += dy; sz += dz;
    }
  }
  if (mode == 0) return;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sx += __shfl_xor_sync(0xffffffffu, sx, o);
    sy += __shfl_xor_sync(0xffffffffu, sy, o);
    sz += __shfl_xor_sync(0xffffffffu, sz, o);
  }
  if (lane == 0) {
    dpos_in[(size_t)row * 3] = dpos_out[(size_t)row * 3] - sx;
    dpos_in[(size_t)row * 3 + 1] = dpos_out[(size_t)row * 3 + 1] - sy;
    dpos_in[(size_t)row * 3 + 2] = dpos_out[(size_t)row * 3 + 2] - sz;
  }
}

// dpos[s, c] += sum over the edges sent by s of dre[e, c]   (sender-sorted CSR, c < 3)
__global__ void egnn_sender_acc_kernel(const float4* __restrict__ dre, const int32_t* __restrict__ rowptr_s,
                                       const int32_t* __restrict__ perm_s, int n_nodes, float* __restrict__ dpos) {
  const int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_nodes) return;
  float ax = 0.f, ay = 0.f, az = 0.f;
  for (int q = rowptr_s[s]; q < rowptr_s[s + 1]; ++q) {
    const float4 v = dre[perm_s[q]];
    ax += v.x; ay += v.y; az += v.z;
  }
  dpos[s * 3] += ax;
  dpos[s * 3 + 1] += ay;
  dpos[s * 3 + 2] += az;
}

// two-stage column sum for tall matrices (bias gradients): partial[cta][n] then fixed-order reduce
__global__ void __launch_bounds__(256) colsum_partial_kernel(const float* __restrict__ x, int64_t M, int N,
                                                             int64_t rows_per_cta, float* __restrict__ partial) {
  __shared__ float sm[8][129];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 8 row-lanes x 32 column-lanes (x4 columns each)
  const int64_t r0 = (int64_t)blockIdx.x * rows_per_cta, r1 = min(M, r0 + rows_per_cta);
  for (int c0 = 0; c0 < N; c0 += 128) {
    float a[4] = {0.f, 0.f, 0.f, 0.f};
    const int c = c0 + tx * 4;
    if (c + 3 < N && (N & 3) == 0) {
      for (int64_t r = r0 + ty; r < r1; r += 8) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(x + r * N + c));
        a[0] += v.x; a[1] += v.y; a[2] += v.z; a[3] += v.w;
      }
    } else {
      for (int64_t r = r0 + ty; r < r1; r += 8)
        for (int q = 0; q < 4; ++q)
          if (c + q < N) a[q] += x[r * N + c + q];
    }
    for (int q = 0; q < 4; ++q) sm[ty][tx * 4 + q] = a[q];
    __syncthreads();
    if (threadIdx.x < 128 && c0 + threadIdx.x < N) {
      float s = 0.f;
      for (int q = 0; q < 8; ++q) s += sm[q][threadIdx.x];
      partial[(size_t)blockIdx.x * N + c0 + threadIdx.x] = s;
    }
    __syncthreads();
  }
}

__global__ void colsum_final_kernel(const float* __restrict__ partial, int n_parts, int N, float* __restrict__ y) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int q = 0; q < n_parts; ++q) s += partial[(size_t)q * N + c];
  y[c] = s;
}

__global__ void softgate_fwd_kernel(const float4* __restrict__ z, const float4* __restrict__ za, int64_t n4,
                                    float4* __restrict__ m) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = z[i], b = za[i];
  m[i] = make_float4(gelu_tanh(a.x) * sigmoid_acc(b.x), gelu_tanh(a.y) * sigmoid_acc(b.y),
                     gelu_tanh(a.z) * sigmoid_acc(b.z), gelu_tanh(a.w) * sigmoid_acc(b.w));
}

__global__ void softgate_bwd_kernel(const float4* __restrict__ z, const float4* __restrict__ za,
                                    const float4* __restrict__ dm, int64_t n4, float4* __restrict__ dza,
                                    float4* __restrict__ g1) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 a = z[i], b = za[i], d = dm[i];
  const float s0 = sigmoid_acc(b.x), s1 = sigmoid_acc(b.y), s2 = sigmoid_acc(b.z), s3 = sigmoid_acc(b.w);
  dza[i] = make_float4(d.x * gelu_tanh(a.x) * s0 * (1.f - s0), d.y * gelu_tanh(a.y) * s1 * (1.f - s1),
                       d.z * gelu_tanh(a.z) * s2 * (1.f - s2), d.w * gelu_tanh(a.w) * s3 * (1.f - s3));
  g1[i] = make_float4(d.x * s0, d.y * s1, d.z * s2, d.w * s3);
}

int make_cell(Cell3& cl, const float* cell, const float* cell_inv) {
  cl.pbc = (cell != nullptr) ? 1 : 0;
  cl.diag = 1;
  for (int i = 0; i < 9; ++i) {
    cl.c[i] = cell ? cell[i] : ((i % 4 == 0) ? 1.f : 0.f);
    cl.ci[i] = cell_inv ? cell_inv[i] : ((i % 4 == 0) ? 1.f : 0.f);
    if (cell && i % 4 != 0 && (cell[i] != 0.f || cell_inv[i] != 0.f)) cl.diag = 0;
  }
  EQNN_CHECK_ARG(!cell || cell_inv, "egnn: cell given without cell_inv");
  return EQNN_OK;
}

constexpr int COLSUM_PARTS = 296;


// ---- node layouts with velocities and scalar attributes, nodes[n] = (x, v, h) (egnn.py:38-60, 198-229) ----

// edge-MLP input U[p] = [feats[p] | h[sender(p)] | h[receiver(p)]] in receiver-sorted order: warp per receiver row
__global__ void __launch_bounds__(256) egnn_edge_input_fwd_kernel(const float* __restrict__ feats, int nf,
                                                                  const float* __restrict__ h, int ld_h, int dh,
                                                                  const float* __restrict__ glob, int n_glob,
                                                                  int nodes_per_graph,
                                                                  const int32_t* __restrict__ rowptr,
                                                                  const int32_t* __restrict__ src, int n_nodes,
                                                                  float* __restrict__ U) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int K0 = nf + 2 * dh + n_glob;
  const float* gl = n_glob ? glob + (size_t)(row / nodes_per_graph) * n_glob : nullptr;   // jraph broadcasts the globals
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float* hr = h + (size_t)row * ld_h;
  for (int p = p0; p < p1; ++p) {                      // lanes run along the columns of one edge row
    const float* hs = h + (size_t)src[p] * ld_h;
    const float* fp = feats + (size_t)p * nf;
    float* u = U + (size_t)p * K0;
    for (int c = lane; c < K0; c += 32) {
      float v;
      if (c < nf) v = fp[c];
      else if (c < nf + dh) v = hs[c - nf];
      else if (c < nf + 2 * dh) v = hr[c - nf - dh];
      else v = gl[c - nf - 2 * dh];
      u[c] = v;
    }
  }
}

// backward of the concatenation: dfeats (copy), receiver part summed over the row (+= into dh_out), sender part
// stored at the ORIGINAL edge id for the sender-sorted gather that follows
__global__ void __launch_bounds__(256) egnn_edge_input_bwd_kernel(const float* __restrict__ dU, int nf, int dh, int n_glob,
                                                                  const int32_t* __restrict__ rowptr,
                                                                  const int32_t* __restrict__ perm, int n_nodes,
                                                                  float* __restrict__ dfeats, float* __restrict__ dhs_e,
                                                                  float* __restrict__ dh_out, int ld_dh) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int K0 = nf + 2 * dh + n_glob;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  // lanes run along the columns (coalesced row segments of dU), the row's edges are walked in order
  for (int p = p0; p < p1; ++p) {
    const float* u = dU + (size_t)p * K0;
    if (dfeats)
      for (int c = lane; c < nf; c += 32) dfeats[(size_t)p * nf + c] = u[c];
    float* ds = dhs_e + (size_t)perm[p] * dh;
    for (int c = lane; c < dh; c += 32) ds[c] = u[nf + c];
  }
  for (int c = lane; c < dh; c += 32) {                // receiver end: fixed-order sum over the row's edges
    float a = 0.f;
    for (int p = p0; p < p1; ++p) a += dU[(size_t)p * K0 + nf + dh + c];
    dh_out[(size_t)row * ld_dh + c] += a;
  }
}

// out[s, :w] += sum over the edges sent by s of rows[e, :w]   (sender-sorted CSR; fixed order)
__global__ void egnn_segment_gather_add_kernel(const float* __restrict__ rows, int w, const int32_t* __restrict__ rowptr_s,
                                               const int32_t* __restrict__ perm_s, int n_nodes, float* __restrict__ out,
                                               int ld_out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * w) return;
  const int64_t s = i / w;
  const int c = (int)(i - s * w);
  float a = 0.f;
  for (int q = rowptr_s[s]; q < rowptr_s[s + 1]; ++q) a += rows[(size_t)perm_s[q] * w + c];
  out[(size_t)s * ld_out + c] += a;
}

// base[n] = sx[n] * x[n] + sv[n] * v[n]   (sx == NULL: 1, coupled update egnn.py:223-224; else :216-221)
__global__ void egnn_xv_base_kernel(const float* __restrict__ nodes, int F, const float* __restrict__ sx,
                                    const float* __restrict__ sv, int64_t n, float* __restrict__ base) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * 3) return;
  const int64_t node = i / 3;
  const int c = (int)(i - node * 3);
  const float a = sx ? sx[node] : 1.f;
  base[i] = a * nodes[node * F + c] + sv[node] * nodes[node * F + 3 + c];
}

// nodes_next[n, 3:] = (pass-through columns [3, h_off) (the old velocity, :229), h + ph); x' is already written
__global__ void egnn_node_assemble_kernel(const float* __restrict__ nodes, int F, int h_off, const float* __restrict__ ph,
                                          int64_t n, float* __restrict__ nodes_next) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int w = F - 3, dh = F - h_off;
  if (i >= n * w) return;
  const int64_t node = i / w;
  const int c = 3 + (int)(i - node * w);
  float v = nodes[node * F + c];
  if (c >= h_off) v += ph[node * dh + (c - h_off)];
  nodes_next[node * F + c] = v;
}

// aggregated messages m_i[r, :] = agg_{e->r} act(rows[e, :]) (jraph segment_sum / segment_mean of m_ij, egnn.py:141,
// read by the (x, h) and (x, v) node functions): warp per receiver row, lanes over the d columns, edges in CSR order
__global__ void __launch_bounds__(256) egnn_segment_rows_fwd_kernel(const float* __restrict__ rows, int d, int act,
                                                                    const int32_t* __restrict__ rowptr, int n_nodes,
                                                                    int agg, float* __restrict__ out, int ld_out,
                                                                    int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float inv = (agg == 1) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
  for (int c = lane; c < d; c += 32) {
    if (agg == 2) {
      // jraph.segment_max: the maximum and the CSR position of the first edge attaining it (gradient routing)
      float a = -INFINITY;
      int am = -1;
      for (int p = p0; p < p1; ++p) {
        float v = rows[(size_t)p * d + c];
        if (act) v = gelu_tanh(v);
        if (v > a || am < 0) {
          a = v;
          am = p;
        }
      }
      out[(size_t)row * ld_out + c] = am < 0 ? 0.f : a;
      argmax[(size_t)row * d + c] = am;
      continue;
    }
    float a = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float v = rows[(size_t)p * d + c];
      a += act ? gelu_tanh(v) : v;
    }
    out[(size_t)row * ld_out + c] = a * inv;
  }
}

// its backward: target[e, :] += dmi[r(e), :] / deg * (act ? gelu'(z[e, :]) : 1)
__global__ void __launch_bounds__(256) egnn_segment_rows_bwd_kernel(const float* __restrict__ dmi, int ld_dmi, int d,
                                                                    const float* __restrict__ z, int act,
                                                                    const int32_t* __restrict__ rowptr, int n_nodes,
                                                                    int agg, float* __restrict__ target,
                                                                    const int32_t* __restrict__ argmax) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float inv = (agg == 1) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
  for (int c = lane; c < d; c += 32) {
    const float g = dmi[(size_t)row * ld_dmi + c] * inv;
    if (agg == 2) {                                   // segment_max: the winning edge takes the whole gradient
      const int p = argmax[(size_t)row * d + c];
      if (p >= 0) {
        const size_t i = (size_t)p * d + c;
        target[i] += act ? g * gelu_tanh_grad(z[i]) : g;
      }
      continue;
    }
    for (int p = p0; p < p1; ++p) {
      const size_t i = (size_t)p * d + c;
      target[i] += act ? g * gelu_tanh_grad(z[i]) : g;
    }
  }
}

// backward of the node update given g = dL/dx' (cols 0..2 of dnext), the pass-through gradients of v and h in dnext,
// and mix = g + geometry gradient of x (NULL in the first step: no input gradient needed, din may then be NULL):
//   dsx = g.x, dsv = g.v ; din = (sx g + (mix - g), sv g + dnext_v, dnext_h)
__global__ void egnn_xv_bwd_kernel(const float* __restrict__ nodes, int F, const float* __restrict__ sx,
                                   const float* __restrict__ sv, const float* __restrict__ dnext, int ld_dnext,
                                   const float* __restrict__ mix, int64_t n, float* __restrict__ din,
                                   float* __restrict__ dsx, float* __restrict__ dsv) {
  const int64_t node = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (node >= n) return;
  const float* nd = nodes + node * F;
  const float* dn = dnext + node * ld_dnext;
  float gx = 0.f, gv = 0.f;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    gx = fmaf(dn[c], nd[c], gx);
    gv = fmaf(dn[c], nd[3 + c], gv);
  }
  if (dsx) dsx[node] = gx;
  dsv[node] = gv;
  if (din) {
    float* o = din + node * F;
    const float a = sx ? sx[node] : 1.f, b = sv[node];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float g = dn[c];
      o[c] = a * g + (mix ? (mix[node * 3 + c] - g) : 0.f);
      o[3 + c] = b * g + dn[3 + c];
    }
    for (int c = 6; c < F; ++c) o[c] = dn[c];
  }
}

// dst[n, c0d + j] (+)= src[n / group, c0s + j], j < w  (strided column block copy; group > 1 broadcasts one source
// row to `group` consecutive destination rows: graph globals -> nodes)
__global__ void copy_cols_kernel(const float* __restrict__ src, int ld_s, int c0s, int w, int64_t n, int group,
                                 int accumulate, float* __restrict__ dst, int ld_d, int c0d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * w) return;
  const int64_t r = i / w;
  const int j = (int)(i - r * w);
  const float v = src[(r / group) * ld_s + c0s + j];
  float* d = dst + r * ld_d + c0d + j;
  *d = accumulate ? *d + v : v;
}


// ---- steerable graph attributes (models/utils/equivariant_graph_utils.py:27-102) -----------------------------
struct ShScale { float f[5]; };      // e3nn `normalization=` factor per degree l

// per edge, in the reference's edge order: r_ij = x[sender] - x[receiver] (minimum image, no EPS; :46-52),
// edge attributes = normalisation * Y_lm(r_ij / |r_ij|) with e3nn's guard for |r| = 0 (:56-61)
template <class SH>
__global__ void steerable_edge_kernel(const float* __restrict__ pos, int ld, int N, int E, int64_t n_edges,
                                      const int32_t* __restrict__ senders, const int32_t* __restrict__ receivers,
                                      Cell3 cl, ShScale sc, float* __restrict__ r_ij, float* __restrict__ edge_attrs) {
  const int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n_edges) return;
  const int64_t b = e / E;
  const float* xs = pos + (size_t)(b * N + senders[e]) * ld;
  const float* xr = pos + (size_t)(b * N + receivers[e]) * ld;
  float rx = xs[0] - xr[0], ry = xs[1] - xr[1], rz = xs[2] - xr[2];
  wrap(rx, ry, rz, cl);
  r_ij[e * 3] = rx; r_ij[e * 3 + 1] = ry; r_ij[e * 3 + 2] = rz;
  const float n2 = rx * rx + ry * ry + rz * rz;
  const float n = sqrtf(n2 == 0.f ? 1.f : n2);
  float Y[SH::NY];
  SH::sh(rx / n, ry / n, rz / n, Y);
#pragma unroll
  for (int l = 0; l <= SH::LMAX; ++l)
#pragma unroll
    for (int m = 0; m < 2 * l + 1; ++m) edge_attrs[e * SH::NY + l * l + m] = sc.f[l] * Y[l * l + m];
}

// node attributes = e3nn.scatter_mean of the edge attributes over the receivers (:64-74; empty rows -> 0),
// plus the harmonics of the node velocities when steerable_velocities (:76-83).  perm: original edge id per CSR slot
template <class SH>
__global__ void steerable_node_kernel(const float* __restrict__ edge_attrs, const int32_t* __restrict__ rowptr,
                                      const int32_t* __restrict__ perm, int n_nodes, const float* __restrict__ vel,
                                      int ld_vel, ShScale sc, float* __restrict__ node_attrs) {
  const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  float a[SH::NY];
#pragma unroll
  for (int i = 0; i < SH::NY; ++i) a[i] = 0.f;
  const int p0 = rowptr[n], p1 = rowptr[n + 1];
  for (int q = p0; q < p1; ++q) {
    const float* row = edge_attrs + (size_t)perm[q] * SH::NY;
#pragma unroll
    for (int i = 0; i < SH::NY; ++i) a[i] += row[i];
  }
  const float cnt = (float)max(p1 - p0, 1);
#pragma unroll
  for (int i = 0; i < SH::NY; ++i) a[i] = a[i] / cnt;
  if (vel) {
    const float vx = vel[n * ld_vel], vy = vel[n * ld_vel + 1], vz = vel[n * ld_vel + 2];
    const float n2 = vx * vx + vy * vy + vz * vz;
    const float nn = sqrtf(n2 == 0.f ? 1.f : n2);
    float Y[SH::NY];
    SH::sh(vx / nn, vy / nn, vz / nn, Y);
#pragma unroll
    for (int l = 0; l <= SH::LMAX; ++l)
#pragma unroll
      for (int m = 0; m < 2 * l + 1; ++m) a[l * l + m] += sc.f[l] * Y[l * l + m];
  }
#pragma unroll
  for (int i = 0; i < SH::NY; ++i) node_attrs[n * SH::NY + i] = a[i];
}

template <class SH>
void launch_steerable(const float* pos, int ld_pos, const float* vel, int ld_vel, int B, int N, int E,
                      const int32_t* senders, const int32_t* receivers, const int32_t* rowptr, const int32_t* perm,
                      const Cell3& cl, const ShScale& sc, float* r_ij, float* edge_attrs, float* node_attrs,
                      cudaStream_t st) {
  const int64_t n_edges = (int64_t)B * E;
  steerable_edge_kernel<SH><<<(unsigned)ceil_div64(n_edges, 256), 256, 0, st>>>(pos, ld_pos, N, E, n_edges, senders,
                                                                                receivers, cl, sc, r_ij, edge_attrs);
  steerable_node_kernel<SH><<<(unsigned)ceil_div64((int64_t)B * N, 256), 256, 0, st>>>(edge_attrs, rowptr, perm, B * N, vel,
                                                                                     ld_vel, sc, node_attrs);
}


// ---- SEGNN TensorProductLinearGate (models/segnn.py:30-64), dense-lowered -----------------------------------
// Linear(TP(x, y)) is bilinear in (x, y): z[r, o] = b[o] + sum_j y[r, j] * (x[r, :] @ M_j)[o].  The GEMM produces
// Xbig[r, j * Do + o] = (x @ [M_0 | M_1 | ...])[r, j * Do + o]; these two kernels contract / expand the y axis.
__global__ void segnn_ycontract_kernel(const float* __restrict__ Xbig, const float* __restrict__ y, int ld_y,
                                       const float* __restrict__ bias, int bias_col, int n_bias, int64_t R, int Dy,
                                       int Do, float* __restrict__ z) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Do) return;
  const int64_t r = i / Do;
  const int o = (int)(i - r * Do);
  float a = (bias && o >= bias_col && o < bias_col + n_bias) ? bias[o - bias_col] : 0.f;
  const float* xb = Xbig + (r * (size_t)Do + o) * Dy;        // the Dy attribute components of one output are adjacent
  for (int j = 0; j < Dy; ++j) a = fmaf(y ? y[r * ld_y + j] : 1.f, xb[j], a);
  z[i] = a;
}

__global__ void segnn_yexpand_kernel(const float* __restrict__ dz, const float* __restrict__ y, int ld_y, int64_t R,
                                     int Dy, int Do, float* __restrict__ dXbig) {
  // one thread per (row, output): dz is read once and fans out to the Dy adjacent attribute columns
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= R * Do) return;
  const int64_t r = i / Do;
  const int o = (int)(i - r * Do);
  const float v = dz[i];
  float* dst = dXbig + ((size_t)r * Do + o) * Dy;
  const float* yr = y ? y + (size_t)r * ld_y : nullptr;
  for (int j = 0; j < Dy; ++j) dst[j] = (yr ? __ldg(yr + j) : 1.f) * v;
}

// Irrep-blocked TensorProductLinearGate: out[r, o] = bias + sum_t coef[t] * y[r, idx[t] >> 24] * in[r, idx[t] & 0xFFFFFF]
// over the CSR range ptr[o] .. ptr[o+1] (Clebsch-Gordan terms).  One warp per row: the input row and the attribute row
// are staged in shared memory, each lane owns the outputs o = lane, lane + 32, ...  The same kernel runs the forward
// (in = T, table sorted by Linear output) and the backward (in = dz, table sorted by T column).
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32)
segnn_cg_kernel(const float* __restrict__ in, int ld_in, int n_in, const float* __restrict__ y, int ld_y, int Dy,
                const int32_t* __restrict__ ptr, const int32_t* __restrict__ idx, const float* __restrict__ coef,
                const float* __restrict__ bias, int bias_col, int n_bias, int64_t R, int n_out,
                float* __restrict__ out, int ld_out) {
  extern __shared__ float cg_sm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* row = cg_sm + (size_t)warp * (n_in + 32);
  float* yr = row + n_in;
  for (int64_t r = (int64_t)blockIdx.x * WARPS + warp; r < R; r += (int64_t)gridDim.x * WARPS) {
    const float* src = in + (size_t)r * ld_in;
    for (int c = lane; c < n_in; c += 32) row[c] = __ldg(src + c);
    if (lane < Dy) yr[lane] = y ? __ldg(y + (size_t)r * ld_y + lane) : 1.f;
    __syncwarp();
    float* dst = out + (size_t)r * ld_out;
    for (int o = lane; o < n_out; o += 32) {
      float acc = (bias != nullptr && o >= bias_col && o < bias_col + n_bias) ? __ldg(bias + (o - bias_col)) : 0.f;
      const int t1 = __ldg(ptr + o + 1);
      for (int t = __ldg(ptr + o); t < t1; ++t) {
        const int q = __ldg(idx + t);
        acc = fmaf(__ldg(coef + t) * yr[q >> 24], row[q & 0xFFFFFF], acc);
      }
      dst[o] = acc;
    }
    __syncwarp();
  }
}

// ---- first edge layer of a SEGNN message-passing step, evaluated at the nodes ------------------------------------
// Its input is concat(add[e], h[sender], h[receiver]) (segnn.py:111-113), so x @ M_cat splits into
// add @ M_a + (h @ M_s)[sender] + (h @ M_r)[receiver]: the two big products are formed once per NODE (P_s, P_r:
// [n_nodes][Do*Dy], 1/k of the per-edge work) and the per-edge part is a gather + attribute contraction:
//   z[p, o] = bias + sum_j y[p, j] (P_r[row, o*Dy+j] + P_s[src[p], o*Dy+j] + sum_a add[p, a] M_a[a, o*Dy+j])
// Warp per receiver row (edges in receiver-sorted order), lanes along the outputs o.
__global__ void __launch_bounds__(256) segnn_edge0_fwd_kernel(const float* __restrict__ Ps, const float* __restrict__ Pr,
                                                              const float* __restrict__ add, int n_add,
                                                              const float* __restrict__ Ma, const float* __restrict__ y,
                                                              int Dy, int Do, const float* __restrict__ bias, int bias_col,
                                                              int n_bias, const int32_t* __restrict__ rowptr,
                                                              const int32_t* __restrict__ src, int n_nodes,
                                                              float* __restrict__ z) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int nc = Do * Dy;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float* pr = Pr + (size_t)row * nc;
  if (Dy == 4 && ((reinterpret_cast<uintptr_t>(Ps) | reinterpret_cast<uintptr_t>(Pr) | reinterpret_cast<uintptr_t>(Ma) |
                   reinterpret_cast<uintptr_t>(y)) & 15) == 0) {
    // attributes l <= 1: the four attribute columns of an output are one 16-byte load; the receiver's own row of
    // P_r stays in registers while the row's edges are walked
    for (int o = lane; o < Do; o += 32) {
      const float b = (bias != nullptr && o >= bias_col && o < bias_col + n_bias) ? __ldg(bias + (o - bias_col)) : 0.f;
      const float4 r4 = __ldg(reinterpret_cast<const float4*>(pr) + o);
      for (int p = p0; p < p1; ++p) {
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(Ps + (size_t)src[p] * nc) + o);
        const float4 y4 = __ldg(reinterpret_cast<const float4*>(y) + p);
        float4 t = make_float4(r4.x + s4.x, r4.y + s4.y, r4.z + s4.z, r4.w + s4.w);
        for (int a = 0; a < n_add; ++a) {
          const float f = __ldg(add + (size_t)p * n_add + a);
          const float4 m4 = __ldg(reinterpret_cast<const float4*>(Ma + (size_t)a * nc) + o);
          t.x = fmaf(f, m4.x, t.x); t.y = fmaf(f, m4.y, t.y); t.z = fmaf(f, m4.z, t.z); t.w = fmaf(f, m4.w, t.w);
        }
        z[(size_t)p * Do + o] = fmaf(y4.w, t.w, fmaf(y4.z, t.z, fmaf(y4.y, t.y, fmaf(y4.x, t.x, b))));
      }
    }
    return;
  }
  for (int p = p0; p < p1; ++p) {
    const float* ps = Ps + (size_t)src[p] * nc;
    const float* yp = y + (size_t)p * Dy;
    const float* ap = add + (size_t)p * n_add;
    for (int o = lane; o < Do; o += 32) {
      float acc = (bias != nullptr && o >= bias_col && o < bias_col + n_bias) ? __ldg(bias + (o - bias_col)) : 0.f;
      for (int j = 0; j < Dy; ++j) {
        const int c = o * Dy + j;
        float t = pr[c] + ps[c];
        for (int a = 0; a < n_add; ++a) t = fmaf(__ldg(ap + a), __ldg(Ma + (size_t)a * nc + c), t);
        acc = fmaf(__ldg(yp + j), t, acc);
      }
      z[(size_t)p * Do + o] = acc;
    }
  }
}

// its backward onto the node-level products: with G[p, o*Dy+j] = y[p, j] dz[p, o] (never written),
//   dPr[n, c] = sum over the edges received by n of G[p, c],   dPs[n, c] = sum over the edges sent by n of G[spos[q], c]
// (spos: sender-sorted position -> receiver-sorted position).  Thread per (node, column), fixed order.
__global__ void segnn_edge0_bwd_kernel(const float* __restrict__ dz, const float* __restrict__ y, int Dy, int Do,
                                       const int32_t* __restrict__ rowptr, const int32_t* __restrict__ rowptr_s,
                                       const int32_t* __restrict__ spos, int n_nodes, float* __restrict__ dPs,
                                       float* __restrict__ dPr) {
  const int nc = Do * Dy;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * nc) return;
  const int64_t n = i / nc;
  const int c = (int)(i - n * nc);
  const int o = c / Dy, j = c - o * Dy;
  float a = 0.f;
  for (int p = rowptr[n]; p < rowptr[n + 1]; ++p) a = fmaf(__ldg(y + (size_t)p * Dy + j), __ldg(dz + (size_t)p * Do + o), a);
  dPr[i] = a;
  a = 0.f;
  for (int q = rowptr_s[n]; q < rowptr_s[n + 1]; ++q) {
    const int p = spos[q];
    a = fmaf(__ldg(y + (size_t)p * Dy + j), __ldg(dz + (size_t)p * Do + o), a);
  }
  dPs[i] = a;
}

// Dy = 4: warp per node, lanes along the outputs o; dz is read once per (edge, o) and the four attribute values of the
// edge are one broadcast 16-byte load; the four columns of an output are written as one 16-byte store
__global__ void __launch_bounds__(256) segnn_edge0_bwd4_kernel(const float* __restrict__ dz, const float* __restrict__ y,
                                                               int Do, const int32_t* __restrict__ rowptr,
                                                               const int32_t* __restrict__ rowptr_s,
                                                               const int32_t* __restrict__ spos, int n_nodes,
                                                               float* __restrict__ dPs, float* __restrict__ dPr) {
  const int lane = threadIdx.x & 31;
  const int64_t n = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (n >= n_nodes) return;
  const int p0 = rowptr[n], p1 = rowptr[n + 1], q0 = rowptr_s[n], q1 = rowptr_s[n + 1];
  const float4* y4 = reinterpret_cast<const float4*>(y);
  for (int o = lane; o < Do; o += 32) {
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = p0; p < p1; ++p) {
      const float d = __ldg(dz + (size_t)p * Do + o);
      const float4 yy = __ldg(y4 + p);
      a.x = fmaf(yy.x, d, a.x); a.y = fmaf(yy.y, d, a.y); a.z = fmaf(yy.z, d, a.z); a.w = fmaf(yy.w, d, a.w);
    }
    reinterpret_cast<float4*>(dPr + (size_t)n * Do * 4)[o] = a;
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int q = q0; q < q1; ++q) {
      const int p = spos[q];
      const float d = __ldg(dz + (size_t)p * Do + o);
      const float4 yy = __ldg(y4 + p);
      a.x = fmaf(yy.x, d, a.x); a.y = fmaf(yy.y, d, a.y); a.z = fmaf(yy.z, d, a.z); a.w = fmaf(yy.w, d, a.w);
    }
    reinterpret_cast<float4*>(dPs + (size_t)n * Do * 4)[o] = a;
  }
}

// gMa[a, o*Dy+j] = sum_p add[p, a] y[p, j] dz[p, o]  (n_add <= 8): per-CTA partials over edge ranges, then a fixed-order sum
constexpr int EDGE0_PARTS = 148 * 16;      // short edge ranges: the per-thread loop is a dependent chain of loads
__global__ void __launch_bounds__(1024) segnn_edge0_wadd_partial_kernel(const float* __restrict__ add, int n_add,
                                                                       const float* __restrict__ y, int Dy,
                                                                       const float* __restrict__ dz, int Do, int64_t E,
                                                                       int64_t per, float* __restrict__ part) {
  const int nc = Do * Dy;
  const int64_t e0 = (int64_t)blockIdx.x * per, e1 = (e0 + per < E) ? e0 + per : E;
  for (int c = threadIdx.x; c < nc; c += blockDim.x) {
    const int o = c / Dy, j = c - o * Dy;
    float acc[8];
#pragma unroll
    for (int a = 0; a < 8; ++a) acc[a] = 0.f;
#pragma unroll 4
    for (int64_t p = e0; p < e1; ++p) {
      const float g = __ldg(y + p * Dy + j) * __ldg(dz + p * Do + o);
#pragma unroll
      for (int a = 0; a < 8; ++a)
        if (a < n_add) acc[a] = fmaf(__ldg(add + p * n_add + a), g, acc[a]);
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
      if (a < n_add) part[((size_t)blockIdx.x * n_add + a) * nc + c] = acc[a];
  }
}
// Dy <= 4: thread per output o keeps all n_add x Dy accumulators, so dz is read once per edge (coalesced) and the
// attribute / message values are warp-wide broadcasts
__global__ void __launch_bounds__(256) segnn_edge0_wadd_partial4_kernel(const float* __restrict__ add, int n_add,
                                                                        const float* __restrict__ y, int Dy,
                                                                        const float* __restrict__ dz, int Do, int64_t E,
                                                                        int64_t per, float* __restrict__ part) {
  const int nc = Do * Dy;
  const int64_t e0 = (int64_t)blockIdx.x * per, e1 = (e0 + per < E) ? e0 + per : E;
  for (int o = threadIdx.x; o < Do; o += blockDim.x) {
    float acc[8][4];
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[a][j] = 0.f;
#pragma unroll 2
    for (int64_t p = e0; p < e1; ++p) {
      const float d = __ldg(dz + p * Do + o);
      float g[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) g[j] = (j < Dy) ? __ldg(y + p * Dy + j) * d : 0.f;
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        if (a < n_add) {
          const float f = __ldg(add + p * n_add + a);
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[a][j] = fmaf(f, g[j], acc[a][j]);
        }
      }
    }
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int j = 0; j < 4; ++j)
        if (a < n_add && j < Dy) part[((size_t)blockIdx.x * n_add + a) * nc + o * Dy + j] = acc[a][j];
  }
}
__global__ void segnn_edge0_wadd_final_kernel(const float* __restrict__ part, int parts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float t = 0.f;
  for (int q = 0; q < parts; ++q) t += part[(size_t)q * n + i];
  out[i] = t;
}

// inv[perm[p]] = p ; spos[q] = inv[perm_s[q]]
__global__ void perm_invert_kernel(const int32_t* __restrict__ perm, int64_t n, int32_t* __restrict__ inv) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) inv[perm[i]] = (int32_t)i;
}
__global__ void perm_gather_kernel(const int32_t* __restrict__ inv, const int32_t* __restrict__ perm_s, int64_t n,
                                   int32_t* __restrict__ spos) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) spos[i] = inv[perm_s[i]];
}

// ---- last edge layer of a SEGNN message-passing step, evaluated at the nodes ------------------------------------
// The last edge layer has no gate (segnn.py:66-77) and its output is only aggregated over the receivers (:218), so
//   agg[n, o] = f sum_{p in n} (b[o] + sum_j y[p, j] (m[p] @ M_cat)[o*Dy+j]) = sum_j (S_j[n] @ M_j)[o] + b[o] w(n)
// with S_j[n, k] = f sum_{p in n} y[p, j] m[p, k] (f = 1 for sum, 1/deg for mean; w(n) = f deg) and
// M_j[k, o] = M_cat[k, o*Dy+j]: one y-weighted segment reduction per edge, the GEMMs run on the node rows.
constexpr int YAGG_MAX_DY = 16;
template <int YAGG_DY>      // attribute slots kept in registers: 4 (l <= 1), 9 (l <= 2) or 16 (l <= 3); Dy <= YAGG_DY
__global__ void __launch_bounds__(256) segnn_yagg_fwd_kernel(const float* __restrict__ m, int Dx, const float* __restrict__ y,
                                                             int Dy, const int32_t* __restrict__ rowptr, int n_nodes, int agg,
                                                             float* __restrict__ S) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float inv = (agg == 1) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
  for (int k = lane; k < Dx; k += 32) {
    float acc[YAGG_DY];
#pragma unroll
    for (int j = 0; j < YAGG_DY; ++j) acc[j] = 0.f;
    for (int p = p0; p < p1; ++p) {
      const float v = m[(size_t)p * Dx + k];
      const float* yp = y + (size_t)p * Dy;
#pragma unroll
      for (int j = 0; j < YAGG_DY; ++j)
        if (j < Dy) acc[j] = fmaf(__ldg(yp + j), v, acc[j]);
    }
#pragma unroll
    for (int j = 0; j < YAGG_DY; ++j)
      if (j < Dy) S[((size_t)j * n_nodes + row) * Dx + k] = acc[j] * inv;
  }
}

// dm[p, k] = f sum_j y[p, j] dS_j[n(p), k]
template <int YAGG_DY>
__global__ void __launch_bounds__(256) segnn_yagg_bwd_kernel(const float* __restrict__ dS, int Dx, const float* __restrict__ y,
                                                             int Dy, const int32_t* __restrict__ rowptr, int n_nodes, int agg,
                                                             float* __restrict__ dm) {
  const int lane = threadIdx.x & 31;
  const int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (row >= n_nodes) return;
  const int p0 = rowptr[row], p1 = rowptr[row + 1];
  const float inv = (agg == 1) ? 1.f / (float)max(p1 - p0, 1) : 1.f;
  for (int k = lane; k < Dx; k += 32) {
    float g[YAGG_DY];
#pragma unroll
    for (int j = 0; j < YAGG_DY; ++j) g[j] = (j < Dy) ? dS[((size_t)j * n_nodes + row) * Dx + k] * inv : 0.f;
    for (int p = p0; p < p1; ++p) {
      const float* yp = y + (size_t)p * Dy;
      float a = 0.f;
#pragma unroll
      for (int j = 0; j < YAGG_DY; ++j)
        if (j < Dy) a = fmaf(__ldg(yp + j), g[j], a);
      dm[(size_t)p * Dx + k] = a;
    }
  }
}

// out[n, c] += b[c] w(n) for the bias columns;  w(n) = deg (sum) or [deg > 0] (mean)
__global__ void segnn_bias_deg_fwd_kernel(float* __restrict__ out, int ld, const float* __restrict__ bias, int n_bias,
                                          const int32_t* __restrict__ rowptr, int n_nodes, int agg) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)n_nodes * n_bias) return;
  const int64_t n = i / n_bias;
  const int c = (int)(i - n * n_bias);
  const int deg = rowptr[n + 1] - rowptr[n];
  const float w = (agg == 1) ? (deg > 0 ? 1.f : 0.f) : (float)deg;
  out[(size_t)n * ld + c] += __ldg(bias + c) * w;
}

// gb[c] = sum_n w(n) d[n, c]: one block per bias column, fixed-order tree
__global__ void __launch_bounds__(256) segnn_bias_deg_bwd_kernel(const float* __restrict__ d, int ld,
                                                                 const int32_t* __restrict__ rowptr, int n_nodes, int agg,
                                                                 float* __restrict__ gb) {
  __shared__ float red[256];
  const int c = blockIdx.x;
  float a = 0.f;
  for (int n = threadIdx.x; n < n_nodes; n += 256) {
    const int deg = rowptr[n + 1] - rowptr[n];
    const float w = (agg == 1) ? (deg > 0 ? 1.f : 0.f) : (float)deg;
    a = fmaf(w, d[(size_t)n * ld + c], a);
  }
  red[threadIdx.x] = a;
  __syncthreads();
  for (int s2 = 128; s2 > 0; s2 >>= 1) {
    if ((int)threadIdx.x < s2) red[threadIdx.x] += red[threadIdx.x + s2];
    __syncthreads();
  }
  if (threadIdx.x == 0) gb[c] = red[0];
}

// dst[p, :] = src[perm[p], :]
__global__ void gather_rows_kernel(const float* __restrict__ src, const int32_t* __restrict__ perm, int64_t n, int w,
                                   float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * w) return;
  const int64_t p = i / w;
  dst[i] = src[(size_t)perm[p] * w + (i - p * w)];
}

// y[i] += a * x[i]
__global__ void axpy_kernel(int64_t n, float a, const float* __restrict__ x, float* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = fmaf(a, x[i], y[i]);
}

}  // namespace

extern "C" int eqnn_egnn_geom_fwd(const float* pos, int ld_pos, int n_nodes, const int32_t* rowptr,
                                  const int32_t* src, const float* cell, const float* cell_inv, int n_rb, double r_max,
                                  float* rij4, float* feats, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(pos && rowptr && src && rij4 && feats && ld_pos >= 3, "egnn_geom_fwd: bad argument");
  EQNN_CHECK_ARG(((uintptr_t)rij4 & 15) == 0, "egnn_geom_fwd: rij4 must be 16-byte aligned");
  EQNN_CHECK_ARG(n_rb == 0 || r_max > 0.0, "egnn_geom_fwd: r_max must be positive");
  if (n_nodes <= 0) return EQNN_OK;
  Cell3 cl;
  int rc = make_cell(cl, cell, cell_inv);
  if (rc) return rc;
  const float pref = n_rb > 0 ? (float)sqrt(2.0 / r_max) : 1.f;
  egnn_geom_fwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      pos, ld_pos, n_nodes, rowptr, src, cl, n_rb, (float)r_max, pref, reinterpret_cast<float4*>(rij4), feats);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_coord_fwd(const float* rij4, const float* t, int use_tanh, int agg, const int32_t* rowptr,
                                   int n_nodes, const float* pos, int ld_pos, const float* cell, const float* cell_inv,
                                   float* pos_out, int ld_out, int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rij4 && t && rowptr && pos && pos_out && ld_pos >= 3 && ld_out >= 3, "egnn_coord_fwd: bad argument");
  EQNN_CHECK_ARG(agg >= 0 && agg <= 2 && (agg != 2 || argmax), "egnn_coord_fwd: agg must be 0 (sum), 1 (mean) or 2 (max, with the argmax buffer)");
  if (n_nodes <= 0) return EQNN_OK;
  Cell3 cl;
  int rc = make_cell(cl, cell, cell_inv);
  if (rc) return rc;
  egnn_coord_fwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(rij4), t, use_tanh, agg, rowptr, n_nodes, pos, ld_pos, cl, pos_out, ld_out, argmax);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_coord_bwd(const float* rij4, const float* t, int use_tanh, int agg, const int32_t* rowptr,
                                   const int32_t* perm, int n_nodes, const float* dpos_out, int n_rb, double r_max,
                                   const float* dfeats, float* dt, float* dre, float* dpos_in, const int32_t* argmax,
                                   eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rij4 && t && rowptr && dpos_out, "egnn_coord_bwd: bad argument");
  EQNN_CHECK_ARG(agg >= 0 && agg <= 2 && (agg != 2 || argmax), "egnn_coord_bwd: agg must be 0 (sum), 1 (mean) or 2 (max, with the argmax buffer)");
  const int mode = dfeats ? 1 : 0;
  EQNN_CHECK_ARG(mode == 1 ? (perm && dre && dpos_in) : (dt != nullptr), "egnn_coord_bwd: missing output for this phase");
  if (n_nodes <= 0) return EQNN_OK;
  const float pref = n_rb > 0 ? (float)sqrt(2.0 / r_max) : 1.f;
  egnn_coord_bwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(rij4), t, use_tanh, agg, rowptr, perm, n_nodes, dpos_out, n_rb, (float)r_max, pref,
      dfeats, mode, dt, reinterpret_cast<float4*>(dre), dpos_in, argmax);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_sender_acc(const float* dre, const int32_t* rowptr_s, const int32_t* perm_s, int n_nodes,
                                    float* dpos, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dre && rowptr_s && perm_s && dpos, "egnn_sender_acc: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  egnn_sender_acc_kernel<<<(unsigned)ceil_div64(n_nodes, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(dre), rowptr_s, perm_s, n_nodes, dpos);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_edge_input_fwd(const float* feats, int nf, const float* h, int ld_h, int dh, const float* glob,
                                        int n_glob, int nodes_per_graph, const int32_t* rowptr, const int32_t* src,
                                        int n_nodes, float* U, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(feats && rowptr && src && U && nf >= 1 && dh >= 0 && (dh == 0 || (h && ld_h >= dh)),
                 "egnn_edge_input_fwd: bad argument");
  EQNN_CHECK_ARG(n_glob >= 0 && (n_glob == 0 || (glob && nodes_per_graph >= 1)), "egnn_edge_input_fwd: bad globals");
  if (n_nodes <= 0) return EQNN_OK;
  egnn_edge_input_fwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      feats, nf, h, ld_h, dh, glob, n_glob, nodes_per_graph, rowptr, src, n_nodes, U);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_edge_input_bwd(const float* dU, int nf, int dh, int n_glob, const int32_t* rowptr, const int32_t* perm,
                                        const int32_t* rowptr_s, const int32_t* perm_s, int n_nodes, float* dfeats,
                                        float* dhs_e, float* dh_out, int ld_dh, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dU && rowptr && perm && rowptr_s && perm_s && dhs_e && dh_out && nf >= 1 && dh >= 1 && ld_dh >= dh,
                 "egnn_edge_input_bwd: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  cudaStream_t st = as_stream(stream);
  egnn_edge_input_bwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, st>>>(
      dU, nf, dh, n_glob, rowptr, perm, n_nodes, dfeats, dhs_e, dh_out, ld_dh);
  egnn_segment_gather_add_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * dh, 256), 256, 0, st>>>(
      dhs_e, dh, rowptr_s, perm_s, n_nodes, dh_out, ld_dh);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" int eqnn_egnn_xv_base(const float* nodes, int F, const float* sx, const float* sv, int64_t n, float* base,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(nodes && sv && base && F >= 6, "egnn_xv_base: bad argument");
  if (n <= 0) return EQNN_OK;
  egnn_xv_base_kernel<<<(unsigned)ceil_div64(n * 3, 256), 256, 0, as_stream(stream)>>>(nodes, F, sx, sv, n, base);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_node_assemble(const float* nodes, int F, int h_off, const float* ph, int64_t n,
                                       float* nodes_next, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(nodes && nodes_next && h_off >= 3 && h_off <= F && (ph || h_off == F), "egnn_node_assemble: bad argument");
  if (n <= 0 || F == 3) return EQNN_OK;
  egnn_node_assemble_kernel<<<(unsigned)ceil_div64(n * (F - 3), 256), 256, 0, as_stream(stream)>>>(nodes, F, h_off, ph, n,
                                                                                                 nodes_next);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_segment_rows_fwd(const float* rows, int d, int act, const int32_t* rowptr, int n_nodes, int agg,
                                          float* out, int ld_out, int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(rows && rowptr && out && d >= 1 && ld_out >= d && agg >= 0 && agg <= 2 && (agg != 2 || argmax),
                 "egnn_segment_rows_fwd: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  egnn_segment_rows_fwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      rows, d, act, rowptr, n_nodes, agg, out, ld_out, argmax);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_segment_rows_bwd(const float* dmi, int ld_dmi, int d, const float* z, int act,
                                          const int32_t* rowptr, int n_nodes, int agg, float* target,
                                          const int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dmi && rowptr && target && d >= 1 && ld_dmi >= d && agg >= 0 && agg <= 2 && (z || !act) &&
                     (agg != 2 || argmax), "egnn_segment_rows_bwd: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  egnn_segment_rows_bwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      dmi, ld_dmi, d, z, act, rowptr, n_nodes, agg, target, argmax);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_egnn_xv_bwd(const float* nodes, int F, const float* sx, const float* sv, const float* dnext,
                                int ld_dnext, const float* mix, int64_t n, float* din, float* dsx, float* dsv,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(nodes && sv && dnext && dsv && F >= 6 && ld_dnext >= F && (!sx || dsx), "egnn_xv_bwd: bad argument");
  if (n <= 0) return EQNN_OK;
  egnn_xv_bwd_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(nodes, F, sx, sv, dnext, ld_dnext, mix,
                                                                                 n, din, dsx, dsv);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_copy_cols(const float* src, int ld_src, int col_src, int width, int64_t n, int group, int accumulate,
                              float* dst, int ld_dst, int col_dst, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(src && dst && width >= 1 && col_src >= 0 && col_dst >= 0 && ld_src >= col_src + width &&
                     ld_dst >= col_dst + width && group >= 1, "copy_cols: bad argument");
  if (n <= 0) return EQNN_OK;
  copy_cols_kernel<<<(unsigned)ceil_div64(n * width, 256), 256, 0, as_stream(stream)>>>(src, ld_src, col_src, width, n,
                                                                                       group, accumulate, dst, ld_dst,
                                                                                       col_dst);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_steerable_attrs(const float* pos, int ld_pos, const float* vel, int ld_vel, int B, int N, int E,
                                    const int32_t* senders, const int32_t* receivers, const int32_t* rowptr,
                                    const int32_t* perm, const float* cell, const float* cell_inv, int lmax,
                                    int sh_norm, float* r_ij, float* edge_attrs, float* node_attrs,
                                    eqnn_stream_t stream) {
  EQNN_CHECK_ARG(pos && senders && receivers && rowptr && perm && r_ij && edge_attrs && node_attrs && ld_pos >= 3,
                 "steerable_attrs: bad argument");
  EQNN_CHECK_ARG(lmax >= 1 && lmax <= 3, "steerable_attrs: lmax_attributes must be 1..3");
  EQNN_CHECK_ARG(sh_norm >= 0 && sh_norm <= 2, "steerable_attrs: sh_norm must be 0 (norm), 1 (component) or 2 (integral)");
  EQNN_CHECK_ARG(!vel || ld_vel >= 3, "steerable_attrs: bad velocity stride");
  if (B <= 0 || N <= 0) return EQNN_OK;
  Cell3 cl;
  int rc = make_cell(cl, cell, cell_inv);
  if (rc) return rc;
  ShScale sc;
  for (int l = 0; l < 5; ++l)
    sc.f[l] = sh_norm == 0 ? 1.f : (sh_norm == 1 ? sqrtf(2.f * l + 1.f) : (float)sqrt((2.0 * l + 1.0) / (4.0 * 3.14159265358979323846)));
  cudaStream_t st = as_stream(stream);
  if (lmax == 1) launch_steerable<SH_1>(pos, ld_pos, vel, ld_vel, B, N, E, senders, receivers, rowptr, perm, cl, sc, r_ij, edge_attrs, node_attrs, st);
  else if (lmax == 2) launch_steerable<SH_2>(pos, ld_pos, vel, ld_vel, B, N, E, senders, receivers, rowptr, perm, cl, sc, r_ij, edge_attrs, node_attrs, st);
  else launch_steerable<SH_3>(pos, ld_pos, vel, ld_vel, B, N, E, senders, receivers, rowptr, perm, cl, sc, r_ij, edge_attrs, node_attrs, st);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" int eqnn_segnn_ycontract(const float* Xbig, const float* y, int ld_y, const float* bias, int bias_col,
                                    int n_bias, int64_t R, int Dy, int Do, float* z, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(Xbig && z && Dy >= 1 && Do >= 1 && (y || Dy == 1) && (!y || ld_y >= Dy) && n_bias >= 0 && bias_col >= 0 &&
                     bias_col + n_bias <= Do, "segnn_ycontract: bad argument");
  if (R <= 0) return EQNN_OK;
  segnn_ycontract_kernel<<<(unsigned)ceil_div64(R * Do, 256), 256, 0, as_stream(stream)>>>(Xbig, y, ld_y, bias, bias_col,
                                                                                         n_bias, R, Dy, Do, z);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_yexpand(const float* dz, const float* y, int ld_y, int64_t R, int Dy, int Do, float* dXbig,
                                  eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dz && dXbig && Dy >= 1 && Do >= 1 && (y || Dy == 1) && (!y || ld_y >= Dy), "segnn_yexpand: bad argument");
  if (R <= 0) return EQNN_OK;
  segnn_yexpand_kernel<<<(unsigned)ceil_div64(R * Do, 256), 256, 0, as_stream(stream)>>>(dz, y, ld_y, R, Dy, Do, dXbig);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_cg_apply(const float* in, int ld_in, int n_in, const float* y, int ld_y, int Dy,
                                   const int32_t* ptr, const int32_t* idx, const float* coef, const float* bias,
                                   int bias_col, int n_bias, int64_t R, int n_out, float* out, int ld_out,
                                   eqnn_stream_t stream) {
  EQNN_CHECK_ARG(in && ptr && idx && coef && out && n_in >= 1 && n_in < (1 << 24) && ld_in >= n_in && n_out >= 1 &&
                     ld_out >= n_out && Dy >= 1 && Dy <= 32 && (y || Dy == 1) && (!y || ld_y >= Dy) && n_bias >= 0 &&
                     bias_col >= 0 && bias_col + n_bias <= n_out,
                 "segnn_cg_apply: bad argument");
  if (R <= 0) return EQNN_OK;
  constexpr int WARPS = 8;
  const size_t smem = (size_t)WARPS * (n_in + 32) * sizeof(float);
  EQNN_CHECK_ARG(smem <= 200 * 1024, "segnn_cg_apply: input row too wide for shared memory");
  if (smem > 48 * 1024)
    EQNN_CUDA(cudaFuncSetAttribute(segnn_cg_kernel<WARPS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int64_t want = ceil_div64(R, WARPS);
  const unsigned grid = (unsigned)(want < 148 * 8 ? want : 148 * 8);
  segnn_cg_kernel<WARPS><<<grid, WARPS * 32, smem, as_stream(stream)>>>(in, ld_in, n_in, y, ld_y, Dy, ptr, idx, coef, bias,
                                                                         bias_col, n_bias, R, n_out, out, ld_out);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_edge0_fwd(const float* Ps, const float* Pr, const float* add, int n_add, const float* Ma,
                                    const float* y, int Dy, int Do, const float* bias, int bias_col, int n_bias,
                                    const int32_t* rowptr, const int32_t* src, int n_nodes, float* z, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(Ps && Pr && y && rowptr && src && z && Dy >= 1 && Do >= 1 && n_add >= 0 && (n_add == 0 || (add && Ma)) &&
                     n_bias >= 0 && bias_col >= 0 && bias_col + n_bias <= Do && (n_bias == 0 || bias),
                 "segnn_edge0_fwd: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  segnn_edge0_fwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
      Ps, Pr, add, n_add, Ma, y, Dy, Do, n_bias ? bias : nullptr, bias_col, n_bias, rowptr, src, n_nodes, z);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_edge0_bwd(const float* dz, const float* y, int Dy, int Do, const int32_t* rowptr,
                                    const int32_t* rowptr_s, const int32_t* spos, int n_nodes, float* dPs, float* dPr,
                                    eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dz && y && rowptr && rowptr_s && spos && dPs && dPr && Dy >= 1 && Do >= 1, "segnn_edge0_bwd: bad argument");
  if (n_nodes <= 0) return EQNN_OK;
  if (Dy == 4 && ((reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(dPs) | reinterpret_cast<uintptr_t>(dPr)) & 15) == 0)
    segnn_edge0_bwd4_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * 32, 256), 256, 0, as_stream(stream)>>>(
        dz, y, Do, rowptr, rowptr_s, spos, n_nodes, dPs, dPr);
  else
    segnn_edge0_bwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * Do * Dy, 256), 256, 0, as_stream(stream)>>>(
        dz, y, Dy, Do, rowptr, rowptr_s, spos, n_nodes, dPs, dPr);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" size_t eqnn_segnn_edge0_wadd_workspace_bytes(int n_add, int Dy, int Do) {
  return (size_t)EDGE0_PARTS * n_add * Dy * Do * sizeof(float);
}

extern "C" int eqnn_segnn_edge0_wadd(const float* add, int n_add, const float* y, int Dy, const float* dz, int Do, int64_t E,
                                     float* gMa, void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(add && y && dz && gMa && n_add >= 1 && n_add <= 8 && Dy >= 1 && Do >= 1 && E >= 0,
                 "segnn_edge0_wadd: bad argument (1 <= n_add <= 8)");
  EQNN_CHECK_ARG(ws && ws_bytes >= eqnn_segnn_edge0_wadd_workspace_bytes(n_add, Dy, Do), "segnn_edge0_wadd: workspace too small");
  // the Dy <= 4 kernel keeps every load coalesced or broadcast: a quarter of the partials is enough to fill the GPU
  const int cap = (Dy <= 4) ? EDGE0_PARTS / 4 : EDGE0_PARTS;
  const int parts = (int)min((int64_t)cap, max((int64_t)1, (E + 63) / 64));
  const int64_t per = (E + parts - 1) / parts;
  const int n = n_add * Dy * Do;
  cudaStream_t st = as_stream(stream);
  const int nthr = (Dy * Do >= 1024) ? 1024 : (Dy * Do + 31) / 32 * 32;
  if (Dy <= 4)
    segnn_edge0_wadd_partial4_kernel<<<parts, Do >= 256 ? 256 : (Do + 31) / 32 * 32, 0, st>>>(
        add, n_add, y, Dy, dz, Do, E, per, reinterpret_cast<float*>(ws));
  else
    segnn_edge0_wadd_partial_kernel<<<parts, nthr, 0, st>>>(add, n_add, y, Dy, dz, Do, E, per, reinterpret_cast<float*>(ws));
  segnn_edge0_wadd_final_kernel<<<(n + 127) / 128, 128, 0, st>>>(reinterpret_cast<float*>(ws), parts, n, gMa);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" int eqnn_perm_compose(const int32_t* perm, const int32_t* perm_s, int64_t n, int32_t* inv, int32_t* spos,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(perm && perm_s && inv && spos && n >= 0, "perm_compose: bad argument");
  if (n == 0) return EQNN_OK;
  cudaStream_t st = as_stream(stream);
  perm_invert_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(perm, n, inv);
  perm_gather_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, st>>>(inv, perm_s, n, spos);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" int eqnn_segnn_yagg_fwd(const float* m, int Dx, const float* y, int Dy, const int32_t* rowptr, int n_nodes,
                                   int agg, float* S, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(m && y && rowptr && S && Dx >= 1 && Dy >= 1 && Dy <= YAGG_MAX_DY && (agg == 0 || agg == 1),
                 "segnn_yagg_fwd: bad argument (Dy <= 16, agg 0 sum | 1 mean)");
  if (n_nodes <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)ceil_div64((int64_t)n_nodes * 32, 256);
  if (Dy <= 4) segnn_yagg_fwd_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(m, Dx, y, Dy, rowptr, n_nodes, agg, S);
  else if (Dy <= 9) segnn_yagg_fwd_kernel<9><<<grid, 256, 0, as_stream(stream)>>>(m, Dx, y, Dy, rowptr, n_nodes, agg, S);
  else segnn_yagg_fwd_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(m, Dx, y, Dy, rowptr, n_nodes, agg, S);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_yagg_bwd(const float* dS, int Dx, const float* y, int Dy, const int32_t* rowptr, int n_nodes,
                                   int agg, float* dm, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dS && y && rowptr && dm && Dx >= 1 && Dy >= 1 && Dy <= YAGG_MAX_DY && (agg == 0 || agg == 1),
                 "segnn_yagg_bwd: bad argument (Dy <= 16, agg 0 sum | 1 mean)");
  if (n_nodes <= 0) return EQNN_OK;
  const unsigned grid = (unsigned)ceil_div64((int64_t)n_nodes * 32, 256);
  if (Dy <= 4) segnn_yagg_bwd_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(dS, Dx, y, Dy, rowptr, n_nodes, agg, dm);
  else if (Dy <= 9) segnn_yagg_bwd_kernel<9><<<grid, 256, 0, as_stream(stream)>>>(dS, Dx, y, Dy, rowptr, n_nodes, agg, dm);
  else segnn_yagg_bwd_kernel<16><<<grid, 256, 0, as_stream(stream)>>>(dS, Dx, y, Dy, rowptr, n_nodes, agg, dm);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_bias_deg_fwd(float* out, int ld, const float* bias, int n_bias, const int32_t* rowptr, int n_nodes,
                                       int agg, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(out && bias && rowptr && n_bias >= 0 && ld >= n_bias && (agg == 0 || agg == 1), "segnn_bias_deg_fwd: bad argument");
  if (n_nodes <= 0 || n_bias == 0) return EQNN_OK;
  segnn_bias_deg_fwd_kernel<<<(unsigned)ceil_div64((int64_t)n_nodes * n_bias, 256), 256, 0, as_stream(stream)>>>(
      out, ld, bias, n_bias, rowptr, n_nodes, agg);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_segnn_bias_deg_bwd(const float* d, int ld, const int32_t* rowptr, int n_nodes, int agg, int n_bias,
                                       float* gb, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(d && rowptr && gb && n_bias >= 0 && ld >= n_bias && n_nodes >= 0 && (agg == 0 || agg == 1),
                 "segnn_bias_deg_bwd: bad argument");
  if (n_bias == 0) return EQNN_OK;
  segnn_bias_deg_bwd_kernel<<<n_bias, 256, 0, as_stream(stream)>>>(d, ld, rowptr, n_nodes, agg, gb);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_gather_rows(const float* src, const int32_t* perm, int64_t n, int width, float* dst,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(src && perm && dst && width >= 1, "gather_rows: bad argument");
  if (n <= 0) return EQNN_OK;
  gather_rows_kernel<<<(unsigned)ceil_div64(n * width, 256), 256, 0, as_stream(stream)>>>(src, perm, n, width, dst);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_axpy(int64_t n, float a, const float* x, float* y, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && y, "axpy: null pointer");
  if (n <= 0) return EQNN_OK;
  axpy_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(n, a, x, y);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_softgate_fwd(const float* z, const float* za, int64_t n, float* m, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(z && za && m && n >= 0 && n % 4 == 0, "softgate_fwd: bad argument (n must be a multiple of 4)");
  if (n == 0) return EQNN_OK;
  softgate_fwd_kernel<<<(unsigned)ceil_div64(n / 4, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(za), n / 4, reinterpret_cast<float4*>(m));
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_softgate_bwd(const float* z, const float* za, const float* dm, int64_t n, float* dza, float* g1,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(z && za && dm && dza && g1 && n >= 0 && n % 4 == 0, "softgate_bwd: bad argument");
  if (n == 0) return EQNN_OK;
  softgate_bwd_kernel<<<(unsigned)ceil_div64(n / 4, 256), 256, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(z), reinterpret_cast<const float4*>(za), reinterpret_cast<const float4*>(dm), n / 4,
      reinterpret_cast<float4*>(dza), reinterpret_cast<float4*>(g1));
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" size_t eqnn_colsum_workspace_bytes(int N) { return (size_t)COLSUM_PARTS * N * sizeof(float); }

extern "C" int eqnn_colsum_tall(const float* x, int64_t M, int N, float* y, void* ws, size_t ws_bytes,
                                eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && y && M >= 0 && N >= 1, "colsum_tall: bad argument");
  EQNN_CHECK_ARG(ws && ws_bytes >= eqnn_colsum_workspace_bytes(N), "colsum_tall: workspace too small");
  EQNN_CHECK_ARG(((uintptr_t)x & 15) == 0, "colsum_tall: x must be 16-byte aligned");
  const int parts = (int)min((int64_t)COLSUM_PARTS, max((int64_t)1, (M + 255) / 256));
  const int64_t rpc = (M + parts - 1) / parts;
  cudaStream_t st = as_stream(stream);
  colsum_partial_kernel<<<parts, 256, 0, st>>>(x, M, N, rpc, reinterpret_cast<float*>(ws));
  colsum_final_kernel<<<(N + 127) / 128, 128, 0, st>>>(reinterpret_cast<float*>(ws), parts, N, y);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}
