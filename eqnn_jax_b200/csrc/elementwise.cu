// Small N-row / parameter-sized kernels around the interaction block (sm_100a):
// e3nn.gate, node pooling, column sums, Linear<->dense weight plumbing, MSE, AdamW.
#include "common.cuh"

namespace {

constexpr int MAX_GATE_CHUNKS = 8;
struct GateDesc {
  int n_extra, n_gates, n_chunks, d_gated;
  int act;                 // scalar activation: EQNN_ACT_GELU (tanh form) or EQNN_ACT_SILU (models/segnn.py:27,56)
  int mul[MAX_GATE_CHUNKS], dim[MAX_GATE_CHUNKS];
  float inv_c_act, inv_c_gate;
};

__device__ __forceinline__ float gate_act(float x, int kind) { return kind == EQNN_ACT_SILU ? x * sigmoid_acc(x) : gelu_tanh(x); }
__device__ __forceinline__ float gate_act_grad(float x, int kind) {
  if (kind == EQNN_ACT_SILU) {
    const float sg = sigmoid_acc(x);
    return sg + x * sg * (1.f - sg);
  }
  return gelu_tanh_grad(x);
}

// gated component j -> index of its gate
__device__ __forceinline__ int gate_index(const GateDesc& d, int j) {
  int g = 0;
#pragma unroll
  for (int c = 0; c < MAX_GATE_CHUNKS; ++c) {
    if (c < d.n_chunks) {
      const int w = d.mul[c] * d.dim[c];
      if (j < w) return g + j / d.dim[c];
      j -= w;
      g += d.mul[c];
    }
  }
  return g;
}

// e3nn.gate defaults (models/nequip.py:88-90): normalised gelu on the leading scalars,
// normalised sigmoid gates multiplying the l>0 irreps.
__global__ void gate_fwd_kernel(const float* __restrict__ z, int64_t n, GateDesc d, float* __restrict__ out) {
  const int d_in = d.n_extra + d.n_gates + d.d_gated, d_out = d.n_extra + d.d_gated;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d_out) return;
  const int64_t row = i / d_out;
  const int c = (int)(i - row * d_out);
  const float* zr = z + row * d_in;
  float v;
  if (c < d.n_extra) {
    v = gate_act(zr[c], d.act) * d.inv_c_act;
  } else {
    const int j = c - d.n_extra;
    const float g = sigmoid_acc(zr[d.n_extra + gate_index(d, j)]) * d.inv_c_gate;
    v = g * zr[d.n_extra + d.n_gates + j];
  }
  out[i] = v;
}

// Tall inputs (per-edge SEGNN layers): thread = column, a block walks GATE_ROWS rows, so the column's role (which
// gate, which components) is resolved once and no 64-bit division runs per element.  Same arithmetic as above.
constexpr int GATE_ROWS = 16;
__global__ void gate_fwd_rows_kernel(const float* __restrict__ z, int64_t n, GateDesc d, float* __restrict__ out) {
  const int d_in = d.n_extra + d.n_gates + d.d_gated, d_out = d.n_extra + d.d_gated;
  const int c = threadIdx.x;
  if (c >= d_out) return;
  const int j = c - d.n_extra;
  const int src = c < d.n_extra ? c : d.n_extra + d.n_gates + j;
  const int gate = c < d.n_extra ? 0 : d.n_extra + gate_index(d, j);
  const int64_t r0 = (int64_t)blockIdx.x * GATE_ROWS;
  const int64_t r1 = r0 + GATE_ROWS < n ? r0 + GATE_ROWS : n;
#pragma unroll 4
  for (int64_t r = r0; r < r1; ++r) {
    const float* zr = z + r * d_in;
    const float x = zr[src];
    out[r * d_out + c] = c < d.n_extra ? gate_act(x, d.act) * d.inv_c_act : (sigmoid_acc(zr[gate]) * d.inv_c_gate) * x;
  }
}

__global__ void gate_bwd_rows_kernel(const float* __restrict__ z, const float* __restrict__ gout, int64_t n, GateDesc d,
                                     float* __restrict__ gz) {
  const int d_in = d.n_extra + d.n_gates + d.d_gated, d_out = d.n_extra + d.d_gated;
  const int c = threadIdx.x;
  if (c >= d_in) return;
  int kind = 0, off = 0, dim = 1, gate = 0;      // 0 activated scalar | 1 gate scalar | 2 gated component
  if (c >= d.n_extra + d.n_gates) {
    kind = 2;
    off = c - d.n_extra - d.n_gates;
    gate = d.n_extra + gate_index(d, off);
  } else if (c >= d.n_extra) {
    kind = 1;
    int gi = c - d.n_extra;
#pragma unroll
    for (int k = 0; k < MAX_GATE_CHUNKS; ++k) {
      if (k < d.n_chunks) {
        if (gi >= 0 && gi < d.mul[k]) { off += gi * d.dim[k]; dim = d.dim[k]; gi = -1; }
        else if (gi >= 0) { gi -= d.mul[k]; off += d.mul[k] * d.dim[k]; }
      }
    }
  }
  const int64_t r0 = (int64_t)blockIdx.x * GATE_ROWS;
  const int64_t r1 = r0 + GATE_ROWS < n ? r0 + GATE_ROWS : n;
#pragma unroll 4
  for (int64_t r = r0; r < r1; ++r) {
    const float* zr = z + r * d_in;
    const float* gr = gout + r * d_out;
    float v;
    if (kind == 0) {
      v = gr[c] * gate_act_grad(zr[c], d.act) * d.inv_c_act;
    } else if (kind == 1) {
      float s = 0.f;
      for (int q = 0; q < dim; ++q) s += gr[d.n_extra + off + q] * zr[d.n_extra + d.n_gates + off + q];
      const float sg = sigmoid_acc(zr[c]);
      v = s * sg * (1.f - sg) * d.inv_c_gate;
    } else {
      const float g = sigmoid_acc(zr[gate]) * d.inv_c_gate;
      v = g * gr[d.n_extra + off];
    }
    gz[r * d_in + c] = v;
  }
}

// one thread per input element of z; gates sum over their (2l+1) components
__global__ void gate_bwd_kernel(const float* __restrict__ z, const float* __restrict__ gout, int64_t n, GateDesc d,
                                float* __restrict__ gz) {
  const int d_in = d.n_extra + d.n_gates + d.d_gated, d_out = d.n_extra + d.d_gated;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * d_in) return;
  const int64_t row = i / d_in;
  const int c = (int)(i - row * d_in);
  const float* zr = z + row * d_in;
  const float* gr = gout + row * d_out;
  float v;
  if (c < d.n_extra) {
    v = gr[c] * gate_act_grad(zr[c], d.act) * d.inv_c_act;
  } else if (c < d.n_extra + d.n_gates) {
    // gate number gi: find its chunk and component range
    int gi = c - d.n_extra, off = 0, dim = 1;
#pragma unroll
    for (int k = 0; k < MAX_GATE_CHUNKS; ++k) {
      if (k < d.n_chunks) {
        if (gi >= 0 && gi < d.mul[k]) { off += gi * d.dim[k]; dim = d.dim[k]; gi = -1; }
        else if (gi >= 0) { gi -= d.mul[k]; off += d.mul[k] * d.dim[k]; }
      }
    }
    float s = 0.f;
    for (int q = 0; q < dim; ++q) s += gr[d.n_extra + off + q] * zr[d.n_extra + d.n_gates + off + q];
    const float sg = sigmoid_acc(zr[c]);
    v = s * sg * (1.f - sg) * d.inv_c_gate;
  } else {
    const int j = c - d.n_extra - d.n_gates;
    const float g = sigmoid_acc(zr[d.n_extra + gate_index(d, j)]) * d.inv_c_gate;
    v = g * gr[d.n_extra + j];
  }
  gz[i] = v;
}

int make_gate_desc(GateDesc& d, int n_extra, int n_gates, int n_chunks, const int* mul, const int* l,
                   float inv_c_act, float inv_c_gate) {
  d.act = EQNN_ACT_GELU;
  EQNN_CHECK_ARG(n_chunks >= 0 && n_chunks <= MAX_GATE_CHUNKS, "gate: too many gated chunks");
  EQNN_CHECK_ARG(n_chunks == 0 || (mul && l), "gate: null chunk arrays");
  d.n_extra = n_extra; d.n_gates = n_gates; d.n_chunks = n_chunks; d.d_gated = 0;
  d.inv_c_act = inv_c_act; d.inv_c_gate = inv_c_gate;
  int tot = 0;
  for (int c = 0; c < MAX_GATE_CHUNKS; ++c) {
    d.mul[c] = c < n_chunks ? mul[c] : 0;
    d.dim[c] = c < n_chunks ? 2 * l[c] + 1 : 1;
    d.d_gated += d.mul[c] * d.dim[c];
    tot += d.mul[c];
  }
  EQNN_CHECK_ARG(tot == n_gates, "gate: number of gates != number of gated irreps");
  return EQNN_OK;
}

__global__ void pool_fwd_kernel(const float* __restrict__ x, int B, int N, int D, float scale,
                                float* __restrict__ out) {
  // grid (ceil(D/32), B), block (32, 8): threadIdx.y strides over nodes, then shared reduce
  __shared__ float sm[8][33];
  const int b = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < D)
    for (int n = threadIdx.y; n < N; n += 8) a += x[((size_t)b * N + n) * D + c];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sm[q][threadIdx.x];
    out[(size_t)b * D + c] = s * scale;
  }
}

__global__ void pool_bwd_kernel(const float* __restrict__ gout, int B, int N, int D, float scale,
                                float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * N * D) return;
  const int c = (int)(i % D);
  const int64_t b = i / ((int64_t)N * D);
  gx[i] = gout[b * D + c] * scale;
}

// jnp.max over the nodes of a graph (readout_agg = "max", models/nequip.py:195-199): value and the index of the FIRST
// node attaining it (the gradient is routed to that node alone)
__global__ void pool_max_fwd_kernel(const float* __restrict__ x, int B, int N, int D, float* __restrict__ out,
                                    int32_t* __restrict__ arg) {
  __shared__ float sv[8][33];
  __shared__ int si[8][33];
  const int b = blockIdx.y, c = blockIdx.x * 32 + threadIdx.x;
  float a = -INFINITY;
  int ia = -1;
  if (c < D)
    for (int n = threadIdx.y; n < N; n += 8) {
      const float v = x[((size_t)b * N + n) * D + c];
      if (v > a || ia < 0) {
        a = v;
        ia = n;
      }
    }
  sv[threadIdx.y][threadIdx.x] = a;
  si[threadIdx.y][threadIdx.x] = ia;
  __syncthreads();
  if (threadIdx.y == 0 && c < D) {
    float m = sv[0][threadIdx.x];
    int im = si[0][threadIdx.x];
#pragma unroll
    for (int q = 1; q < 8; ++q) {
      const float v = sv[q][threadIdx.x];
      const int iv = si[q][threadIdx.x];
      if (iv >= 0 && (im < 0 || v > m || (v == m && iv < im))) {
        m = v;
        im = iv;
      }
    }
    out[(size_t)b * D + c] = m;
    arg[(size_t)b * D + c] = im;
  }
}

__global__ void pool_max_bwd_kernel(const float* __restrict__ gout, const int32_t* __restrict__ arg, int B, int N, int D,
                                    float* __restrict__ gx) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * N * D) return;
  const int c = (int)(i % D);
  const int64_t bn = i / D;
  const int64_t b = bn / N;
  const int n = (int)(bn - b * N);
  gx[i] = (arg[b * D + c] == n) ? gout[b * D + c] : 0.f;
}

__global__ void colsum_kernel(const float* __restrict__ x, int64_t M, int N, float* __restrict__ y) {
  __shared__ float sm[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  float a = 0.f;
  if (c < N)
    for (int64_t m = threadIdx.y; m < M; m += 8) a += x[m * N + c];
  sm[threadIdx.y][threadIdx.x] = a;
  __syncthreads();
  if (threadIdx.y == 0 && c < N) {
    float s = 0.f;
#pragma unroll
    for (int q = 0; q < 8; ++q) s += sm[q][threadIdx.x];
    y[c] = s;
  }
}

__global__ void expand_kernel(const float* __restrict__ params, const int32_t* __restrict__ idx,
                              const float* __restrict__ scale, int64_t n, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int j = idx[i];
  dst[i] = j == -1 ? 0.f : (j == -2 ? scale[i] : scale[i] * params[j]);   // -1: structural zero, -2: constant
}

__global__ void contract_kernel(const float* __restrict__ ddst, const int32_t* __restrict__ pidx,
                                const int32_t* __restrict__ ptr, const int32_t* __restrict__ tidx,
                                const float* __restrict__ scale_t, int64_t n, float* __restrict__ grads) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  float s = 0.f;
  for (int t = ptr[j]; t < ptr[j + 1]; ++t) s += scale_t[t] * ddst[tidx[t]];
  grads[pidx[j]] = s;
}

__global__ void mse_kernel(const float* __restrict__ pred, const float* __restrict__ target, int n, float gscale,
                           float* __restrict__ loss, float* __restrict__ gpred) {
  // single CTA (n = B * n_outputs is tiny)
  __shared__ float sm[256];
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += 256) {
    const float d = pred[i] - target[i];
    a += d * d;
    if (gpred) gpred[i] = 2.f * d / (float)n * gscale;
  }
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sm[threadIdx.x] += sm[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0 && loss) *loss = sm[0] / (float)n;
}

__global__ void adamw_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m,
                             float* __restrict__ v, int64_t n, float lr, float b1, float b2, float eps, float wd,
                             float bc1, float bc2) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float gi = g[i];
  const float mi = b1 * m[i] + (1.f - b1) * gi;
  const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
  m[i] = mi;
  v[i] = vi;
  const float mh = mi / bc1, vh = vi / bc2;
  const float pi = p[i];
  p[i] = pi - lr * (mh / (sqrtf(vh) + eps) + wd * pi);
}


// ---- LayerNorm over [pooled | globals] rows (nequip.py:201-205; flax.linen.LayerNorm: eps 1e-6, scale + bias) ----
// one warp per row; x = concat(a[b, :da], g[b, :dg]); stats[b] = (mean, rstd)
__global__ void layernorm_fwd_kernel(const float* __restrict__ a, int da, const float* __restrict__ g, int dg, int B,
                                     const float* __restrict__ scale, const float* __restrict__ bias, float eps,
                                     float* __restrict__ y, float* __restrict__ stats) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const int D = da + dg;
  auto x = [&](int c) { return c < da ? a[(size_t)row * da + c] : g[(size_t)row * dg + (c - da)]; };
  float s = 0.f;
  for (int c = lane; c < D; c += 32) s += x(c);
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mu = s / (float)D;
  float v = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float d = x(c) - mu;
    v = fmaf(d, d, v);
  }
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const float rstd = rsqrtf(v / (float)D + eps);
  for (int c = lane; c < D; c += 32) y[(size_t)row * D + c] = (x(c) - mu) * rstd * scale[c] + bias[c];
  if (lane == 0) {
    stats[2 * row] = mu;
    stats[2 * row + 1] = rstd;
  }
}

// dx = rstd * (dyh - mean(dyh) - xh * mean(dyh * xh)), dyh = dy * scale; only the first da columns are returned
__global__ void layernorm_bwd_kernel(const float* __restrict__ a, int da, const float* __restrict__ g, int dg, int B,
                                     const float* __restrict__ scale, const float* __restrict__ stats,
                                     const float* __restrict__ dy, float* __restrict__ d_a) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= B) return;
  const int D = da + dg;
  const float mu = stats[2 * row], rstd = stats[2 * row + 1];
  auto xh = [&](int c) { return ((c < da ? a[(size_t)row * da + c] : g[(size_t)row * dg + (c - da)]) - mu) * rstd; };
  float s1 = 0.f, s2 = 0.f;
  for (int c = lane; c < D; c += 32) {
    const float t = dy[(size_t)row * D + c] * scale[c];
    s1 += t;
    s2 = fmaf(t, xh(c), s2);
  }
  for (int o = 16; o > 0; o >>= 1) {
    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
  }
  s1 /= (float)D;
  s2 /= (float)D;
  for (int c = lane; c < da; c += 32)
    d_a[(size_t)row * da + c] = rstd * (dy[(size_t)row * D + c] * scale[c] - s1 - xh(c) * s2);
}

// dscale[c] = sum_b dy[b,c] * xh[b,c];  dbias[c] = sum_b dy[b,c]   (rows in ascending order: deterministic)
__global__ void layernorm_wgrad_kernel(const float* __restrict__ a, int da, const float* __restrict__ g, int dg, int B,
                                       const float* __restrict__ stats, const float* __restrict__ dy,
                                       float* __restrict__ dscale, float* __restrict__ dbias) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x, D = da + dg;
  if (c >= D) return;
  float ds = 0.f, db = 0.f;
  for (int b = 0; b < B; ++b) {
    const float x = c < da ? a[(size_t)b * da + c] : g[(size_t)b * dg + (c - da)];
    const float t = dy[(size_t)b * D + c];
    ds = fmaf(t, (x - stats[2 * b]) * stats[2 * b + 1], ds);
    db += t;
  }
  dscale[c] = ds;
  dbias[c] = db;
}

}  // namespace

extern "C" int eqnn_gate_act_fwd(const float* z, int64_t n, int n_extra, int n_gates, int n_chunks, const int* mul,
                                 const int* l, int activation, float inv_c_act, float inv_c_gate, float* out,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(z && out && n >= 0, "gate_fwd: bad argument");
  EQNN_CHECK_ARG(activation == EQNN_ACT_GELU || activation == EQNN_ACT_SILU, "gate_fwd: unknown activation");
  GateDesc d;
  int rc = make_gate_desc(d, n_extra, n_gates, n_chunks, mul, l, inv_c_act, inv_c_gate);
  if (rc) return rc;
  d.act = activation;
  const int64_t tot = n * (d.n_extra + d.d_gated);
  if (tot == 0) return EQNN_OK;
  const int d_out = d.n_extra + d.d_gated;
  if (n >= 4096 && d_out <= 1024)
    gate_fwd_rows_kernel<<<(unsigned)ceil_div64(n, GATE_ROWS), (d_out + 31) / 32 * 32, 0, as_stream(stream)>>>(z, n, d, out);
  else
    gate_fwd_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, as_stream(stream)>>>(z, n, d, out);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_gate_act_bwd(const float* z, const float* gout, int64_t n, int n_extra, int n_gates, int n_chunks,
                                 const int* mul, const int* l, int activation, float inv_c_act, float inv_c_gate, float* gz,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(z && gout && gz && n >= 0, "gate_bwd: bad argument");
  EQNN_CHECK_ARG(activation == EQNN_ACT_GELU || activation == EQNN_ACT_SILU, "gate_bwd: unknown activation");
  GateDesc d;
  int rc = make_gate_desc(d, n_extra, n_gates, n_chunks, mul, l, inv_c_act, inv_c_gate);
  if (rc) return rc;
  d.act = activation;
  const int64_t tot = n * (d.n_extra + d.n_gates + d.d_gated);
  if (tot == 0) return EQNN_OK;
  const int d_in = d.n_extra + d.n_gates + d.d_gated;
  if (n >= 4096 && d_in <= 1024)
    gate_bwd_rows_kernel<<<(unsigned)ceil_div64(n, GATE_ROWS), (d_in + 31) / 32 * 32, 0, as_stream(stream)>>>(z, gout, n, d, gz);
  else
    gate_bwd_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, as_stream(stream)>>>(z, gout, n, d, gz);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_gate_fwd(const float* z, int64_t n, int n_extra, int n_gates, int n_chunks, const int* mul,
                             const int* l, float inv_c_act, float inv_c_gate, float* out, eqnn_stream_t stream) {
  return eqnn_gate_act_fwd(z, n, n_extra, n_gates, n_chunks, mul, l, EQNN_ACT_GELU, inv_c_act, inv_c_gate, out, stream);
}
extern "C" int eqnn_gate_bwd(const float* z, const float* gout, int64_t n, int n_extra, int n_gates, int n_chunks,
                             const int* mul, const int* l, float inv_c_act, float inv_c_gate, float* gz,
                             eqnn_stream_t stream) {
  return eqnn_gate_act_bwd(z, gout, n, n_extra, n_gates, n_chunks, mul, l, EQNN_ACT_GELU, inv_c_act, inv_c_gate, gz, stream);
}

extern "C" int eqnn_layernorm_fwd(const float* a, int da, const float* g, int dg, int B, const float* scale,
                                  const float* bias, float eps, float* y, float* stats, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(a && (g || dg == 0) && scale && bias && y && stats && B >= 1 && da >= 1 && dg >= 0,
                 "layernorm_fwd: bad argument");
  layernorm_fwd_kernel<<<(B + 7) / 8, 256, 0, as_stream(stream)>>>(a, da, g, dg, B, scale, bias, eps, y, stats);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_layernorm_bwd(const float* a, int da, const float* g, int dg, int B, const float* scale,
                                  const float* stats, const float* dy, float* d_a, float* dscale, float* dbias,
                                  eqnn_stream_t stream) {
  EQNN_CHECK_ARG(a && (g || dg == 0) && scale && stats && dy && d_a && dscale && dbias && B >= 1 && da >= 1 && dg >= 0,
                 "layernorm_bwd: bad argument");
  cudaStream_t st = as_stream(stream);
  layernorm_bwd_kernel<<<(B + 7) / 8, 256, 0, st>>>(a, da, g, dg, B, scale, stats, dy, d_a);
  layernorm_wgrad_kernel<<<(da + dg + 127) / 128, 128, 0, st>>>(a, da, g, dg, B, stats, dy, dscale, dbias);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}

extern "C" int eqnn_pool_fwd(const float* x, int B, int N, int D, int mode, float* out, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && out && B >= 1 && N >= 1 && D >= 1, "pool_fwd: bad argument");
  EQNN_CHECK_ARG(mode == 0 || mode == 1, "pool_fwd: mode must be 0 (sum) or 1 (mean)");
  dim3 grid((D + 31) / 32, B), block(32, 8);
  pool_fwd_kernel<<<grid, block, 0, as_stream(stream)>>>(x, B, N, D, mode == 1 ? 1.f / (float)N : 1.f, out);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_pool_bwd(const float* gout, int B, int N, int D, int mode, float* gx, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(gout && gx && B >= 1 && N >= 1 && D >= 1, "pool_bwd: bad argument");
  EQNN_CHECK_ARG(mode == 0 || mode == 1, "pool_bwd: mode must be 0 (sum) or 1 (mean)");
  const int64_t tot = (int64_t)B * N * D;
  pool_bwd_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, as_stream(stream)>>>(gout, B, N, D,
                                                                                mode == 1 ? 1.f / (float)N : 1.f, gx);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_pool_max_fwd(const float* x, int B, int N, int D, float* out, int32_t* argmax, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && out && argmax && B >= 1 && N >= 1 && D >= 1, "pool_max_fwd: bad argument");
  dim3 grid((D + 31) / 32, B), block(32, 8);
  pool_max_fwd_kernel<<<grid, block, 0, as_stream(stream)>>>(x, B, N, D, out, argmax);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_pool_max_bwd(const float* gout, const int32_t* argmax, int B, int N, int D, float* gx,
                                 eqnn_stream_t stream) {
  EQNN_CHECK_ARG(gout && argmax && gx && B >= 1 && N >= 1 && D >= 1, "pool_max_bwd: bad argument");
  const int64_t tot = (int64_t)B * N * D;
  pool_max_bwd_kernel<<<(unsigned)ceil_div64(tot, 256), 256, 0, as_stream(stream)>>>(gout, argmax, B, N, D, gx);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_colsum(const float* x, int64_t M, int N, float* y, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(x && y && M >= 0 && N >= 1, "colsum: bad argument");
  dim3 grid((N + 31) / 32), block(32, 8);
  colsum_kernel<<<grid, block, 0, as_stream(stream)>>>(x, M, N, y);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_weights_expand(const float* params, const int32_t* idx, const float* scale, int64_t n_dst,
                                   float* dst, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(params && idx && scale && dst && n_dst >= 0, "weights_expand: bad argument");
  if (n_dst == 0) return EQNN_OK;
  expand_kernel<<<(unsigned)ceil_div64(n_dst, 256), 256, 0, as_stream(stream)>>>(params, idx, scale, n_dst, dst);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_weights_contract(const float* ddst, const int32_t* pidx, const int32_t* ptr,
                                     const int32_t* tidx, const float* scale_t, int64_t n_params, float* grads,
                                     eqnn_stream_t stream) {
  EQNN_CHECK_ARG(ddst && pidx && ptr && tidx && scale_t && grads && n_params >= 0, "weights_contract: bad argument");
  if (n_params == 0) return EQNN_OK;
  contract_kernel<<<(unsigned)ceil_div64(n_params, 256), 256, 0, as_stream(stream)>>>(ddst, pidx, ptr, tidx, scale_t,
                                                                                     n_params, grads);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_mse_loss(const float* pred, const float* target, int B, int O, float gscale, float* loss,
                             float* gpred, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(pred && target && B >= 1 && O >= 1, "mse_loss: bad argument");
  mse_kernel<<<1, 256, 0, as_stream(stream)>>>(pred, target, B * O, gscale, loss, gpred);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_adamw(float* params, const float* grads, float* m, float* v, int64_t n, float lr, float b1,
                          float b2, float eps, float wd, int step, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(params && grads && m && v && n >= 0 && step >= 1, "adamw: bad argument");
  if (n == 0) return EQNN_OK;
  const float bc1 = 1.f - powf(b1, (float)step), bc2 = 1.f - powf(b2, (float)step);
  adamw_kernel<<<(unsigned)ceil_div64(n, 256), 256, 0, as_stream(stream)>>>(params, grads, m, v, n, lr, b1, b2, eps,
                                                                           wd, bc1, bc2);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

// ---- error plumbing -------------------------------------------------------------------
static thread_local char g_err[256] = "";
int eqnn_fail(int code, const char* msg) {
  snprintf(g_err, sizeof(g_err), "%s", msg ? msg : "unknown error");
  return code;
}
#include <atomic>
static std::atomic<long long> g_launches{0};
void eqnn_count_launches(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
extern "C" long long eqnn_launch_count(void) { return g_launches.load(std::memory_order_relaxed); }
extern "C" const char* eqnn_last_error_string(void) { return g_err; }
extern "C" int eqnn_version(void) { return 100; }
