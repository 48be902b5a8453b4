// Fused node update of the NequIP interaction block (sm_100a):
//     z  = alpha1 * A @ W1 + h @ W2          Linear on the aggregated messages + self-connection (nequip.py:80-85)
//     g  = gate(z)                           e3nn.gate, default activations                        (nequip.py:88-90)
//     h' = g @ W3                            re-projection to the input irreps                     (nequip.py:162)
//     xt = h' @ W0n                          next step's sender record (Linear of nequip.py:43)     (optional)
// in ONE N-row kernel: z is written once (saved for the backward pass), g never leaves the SM.  The irreps
// Linear layers arrive lowered to small dense matrices (<= 32 x 256), which live in shared memory; persistent
// CTAs sweep tiles of 16 nodes.  The backward kernel mirrors it and reduces the three weight gradients in
// per-thread registers -> per-CTA partials -> fixed-order sum (deterministic, no atomics).
#include "common.cuh"

namespace {

constexpr int NT = 256;        // threads per CTA (one z column per thread)
constexpr int TILE = 16;       // nodes per tile
constexpr int MAX_IN = 32;     // d_a + d_h
constexpr int MAX_DH = 8;
constexpr int MAX_CH = 8;

struct NodeP {
  const float* A; int ld_a, d_a;
  const float* h; int d_h;
  const float* W1; float alpha1;
  const float* W2;
  const float* W3;
  const float* W0n; int dxp;
  int n_extra, n_gates, n_chunks, d_gated, d_z, d_g;
  int mul[MAX_CH], dim[MAX_CH];
  float inv_c_act, inv_c_gate;
  int64_t n;
  // forward outputs
  float* z; float* h_out; float* xt_out;
  // backward
  const float* dh_out; float* dA; float* dh_in; float* partial;
};

__device__ __forceinline__ int gate_of(const NodeP& p, int j) {   // gated component j -> gate index
  int g = 0;
#pragma unroll
  for (int c = 0; c < MAX_CH; ++c) {
    if (c < p.n_chunks) {
      const int w = p.mul[c] * p.dim[c];
      if (j < w) return g + j / p.dim[c];
      j -= w;
      g += p.mul[c];
    }
  }
  return g;
}

// shared-memory carve-up (floats)
struct Smem {
  float* W12;   // [d_in][d_z]   rows 0..d_a-1 = alpha1*W1, rows d_a.. = W2
  float* W3;    // [d_g][d_h]
  float* W0n;   // [d_h][dxp]
  float* in;    // [TILE][d_in]
  float* z;     // [TILE][d_z]
  float* g;     // [TILE][d_g]   (forward: gate output; backward: dg then reused)
  float* t;     // [TILE][MAX_DH] small per-node vectors (h', dh_out)
};

// odd row stride for the matrices that are read "one row per lane" (conflict-free shared-memory banks)
__host__ __device__ __forceinline__ int odd(int n) { return n | 1; }

__device__ __forceinline__ Smem carve(float* base, const NodeP& p, int d_in) {
  Smem s;
  s.W12 = base;
  s.W3 = s.W12 + d_in * odd(p.d_z);
  s.W0n = s.W3 + p.d_g * p.d_h;
  s.in = s.W0n + p.d_h * (p.dxp > 0 ? p.dxp : 1);
  s.z = s.in + TILE * d_in;
  s.g = s.z + TILE * p.d_z;
  s.t = s.g + TILE * p.d_g;
  return s;
}

size_t smem_floats(const NodeP& p) {
  const int d_in = p.d_a + p.d_h;
  return (size_t)d_in * odd(p.d_z) + (size_t)p.d_g * p.d_h + (size_t)p.d_h * (p.dxp > 0 ? p.dxp : 1) +
         (size_t)TILE * (d_in + p.d_z + p.d_g + MAX_DH) + 2 * (size_t)TILE * odd(p.d_z);   // + dz, gact in backward
}

__device__ __forceinline__ void load_weights(const NodeP& p, const Smem& s, int d_in) {
  for (int i = threadIdx.x; i < d_in * p.d_z; i += NT) {
    const int r = i / p.d_z, c = i - r * p.d_z;
    float v;
    if (r < p.d_a) v = p.alpha1 * p.W1[(size_t)r * p.d_z + c];
    else v = p.W2 ? p.W2[(size_t)(r - p.d_a) * p.d_z + c] : 0.f;
    s.W12[r * odd(p.d_z) + c] = v;
  }
  for (int i = threadIdx.x; i < p.d_g * p.d_h; i += NT) s.W3[i] = p.W3[i];
  if (p.W0n)
    for (int i = threadIdx.x; i < p.d_h * p.dxp; i += NT) s.W0n[i] = p.W0n[i];
}

__device__ __forceinline__ void load_inputs(const NodeP& p, const Smem& s, int d_in, int64_t n0, int nn) {
  for (int i = threadIdx.x; i < TILE * d_in; i += NT) {
    const int node = i / d_in, r = i - node * d_in;
    float v = 0.f;
    if (node < nn) v = (r < p.d_a) ? p.A[(n0 + node) * p.ld_a + r] : p.h[(n0 + node) * p.d_h + (r - p.d_a)];
    s.in[i] = v;
  }
}

__global__ void __launch_bounds__(NT) node_fwd_kernel(const NodeP p) {
  extern __shared__ __align__(16) float smem_f[];
  const int d_in = p.d_a + p.d_h;
  const Smem s = carve(smem_f, p, d_in);
  load_weights(p, s, d_in);
  const int tid = threadIdx.x;
  const int64_t n_tiles = (p.n + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * TILE;
    const int nn = (int)min((int64_t)TILE, p.n - n0);
    __syncthreads();
    load_inputs(p, s, d_in, n0, nn);
    __syncthreads();
    // z[node][c]: thread = column, weights of the column in registers would need d_in <= 32 registers; stream from smem
    if (tid < p.d_z) {
      float acc[TILE];
#pragma unroll
      for (int q = 0; q < TILE; ++q) acc[q] = 0.f;
      for (int r = 0; r < d_in; ++r) {
        const float w = s.W12[r * odd(p.d_z) + tid];
#pragma unroll
        for (int q = 0; q < TILE; ++q) acc[q] = fmaf(s.in[q * d_in + r], w, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < TILE; ++q) s.z[q * p.d_z + tid] = acc[q];
    }
    __syncthreads();
    // save z (coalesced), gate
    for (int i = tid; i < nn * p.d_z; i += NT) p.z[n0 * p.d_z + i] = s.z[i];
    for (int i = tid; i < TILE * p.d_g; i += NT) {
      const int node = i / p.d_g, c = i - node * p.d_g;
      const float* zr = s.z + node * p.d_z;
      float v;
      if (c < p.n_extra) v = gelu_tanh(zr[c]) * p.inv_c_act;
      else {
        const int j = c - p.n_extra;
        v = sigmoid_acc(zr[p.n_extra + gate_of(p, j)]) * p.inv_c_gate * zr[p.n_extra + p.n_gates + j];
      }
      s.g[i] = v;
    }
    __syncthreads();
    // h'[node][o] = sum_j g[node][j] W3[j][o]
    if (tid < TILE * p.d_h) {
      const int node = tid / p.d_h, o = tid - node * p.d_h;
      const float* gr = s.g + node * p.d_g;
      float a = 0.f;
      for (int j = 0; j < p.d_g; ++j) a = fmaf(gr[j], s.W3[j * p.d_h + o], a);
      s.t[node * MAX_DH + o] = a;
      if (node < nn) p.h_out[(n0 + node) * p.d_h + o] = a;
    }
    __syncthreads();
    if (p.W0n && tid < TILE * p.dxp) {
      const int node = tid / p.dxp, c = tid - node * p.dxp;
      float a = 0.f;
      for (int r = 0; r < p.d_h; ++r) a = fmaf(s.t[node * MAX_DH + r], s.W0n[r * p.dxp + c], a);
      if (node < nn) p.xt_out[(n0 + node) * p.dxp + c] = a;
    }
  }
}

// Backward: given dh_out, produces dA, dh_in (self-connection part) and per-CTA partial weight gradients
// laid out as [dW12 (d_in x d_z) | dW3 (d_g x d_h)].
__global__ void __launch_bounds__(NT) node_bwd_kernel(const NodeP p) {
  extern __shared__ __align__(16) float smem_f[];
  const int d_in = p.d_a + p.d_h;
  const Smem s = carve(smem_f, p, d_in);
  const int ldz = odd(p.d_z);
  float* dz = s.t + TILE * MAX_DH;            // [TILE][ldz]
  float* gact = dz + TILE * ldz;              // [TILE][d_g]: dg, reused between phases
  load_weights(p, s, d_in);
  const int tid = threadIdx.x;
  // weight-gradient accumulators: thread c < d_z owns column c of dW12; thread j < d_g owns row j of dW3
  float aw[MAX_IN], a3[MAX_DH];
#pragma unroll
  for (int r = 0; r < MAX_IN; ++r) aw[r] = 0.f;
#pragma unroll
  for (int o = 0; o < MAX_DH; ++o) a3[o] = 0.f;
  const int64_t n_tiles = (p.n + TILE - 1) / TILE;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * TILE;
    const int nn = (int)min((int64_t)TILE, p.n - n0);
    __syncthreads();
    load_inputs(p, s, d_in, n0, nn);
    for (int i = tid; i < TILE * p.d_z; i += NT) s.z[i] = (i < nn * p.d_z) ? p.z[n0 * p.d_z + i] : 0.f;
    for (int i = tid; i < TILE * MAX_DH; i += NT) {
      const int node = i / MAX_DH, o = i - node * MAX_DH;
      s.t[i] = (node < nn && o < p.d_h) ? p.dh_out[(n0 + node) * p.d_h + o] : 0.f;
    }
    __syncthreads();
    // g (recomputed) and dg[node][j] = sum_o dh_out[node][o] W3[j][o]
    for (int i = tid; i < TILE * p.d_g; i += NT) {
      const int node = i / p.d_g, c = i - node * p.d_g;
      const float* zr = s.z + node * p.d_z;
      float dg = 0.f;
      for (int o = 0; o < p.d_h; ++o) dg = fmaf(s.t[node * MAX_DH + o], s.W3[c * p.d_h + o], dg);
      float gv;
      if (c < p.n_extra) {
        gv = gelu_tanh(zr[c]) * p.inv_c_act;
        dz[node * ldz + c] = dg * gelu_tanh_grad(zr[c]) * p.inv_c_act;
      } else {
        const int j = c - p.n_extra;
        const float sg = sigmoid_acc(zr[p.n_extra + gate_of(p, j)]) * p.inv_c_gate;
        gv = sg * zr[p.n_extra + p.n_gates + j];
        dz[node * ldz + p.n_extra + p.n_gates + j] = dg * sg;
      }
      s.g[i] = gv;        // gate output (for dW3)
      gact[i] = dg;       // dg (for the gate-scalar gradients)
    }
    __syncthreads();
    // gradients of the gate scalars: sum over the components they gate
    for (int i = tid; i < TILE * p.n_gates; i += NT) {
      const int node = i / p.n_gates, gi = i - node * p.n_gates;
      int rem = gi, off = 0, dim = 1;
#pragma unroll
      for (int k = 0; k < MAX_CH; ++k) {
        if (k < p.n_chunks) {
          if (rem >= 0 && rem < p.mul[k]) { off += rem * p.dim[k]; dim = p.dim[k]; rem = -1; }
          else if (rem >= 0) { rem -= p.mul[k]; off += p.mul[k] * p.dim[k]; }
        }
      }
      const float* zr = s.z + node * p.d_z;
      float acc = 0.f;
      for (int q = 0; q < dim; ++q) acc += gact[node * p.d_g + p.n_extra + off + q] * zr[p.n_extra + p.n_gates + off + q];
      const float sg = sigmoid_acc(zr[p.n_extra + gi]);
      dz[node * ldz + p.n_extra + gi] = acc * sg * (1.f - sg) * p.inv_c_gate;
    }
    __syncthreads();
    // weight gradients (registers) ...
    if (tid < p.d_z) {
#pragma unroll
      for (int r = 0; r < MAX_IN; ++r) {
        if (r < d_in) {
          float a = aw[r];
#pragma unroll
          for (int q = 0; q < TILE; ++q) a = fmaf(s.in[q * d_in + r], dz[q * ldz + tid], a);
          aw[r] = a;
        }
      }
    }
    if (tid < p.d_g) {
#pragma unroll
      for (int o = 0; o < MAX_DH; ++o) {
        if (o < p.d_h) {
          float a = a3[o];
#pragma unroll
          for (int q = 0; q < TILE; ++q) a = fmaf(s.g[q * p.d_g + tid], s.t[q * MAX_DH + o], a);
          a3[o] = a;
        }
      }
    }
    // ... and data gradients d_in[node][r] = sum_c dz[node][c] W12[r][c]
    for (int i = tid; i < TILE * d_in; i += NT) {
      const int node = i / d_in, r = i - node * d_in;
      const float* dzr = dz + node * ldz;
      const float* wr = s.W12 + r * odd(p.d_z);
      float a = 0.f;
      for (int c = 0; c < p.d_z; ++c) a = fmaf(dzr[c], wr[c], a);
      if (node < nn) {
        if (r < p.d_a) p.dA[(n0 + node) * p.d_a + r] = a;
        else p.dh_in[(n0 + node) * p.d_h + (r - p.d_a)] = p.W2 ? a : 0.f;
      }
    }
  }
  // per-CTA partials
  float* part = p.partial + (size_t)blockIdx.x * ((size_t)d_in * p.d_z + (size_t)p.d_g * p.d_h);
  if (tid < p.d_z) {
#pragma unroll
    for (int r = 0; r < MAX_IN; ++r)
      if (r < d_in) part[(size_t)r * p.d_z + tid] = aw[r];
  }
  if (tid < p.d_g) {
#pragma unroll
    for (int o = 0; o < MAX_DH; ++o)
      if (o < p.d_h) part[(size_t)d_in * p.d_z + (size_t)tid * p.d_h + o] = a3[o];
  }
}

// dW1 = alpha1 * sum partial rows < d_a ; dW2 = sum partial rows >= d_a ; dW3 = sum partial tail
__global__ void node_wgrad_reduce_kernel(const float* __restrict__ partial, int n_parts, int d_a, int d_h, int d_z,
                                         int d_g, float alpha1, float* __restrict__ dW1, float* __restrict__ dW2,
                                         float* __restrict__ dW3) {
  const int d_in = d_a + d_h;
  const int tot = d_in * d_z + d_g * d_h;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= tot) return;
  float sum = 0.f;
  for (int q = 0; q < n_parts; ++q) sum += partial[(size_t)q * tot + i];
  if (i < d_a * d_z) dW1[i] = alpha1 * sum;     // forward used alpha1*W1: d/dW1 = alpha1 * d/d(alpha1 W1)
  else if (i < d_in * d_z) { if (dW2) dW2[i - d_a * d_z] = sum; }
  else dW3[i - d_in * d_z] = sum;
}

int fill_params(NodeP& p, const float* A, int ld_a, int d_a, const float* h, int d_h, const float* W1, float alpha1,
                const float* W2, int n_extra, int n_gates, int n_chunks, const int* mul, const int* l,
                float inv_c_act, float inv_c_gate, const float* W3, int64_t n) {
  EQNN_CHECK_ARG(A && h && W1 && W3, "node_update: null pointer");
  EQNN_CHECK_ARG(n_chunks >= 0 && n_chunks <= MAX_CH, "node_update: too many gated chunks");
  EQNN_CHECK_ARG(d_h >= 1 && d_h <= MAX_DH && d_a >= 1 && d_a + d_h <= MAX_IN, "node_update: input width not supported");
  p.A = A; p.ld_a = ld_a; p.d_a = d_a; p.h = h; p.d_h = d_h; p.W1 = W1; p.alpha1 = alpha1; p.W2 = W2; p.W3 = W3;
  p.n_extra = n_extra; p.n_gates = n_gates; p.n_chunks = n_chunks; p.d_gated = 0;
  p.inv_c_act = inv_c_act; p.inv_c_gate = inv_c_gate; p.n = n;
  int tot = 0;
  for (int c = 0; c < MAX_CH; ++c) {
    p.mul[c] = c < n_chunks ? mul[c] : 0;
    p.dim[c] = c < n_chunks ? 2 * l[c] + 1 : 1;
    p.d_gated += p.mul[c] * p.dim[c];
    tot += p.mul[c];
  }
  EQNN_CHECK_ARG(tot == n_gates, "node_update: number of gates != number of gated irreps");
  p.d_z = n_extra + n_gates + p.d_gated;
  p.d_g = n_extra + p.d_gated;
  EQNN_CHECK_ARG(p.d_z <= NT && TILE * d_h <= NT, "node_update: hidden width not supported");
  p.W0n = nullptr; p.dxp = 0; p.z = nullptr; p.h_out = nullptr; p.xt_out = nullptr;
  p.dh_out = nullptr; p.dA = nullptr; p.dh_in = nullptr; p.partial = nullptr;
  return EQNN_OK;
}

int grid_for(int64_t n) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int64_t tiles = (n + TILE - 1) / TILE;
  const int64_t g = 2 * (int64_t)sms;
  return (int)(tiles < g ? tiles : g);
}

}  // namespace

extern "C" size_t eqnn_node_update_workspace_bytes(int d_a, int d_h, int d_z, int d_g) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (size_t)2 * sms * ((size_t)(d_a + d_h) * d_z + (size_t)d_g * d_h) * sizeof(float);
}

extern "C" int eqnn_node_update_fwd(const float* A, int ld_a, int d_a, const float* h, int d_h, const float* W1,
                                    float alpha1, const float* W2, int n_extra, int n_gates, int n_chunks,
                                    const int* mul, const int* l, float inv_c_act, float inv_c_gate, const float* W3,
                                    const float* W0n, int dxp, int64_t n, float* z, float* h_out, float* xt_out,
                                    eqnn_stream_t stream) {
  NodeP p;
  int rc = fill_params(p, A, ld_a, d_a, h, d_h, W1, alpha1, W2, n_extra, n_gates, n_chunks, mul, l, inv_c_act,
                       inv_c_gate, W3, n);
  if (rc) return rc;
  EQNN_CHECK_ARG(z && h_out, "node_update_fwd: null output");
  EQNN_CHECK_ARG(!W0n || (xt_out && dxp >= 1 && TILE * dxp <= NT), "node_update_fwd: bad next-step record");
  if (n <= 0) return EQNN_OK;
  p.W0n = W0n; p.dxp = W0n ? dxp : 0; p.z = z; p.h_out = h_out; p.xt_out = xt_out;
  const size_t smem = smem_floats(p) * sizeof(float);
  EQNN_CHECK_ARG(smem <= 200 * 1024, "node_update_fwd: shared memory budget exceeded");
  EQNN_CUDA(cudaFuncSetAttribute(node_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  node_fwd_kernel<<<grid_for(n), NT, smem, as_stream(stream)>>>(p);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_node_update_bwd(const float* A, int ld_a, int d_a, const float* h, int d_h, const float* W1,
                                    float alpha1, const float* W2, int n_extra, int n_gates, int n_chunks,
                                    const int* mul, const int* l, float inv_c_act, float inv_c_gate, const float* W3,
                                    const float* z, const float* dh_out, int64_t n, float* dA, float* dh_in,
                                    float* dW1, float* dW2, float* dW3, void* ws, size_t ws_bytes,
                                    eqnn_stream_t stream) {
  NodeP p;
  int rc = fill_params(p, A, ld_a, d_a, h, d_h, W1, alpha1, W2, n_extra, n_gates, n_chunks, mul, l, inv_c_act,
                       inv_c_gate, W3, n);
  if (rc) return rc;
  EQNN_CHECK_ARG(z && dh_out && dA && dh_in && dW1 && dW3 && (dW2 || !W2), "node_update_bwd: null pointer");
  if (n <= 0) return EQNN_OK;
  p.z = const_cast<float*>(z); p.dh_out = dh_out; p.dA = dA; p.dh_in = dh_in;
  const int grid = grid_for(n);
  const size_t per = (size_t)(d_a + d_h) * p.d_z + (size_t)p.d_g * d_h;
  EQNN_CHECK_ARG(ws && ws_bytes >= (size_t)grid * per * sizeof(float), "node_update_bwd: workspace too small");
  p.partial = reinterpret_cast<float*>(ws);
  const size_t smem = smem_floats(p) * sizeof(float);
  EQNN_CHECK_ARG(smem <= 200 * 1024, "node_update_bwd: shared memory budget exceeded");
  EQNN_CUDA(cudaFuncSetAttribute(node_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = as_stream(stream);
  node_bwd_kernel<<<grid, NT, smem, st>>>(p);
  node_wgrad_reduce_kernel<<<(unsigned)((per + 255) / 256), 256, 0, st>>>(p.partial, grid, d_a, d_h, p.d_z, p.d_g,
                                                                        alpha1, dW1, dW2, dW3);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}
