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
#include <stdlib.h>
#include <type_traits>

namespace {

constexpr int NT = 256;        // threads per CTA (one z column per thread)
constexpr int TILE = 16;       // nodes per tile
constexpr int CTAS_PER_SM = 3; // resident CTAs (registers <= 85, shared memory ~58 KB each)
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

// shared-memory carve-up (floats).  Layout rules that make the inner loops cheap:
//   * matrices that are walked along a row with 128-bit loads by lanes holding DIFFERENT rows (W12, dz) use a row
//     stride of 4*odd floats: eight consecutive rows then start in eight different 16-byte bank groups;
//   * per-tile operands that every lane reads at the same address (inputs, dh_out) are stored TRANSPOSED,
//     [feature][node], so that the 16 nodes of a tile are four broadcast 128-bit loads.
struct Smem {
  float* W12;   // [d_in][ldw]    rows 0..d_a-1 = alpha1*W1, rows d_a.. = W2, pad columns zero
  float* W3;    // [2][d_g][4]    two planes (outputs 0..3 / 4..7, zero padded): lanes that read consecutive rows read
                //                consecutive 16-byte groups -- no bank conflicts (one [d_g][8] array: 2-way)
  float* W0n;   // [d_h][dxp]
  float* inT;   // [d_in][TILE]
  float* z;     // [TILE][d_z]
  float* z2;    // [TILE][d_z]    backward only: the other half of the z double buffer (cp.async prefetch of the next tile)
  float* g;     // [TILE][d_g]    gate output
  float* tT;    // [MAX_DH][TILE] h' (forward) / dh_out (backward)
  float* dz;    // [TILE][ldw]    backward only, pad columns zero
  float* dg;    // [TILE][d_g]    backward only
  int* gidx;    // [d_gated]      gated component -> gate index
  int* goff;    // [n_gates]      gate -> first gated component
  int* gdim;    // [n_gates]      gate -> number of components it gates
  float* sgate; // [NT / 32][n_gates]  per-warp sigmoid of the node's gate scalars (evaluated once per gate, not once per
                //                     gated component)
};

__host__ __device__ __forceinline__ int row4odd(int n) {   // smallest multiple of 4 >= n whose quarter is odd
  int r = (n + 3) / 4;
  if ((r & 1) == 0) ++r;
  return 4 * r;
}

__host__ __device__ __forceinline__ size_t smem_layout(const NodeP& p, Smem* s, float* base) {
  const int d_in = p.d_a + p.d_h, ldw = row4odd(p.d_z);
  size_t o = 0;
  auto take = [&](size_t n) { size_t at = o; o += (n + 3) / 4 * 4; return at; };   // keep every array 16-byte aligned
  const size_t oW12 = take((size_t)d_in * ldw), oW3 = take((size_t)p.d_g * MAX_DH);
  const size_t oW0 = take((size_t)p.d_h * (p.dxp > 0 ? p.dxp : 1)), oIn = take((size_t)d_in * TILE);
  const size_t oZ = take((size_t)TILE * p.d_z), oZ2 = take((size_t)TILE * p.d_z), oG = take((size_t)TILE * p.d_g), oT = take((size_t)MAX_DH * TILE);
  const size_t oDz = take((size_t)TILE * ldw), oDg = take((size_t)TILE * p.d_g > 2 * TILE * 32 ? (size_t)TILE * p.d_g : (size_t)2 * TILE * 32);   // also the data-gradient scratch [2][TILE][32]
  const size_t oGi = take(p.d_gated > 0 ? p.d_gated : 1), oGo = take(p.n_gates > 0 ? p.n_gates : 1);
  const size_t oGd = take(p.n_gates > 0 ? p.n_gates : 1);
  const size_t oSg = take((size_t)(NT / 32) * (p.n_gates > 0 ? p.n_gates : 1));
  if (s) {
    s->sgate = base + oSg;
    s->z2 = base + oZ2;
    s->W12 = base + oW12; s->W3 = base + oW3; s->W0n = base + oW0; s->inT = base + oIn; s->z = base + oZ;
    s->g = base + oG; s->tT = base + oT; s->dz = base + oDz; s->dg = base + oDg;
    s->gidx = reinterpret_cast<int*>(base + oGi); s->goff = reinterpret_cast<int*>(base + oGo);
    s->gdim = reinterpret_cast<int*>(base + oGd);
  }
  return o;
}

size_t smem_floats(const NodeP& p) { return smem_layout(p, nullptr, nullptr); }

__device__ __forceinline__ void load_weights(const NodeP& p, const Smem& s, int d_in) {
  const int ldw = row4odd(p.d_z);
  for (int i = threadIdx.x; i < d_in * ldw; i += NT) {
    const int r = i / ldw, c = i - r * ldw;
    float v = 0.f;
    if (c < p.d_z) {
      if (r < p.d_a) v = p.alpha1 * p.W1[(size_t)r * p.d_z + c];
      else v = p.W2 ? p.W2[(size_t)(r - p.d_a) * p.d_z + c] : 0.f;
    }
    s.W12[i] = v;
  }
  for (int i = threadIdx.x; i < TILE * ldw; i += NT) s.dz[i] = 0.f;
  for (int i = threadIdx.x; i < p.d_g * MAX_DH; i += NT) {
    const int j = i / MAX_DH, o = i - j * MAX_DH;
    s.W3[(o >> 2) * (p.d_g * 4) + j * 4 + (o & 3)] = o < p.d_h ? p.W3[j * p.d_h + o] : 0.f;
  }
  if (p.W0n)
    for (int i = threadIdx.x; i < p.d_h * p.dxp; i += NT) s.W0n[i] = p.W0n[i];
  for (int j = threadIdx.x; j < p.d_gated; j += NT) s.gidx[j] = gate_of(p, j);
  for (int gi = threadIdx.x; gi < p.n_gates; gi += NT) {
    int rem = gi, off = 0, dim = 1;
#pragma unroll
    for (int k = 0; k < MAX_CH; ++k) {
      if (k < p.n_chunks) {
        if (rem >= 0 && rem < p.mul[k]) { off += rem * p.dim[k]; dim = p.dim[k]; rem = -1; }
        else if (rem >= 0) { rem -= p.mul[k]; off += p.mul[k] * p.dim[k]; }
      }
    }
    s.goff[gi] = off;
    s.gdim[gi] = dim;
  }
}

// Inputs of a tile, software-pipelined: the global loads of tile t + 1 are issued at the top of tile t into registers
// (TILE * d_in <= 512 elements = two per thread) and stored to shared memory at the top of tile t + 1 -- the per-tile
// chain load -> barrier -> GEMM was 15 % of the forward kernel's samples on the stores that waited for these loads.
__device__ __forceinline__ void fetch_inputs(const NodeP& p, int d_in, int64_t n0, int nn, float (&pre)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x + k * NT;
    float v = 0.f;
    if (i < TILE * d_in) {
      const int node = i / d_in, r = i - node * d_in;
      if (node < nn) v = (r < p.d_a) ? __ldg(p.A + (n0 + node) * p.ld_a + r) : __ldg(p.h + (n0 + node) * p.d_h + (r - p.d_a));
    }
    pre[k] = v;
  }
}
__device__ __forceinline__ void store_inputs(const Smem& s, int d_in, const float (&pre)[2]) {
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = threadIdx.x + k * NT;
    if (i < TILE * d_in) {
      const int node = i / d_in, r = i - node * d_in;
      s.inT[r * TILE + node] = pre[k];
    }
  }
}
__device__ __forceinline__ void cp_async16_zfill(float* dst_smem, const float* src, int src_bytes) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(dst_smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// the TILE node values of feature row `r` of a transposed operand: broadcast 128-bit loads
__device__ __forceinline__ void load_tile_row(const float* rowT, float (&v)[TILE]) {
#pragma unroll
  for (int q = 0; q < TILE; q += 4) {
    const float4 t = *reinterpret_cast<const float4*>(rowT + q);
    v[q] = t.x; v[q + 1] = t.y; v[q + 2] = t.z; v[q + 3] = t.w;
  }
}

__device__ __forceinline__ float dot4(const float4& a, const float4& b, float acc) {
  acc = fmaf(a.x, b.x, acc);
  acc = fmaf(a.y, b.y, acc);
  acc = fmaf(a.z, b.z, acc);
  return fmaf(a.w, b.w, acc);
}

// Shape specialisation.  The kernels are written for run-time shapes (any lowered irreps); with every width a run-time
// value the index arithmetic dominates them (2 400 SASS instructions, 134 of them FFMA: integer divisions, address
// products, loop control).  For the shapes the benchmarked models produce, the same code is instantiated with the widths
// as compile-time constants: the kernel copies its parameter block and overwrites the shape fields with the constants,
// from where the compiler folds strides and divisions and unrolls the loops.  SH = void keeps the run-time shape.
template <int DA, int DH, int NEXTRA, int NGATES, int MUL0, int DIM0>
struct NodeShape {
  static constexpr int d_a = DA, d_h = DH, n_extra = NEXTRA, n_gates = NGATES, mul0 = MUL0, dim0 = DIM0;
};
template <class SH>
__device__ __forceinline__ NodeP specialise(const NodeP& in) {
  NodeP p = in;
  if constexpr (!std::is_void<SH>::value) {
    p.d_a = SH::d_a; p.d_h = SH::d_h; p.n_extra = SH::n_extra; p.n_gates = SH::n_gates;
    p.n_chunks = 1; p.mul[0] = SH::mul0; p.dim[0] = SH::dim0;
#pragma unroll
    for (int c = 1; c < MAX_CH; ++c) { p.mul[c] = 0; p.dim[c] = 1; }
    p.d_gated = SH::mul0 * SH::dim0;
    p.d_z = SH::n_extra + SH::n_gates + SH::mul0 * SH::dim0;
    p.d_g = SH::n_extra + SH::mul0 * SH::dim0;
  }
  return p;
}
template <class SH>
bool shape_matches(const NodeP& p) {
  if constexpr (std::is_void<SH>::value) return true;
  else return p.d_a == SH::d_a && p.d_h == SH::d_h && p.n_extra == SH::n_extra && p.n_gates == SH::n_gates &&
              p.n_chunks == 1 && p.mul[0] == SH::mul0 && p.dim[0] == SH::dim0;
}
// the interaction blocks of the benchmarked models (bench.py MODEL_KW; messages 0e + 1o = 18 live components, node
// features 1o + 1o + 0e = 7): l_max 2 -> 88 scalars + 14 gates + 14 x 1o; l_max 3 -> 100 scalars + 10 gates + 10 x 1o
using ShapeCfg2 = NodeShape<18, 7, 88, 14, 14, 3>;
using ShapeCfg5 = NodeShape<18, 7, 100, 10, 10, 3>;

template <class SH>
__global__ void __launch_bounds__(NT, 3) node_fwd_kernel(const NodeP p_in) {
  const NodeP p = specialise<SH>(p_in);
  extern __shared__ __align__(16) float smem_f[];
  const int d_in = p.d_a + p.d_h, ldw = row4odd(p.d_z);
  Smem s;
  smem_layout(p, &s, smem_f);
  load_weights(p, s, d_in);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t n_tiles = (p.n + TILE - 1) / TILE;
  float pre[2];
  if ((int64_t)blockIdx.x < n_tiles)
    fetch_inputs(p, d_in, (int64_t)blockIdx.x * TILE, (int)min((int64_t)TILE, p.n - (int64_t)blockIdx.x * TILE), pre);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * TILE;
    const int nn = (int)min((int64_t)TILE, p.n - n0);
    __syncthreads();
    store_inputs(s, d_in, pre);
    if (tile + gridDim.x < n_tiles) {                          // next tile's inputs: in flight across this tile
      const int64_t m0 = (tile + gridDim.x) * TILE;
      fetch_inputs(p, d_in, m0, (int)min((int64_t)TILE, p.n - m0), pre);
    }
    __syncthreads();
    // z[node][c]: thread = column c, 16 nodes in registers; per input feature 1 weight + 4 broadcast loads, 16 FMAs
    if (tid < p.d_z) {
      float acc[TILE];
#pragma unroll
      for (int q = 0; q < TILE; ++q) acc[q] = 0.f;
      for (int r = 0; r < d_in; ++r) {
        const float w = s.W12[r * ldw + tid];
        float in[TILE];
        load_tile_row(s.inT + r * TILE, in);
#pragma unroll
        for (int q = 0; q < TILE; ++q) acc[q] = fmaf(in[q], w, acc[q]);
      }
#pragma unroll
      for (int q = 0; q < TILE; ++q) s.z[q * p.d_z + tid] = acc[q];
    }
    __syncthreads();
    // save z (coalesced), gate
    for (int i = tid; i < nn * p.d_z; i += NT) p.z[n0 * p.d_z + i] = s.z[i];
    for (int node = warp; node < TILE; node += NT / 32) {       // warp per node, lanes over components: no divisions
      const float* zr = s.z + node * p.d_z;
      float* gr = s.g + node * p.d_g;
      float* sgw = s.sgate + warp * p.n_gates;
      for (int gi = lane; gi < p.n_gates; gi += 32) sgw[gi] = sigmoid_acc(zr[p.n_extra + gi]) * p.inv_c_gate;
      for (int c = lane; c < p.n_extra; c += 32) gr[c] = gelu_tanh(zr[c]) * p.inv_c_act;
      __syncwarp();
      for (int j = lane; j < p.d_gated; j += 32) gr[p.n_extra + j] = sgw[s.gidx[j]] * zr[p.n_extra + p.n_gates + j];
      __syncwarp();                                              // sgw is rewritten for the warp's next node
    }
    __syncthreads();
    // h'[node][o] = sum_j g[node][j] W3[j][o]: 16 lanes per node split j, every lane keeps all d_h outputs (one g value
    // and one padded W3 row -- two 128-bit loads -- per 7 FMAs), then a 4-step shuffle tree over the 16 lanes
    {
      constexpr int PARTS = NT / TILE;                           // 16
      const int node = tid / PARTS, part = tid % PARTS;
      const float* gr = s.g + node * p.d_g;
      float a[MAX_DH];
#pragma unroll
      for (int o = 0; o < MAX_DH; ++o) a[o] = 0.f;
      for (int j = part; j < p.d_g; j += PARTS) {
        const float gv = gr[j];
        const float4 w0 = *reinterpret_cast<const float4*>(s.W3 + j * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(s.W3 + p.d_g * 4 + j * 4);
        a[0] = fmaf(gv, w0.x, a[0]); a[1] = fmaf(gv, w0.y, a[1]); a[2] = fmaf(gv, w0.z, a[2]); a[3] = fmaf(gv, w0.w, a[3]);
        a[4] = fmaf(gv, w1.x, a[4]); a[5] = fmaf(gv, w1.y, a[5]); a[6] = fmaf(gv, w1.z, a[6]); a[7] = fmaf(gv, w1.w, a[7]);
      }
#pragma unroll
      for (int off = PARTS / 2; off > 0; off >>= 1) {
#pragma unroll
        for (int o = 0; o < MAX_DH; ++o) a[o] += __shfl_xor_sync(0xffffffffu, a[o], off);
      }
      if (part < p.d_h) {
        float v = a[0];
#pragma unroll
        for (int o = 1; o < MAX_DH; ++o) v = (part == o) ? a[o] : v;
        s.tT[part * TILE + node] = v;
        if (node < nn) p.h_out[(n0 + node) * p.d_h + part] = v;
      }
    }
    __syncthreads();
    if (p.W0n && tid < TILE * p.dxp) {
      const int node = tid / p.dxp, c = tid - node * p.dxp;
      float a = 0.f;
      for (int r = 0; r < p.d_h; ++r) a = fmaf(s.tT[r * TILE + node], s.W0n[r * p.dxp + c], a);
      if (node < nn) p.xt_out[(n0 + node) * p.dxp + c] = a;
    }
  }
}

// Backward: given dh_out, produces dA, dh_in (self-connection part) and per-CTA partial weight gradients
// laid out as [dW12 (d_in x d_z) | dW3 (d_g x d_h)].
template <class SH>
__global__ void __launch_bounds__(NT, 3) node_bwd_kernel(const NodeP p_in) {
  const NodeP p = specialise<SH>(p_in);
  extern __shared__ __align__(16) float smem_f[];
  const int d_in = p.d_a + p.d_h, ldw = row4odd(p.d_z);
  Smem s;
  smem_layout(p, &s, smem_f);
  load_weights(p, s, d_in);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  // weight-gradient accumulators: thread c < d_z owns column c of dW12; thread j < d_g owns row j of dW3
  float aw[MAX_IN], a3[MAX_DH];
#pragma unroll
  for (int r = 0; r < MAX_IN; ++r) aw[r] = 0.f;
#pragma unroll
  for (int o = 0; o < MAX_DH; ++o) a3[o] = 0.f;
  const int64_t n_tiles = (p.n + TILE - 1) / TILE;
  // Software pipeline over tiles: inputs and dh_out of tile t + 1 travel in registers, its z block (TILE * d_z floats,
  // contiguous and 64-byte aligned in global memory) lands in the other half of a shared-memory double buffer by cp.async
  // (zero fill past the last node) while tile t is processed.
  float pre[2], pre_t = 0.f;
  float* zbuf[2] = {s.z, s.z2};
  int cur = 0;
  auto fetch_tile = [&](int64_t t, float* zdst) {
    const int64_t m0 = t * TILE;
    const int mn = (int)min((int64_t)TILE, p.n - m0);
    fetch_inputs(p, d_in, m0, mn, pre);
    if (tid < TILE * MAX_DH) {
      const int node = tid / MAX_DH, o = tid - node * MAX_DH;
      pre_t = (node < mn && o < p.d_h) ? __ldg(p.dh_out + (m0 + node) * p.d_h + o) : 0.f;
    }
    const int valid = mn * p.d_z;                              // floats of the block that exist
    for (int f = tid; f < TILE * p.d_z / 4; f += NT) {
      const int vf = min(max(valid - 4 * f, 0), 4);
      cp_async16_zfill(zdst + 4 * f, p.z + m0 * p.d_z + (vf > 0 ? 4 * f : 0), 4 * vf);
    }
  };
  if ((int64_t)blockIdx.x < n_tiles) fetch_tile(blockIdx.x, zbuf[0]);
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t n0 = tile * TILE;
    const int nn = (int)min((int64_t)TILE, p.n - n0);
    cp_async_wait_all();
    __syncthreads();                                           // previous tile's readers are done; this tile's z has landed
    store_inputs(s, d_in, pre);
    if (tid < TILE * MAX_DH) s.tT[(tid % MAX_DH) * TILE + tid / MAX_DH] = pre_t;
    s.z = zbuf[cur];
    if (tile + gridDim.x < n_tiles) fetch_tile(tile + gridDim.x, zbuf[cur ^ 1]);
    cur ^= 1;
    __syncthreads();
    // g (recomputed), dg[node][j] = sum_o dh_out[node][o] W3[j][o], dz -- one warp per node, no block barrier inside:
    // gate sigmoids once per gate (per-warp scratch), scalars with one exponential for gelu and gelu', gated components,
    // then the gate scalars' gradients from the products dg * z the same warp has just stored
    for (int node = warp; node < TILE; node += NT / 32) {       // warp per node, lanes over components
      const float* zr = s.z + node * p.d_z;
      float* sgw = s.sgate + warp * p.n_gates;
      float dho[MAX_DH];
#pragma unroll
      for (int o = 0; o < MAX_DH; ++o) dho[o] = s.tT[o * TILE + node];
      for (int gi = lane; gi < p.n_gates; gi += 32) sgw[gi] = sigmoid_acc(zr[p.n_extra + gi]);
      for (int c = lane; c < p.n_extra; c += 32) {
        const float4 w0 = *reinterpret_cast<const float4*>(s.W3 + c * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(s.W3 + p.d_g * 4 + c * 4);
        float dg = dho[0] * w0.x;                                  // (padded W3 columns and dh_out rows are zero)
        dg = fmaf(dho[1], w0.y, dg); dg = fmaf(dho[2], w0.z, dg); dg = fmaf(dho[3], w0.w, dg);
        dg = fmaf(dho[4], w1.x, dg); dg = fmaf(dho[5], w1.y, dg); dg = fmaf(dho[6], w1.z, dg); dg = fmaf(dho[7], w1.w, dg);
        float gv, gp;
        gelu_tanh_both(zr[c], gv, gp);
        s.g[node * p.d_g + c] = gv * p.inv_c_act;                // gate output (for dW3)
        s.dz[node * ldw + c] = dg * gp * p.inv_c_act;
      }
      __syncwarp();
      for (int j = lane; j < p.d_gated; j += 32) {
        const int c = p.n_extra + j;
        const float4 w0 = *reinterpret_cast<const float4*>(s.W3 + c * 4);
        const float4 w1 = *reinterpret_cast<const float4*>(s.W3 + p.d_g * 4 + c * 4);
        float dg = dho[0] * w0.x;                                  // (padded W3 columns and dh_out rows are zero)
        dg = fmaf(dho[1], w0.y, dg); dg = fmaf(dho[2], w0.z, dg); dg = fmaf(dho[3], w0.w, dg);
        dg = fmaf(dho[4], w1.x, dg); dg = fmaf(dho[5], w1.y, dg); dg = fmaf(dho[6], w1.z, dg); dg = fmaf(dho[7], w1.w, dg);
        const float sg = sgw[s.gidx[j]] * p.inv_c_gate;
        const float zv = zr[p.n_extra + p.n_gates + j];
        s.g[node * p.d_g + c] = sg * zv;
        s.dz[node * ldw + p.n_extra + p.n_gates + j] = dg * sg;
        s.dg[node * p.d_g + c] = dg * zv;                        // summed per gate below
      }
      __syncwarp();
      // gradients of the gate scalars: sum over the components they gate
      for (int gi = lane; gi < p.n_gates; gi += 32) {
        const int off = s.goff[gi], dim = s.gdim[gi];
        float acc = 0.f;
        for (int q = 0; q < dim; ++q) acc += s.dg[node * p.d_g + p.n_extra + off + q];
        const float sg = sgw[gi];
        s.dz[node * ldw + p.n_extra + gi] = acc * sg * (1.f - sg) * p.inv_c_gate;
      }
      __syncwarp();                                              // sgw is rewritten for the warp's next node
    }
    __syncthreads();
    // weight gradients (registers): the thread's column of dz for the 16 nodes, then per input feature
    // 4 broadcast loads + 16 FMAs
    if (tid < p.d_z) {
      float dzc[TILE];
#pragma unroll
      for (int q = 0; q < TILE; ++q) dzc[q] = s.dz[q * ldw + tid];
#pragma unroll
      for (int r = 0; r < MAX_IN; ++r) {
        if (r < d_in) {
          float in[TILE];
          load_tile_row(s.inT + r * TILE, in);
          float a = aw[r];
#pragma unroll
          for (int q = 0; q < TILE; ++q) a = fmaf(in[q], dzc[q], a);
          aw[r] = a;
        }
      }
    }
    if (tid < p.d_g) {
      float gq[TILE];
#pragma unroll
      for (int q = 0; q < TILE; ++q) gq[q] = s.g[q * p.d_g + tid];
#pragma unroll
      for (int o = 0; o < MAX_DH; ++o) {
        if (o < p.d_h) {
          float t[TILE];
          load_tile_row(s.tT + o * TILE, t);
          float a = a3[o];
#pragma unroll
          for (int q = 0; q < TILE; ++q) a = fmaf(gq[q], t[q], a);
          a3[o] = a;
        }
      }
    }
    // ... and data gradients d_in[node][r] = sum_c dz[node][c] W12[r][c]  (128-bit loads along c; pads are zero).
    // The kernel is bound by shared-memory wavefronts (73 % of the L1 data pipe's peak, profiles/r2h_tpconv_node_raw.csv), and
    // this product was 45 % of them when every node re-read all of W12: a warp now owns FOUR nodes and one half of the
    // columns -- lane r reads its W12 row once per four nodes, the dz rows are warp-wide broadcasts -- and the two column
    // halves meet in a scratch (the dg array, dead after the gate phase) from which the results leave coalesced.
    {
      static_assert(NT / 32 == 8 && TILE == 16, "node-group x column-half mapping of the data-gradient product");
      const int ng = warp & 3, ch = warp >> 2;
      const int n4 = (p.d_z + 3) / 4, half = (n4 + 1) / 2;
      const int c_lo = ch * half, c_hi = min(n4, c_lo + half);
      float* scratch = s.dg;                                     // [2][TILE][32]
      const float4* dz0 = reinterpret_cast<const float4*>(s.dz + (4 * ng + 0) * ldw);
      const float4* dz1 = reinterpret_cast<const float4*>(s.dz + (4 * ng + 1) * ldw);
      const float4* dz2 = reinterpret_cast<const float4*>(s.dz + (4 * ng + 2) * ldw);
      const float4* dz3 = reinterpret_cast<const float4*>(s.dz + (4 * ng + 3) * ldw);
      if (lane < d_in) {
        const float4* wr = reinterpret_cast<const float4*>(s.W12 + lane * ldw);
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int c = c_lo; c < c_hi; ++c) {
          const float4 w4 = wr[c];
          a0 = dot4(dz0[c], w4, a0);
          a1 = dot4(dz1[c], w4, a1);
          a2 = dot4(dz2[c], w4, a2);
          a3 = dot4(dz3[c], w4, a3);
        }
        float* o = scratch + (ch * TILE + 4 * ng) * 32 + lane;
        o[0] = a0; o[32] = a1; o[64] = a2; o[96] = a3;
      }
    }
    __syncthreads();
    for (int i = tid; i < TILE * d_in; i += NT) {
      const int node = i / d_in, r = i - node * d_in;
      const float a = s.dg[node * 32 + r] + s.dg[(TILE + node) * 32 + r];
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
__global__ void __launch_bounds__(256) node_wgrad_reduce_kernel(const float* __restrict__ partial, int n_parts, int d_a,
                                                                int d_h, int d_z, int d_g, float alpha1,
                                                                float* __restrict__ dW1, float* __restrict__ dW2,
                                                                float* __restrict__ dW3) {
  // 32 outputs x 8 part-groups per block; the 8 group sums are combined in a fixed order (deterministic)
  __shared__ float red[8][33];
  const int d_in = d_a + d_h;
  const int tot = d_in * d_z + d_g * d_h;
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;
  float s = 0.f;
  if (i < tot)
    for (int q = grp; q < n_parts; q += 8) s += partial[(size_t)q * tot + i];
  red[grp][lane] = s;
  __syncthreads();
  if (grp != 0 || i >= tot) return;
  float sum = red[0][lane];
#pragma unroll
  for (int g = 1; g < 8; ++g) sum += red[g][lane];
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
  const int64_t g = CTAS_PER_SM * (int64_t)sms;
  return (int)(tiles < g ? tiles : g);
}

}  // namespace

extern "C" size_t eqnn_node_update_workspace_bytes(int d_a, int d_h, int d_z, int d_g) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return (size_t)CTAS_PER_SM * sms * ((size_t)(d_a + d_h) * d_z + (size_t)d_g * d_h) * sizeof(float);
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
#define EQNN_NODE_FWD(SH)                                                                                         \
  do {                                                                                                            \
    EQNN_CUDA(cudaFuncSetAttribute(node_fwd_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    node_fwd_kernel<SH><<<grid_for(n), NT, smem, as_stream(stream)>>>(p);                                         \
  } while (0)
  const bool generic = getenv("EQNN_NODE_GENERIC") != nullptr;
  if (!generic && shape_matches<ShapeCfg2>(p)) EQNN_NODE_FWD(ShapeCfg2);
  else if (!generic && shape_matches<ShapeCfg5>(p)) EQNN_NODE_FWD(ShapeCfg5);
  else EQNN_NODE_FWD(void);
#undef EQNN_NODE_FWD
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
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(z) & 15) == 0, "node_update_bwd: z must be 16-byte aligned");
  if (n <= 0) return EQNN_OK;
  p.z = const_cast<float*>(z); p.dh_out = dh_out; p.dA = dA; p.dh_in = dh_in;
  const int grid = grid_for(n);
  const size_t per = (size_t)(d_a + d_h) * p.d_z + (size_t)p.d_g * d_h;
  EQNN_CHECK_ARG(ws && ws_bytes >= (size_t)grid * per * sizeof(float), "node_update_bwd: workspace too small");
  p.partial = reinterpret_cast<float*>(ws);
  const size_t smem = smem_floats(p) * sizeof(float);
  EQNN_CHECK_ARG(smem <= 200 * 1024, "node_update_bwd: shared memory budget exceeded");
  cudaStream_t st = as_stream(stream);
#define EQNN_NODE_BWD(SH)                                                                                         \
  do {                                                                                                            \
    EQNN_CUDA(cudaFuncSetAttribute(node_bwd_kernel<SH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    node_bwd_kernel<SH><<<grid, NT, smem, st>>>(p);                                                               \
  } while (0)
  const bool generic = getenv("EQNN_NODE_GENERIC") != nullptr;
  if (!generic && shape_matches<ShapeCfg2>(p)) EQNN_NODE_BWD(ShapeCfg2);
  else if (!generic && shape_matches<ShapeCfg5>(p)) EQNN_NODE_BWD(ShapeCfg5);
  else EQNN_NODE_BWD(void);
#undef EQNN_NODE_BWD
  node_wgrad_reduce_kernel<<<(unsigned)((per + 31) / 32), 256, 0, st>>>(p.partial, grid, d_a, d_h, p.d_z, p.d_g,
                                                                        alpha1, dW1, dW2, dW3);
  EQNN_CHECK_LAUNCH_N(2);
  return EQNN_OK;
}
