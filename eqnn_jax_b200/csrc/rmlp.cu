// Fused radial MLP of the NequIP interaction block (models/nequip.py:55-68), sm_100a.
//
//   W_e = MLP(bessel(|r_e|))      R = e3nn.bessel(d, n_rb, r_cut); n_hidden x [Dense(128)/sqrt(fan_in), gelu/c];
//                                  Dense(n_irr)/sqrt(128) (only the live output columns, SH norm folded in)
//
// eqnn_radial_mlp_fwd: ONE kernel from the edge length to the radial weights.  No activation ever touches HBM:
// per edge the kernel reads 4 bytes (d) and writes 4*n_out bytes (w_t).
//   * persistent CTA per SM, 17 warps: 16 compute warps + 1 MMA warp (one elected thread issues tcgen05.mma);
//   * all weight matrices resident in shared memory as fp16 (hi, lo) images (168 KB for 64 -> 128 -> 128 -> 128 -> 16);
//   * activations live in tensor memory: a 128-edge tile owns two 128-column regions.  A layer's fp32 accumulator
//     sits in one region; the compute warps read it (tcgen05.ld), apply alpha, gelu and the (hi, lo) split and write
//     the next layer's A operand IN PLACE over the columns they just read (tcgen05.st: 32 fp32 columns become 16
//     columns of hi pairs + 16 columns of lo pairs); the next layer's MMAs (A from tensor memory) accumulate into
//     the other region.  The Bessel basis is computed by the compute warps from d and written the same way;
//   * two tiles are in flight per CTA (2 x 256 = all 512 tensor-memory columns): while the compute warps work on one
//     tile's activations the tensor core runs the other tile's layer.
// Three MMAs per product (hi*hi + lo*hi + hi*lo): see rmlp_common.cuh for the precision model.
#include "common.cuh"
#include "rmlp_common.cuh"

#include <stdlib.h>

namespace {

using namespace rmlp;

constexpr int RM_CW = 16;                       // compute warps = CTA warps 0..15: lane quarter w % 4, column group w / 4
constexpr int RM_MMA_WARP = RM_CW;              // warp 16
constexpr int RM_THREADS = (RM_CW + 1) * 32;    // 544
constexpr int RM_CT = RM_CW * 32;               // 512 compute threads
constexpr int RM_MAXL = 4;                      // GEMM layers: n_hidden (<= 3) + the output layer
constexpr int RM_TILE = 128;
constexpr uint32_t RM_FINAL_COL = 112;          // the narrow last accumulator sits in columns [112, 128) of its region

struct RmlpP {
  const float* dist;        // edge length of edge e at dist[e * dist_stride]
  int64_t dist_stride;
  int64_t E;
  float r_cut, pref;        // bessel: pref * sin(j pi d / r_cut) / d
  int n_rb;                 // 16, 32 or 64
  int n_layers;             // GEMM layers F = n_hidden + 1 (2..4)
  int n_out;                // live output columns (<= 16)
  const float* W[RM_MAXL];  // W[l]: element (k, n) at W[l][k * ldw[l] + n]
  int64_t ldw[RM_MAXL];
  float alpha[RM_MAXL];     // 1 / sqrt(fan_in) of layer l
  float act_scale;          // 1 / c of e3nn.normalize_function(activation)
  int act_kind;             // 0 gelu (tanh form), 1 silu
  float* w_t;               // [n_out][ld_wt] column-major radial weights
  int64_t ld_wt;
  // optional stash for eqnn_radial_mlp_bwd: z_stash[l - 1] = z_l for l = 1..n_hidden-1; the last hidden layer stores
  // z_stash[n_hidden - 1] = gelu(z_H) and z_stash[n_hidden] = gelu'(z_H) instead of z_H (see the kernel); each
  // [ceil(E/32)] chunks of [32 column groups][32 rows][4 floats] (the blocked chunk layout of tc_mlp.cu's WgradP)
  float* z_stash[RM_MAXL];
  // shared radial evaluations (pairs.cu): the live row count is a device word (<= E, the capacity every buffer is sized
  // for), and the radial weights are also / instead written as row-major records w_rows[e][ld_rows] (gathered by
  // eqnn_tpconv_fwd_shared through the edge -> evaluation map)
  const int32_t* n_dev;
  float* w_rows;
  int ld_rows;
};

__host__ __device__ inline int rm_K(const RmlpP& p, int l) { return l == 0 ? p.n_rb : 128; }
__host__ __device__ inline int rm_N(const RmlpP& p, int l) { return l == p.n_layers - 1 ? 16 : 128; }

__host__ inline uint32_t rm_smem_bytes(const RmlpP& p) {
  uint32_t b = 0;
  for (int l = 0; l < p.n_layers; ++l) b += 2 * w_half_bytes(rm_K(p, l), rm_N(p, l));
  return b + 1024 /*align*/ + 256 /*bessel frequencies*/ + 64 /*barriers*/;
}

// The scalar factors of the MLP are folded into the weight images: image l holds W_l * alpha_l (* act_scale for
// l >= 1, the factor of the activation that feeds it), so an accumulator IS the pre-activation z_l and the operand
// written back is the bare gelu_tanh(z_l).
__host__ __device__ inline float rm_wscale(const float* alpha, float act_scale, int l) {
  return l == 0 ? alpha[0] : alpha[l] * act_scale;
}

// z -> gelu_tanh(z) -> (hi, lo) words, 32 accumulator columns of one row.  (A variant with one reciprocal shared by
// four sigmoids -- 5 instead of 8 MUFU per four values -- measured 3.5 % slower: the kernel is bound by issue slots
// and latency, not by the special-function unit.)
__device__ __forceinline__ void act_split32(const float* v, const ActC& ac, uint32_t* hiw, uint32_t* low) {
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const f32x2 x = pk2(v[2 * j], v[2 * j + 1]);
    const f32x2 a = mul2(x, act_sigmoid2(x, mul2(x, x), ac));
    float a0, a1;
    upk2(a, a0, a1);
    split_h2(a0, a1, hiw[j], low[j]);
  }
}

// NV values of row e, columns [c0, c0 + NV) of a [E][ncols] matrix
template <int NV>
__device__ __forceinline__ void emit_cols(float* basep, int blocked, int64_t e, int ncols, int c0, const float* v) {
  if (blocked) {
    float* dst = basep + (e >> 5) * (int64_t)(32 * ncols) + (int64_t)(c0 >> 2) * 128 + (e & 31) * 4;
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      *reinterpret_cast<float4*>(dst + i * 128) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  } else {
    float* dst = basep + e * ncols + c0;
#pragma unroll
    for (int i = 0; i < NV / 4; ++i)
      *reinterpret_cast<float4*>(dst + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
  }
}

__global__ void __launch_bounds__(RM_THREADS, 1) rmlp_fwd_kernel(const RmlpP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = p.n_layers;
  const int64_t E = p.n_dev ? min((int64_t)__ldg(p.n_dev), p.E) : p.E;   // live rows
  const ActC ac = act_constants(p.act_kind);

  // shared-memory map: weight images | bessel frequencies | barriers
  uint32_t w_off[RM_MAXL], w_half[RM_MAXL];
  uint32_t off = 0;
#pragma unroll
  for (int l = 0; l < RM_MAXL; ++l) {
    if (l < F) {
      w_off[l] = off;
      w_half[l] = w_half_bytes(rm_K(p, l), rm_N(p, l));
      off += 2 * w_half[l];
    } else {
      w_off[l] = w_half[l] = 0;
    }
  }
  float* wfreq = reinterpret_cast<float*>(gen + off);
  const uint32_t bars = base + off + 256;
  auto opnd_full = [&](int x) { return bars + 8u * x; };          // compute -> MMA: A operand of slot x written
  auto acc_full = [&](int x) { return bars + 8u * (2 + x); };     // MMA -> compute: accumulator of slot x complete
  const uint32_t tmem_slot = bars + 32;

  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(opnd_full(x), RM_CT);
      mbar_init(acc_full(x), 1);
    }
    mbar_fence_init();
  }
  if (warp == RM_MMA_WARP) tmem_alloc(tmem_slot, 512);
  for (int l = 0; l < F; ++l)
    fill_weight_image(gen + w_off[l], gen + w_off[l] + w_half[l], p.W[l], p.ldw[l], rm_K(p, l), rm_N(p, l),
                      l == F - 1 ? p.n_out : 128, rm_wscale(p.alpha, p.act_scale, l), tid, RM_THREADS);
  if (tid < p.n_rb) wfreq[tid] = __fdiv_rn(__fmul_rn((float)(tid + 1), 3.14159265358979323846f), p.r_cut);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int64_t n_tiles = (E + RM_TILE - 1) / RM_TILE;
  const int64_t my_tiles = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_pairs = (my_tiles + 1) / 2;
  // region r (0 / 1) of slot x; layer l (1-based) accumulates into region (l - 1) & 1 and reads its operand from the other
  auto region = [&](int x, int r) { return tmem_base + (uint32_t)(x * 256 + r * 128); };

  if (warp < RM_CW) {
    // =========================== compute warps ===========================
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    uint32_t n_acc[2] = {0u, 0u};      // completed accumulator phases per slot
    for (int64_t pr = 0; pr < n_pairs; ++pr) {
      const int nx = (2 * pr + 1 < my_tiles) ? 2 : 1;
      // ---- Bessel basis -> A operand of layer 1 (region 1) ----
      for (int x = 0; x < nx; ++x) {
        const int64_t tile = blockIdx.x + (2 * pr + x) * (int64_t)gridDim.x;
        const int64_t e = tile * RM_TILE + row;
        if (16 * cg < p.n_rb) {
          const float d = (e < E) ? __ldg(p.dist + e * p.dist_stride) : 0.f;
          float v[16];
          bessel16(wfreq, 16 * cg, d, p.pref, v);
          uint32_t hiw[8], low[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_h2(v[2 * j], v[2 * j + 1], hiw[j], low[j]);
          const uint32_t t = region(x, 1) + lane_sel + (uint32_t)(32 * (cg >> 1) + 8 * (cg & 1));
          tmem_st8(t, hiw);
          tmem_st8(t + 16, low);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(opnd_full(x));
      }
      // ---- hidden layers: accumulator -> gelu -> next operand, in place ----
      for (int l = 1; l < F; ++l) {
        for (int x = 0; x < nx; ++x) {
          mbar_wait(acc_full(x), n_acc[x] & 1u);
          ++n_acc[x];
          tc_fence_after();
          const uint32_t t = region(x, (l - 1) & 1) + lane_sel + (uint32_t)(32 * cg);
          float v[32];
          tmem_ld32(t, v);
          tmem_ld_wait();
          uint32_t hiw[16], low[16];
          if (p.z_stash[l - 1] && l == F - 1) {
            // last hidden layer: the backward pass needs gelu(z_H) (output layer's weight gradient) and gelu'(z_H)
            // (g_H = (dw @ W_out^T) * gelu'(z_H)) and never z_H itself: both come out of the sigmoid evaluated here anyway
            const int64_t e = (blockIdx.x + (2 * pr + x) * (int64_t)gridDim.x) * RM_TILE + row;
            float* da = p.z_stash[l - 1] + (e >> 5) * (int64_t)(32 * 128) + (int64_t)(8 * cg) * 128 + (e & 31) * 4;
            float* dg = p.z_stash[l] + (e >> 5) * (int64_t)(32 * 128) + (int64_t)(8 * cg) * 128 + (e & 31) * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              f32x2 a01, g01, a23, g23;
              act_both2(pk2(v[4 * i], v[4 * i + 1]), ac, a01, g01);
              act_both2(pk2(v[4 * i + 2], v[4 * i + 3]), ac, a23, g23);
              float4 a4, g4;
              upk2(a01, a4.x, a4.y); upk2(a23, a4.z, a4.w);
              upk2(g01, g4.x, g4.y); upk2(g23, g4.z, g4.w);
              if (e < E) {
                *reinterpret_cast<float4*>(da + i * 128) = a4;
                *reinterpret_cast<float4*>(dg + i * 128) = g4;
              }
              split_h2(a4.x, a4.y, hiw[2 * i], low[2 * i]);
              split_h2(a4.z, a4.w, hiw[2 * i + 1], low[2 * i + 1]);
            }
          } else {
            if (p.z_stash[l - 1]) {
              const int64_t e = (blockIdx.x + (2 * pr + x) * (int64_t)gridDim.x) * RM_TILE + row;
              if (e < E) emit_cols<32>(p.z_stash[l - 1], 1, e, 128, 32 * cg, v);
            }
            act_split32(v, ac, hiw, low);
          }
          tmem_st16(t, hiw);
          tmem_st16(t + 16, low);
          tmem_st_wait();
          tc_fence_before();
          mbar_arrive(opnd_full(x));
        }
      }
      // ---- output layer: alpha * accumulator -> w_t ----
      for (int x = 0; x < nx; ++x) {
        mbar_wait(acc_full(x), n_acc[x] & 1u);
        ++n_acc[x];
        tc_fence_after();
        if (cg == 0) {
          const int64_t tile = blockIdx.x + (2 * pr + x) * (int64_t)gridDim.x;
          const int64_t e = tile * RM_TILE + row;
          float v[16];
          tmem_ld16(region(x, (F - 1) & 1) + lane_sel + RM_FINAL_COL, v);
          tmem_ld_wait();
          if (e < E) {
            if (p.w_t) {
#pragma unroll
              for (int c = 0; c < 16; ++c)
                if (c < p.n_out) p.w_t[(int64_t)c * p.ld_wt + e] = v[c];
            }
            if (p.w_rows) {
              float4* wr = reinterpret_cast<float4*>(p.w_rows + e * p.ld_rows);
#pragma unroll
              for (int i = 0; i < 4; ++i)
                if (4 * i < p.ld_rows) wr[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
            }
          }
        }
      }
    }
  } else if (lane == 0) {
    // =========================== MMA issuer ===========================
    uint32_t n_op[2] = {0u, 0u};
    for (int64_t pr = 0; pr < n_pairs; ++pr) {
      const int nx = (2 * pr + 1 < my_tiles) ? 2 : 1;
      for (int l = 1; l <= F; ++l) {
        const int K = rm_K(p, l - 1), N = rm_N(p, l - 1);
        const uint32_t idesc = make_idesc_f16(128, N, 0, 0);
        const uint32_t b_hi = base + w_off[l - 1], b_lo = b_hi + w_half[l - 1];
        for (int x = 0; x < nx; ++x) {
          mbar_wait(opnd_full(x), n_op[x] & 1u);
          ++n_op[x];
          tc_fence_after();
          const uint32_t d_tmem = region(x, (l - 1) & 1) + (l == F ? RM_FINAL_COL : 0u);
          const uint32_t a_reg = region(x, l & 1);
          // The accumulator in tensor memory truncates on every update, in proportion to its current magnitude: the
          // small cross terms (lo*hi, hi*lo) go first, the hi*hi terms that carry the value go last (pass 0 / pass 1).
          for (int pass = 0; pass < 2; ++pass)
            for (int s = 0; s < K / 16; ++s) {
              const uint32_t a_hi = a_reg + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
              const uint32_t boff = (uint32_t)(s >> 2) * (uint32_t)N * 128u + (uint32_t)(s & 3) * 32u;
              const uint64_t dbh = make_desc(b_hi + boff, 16, 1024), dbl = make_desc(b_lo + boff, 16, 1024);
              if (pass == 0) {
                mma_f16_ts(d_tmem, a_lo, dbh, idesc, s ? 1u : 0u);
                mma_f16_ts(d_tmem, a_hi, dbl, idesc, 1u);
              } else {
                mma_f16_ts(d_tmem, a_hi, dbh, idesc, 1u);
              }
            }
          mma_commit(acc_full(x));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Backward chain: per 128-edge tile, recompute the forward activations from d, then run the data-gradient chain
//   g_H = (dw @ W_out^T) * gelu'(z_H);  g_{l-1} = (g_l @ W_l^T) * gelu'(z_{l-1})        (weight images already carry
// the alpha / act_scale factors, which are linear and therefore the same in both directions), everything in tensor
// memory.  The kernel emits what the weight gradients need: the basis a_0, the activations gelu(z_l) and the
// gradients g_l = dL/dz_l (fp32, row-major).  One tile is in flight: it owns all four 128-column regions --
// two for the ping-pong chain (accumulator / operand in place, as in the forward kernel) and two that stash
// gelu'(z_1), gelu'(z_2) from the recompute phase (the sigmoid is evaluated once per pre-activation).
// Gradients are brought into fp16 range by one power-of-two scale per launch (from max |dw|): only sums over edges
// are ever formed from them, so what matters is the error relative to the LARGEST rows, which keeps 22 bits.
// ------------------------------------------------------------------------------------------------------------------
struct RmlpBwdP {
  RmlpP f;                   // forward description (f.w_t unused)
  const float* dw;           // dL/dw_t, [n_out][ld_dw] column-major
  int64_t ld_dw;
  const uint32_t* absmax;    // bit pattern of max |dw| (absmax_kernel)
  float* act[RM_MAXL];       // act[0] = bessel basis [E][n_rb]; act[l] = gelu(z_l) [E][128], l = 1..H
  float* grad[RM_MAXL];      // grad[l] = dL/dz_l [E][128], l = 1..H
  // 1: emit in the blocked chunk layout of tc_mlp.cu's weight-gradient kernel -- per 32-edge chunk
  // [16-byte column group][row][4 floats]: a warp (one row per lane) writes 512 contiguous bytes per store;
  // 0: plain row-major (small problems that take the CUDA-core GEMM)
  // blk[l] is the layout of the two operands of layer l's weight gradient, act[l] and grad[l + 1]
  int blk[RM_MAXL];
};

constexpr uint32_t RM_GSLAB = 128 * 128;   // dw operand: 128 rows x 128 B (16 halves used), K-major, 128B swizzle

__global__ void absmax_kernel(const float* __restrict__ x, int64_t ld, int rows, int64_t cols, uint32_t* __restrict__ out,
                              const int32_t* __restrict__ n_dev) {
  if (n_dev) cols = min((int64_t)__ldg(n_dev), cols);
  uint32_t m = 0;
  for (int r = 0; r < rows; ++r)
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < cols; i += (int64_t)gridDim.x * blockDim.x) {
      const uint32_t b = __float_as_uint(__ldg(x + (int64_t)r * ld + i)) & 0x7fffffffu;
      m = b > m ? b : m;
    }
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__global__ void __launch_bounds__(RM_THREADS, 1) rmlp_bwd_kernel(const RmlpBwdP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const RmlpP& f = p.f;
  const int F = f.n_layers, H = F - 1;
  const ActC ac = act_constants(f.act_kind);

  // shared-memory map: weight images | dw operand slabs (hi, lo) | bessel frequencies | barriers
  uint32_t w_off[RM_MAXL], w_half[RM_MAXL];
  uint32_t off = 0;
#pragma unroll
  for (int l = 0; l < RM_MAXL; ++l) {
    if (l < F) {
      w_off[l] = off;
      w_half[l] = w_half_bytes(rm_K(f, l), rm_N(f, l));
      off += 2 * w_half[l];
    } else {
      w_off[l] = w_half[l] = 0;
    }
  }
  const uint32_t gslab = off;
  off += 2 * RM_GSLAB;
  float* wfreq = reinterpret_cast<float*>(gen + off);
  const uint32_t bars = base + off + 256;
  const uint32_t opnd_full = bars, acc_full = bars + 8, mma_done = bars + 16, tmem_slot = bars + 24;

  if (tid == 0) {
    mbar_init(opnd_full, RM_CT);
    mbar_init(acc_full, 1);
    mbar_init(mma_done, 1);
    mbar_fence_init();
  }
  if (warp == RM_MMA_WARP) tmem_alloc(tmem_slot, 512);
  for (int l = 0; l < F; ++l)
    fill_weight_image(gen + w_off[l], gen + w_off[l] + w_half[l], f.W[l], f.ldw[l], rm_K(f, l), rm_N(f, l),
                      l == F - 1 ? f.n_out : 128, rm_wscale(f.alpha, f.act_scale, l), tid, RM_THREADS);
  for (uint32_t o = tid * 16; o < 2 * RM_GSLAB; o += RM_THREADS * 16)
    *reinterpret_cast<float4*>(gen + gslab + o) = make_float4(0.f, 0.f, 0.f, 0.f);
  if (tid < f.n_rb) wfreq[tid] = __fdiv_rn(__fmul_rn((float)(tid + 1), 3.14159265358979323846f), f.r_cut);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int64_t n_tiles = (f.E + RM_TILE - 1) / RM_TILE;
  auto region = [&](int r) { return tmem_base + (uint32_t)((r & 1) * 128); };     // chain ping-pong
  auto stash = [&](int l) { return tmem_base + 256u + (uint32_t)((l - 1) * 128); };   // gelu'(z_l), l = 1, 2

  // gradient scale: max |dw| * S in [4, 8)
  const uint32_t mbits = __ldg(p.absmax);
  const int mexp = (int)((mbits >> 23) & 0xffu);
  int sexp = 127 + 2 - (mexp - 127);
  sexp = sexp < 1 ? 1 : (sexp > 254 ? 254 : sexp);
  const float gS0 = (mexp == 0) ? 1.f : __uint_as_float((uint32_t)sexp << 23);
  const float gSinv0 = (mexp == 0) ? 1.f : __uint_as_float((uint32_t)(254 - sexp) << 23);

  if (warp < RM_CW) {
    // =========================== compute warps ===========================
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    uint32_t n_acc = 0;
    auto wait_acc = [&]() {
      mbar_wait(acc_full, n_acc & 1u);
      ++n_acc;
      tc_fence_after();
    };
    auto publish = [&]() {
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(opnd_full);
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t e = tile * RM_TILE + row;
      const bool live = e < f.E;
      const float gS = gS0, gSinv = gSinv0;
      // ---- c0: Bessel basis -> operand of layer 1 (k-step s of the basis sits at column 32 s of region 1) ----
      if (16 * cg < f.n_rb) {
        const float d = live ? __ldg(f.dist + e * f.dist_stride) : 0.f;
        uint32_t hiw[8], low[8];
        float v[16];
        bessel16(wfreq, 16 * cg, d, f.pref, v);
#pragma unroll
        for (int j = 0; j < 8; ++j) split_h2(v[2 * j], v[2 * j + 1], hiw[j], low[j]);
        const uint32_t t = region(1) + lane_sel + (uint32_t)(32 * cg);
        tmem_st8(t, hiw);
        tmem_st8(t + 16, low);
        if (live) emit_cols<16>(p.act[0], p.blk[0], e, f.n_rb, 16 * cg, v);
      }
      if (cg == 0) {
        // dw row (scaled) -> K-major operand slab: 16 halves = chunks 0, 1 of this row's 128-byte line
        float g[16];
#pragma unroll
        for (int c = 0; c < 16; ++c) g[c] = (live && c < f.n_out) ? gS * __ldg(p.dw + (int64_t)c * p.ld_dw + e) : 0.f;
        uint32_t hiw[8], low[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) split_h2(g[2 * j], g[2 * j + 1], hiw[j], low[j]);
        uint8_t* sl = gen + gslab;
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          const uint32_t o = sw128((uint32_t)row, (uint32_t)c);
          *reinterpret_cast<uint4*>(sl + o) = make_uint4(hiw[4 * c], hiw[4 * c + 1], hiw[4 * c + 2], hiw[4 * c + 3]);
          *reinterpret_cast<uint4*>(sl + RM_GSLAB + o) = make_uint4(low[4 * c], low[4 * c + 1], low[4 * c + 2], low[4 * c + 3]);
        }
        fence_proxy_async_smem();
      }
      publish();
      // ---- recompute: z_l -> gelu (operand, in place) and gelu' (stash); the last hidden layer meets the gradient ----
      for (int l = 1; l <= H; ++l) {
        wait_acc();
        const uint32_t tz = region(l - 1) + lane_sel + (uint32_t)(32 * cg);
        float z[32], a[32], gp[32];
        tmem_ld32(tz, z);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          f32x2 g2, gp2;
          act_both2(pk2(z[2 * j], z[2 * j + 1]), ac, g2, gp2);
          upk2(g2, a[2 * j], a[2 * j + 1]);
          upk2(gp2, gp[2 * j], gp[2 * j + 1]);
        }
        if (live) emit_cols<32>(p.act[l], p.blk[l], e, 128, 32 * cg, a);
        if (l < H) {
          uint32_t hiw[16], low[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_h2(a[2 * j], a[2 * j + 1], hiw[j], low[j]);
          tmem_st16(tz, hiw);
          tmem_st16(tz + 16, low);
          tmem_st32(stash(l) + lane_sel + (uint32_t)(32 * cg), reinterpret_cast<const uint32_t*>(gp));
          publish();
        } else {
          // g_H = (dw @ W_out^T) * gelu'(z_H): the product sits in the other region
          const uint32_t tg = region(l) + lane_sel + (uint32_t)(32 * cg);
          float pre[32];
          tmem_ld32(tg, pre);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) pre[j] *= gp[j];
          if (H >= 2) {
            uint32_t hiw[16], low[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_h2(pre[2 * j], pre[2 * j + 1], hiw[j], low[j]);
            tmem_st16(tg, hiw);
            tmem_st16(tg + 16, low);
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) pre[j] *= gSinv;
          if (live) emit_cols<32>(p.grad[l], p.blk[l - 1], e, 128, 32 * cg, pre);
          if (H >= 2) publish();
        }
      }
      // ---- data-gradient chain: g_l = (g_{l+1} @ W_{l+1}^T) * gelu'(z_l), l = H-1 .. 1 ----
      for (int l = H - 1; l >= 1; --l) {
        wait_acc();
        // g_H sits in region(H); every product lands in the other region: g_l in region(H + (H - l))
        const uint32_t tg = region(2 * H - l) + lane_sel + (uint32_t)(32 * cg);
        float pre[32], gp[32];
        tmem_ld32(tg, pre);
        tmem_ld32(stash(l) + lane_sel + (uint32_t)(32 * cg), gp);
        tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; ++j) pre[j] *= gp[j];
        if (l >= 2) {
          uint32_t hiw[16], low[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) split_h2(pre[2 * j], pre[2 * j + 1], hiw[j], low[j]);
          tmem_st16(tg, hiw);
          tmem_st16(tg + 16, low);
        }
#pragma unroll
        for (int j = 0; j < 32; ++j) pre[j] *= gSinv;
        if (live) emit_cols<32>(p.grad[l], p.blk[l - 1], e, 128, 32 * cg, pre);
        if (l >= 2) publish();
      }
    }
  } else if (lane == 0) {
    // =========================== MMA issuer ===========================
    uint32_t n_op = 0, n_done = 0;
    auto wait_opnd = [&]() {
      mbar_wait(opnd_full, n_op & 1u);
      ++n_op;
      tc_fence_after();
    };
    // forward layer l (1-based): D = region(l-1), A = region(l) [basis: k-step s at column 32 s], B = image l-1 (K-major)
    auto issue_fwd = [&](int l) {
      const int K = rm_K(f, l - 1);
      const uint32_t idesc = make_idesc_f16(128, 128, 0, 0);
      const uint32_t b_hi = base + w_off[l - 1], b_lo = b_hi + w_half[l - 1];
      const uint32_t d_tmem = region(l - 1), a_reg = region(l);
      for (int pass = 0; pass < 2; ++pass)     // cross terms first, hi*hi last (see rmlp_fwd_kernel)
        for (int s = 0; s < K / 16; ++s) {
          const uint32_t a_hi = a_reg + (l == 1 ? (uint32_t)(32 * s) : (uint32_t)(32 * (s >> 1) + 8 * (s & 1))), a_lo = a_hi + 16;
          const uint32_t boff = (uint32_t)(s >> 2) * 128u * 128u + (uint32_t)(s & 3) * 32u;
          const uint64_t dbh = make_desc(b_hi + boff, 16, 1024), dbl = make_desc(b_lo + boff, 16, 1024);
          if (pass == 0) {
            mma_f16_ts(d_tmem, a_lo, dbh, idesc, s ? 1u : 0u);
            mma_f16_ts(d_tmem, a_hi, dbl, idesc, 1u);
          } else {
            mma_f16_ts(d_tmem, a_hi, dbh, idesc, 1u);
          }
        }
    };
    // data gradient through layer l (1-based, 2 <= l <= H): D (N = 128 inputs k) = A (g_l, K = 128 outputs n) x image l-1
    // read MN-major: k-step s = rows n = 16 s .. 16 s + 15 of both 64-k slabs (LBO = slab stride, SBO = 8 rows)
    auto issue_dgrad = [&](int l, uint32_t a_reg, uint32_t d_tmem) {
      const uint32_t idesc = make_idesc_f16(128, 128, 0, 1);
      const uint32_t b_hi = base + w_off[l - 1], b_lo = b_hi + w_half[l - 1];
      for (int pass = 0; pass < 2; ++pass)
        for (int s = 0; s < 8; ++s) {
          const uint32_t a_hi = a_reg + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
          const uint64_t dbh = make_desc(b_hi + s * 2048u, 128u * 128u, 1024), dbl = make_desc(b_lo + s * 2048u, 128u * 128u, 1024);
          if (pass == 0) {
            mma_f16_ts(d_tmem, a_lo, dbh, idesc, s ? 1u : 0u);
            mma_f16_ts(d_tmem, a_hi, dbl, idesc, 1u);
          } else {
            mma_f16_ts(d_tmem, a_hi, dbh, idesc, 1u);
          }
        }
    };
    for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      for (int l = 1; l <= H; ++l) {
        wait_opnd();
        issue_fwd(l);
        if (l < H) {
          mma_commit(acc_full);
        } else {
          // the output layer's data gradient overwrites region(H), which layer H is still reading: wait for it
          mma_commit(mma_done);
          mbar_wait(mma_done, n_done & 1u);
          ++n_done;
          tc_fence_after();
          const uint32_t idesc = make_idesc_f16(128, 128, 0, 1);
          const uint32_t g_hi = base + gslab, g_lo = g_hi + RM_GSLAB;
          const uint32_t b_hi = base + w_off[F - 1], b_lo = b_hi + w_half[F - 1];
          const uint64_t dah = make_desc(g_hi, 16, 1024), dal = make_desc(g_lo, 16, 1024);
          const uint64_t dbh = make_desc(b_hi, 16u * 128u, 1024), dbl = make_desc(b_lo, 16u * 128u, 1024);
          mma_f16_ss(region(H), dal, dbh, idesc, 0u);
          mma_f16_ss(region(H), dah, dbl, idesc, 1u);
          mma_f16_ss(region(H), dah, dbh, idesc, 1u);
          mma_commit(acc_full);
        }
      }
      for (int l = H; l >= 2; --l) {
        wait_opnd();
        issue_dgrad(l, region(2 * H - l), region(2 * H - l + 1));
        mma_commit(acc_full);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}


// ------------------------------------------------------------------------------------------------------------------
// Backward chain from stashed pre-activations (the default): the forward kernel has written z_1 .. z_H once; this
// kernel runs only the data-gradient chain, g_H = (dw @ W_out^T) * gelu'(z_H), g_{l-1} = (g_l @ W_l^T) * gelu'(z_{l-1}),
// with the structure of the forward kernel: accumulator and operand in place in tensor memory, two tiles in flight
// (no stash regions are needed: gelu'(z_l) is formed from the z_l rows loaded straight into registers, prefetched
// before the accumulator barrier).  It emits g_l for the weight-gradient kernel.
// ------------------------------------------------------------------------------------------------------------------
struct RmlpDgradP {
  RmlpP f;
  const float* dw;
  int64_t ld_dw;
  const uint32_t* absmax;
  const float* z[RM_MAXL];   // z[l] = z_l, l = 1..H-1, and z[H] = gelu'(z_H); blocked chunk layout
  float* grad[RM_MAXL];      // grad[l] = dL/dz_l, l = 1..H
  int blk[RM_MAXL];          // blk[l - 1]: layout of grad[l] (1 blocked, 0 row-major)
};

__device__ __forceinline__ void load_blocked32(const float* basep, int64_t e, int c0, float* v) {
  const float* src = basep + (e >> 5) * (int64_t)(32 * 128) + (int64_t)(c0 >> 2) * 128 + (e & 31) * 4;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4*>(src + i * 128));
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}

__global__ void __launch_bounds__(RM_THREADS, 1) rmlp_dgrad_kernel(const RmlpDgradP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const RmlpP& f = p.f;
  const int F = f.n_layers, H = F - 1;
  const ActC ac = act_constants(f.act_kind);

  // shared-memory map: weight images of layers 2..F (layer 1 is not differentiated through) | barriers
  uint32_t w_off[RM_MAXL], w_half[RM_MAXL];
  uint32_t off = 0;
#pragma unroll
  for (int l = 0; l < RM_MAXL; ++l) {
    if (l >= 1 && l < F) {
      w_off[l] = off;
      w_half[l] = w_half_bytes(rm_K(f, l), rm_N(f, l));
      off += 2 * w_half[l];
    } else {
      w_off[l] = w_half[l] = 0;
    }
  }
  const uint32_t bars = base + off;
  auto opnd_full = [&](int x) { return bars + 8u * x; };
  auto acc_full = [&](int x) { return bars + 8u * (2 + x); };
  const uint32_t tmem_slot = bars + 32;

  if (tid == 0) {
    for (int x = 0; x < 2; ++x) {
      mbar_init(opnd_full(x), RM_CT);
      mbar_init(acc_full(x), 1);
    }
    mbar_fence_init();
  }
  if (warp == RM_MMA_WARP) tmem_alloc(tmem_slot, 512);
  for (int l = 1; l < F; ++l)
    fill_weight_image(gen + w_off[l], gen + w_off[l] + w_half[l], f.W[l], f.ldw[l], rm_K(f, l), rm_N(f, l),
                      l == F - 1 ? f.n_out : 128, rm_wscale(f.alpha, f.act_scale, l), tid, RM_THREADS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int64_t n_tiles = (f.E + RM_TILE - 1) / RM_TILE;
  const int64_t my_tiles = (blockIdx.x < n_tiles) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const int64_t n_pairs = (my_tiles + 1) / 2;
  auto region = [&](int x, int r) { return tmem_base + (uint32_t)(x * 256 + (r & 1) * 128); };

  const uint32_t mbits = __ldg(p.absmax);
  const int mexp = (int)((mbits >> 23) & 0xffu);
  int sexp = 127 + 2 - (mexp - 127);
  sexp = sexp < 1 ? 1 : (sexp > 254 ? 254 : sexp);
  const float gS0 = (mexp == 0) ? 1.f : __uint_as_float((uint32_t)sexp << 23);
  const float gSinv0 = (mexp == 0) ? 1.f : __uint_as_float((uint32_t)(254 - sexp) << 23);

  if (warp < RM_CW) {
    // =========================== compute warps ===========================
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    uint32_t n_acc[2] = {0u, 0u};
    for (int64_t pr = 0; pr < n_pairs; ++pr) {
      const int nx = (2 * pr + 1 < my_tiles) ? 2 : 1;
      // ---- dw row (scaled) -> operand in region 1, columns [0, 8) + [16, 24) ----
      for (int x = 0; x < nx; ++x) {
        const int64_t tile = blockIdx.x + (2 * pr + x) * (int64_t)gridDim.x;
        const int64_t e = tile * RM_TILE + row;
        if (cg == 0) {
          const float gS = gS0;
          float g[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) g[c] = (e < f.E && c < f.n_out) ? gS * __ldg(p.dw + (int64_t)c * p.ld_dw + e) : 0.f;
          uint32_t hiw[8], low[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) split_h2(g[2 * j], g[2 * j + 1], hiw[j], low[j]);
          const uint32_t t = region(x, 1) + lane_sel;
          tmem_st8(t, hiw);
          tmem_st8(t + 16, low);
          tmem_st_wait();
        }
        tc_fence_before();
        mbar_arrive(opnd_full(x));
      }
      // ---- g_l = product * gelu'(z_l), l = H .. 1; g_l sits in region (H - l) ----
      for (int l = H; l >= 1; --l) {
        for (int x = 0; x < nx; ++x) {
          const int64_t tile = blockIdx.x + (2 * pr + x) * (int64_t)gridDim.x;
          const int64_t e = tile * RM_TILE + row;
          const bool live = e < f.E;
          float z[32];
          if (live) {
            load_blocked32(p.z[l], e, 32 * cg, z);
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) z[j] = 0.f;
          }
          mbar_wait(acc_full(x), n_acc[x] & 1u);
          ++n_acc[x];
          tc_fence_after();
          const uint32_t t = region(x, H - l) + lane_sel + (uint32_t)(32 * cg);
          float pre[32];
          tmem_ld32(t, pre);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const f32x2 zz = pk2(z[2 * j], z[2 * j + 1]);
            const f32x2 g2 = mul2(pk2(pre[2 * j], pre[2 * j + 1]), l == H ? zz : act_grad2(zz, ac));   // z[H] holds gelu'(z_H)
            upk2(g2, pre[2 * j], pre[2 * j + 1]);
          }
          if (l >= 2) {
            uint32_t hiw[16], low[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) split_h2(pre[2 * j], pre[2 * j + 1], hiw[j], low[j]);
            tmem_st16(t, hiw);
            tmem_st16(t + 16, low);
            tmem_st_wait();
            tc_fence_before();
            mbar_arrive(opnd_full(x));
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) pre[j] *= gSinv0;
          if (live) emit_cols<32>(p.grad[l], p.blk[l - 1], e, 128, 32 * cg, pre);
        }
      }
    }
  } else if (lane == 0) {
    // =========================== MMA issuer ===========================
    uint32_t n_op[2] = {0u, 0u};
    const uint32_t idesc = make_idesc_f16(128, 128, 0, 1);
    for (int64_t pr = 0; pr < n_pairs; ++pr) {
      const int nx = (2 * pr + 1 < my_tiles) ? 2 : 1;
      for (int l = F; l >= 2; --l) {
        // data gradient through layer l (1-based): A = g_l (the dw rows for l = F), B = image of layer l read MN-major
        const uint32_t b_hi = base + w_off[l - 1], b_lo = b_hi + w_half[l - 1];
        for (int x = 0; x < nx; ++x) {
          mbar_wait(opnd_full(x), n_op[x] & 1u);
          ++n_op[x];
          tc_fence_after();
          if (l == F) {
            const uint64_t dbh = make_desc(b_hi, 16u * 128u, 1024), dbl = make_desc(b_lo, 16u * 128u, 1024);
            const uint32_t a_hi = region(x, 1), a_lo = a_hi + 16, d_tmem = region(x, 0);
            mma_f16_ts(d_tmem, a_lo, dbh, idesc, 0u);
            mma_f16_ts(d_tmem, a_hi, dbl, idesc, 1u);
            mma_f16_ts(d_tmem, a_hi, dbh, idesc, 1u);
          } else {
            const uint32_t a_reg = region(x, H - l), d_tmem = region(x, H - l + 1);
            for (int pass = 0; pass < 2; ++pass)     // cross terms first, hi*hi last (see rmlp_fwd_kernel)
              for (int s = 0; s < 8; ++s) {
                const uint32_t a_hi = a_reg + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
                const uint64_t dbh = make_desc(b_hi + s * 2048u, 128u * 128u, 1024), dbl = make_desc(b_lo + s * 2048u, 128u * 128u, 1024);
                if (pass == 0) {
                  mma_f16_ts(d_tmem, a_lo, dbh, idesc, s ? 1u : 0u);
                  mma_f16_ts(d_tmem, a_hi, dbl, idesc, 1u);
                } else {
                  mma_f16_ts(d_tmem, a_hi, dbh, idesc, 1u);
                }
              }
          }
          mma_commit(acc_full(x));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == RM_MMA_WARP) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Fused backward of ONE hidden layer l (2 <= l <= H; W_l maps gelu(z_{l-1}) to z_l), from the z stash:
//     g_{l-1} = (g_l @ W_l^T) * gelu'(z_{l-1})                      data gradient   (tcgen05, A = g_l in tensor memory)
//     dW_l    = alpha_l c * gelu(z_{l-1})^T @ g_l                   weight gradient (tcgen05, K = edges: both operands
//                                                                   edge-major in shared memory, read MN-major)
// in one pass over the edges: z_{l-1} and g_l cross HBM once (the separate chain + weight-gradient kernels read z and
// g twice and wrote every g_l for the second reader).  For l = H the kernel forms g_H itself,
// g_H = (dw @ W_out^T) * gelu'(z_H), so g_H never touches memory at all.
// Per 128-edge tile (one in flight), tensor memory = OP | PRE | ACC | X (4 x 128 columns), 20 warps:
//   g path (warps 16..19, one row per thread, four passes of 32 columns; the rows of a warp are one 32-edge chunk of
//            the blocked layout, so a pass is ONE 4 KB bulk copy into the warp's landing buffer):
//            g_l -> (hi, lo) -> G tile in shared memory AND in place into OP                           [publish g]
//            (l = H: g_H = OP * gelu'(z_H) with OP = dw @ W_out^T from a small prologue MMA issued one tile ahead)
//   MMA (lane 0 of warp 0, issued where the a path waits for PRE):  PRE = OP @ W_l^T (24 MMAs);  ACC += A^T G (24 MMAs)
//   a path (warps 0..15, 32 columns of one row per thread, z_{l-1} prefetched into registers one tile ahead):
//            gelu(z_{l-1}) -> (hi, lo) -> A tile in shared memory, gelu' kept in registers              [publish a]
//            g_{l-1} = PRE * gelu'(z_{l-1}) -> global (blocked layout) while the weight-gradient MMAs run.
// The tensor core's fp32 accumulation truncates, so ACC spans RM_LBWD_SEG tiles (384 updates) and is then added to
// the CTA's partial sums in global memory with IEEE adds.
// Measured (in-kernel cycle counters, 3.2 M edges): a global->SM round trip costs 2-3 k cycles under load for loads and
// bulk copies alike, and L2 prefetches (bulk or per line) only add to the queues -- so every stream is prefetched by the
// threads that consume it (z_{l-1}: registers, one tile ahead; g_l / z_H: the landing ring) and nothing else.  The
// kernels run at 1.18 ms (l < H) and 1.45 ms (l = H) against HBM floors of 0.93 / 0.95 ms; the a path's own work
// (about 10 k cycles per tile) equals the HBM time of a tile, the g path of l = H (gelu' of a whole tile on 4 warps) exceeds it.
// ------------------------------------------------------------------------------------------------------------------
struct RmlpLbwdP {
  int64_t E;
  const float* W;            // W_l: element (k, n) at W[k * ldw + n], 128 x 128
  int64_t ldw;
  float wscale;              // alpha_l * act_scale (folded into the image, as in the other kernels)
  const float* Wout;         // first = 1: output layer 128 x n_out
  int64_t ldwout;
  float wout_scale;
  int n_out;
  int first;                 // 1: l = H, g_l formed from dw and z_l;  0: g_l read from g_in
  const float* dw;           // [n_out][ld_dw] column-major
  int64_t ld_dw;
  const uint32_t* absmax;
  const float* z_l;          // first = 1: gelu'(z_H) (blocked chunk layout; the forward kernel's stash)
  const float* g_in;         // first = 0: g_l (blocked, unscaled)
  const float* z_prev;       // z_{l-1} (blocked)
  float* g_out;              // g_{l-1} (blocked, unscaled)
  float* partial;            // [gridDim.x][n = 128][k = 128] per-CTA weight-gradient sums (unscaled by alpha)
  int act_kind;
  const int32_t* n_dev;      // live row count on the device (<= E), or NULL
};

constexpr uint32_t RM_LBWD_SEG = 16;         // tiles per tensor-memory accumulation of the weight gradient (384 updates)
constexpr uint32_t RM_OPT = 2 * 128 * 128; // one (hi or lo) half of a 128-edge x 128-column fp16 operand tile: 2 slabs of 16 KB

constexpr int LB_GW0 = RM_CW;                    // warps 16..19: the g path (one warp per tensor-memory lane quarter);
                                                 // lane 0 of warp 0 also issues the MMAs (20 warps keep 96 registers/thread)
constexpr int LB_THREADS = (LB_GW0 + 4) * 32;    // 640

__device__ __forceinline__ void sts_chunks(uint8_t* tile, uint32_t row, uint32_t chunk0, const uint32_t* hiw, const uint32_t* low) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t o = sw128(row, chunk0 + i);
    *reinterpret_cast<uint4*>(tile + o) = make_uint4(hiw[4 * i], hiw[4 * i + 1], hiw[4 * i + 2], hiw[4 * i + 3]);
    *reinterpret_cast<uint4*>(tile + RM_OPT + o) = make_uint4(low[4 * i], low[4 * i + 1], low[4 * i + 2], low[4 * i + 3]);
  }
}

__global__ void __launch_bounds__(LB_THREADS, 1) rmlp_lbwd_kernel(const RmlpLbwdP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const ActC ac = act_constants(p.act_kind);

  // shared-memory map: W_l image (hi, lo) | A tile (hi, lo) | G tile (hi, lo) | [W_out image (hi, lo)] | barriers
  const uint32_t w_half = w_half_bytes(128, 128);                 // 32 KB
  const uint32_t wo_half = w_half_bytes(128, 16);                 // 4 KB
  const uint32_t o_w = 0, o_a = 2 * w_half, o_g = o_a + 2 * RM_OPT, o_wo = o_g + 2 * RM_OPT;
  // landing buffers of the g-path bulk copies: 4 warps x NST stages x 4 KB (two stages where W_out is not resident)
  const uint32_t NST = p.first ? 1u : 2u;
  const uint32_t o_land = o_wo + (p.first ? 2 * wo_half : 0u);
  const uint32_t o_bar = o_land + 4 * NST * 4096;
  const uint32_t bars = base + o_bar;
  const uint32_t opnd0_full = bars, pro_full = bars + 8, opnd_g = bars + 16, opnd_a = bars + 24, pre_full = bars + 32,
                 wg_done = bars + 40;
  auto land_full = [&](int q, uint32_t st) { return bars + 48u + 8u * (2u * q + st); };
  const uint32_t tmem_slot = bars + 112;

  if (tid == 0) {
    mbar_init(opnd0_full, 128);
    mbar_init(pro_full, 1);
    mbar_init(opnd_g, 128);
    mbar_init(opnd_a, RM_CT);
    mbar_init(pre_full, 1);
    mbar_init(wg_done, 1);
    for (int q = 0; q < 4; ++q)
      for (uint32_t st = 0; st < 2; ++st) mbar_init(land_full(q, st), 1);
    mbar_fence_init();
  }
  if (warp == LB_GW0) tmem_alloc(tmem_slot, 512);
  fill_weight_image(gen + o_w, gen + o_w + w_half, p.W, p.ldw, 128, 128, 128, p.wscale, tid, LB_THREADS);
  if (p.first)
    fill_weight_image(gen + o_wo, gen + o_wo + wo_half, p.Wout, p.ldwout, 128, 16, p.n_out, p.wout_scale, tid, LB_THREADS);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));
  const uint32_t T_OP = tmem_base, T_PRE = tmem_base + 128u, T_ACC = tmem_base + 256u, T_X = tmem_base + 384u;

  const int64_t E = p.n_dev ? min((int64_t)__ldg(p.n_dev), p.E) : p.E;   // live rows
  const int64_t n_tiles = (E + RM_TILE - 1) / RM_TILE;
  const uint32_t my_tiles = (blockIdx.x < n_tiles) ? (uint32_t)((n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0u;

  const uint32_t mbits = __ldg(p.absmax);
  const int mexp = (int)((mbits >> 23) & 0xffu);
  int sexp = 127 + 2 - (mexp - 127);
  sexp = sexp < 1 ? 1 : (sexp > 254 ? 254 : sexp);
  const float gS = (mexp == 0) ? 1.f : __uint_as_float((uint32_t)sexp << 23);
  const float gSinv = (mexp == 0) ? 1.f : __uint_as_float((uint32_t)(254 - sexp) << 23);

  // ---- MMA issue: lanes 0 of a-path warps 0 and 1, in the gap where the a path waits for the data-gradient product ----
  const bool issuer = tid == 0;          // prologue + data-gradient MMAs
  const bool issuer_w = tid == 32;       // weight-gradient MMAs (independent of the former: other operands, other accumulator)
  const uint32_t idesc_d = make_idesc_f16(128, 128, 0, 1);      // A in tensor memory, B = weight image read MN-major
  const uint32_t idesc_w = make_idesc_f16(128, 128, 1, 1);      // both operands edge-major shared memory (MN-major)
  const uint32_t b_hi = base + o_w, b_lo = b_hi + w_half;
  const uint32_t a_hi_s = base + o_a, a_lo_s = a_hi_s + RM_OPT, g_hi_s = base + o_g, g_lo_s = g_hi_s + RM_OPT;
  // pre-activation gradient of the last hidden layer, (dw @ W_out^T), for tile `it` -> OP
  auto prologue = [&](uint32_t it) {
    mbar_wait(opnd0_full, it & 1u);
    tc_fence_after();
    const uint32_t wo_hi = base + o_wo, wo_lo = wo_hi + wo_half;
    const uint64_t dbh = make_desc(wo_hi, 16u * 128u, 1024), dbl = make_desc(wo_lo, 16u * 128u, 1024);
    const uint32_t a_hi = T_X, a_lo = a_hi + 16;
    mma_f16_ts(T_OP, a_lo, dbh, idesc_d, 0u);
    mma_f16_ts(T_OP, a_hi, dbl, idesc_d, 1u);
    mma_f16_ts(T_OP, a_hi, dbh, idesc_d, 1u);
    mma_commit(pro_full);
  };
  auto issue_dgrad = [&](uint32_t it) {
    mbar_wait(opnd_g, it & 1u);
    tc_fence_after();
    for (int pass = 0; pass < 2; ++pass)       // cross terms first, hi*hi last (see rmlp_fwd_kernel)
      for (int s = 0; s < 8; ++s) {
        const uint32_t a_hi = T_OP + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
        const uint64_t dbh = make_desc(b_hi + s * 2048u, 128u * 128u, 1024), dbl = make_desc(b_lo + s * 2048u, 128u * 128u, 1024);
        if (pass == 0) {
          mma_f16_ts(T_PRE, a_lo, dbh, idesc_d, s ? 1u : 0u);
          mma_f16_ts(T_PRE, a_hi, dbl, idesc_d, 1u);
        } else {
          mma_f16_ts(T_PRE, a_hi, dbh, idesc_d, 1u);
        }
      }
    mma_commit(pre_full);
  };
  // weight gradient of a tile: D[k][n] += sum_e a[e][k] g[e][n]; k-step s = edges 16 s .. 16 s + 15
  auto issue_wgrad = [&](uint32_t it) {
    mbar_wait(opnd_g, it & 1u);
    mbar_wait(opnd_a, it & 1u);
    tc_fence_after();
    const uint32_t fresh = (it % RM_LBWD_SEG) == 0 ? 0u : 1u;
    for (int pass = 0; pass < 2; ++pass)
      for (int s = 0; s < 8; ++s) {
        const uint64_t dah = make_desc(a_hi_s + s * 2048u, 16384u, 1024), dal = make_desc(a_lo_s + s * 2048u, 16384u, 1024);
        const uint64_t dgh = make_desc(g_hi_s + s * 2048u, 16384u, 1024), dgl = make_desc(g_lo_s + s * 2048u, 16384u, 1024);
        if (pass == 0) {
          mma_f16_ss(T_ACC, dal, dgh, idesc_w, s ? 1u : fresh);
          mma_f16_ss(T_ACC, dah, dgl, idesc_w, 1u);
        } else {
          mma_f16_ss(T_ACC, dah, dgh, idesc_w, 1u);
        }
      }
    mma_commit(wg_done);
  };
  if (warp < RM_CW) {
    // ============ a path (16 warps): gelu(z_{l-1}) -> A tile, then g_{l-1} = PRE * gelu'(z_{l-1}) -> global ============
    const int q = warp & 3, cg = warp >> 2;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    const uint32_t tile_off = (uint32_t)(cg >> 1) * 16384u;     // 64-column slab of an operand tile
    const uint32_t chunk0 = 4u * (uint32_t)(cg & 1);            // this thread's four 16-byte chunks of its row
    // per-CTA partial sums, [n][k]: this thread owns input index k = row (= tensor-memory lane) and outputs n = 32 cg + j,
    // so every access of a warp is one full 128-byte line
    float* part = p.partial + ((size_t)blockIdx.x * 128 + 32 * cg) * 128 + row;
    // the accumulator of the weight gradient spans RM_LBWD_SEG tiles, then moves to the partial sums with IEEE adds
    auto drain = [&](bool first_seg) {
      float v[32];
      tmem_ld32(T_ACC + lane_sel + (uint32_t)(32 * cg), v);
      tmem_ld_wait();
      if (first_seg) {
#pragma unroll
        for (int j = 0; j < 32; ++j) part[j * 128] = v[j] * gSinv;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) part[j * 128] = fmaf(v[j], gSinv, part[j * 128]);
      }
      tc_fence_before();
    };
    if (my_tiles == 0) {                      // (possible only with a device-side row count below the launch bound)
#pragma unroll
      for (int j = 0; j < 32; ++j) part[j * 128] = 0.f;
    }
    auto load_z = [&](uint32_t it, float* z) {
      const int64_t e = (blockIdx.x + (int64_t)it * gridDim.x) * RM_TILE + row;
      if (e < E) {
        load_blocked32(p.z_prev, e, 32 * cg, z);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.f;
      }
    };
    float z[32];
    if (my_tiles > 0) load_z(0, z);
    if (p.first && my_tiles > 0) {
      if (issuer) prologue(0);
      __syncwarp();
    }
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1u;
      const int64_t e = (blockIdx.x + (int64_t)it * gridDim.x) * RM_TILE + row;
      const bool live = e < E;
      float gp[32];
      uint32_t hiw[16], low[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        f32x2 a2, gp2;
        act_both2(pk2(z[2 * j], z[2 * j + 1]), ac, a2, gp2);
        float a0, a1;
        upk2(a2, a0, a1);
        upk2(gp2, gp[2 * j], gp[2 * j + 1]);
        split_h2(a0, a1, hiw[j], low[j]);
      }
      // the A tile is free once the previous tile's weight-gradient MMAs have completed
      if (it > 0) {
        mbar_wait(wg_done, (it - 1) & 1u);
        tc_fence_after();
        if ((it % RM_LBWD_SEG) == 0) drain(it == RM_LBWD_SEG);
      }
      sts_chunks(gen + o_a + tile_off, (uint32_t)row, chunk0, hiw, low);
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(opnd_a);
      // next tile's z_{l-1}: in flight across the wait below
      if (it + 1 < my_tiles) load_z(it + 1, z);
      if (issuer) {
        issue_dgrad(it);
        // the next tile's prologue: OP is free (the data-gradient MMAs above precede it in the tensor pipe)
        if (p.first && it + 1 < my_tiles) prologue(it + 1);
      }
      if (issuer_w) issue_wgrad(it);
      __syncwarp();
      mbar_wait(pre_full, par);
      tc_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float pre[16];
        tmem_ld16(T_PRE + lane_sel + (uint32_t)(32 * cg + 16 * h), pre);
        tmem_ld_wait();
        const f32x2 s2 = pk2(gSinv, gSinv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const f32x2 o2 = mul2(mul2(pk2(pre[2 * j], pre[2 * j + 1]), pk2(gp[16 * h + 2 * j], gp[16 * h + 2 * j + 1])), s2);
          upk2(o2, pre[2 * j], pre[2 * j + 1]);
        }
        if (live) emit_cols<16>(p.g_out, 1, e, 128, 32 * cg + 16 * h, pre);
      }
      tc_fence_before();
    }
    if (my_tiles > 0) {
      mbar_wait(wg_done, (my_tiles - 1) & 1u);
      tc_fence_after();
      drain(my_tiles <= RM_LBWD_SEG);
    }
  } else {
    // ============ g path (4 warps, one row per thread, four passes of 32 columns): g_l -> G tile and OP ============
    const int q = warp & 3;
    const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
    const int row = q * 32 + lane;
    const float* gsrc = p.first ? p.z_l : p.g_in;
    auto publish_dw = [&](const float* d) {  // a dw row (scaled here) -> A operand of the prologue MMA (T_X)
      uint32_t hiw[8], low[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) split_h2(d[2 * j] * gS, d[2 * j + 1] * gS, hiw[j], low[j]);
      tmem_st8(T_X + lane_sel, hiw);
      tmem_st8(T_X + lane_sel + 16, low);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(opnd0_full);
    };
    // The rows of this warp are one 32-edge chunk of the blocked layout; a pass (32 columns) of it is 4 KB contiguous.
    // One bulk copy per pass lands in this warp's 4 KB buffer; the copy of the next pass is issued as soon as the
    // current one sits in registers, so its latency is covered by the pass's arithmetic.
    const uint32_t land = o_land + (uint32_t)q * NST * 4096u;
    auto copy_pass = [&](uint32_t n) {           // n = 4 * it + c, stage n % NST
      if (lane == 0 && n < 4u * my_tiles) {
        const int64_t chunk = (blockIdx.x + (int64_t)(n >> 2) * gridDim.x) * 4 + q;
        const uint32_t st = n % NST;
        mbar_expect_tx(land_full(q, st), 4096u);
        bulk_g2s(base + land + st * 4096u, gsrc + chunk * (int64_t)(32 * 128) + (int64_t)(n & 3u) * 1024, 4096u, land_full(q, st));
      }
    };
    for (uint32_t n = 0; n < NST; ++n) copy_pass(n);
    float dnext[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) dnext[c] = 0.f;
    auto load_dw = [&](uint32_t it) {
      const int64_t e = (blockIdx.x + (int64_t)it * gridDim.x) * RM_TILE + row;
#pragma unroll
      for (int c = 0; c < 16; ++c)
        dnext[c] = (it < my_tiles && e < E && c < p.n_out) ? __ldg(p.dw + (int64_t)c * p.ld_dw + e) : 0.f;
    };
    if (p.first && my_tiles > 0) {
      load_dw(0);
      publish_dw(dnext);
    }
    for (uint32_t it = 0; it < my_tiles; ++it) {
      const uint32_t par = it & 1u;
      const int64_t e = (blockIdx.x + (int64_t)it * gridDim.x) * RM_TILE + row;
      const bool live = e < E;
      if (p.first) load_dw(it + 1);            // the next tile's dw row: in flight across this tile's passes
      if (p.first) {
        mbar_wait(pro_full, par);
        tc_fence_after();
      }
      // the G tile and OP are free once the previous tile's MMAs (weight gradient, data gradient) have completed
      if (it > 0) {
        mbar_wait(wg_done, (it - 1) & 1u);
        mbar_wait(pre_full, (it - 1) & 1u);
        tc_fence_after();
      }
#pragma unroll 1
      for (int c = 0; c < 4; ++c) {
        const uint32_t n = 4u * it + (uint32_t)c;
        float v[32];
        mbar_wait(land_full(q, n % NST), (n / NST) & 1u);
        const float* lbuf = reinterpret_cast<const float*>(gen + land + (n % NST) * 4096u) + lane * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 t4 = *reinterpret_cast<const float4*>(lbuf + i * 128);
          v[4 * i] = live ? t4.x : 0.f; v[4 * i + 1] = live ? t4.y : 0.f; v[4 * i + 2] = live ? t4.z : 0.f; v[4 * i + 3] = live ? t4.w : 0.f;
        }
        __syncwarp();
        copy_pass(n + NST);
        uint32_t hiw[16], low[16];
        if (p.first) {
          float pg[32];
          tmem_ld32(T_OP + lane_sel + (uint32_t)(32 * c), pg);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const f32x2 g2 = mul2(pk2(pg[2 * j], pg[2 * j + 1]), pk2(v[2 * j], v[2 * j + 1]));     // v = gelu'(z_H)
            float g0, g1;
            upk2(g2, g0, g1);
            split_h2(g0, g1, hiw[j], low[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 16; ++j) split_h2(v[2 * j] * gS, v[2 * j + 1] * gS, hiw[j], low[j]);
        }
        sts_chunks(gen + o_g + (uint32_t)(c >> 1) * 16384u, (uint32_t)row, 4u * (uint32_t)(c & 1), hiw, low);
        const uint32_t t = T_OP + lane_sel + (uint32_t)(32 * c);
        tmem_st16(t, hiw);
        tmem_st16(t + 16, low);
      }
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(opnd_g);
      if (p.first && it + 1 < my_tiles) publish_dw(dnext);   // T_X is free: this tile's prologue completed before its passes
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == LB_GW0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// dW_l = alpha * sum over the per-CTA partials of rmlp_lbwd_kernel, in CTA order (deterministic): 32 outputs x 8
// part-groups per block
__global__ void __launch_bounds__(256) lbwd_reduce_kernel(const float* __restrict__ partial, int n_parts, float alpha,
                                                          float* __restrict__ out, int64_t ldo) {
  __shared__ float red[8][33];
  const int lane = threadIdx.x & 31, grp = threadIdx.x >> 5;
  const int i = blockIdx.x * 32 + lane;                 // < 128 * 128
  const int c = i >> 7, r = i & 127;                    // partials are stored [n = c][k = r]
  float s = 0.f;
  for (int q = grp; q < n_parts; q += 8) s += partial[((size_t)q * 128 + c) * 128 + r];
  red[grp][lane] = s;
  __syncthreads();
  if (grp == 0) {
    float t = red[0][lane];
#pragma unroll
    for (int g = 1; g < 8; ++g) t += red[g][lane];
    out[(int64_t)r * ldo + c] = alpha * t;
  }
}

// blocked [chunk][c4][row][4] -> row-major [E][128] (small problems only)
__global__ void unblock_kernel(const float* __restrict__ src, int64_t n_edges, float* __restrict__ dst) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // float4 index of the row-major matrix
  if (i >= n_edges * 32) return;
  const int64_t e = i >> 5;
  const int c4 = (int)(i & 31);
  reinterpret_cast<float4*>(dst)[i] =
      __ldg(reinterpret_cast<const float4*>(src + (e >> 5) * (int64_t)(32 * 128) + (int64_t)c4 * 128 + (e & 31) * 4));
}

}  // namespace

// Shapes the fused kernels cover: 128-wide hidden layers, 1..3 of them, a Bessel basis of 16, 32 or 64 functions,
// at most 16 live output columns.  Anything else: the caller keeps the per-layer GEMMs (eqnn_gemm_f32).
extern "C" int eqnn_radial_mlp_supported(int n_rb, int d_hidden, int n_hidden, int n_out) {
  return (n_rb == 16 || n_rb == 32 || n_rb == 64) && d_hidden == 128 && n_hidden >= 1 && n_hidden <= 3 && n_out >= 1 &&
         n_out <= 16;
}

// one layer of the z stash: whole 128-edge tiles of 128 floats
static size_t rm_stash_layer_floats(int64_t n_edges) { return align_up((size_t)n_edges, 128) * 128; }
extern "C" size_t eqnn_radial_mlp_stash_bytes(int64_t n_edges, int n_hidden) {
  return (size_t)(n_hidden + 1) * rm_stash_layer_floats(n_edges) * sizeof(float);   // z_1 .. z_{H-1}, gelu(z_H), gelu'(z_H)
}

static int rm_fwd_impl(const float* dist, int64_t dist_stride, int64_t n_edges, const int32_t* n_dev, int n_rb, double r_cut,
                       int n_hidden, int d_hidden, const float* const* weights, const int64_t* ldw, int n_out, int activation,
                       float act_scale, float* w_t, int64_t ld_wt, float* w_rows, int ld_rows, float* z_stash,
                       eqnn_stream_t stream) {
  EQNN_CHECK_ARG(activation == EQNN_ACT_GELU || activation == EQNN_ACT_SILU, "radial_mlp_fwd: unknown activation");
  EQNN_CHECK_ARG(dist && weights && ldw && (w_t || w_rows) && n_edges >= 0 && dist_stride >= 1, "radial_mlp_fwd: bad argument");
  EQNN_CHECK_ARG(eqnn_radial_mlp_supported(n_rb, d_hidden, n_hidden, n_out), "radial_mlp_fwd: unsupported shape");
  EQNN_CHECK_ARG(r_cut > 0.0 && (!w_t || ld_wt >= n_edges), "radial_mlp_fwd: bad r_cut / ld_wt");
  EQNN_CHECK_ARG(!w_rows || (ld_rows % 4 == 0 && ld_rows >= n_out && ld_rows <= 16 && (reinterpret_cast<uintptr_t>(w_rows) & 15) == 0),
                 "radial_mlp_fwd: w_rows needs ld_rows = n_out rounded up to 4 and 16-byte alignment");
  // bessel values are bounded by sqrt(2/c) * n_rb * pi / c: they must stay inside the fp16 range of the operand split
  EQNN_CHECK_ARG(sqrt(2.0 / r_cut) * n_rb * 3.14159265358979323846 / r_cut < 60000.0, "radial_mlp_fwd: r_cut too small for the fp16 operand split");
  if (n_edges == 0) return EQNN_OK;
  RmlpP p;
  p.dist = dist; p.dist_stride = dist_stride; p.E = n_edges;
  p.r_cut = (float)r_cut; p.pref = (float)sqrt(2.0 / r_cut);   // jnp.sqrt(2.0 / x_max): a Python double folded to fp32
  p.n_rb = n_rb; p.n_layers = n_hidden + 1; p.n_out = n_out;
  for (int l = 0; l < RM_MAXL; ++l) {
    p.W[l] = nullptr; p.ldw[l] = 0; p.alpha[l] = 0.f;
  }
  for (int l = 0; l < p.n_layers; ++l) {
    EQNN_CHECK_ARG(weights[l] != nullptr, "radial_mlp_fwd: null weight");
    p.W[l] = weights[l]; p.ldw[l] = ldw[l];
    p.alpha[l] = (float)(1.0 / sqrt((double)rm_K(p, l)));
  }
  p.act_scale = act_scale; p.act_kind = activation; p.w_t = w_t; p.ld_wt = ld_wt;
  p.n_dev = n_dev; p.w_rows = w_rows; p.ld_rows = ld_rows;
  EQNN_CHECK_ARG(!z_stash || (reinterpret_cast<uintptr_t>(z_stash) & 15) == 0, "radial_mlp_fwd: z_stash must be 16-byte aligned");
  for (int l = 0; l < RM_MAXL; ++l)
    p.z_stash[l] = (z_stash && l <= n_hidden) ? z_stash + (size_t)l * rm_stash_layer_floats(n_edges) : nullptr;
  const uint32_t smem = rm_smem_bytes(p);
  static DeviceOnce once;
  EQNN_CUDA(configure_smem(once, rmlp_fwd_kernel, 227 * 1024));
  const int64_t n_tiles = (n_edges + RM_TILE - 1) / RM_TILE;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  rmlp_fwd_kernel<<<grid, RM_THREADS, smem, as_stream(stream)>>>(p);
  EQNN_CHECK_LAUNCH();
  return EQNN_OK;
}

extern "C" int eqnn_radial_mlp_fwd(const float* dist, int64_t dist_stride, int64_t n_edges, int n_rb, double r_cut,
                                   int n_hidden, int d_hidden, const float* const* weights, const int64_t* ldw,
                                   int n_out, int activation, float act_scale, float* w_t, int64_t ld_wt, float* z_stash,
                                   eqnn_stream_t stream) {
  EQNN_CHECK_ARG(w_t, "radial_mlp_fwd: bad argument");
  return rm_fwd_impl(dist, dist_stride, n_edges, nullptr, n_rb, r_cut, n_hidden, d_hidden, weights, ldw, n_out, activation,
                     act_scale, w_t, ld_wt, nullptr, 0, z_stash, stream);
}

// Same MLP over a compact array of rows whose live count sits in device memory (shared radial evaluations, pairs.cu):
// n_cap sizes every buffer and the launch, *n_rows (<= n_cap) rows are evaluated; the result is written as row-major
// records w_rows [n_cap][ld_rows] (ld_rows = n_out rounded up to 4) and / or column-major w_t.
extern "C" int eqnn_radial_mlp_fwd_rows(const float* dist, int64_t dist_stride, int64_t n_cap, const int32_t* n_rows, int n_rb,
                                        double r_cut, int n_hidden, int d_hidden, const float* const* weights,
                                        const int64_t* ldw, int n_out, int activation, float act_scale, float* w_t,
                                        int64_t ld_wt, float* w_rows, int ld_rows, float* z_stash, eqnn_stream_t stream) {
  return rm_fwd_impl(dist, dist_stride, n_cap, n_rows, n_rb, r_cut, n_hidden, d_hidden, weights, ldw, n_out, activation,
                     act_scale, w_t, ld_wt, w_rows, ld_rows, z_stash, stream);
}

int eqnn_tc_wgrad_try(int64_t M, int64_t N, int64_t K, float alpha, const float* A, int64_t sam, int64_t sak,
                      const float* B, int64_t sbk, int64_t sbn, float* C, int64_t scm, int64_t scn, int accumulate,
                      int pro, float pro_scale, int epi, void* ws, size_t ws_bytes, cudaStream_t st,
                      float* colsum_out, int blocked, const int32_t* n_dev);
int eqnn_tc_wgrad_bessel_try(int n_rb, int64_t E, float alpha, const float* geom, double r_cut, const float* G, int g_blocked,
                             float* C, int64_t scm, void* ws, size_t ws_bytes, cudaStream_t st, const int32_t* n_dev);

// workspace: [absmax 256 B | bessel basis E*n_rb | n_hidden x gelu(z_l) E*128 | n_hidden x g_l E*128 | GEMM workspace]
static size_t rm_bwd_gemm_ws(int64_t n_edges) {
  size_t a = eqnn_gemm_workspace_bytes(128, 128, n_edges), b = eqnn_gemm_workspace_bytes(64, 128, n_edges);
  size_t c = eqnn_gemm_workspace_bytes(128, 16, n_edges);
  a = a > b ? a : b;
  return a > c ? a : c;
}
extern "C" size_t eqnn_radial_mlp_bwd_workspace_bytes(int64_t n_edges, int n_rb, int n_hidden) {
  const size_t ep = align_up((size_t)n_edges, 128);   // whole 32-edge chunks (blocked layout)
  return 256 + ep * n_rb * 4 + 2 * (size_t)n_hidden * ep * 512 + align_up(rm_bwd_gemm_ws(n_edges), 256);
}


// Backward from the forward kernel's z stash: absmax -> data-gradient chain kernel -> weight gradients on the tcgen05
// weight-gradient kernel (X = z_{l-1} with the gelu prologue, or the Bessel basis for the first layer; G = g_l).
// workspace: [absmax 256 B | n_hidden x g_l | (bessel basis when radial == NULL) | GEMM workspace]
static int rm_bwd_from_stash(const float* dist, int64_t dist_stride, int64_t n_edges, const int32_t* n_dev, int n_rb, double r_cut, int n_hidden,
                             int d_hidden, const float* const* weights, const int64_t* ldw, int n_out, int activation,
                             float act_scale, const float* dw_t, int64_t ld_dwt, const float* z_stash, const float* radial,
                             float* const* grad_weights, void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  // the weight-gradient kernel's prologue (tc_mlp.cu) evaluates gelu on the stashed pre-activations: other activations
  // take the recompute variant, which hands it bare activations
  EQNN_CHECK_ARG(activation == EQNN_ACT_GELU, "radial_mlp_bwd: the z-stash variant supports gelu only (pass z_stash = NULL)");
  EQNN_CHECK_ARG(dist && weights && ldw && dw_t && grad_weights && n_edges >= 0 && dist_stride >= 1, "radial_mlp_bwd: bad argument");
  EQNN_CHECK_ARG(eqnn_radial_mlp_supported(n_rb, d_hidden, n_hidden, n_out), "radial_mlp_bwd: unsupported shape");
  EQNN_CHECK_ARG(r_cut > 0.0 && ld_dwt >= n_edges, "radial_mlp_bwd: bad r_cut / ld_dwt");
  // The first layer's weight gradient can generate the Bessel basis itself (tc_wgrad_kernel PRO = 2) when the edge lengths
  // sit in (r_hat, d) records (dist = geom + 3, stride 4): taken when no radial array is given, or on request
  // (EQNN_RMLP_BESSEL_WGRAD=1).  Measured at 3.2 M edges: 0.68 ms against 0.52 ms reading eqnn_edge_geom's radial -- the
  // transform warps pay 64 sines per edge four times per training step, more than the 256 B / edge they save -- so
  // callers that hold the basis anyway (NequIP: one array per batch) pass it.
  const bool bessel_ok = dist_stride == 4 && (n_rb == 32 || n_rb == 64) && n_edges >= 4096 &&
                         (reinterpret_cast<uintptr_t>(dist - 3) & 15) == 0 && !getenv("EQNN_RMLP_NO_TC_WGRAD");
  const bool bessel_wgrad = bessel_ok && (radial == nullptr || getenv("EQNN_RMLP_BESSEL_WGRAD"));
  EQNN_CHECK_ARG(radial != nullptr || bessel_wgrad || n_edges == 0,
                 "radial_mlp_bwd: this shape needs the Bessel basis (eqnn_edge_geom's radial) for the first layer's weight gradient");
  EQNN_CHECK_ARG((reinterpret_cast<uintptr_t>(z_stash) & 15) == 0, "radial_mlp_bwd: z_stash must be 16-byte aligned");
  if (ws_bytes < eqnn_radial_mlp_bwd_workspace_bytes(n_edges, n_rb, n_hidden) || !ws) return eqnn_fail(EQNN_ERR_WORKSPACE, "radial_mlp_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int F = n_hidden + 1;
  RmlpDgradP p;
  RmlpP& f = p.f;
  f.dist = dist; f.dist_stride = dist_stride; f.E = n_edges;
  f.r_cut = (float)r_cut; f.pref = (float)sqrt(2.0 / r_cut);
  f.n_rb = n_rb; f.n_layers = F; f.n_out = n_out;
  for (int l = 0; l < RM_MAXL; ++l) {
    f.W[l] = nullptr; f.ldw[l] = 0; f.alpha[l] = 0.f; f.z_stash[l] = nullptr;
    p.z[l] = nullptr; p.grad[l] = nullptr; p.blk[l] = 0;
  }
  for (int l = 0; l < F; ++l) {
    EQNN_CHECK_ARG(weights[l] != nullptr && grad_weights[l] != nullptr, "radial_mlp_bwd: null weight");
    f.W[l] = weights[l]; f.ldw[l] = ldw[l];
    f.alpha[l] = (float)(1.0 / sqrt((double)rm_K(f, l)));
  }
  f.act_scale = act_scale; f.act_kind = activation; f.w_t = nullptr; f.ld_wt = 0;
  f.n_dev = nullptr; f.w_rows = nullptr; f.ld_rows = 0;
  p.dw = dw_t; p.ld_dw = ld_dwt;
  char* w = reinterpret_cast<char*>(ws);
  uint32_t* absmax = reinterpret_cast<uint32_t*>(w);
  w += 256;
  p.absmax = absmax;
  const size_t ep = align_up((size_t)n_edges, 128);
  // stash slots (eqnn_radial_mlp_fwd): z_1 .. z_{H-1}, gelu(z_H), gelu'(z_H)
  const float* a_last = z_stash + (size_t)(n_hidden - 1) * rm_stash_layer_floats(n_edges);
  for (int l = 1; l <= n_hidden; ++l) {
    p.z[l] = z_stash + (size_t)(l < n_hidden ? l - 1 : n_hidden) * rm_stash_layer_floats(n_edges);   // z[H] = gelu'(z_H)
    p.grad[l] = reinterpret_cast<float*>(w);
    w += ep * 512;
  }
  w += ep * n_rb * 4 + (size_t)n_hidden * ep * 512;   // (the recompute variant's activations live here)
  void* gws = w;
  const size_t gws_bytes = rm_bwd_gemm_ws(n_edges);
  if (n_edges == 0) {
    for (int l = 0; l < F; ++l) {
      const int K = rm_K(f, l), N = l == F - 1 ? n_out : 128;
      for (int k = 0; k < K; ++k) EQNN_CUDA(cudaMemsetAsync(grad_weights[l] + (int64_t)k * ldw[l], 0, (size_t)N * 4, st));
    }
    return EQNN_OK;
  }
  // layer l's weight gradient (0-based): X = radial (l = 0, row-major) or z_l (blocked), G = grad[l + 1] or dw_t
  bool tc_ok[RM_MAXL];
  for (int l = 0; l < F; ++l) {
    const int K = rm_K(f, l);
    bool ok = n_edges >= 4096 && (K == 32 || K == 64 || K == 128) && !getenv("EQNN_RMLP_NO_TC_WGRAD");
    if (l == 0) ok = ok && (bessel_wgrad || (radial && (reinterpret_cast<uintptr_t>(radial) & 15) == 0));
    if (l == F - 1) ok = ok && ld_dwt % 4 == 0 && (reinterpret_cast<uintptr_t>(dw_t) & 15) == 0;
    tc_ok[l] = ok;
    if (l < F - 1) p.blk[l] = ok ? 1 : 0;     // layout of grad[l + 1]
  }
  EQNN_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
  {
    const int64_t want = (n_edges + 255) / 256;
    const unsigned grid = (unsigned)(want < 1184 ? want : 1184);
    absmax_kernel<<<grid, 256, 0, st>>>(dw_t, ld_dwt, n_out, n_edges, absmax, n_dev);
  }
  const int64_t n_tiles = (n_edges + RM_TILE - 1) / RM_TILE;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  // Hidden layers 2..H: one fused kernel per layer (data gradient + weight gradient in one pass over the edges, 40 %
  // fewer bytes than the chain kernel + separate weight gradients: 3.8 vs 4.4 ms per 3.2 M edges).  EQNN_RMLP_NO_LBWD=1
  // keeps the chain kernel (also taken when a weight gradient cannot use the blocked tensor-core route).
  bool fused = n_hidden >= 2 && !getenv("EQNN_RMLP_NO_LBWD") && gws_bytes >= (size_t)grid * 128 * 128 * sizeof(float);
  for (int l = 0; l < F; ++l) fused = fused && tc_ok[l];
  if (fused) {
    static DeviceOnce once;
    EQNN_CUDA(configure_smem(once, rmlp_lbwd_kernel, 227 * 1024));
    EQNN_CHECK_LAUNCH();   // (absmax)
    for (int l = n_hidden; l >= 2; --l) {                  // 1-based hidden layer: W[l - 1] maps gelu(z_{l-1}) to z_l
      RmlpLbwdP q;
      q.E = n_edges; q.n_dev = n_dev;
      q.W = weights[l - 1]; q.ldw = ldw[l - 1]; q.wscale = rm_wscale(f.alpha, act_scale, l - 1);
      q.Wout = weights[F - 1]; q.ldwout = ldw[F - 1]; q.wout_scale = rm_wscale(f.alpha, act_scale, F - 1); q.n_out = n_out;
      q.first = (l == n_hidden) ? 1 : 0;
      q.dw = dw_t; q.ld_dw = ld_dwt; q.absmax = absmax;
      q.z_l = p.z[l]; q.g_in = p.grad[l]; q.z_prev = p.z[l - 1]; q.g_out = p.grad[l - 1];
      q.partial = reinterpret_cast<float*>(gws);
      q.act_kind = activation;
      const uint32_t smem_l = 2 * w_half_bytes(128, 128) + 4 * RM_OPT + (q.first ? 2 * w_half_bytes(128, 16) + 4 * 4096 : 8 * 4096) + 1024 + 128;
      rmlp_lbwd_kernel<<<grid, LB_THREADS, smem_l, st>>>(q);
      EQNN_CHECK_LAUNCH();
      lbwd_reduce_kernel<<<128 * 128 / 32, 256, 0, st>>>(q.partial, (int)grid, q.wscale, grad_weights[l - 1], ldw[l - 1]);
      EQNN_CHECK_LAUNCH();
    }
  } else {
    if (n_dev) return eqnn_fail(EQNN_ERR_ARG, "radial_mlp_bwd_rows: a device-side row count needs the fused layer kernels (n_hidden >= 2, n_cap >= 4096, blocked tensor-core weight gradients)");
    uint32_t smem = 1024 + 64;
    for (int l = 1; l < F; ++l) smem += 2 * w_half_bytes(rm_K(f, l), rm_N(f, l));
    static DeviceOnce once;
    EQNN_CUDA(configure_smem(once, rmlp_dgrad_kernel, 227 * 1024));
    rmlp_dgrad_kernel<<<grid, RM_THREADS, smem, st>>>(p);
    EQNN_CHECK_LAUNCH_N(2);
  }
  for (int l = 0; l < F; ++l) {
    if (fused && l >= 1 && l < F - 1) continue;           // done by the fused layer kernels
    const int K = rm_K(f, l);
    // X = input of layer l: the Bessel basis, z_l (gelu applied by the kernel's prologue), or the stashed gelu(z_H)
    const float* X = l == 0 ? radial : (l == F - 1 ? a_last : p.z[l]);
    const int pro = (l == 0 || l == F - 1) ? 0 : 1;
    const float al = l == F - 1 ? f.alpha[l] * act_scale : f.alpha[l];
    const float* G = l < F - 1 ? p.grad[l + 1] : dw_t;
    const int64_t N = l < F - 1 ? 128 : n_out, sbk = l < F - 1 ? 128 : 1, sbn = l < F - 1 ? 1 : ld_dwt;
    if (tc_ok[l] && l == 0 && bessel_wgrad && F >= 2) {
      const int r = eqnn_tc_wgrad_bessel_try(n_rb, n_edges, f.alpha[0], dist - 3, r_cut, G, 1, grad_weights[0], ldw[0], gws, gws_bytes, st, n_dev);
      if (r < 0) return -r;
      if (r == 1) continue;
      if (!radial) return eqnn_fail(EQNN_ERR_ARG, "radial_mlp_bwd: first-layer weight gradient not covered and no radial given");
    }
    if (tc_ok[l]) {
      const int blocked = (l > 0 ? 1 : 0) | (l < F - 1 ? 2 : 0);
      const int r = eqnn_tc_wgrad_try(K, N, n_edges, al, X, 1, K, G, sbk, sbn, grad_weights[l], ldw[l], 1, 0, pro,
                                      act_scale, 0, gws, gws_bytes, st, nullptr, blocked, n_dev);
      if (r < 0) return -r;
      if (r != 1) return eqnn_fail(EQNN_ERR_ARG, "radial_mlp_bwd: weight-gradient shape not covered by the tensor-core kernel");
      continue;
    }
    // small problems: CUDA-core GEMM on row-major operands; a blocked z_l is first copied out row-major
    if (n_dev) return eqnn_fail(EQNN_ERR_ARG, "radial_mlp_bwd_rows: weight-gradient shape not covered by the tensor-core kernel");
    const float* Xr = X;
    if (l > 0) {
      float* tmp = reinterpret_cast<float*>(reinterpret_cast<char*>(ws) + 256 + (size_t)n_hidden * ep * 512);
      unblock_kernel<<<(unsigned)ceil_div64(n_edges * 32, 256), 256, 0, st>>>(X, n_edges, tmp);
      EQNN_CHECK_LAUNCH();
      Xr = tmp;
    }
    const int rc = eqnn_gemm_f32(K, N, n_edges, al, Xr, 1, K, G, sbk, sbn, grad_weights[l], ldw[l], 1, 0, pro,
                                 act_scale, 0, nullptr, 0, 1.f, gws, gws_bytes, stream);
    if (rc) return rc;
  }
  return EQNN_OK;
}

// Backward over a compact row array with a device-side live count (see eqnn_radial_mlp_fwd_rows): the z-stash variant on
// the fused layer kernels only.
extern "C" int eqnn_radial_mlp_bwd_rows(const float* dist, int64_t dist_stride, int64_t n_cap, const int32_t* n_rows, int n_rb,
                                        double r_cut, int n_hidden, int d_hidden, const float* const* weights,
                                        const int64_t* ldw, int n_out, int activation, float act_scale, const float* dw_t,
                                        int64_t ld_dwt, const float* z_stash, const float* radial, float* const* grad_weights,
                                        void* ws, size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(z_stash && radial, "radial_mlp_bwd_rows: needs the forward kernel's z stash and the Bessel basis of the rows");
  return rm_bwd_from_stash(dist, dist_stride, n_cap, n_rows, n_rb, r_cut, n_hidden, d_hidden, weights, ldw, n_out, activation,
                           act_scale, dw_t, ld_dwt, z_stash, radial, grad_weights, ws, ws_bytes, stream);
}

extern "C" int eqnn_radial_mlp_bwd(const float* dist, int64_t dist_stride, int64_t n_edges, int n_rb, double r_cut,
                                   int n_hidden, int d_hidden, const float* const* weights, const int64_t* ldw,
                                   int n_out, int activation, float act_scale, const float* dw_t, int64_t ld_dwt,
                                   const float* z_stash, const float* radial, float* const* grad_weights, void* ws,
                                   size_t ws_bytes, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(activation == EQNN_ACT_GELU || activation == EQNN_ACT_SILU, "radial_mlp_bwd: unknown activation");
  if (z_stash)
    return rm_bwd_from_stash(dist, dist_stride, n_edges, nullptr, n_rb, r_cut, n_hidden, d_hidden, weights, ldw, n_out, activation,
                             act_scale, dw_t, ld_dwt, z_stash, radial, grad_weights, ws, ws_bytes, stream);
  EQNN_CHECK_ARG(dist && weights && ldw && dw_t && grad_weights && n_edges >= 0 && dist_stride >= 1, "radial_mlp_bwd: bad argument");
  EQNN_CHECK_ARG(eqnn_radial_mlp_supported(n_rb, d_hidden, n_hidden, n_out), "radial_mlp_bwd: unsupported shape");
  EQNN_CHECK_ARG(r_cut > 0.0 && ld_dwt >= n_edges, "radial_mlp_bwd: bad r_cut / ld_dwt");
  EQNN_CHECK_ARG(sqrt(2.0 / r_cut) * n_rb * 3.14159265358979323846 / r_cut < 60000.0, "radial_mlp_bwd: r_cut too small for the fp16 operand split");
  if (ws_bytes < eqnn_radial_mlp_bwd_workspace_bytes(n_edges, n_rb, n_hidden) || !ws) return eqnn_fail(EQNN_ERR_WORKSPACE, "radial_mlp_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const int F = n_hidden + 1;
  RmlpBwdP p;
  RmlpP& f = p.f;
  f.dist = dist; f.dist_stride = dist_stride; f.E = n_edges;
  f.r_cut = (float)r_cut; f.pref = (float)sqrt(2.0 / r_cut);
  f.n_rb = n_rb; f.n_layers = F; f.n_out = n_out;
  for (int l = 0; l < RM_MAXL; ++l) {
    f.W[l] = nullptr; f.ldw[l] = 0; f.alpha[l] = 0.f;
    p.act[l] = nullptr; p.grad[l] = nullptr;
  }
  for (int l = 0; l < F; ++l) {
    EQNN_CHECK_ARG(weights[l] != nullptr && grad_weights[l] != nullptr, "radial_mlp_bwd: null weight");
    f.W[l] = weights[l]; f.ldw[l] = ldw[l];
    f.alpha[l] = (float)(1.0 / sqrt((double)rm_K(f, l)));
  }
  f.act_scale = act_scale; f.act_kind = activation; f.w_t = nullptr; f.ld_wt = 0;
  f.n_dev = nullptr; f.w_rows = nullptr; f.ld_rows = 0;
  for (int l = 0; l < RM_MAXL; ++l) f.z_stash[l] = nullptr;
  p.dw = dw_t; p.ld_dw = ld_dwt;
  char* w = reinterpret_cast<char*>(ws);
  uint32_t* absmax = reinterpret_cast<uint32_t*>(w);
  w += 256;
  p.absmax = absmax;
  const size_t ep = align_up((size_t)n_edges, 128);
  p.act[0] = reinterpret_cast<float*>(w);
  w += ep * n_rb * 4;
  for (int l = 1; l <= n_hidden; ++l) {
    p.act[l] = reinterpret_cast<float*>(w);
    w += ep * 512;
  }
  for (int l = 1; l <= n_hidden; ++l) {
    p.grad[l] = reinterpret_cast<float*>(w);
    w += ep * 512;
  }
  void* gws = w;
  const size_t gws_bytes = rm_bwd_gemm_ws(n_edges);
  // the tensor-core weight-gradient kernel (tc_mlp.cu) takes layers with 32 / 64 / 128 inputs from 4096 edges on; the
  // output layer reads dw_t column-major and needs 16-byte aligned columns.  Other layers: row-major + CUDA-core GEMM.
  for (int l = 0; l < RM_MAXL; ++l) p.blk[l] = 0;
  for (int l = 0; l < F; ++l) {
    const int K = rm_K(f, l);
    bool ok = n_edges >= 4096 && (K == 32 || K == 64 || K == 128);
    if (l == F - 1) ok = ok && ld_dwt % 4 == 0 && (reinterpret_cast<uintptr_t>(dw_t) & 15) == 0;
    p.blk[l] = ok ? 1 : 0;
  }
  if (n_edges == 0) {
    for (int l = 0; l < F; ++l) {
      const int K = rm_K(f, l), N = l == F - 1 ? n_out : 128;
      for (int k = 0; k < K; ++k) EQNN_CUDA(cudaMemsetAsync(grad_weights[l] + (int64_t)k * ldw[l], 0, (size_t)N * 4, st));
    }
    return EQNN_OK;
  }
  EQNN_CUDA(cudaMemsetAsync(absmax, 0, 4, st));
  {
    const int64_t want = (n_edges + 255) / 256;
    const unsigned grid = (unsigned)(want < 1184 ? want : 1184);
    absmax_kernel<<<grid, 256, 0, st>>>(dw_t, ld_dwt, n_out, n_edges, absmax, nullptr);
  }
  uint32_t smem = rm_smem_bytes(f) + 2 * RM_GSLAB;
  static DeviceOnce once;
  EQNN_CUDA(configure_smem(once, rmlp_bwd_kernel, 227 * 1024));
  const int64_t n_tiles = (n_edges + RM_TILE - 1) / RM_TILE;
  const int sms = sm_count();
  const unsigned grid = (unsigned)(n_tiles < sms ? n_tiles : sms);
  rmlp_bwd_kernel<<<grid, RM_THREADS, smem, st>>>(p);
  EQNN_CHECK_LAUNCH_N(2);
  // weight gradients: dW_l = alpha_l * c_l * gelu(z_{l-1})^T g_l (c_l = act_scale for l >= 1: the emitted activations are bare)
  for (int l = 0; l < F; ++l) {
    const int K = rm_K(f, l);
    const float al = rm_wscale(f.alpha, f.act_scale, l);
    int rc = 0;
    const float* G = l < F - 1 ? p.grad[l + 1] : dw_t;
    const int64_t N = l < F - 1 ? 128 : n_out, sbk = l < F - 1 ? 128 : 1, sbn = l < F - 1 ? 1 : ld_dwt;
    if (p.blk[l]) {
      const int w = eqnn_tc_wgrad_try(K, N, n_edges, al, p.act[l], 1, K, G, sbk, sbn, grad_weights[l], ldw[l], 1, 0, 0, 1.f,
                                      0, gws, gws_bytes, st, nullptr, l < F - 1 ? 3 : 1, nullptr);
      if (w < 0) return -w;
      if (w != 1) return eqnn_fail(EQNN_ERR_ARG, "radial_mlp_bwd: weight-gradient shape not covered by the tensor-core kernel");
      continue;
    }
    rc = eqnn_gemm_f32(K, N, n_edges, al, p.act[l], 1, K, G, sbk, sbn, grad_weights[l], ldw[l], 1, 0, 0, 1.f, 0, nullptr, 0,
                       1.f, gws, gws_bytes, stream);
    if (rc) return rc;
  }
  return EQNN_OK;
}
