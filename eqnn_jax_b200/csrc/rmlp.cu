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
  float act_scale;          // 1 / c of e3nn.normalize_function(gelu)
  float* w_t;               // [n_out][ld_wt] column-major radial weights
  int64_t ld_wt;
};

__host__ __device__ inline int rm_K(const RmlpP& p, int l) { return l == 0 ? p.n_rb : 128; }
__host__ __device__ inline int rm_N(const RmlpP& p, int l) { return l == p.n_layers - 1 ? 16 : 128; }

__host__ inline uint32_t rm_smem_bytes(const RmlpP& p) {
  uint32_t b = 0;
  for (int l = 0; l < p.n_layers; ++l) b += 2 * w_half_bytes(rm_K(p, l), rm_N(p, l));
  return b + 1024 /*align*/ + 256 /*bessel frequencies*/ + 64 /*barriers*/;
}

// alpha * acc -> gelu * act_scale -> (hi, lo) words, 32 accumulator columns of one row
__device__ __forceinline__ void act_split32(const float* v, float alpha, float act_scale, uint32_t* hiw, uint32_t* low) {
  const f32x2 al2 = pk2(alpha, alpha), sc2 = pk2(act_scale, act_scale);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const f32x2 a = gelu2(mul2(pk2(v[2 * j], v[2 * j + 1]), al2), sc2);
    float a0, a1;
    upk2(a, a0, a1);
    split_h2(a0, a1, hiw[j], low[j]);
  }
}

__global__ void __launch_bounds__(RM_THREADS, 1) rmlp_fwd_kernel(const RmlpP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* gen = smem_raw + (base - smem_u32(smem_raw));
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int F = p.n_layers;

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
                      l == F - 1 ? p.n_out : 128, 1.0f, tid, RM_THREADS);
  if (tid < p.n_rb) wfreq[tid] = __fdiv_rn(__fmul_rn((float)(tid + 1), 3.14159265358979323846f), p.r_cut);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(gen + (tmem_slot - base));

  const int64_t n_tiles = (p.E + RM_TILE - 1) / RM_TILE;
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
          const float d = (e < p.E) ? __ldg(p.dist + e * p.dist_stride) : 0.f;
          const float inv_d = (d == 0.f) ? 0.f : __fdiv_rn(1.0f, d);
          uint32_t hiw[8], low[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            float v[2];
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float w = wfreq[16 * cg + 2 * j + u];
              const float s = sin_reduced(__fmul_rn(w, d)) * inv_d;
              v[u] = __fmul_rn(p.pref, (d == 0.f) ? w : s);
            }
            split_h2(v[0], v[1], hiw[j], low[j]);
          }
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
          act_split32(v, p.alpha[l - 1], p.act_scale, hiw, low);
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
          if (e < p.E) {
#pragma unroll
            for (int c = 0; c < 16; ++c)
              if (c < p.n_out) p.w_t[(int64_t)c * p.ld_wt + e] = p.alpha[F - 1] * v[c];
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
          for (int s = 0; s < K / 16; ++s) {
            const uint32_t a_hi = a_reg + (uint32_t)(32 * (s >> 1) + 8 * (s & 1)), a_lo = a_hi + 16;
            const uint32_t boff = (uint32_t)(s >> 2) * (uint32_t)N * 128u + (uint32_t)(s & 3) * 32u;
            const uint64_t dbh = make_desc(b_hi + boff, 16, 1024), dbl = make_desc(b_lo + boff, 16, 1024);
            mma_f16_ts(d_tmem, a_hi, dbh, idesc, s ? 1u : 0u);
            mma_f16_ts(d_tmem, a_lo, dbh, idesc, 1u);
            mma_f16_ts(d_tmem, a_hi, dbl, idesc, 1u);
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

}  // namespace

// Shapes the fused kernels cover: 128-wide hidden layers, 1..3 of them, a Bessel basis of 16, 32 or 64 functions,
// at most 16 live output columns.  Anything else: the caller keeps the per-layer GEMMs (eqnn_gemm_f32).
extern "C" int eqnn_radial_mlp_supported(int n_rb, int d_hidden, int n_hidden, int n_out) {
  return (n_rb == 16 || n_rb == 32 || n_rb == 64) && d_hidden == 128 && n_hidden >= 1 && n_hidden <= 3 && n_out >= 1 &&
         n_out <= 16;
}

extern "C" int eqnn_radial_mlp_fwd(const float* dist, int64_t dist_stride, int64_t n_edges, int n_rb, double r_cut,
                                   int n_hidden, int d_hidden, const float* const* weights, const int64_t* ldw,
                                   int n_out, float act_scale, float* w_t, int64_t ld_wt, eqnn_stream_t stream) {
  EQNN_CHECK_ARG(dist && weights && ldw && w_t && n_edges >= 0 && dist_stride >= 1, "radial_mlp_fwd: bad argument");
  EQNN_CHECK_ARG(eqnn_radial_mlp_supported(n_rb, d_hidden, n_hidden, n_out), "radial_mlp_fwd: unsupported shape");
  EQNN_CHECK_ARG(r_cut > 0.0 && ld_wt >= n_edges, "radial_mlp_fwd: bad r_cut / ld_wt");
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
  p.act_scale = act_scale; p.w_t = w_t; p.ld_wt = ld_wt;
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
