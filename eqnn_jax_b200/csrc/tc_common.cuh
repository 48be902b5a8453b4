// Thin inline-PTX layer for the Blackwell (sm_100a) tensor-core kernels: mbarrier, proxy fences,
// TMEM allocation, tcgen05.mma (kind::tf32), tcgen05.commit / ld, shared-memory matrix descriptors.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

namespace tc {

// ---- host side: per-device launch configuration ------------------------------------------
// cudaFuncAttributeMaxDynamicSharedMemorySize belongs to the (function, device) pair, and one process may drive
// several devices from several host threads (jax.pmap: one thread per device).  Every kernel instantiation owns one
// DeviceOnce: the attribute is set the first time the instantiation is launched on a device; concurrent first
// launches may both set it (idempotent), later launches cost one relaxed atomic load.
constexpr int kMaxDevices = 64;
struct DeviceOnce {
  std::atomic<unsigned char> done[kMaxDevices];
};
template <typename Kernel>
inline cudaError_t configure_smem(DeviceOnce& once, Kernel kern, int smem_bytes, int* device_out = nullptr) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  if (device_out) *device_out = dev;
  if (dev < 0 || dev >= kMaxDevices || !once.done[dev].load(std::memory_order_acquire)) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < kMaxDevices) once.done[dev].store(1, std::memory_order_release);
  }
  return cudaSuccess;
}
// SM count of the current device, cached per device (cudaDeviceGetAttribute is a driver round trip)
inline int sm_count() {
  static std::atomic<int> cache[kMaxDevices];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev >= 0 && dev < kMaxDevices) {
    const int c = cache[dev].load(std::memory_order_relaxed);
    if (c > 0) return c;
  }
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (dev >= 0 && dev < kMaxDevices) cache[dev].store(sms, std::memory_order_relaxed);
  return sms;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- mbarrier ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug traps (reported as a launch failure) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
    if (spin > (1u << 26)) __trap();
}

// generic-proxy shared-memory writes -> visible to the async proxy (tensor core reads)
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---- bulk asynchronous copies (TMA, non-tensor form): global -> shared, completion counted on an mbarrier ----
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// dst (shared), src (global) and bytes are multiples of 16
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_smem),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// contiguous block -> L2 (no data returned to the SM); addr and bytes multiples of 16
__device__ __forceinline__ void l2_prefetch_bulk(const void* gptr, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(gptr), "r"(bytes) : "memory");
}

// ---- tensor memory ----------------------------------------------------------------------
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {    // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 32 lanes x 32 consecutive fp32 columns: thread i of the warp gets lane (warp%4)*32+i
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
  uint32_t* r = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- descriptors --------------------------------------------------------------------------
// Shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, sm_100): 128-byte swizzle, version 1.
//   K-major : 8-row groups of 128-byte rows; SBO = bytes between 8-row groups; LBO unused (1)
//   MN-major: atoms of 8 K-rows x 128 bytes of M/N; LBO = bytes between atoms along M/N,
//             SBO = bytes between atoms along K
//   MN-major fp32/tf32 operands only exist as SWIZZLE_128B_BASE32B (layout type 1): atoms of 4 K-rows x
//             128 bytes, 32-byte chunks XOR-swizzled with the row index (cute Layout_MN_SW128_32B_Atom).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type = 2) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((saddr >> 4) & 0x3fff);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= 1ull << 46;   // version = 1 (Blackwell)
  d |= static_cast<uint64_t>(layout_type) << 61;   // 2 = SWIZZLE_128B, 1 = SWIZZLE_128B_BASE32B
  return d;
}
// Instruction descriptor (cute::UMMA::InstrDescriptor) for kind::tf32, fp32 accumulate.
__host__ __device__ constexpr uint32_t make_idesc_tf32(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(a_mn_major) << 15) |
         (static_cast<uint32_t>(b_mn_major) << 16) | (static_cast<uint32_t>(N >> 3) << 17) |
         (static_cast<uint32_t>(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
      : "memory");
}
// arrive on an mbarrier once every previously issued MMA of this thread has completed
__device__ __forceinline__ void mma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}

// 256-bit global accesses (sm_100+): one thread moves a whole 32-byte sector, so a warp that owns one matrix ROW
// per lane (the tcgen05.ld register layout) still reads/writes complete sectors without a shared-memory transpose.
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%8], {%0,%1,%2,%3,%4,%5,%6,%7};" ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]), "l"(p)
               : "memory");
}

// fp32 -> (hi, lo) with hi exactly representable in tf32 (10 explicit mantissa bits) and lo = a - hi exact.
// a*b ~= hi_a*hi_b + lo_a*hi_b + hi_a*lo_b with relative error ~2^-21: fp32-class accuracy on the tf32 pipe.
__device__ __forceinline__ void split_tf32(float a, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(a) & 0xffffe000u);
  lo = a - hi;
}

// ---- packed fp32x2 arithmetic (FFMA2 / FMUL2 / FADD2: one issue slot for two lanes' worth of fp32) ----------
// The producer and epilogue warps of the GEMM kernels are issue-bound, so their element-wise math is written on
// register pairs.  ex2/rcp stay scalar (MUFU).
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk2(float a, float b) {
  f32x2 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}
__device__ __forceinline__ void upk2(f32x2 v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ float ex2_ftz(float x) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
// sigmoid(2*k0*(x + k1*x^3)) for a pair: 1 / (1 + 2^(x*(C0 + C1*x^2))), C0 = -2*k0*log2(e), C1 = C0*k1
__device__ __forceinline__ f32x2 gelu_sigmoid2(f32x2 x, f32x2 xx) {
  constexpr float C0 = -2.0f * 0.7978845608028654f * 1.4426950408889634f, C1 = C0 * 0.044715f;
  const f32x2 t = mul2(fma2(xx, pk2(C1, C1), pk2(C0, C0)), x);
  float t0, t1;
  upk2(t, t0, t1);
  const f32x2 d = add2(pk2(ex2_ftz(t0), ex2_ftz(t1)), pk2(1.f, 1.f));
  float d0, d1;
  upk2(d, d0, d1);
  return pk2(rcp_ftz(d0), rcp_ftz(d1));
}
// gelu_tanh(x) * scale for a pair (same function as common.cuh gelu_tanh)
__device__ __forceinline__ f32x2 gelu2(f32x2 x, f32x2 scale) {
  return mul2(mul2(x, scale), gelu_sigmoid2(x, mul2(x, x)));
}
// d/dx gelu_tanh(x) for a pair: s + x*s*(1-s)*2*k0*(1 + 3*k1*x^2)
__device__ __forceinline__ f32x2 gelu_grad2(f32x2 x) {
  constexpr float D0 = 2.0f * 0.7978845608028654f, D1 = D0 * 3.0f * 0.044715f;
  const f32x2 xx = mul2(x, x);
  const f32x2 s = gelu_sigmoid2(x, xx);
  const f32x2 w = mul2(x, fma2(xx, pk2(D1, D1), pk2(D0, D0)));
  const f32x2 q = mul2(s, fma2(s, pk2(-1.f, -1.f), pk2(1.f, 1.f)));
  return fma2(q, w, s);
}
// (hi, lo) split of a pair: hi = tf32-representable head, lo = exact remainder
__device__ __forceinline__ void split2(f32x2 a, f32x2& hi, f32x2& lo) {
  hi = a & 0xffffe000ffffe000ull;
  lo = fma2(hi, pk2(-1.f, -1.f), a);
}
// float4 -> optional gelu*scale -> (hi, lo) float4s
template <bool ACT>
__device__ __forceinline__ void act_split4(const float4& v, float scale, float4& hi, float4& lo) {
  f32x2 a = pk2(v.x, v.y), b = pk2(v.z, v.w);
  if (ACT) {
    const f32x2 sc = pk2(scale, scale);
    a = gelu2(a, sc);
    b = gelu2(b, sc);
  }
  f32x2 ah, al, bh, bl;
  split2(a, ah, al);
  split2(b, bh, bl);
  upk2(ah, hi.x, hi.y);
  upk2(bh, hi.z, hi.w);
  upk2(al, lo.x, lo.y);
  upk2(bl, lo.z, lo.w);
}

// byte offset of element (row, 16-byte chunk c) inside a 128-byte-swizzled tile whose rows are 128 bytes
__device__ __forceinline__ uint32_t sw128(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}

// byte offset of (k-row r, 16-byte chunk c of a 128-byte row) in an MN-major SWIZZLE_128B_BASE32B slab whose
// 4-row atoms are packed 512 bytes apart along K
__device__ __forceinline__ uint32_t sw128_base32(uint32_t r, uint32_t c) {
  return (r >> 2) * 512u + (r & 3u) * 128u + ((((c >> 1) ^ (r & 3u)) << 5) | ((c & 1u) << 4));
}

}  // namespace tc
