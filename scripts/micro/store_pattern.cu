// micro-test: store/load patterns of the row-GEMM epilogue, 148 persistent CTAs x 256 threads, 128-row x 512-B tiles
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void stg256(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7]) : "memory");
}
__device__ __forceinline__ void ldg256(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];" : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]),
               "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
}
// MODE 0: row-per-lane 256-bit stores (current epilogue).  MODE 1: coalesced float4 (lane = 16 B of a row, warp = 512 B row)
// MODE 2: row-per-lane 256-bit LOADS (aux pattern) + reduce.  MODE 3: coalesced float4 loads.
template <int MODE>
__global__ void __launch_bounds__(256) k(float* C, int64_t n_tiles, float* sink) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float acc = 0.f;
  for (int64_t t = blockIdx.x; t < n_tiles; t += gridDim.x) {
    float* tile = C + t * 128 * 128;
    if (MODE == 0 || MODE == 2) {
      const int row = (warp & 3) * 32 + lane, col0 = (warp >> 2) * 64;
#pragma unroll
      for (int cb = 0; cb < 4; ++cb) {
        float v[16];
        if (MODE == 0) {
#pragma unroll
          for (int j = 0; j < 16; ++j) v[j] = (float)(t + j);
          stg256(tile + row * 128 + col0 + cb * 16, v);
          stg256(tile + row * 128 + col0 + cb * 16 + 8, v + 8);
        } else {
          ldg256(tile + row * 128 + col0 + cb * 16, v);
          ldg256(tile + row * 128 + col0 + cb * 16 + 8, v + 8);
#pragma unroll
          for (int j = 0; j < 16; ++j) acc += v[j];
        }
      }
    } else {
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const int row = warp * 16 + i;
        float4* p = reinterpret_cast<float4*>(tile + row * 128) + lane;
        if (MODE == 1) *p = make_float4((float)t, 1.f, 2.f, 3.f);
        else { float4 v = __ldg(p); acc += v.x + v.y + v.z + v.w; }
      }
    }
  }
  if (acc == 123.456f) *sink = acc;
}
template <int MODE> float run(float* C, int64_t n_tiles, float* sink, int grid) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<grid, 256>>>(C, n_tiles, sink);
  cudaEventRecord(e0);
  for (int i = 0; i < 5; ++i) k<MODE><<<grid, 256>>>(C, n_tiles, sink);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms / 5;
}
int main() {
  const int64_t E = 3200000, n_tiles = E / 128;
  float *C, *sink; cudaMalloc(&C, E * 128 * 4); cudaMalloc(&sink, 4);
  cudaMemset(C, 0, E * 128 * 4);
  for (int grid : {148, 296, 592}) {
    printf("grid %d: 1.64 GB  rowlane-st256 %.3f ms | coalesced-st128 %.3f ms | rowlane-ld256 %.3f ms | coalesced-ld128 %.3f ms\n", grid,
           run<0>(C, n_tiles, sink, grid), run<1>(C, n_tiles, sink, grid), run<2>(C, n_tiles, sink, grid), run<3>(C, n_tiles, sink, grid));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
