// Dev tool: the phase-3 recurrence of the attention kernel in isolation (one CTA),
// to separate the cost of the rounding chain from the operand fetch and the barrier.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 tools/p3_bench.cu -o /tmp/p3 && /tmp/p3
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }
constexpr int D = 256, ROWS = 64;

template <int MODE>
__global__ void __launch_bounds__(1024) p3(float* out, long long* cyc, int reps) {
  extern __shared__ __align__(16) uint8_t sm[];
  __half* tile = reinterpret_cast<__half*>(sm);                 // [ROWS][D]
  float* se = reinterpret_cast<float*>(sm + ROWS * D * 2);      // [ROWS]
  uint8_t* nm = reinterpret_cast<uint8_t*>(se + ROWS);          // [ROWS]
  for (int i = threadIdx.x; i < ROWS * D; i += blockDim.x) tile[i] = __float2half(0.001f * float(i % 97) - 0.04f);
  for (int i = threadIdx.x; i < ROWS; i += blockDim.x) {
    se[i] = 0.5f + 0.001f * i;
    nm[i] = 0;
  }
  __syncthreads();
  const uint32_t e = threadIdx.x;
  const bool active = e < D;
  float v = 0.0f;
  const long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (active) {
      const __half* col = tile + e;
      if (MODE == 0) {  // straightforward: fetch a chunk, fold it
        for (int r = 0; r + 8 <= ROWS; r += 8) {
          float x8[8], e8[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            x8[k] = __half2float(col[(r + k) * D]);
            e8[k] = se[r + k];
          }
          const uint2 m8 = *reinterpret_cast<const uint2*>(nm + r);
          if ((m8.x | m8.y) == 0) {
#pragma unroll
            for (int k = 0; k < 8; ++k) v = r16(__fmaf_rn(x8[k], e8[k], v));
          } else {
            v += 1.0f;
          }
        }
      } else if (MODE == 1) {  // all operands of the tile fetched up front (registers), then one long chain
        float x[ROWS];
#pragma unroll
        for (int r = 0; r < ROWS; ++r) x[r] = __half2float(col[r * D]);
#pragma unroll
        for (int r = 0; r < ROWS; ++r) v = r16(__fmaf_rn(x[r], se[r], v));
      } else if (MODE == 2) {  // chain only, operands loop-invariant
        const float x = __half2float(col[0]), s = se[1];
#pragma unroll 8
        for (int r = 0; r < ROWS; ++r) v = r16(__fmaf_rn(x, s, v));
      } else if (MODE == 3) {  // 16 rows per chunk, packed fetch first, convert late
        for (int r = 0; r + 16 <= ROWS; r += 16) {
          __half xr[16];
          float e16[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) xr[k] = col[(r + k) * D];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 f = *reinterpret_cast<const float4*>(se + r + 4 * k);
            e16[4 * k] = f.x; e16[4 * k + 1] = f.y; e16[4 * k + 2] = f.z; e16[4 * k + 3] = f.w;
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) v = r16(__fmaf_rn(__half2float(xr[k]), e16[k], v));
        }
      }
    }
    __syncthreads();
  }
  const long long t1 = clock64();
  if (active) out[e] = v;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* out;
  long long* cyc;
  cudaMalloc(&out, 4096);
  cudaMalloc(&cyc, 8);
  const int smem = ROWS * D * 2 + ROWS * 5 + 64, reps = 50;
  const char* names[] = {"chunk of 8: fetch then fold", "whole tile to registers, then chain", "chain only (invariant operands)",
                         "chunk of 16, packed fetch"};
  for (int m = 0; m < 4; ++m) {
    for (int th : {256, 1024}) {
      for (int w = 0; w < 2; ++w) {
        if (m == 0) p3<0><<<1, th, smem>>>(out, cyc, reps);
        if (m == 1) p3<1><<<1, th, smem>>>(out, cyc, reps);
        if (m == 2) p3<2><<<1, th, smem>>>(out, cyc, reps);
        if (m == 3) p3<3><<<1, th, smem>>>(out, cyc, reps);
        cudaDeviceSynchronize();
      }
      long long c;
      cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-40s %4d threads: %.1f cycles/position (%s)\n", names[m], th, double(c) / reps / ROWS,
             cudaGetErrorString(cudaGetLastError()));
    }
  }
  return 0;
}
