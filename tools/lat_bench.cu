// Dev tool: dependent-chain latency (cycles per step) of the candidate inner loops.
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
__device__ __forceinline__ float r16_cvt(float x) { return __half2float(__float2half_rn(x)); }
__device__ __forceinline__ float r16_alu(float x) {
  const uint32_t b = __float_as_uint(x);
  const uint32_t ex = b & 0x7f800000u;
  float c = __uint_as_float((b & 0xff800000u) + (13u << 23)) * 1.5f;
  c = ex >= (113u << 23) ? c : copysignf(0.75f, x);
  return __fsub_rn(__fadd_rn(x, c), c);
}
template <int MODE>
__global__ void chain(float* out, const float* in, int n, long long* cyc) {
  float v = in[threadIdx.x & 31], x = in[32 + (threadIdx.x & 31)], e = in[64 + (threadIdx.x & 31)];
  double d = in[96 + (threadIdx.x & 31)];
  const long long t0 = clock64();
  for (int i = 0; i < n; ++i) {
    if (MODE == 0) v = r16_cvt(__fmaf_rn(x, e, v));
    if (MODE == 1) v = r16_alu(__fmaf_rn(x, e, v));
    if (MODE == 2) v = __fmaf_rn(x, e, v);
    if (MODE == 3) d += double(__fmul_rn(x, v));
    if (MODE == 4) d = __dadd_rn(d, 1.0000001);
    if (MODE == 5) { v = __fmul_rn(v, x); d += double(v); }             // fmul + F2F.F64.F32 + dadd, not hoistable
    if (MODE == 6) d = __fma_rn(d, 1.0000001, double(e));                // dfma chain
    if (MODE == 7) { v = __fmul_rn(v, x); e += __half2float(__float2half_rn(v)); }  // F2FP + HADD2 off the critical chain
  }
  const long long t1 = clock64();
  out[threadIdx.x] = v + float(d) + e;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void check_r16(unsigned long long* bad) {
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long u = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; u < (1ull << 32); u += stride) {
    const float x = __uint_as_float((uint32_t)u);
    if (!(fabsf(x) < 65520.0f)) continue;
    if (__float_as_uint(r16_cvt(x)) != __float_as_uint(r16_alu(x))) atomicAdd(bad, 1ull);
  }
}
int main() {
  float *in, *out; long long* cyc; unsigned long long* bad;
  cudaMalloc(&in, 1024); cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8); cudaMalloc(&bad, 8);
  float h[256]; for (int i = 0; i < 256; ++i) h[i] = 0.001f * (i % 7 + 1);
  cudaMemcpy(in, h, 1024, cudaMemcpyHostToDevice); cudaMemset(bad, 0, 8);
  const int n = 4096; long long c;
  const char* names[] = {"fma + F2FP + HADD2 (cvt round trip)", "fma + ALU r16", "fma only", "fmul + cvt.f64 + dadd", "dadd only"};
  for (int m = 0; m < 5; ++m) {
    for (int rep = 0; rep < 2; ++rep) {
      if (m == 0) chain<0><<<1, 32>>>(out, in, n, cyc); if (m == 1) chain<1><<<1, 32>>>(out, in, n, cyc);
      if (m == 2) chain<2><<<1, 32>>>(out, in, n, cyc); if (m == 3) chain<3><<<1, 32>>>(out, in, n, cyc);
      if (m == 4) chain<4><<<1, 32>>>(out, in, n, cyc);
      cudaDeviceSynchronize();
    }
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-40s %.1f cycles/step (1 warp)\n", names[m], double(c) / n);
  }
  // throughput with 8 and 32 warps on one SM
  const char* n2[] = {"fma + F2FP + HADD2 chain", "", "fma only", "", "dadd only", "fmul + F2F.F64.F32 + dadd", "dfma chain", "fmul; F2FP+HADD2+fadd"};
  for (int m : {0, 2, 4, 5, 6, 7}) {
    for (int th : {32, 256, 1024}) {
      if (m == 0) chain<0><<<1, th>>>(out, in, n, cyc); if (m == 2) chain<2><<<1, th>>>(out, in, n, cyc);
      if (m == 4) chain<4><<<1, th>>>(out, in, n, cyc); if (m == 5) chain<5><<<1, th>>>(out, in, n, cyc);
      if (m == 6) chain<6><<<1, th>>>(out, in, n, cyc); if (m == 7) chain<7><<<1, th>>>(out, in, n, cyc);
      cudaDeviceSynchronize(); cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("%-32s %6.1f cycles/step (%d warps on one SM)\n", n2[m], double(c) / n, th / 32);
    }
  }
  check_r16<<<148 * 8, 256>>>(bad); cudaDeviceSynchronize();
  unsigned long long nb; cudaMemcpy(&nb, bad, 8, cudaMemcpyDeviceToHost);
  printf("r16_alu vs cvt mismatches over all |x| < 65520: %llu (%s)\n", nb, cudaGetErrorString(cudaGetLastError()));
  return 0;
}
