// Warp-level activation quantizers shared by the standalone quantize kernels
// and the fused glue kernels (norm -> quantize, GEGLU -> quantize, attention ->
// quantize).  Bit-exact with the reference (ops.cpp:116-178), see quantize.cu.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

namespace llmi_dev {

// nearest_int(a*b) as the reference's object code evaluates it: one fused
// multiply-add with the 12582912.f magic constant.
__device__ __forceinline__ int nearest_int_fma(float a, float b) {
  return (__float_as_int(__fmaf_rn(a, b, 12582912.0f)) & 0x007fffff) - 0x00400000;
}

// One warp quantizes block `b` (32 values, lane = element) into the ACT_Q8_0
// layout: [n int8][n/32 x {f16 d, int16 sum}].
__device__ __forceinline__ void warp_quantize_q8_0(float v, uint32_t b, uint32_t n, uint8_t* buf, int lane) {
  float amax = fabsf(v);
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float d = __fdiv_rn(amax, 127.0f);
  const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
  const int q = nearest_int_fma(v, id);
  int sum = q;
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  reinterpret_cast<int8_t*>(buf)[b * 32 + lane] = (int8_t)q;
  if (lane == 0)
    reinterpret_cast<uint32_t*>(buf + n)[b] =
        uint32_t(__half_as_ushort(__float2half_rn(d))) | (uint32_t(uint16_t(int16_t(sum))) << 16);
}

// NK blocks at once (lane = element of each): the same arithmetic with the NK shuffle chains interleaved, so the
// warp pays the latency of one chain instead of NK.  Every lane runs every shuffle; only the stores look at live[k].
template <int NK>
__device__ __forceinline__ void warp_quantize_q8_0_multi(const float (&v)[NK], const bool (&live)[NK],
                                                         const uint32_t (&b)[NK], uint32_t n, uint8_t* buf, int lane) {
  float amax[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) amax[k] = fabsf(v[k]);
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int k = 0; k < NK; ++k) amax[k] = fmaxf(amax[k], __shfl_xor_sync(0xffffffffu, amax[k], o));
  float d[NK];
  int q[NK], sum[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    d[k] = __fdiv_rn(amax[k], 127.0f);
    const float id = d[k] != 0.0f ? __fdiv_rn(1.0f, d[k]) : 0.0f;
    q[k] = nearest_int_fma(v[k], id);
    sum[k] = q[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int k = 0; k < NK; ++k) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    if (live[k]) {
      reinterpret_cast<int8_t*>(buf)[b[k] * 32 + lane] = (int8_t)q[k];
      if (lane == 0)
        reinterpret_cast<uint32_t*>(buf + n)[b[k]] =
            uint32_t(__half_as_ushort(__float2half_rn(d[k]))) | (uint32_t(uint16_t(int16_t(sum[k]))) << 16);
    }
  }
}

// One warp quantizes super-block `sb` (256 values; lane holds elements
// 8*lane .. 8*lane+7) into the ACT_Q8_K layout:
// [n int8][n/16 int16 bsums][n/256 fp32 d].  The scale comes from the signed
// value at the FIRST index of largest magnitude (ops.cpp:149-157).
__device__ __forceinline__ void warp_quantize_q8_k(const float (&v)[8], uint32_t sb, uint32_t n, uint8_t* buf,
                                                   int lane) {
  unsigned long long key = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    float av = fabsf(v[i]);
    if (av != av) av = 0.0f;
    const unsigned long long k = (uint64_t(__float_as_uint(av)) << 32) | uint32_t(255 - (lane * 8 + i));
    key = k > key ? k : key;
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  const int idx = 255 - int(uint32_t(key));
  const float amax = __uint_as_float(uint32_t(key >> 32));
  float mine = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if ((idx & 7) == i) mine = v[i];
  const float vmax = __shfl_sync(0xffffffffu, mine, idx >> 3);
  int q[8];
  float d = 0.0f;
  if (amax != 0.0f) {
    const float iscale = __fdiv_rn(-127.0f, vmax);
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = max(-128, min(127, nearest_int_fma(iscale, v[i])));
    d = __fdiv_rn(1.0f, iscale);
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) q[i] = 0;
  }
  int sum = 0;
  uint32_t lo = 0, hi = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lo |= uint32_t(uint8_t(int8_t(q[i]))) << (8 * i);
    hi |= uint32_t(uint8_t(int8_t(q[4 + i]))) << (8 * i);
    sum += q[i] + q[4 + i];
  }
  reinterpret_cast<uint2*>(buf)[sb * 32 + lane] = make_uint2(lo, hi);
  sum += __shfl_xor_sync(0xffffffffu, sum, 1);  // group of 16 = two lanes
  if ((lane & 1) == 0) reinterpret_cast<int16_t*>(buf + n)[sb * 16 + (lane >> 1)] = (int16_t)sum;
  if (lane == 0) reinterpret_cast<float*>(buf + n + 2 * (n / 16))[sb] = d;
}

}  // namespace llmi_dev
