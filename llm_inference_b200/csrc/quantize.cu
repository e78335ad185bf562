// Activation quantizers on the device — bit-exact with the reference.
//
//   quantize_row_q8_0  ops.cpp:116-139 (ops.h:89-95)
//   quantize_row_q8_k  ops.cpp:142-178 (ops.h:98-105)
//   x -> f16 rounding  ops.cpp:542-551 (mat_vec_mul_fp16 prologue)
//
// Bit-exactness notes: every fp32 operation below is the IEEE one the CPU code
// performs (__fdiv_rn, never the fast-math forms).  nearest_int(x*id)
// (ops.cpp:107-113,135-136,166) is inlined in the reference build and, under
// the ops library's -mfma (BUILD:45-50), compiles to ONE fused multiply-add
// with the 12582912.f magic constant (vfmadd132ss in the object code): the
// product is not rounded before the round-to-integer.  nearest_int_fma() below
// reproduces exactly that; a separate multiply would differ on ~1e-5 of the
// elements (those within an fp32 rounding error of k+0.5).  f32_to_f16 (gguf.cpp:68-97) is value-identical to IEEE RNE ==
// __float2half_rn for every non-NaN input (checked exhaustively in the survey
// and again for the port oracle in tests/test_oracle.py).
#include <cuda_fp16.h>

#include "llmi_internal.h"

namespace {

__device__ __forceinline__ int nearest_int_fma(float a, float b) {
  return (__float_as_int(__fmaf_rn(a, b, 12582912.0f)) & 0x007fffff) - 0x00400000;
}

// One warp per block of 32.  Output layout ACT_Q8_0 (llmi_internal.h):
//   [n int8 quants][n/32 x {f16 d, int16 sum of the 32 quants}]
// The per-block sum lets the Q4_0 kernel use unsigned nibbles:
//   sum (nib-8)*q = sum nib*q - 8*sum q     (exact in integers).
__global__ void quantize_q8_0_kernel(const float* __restrict__ x, uint32_t n_blocks, uint32_t n, uint8_t* buf) {
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n_blocks) return;
  const float v = x[warp * 32 + lane];
  float amax = fabsf(v);
#pragma unroll
  for (int o = 16; o; o >>= 1) amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
  const float d = __fdiv_rn(amax, 127.0f);
  const float id = d != 0.0f ? __fdiv_rn(1.0f, d) : 0.0f;
  const int q = nearest_int_fma(v, id);
  int sum = q;
#pragma unroll
  for (int o = 16; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  reinterpret_cast<int8_t*>(buf)[warp * 32 + lane] = (int8_t)q;
  if (lane == 0) {
    const uint32_t meta = uint32_t(__half_as_ushort(__float2half_rn(d))) | (uint32_t(uint16_t(int16_t(sum))) << 16);
    reinterpret_cast<uint32_t*>(buf + n)[warp] = meta;
  }
}

// One CTA of 256 threads per super-block of 256.  Output layout ACT_Q8_K:
//   [n int8 quants][n/16 int16 bsums][n/256 fp32 d]
// "max" is the signed value at the FIRST index of largest |x| (strict > in the
// reference loop), so ties are broken towards the smaller index.
__global__ void quantize_q8_k_kernel(const float* __restrict__ x, uint32_t n, uint8_t* buf) {
  __shared__ unsigned long long s_key[8];
  __shared__ float s_max;
  const uint32_t sb = blockIdx.x;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
  const float v = x[sb * 256 + t];
  float av = fabsf(v);
  if (av != av) av = 0.0f;  // NaN never wins the `ax > amax` test
  unsigned long long key = (uint64_t(__float_as_uint(av)) << 32) | uint32_t(255 - t);
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, o);
    key = other > key ? other : key;
  }
  if (lane == 0) s_key[wid] = key;
  __syncthreads();
  key = s_key[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) key = s_key[i] > key ? s_key[i] : key;
  const int idx = 255 - int(uint32_t(key));
  const float amax = __uint_as_float(uint32_t(key >> 32));
  if (t == idx) s_max = v;
  __syncthreads();
  int q = 0;
  float d = 0.0f;
  if (amax != 0.0f) {
    const float iscale = __fdiv_rn(-127.0f, s_max);
    q = nearest_int_fma(iscale, v);
    q = max(-128, min(127, q));
    d = __fdiv_rn(1.0f, iscale);
  }
  reinterpret_cast<int8_t*>(buf)[sb * 256 + t] = (int8_t)q;
  int sum = q;
#pragma unroll
  for (int o = 8; o; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);  // groups of 16 lanes
  if ((t & 15) == 0) reinterpret_cast<int16_t*>(buf + n)[sb * 16 + (t >> 4)] = (int16_t)sum;
  if (t == 0) reinterpret_cast<float*>(buf + n + 2 * (n / 16))[sb] = d;
}

__global__ void round_f16_kernel(const float* __restrict__ x, uint32_t n, uint32_t n_pad, __half* out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n_pad) out[i] = i < n ? __float2half_rn(x[i]) : __ushort_as_half(0);
}

// Reference record layouts, for the bit-exact tests.
__global__ void export_q8_0_kernel(const uint8_t* __restrict__ buf, uint32_t n, uint8_t* out) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n / 32) return;
  const uint32_t meta = reinterpret_cast<const uint32_t*>(buf + n)[b];
  uint8_t* o = out + b * 34;
  o[0] = meta & 0xff;
  o[1] = (meta >> 8) & 0xff;
  for (int j = 0; j < 32; ++j) o[2 + j] = buf[b * 32 + j];
}
__global__ void export_q8_k_kernel(const uint8_t* __restrict__ buf, uint32_t n, uint8_t* out) {
  const uint32_t sb = blockIdx.x * blockDim.x + threadIdx.x;
  if (sb >= n / 256) return;
  uint8_t* o = out + sb * 292;
  const uint8_t* d = buf + n + 2 * (n / 16) + 4 * sb;
  for (int j = 0; j < 4; ++j) o[j] = d[j];
  for (int j = 0; j < 256; ++j) o[4 + j] = buf[sb * 256 + j];
  const uint8_t* bs = buf + n + 32 * sb;
  for (int j = 0; j < 32; ++j) o[260 + j] = bs[j];
}

}  // namespace

cudaError_t llmi_launch_quantize_q8_0(const float* x, uint64_t n, uint8_t* buf, cudaStream_t s) {
  const uint32_t nb = uint32_t(n / 32);
  if (!nb) return cudaSuccess;
  const int threads = 128;  // 4 blocks of 32 per CTA
  quantize_q8_0_kernel<<<(nb + 3) / 4, threads, 0, s>>>(x, nb, uint32_t(n), buf);
  return cudaGetLastError();
}

cudaError_t llmi_launch_quantize_q8_k(const float* x, uint64_t n, uint8_t* buf, cudaStream_t s) {
  const uint32_t nsb = uint32_t(n / 256);
  if (!nsb) return cudaSuccess;
  quantize_q8_k_kernel<<<nsb, 256, 0, s>>>(x, uint32_t(n), buf);
  return cudaGetLastError();
}

cudaError_t llmi_launch_round_f16(const float* x, uint64_t n, uint8_t* buf, cudaStream_t s) {
  const uint32_t n_pad = uint32_t(round_up(n, 8));
  if (!n_pad) return cudaSuccess;
  round_f16_kernel<<<(n_pad + 255) / 256, 256, 0, s>>>(x, uint32_t(n), n_pad, reinterpret_cast<__half*>(buf));
  return cudaGetLastError();
}

cudaError_t llmi_launch_export_q8_0(const uint8_t* buf, uint64_t n, uint8_t* out34, cudaStream_t s) {
  const uint32_t nb = uint32_t(n / 32);
  if (!nb) return cudaSuccess;
  export_q8_0_kernel<<<(nb + 127) / 128, 128, 0, s>>>(buf, uint32_t(n), out34);
  return cudaGetLastError();
}

cudaError_t llmi_launch_export_q8_k(const uint8_t* buf, uint64_t n, uint8_t* out292, cudaStream_t s) {
  const uint32_t nsb = uint32_t(n / 256);
  if (!nsb) return cudaSuccess;
  export_q8_k_kernel<<<(nsb + 63) / 64, 64, 0, s>>>(buf, uint32_t(n), out292);
  return cudaGetLastError();
}
