// Device-resident glue between the mat-vecs of one decode token:
// embedding-row gather, RMSNorm (+ residual) fused with the activation
// quantizer of the next mat-vec, per-head q/k norm + RoPE + KV append,
// attention with the reference's exact online-softmax rounding sequence,
// GEGLU fused with the quantizer, logit soft-cap + argmax.  These exist only so
// that activations and the KV cache never leave the device between mat-vecs;
// each follows the reference's arithmetic (file:line cited per kernel),
// including which operations its object code fuses.
#include <cuda_fp16.h>
#include <math.h>

#include "glue.h"
#include "launch.cuh"
#include "quant_device.cuh"

namespace {

using namespace llmi_dev;

__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// ---------------------------------------------------------------- embedding
// One element of row `row` of a repacked matrix, dequantized as the reference's
// row dequantizers do (model.cpp:251-322 -> ops.cpp:1005-1082): F16 direct,
// Q8_0 d*q, Q5_0 d*(q-16), Q6_K d*sc*q (left to right).  Plane item order:
// repack.cu.
__device__ float dequant_elem(const EmbedArgs& a, uint32_t row, uint32_t e) {
  const uint32_t s = row >> 3, r = row & 7;
  const uint64_t nb = a.nb;
  switch (a.type) {
    case LLMI_F16: {
      const uint64_t cell = (uint64_t(s) * nb + (e >> 3)) * 8 + r;
      return h2f(reinterpret_cast<const uint16_t*>(a.q)[cell * 8 + (e & 7)]);
    }
    case LLMI_Q8_0: {
      const uint32_t b = e >> 5, i = e & 31;
      const uint64_t cell = (uint64_t(s) * nb + b) * 8 + r;
      const int8_t qv = reinterpret_cast<const int8_t*>(a.q)[(((uint64_t(s) * nb + b) * 2 + (i >> 4)) * 8 + r) * 16 + (i & 15)];
      return h2f(reinterpret_cast<const uint16_t*>(a.d)[cell]) * float(qv);
    }
    case LLMI_Q5_0: {
      const uint32_t b = e >> 5, i = e & 31;
      const uint64_t cell = (uint64_t(s) * nb + b) * 8 + r;
      const uint8_t byte = a.q[cell * 16 + (i & 15)];
      const uint32_t qh = reinterpret_cast<const uint32_t*>(a.x)[cell];
      const int qv = int((i < 16 ? (byte & 0x0f) : (byte >> 4)) | (((qh >> i) & 1u) << 4));
      return h2f(reinterpret_cast<const uint16_t*>(a.d)[cell]) * float(qv - 16);
    }
    case LLMI_Q6_K: {
      const uint32_t sb = e >> 8, w = e & 255, n = w >> 7, l = w & 127, g = l >> 5, ll = l & 31;
      const uint32_t hh = ll >> 4, bi = ll & 15, sub = 2 * n + hh;
      const uint64_t su = uint64_t(s) * nb + sb;
      const uint8_t* q = a.q;
      const uint8_t ql = q[(((su * 3 + ((g & 1) ? 1 : 0)) * 4 + sub) * 8 + r) * 16 + bi];
      const uint8_t qh = q[(((su * 3 + 2) * 4 + sub) * 8 + r) * 16 + bi];
      const int lo = (g >= 2) ? (ql >> 4) : (ql & 0x0f);
      const int qv = int(int8_t(lo | (((qh >> (2 * g)) & 3) << 4))) - 32;
      const int8_t sc = reinterpret_cast<const int8_t*>(a.x)[(su * 8 + r) * 16 + 8 * n + hh + 2 * g];
      const float d = h2f(reinterpret_cast<const uint16_t*>(a.d)[su * 8 + r]);
      return d * float(sc) * float(qv);
    }
    default: return 0.0f;
  }
}

// embed_tokens + scale_embeddings (model.cpp:240-344): h = dequant(row) * sqrt(float(E))
__global__ void embed_kernel(EmbedArgs a, const int32_t* __restrict__ token, float scale, float* __restrict__ h) {
  pdl_trigger();
  pdl_wait();
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.n_cols) return;
  h[e] = dequant_elem(a, uint32_t(*token), e) * scale;
}

// ----------------------------------------------------------------- reductions
// Deterministic block sum (fixed tree).  The reference sums squares
// sequentially in fp32 (ops.cpp:33-36); a GPU cannot afford a 1152-5376 long
// dependent chain per norm, so the order differs (~1e-7 relative on the scale).
__device__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  float s = 0.0f;
  for (int i = 0; i < nw; ++i) s += red[i];
  return s;
}

// rms_norm's scale (ops.cpp:37-38): mean = sum/size in fp32, eps added in
// DOUBLE, rounded to fp32, sqrtf, 1.0f/.
__device__ __forceinline__ float rms_scale(float sum, uint32_t n, double eps) {
  const float mean = __fdiv_rn(sum, float(n));
  return __fdiv_rn(1.0f, __fsqrt_rn(float(double(mean) + eps)));
}

// Writes the activation of kind `kind` for the float vector xs[0..n) held in
// shared memory (whole CTA participates).
__device__ void emit_act(int kind, const float* xs, uint32_t n, uint8_t* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (kind == ACT_Q8_0) {
    for (uint32_t b = warp; b < n / 32; b += nw) warp_quantize_q8_0(xs[b * 32 + lane], b, n, buf, lane);
  } else if (kind == ACT_Q8_K) {
    for (uint32_t sb = warp; sb < n / 256; sb += nw) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = xs[sb * 256 + lane * 8 + i];
      warp_quantize_q8_k(v, sb, n, buf, lane);
    }
  } else if (kind == ACT_F16) {
    const uint32_t n_pad = (n + 7) & ~7u;
    for (uint32_t i = threadIdx.x; i < n_pad; i += blockDim.x)
      reinterpret_cast<uint16_t*>(buf)[i] = i < n ? f2h(xs[i]) : uint16_t(0);
  } else if (kind == ACT_F32) {
    const uint32_t n_pad = (n + 3) & ~3u;
    for (uint32_t i = threadIdx.x; i < n_pad; i += blockDim.x) reinterpret_cast<float*>(buf)[i] = i < n ? xs[i] : 0.0f;
  }
}

// ------------------------------------------------------- norm (+ residual) + act
// Single CTA.  Optional first stage (post-norm + residual, model.cpp:843-854 /
// 915-924):   h += (rms_scale(y) * y) * w_post
// Optional second stage (run_norm, model.cpp:346-386 + the quantizer of the
// next mat-vec):   xn = (rms_scale(h) * h) * w ; act = quantize(xn)
// Every global input (y, h, both norm weights) is loaded into registers up
// front — one memory round trip instead of one per stage; the rest is two block
// reductions and the quantizer.  NORM_PER elements per thread (n <= 6 * 1024).
constexpr int NORM_PER = 6;
__global__ void __launch_bounds__(1024) norm_act_kernel(NormArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float xs[];  // n floats
  __shared__ float red[32];
  const uint32_t n = a.n;
  float yv[NORM_PER], hv[NORM_PER], wp[NORM_PER], wn[NORM_PER];
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {
    const uint32_t i = threadIdx.x + k * blockDim.x;
    const bool ok = i < n;
    hv[k] = ok ? a.h[i] : 0.0f;
    yv[k] = (ok && a.y) ? a.y[i] : 0.0f;
    wp[k] = (ok && a.y && a.w_post) ? a.w_post[i] : 0.0f;
    wn[k] = (ok && a.w) ? a.w[i] : 0.0f;
  }
  if (a.pos_inc && threadIdx.x == 0) *a.pos_inc += 1;
  if (a.y) {
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) ss += __fmul_rn(yv[k], yv[k]);
    const float sc = rms_scale(block_sum(ss, red), n, a.eps);
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) {
      const uint32_t i = threadIdx.x + k * blockDim.x;
      // with a post-norm: h += (scale*y)*w (model.cpp:843-854); without: h += y
      const float add = a.w_post ? __fmul_rn(__fmul_rn(sc, yv[k]), wp[k]) : yv[k];
      hv[k] = __fadd_rn(hv[k], add);
      if (i < n) a.h[i] = hv[k];
    }
  }
  if (!a.w) return;
  float ss = 0.0f;
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) ss += __fmul_rn(hv[k], hv[k]);
  const float sc = rms_scale(block_sum(ss, red), n, a.eps);
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {
    const uint32_t i = threadIdx.x + k * blockDim.x;
    if (i < n) {
      const float v = __fmul_rn(__fmul_rn(sc, hv[k]), wn[k]);
      xs[i] = v;
      if (a.xn_out) a.xn_out[i] = v;
    }
  }
  __syncthreads();
  emit_act(a.act_kind, xs, n, a.act_buf);
}

// Generic multi-CTA quantizer of a device vector (after attention).
__global__ void act_kernel(const float* __restrict__ x, uint32_t n, int kind, uint8_t* buf) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  if (kind == ACT_Q8_0) {
    for (uint32_t b = gw; b < n / 32; b += nw) warp_quantize_q8_0(x[b * 32 + lane], b, n, buf, lane);
  } else if (kind == ACT_Q8_K) {
    for (uint32_t sb = gw; sb < n / 256; sb += nw) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = x[sb * 256 + lane * 8 + i];
      warp_quantize_q8_k(v, sb, n, buf, lane);
    }
  } else if (kind == ACT_F16) {
    const uint32_t n_pad = (n + 7) & ~7u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
      reinterpret_cast<uint16_t*>(buf)[i] = i < n ? f2h(x[i]) : uint16_t(0);
  } else if (kind == ACT_F32) {
    const uint32_t n_pad = (n + 3) & ~3u;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x)
      reinterpret_cast<float*>(buf)[i] = i < n ? x[i] : 0.0f;
  }
}

// ------------------------------------------------------- q/k norm, RoPE, KV append
// One CTA per head job: blockIdx.x in [0,H) = q head, [H,H+HK) = k head,
// [H+HK, H+2HK) = v head.  blockDim = D/2; thread i owns elements i and i+D/2
// (the NEOX rotation pair, ops.cpp:85-92).
//   q: run_norm (model.cpp:388-423) -> rope (ops.cpp:67-95) -> scale (ops.cpp:97-105)
//   k: run_norm -> rope -> f32_to_f16 -> cache[pos] (model.cpp:442-474)
//   v: f32_to_f16 -> cache[pos]
// rope in the reference's object code: x0' = fma(v0, cos, -(v1*sin)),
// x1' = fma(v0, sin, v1*cos); angle = (float(pos) * (1/powf(base, 2i/n_rot))) / scale.
__global__ void qkv_post_kernel(QkvArgs a) {
  pdl_trigger();
  pdl_wait();
  __shared__ float red[32];
  const uint32_t D = a.D, half = D / 2, i = threadIdx.x;
  const uint32_t job = blockIdx.x;
  const int pos = *a.pos;
  if (job >= a.H + a.HK) {  // v head
    const uint32_t hv = job - a.H - a.HK;
    const float* v = a.v + hv * D;
    __half* dst = a.vcache + (size_t(pos) * a.HK + hv) * D;
    dst[i] = __float2half_rn(v[i]);
    dst[i + half] = __float2half_rn(v[i + half]);
    return;
  }
  const bool is_q = job < a.H;
  const uint32_t hd = is_q ? job : job - a.H;
  const float* src = (is_q ? a.q : a.k) + hd * D;
  const float* w = is_q ? a.wq_norm : a.wk_norm;
  float v0 = src[i], v1 = src[i + half];
  const float ss = block_sum(__fadd_rn(__fmul_rn(v0, v0), __fmul_rn(v1, v1)), red);
  const float sc = rms_scale(ss, D, a.eps);
  v0 = __fmul_rn(__fmul_rn(sc, v0), w[i]);
  v1 = __fmul_rn(__fmul_rn(sc, v1), w[i + half]);
  const float freq = __fdiv_rn(1.0f, powf(a.rope_base, __fdiv_rn(float(2 * i), float(int(D)))));
  const float ang = __fdiv_rn(__fmul_rn(float(pos), freq), a.rope_scale);
  float sn, cs;
  sincosf(ang, &sn, &cs);
  float x0 = __fmaf_rn(v0, cs, -__fmul_rn(v1, sn));
  float x1 = __fmaf_rn(v0, sn, __fmul_rn(v1, cs));
  if (is_q) {
    a.q_out[hd * D + i] = __fmul_rn(x0, a.attn_scale);
    a.q_out[hd * D + i + half] = __fmul_rn(x1, a.attn_scale);
  } else {
    __half* dst = a.kcache + (size_t(pos) * a.HK + hd) * D;
    dst[i] = __float2half_rn(x0);
    dst[i + half] = __float2half_rn(x1);
  }
}

// ---------------------------------------------------------------- attention
// Model::run_attn (model.cpp:476-550) for one query token, one CTA per head.
// The reference walks the cached positions sequentially with an fp16 value
// accumulator that is rounded at every step (vec_mad_f16 / vec_scale_f16,
// ops.cpp:1084-1099) — that recurrence is kept element by element (phase 3);
// everything that does not depend on it is computed in parallel first:
//   phase 1  score[t] = sum_i double(f16(k[t][i]) * f16(q[i]))      (:504-509)
//   phase 2  running max M (prefix max of float(score)), and per position
//            new_max / score_exp / prev_score_exp exactly as :520-533
//   phase 3  per element: v = f16(v*pse) on a new max; v = f16(fma(x, se, v))
//   phase 4  out = f32(v) / s_acc                                     (:543-547)
__global__ void attention_kernel(AttnArgs a) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) uint8_t smraw[];
  const uint32_t D = a.D, h = blockIdx.x, hkv = h / (a.H / a.HK);
  const int T = *a.pos + 1;
  double* sc = reinterpret_cast<double*>(smraw);                 // [Tmax]
  float* se = reinterpret_cast<float*>(sc + a.t_max);            // [Tmax]
  float* pse = se + a.t_max;                                     // [Tmax] (first holds M_prev)
  float* qh = pse + a.t_max;                                     // [D]
  uint8_t* nm = reinterpret_cast<uint8_t*>(qh + D);              // [Tmax]
  __shared__ float s_inv;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) qh[i] = __half2float(__float2half_rn(a.q[h * D + i]));
  __syncthreads();
  // phase 1
  for (int t = warp; t < T; t += nw) {
    const __half* kp = a.kcache + (size_t(t) * a.HK + hkv) * D;
    double s = 0.0;
    for (uint32_t i = lane; i < D; i += 32) s += double(__fmul_rn(__half2float(kp[i]), qh[i]));
#pragma unroll
    for (int o = 16; o; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
      if (a.softcap > 0.0f) s = double(__fmul_rn(a.softcap, tanhf(float(s / double(a.softcap)))));
      sc[t] = s;
    }
  }
  __syncthreads();
  // phase 2a: exclusive prefix max of float(score) (warp 0)
  if (warp == 0) {
    const int chunk = (T + 31) / 32, t0 = lane * chunk, t1 = min(T, t0 + chunk);
    float m = -INFINITY;
    for (int t = t0; t < t1; ++t) m = fmaxf(m, float(sc[t]));
    float run = m;  // inclusive scan of chunk maxima
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float other = __shfl_up_sync(0xffffffffu, run, o);
      if (lane >= o) run = fmaxf(run, other);
    }
    float prev = __shfl_up_sync(0xffffffffu, run, 1);
    if (lane == 0) prev = -INFINITY;
    for (int t = t0; t < t1; ++t) {
      pse[t] = prev;
      prev = fmaxf(prev, float(sc[t]));
    }
  }
  __syncthreads();
  // phase 2b
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float M = pse[t];
    const double s = sc[t];
    if (s > double(M)) {
      const float fs = float(s);
      nm[t] = 1;
      se[t] = 1.0f;
      pse[t] = expf(__fsub_rn(M, fs));
    } else {
      nm[t] = 0;
      se[t] = expf(float(s - double(M)));
      pse[t] = 1.0f;
    }
  }
  __syncthreads();
  // phase 2c: s_acc = s_acc*pse + se, sequential, no FMA (model.cpp:540)
  if (threadIdx.x == blockDim.x - 1) {
    float s = 0.0f;
    for (int t = 0; t < T; ++t) s = __fadd_rn(__fmul_rn(s, pse[t]), se[t]);
    s_inv = s == 0.0f ? 0.0f : __fdiv_rn(1.0f, s);
  }
  // phase 3
  for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) {
    const __half* vp = a.vcache + size_t(hkv) * D + i;
    const size_t stride = size_t(a.HK) * D;
    __half v = __float2half_rn(0.0f);
    for (int t = 0; t < T; ++t) {
      const float x = __half2float(vp[t * stride]);
      if (nm[t]) v = __float2half_rn(__fmul_rn(__half2float(v), pse[t]));
      v = __float2half_rn(__fmaf_rn(x, se[t], __half2float(v)));
    }
    qh[i] = __half2float(v);  // q no longer needed
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) {
    const float o = __fmul_rn(qh[i], s_inv);
    qh[i] = o;
    a.out[h * D + i] = o;
  }
  // Fused quantizer of the attn_output mat-vec: this head's D outputs are whole
  // Q8_0 blocks (and whole Q8_K super-blocks when D % 256 == 0).
  if (a.act_kind == ACT_NONE) return;
  __syncthreads();
  const uint32_t n = a.H * D;
  if (a.act_kind == ACT_Q8_0) {
    for (uint32_t b = warp; b < D / 32; b += nw) warp_quantize_q8_0(qh[b * 32 + lane], h * (D / 32) + b, n, a.act_buf, lane);
  } else if (a.act_kind == ACT_Q8_K) {
    for (uint32_t sb = warp; sb < D / 256; sb += nw) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = qh[sb * 256 + lane * 8 + i];
      warp_quantize_q8_k(v, h * (D / 256) + sb, n, a.act_buf, lane);
    }
  } else if (a.act_kind == ACT_F16) {
    for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) reinterpret_cast<uint16_t*>(a.act_buf)[h * D + i] = f2h(qh[i]);
  } else if (a.act_kind == ACT_F32) {
    for (uint32_t i = threadIdx.x; i < D; i += blockDim.x) reinterpret_cast<float*>(a.act_buf)[h * D + i] = qh[i];
  }
}

// ------------------------------------------------------------------- GEGLU + act
// model.cpp:887-901: gelu_x = 0.5f*x*(1.0f + tanhf(sqrtf(2/pi)*(x + 0.044715f*x*x*x)));
// hidden = gelu_x * up.  model.cpp is built without FMA: every operation rounds.
__device__ __forceinline__ float geglu(float x, float up) {
  const float c = 0.7978845608028654f;  // sqrtf(float(2.0f / M_PI))
  const float x3 = __fmul_rn(__fmul_rn(__fmul_rn(0.044715f, x), x), x);
  const float inner = __fmul_rn(c, __fadd_rn(x, x3));
  const float g = __fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, tanhf(inner)));
  return __fmul_rn(g, up);
}

__global__ void geglu_act_kernel(const float* __restrict__ gate, const float* __restrict__ up, uint32_t n, int kind,
                                 uint8_t* buf, float* hidden_out) {
  pdl_trigger();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  if (kind == ACT_Q8_0) {
    for (uint32_t b = gw; b < n / 32; b += nw) {
      const float v = geglu(gate[b * 32 + lane], up[b * 32 + lane]);
      if (hidden_out) hidden_out[b * 32 + lane] = v;
      warp_quantize_q8_0(v, b, n, buf, lane);
    }
  } else if (kind == ACT_Q8_K) {
    for (uint32_t sb = gw; sb < n / 256; sb += nw) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t e = sb * 256 + lane * 8 + i;
        v[i] = geglu(gate[e], up[e]);
        if (hidden_out) hidden_out[e] = v[i];
      }
      warp_quantize_q8_k(v, sb, n, buf, lane);
    }
  } else {
    const uint32_t n_pad = kind == ACT_F16 ? ((n + 7) & ~7u) : ((n + 3) & ~3u);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
      const float v = i < n ? geglu(gate[i], up[i]) : 0.0f;
      if (hidden_out && i < n) hidden_out[i] = v;
      if (kind == ACT_F16) reinterpret_cast<uint16_t*>(buf)[i] = f2h(v);
      else reinterpret_cast<float*>(buf)[i] = v;
    }
  }
}

// ------------------------------------------------------------ soft-cap + argmax
// model.cpp:1036-1041 (final logit soft-cap) and main.cpp:193-194 (greedy:
// std::max_element = FIRST index of the maximum) are fused into the epilogue of
// the logits mat-vec (gemv.cu), which leaves a 64-bit key = ordered value bits
// << 32 | ~index.  This kernel turns the key into the next token, appends it to
// the generated list and re-arms the key.
__global__ void finish_token_kernel(unsigned long long* key, int32_t* cur_tok, int32_t* gen, int32_t* gen_count) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    const int32_t tok = int32_t(0xffffffffu - uint32_t(*key));
    *key = 0ull;
    if (cur_tok) *cur_tok = tok;
    if (gen && gen_count) {
      gen[*gen_count] = tok;
      *gen_count += 1;
    }
  }
}

__global__ void softcap_kernel(float* logits, uint32_t n, float softcap) {
  pdl_trigger();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) logits[i] = __fmul_rn(softcap, tanhf(__fdiv_rn(logits[i], softcap)));
}

}  // namespace

// ------------------------------------------------------------------ launchers

cudaError_t llmi_launch_embed(const EmbedArgs& a, const int32_t* token, float scale, float* h, cudaStream_t s) {
  return llmi_launch(embed_kernel, dim3((a.n_cols + 255) / 256), dim3(256), 0, s, a, token, scale, h);
}

cudaError_t llmi_launch_norm_act(const NormArgs& a, cudaStream_t s) {
  const int threads = a.n >= 2048 ? 1024 : 512;
  return llmi_launch(norm_act_kernel, dim3(1), dim3(threads), a.n * sizeof(float), s, a);
}

cudaError_t llmi_launch_act(const float* x, uint32_t n, int kind, uint8_t* buf, cudaStream_t s) {
  const uint32_t warps = kind == ACT_Q8_0 ? n / 32 : (kind == ACT_Q8_K ? n / 256 : (n + 31) / 32);
  const uint32_t blocks = (warps + 7) / 8 ? (warps + 7) / 8 : 1;
  return llmi_launch(act_kernel, dim3(blocks), dim3(256), 0, s, x, n, kind, buf);
}

cudaError_t llmi_launch_qkv_post(const QkvArgs& a, cudaStream_t s) {
  return llmi_launch(qkv_post_kernel, dim3(a.H + 2 * a.HK), dim3(a.D / 2), 0, s, a);
}

size_t llmi_attention_smem(uint32_t t_max, uint32_t D) { return size_t(t_max) * (8 + 4 + 4 + 1) + size_t(D) * 4 + 16; }

cudaError_t llmi_attention_init(uint32_t t_max, uint32_t D) {
  return cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              int(llmi_attention_smem(t_max, D)));
}

cudaError_t llmi_launch_attention(const AttnArgs& a, cudaStream_t s) {
  return llmi_launch(attention_kernel, dim3(a.H), dim3(256), llmi_attention_smem(a.t_max, a.D), s, a);
}

cudaError_t llmi_launch_geglu_act(const float* gate, const float* up, uint32_t n, int kind, uint8_t* buf,
                                  float* hidden_out, cudaStream_t s) {
  const uint32_t warps = kind == ACT_Q8_0 ? n / 32 : (kind == ACT_Q8_K ? n / 256 : (n + 31) / 32);
  const uint32_t blocks = (warps + 3) / 4 ? (warps + 3) / 4 : 1;
  return llmi_launch(geglu_act_kernel, dim3(blocks), dim3(128), 0, s, gate, up, n, kind, buf, hidden_out);
}

cudaError_t llmi_launch_finish_token(unsigned long long* key, int32_t* cur_tok, int32_t* gen, int32_t* gen_count,
                                     cudaStream_t s) {
  return llmi_launch(finish_token_kernel, dim3(1), dim3(32), 0, s, key, cur_tok, gen, gen_count);
}

cudaError_t llmi_launch_softcap(float* logits, uint32_t n, float softcap, cudaStream_t s) {
  return llmi_launch(softcap_kernel, dim3((n + 255) / 256), dim3(256), 0, s, logits, n, softcap);
}
