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

#include <type_traits>

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
// Row-sharded model (ll.peers.n > 0, one token): the rank that holds the token's row sends it to everybody.
__global__ void embed_kernel(EmbedArgs a, const int32_t* __restrict__ token, float scale, float* __restrict__ h,
                             LLCtx ll, uint32_t ll_off) {
  pdl_trigger();
  pdl_wait();
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;  // m: token of a prefill batch
  if (e >= a.n_cols) return;
  const uint32_t row = uint32_t(token[m]);
  if (ll.peers.n == 0) {
    h[size_t(m) * a.n_cols + e] = dequant_elem(a, row, e) * scale;
    return;
  }
  const uint32_t tag = ll_tag(ll.tag);
  if (row >= a.row_begin && row < a.row_end) {
    const float v = dequant_elem(a, row - a.row_begin, e) * scale;
    for (uint32_t p = 0; p < ll.peers.n; ++p) ll_store(ll.peers.base[p] + ll_off + e, __float_as_uint(v), tag);
  }
  h[e] = ll_waitf(ll.peers.base[ll.rank] + ll_off + e, tag, ll.tag.err);
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
  extern __shared__ float xs[];  // n floats
  __shared__ float red[32];
  const uint32_t n = a.n;
  {  // one CTA per token of a prefill batch (a single CTA when decoding)
    const size_t off = size_t(blockIdx.x) * n;
    if (a.y) a.y += off;
    a.h += off;
    if (a.xn_out) a.xn_out += off;
    if (a.act_buf) a.act_buf += size_t(blockIdx.x) * a.act_stride;
  }
  float yv[NORM_PER], hv[NORM_PER], wp[NORM_PER], wn[NORM_PER];
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {  // norm weights are static: fetch them under the predecessor's tail
    const uint32_t i = threadIdx.x + k * blockDim.x;
    const bool ok = i < n;
    wp[k] = (ok && (a.y || a.ll_y) && a.w_post) ? a.w_post[i] : 0.0f;
    wn[k] = (ok && a.w) ? a.w[i] : 0.0f;
  }
  pdl_wait();
  const uint32_t tag = a.ll_y ? ll_tag(a.ll_tag) : 0u;
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {
    const uint32_t i = threadIdx.x + k * blockDim.x;
    const bool ok = i < n;
    hv[k] = ok ? a.h[i] : 0.0f;
    if (a.ll_y) yv[k] = ok ? ll_waitf(a.ll_y + i, tag, a.ll_tag.err) : 0.0f;  // one decode token, grid = 1
    else yv[k] = (ok && a.y) ? a.y[i] : 0.0f;
  }
  if (a.pos_inc && threadIdx.x == 0 && blockIdx.x == 0) *a.pos_inc += int32_t(gridDim.x);
  if (a.y || a.ll_y) {
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
    // every thread read the epoch before block_sum's barriers: safe to open the next step's epoch now
    if (a.epoch_inc && threadIdx.x == 0) *a.epoch_inc += 1u;
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
__global__ void act_kernel(const float* __restrict__ x, uint32_t n, int kind, uint8_t* buf, uint32_t act_stride) {
  pdl_trigger();
  pdl_wait();
  x += size_t(blockIdx.y) * n;  // blockIdx.y: vector of a token batch
  buf += size_t(blockIdx.y) * act_stride;
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

// ------------------------------------- q/k norm, RoPE, KV append, attention
// One CTA per query head for one token (grid = H, 1024 threads; D <= 512):
//   q: run_norm (model.cpp:388-423) -> rope (ops.cpp:67-95) -> scale (:97-105)
//   k: run_norm -> rope -> f32_to_f16 ; v: f32_to_f16 ; append to the cache at
//      `pos` (model.cpp:442-474).  Every query head of a GQA group derives the
//      new K/V row itself (cheap) and keeps it in shared memory, the first head
//      of the group also writes it to the cache — so no CTA depends on another.
//   Model::run_attn (model.cpp:476-550): the reference walks the cached
//   positions sequentially with an fp16 value accumulator that is rounded at
//   every step (vec_mad_f16 / vec_scale_f16, ops.cpp:1084-1099).  That
//   recurrence is kept element by element (phase 3); everything that does not
//   depend on it is computed in parallel first:
//     phase 1  score[t] = sum_i double(f16(k[t][i]) * f16(q[i]))      (:504-509)
//     phase 2  running max M (prefix max of float(score)), and per position
//              new_max / score_exp / prev_score_exp exactly as :520-533
//     phase 3  per element: v = f16(v*pse) on a new max; v = f16(fma(x, se, v))
//     phase 4  out = f32(v) / s_acc (:543-547) + the quantizer of attn_output
// rope in the reference's object code: x0' = fma(v0, cos, -(v1*sin)),
// x1' = fma(v0, sin, v1*cos); angle = (float(pos) * (1/powf(base, 2i/n_rot))) / scale.
// K/V rows stream through a ring of shared-memory tiles of ATT_TILE_BYTES each
// (16384/D positions per tile), filled by per-row bulk async copies
// (cp.async.bulk -> UBLKCP) that complete on one mbarrier per tile.
constexpr int ATT_TILE_BYTES = 32768;
constexpr int ATT_MAX_BUF = 6;

#ifdef LLMI_ATTN_TIMING  // dev only (tools/attn_bench.cu): cycle stamps of CTA 0 at the phase boundaries.
// BAR.SYNC does not block at issue, so the stamp is made to depend on the barrier's result.
__device__ long long g_attn_stamp[16];
#define ATTN_STAMP(i)                                                     \
  do {                                                                    \
    const int c_ = __syncthreads_count(1);                                \
    if (c_ < 0) return;                                                   \
    if (blockIdx.x == 0 && threadIdx.x == 0) g_attn_stamp[i] = clock64(); \
  } while (0)
#define ATTN_RAW(i, tid) do { if (blockIdx.x == 0 && threadIdx.x == (tid)) g_attn_stamp[i] = clock64(); } while (0)
#else
#define ATTN_STAMP(i) do { } while (0)
#define ATTN_RAW(i, tid) do { } while (0)
#endif

// RoPE factors for every (position, pair): ops.cpp:80-83 —
//   freq = 1.0f / powf(base, float(2i)/n_rot); val = float(pos) * freq / scale; cosf(val), sinf(val)
__global__ void rope_table_kernel(float2* table, uint32_t t_max, uint32_t D, float base, float scale) {
  const uint32_t half = D / 2;
  const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (idx >= uint64_t(t_max) * half) return;
  const uint32_t pos = uint32_t(idx / half), i = uint32_t(idx % half);
  const float freq = __fdiv_rn(1.0f, powf(base, __fdiv_rn(float(2 * i), float(int(D)))));
  const float ang = __fdiv_rn(__fmul_rn(float(pos), freq), scale);
  float sn, cs;
  sincosf(ang, &sn, &cs);
  table[idx] = make_float2(cs, sn);
}

__device__ __forceinline__ float r16(float x) { return __half2float(__float2half_rn(x)); }

// The K cache holds, per element, the HIGH WORD OF THE DOUBLE that equals the
// f16-rounded key (the low word of such a double is zero).  The reference's
// score is sum_i double(f32(k_i) * f32(q_i)) (model.cpp:504-509); the product
// of two f16 values is exact in fp32, so fma(double(k_i), double(q_i), s)
// rounds exactly the same sum — one DFMA per element and no conversion in the
// loop (F2F.F64.F32 runs at 16 lanes/clk/SM and would bound the phase).
__device__ __forceinline__ uint32_t f16_as_double_hi(__half h) { return uint32_t(__double2hiint(double(__half2float(h)))); }

// MODE 0: decode — prologue (q/k norm, RoPE, KV append) and attention for one token in one kernel.
// Prefill processes a batch of tokens (blockIdx.y) in two kernels, because a token attends to rows the
// other CTAs of the batch append:  MODE 1 = prologue only (appends K/V, leaves f16(q) as double high
// words in a.qbuf);  MODE 2 = attention only (q from a.qbuf, every row — its own included — from the cache).
template <int D, int MODE>
__global__ void __launch_bounds__(1024) attention_kernel(AttnArgs a, uint32_t nbuf) {
  pdl_trigger();
  constexpr int HALF = D / 2, VEC = D / 32;           // elements per lane in phase 1
  constexpr int RTK = ATT_TILE_BYTES / (4 * D);       // K rows per tile (4 bytes per element)
  constexpr int RTV = ATT_TILE_BYTES / (2 * D);       // V rows per tile (f16)
  constexpr int PIECES = VEC >= 4 ? VEC / 4 : 1, PW = VEC >= 4 ? 4 : VEC;  // 16-byte pieces per lane / words per piece
  extern __shared__ __align__(128) uint8_t smraw[];
  __shared__ float red[32];
  __shared__ float wmax[32];
  __shared__ float s_inv;
  __shared__ __align__(8) uint64_t bars[ATT_MAX_BUF];
  const uint32_t h = blockIdx.x, group = a.H / a.HK, hkv = h / group;
  const uint32_t tp = (a.t_max + 15) & ~15u;
  uint8_t* tiles = smraw;                                                    // [nbuf][ATT_TILE_BYTES]
  double* sc = reinterpret_cast<double*>(smraw + size_t(nbuf) * ATT_TILE_BYTES);  // [tp] scores
  float* se = reinterpret_cast<float*>(sc + tp);                             // [tp] score_exp
  float* pse = se + tp;                                                      // [tp] prev_score_exp
  float* qh = pse + tp;                                                      // [D] output staging
  uint32_t* qhi = reinterpret_cast<uint32_t*>(qh + D);                       // [D] f16(q) as double high words
  uint32_t* knew = qhi + D;                                                  // [D] this token's K row (double high words)
  __half* vnew = reinterpret_cast<__half*>(knew + D);                        // [D] this token's V row (f16)
  uint8_t* nm = reinterpret_cast<uint8_t*>(vnew + D);                        // [tp] 1 where the running max moves
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const uint32_t i = threadIdx.x;
  const bool pair = i < HALF;  // thread i owns the rotation pair (i, i + D/2)
  // static inputs, fetched under the predecessor's tail
  float wq0 = 0.0f, wq1 = 0.0f, wk0 = 0.0f, wk1 = 0.0f;
  if (pair) {
    wq0 = a.wq_norm[i];
    wq1 = a.wq_norm[i + HALF];
    wk0 = a.wk_norm[i];
    wk1 = a.wk_norm[i + HALF];
  }
  if (threadIdx.x == 0)
    for (uint32_t b = 0; b < nbuf; ++b) mbar_init(&bars[b], 1);
  ATTN_STAMP(0);
  pdl_wait();
  const uint32_t tok = blockIdx.y;  // token of a prefill batch (0 when decoding)
  const int pos = *a.pos + int(tok), T = pos + 1;
  a.q += size_t(tok) * a.H * D;
  a.k += size_t(tok) * a.HK * D;
  a.v += size_t(tok) * a.HK * D;
  a.out += size_t(tok) * a.H * D;
  if (a.act_buf) a.act_buf += size_t(tok) * a.act_stride;
  uint32_t* qglob = a.qbuf ? a.qbuf + (size_t(tok) * a.H + h) * D : nullptr;
  constexpr int OWN = MODE == 0 ? 1 : 0;  // the token's own K/V row comes from shared memory, not from the cache
  // The tile stream: K tiles 0..n_tk-1, then V tiles 0..n_tv-1, through the ring.
  // Rows of one KV head are contiguous ([HK][t_max][D]): one bulk copy per tile.
  // Row `pos` (always the last row of the last tile) is not in the cache yet: it
  // is taken from knew/vnew, so a tile's copy never depends on this kernel's stores.
  const int n_tk = (T + RTK - 1) / RTK, n_tv = (T + RTV - 1) / RTV, n_stream = n_tk + n_tv;
  auto issue_tile = [&](int j, int buf) {  // one thread
    uint64_t* bar = &bars[buf];
    uint8_t* dst = tiles + size_t(buf) * ATT_TILE_BYTES;
    if (j < n_tk) {
      const int t0 = j * RTK, n_cached = min(RTK, T - t0) - (j == n_tk - 1 ? OWN : 0);
      mbar_expect_tx(bar, uint32_t(n_cached) * D * 4);
      if (n_cached) bulk_g2s(dst, a.kcache + (size_t(hkv) * a.t_max + t0) * D, uint32_t(n_cached) * D * 4, bar);
    } else {
      const int jj = j - n_tk, t0 = jj * RTV, n_cached = min(RTV, T - t0) - (jj == n_tv - 1 ? OWN : 0);
      mbar_expect_tx(bar, uint32_t(n_cached) * D * 2);
      if (n_cached) bulk_g2s(dst, a.vcache + (size_t(hkv) * a.t_max + t0) * D, uint32_t(n_cached) * D * 2, bar);
    }
  };
  if (MODE != 1 && threadIdx.x == 0)
    for (int j = 0; j < min(int(nbuf), n_stream); ++j) issue_tile(j, j);
  int cbuf = 0;          // ring position of the tile being consumed
  uint32_t cpar = 0;     // its mbarrier phase parity
  int cj = 0;            // its index in the stream
  auto tile_done = [&]() {  // all threads; the consumed tile is refilled with the tile nbuf further down
    __syncthreads();
    if (threadIdx.x == 0 && cj + int(nbuf) < n_stream) issue_tile(cj + int(nbuf), cbuf);
    ++cj;
    if (++cbuf == int(nbuf)) {
      cbuf = 0;
      cpar ^= 1;
    }
  };
  if (MODE == 2) {  // q was normalized, rotated and rounded by the MODE 1 kernel
    for (uint32_t e = threadIdx.x; e < D; e += blockDim.x) qhi[e] = qglob[e];
  } else {
  float q0 = 0.0f, q1 = 0.0f, k0 = 0.0f, k1 = 0.0f, v0 = 0.0f, v1 = 0.0f;
  float2 csn = make_float2(1.0f, 0.0f);
  if (pair && a.ll_q) {  // row-sharded model: the q/k/v rows of every rank land in the exchange buffer
    const uint32_t tag = ll_tag(a.ll_tag);
    q0 = ll_waitf(a.ll_q + h * D + i, tag, a.ll_tag.err);
    q1 = ll_waitf(a.ll_q + h * D + i + HALF, tag, a.ll_tag.err);
    k0 = ll_waitf(a.ll_k + hkv * D + i, tag, a.ll_tag.err);
    k1 = ll_waitf(a.ll_k + hkv * D + i + HALF, tag, a.ll_tag.err);
    v0 = ll_waitf(a.ll_v + hkv * D + i, tag, a.ll_tag.err);
    v1 = ll_waitf(a.ll_v + hkv * D + i + HALF, tag, a.ll_tag.err);
    csn = a.rope_table[size_t(pos) * HALF + i];
  } else if (pair) {
    q0 = a.q[h * D + i];
    q1 = a.q[h * D + i + HALF];
    k0 = a.k[hkv * D + i];
    k1 = a.k[hkv * D + i + HALF];
    v0 = a.v[hkv * D + i];
    v1 = a.v[hkv * D + i + HALF];
    // (cos, sin) of (float(pos) * (1/powf(base, 2i/D))) / scale: tabulated per position at
    // load time by rope_table_kernel with exactly this arithmetic
    csn = a.rope_table[size_t(pos) * HALF + i];
  }
  ATTN_STAMP(1);
  // sums of squares of the q and k head: only the first D/64 warps hold data
  if (warp < (HALF + 31) / 32) {
    float sq = __fadd_rn(__fmul_rn(q0, q0), __fmul_rn(q1, q1)), sk = __fadd_rn(__fmul_rn(k0, k0), __fmul_rn(k1, k1));
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      sq += __shfl_xor_sync(0xffffffffu, sq, o);
      sk += __shfl_xor_sync(0xffffffffu, sk, o);
    }
    if (lane == 0) {
      red[warp] = sq;
      red[16 + warp] = sk;
    }
  }
  __syncthreads();
  if (pair) {
    float ssq = 0.0f, ssk = 0.0f;
#pragma unroll
    for (int w = 0; w < (HALF + 31) / 32; ++w) {
      ssq += red[w];
      ssk += red[16 + w];
    }
    const float scq = rms_scale(ssq, D, a.eps), sck = rms_scale(ssk, D, a.eps);
    q0 = __fmul_rn(__fmul_rn(scq, q0), wq0);
    q1 = __fmul_rn(__fmul_rn(scq, q1), wq1);
    k0 = __fmul_rn(__fmul_rn(sck, k0), wk0);
    k1 = __fmul_rn(__fmul_rn(sck, k1), wk1);
    const float cs = csn.x, sn = csn.y;
    const float qa = __fmul_rn(__fmaf_rn(q0, cs, -__fmul_rn(q1, sn)), a.attn_scale);
    const float qb = __fmul_rn(__fmaf_rn(q0, sn, __fmul_rn(q1, cs)), a.attn_scale);
    const uint32_t qah = f16_as_double_hi(__float2half_rn(qa));  // Q is rounded to f16 for the scores (model.cpp:506)
    const uint32_t qbh = f16_as_double_hi(__float2half_rn(qb));
    if (MODE == 1) {
      qglob[i] = qah;
      qglob[i + HALF] = qbh;
    } else {
      qhi[i] = qah;
      qhi[i + HALF] = qbh;
    }
    const __half ka = __float2half_rn(__fmaf_rn(k0, cs, -__fmul_rn(k1, sn)));
    const __half kb = __float2half_rn(__fmaf_rn(k0, sn, __fmul_rn(k1, cs)));
    const __half va = __float2half_rn(v0), vb = __float2half_rn(v1);
    const uint32_t kah = f16_as_double_hi(ka), kbh = f16_as_double_hi(kb);
    if (MODE == 0) {
      knew[i] = kah;
      knew[i + HALF] = kbh;
      vnew[i] = va;
      vnew[i + HALF] = vb;
    }
    if (h % group == 0) {  // one writer per KV head
      uint32_t* kd = a.kcache + (size_t(hkv) * a.t_max + pos) * D;
      __half* vd = a.vcache + (size_t(hkv) * a.t_max + pos) * D;
      kd[i] = kah;
      kd[i + HALF] = kbh;
      vd[i] = va;
      vd[i + HALF] = vb;
    }
  }
  if (MODE == 1) return;
  }
  __syncthreads();
  ATTN_STAMP(2);
  // phase 1: scores.  One warp per position, four positions in flight per warp.
  // A lane owns the 16-byte pieces {c*128 + 4*lane .. +3} of the row (conflict-free
  // LDS.128); its slice of q stays in registers as doubles.
  {
    double qd[VEC];
#pragma unroll
    for (int c = 0; c < PIECES; ++c)
#pragma unroll
      for (int w = 0; w < PW; ++w) qd[c * PW + w] = __hiloint2double(int(qhi[c * 32 * PW + lane * PW + w]), 0);
    auto lane_dot = [&](const uint32_t* row) {
      double s = 0.0;
#pragma unroll
      for (int c = 0; c < PIECES; ++c) {
        uint32_t w[4] = {0, 0, 0, 0};
        const uint32_t* p = row + c * 32 * PW + lane * PW;
        if (PW == 4) {
          const uint4 u = *reinterpret_cast<const uint4*>(p);
          w[0] = u.x; w[1] = u.y; w[2] = u.z; w[3] = u.w;
        } else {
          const uint2 u = *reinterpret_cast<const uint2*>(p);
          w[0] = u.x; w[1] = u.y;
        }
#pragma unroll
        for (int k = 0; k < PW; ++k) s = __fma_rn(__hiloint2double(int(w[k]), 0), qd[c * PW + k], s);
      }
      return s;
    };
    // K tiles are consumed in sweeps of up to 4 tiles (all of them in flight
    // since the kernel started or since the previous sweep), so that every warp
    // has four positions and the CTA synchronizes once per sweep.
    const uint32_t* tiles32 = reinterpret_cast<const uint32_t*>(tiles);
    const int G = min(4, int(nbuf));
    while (cj < n_tk) {
      const int g = min(n_tk - cj, G);
      {
        int b2 = cbuf;
        uint32_t p2 = cpar;
        for (int x = 0; x < g; ++x) {
          mbar_wait(&bars[b2], p2);
          if (++b2 == int(nbuf)) {
            b2 = 0;
            p2 ^= 1;
          }
        }
      }
      const int t0 = cj * RTK, nt = min(g * RTK, T - t0);
      for (int r = 4 * warp; r < nt; r += 4 * nw) {
        double s[4];
#pragma unroll
        for (int p = 0; p < 4; ++p) {
          const int rr = min(r + p, nt - 1);
          int slot = cbuf + rr / RTK;
          if (slot >= int(nbuf)) slot -= int(nbuf);
          s[p] = lane_dot(OWN && t0 + rr == pos ? knew
                                                : tiles32 + size_t(slot) * (ATT_TILE_BYTES / 4) + size_t(rr % RTK) * D);
        }
#pragma unroll
        for (int o = 16; o; o >>= 1)
#pragma unroll
          for (int p = 0; p < 4; ++p) s[p] += __shfl_xor_sync(0xffffffffu, s[p], o);
        if (lane < 4 && r + lane < nt) {
          double v = lane == 0 ? s[0] : (lane == 1 ? s[1] : (lane == 2 ? s[2] : s[3]));
          if (a.softcap > 0.0f) v = double(__fmul_rn(a.softcap, tanhf(float(v / double(a.softcap)))));
          sc[t0 + r + lane] = v;
        }
      }
      __syncthreads();  // the sweep's tiles are free: refill them with the tiles nbuf further down the stream
      for (int x = 0; x < g; ++x) {
        if (threadIdx.x == 0 && cj + int(nbuf) < n_stream) issue_tile(cj + int(nbuf), cbuf);
        ++cj;
        if (++cbuf == int(nbuf)) {
          cbuf = 0;
          cpar ^= 1;
        }
      }
    }
  }
  ATTN_STAMP(3);
  // phase 2a/2b: one position per thread (blocks of blockDim positions).  The
  // running max M before position t (model.cpp:520-533) is an exclusive prefix
  // max of float(score): warp scan, then the maxima of the preceding warps.
  {
    float carry = -INFINITY;  // max over all earlier blocks
    for (int base = 0; base < T; base += int(blockDim.x)) {
      const int t = base + int(threadIdx.x);
      const double s = t < T ? sc[t] : 0.0;
      const float fs = t < T ? float(s) : -INFINITY;
      float inc = fs;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float other = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc = fmaxf(inc, other);
      }
      float exc = __shfl_up_sync(0xffffffffu, inc, 1);
      if (lane == 0) exc = -INFINITY;
      if (lane == 31) wmax[warp] = inc;
      __syncthreads();
      float wm = wmax[lane];  // blockDim is 1024: one entry per warp
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const float other = __shfl_up_sync(0xffffffffu, wm, o);
        if (lane >= o) wm = fmaxf(wm, other);
      }
      float prevw = __shfl_sync(0xffffffffu, wm, warp ? warp - 1 : 0);
      if (warp == 0) prevw = -INFINITY;
      const float M = fmaxf(fmaxf(carry, prevw), exc);
      carry = fmaxf(carry, __shfl_sync(0xffffffffu, wm, 31));
      if (t < T) {
        if (s > double(M)) {
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
    }
  }
  ATTN_STAMP(4);
  // phase 2c: s_acc = s_acc*pse + se, sequential, no FMA (model.cpp:540), on a
  // thread of a warp that idles in phase 3.  The operands are fetched 8 positions
  // at a time so that only the mul+add chain is serial, not the shared-memory latency.
  if (threadIdx.x == blockDim.x - 1) {
    ATTN_RAW(8, 1023);
    float s = 0.0f;
    int t = 0;
    for (; t + 8 <= T; t += 8) {
      const float4 pa = *reinterpret_cast<const float4*>(pse + t), pb = *reinterpret_cast<const float4*>(pse + t + 4);
      const float4 ea = *reinterpret_cast<const float4*>(se + t), eb = *reinterpret_cast<const float4*>(se + t + 4);
      s = __fadd_rn(__fmul_rn(s, pa.x), ea.x);
      s = __fadd_rn(__fmul_rn(s, pa.y), ea.y);
      s = __fadd_rn(__fmul_rn(s, pa.z), ea.z);
      s = __fadd_rn(__fmul_rn(s, pa.w), ea.w);
      s = __fadd_rn(__fmul_rn(s, pb.x), eb.x);
      s = __fadd_rn(__fmul_rn(s, pb.y), eb.y);
      s = __fadd_rn(__fmul_rn(s, pb.z), eb.z);
      s = __fadd_rn(__fmul_rn(s, pb.w), eb.w);
    }
    for (; t < T; ++t) s = __fadd_rn(__fmul_rn(s, pse[t]), se[t]);
    s_inv = s == 0.0f ? 0.0f : __fdiv_rn(1.0f, s);
    ATTN_RAW(9, 1023);
  }
  // phase 3: the fp16 accumulator recurrence, one thread per element.  The
  // accumulator is an fp32 register that always holds an f16-representable
  // value; r16(x) = f32(f16(x)) is the rounding the reference applies at every
  // step (vec_mad_f16, ops.cpp:1084-1099).  A new running maximum (rare: ~ln T
  // positions) rescales first; chunks of 8 positions without one take the short
  // chain.  The raw operands of the next chunk are fetched before this one folds
  // and converted after, so the shared-memory latency hides under the chain.
  {
    const uint32_t e = threadIdx.x;
    const bool active = e < D;
    float v = 0.0f;
    for (int jj = 0; jj < n_tv; ++jj) {
      const int t0 = jj * RTV, nt = min(RTV, T - t0);
      const int nc = nt - (jj == n_tv - 1 ? OWN : 0);  // rows that came from the cache
      const __half* col = reinterpret_cast<const __half*>(tiles + size_t(cbuf) * ATT_TILE_BYTES) + e;
      mbar_wait(&bars[cbuf], cpar);
      if (jj == 0) ATTN_RAW(10, 0);
      if (active) {
        // chunk of up to 16 positions at rows r..r+15 (t0 + r is a multiple of 16): fetch, then fold
        auto chunk = [&](int r, int rem, auto full) {
          constexpr bool FULL = decltype(full)::value;
          __half xr[16];
#pragma unroll
          for (int k = 0; k < 16; ++k) xr[k] = col[(FULL ? r + k : min(r + k, nc - 1)) * D];
          float e16[16];
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const float4 f = *reinterpret_cast<const float4*>(se + t0 + r + 4 * k);
            e16[4 * k] = f.x; e16[4 * k + 1] = f.y; e16[4 * k + 2] = f.z; e16[4 * k + 3] = f.w;
          }
          const uint4 mk = *reinterpret_cast<const uint4*>(nm + t0 + r);
          if ((mk.x | mk.y | mk.z | mk.w) == 0) {
#pragma unroll
            for (int k = 0; k < 16; ++k)
              if (FULL || k < rem) v = r16(__fmaf_rn(__half2float(xr[k]), e16[k], v));
          } else {
            // a new running maximum in the chunk: v = f16(v * prev_score_exp) first (model.cpp:528-533);
            // prev_score_exp is 1.0f elsewhere, which leaves the f16-valued v unchanged
            float p16[16];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const float4 f = *reinterpret_cast<const float4*>(pse + t0 + r + 4 * k);
              p16[4 * k] = f.x; p16[4 * k + 1] = f.y; p16[4 * k + 2] = f.z; p16[4 * k + 3] = f.w;
            }
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              if (FULL || k < rem) {
                v = r16(__fmul_rn(v, p16[k]));
                v = r16(__fmaf_rn(__half2float(xr[k]), e16[k], v));
              }
            }
          }
        };
        int r = 0;
        for (; r + 16 <= nc; r += 16) chunk(r, 16, std::true_type{});
        if (r < nc) chunk(r, nc - r, std::false_type{});
      }
      if (jj == 0) ATTN_RAW(11, 0);
      tile_done();
      if (jj == 0) ATTN_RAW(12, 0);
    }
    if (active) {
      if (OWN) {  // the current token's own row
        if (nm[pos]) v = r16(__fmul_rn(v, pse[pos]));
        v = r16(__fmaf_rn(__half2float(vnew[e]), se[pos], v));
      }
      qh[e] = v;
    }
  }
  ATTN_STAMP(5);
  __syncthreads();
  if (threadIdx.x < D) {
    const float o = __fmul_rn(qh[threadIdx.x], s_inv);
    qh[threadIdx.x] = o;
    a.out[h * D + threadIdx.x] = o;
  }
  ATTN_STAMP(6);
  // Fused quantizer of the attn_output mat-vec: this head's D outputs are whole
  // Q8_0 blocks (and whole Q8_K super-blocks when D % 256 == 0).
  if (a.act_kind == ACT_NONE) return;
  __syncthreads();
  const uint32_t n = a.H * D;
  if (a.act_kind == ACT_Q8_0) {
    for (uint32_t b = warp; b < D / 32; b += nw) warp_quantize_q8_0(qh[b * 32 + lane], h * (D / 32) + b, n, a.act_buf, lane);
  } else if (a.act_kind == ACT_Q8_K) {
    if constexpr (D >= 256) {
      for (uint32_t sb = warp; sb < uint32_t(D / 256); sb += nw) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = qh[sb * 256 + lane * 8 + j];
        warp_quantize_q8_k(v, h * (D / 256) + sb, n, a.act_buf, lane);
      }
    }
  } else if (a.act_kind == ACT_F16) {
    for (uint32_t e = threadIdx.x; e < D; e += blockDim.x) reinterpret_cast<uint16_t*>(a.act_buf)[h * D + e] = f2h(qh[e]);
  } else if (a.act_kind == ACT_F32) {
    for (uint32_t e = threadIdx.x; e < D; e += blockDim.x) reinterpret_cast<float*>(a.act_buf)[h * D + e] = qh[e];
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

// One element pair (gate, up): plain vectors, or the flagged exchange buffers of a row-sharded model.
struct GegluIn {
  const float *gate, *up;
  const uint2 *ll_gate, *ll_up;
  uint32_t tag;
  uint32_t* err;
  __device__ __forceinline__ float operator()(uint32_t e) const {
    if (ll_gate) return geglu(ll_waitf(ll_gate + e, tag, err), ll_waitf(ll_up + e, tag, err));
    return geglu(gate[e], up[e]);
  }
};

__global__ void geglu_act_kernel(const float* __restrict__ gate, const float* __restrict__ up, uint32_t n, int kind,
                                 uint8_t* buf, float* hidden_out, uint32_t act_stride, const uint2* ll_gate,
                                 const uint2* ll_up, LLTag lltag) {
  pdl_trigger();
  pdl_wait();
  gate += size_t(blockIdx.y) * n;  // blockIdx.y: token of a prefill batch
  up += size_t(blockIdx.y) * n;
  const GegluIn in{gate, up, ll_gate, ll_up, ll_gate ? ll_tag(lltag) : 0u, lltag.err};
  buf += size_t(blockIdx.y) * act_stride;
  if (hidden_out) hidden_out += size_t(blockIdx.y) * n;
  const int lane = threadIdx.x & 31;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  if (kind == ACT_Q8_0) {
    for (uint32_t b = gw; b < n / 32; b += nw) {
      const float v = in(b * 32 + lane);
      if (hidden_out) hidden_out[b * 32 + lane] = v;
      warp_quantize_q8_0(v, b, n, buf, lane);
    }
  } else if (kind == ACT_Q8_K) {
    for (uint32_t sb = gw; sb < n / 256; sb += nw) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint32_t e = sb * 256 + lane * 8 + i;
        v[i] = in(e);
        if (hidden_out) hidden_out[e] = v[i];
      }
      warp_quantize_q8_k(v, sb, n, buf, lane);
    }
  } else {
    const uint32_t n_pad = kind == ACT_F16 ? ((n + 7) & ~7u) : ((n + 3) & ~3u);
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pad; i += gridDim.x * blockDim.x) {
      const float v = i < n ? in(i) : 0.0f;
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
// Row-sharded model: *key covers this rank's rows only; the ranks swap keys through the flagged exchange buffer
// (two words per rank) and every rank takes the maximum — the same token everywhere.
__global__ void finish_token_kernel(unsigned long long* key, int32_t* cur_tok, int32_t* gen, int32_t* gen_count,
                                    LLCtx ll, uint32_t ll_off) {
  pdl_trigger();
  pdl_wait();
  if (threadIdx.x == 0) {
    unsigned long long best = *key;
    if (ll.peers.n) {
      const uint32_t tag = ll_tag(ll.tag);
      for (uint32_t p = 0; p < ll.peers.n; ++p) {
        ll_store(ll.peers.base[p] + ll_off + 2 * ll.rank, uint32_t(best >> 32), tag);
        ll_store(ll.peers.base[p] + ll_off + 2 * ll.rank + 1, uint32_t(best), tag);
      }
      const uint2* mine = ll.peers.base[ll.rank] + ll_off;
      best = 0ull;
      for (uint32_t r = 0; r < ll.peers.n; ++r) {
        const unsigned long long hi = ll_wait(mine + 2 * r, tag, ll.tag.err), lo = ll_wait(mine + 2 * r + 1, tag, ll.tag.err);
        const unsigned long long k = (hi << 32) | lo;
        best = k > best ? k : best;
      }
    }
    const int32_t tok = int32_t(0xffffffffu - uint32_t(best));
    *key = 0ull;
    if (cur_tok) *cur_tok = tok;
    if (gen && gen_count) {
      gen[*gen_count] = tok;
      *gen_count += 1;
    }
  }
}

__global__ void ll_unpack_kernel(const uint2* ll, LLTag lltag, float* out, uint32_t n, float softcap) {
  pdl_trigger();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float v = ll_waitf(ll + i, ll_tag(lltag), lltag.err);
  if (softcap > 0.0f) v = __fmul_rn(softcap, tanhf(__fdiv_rn(v, softcap)));  // model.cpp:1036-1041
  out[i] = v;
}

__global__ void softcap_kernel(float* logits, uint32_t n, float softcap) {
  pdl_trigger();
  pdl_wait();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) logits[i] = __fmul_rn(softcap, tanhf(__fdiv_rn(logits[i], softcap)));
}

}  // namespace

// ------------------------------------------------------------------ launchers

cudaError_t llmi_launch_embed(const EmbedArgs& a, const int32_t* token, float scale, float* h, cudaStream_t s,
                              uint32_t n_tok, const LLCtx* ll, uint32_t ll_off) {
  if (ll && ll->peers.n && n_tok != 1) return cudaErrorInvalidValue;
  return llmi_launch(embed_kernel, dim3((a.n_cols + 255) / 256, n_tok), dim3(256), 0, s, a, token, scale, h,
                     ll ? *ll : LLCtx(), ll_off);
}

cudaError_t llmi_launch_norm_act(const NormArgs& a, cudaStream_t s) {
  const int threads = a.n >= 2048 ? 1024 : 512;
  return llmi_launch(norm_act_kernel, dim3(a.n_tok ? a.n_tok : 1), dim3(threads), a.n * sizeof(float), s, a);
}

cudaError_t llmi_launch_act(const float* x, uint32_t n, int kind, uint8_t* buf, cudaStream_t s, uint32_t n_tok,
                            uint32_t act_stride) {
  const uint32_t warps = kind == ACT_Q8_0 ? n / 32 : (kind == ACT_Q8_K ? n / 256 : (n + 31) / 32);
  const uint32_t blocks = (warps + 7) / 8 ? (warps + 7) / 8 : 1;
  return llmi_launch(act_kernel, dim3(blocks, n_tok), dim3(256), 0, s, x, n, kind, buf, act_stride);
}

cudaError_t llmi_launch_rope_table(float2* table, uint32_t t_max, uint32_t D, float base, float scale, cudaStream_t s) {
  const uint64_t n = uint64_t(t_max) * (D / 2);
  rope_table_kernel<<<unsigned((n + 255) / 256), 256, 0, s>>>(table, t_max, D, base, scale);
  return cudaGetLastError();
}

// Dynamic shared memory of the attention kernel: the per-position arrays plus as
// many K/V tiles (2..ATT_MAX_BUF) as fit next to them.  0 = t_max too large.
static size_t attention_fixed_smem(uint32_t t_max, uint32_t D) {
  const size_t tp = (t_max + 15) & ~15u;
  return tp * (8 + 4 + 4 + 1) + size_t(D) * (4 + 4 + 4 + 2);
}

static uint32_t attention_nbuf(uint32_t t_max, uint32_t D) {
  const size_t fixed = attention_fixed_smem(t_max, D);
  const size_t room = 232448 - 1024;  // 227 KB per CTA minus the static part
  if (fixed + 2 * size_t(ATT_TILE_BYTES) > room) return 0;
  const size_t n = (room - fixed) / ATT_TILE_BYTES;
  return uint32_t(n > ATT_MAX_BUF ? ATT_MAX_BUF : n);
}

size_t llmi_attention_smem(uint32_t t_max, uint32_t D) {
  const uint32_t nbuf = attention_nbuf(t_max, D);
  return nbuf ? attention_fixed_smem(t_max, D) + size_t(nbuf) * ATT_TILE_BYTES : 0;
}

template <int D>
static cudaError_t attention_set_smem(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(attention_kernel<D, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(attention_kernel<D, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem));
}

cudaError_t llmi_attention_init(uint32_t t_max, uint32_t D) {
  const size_t smem = llmi_attention_smem(t_max, D);
  if (!smem) return cudaErrorInvalidValue;
  switch (D) {
    case 64: return attention_set_smem<64>(smem);
    case 128: return attention_set_smem<128>(smem);
    case 256: return attention_set_smem<256>(smem);
    case 512: return attention_set_smem<512>(smem);
    default: return cudaErrorInvalidValue;
  }
}

template <int D>
static cudaError_t attention_launch(const AttnArgs& a, uint32_t n_tok, cudaStream_t s) {
  const size_t smem = llmi_attention_smem(a.t_max, a.D);
  const uint32_t nbuf = attention_nbuf(a.t_max, a.D);
  if (n_tok <= 1 && !a.qbuf) return llmi_launch(attention_kernel<D, 0>, dim3(a.H), dim3(1024), smem, s, a, nbuf);
  if (!a.qbuf) return cudaErrorInvalidValue;
  cudaError_t e = llmi_launch(attention_kernel<D, 1>, dim3(a.H, n_tok), dim3(1024), 0, s, a, nbuf);
  if (e != cudaSuccess) return e;
  return llmi_launch(attention_kernel<D, 2>, dim3(a.H, n_tok), dim3(1024), smem, s, a, nbuf);
}

// n_tok > 1 (prefill batch): two launches, a.qbuf required (n_tok * H * D words).
cudaError_t llmi_launch_attention(const AttnArgs& a, cudaStream_t s, uint32_t n_tok) {
  switch (a.D) {
    case 64: return attention_launch<64>(a, n_tok, s);
    case 128: return attention_launch<128>(a, n_tok, s);
    case 256: return attention_launch<256>(a, n_tok, s);
    case 512: return attention_launch<512>(a, n_tok, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t llmi_launch_geglu_act(const float* gate, const float* up, uint32_t n, int kind, uint8_t* buf,
                                  float* hidden_out, cudaStream_t s, uint32_t n_tok, uint32_t act_stride,
                                  const uint2* ll_gate, const uint2* ll_up, const LLTag* tag) {
  const uint32_t warps = kind == ACT_Q8_0 ? n / 32 : (kind == ACT_Q8_K ? n / 256 : (n + 31) / 32);
  const uint32_t blocks = (warps + 3) / 4 ? (warps + 3) / 4 : 1;
  if (ll_gate && (n_tok != 1 || !ll_up || !tag)) return cudaErrorInvalidValue;
  return llmi_launch(geglu_act_kernel, dim3(blocks, n_tok), dim3(128), 0, s, gate, up, n, kind, buf, hidden_out,
                     act_stride, ll_gate, ll_up, tag ? *tag : LLTag());
}

cudaError_t llmi_launch_ll_unpack(const uint2* ll, const LLTag& tag, float* out, uint32_t n, float softcap,
                                  cudaStream_t s) {
  return llmi_launch(ll_unpack_kernel, dim3((n + 255) / 256), dim3(256), 0, s, ll, tag, out, n, softcap);
}

cudaError_t llmi_launch_finish_token(unsigned long long* key, int32_t* cur_tok, int32_t* gen, int32_t* gen_count,
                                     cudaStream_t s, const LLCtx* ll, uint32_t ll_off) {
  return llmi_launch(finish_token_kernel, dim3(1), dim3(32), 0, s, key, cur_tok, gen, gen_count, ll ? *ll : LLCtx(),
                     ll_off);
}

cudaError_t llmi_launch_softcap(float* logits, uint32_t n, float softcap, cudaStream_t s) {
  return llmi_launch(softcap_kernel, dim3((n + 255) / 256), dim3(256), 0, s, logits, n, softcap);
}
