// Device functions of the glue between the mat-vecs: embedding-row
// dequantization, the RMSNorm reduction, the activation quantizer front-end, the
// attention body and GEGLU.  Shared by the one-launch-per-stage kernels (glue.cu)
// and the persistent decode kernel (mega.cu) so that both produce the same bits.
#pragma once

#include <cuda_fp16.h>
#include <math.h>

#include <type_traits>

#include "glue.h"
#include "launch.cuh"
#include "quant_device.cuh"

namespace {

using namespace llmi_dev;

#ifndef LLMI_H2F_DEFINED
#define LLMI_H2F_DEFINED
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
#endif
__device__ __forceinline__ uint16_t f2h(float f) { return __half_as_ushort(__float2half_rn(f)); }

// ---------------------------------------------------------------- embedding
// One element of row `row` of a repacked matrix, dequantized as the reference's
// row dequantizers do (model.cpp:251-322 -> ops.cpp:1005-1082): F16 direct,
// Q8_0 d*q, Q5_0 d*(q-16), Q6_K d*sc*q (left to right).  Plane item order:
// repack.cu.
// TM: the formats the caller is compiled for (llmi_type_bit); others never occur there and carry no code.
template <uint32_t TM = LLMI_ALL_TYPES>
__device__ float dequant_elem(const EmbedArgs& a, uint32_t row, uint32_t e) {
  const uint32_t s = row >> 3, r = row & 7;
  const uint64_t nb = a.nb;
  if (!(TM & llmi_type_bit(a.type))) return 0.0f;
  switch (a.type) {
    case LLMI_F16: {
      if (!(TM & llmi_type_bit(LLMI_F16))) return 0.0f;
      const uint64_t cell = (uint64_t(s) * nb + (e >> 3)) * 8 + r;
      return h2f(reinterpret_cast<const uint16_t*>(a.q)[cell * 8 + (e & 7)]);
    }
    case LLMI_Q8_0: {
      if (!(TM & llmi_type_bit(LLMI_Q8_0))) return 0.0f;
      const uint32_t b = e >> 5, i = e & 31;
      const uint64_t cell = (uint64_t(s) * nb + b) * 8 + r;
      const int8_t qv = reinterpret_cast<const int8_t*>(a.q)[(((uint64_t(s) * nb + b) * 2 + (i >> 4)) * 8 + r) * 16 + (i & 15)];
      return h2f(reinterpret_cast<const uint16_t*>(a.d)[cell]) * float(qv);
    }
    case LLMI_Q5_0: {
      if (!(TM & llmi_type_bit(LLMI_Q5_0))) return 0.0f;
      const uint32_t b = e >> 5, i = e & 31;
      const uint64_t cell = (uint64_t(s) * nb + b) * 8 + r;
      const uint8_t byte = a.q[cell * 16 + (i & 15)];
      const uint32_t qh = reinterpret_cast<const uint32_t*>(a.x)[cell];
      const int qv = int((i < 16 ? (byte & 0x0f) : (byte >> 4)) | (((qh >> i) & 1u) << 4));
      return h2f(reinterpret_cast<const uint16_t*>(a.d)[cell]) * float(qv - 16);
    }
    case LLMI_Q6_K: {
      if (!(TM & llmi_type_bit(LLMI_Q6_K))) return 0.0f;
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
template <uint32_t KM = 0xffu>  // KM: the kinds (1 << kind) the caller can ask for
__device__ void emit_act(int kind, const float* xs, uint32_t n, uint8_t* buf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (!(KM & (1u << kind))) return;
  if ((KM & (1u << ACT_Q8_0)) && kind == ACT_Q8_0) {
    for (uint32_t b = warp; b < n / 32; b += nw) warp_quantize_q8_0(xs[b * 32 + lane], b, n, buf, lane);
  } else if ((KM & (1u << ACT_Q8_K)) && kind == ACT_Q8_K) {
    for (uint32_t sb = warp; sb < n / 256; sb += nw) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = xs[sb * 256 + lane * 8 + i];
      warp_quantize_q8_k(v, sb, n, buf, lane);
    }
  } else if ((KM & (1u << ACT_F16)) && kind == ACT_F16) {
    const uint32_t n_pad = (n + 7) & ~7u;
    for (uint32_t i = threadIdx.x; i < n_pad; i += blockDim.x)
      reinterpret_cast<uint16_t*>(buf)[i] = i < n ? f2h(xs[i]) : uint16_t(0);
  } else if ((KM & (1u << ACT_F32)) && kind == ACT_F32) {
    const uint32_t n_pad = (n + 3) & ~3u;
    for (uint32_t i = threadIdx.x; i < n_pad; i += blockDim.x) reinterpret_cast<float*>(buf)[i] = i < n ? xs[i] : 0.0f;
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
#elif defined(LLMI_TIMELINE)  // rough phase boundaries of CTA 0 on the step timeline (tools/step_timeline.py)
static __device__ int g_tl_attn_slot;  // the running attention launch's record (set by its wrapper)
#define ATTN_STAMP(i)                                                                                      \
  do {                                                                                                     \
    if (!MEGA && !(blockIdx.x | blockIdx.y | threadIdx.x)) g_tl[g_tl_attn_slot & 8191].t[3 + (i)] = tl_now(); \
  } while (0)
#define ATTN_RAW(i, tid) do { } while (0)
#else
#define ATTN_STAMP(i) do { } while (0)
#define ATTN_RAW(i, tid) do { } while (0)
#endif

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
//
// attention_body is the whole CTA's work for query head `h` (token `tok` of a prefill batch).  MEGA = the
// persistent decode kernel's view (mega.cu): q/k/v arrive flagged under a tag VALUE, the position comes by
// value, the head's D outputs leave as flagged stores (to every rank of a head-sharded model) and the
// consumer quantizes; the barriers are invalidated on exit because the same CTA calls again next layer.
struct AttnMega {
  int pos = 0;
  uint32_t in_tag = 0;                // tag of the flagged q/k/v rows
  uint32_t out_tag = 0, out_off = 0;  // flagged output vector [H*D] at out_off of every buffer in out_peers
  LLPeers out_peers;
  uint32_t* err = nullptr;
};

template <int D, int MODE, bool MEGA>
__device__ __forceinline__ void attention_body(AttnArgs a, const uint32_t nbuf, const uint32_t h, const uint32_t tok,
                                               uint8_t* smraw, const AttnMega* mg) {
  if (!MEGA) pdl_trigger();
  constexpr int HALF = D / 2, VEC = D / 32;           // elements per lane in phase 1
  constexpr int RTK = ATT_TILE_BYTES / (4 * D);       // K rows per tile (4 bytes per element)
  constexpr int RTV = ATT_TILE_BYTES / (2 * D);       // V rows per tile (f16)
  constexpr int PIECES = VEC >= 4 ? VEC / 4 : 1, PW = VEC >= 4 ? 4 : VEC;  // 16-byte pieces per lane / words per piece
  __shared__ float red[32];
  __shared__ float wmax[32];
  __shared__ float s_inv;
  __shared__ __align__(8) uint64_t bars[ATT_MAX_BUF];
  const uint32_t group = a.H / a.HK, hkv = h / group;
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
  if (!MEGA) pdl_wait();
#ifdef LLMI_TIMELINE
  if (!MEGA && !(blockIdx.x | blockIdx.y | threadIdx.x)) g_tl[g_tl_attn_slot & 8191].t[1] = tl_now();  // dev only
#endif
  const int pos = MEGA ? mg->pos : *a.pos + int(tok), T = pos + 1;  // tok: token of a prefill batch (0 when decoding)
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
    const uint32_t tag = MEGA ? mg->in_tag : ll_tag(a.ll_tag);
    uint32_t* const err = MEGA ? mg->err : a.ll_tag.err;
    const uint2* const src[6] = {a.ll_q + h * D + i,   a.ll_q + h * D + i + HALF,   a.ll_k + hkv * D + i,
                                 a.ll_k + hkv * D + i + HALF, a.ll_v + hkv * D + i, a.ll_v + hkv * D + i + HALF};
    const bool all6[6] = {true, true, true, true, true, true};
    uint32_t w6[6];
    ll_wait_many<6>(src, all6, tag, err, w6);
    q0 = __uint_as_float(w6[0]);
    q1 = __uint_as_float(w6[1]);
    k0 = __uint_as_float(w6[2]);
    k1 = __uint_as_float(w6[3]);
    v0 = __uint_as_float(w6[4]);
    v1 = __uint_as_float(w6[5]);
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
      if (MEGA) {  // the row is read by a LATER step of the same kernel through bulk copies (async proxy), by
                   // any CTA: make it visible device-wide before this CTA's flagged outputs can be observed
        __threadfence();
        asm volatile("fence.proxy.async;" ::: "memory");
      }
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
      float wm = lane < nw ? wmax[lane] : -INFINITY;  // one entry per warp
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
    if (MEGA) {
      for (uint32_t p = 0; p < mg->out_peers.n; ++p)
        ll_store(mg->out_peers.base[p] + mg->out_off + h * D + threadIdx.x, __float_as_uint(o), mg->out_tag);
    } else {
      a.out[h * D + threadIdx.x] = o;
    }
  }
  ATTN_STAMP(6);
  // Fused quantizer of the attn_output mat-vec: this head's D outputs are whole
  // Q8_0 blocks (and whole Q8_K super-blocks when D % 256 == 0).
  if (MEGA) {  // every copy into the ring has completed and was waited for: retire the barriers
    __syncthreads();
    if (threadIdx.x == 0)
      for (uint32_t b = 0; b < nbuf; ++b) mbar_inval(&bars[b]);
    return;
  }
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

}  // namespace
