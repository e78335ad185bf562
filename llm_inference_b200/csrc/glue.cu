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

#include <algorithm>
#include <cstdlib>

#include <mma.h>

#include "glue.h"
#include "launch.cuh"
#include "quant_device.cuh"
#include "glue_device.cuh"

namespace {
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
  TL_ENTER(3);
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
  TL_MARK(1);
  const uint32_t tag = a.ll_y ? ll_tag(a.ll_tag) : 0u;
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {
    const uint32_t i = threadIdx.x + k * blockDim.x;
    const bool ok = i < n;
    hv[k] = ok ? a.h[i] : 0.0f;
    if (a.ll_y) yv[k] = ok ? ll_waitf(a.ll_y + i, tag, a.ll_tag.err) : 0.0f;  // one decode token, grid = 1
    else yv[k] = (ok && a.y) ? a.y[i] : 0.0f;
  }
  if (a.pos_inc && threadIdx.x == 0 && blockIdx.x == 0) *a.pos_inc += int32_t(a.pos_inc_by ? a.pos_inc_by : gridDim.x);
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
  if (!a.w) {
    TL_MARK(2);
    return;
  }
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
  TL_MARK(2);
}

// ---- the same stage on a thread-block CLUSTER of 8 CTAs (distributed shared memory) ----------------------------------
// norm_act_kernel is ONE CTA of T = 512 / 1024 threads: on the timeline (tools/step_timeline.py) it holds the decode
// step for 3.0 us (E = 1152) to 6.0 us (E = 5376) — issue-bound on a single SM (two reductions, IEEE divisions, the
// quantizer: ~19k warp-instructions on one SM), 13 us of every 75 us gemma-3-27b layer.  Here the T LOGICAL threads
// of that kernel are spread over 8 CTAs of T/8 threads: thread t still owns elements t, t + T, ..., a warp is still
// the same 32 logical threads with the same xor-shuffle tree, and the T/32 warp sums — written by every warp into
// the shared memory of ALL 8 CTAs (mapa + st.shared::cluster) — are still added left to right by every thread after
// a cluster barrier.  Same partial sums, same tree, same order: bit-identical to norm_act_kernel (and to the persistent
// kernel's norm_sum_sq).  A logical warp holds exactly one 32-element block per pass, so Q8_0 / F16 / F32 activations
// are emitted straight from registers; Q8_K (256-element super-blocks span 8 warps) goes through the fp32 copy.
__device__ __forceinline__ uint32_t cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void dsmem_store_all(float* local, float v) {  // *local = v in every CTA of the cluster
  const uint32_t a = smem_u32(local);
#pragma unroll
  for (uint32_t r = 0; r < 8; ++r) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(a), "r"(r));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(ra), "f"(v) : "memory");
  }
}
constexpr int NORM_CL = 8;
template <int T>  // logical threads of norm_act_kernel: 512 or 1024
__global__ void __launch_bounds__(T / NORM_CL) norm_act_cluster_kernel(NormArgs a) {
  pdl_trigger();
  TL_ENTER(3);
  constexpr int TC = T / NORM_CL, NW = T / 32;
  __shared__ float red1[NW], red2[NW];
  const uint32_t n = a.n;
  const uint32_t rank = cluster_rank(), tok = blockIdx.x / NORM_CL;
  const uint32_t lt = rank * TC + threadIdx.x;  // logical thread
  const int lane = threadIdx.x & 31, lwarp = int(lt >> 5);
  {
    const size_t off = size_t(tok) * n;
    if (a.y) a.y += off;
    a.h += off;
    if (a.xn_out) a.xn_out += off;
    if (a.act_buf) a.act_buf += size_t(tok) * a.act_stride;
  }
  float yv[NORM_PER], hv[NORM_PER], wp[NORM_PER], wn[NORM_PER];
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {  // static inputs: under the predecessor's tail
    const uint32_t i = lt + k * T;
    const bool ok = i < n;
    wp[k] = (ok && (a.y || a.ll_y) && a.w_post) ? a.w_post[i] : 0.0f;
    wn[k] = (ok && a.w) ? a.w[i] : 0.0f;
  }
  cluster_arrive();  // every CTA of the cluster runs (its shared memory exists) before anyone stores into it
  pdl_wait();
  TL_MARK(1);
  const uint32_t tag = a.ll_y ? ll_tag(a.ll_tag) : 0u;
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {
    const uint32_t i = lt + k * T;
    const bool ok = i < n;
    hv[k] = ok ? a.h[i] : 0.0f;
    if (a.ll_y) yv[k] = ok ? ll_waitf(a.ll_y + i, tag, a.ll_tag.err) : 0.0f;
    else yv[k] = (ok && a.y) ? a.y[i] : 0.0f;
  }
  if (a.pos_inc && lt == 0 && tok == 0) *a.pos_inc += int32_t(a.pos_inc_by ? a.pos_inc_by : gridDim.x / NORM_CL);
  cluster_wait();
  TL_MARK(3);
  if (a.y || a.ll_y) {
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) ss += __fmul_rn(yv[k], yv[k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) dsmem_store_all(&red1[lwarp], ss);
    cluster_arrive();
    cluster_wait();
    if (__float_as_uint(yv[0]) == 0x7fc12345u) return;  // (timeline builds: makes stamp 4 wait for the loads)
    TL_MARK(4);
    float tot = 0.0f;
#pragma unroll
    for (int i = 0; i < NW; ++i) tot += red1[i];
    const float sc = rms_scale(tot, n, a.eps);
    if (sc == 12345.678f) return;
    TL_MARK(5);
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) {
      const uint32_t i = lt + k * T;
      const float add = a.w_post ? __fmul_rn(__fmul_rn(sc, yv[k]), wp[k]) : yv[k];
      hv[k] = __fadd_rn(hv[k], add);
      if (i < n) a.h[i] = hv[k];
    }
    if (a.epoch_inc && lt == 0) *a.epoch_inc += 1u;  // (every thread of the cluster read the epoch before the barrier)
  }
  if (!a.w) {
    TL_MARK(2);
    return;
  }
  float ss = 0.0f;
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) ss += __fmul_rn(hv[k], hv[k]);
#pragma unroll
  for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  if (lane == 0) dsmem_store_all(&red2[lwarp], ss);
  cluster_arrive();
  cluster_wait();
  TL_MARK(6);
  float tot = 0.0f;
#pragma unroll
  for (int i = 0; i < NW; ++i) tot += red2[i];
  const float sc = rms_scale(tot, n, a.eps);
  if (sc == 12345.678f) return;
  TL_MARK(7);
  float xv[NORM_PER];
#pragma unroll
  for (int k = 0; k < NORM_PER; ++k) {
    const uint32_t i = lt + k * T;
    xv[k] = __fmul_rn(__fmul_rn(sc, hv[k]), wn[k]);
    if (i < n && a.xn_out) a.xn_out[i] = xv[k];
  }
  if (a.act_kind == ACT_Q8_0) {  // pass k of logical warp w = block k * T/32 + w (n is a multiple of 32)
    bool live[NORM_PER];
    uint32_t blk[NORM_PER];
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) {
      live[k] = uint32_t(lwarp) * 32u + uint32_t(k) * T < n;
      blk[k] = uint32_t(k) * NW + lwarp;
    }
    warp_quantize_q8_0_multi<NORM_PER>(xv, live, blk, n, a.act_buf, lane);
  } else if (a.act_kind == ACT_F16) {
    const uint32_t n_pad = (n + 7) & ~7u;
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) {
      const uint32_t i = lt + k * T;
      if (i < n_pad) reinterpret_cast<uint16_t*>(a.act_buf)[i] = i < n ? f2h(xv[k]) : uint16_t(0);
    }
  } else if (a.act_kind == ACT_F32) {
    const uint32_t n_pad = (n + 3) & ~3u;
#pragma unroll
    for (int k = 0; k < NORM_PER; ++k) {
      const uint32_t i = lt + k * T;
      if (i < n_pad) reinterpret_cast<float*>(a.act_buf)[i] = i < n ? xv[k] : 0.0f;
    }
  } else if (a.act_kind == ACT_Q8_K) {  // super-blocks span 8 logical warps: through the fp32 copy (launcher: xn_out set)
    cluster_arrive();
    cluster_wait();  // release / acquire at cluster scope: the xn_out stores of all 8 CTAs are visible
    const uint32_t gw = lt >> 5;
    for (uint32_t sb = gw; sb < n / 256; sb += NW) {
      float v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = __ldcg(a.xn_out + sb * 256 + lane * 8 + i);
      warp_quantize_q8_k(v, sb, n, a.act_buf, lane);
    }
  }
  TL_MARK(2);
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

template <int D, int MODE>
__global__ void __launch_bounds__(1024) attention_kernel(AttnArgs a, uint32_t nbuf) {
  extern __shared__ __align__(128) uint8_t att_smem[];
  TL_ENTER(4);
#ifdef LLMI_TIMELINE
  if (tl_slot >= 0) g_tl_attn_slot = tl_slot;
#endif
  // (MODE 2 of a row-sharded batch: this rank's heads only, AttnArgs::hk_begin)
  attention_body<D, MODE, false>(a, nbuf, blockIdx.x + (MODE == 2 ? a.hk_begin * (a.H / a.HK) : 0u), blockIdx.y, att_smem, nullptr);
  TL_MARK(2);
}

// ---- throughput prefill (opt-in, llmi_set_prefill_mode(1)): attention of a token batch without the reference's
// position-by-position fp16 recurrence ------------------------------------------------------------------------------
// attention_kernel<D, 2> reproduces model.cpp:476-550 operation for operation — an fp16 accumulator rounded after
// every cached position — one CTA per (head, token) walking the whole context: 1.1 of the 1.33 s of a 2048-token
// gemma-3-27b prompt once the mat-vecs run on the tensor cores.  This kernel computes the same softmax(q.K)V with an
// fp32 online softmax: CTA = one KV head x QB consecutive tokens x the G query heads that share the KV head (16
// warps, one per (head, token)); K / V tiles of 32 positions go through shared memory once for all of them; a lane
// owns a POSITION for the scores (q broadcast from shared memory, K rows padded to D + 1 words: conflict-free) and
// a run of D/32 ELEMENTS for the value accumulation.  Inputs are what the exact prologue (MODE 1) left: f16(q),
// f16(k) as double high words, f16 values; the output feeds the generic quantizer.  Not bit-exact (the reference's
// fp16 recurrence is itself ~1e-3 from the arithmetic it approximates); covered by the fast mode's stated tolerance.
__device__ __forceinline__ float double_hi_to_float(uint32_t hi) {  // hi = high word of double(x), x an f16 value
  const uint32_t e = (hi >> 20) & 0x7ffu;
  return e == 0 ? 0.0f : __uint_as_float((hi & 0x80000000u) | ((e - 896u) << 23) | ((hi & 0xfffffu) << 3));
}
// FA_QPW = queries (consecutive tokens) per warp.  4 was measured and is slower (gemma-3-27b 2048-token prompt 426 -> 470 ms,
// 1b 69 -> 91 ms): the kernel is bound by the per-warp score / value loops, not by the tile loads, and fewer, fatter
// CTAs lose more in parallelism than they save in tile traffic.
constexpr int FA_WARPS = 16, FA_TILE = 32, FA_QPW = 1;  // warps per CTA, positions per tile, queries per warp
template <int D>
__global__ void __launch_bounds__(FA_WARPS * 32) fast_attention_kernel(AttnArgs a, uint32_t n_tok, uint32_t qb_tokens) {
  extern __shared__ __align__(16) uint8_t fa_smem[];
  constexpr int KS = D + 1, EPL = D / 32;  // padded K row (words); elements per lane
  float* kt = reinterpret_cast<float*>(fa_smem);                                   // [32][D + 1]
  __half* vt = reinterpret_cast<__half*>((reinterpret_cast<uintptr_t>(kt + FA_TILE * KS) + 15) & ~uintptr_t(15));  // [32][D]
  float* qs = reinterpret_cast<float*>(vt + FA_TILE * D);                          // [FA_WARPS][FA_QPW][D]
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t hkv = blockIdx.x, G = a.H / a.HK;
  const uint32_t g = warp % G, tw = warp / G;  // query head within the KV group; this warp's run of FA_QPW tokens
  const uint32_t h = hkv * G + g;
  const uint32_t tok0 = blockIdx.y * qb_tokens + tw * FA_QPW;
  pdl_wait();
  const int pos0 = *a.pos;  // position of the batch's first token
  const int P_max = pos0 + int(min(n_tok, (blockIdx.y + 1) * qb_tokens)) - 1;  // the CTA's last token attends to 0..P_max
  int P[FA_QPW];      // query qi attends to positions 0..P[qi]; -1: no such token
  float m[FA_QPW], l[FA_QPW], acc[FA_QPW][EPL];
#pragma unroll
  for (int qi = 0; qi < FA_QPW; ++qi) {
    const uint32_t tok = tok0 + qi;
    P[qi] = tok < n_tok ? pos0 + int(tok) : -1;
    m[qi] = -INFINITY;
    l[qi] = 0.0f;
#pragma unroll
    for (int e = 0; e < EPL; ++e) acc[qi][e] = 0.0f;
    if (P[qi] >= 0) {
      const uint32_t* q = a.qbuf + (size_t(tok) * a.H + h) * D;
      for (int i = lane; i < D; i += 32) qs[(warp * FA_QPW + qi) * D + i] = double_hi_to_float(q[i]);
    }
  }
  const uint32_t* kc = a.kcache + size_t(hkv) * a.t_max * D;
  const __half* vc = a.vcache + size_t(hkv) * a.t_max * D;
  for (int t0 = 0; t0 <= P_max; t0 += FA_TILE) {
    const int nt = min(FA_TILE, P_max + 1 - t0);
    __syncthreads();  // the previous tile has been consumed (and qs is written)
    for (int i = threadIdx.x; i < nt * D / 4; i += FA_WARPS * 32) {
      const int r = (i * 4) / D, c = i * 4 - r * D;
      const uint4 w = *reinterpret_cast<const uint4*>(kc + size_t(t0 + r) * D + c);
      float* dst = kt + r * KS + c;
      dst[0] = double_hi_to_float(w.x);
      dst[1] = double_hi_to_float(w.y);
      dst[2] = double_hi_to_float(w.z);
      dst[3] = double_hi_to_float(w.w);
    }
    for (int i = threadIdx.x; i < nt * D / 8; i += FA_WARPS * 32)
      reinterpret_cast<uint4*>(vt)[i] = reinterpret_cast<const uint4*>(vc + size_t(t0) * D)[i];
    __syncthreads();
    const float* kr = kt + min(lane, nt - 1) * KS;  // scores: lane = position t0 + lane
#pragma unroll
    for (int qi = 0; qi < FA_QPW; ++qi) {
      if (t0 > P[qi]) continue;  // warp-uniform
      float sc = 0.0f;
      const float4* q4 = reinterpret_cast<const float4*>(qs + (warp * FA_QPW + qi) * D);
#pragma unroll 8
      for (int i = 0; i < D / 4; ++i) {
        const float4 qv = q4[i];
        sc = fmaf(qv.x, kr[4 * i], sc);
        sc = fmaf(qv.y, kr[4 * i + 1], sc);
        sc = fmaf(qv.z, kr[4 * i + 2], sc);
        sc = fmaf(qv.w, kr[4 * i + 3], sc);
      }
      if (a.softcap > 0.0f) sc = a.softcap * tanhf(sc / a.softcap);
      const bool ok = t0 + lane <= P[qi];
      float tm = ok ? sc : -INFINITY;
#pragma unroll
      for (int o = 16; o; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
      const float m_new = fmaxf(m[qi], tm);
      const float p = ok ? __expf(sc - m_new) : 0.0f;
      const float corr = __expf(m[qi] - m_new);  // m = -inf on the first tile: 0
      float ps = p;
#pragma unroll
      for (int o = 16; o; o >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, o);
      l[qi] = l[qi] * corr + ps;
      m[qi] = m_new;
#pragma unroll
      for (int e = 0; e < EPL; ++e) acc[qi][e] *= corr;
      // values: lane = elements EPL * lane .. + EPL - 1
      const int jn = min(nt, P[qi] - t0 + 1);
      for (int j = 0; j < jn; ++j) {
        const float pj = __shfl_sync(0xffffffffu, p, j);
        const __half* vr = vt + j * D + lane * EPL;
#pragma unroll
        for (int e = 0; e < EPL; e += 2) {
          const float2 v = __half22float2(*reinterpret_cast<const __half2*>(vr + e));
          acc[qi][e] = fmaf(pj, v.x, acc[qi][e]);
          acc[qi][e + 1] = fmaf(pj, v.y, acc[qi][e + 1]);
        }
      }
    }
  }
#pragma unroll
  for (int qi = 0; qi < FA_QPW; ++qi) {
    if (P[qi] < 0) continue;
    const float inv = l[qi] > 0.0f ? 1.0f / l[qi] : 0.0f;
    float* o = a.out + size_t(tok0 + qi) * a.H * D + size_t(h) * D + lane * EPL;
#pragma unroll
    for (int e = 0; e < EPL; ++e) o[e] = acc[qi][e] * inv;
  }
}

// The prologue of a token batch in the throughput mode: q/k RMSNorm, RoPE, q scaling, f16 rounding and the K/V append of
// attention_body's MODE 1 — the same operations in the same order (two pair sums per lane, the xor-shuffle tree per 32
// pairs, the 32-pair sums left to right), but one WARP per (token, head) instead of one 1024-thread CTA: MODE 1 was
// 461 us per 1024-token gemma-3-27b layer batch, 10 % of the fast prompt (profiles/r02_notes.md).
template <int D>
__global__ void __launch_bounds__(256) fast_attn_prologue_kernel(AttnArgs a, uint32_t n_tok) {
  constexpr int HALF = D / 2, NC = (HALF + 31) / 32;
  pdl_trigger();
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const uint32_t slots = a.H + a.HK, tok = gw / slots, slot = gw % slots;
  pdl_wait();
  if (tok >= n_tok) return;
  const int pos = *a.pos + int(tok);
  const bool is_q = slot < a.H;
  const uint32_t head = is_q ? slot : slot - a.H;
  if (is_q && a.hk_count) {  // row-sharded batch: only this rank's query heads are attended to (and only their q rows exist here)
    const uint32_t hk = head / (a.H / a.HK);
    if (hk < a.hk_begin || hk >= a.hk_begin + a.hk_count) return;
  }
  const float* x = is_q ? a.q + (size_t(tok) * a.H + head) * D : a.k + (size_t(tok) * a.HK + head) * D;
  const float* wn = is_q ? a.wq_norm : a.wk_norm;
  float x0[NC], x1[NC];
  float ss = 0.0f;
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int i = int(lane) + 32 * c;
    x0[c] = i < HALF ? x[i] : 0.0f;
    x1[c] = i < HALF ? x[i + HALF] : 0.0f;
    float sq = __fadd_rn(__fmul_rn(x0[c], x0[c]), __fmul_rn(x1[c], x1[c]));
#pragma unroll
    for (int o = 16; o; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
    ss += sq;
  }
  const float sc = rms_scale(ss, D, a.eps);
#pragma unroll
  for (int c = 0; c < NC; ++c) {
    const int i = int(lane) + 32 * c;
    if (i >= HALF) continue;
    const float v0 = __fmul_rn(__fmul_rn(sc, x0[c]), wn[i]), v1 = __fmul_rn(__fmul_rn(sc, x1[c]), wn[i + HALF]);
    const float2 csn = a.rope_table[size_t(pos) * HALF + i];
    const float cs = csn.x, sn = csn.y;
    if (is_q) {
      const float qa = __fmul_rn(__fmaf_rn(v0, cs, -__fmul_rn(v1, sn)), a.attn_scale);
      const float qb = __fmul_rn(__fmaf_rn(v0, sn, __fmul_rn(v1, cs)), a.attn_scale);
      uint32_t* qd = a.qbuf + (size_t(tok) * a.H + head) * D;
      qd[i] = f16_as_double_hi(__float2half_rn(qa));
      qd[i + HALF] = f16_as_double_hi(__float2half_rn(qb));
    } else {
      const __half ka = __float2half_rn(__fmaf_rn(v0, cs, -__fmul_rn(v1, sn)));
      const __half kb = __float2half_rn(__fmaf_rn(v0, sn, __fmul_rn(v1, cs)));
      uint32_t* kd = a.kcache + (size_t(head) * a.t_max + pos) * D;
      __half* vd = a.vcache + (size_t(head) * a.t_max + pos) * D;
      const float* v = a.v + (size_t(tok) * a.HK + head) * D;
      kd[i] = f16_as_double_hi(ka);
      kd[i + HALF] = f16_as_double_hi(kb);
      vd[i] = __float2half_rn(v[i]);
      vd[i + HALF] = __float2half_rn(v[i + HALF]);
    }
  }
}

// ---- the same attention on the tensor cores (f16 mma through nvcuda::wmma, fp32 accumulation) ---------------------------
// The CUDA-core kernel above was 52 % of a 1024-token gemma-3-27b batch once the mat-vecs ran on tcgen05
// (profiles/r02_notes.md).  Here a CTA of 4 warps owns 64 query rows = 64/G consecutive tokens x the G query heads of one
// KV head, a warp 16 of them; K / V tiles of 64 positions go through shared memory as f16 (q and k ARE f16 values:
// the exact prologue rounded them; fast_attn_prep_kernel turns their double-high-word storage into halves once per
// batch).  Two passes over the KV tiles, so the opaque wmma accumulator layout never has to be rescaled: pass 1
// S = Q.K^T -> row maxima; pass 2 S again, P = exp(S - max) (f16), row sums, O += P.V in D/16 accumulator fragments
// that live across the tiles; O / sum -> out.  Same tolerance class as the kernel above (tests cover both).
__global__ void fast_attn_prep_kernel(AttnArgs a, uint32_t n_tok, __half* __restrict__ qh, __half* __restrict__ kh) {
  pdl_trigger();
  pdl_wait();
  const int T = *a.pos + int(n_tok);  // positions in the cache after this batch's prologue
  const uint64_t nq = uint64_t(n_tok) * a.H * a.D, nk = uint64_t(a.HK) * a.t_max * a.D;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < nq + nk; i += uint64_t(gridDim.x) * blockDim.x) {
    if (i < nq) {
      qh[i] = __float2half_rn(double_hi_to_float(a.qbuf[i]));
    } else {
      const uint64_t j = i - nq;
      const uint32_t t = uint32_t((j / a.D) % a.t_max);
      if (int(t) < T) kh[j] = __float2half_rn(double_hi_to_float(a.kcache[j]));
    }
  }
}

template <int D>
__global__ void __launch_bounds__(128) fast_attention_tc_kernel(AttnArgs a, uint32_t n_tok, const __half* __restrict__ qh,
                                                                const __half* __restrict__ kh) {
  using namespace nvcuda;
  constexpr int LD = D + 8, SLD = 68, PLD = 72, OLD = D + 4;  // leading dimensions: halves, floats, halves, floats
  extern __shared__ __align__(128) uint8_t fat_smem[];
  __half* Qs = reinterpret_cast<__half*>(fat_smem);        // [64][LD]
  __half* Ks = Qs + 64 * LD;                               // [64][LD]
  __half* Vs = Ks + 64 * LD;                               // [64][LD]
  float* Ss = reinterpret_cast<float*>(Vs + 64 * LD);      // [4 warps][16][SLD]
  __half* Ps = reinterpret_cast<__half*>(Ss + 4 * 16 * SLD);  // [4 warps][16][PLD]
  pdl_trigger();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t hkv = blockIdx.x + a.hk_begin, G = a.H / a.HK, TPB = 64 / G;
  const uint32_t tokb = blockIdx.y * TPB;
  pdl_wait();
  const int pos0 = *a.pos;
  const int n_here = int(min(TPB, n_tok - tokb));
  const int P_max = pos0 + int(tokb) + n_here - 1;
  for (int i = threadIdx.x; i < 64 * D / 8; i += 128) {  // row r = (token in block) * G + (head in group)
    const int r = i / (D / 8), c = (i % (D / 8)) * 8;
    const int tl = r / int(G), g = r % int(G);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tl < n_here) v = *reinterpret_cast<const uint4*>(qh + (size_t(tokb + tl) * a.H + hkv * G + g) * D + c);
    *reinterpret_cast<uint4*>(Qs + r * LD + c) = v;
  }
  const int rr = lane >> 1, hc = lane & 1;  // this lane's row of the warp tile and its half of the 64 columns
  const int tl_row = (warp * 16 + rr) / int(G);
  const int P_row = tl_row < n_here ? pos0 + int(tokb) + tl_row : -1;
  float* Sw = Ss + warp * 16 * SLD;
  __half* Pw = Ps + warp * 16 * PLD;
  const __half* kbase = kh + size_t(hkv) * a.t_max * D;
  const __half* vbase = a.vcache + size_t(hkv) * a.t_max * D;
  auto load_tile = [&](const __half* base, __half* dst, int t0, int nt) {
    for (int i = threadIdx.x; i < 64 * D / 8; i += 128) {
      const int r = i / (D / 8), c = (i % (D / 8)) * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (r < nt) v = *reinterpret_cast<const uint4*>(base + size_t(t0 + r) * D + c);
      *reinterpret_cast<uint4*>(dst + r * LD + c) = v;
    }
  };
  auto scores = [&]() {  // Sw[16][64] = Q_w (16 x D) . K^T (D x 64)
    wmma::fragment<wmma::accumulator, 16, 16, 16, float> acc[4];
#pragma unroll
    for (int n = 0; n < 4; ++n) wmma::fill_fragment(acc[n], 0.0f);
#pragma unroll 4
    for (int k = 0; k < D / 16; ++k) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> af;
      wmma::load_matrix_sync(af, Qs + warp * 16 * LD + k * 16, LD);
#pragma unroll
      for (int n = 0; n < 4; ++n) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::col_major> bf;  // (k, n) = Ks[n][k]
        wmma::load_matrix_sync(bf, Ks + n * 16 * LD + k * 16, LD);
        wmma::mma_sync(acc[n], af, bf, acc[n]);
      }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) wmma::store_matrix_sync(Sw + n * 16, acc[n], SLD, wmma::mem_row_major);
    __syncwarp();
  };
  auto sval = [&](float sc) { return a.softcap > 0.0f ? a.softcap * tanhf(sc / a.softcap) : sc; };
  // pass 1: row maxima
  float m_row = -INFINITY;
  for (int t0 = 0; t0 <= P_max; t0 += 64) {
    const int nt = min(64, P_max + 1 - t0);
    __syncthreads();
    load_tile(kbase, Ks, t0, nt);
    __syncthreads();
    if (t0 > pos0 + int(tokb) + (warp * 16 + 15) / int(G)) continue;  // no row of this warp reaches the tile (warp-uniform)
    scores();
    float mx = -INFINITY;
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const int col = hc * 32 + c;
      if (t0 + col <= P_row) mx = fmaxf(mx, sval(Sw[rr * SLD + col]));
    }
    mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
    m_row = fmaxf(m_row, mx);
    __syncwarp();
  }
  // pass 2: P = exp(S - max), row sums, O += P.V
  wmma::fragment<wmma::accumulator, 16, 16, 16, float> oacc[D / 16];
#pragma unroll
  for (int n = 0; n < D / 16; ++n) wmma::fill_fragment(oacc[n], 0.0f);
  float l_row = 0.0f;
  for (int t0 = 0; t0 <= P_max; t0 += 64) {
    const int nt = min(64, P_max + 1 - t0);
    __syncthreads();
    load_tile(kbase, Ks, t0, nt);
    load_tile(vbase, Vs, t0, nt);
    __syncthreads();
    if (t0 > pos0 + int(tokb) + (warp * 16 + 15) / int(G)) continue;
    scores();
    float ls = 0.0f;
#pragma unroll 8
    for (int c = 0; c < 32; ++c) {
      const int col = hc * 32 + c;
      const float p = t0 + col <= P_row ? __expf(sval(Sw[rr * SLD + col]) - m_row) : 0.0f;
      ls += p;
      Pw[rr * PLD + col] = __float2half_rn(p);
    }
    ls += __shfl_xor_sync(0xffffffffu, ls, 1);
    l_row += ls;
    __syncwarp();
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      wmma::fragment<wmma::matrix_a, 16, 16, 16, __half, wmma::row_major> af;
      wmma::load_matrix_sync(af, Pw + kk * 16, PLD);
#pragma unroll
      for (int n = 0; n < D / 16; ++n) {
        wmma::fragment<wmma::matrix_b, 16, 16, 16, __half, wmma::row_major> bf;  // (k, n) = Vs[k][n]
        wmma::load_matrix_sync(bf, Vs + kk * 16 * LD + n * 16, LD);
        wmma::mma_sync(oacc[n], af, bf, oacc[n]);
      }
    }
    __syncwarp();
  }
  __syncthreads();  // every warp is done with K / V: their space takes the fp32 output tiles
  float* Ow = reinterpret_cast<float*>(Ks) + warp * 16 * OLD;
#pragma unroll
  for (int n = 0; n < D / 16; ++n) wmma::store_matrix_sync(Ow + n * 16, oacc[n], OLD, wmma::mem_row_major);
  __syncwarp();
  if (P_row >= 0) {
    const float inv = l_row > 0.0f ? 1.0f / l_row : 0.0f;
    const uint32_t g = uint32_t(warp * 16 + rr) % G;
    float* o = a.out + size_t(tokb + tl_row) * a.H * D + size_t(hkv * G + g) * D + hc * (D / 2);
    const float* src = Ow + rr * OLD + hc * (D / 2);
#pragma unroll 4
    for (int c = 0; c < D / 2; c += 4) {
      const float4 v = *reinterpret_cast<const float4*>(src + c);
      *reinterpret_cast<float4*>(o + c) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
    }
  }
}

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

// One element pair (gate, up): plain vectors, or the flagged exchange buffers of a row-sharded model.
struct GegluIn {
  const float *gate, *up;
  const uint2 *ll_gate, *ll_up;
  uint32_t tag;
  uint32_t* err;
  __device__ __forceinline__ float operator()(uint32_t e) const {
    if (ll_gate && !ll_up) return ll_waitf(ll_gate + e, tag, err);  // hidden rows already combined by their producers
    if (ll_gate) return geglu(ll_waitf(ll_gate + e, tag, err), ll_waitf(ll_up + e, tag, err));
    return geglu(gate[e], up[e]);
  }
};

__global__ void geglu_act_kernel(const float* __restrict__ gate, const float* __restrict__ up, uint32_t n, int kind,
                                 uint8_t* buf, float* hidden_out, uint32_t act_stride, const uint2* ll_gate,
                                 const uint2* ll_up, LLTag lltag) {
  pdl_trigger();
  TL_ENTER(5);
  pdl_wait();
  TL_MARK(1);
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
  TL_MARK(2);
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

// ---- token batches of a row-sharded model (prefill): the all-gather of [token][row] tiles -----------
// Every rank computes its own columns (= matrix rows) of a [n_tok][stride] fp32 batch in place; the batch buffers sit
// at the same byte offset of every rank's exchange allocation.  bx_exchange_kernel copies this rank's columns into
// every peer's buffer (plain 16-byte stores over NVLink peer memory) and ends in a barrier: the last CTA to finish
// (fence + ticket) raises this rank's flag on every peer and waits for every peer's flag, so when the kernel
// completes all columns of all ranks are in place.  Consecutive exchanges use different buffers and a rank can be at
// most one exchange ahead of its slowest peer (it needs that peer's flag), so a buffer is never overwritten while a
// peer still reads it.  n_seg == 0: the barrier alone.
__global__ void __launch_bounds__(256) bx_exchange_kernel(BxArgs a) {
  pdl_trigger();
  pdl_wait();
  __shared__ uint32_t s_last;
  const uint32_t world = a.peers.n, tid = threadIdx.x;
  for (uint32_t g = 0; g < a.n_seg; ++g) {
    const BxSeg sg = a.seg[g];
    if (sg.cols == 0) continue;
    const bool vec = ((sg.col0 | sg.cols | sg.stride) & 3u) == 0;
    const uint32_t per = vec ? sg.cols / 4 : sg.cols;
    const uint64_t items = uint64_t(a.n_tok) * per;
    for (uint64_t i = uint64_t(blockIdx.x) * blockDim.x + tid; i < items; i += uint64_t(gridDim.x) * blockDim.x) {
      const uint32_t t = uint32_t(i / per), c = uint32_t(i % per);
      const uint64_t elem = uint64_t(t) * sg.stride + sg.col0 + (vec ? c * 4 : c);
      const uint64_t byte = sg.byte_off + elem * 4;
      const char* mine = reinterpret_cast<const char*>(a.peers.base[a.rank]) + byte;
      if (vec) {
        const float4 v = *reinterpret_cast<const float4*>(mine);
        for (uint32_t p = 0; p < world; ++p)
          if (p != a.rank) *reinterpret_cast<float4*>(reinterpret_cast<char*>(a.peers.base[p]) + byte) = v;
      } else {
        const float v = *reinterpret_cast<const float*>(mine);
        for (uint32_t p = 0; p < world; ++p)
          if (p != a.rank) *reinterpret_cast<float*>(reinterpret_cast<char*>(a.peers.base[p]) + byte) = v;
      }
    }
  }
  __threadfence_system();
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(a.counter, 1u) == gridDim.x - 1 ? 1u : 0u;
  __syncthreads();
  if (!s_last) return;
  if (tid == 0) *a.counter = 0;
  __threadfence_system();  // (acquire side of the ticket: every CTA's stores are ordered before the flags below)
  if (tid < world && tid != a.rank) {
    uint32_t* theirs = reinterpret_cast<uint32_t*>(a.peers.base[tid] + a.flag_off + a.rank);
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(a.seq) : "memory");
    const uint32_t* mine = reinterpret_cast<const uint32_t*>(a.peers.base[a.rank] + a.flag_off + tid);
    unsigned long long limit_ns = 4000000000ull, t0;
    bool dead = false;
    if (a.err) {
      uint32_t d, ms;
      asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(d) : "l"(a.err) : "memory");
      asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(ms) : "l"(a.err + 1) : "memory");
      dead = d != 0;
      if (ms) limit_ns = (unsigned long long)ms * 1000000ull;
    }
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (uint32_t spins = 1; !dead; ++spins) {
      uint32_t f;
      asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(f) : "l"(mine) : "memory");
      if (int32_t(f - a.seq) >= 0) break;  // a peer may already have raised the next exchange's flag
      if ((spins & 255u) == 0) {
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > limit_ns) {
          if (a.err) *a.err = 1;
          break;
        }
      }
      __nanosleep(64);
    }
  }
  __threadfence_system();
}

// embed_tokens + scale_embeddings for a token batch of a row-sharded model: the rank that holds a token's row writes
// it into the residual batch of every rank (a bx_exchange barrier follows).
__global__ void embed_shard_batch_kernel(EmbedArgs a, const int32_t* __restrict__ token, float scale, LLPeers peers,
                                         uint64_t h_byte_off) {
  pdl_trigger();
  pdl_wait();
  const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x, m = blockIdx.y;
  if (e >= a.n_cols) return;
  const uint32_t row = uint32_t(token[m]);
  if (row < a.row_begin || row >= a.row_end) return;
  const float v = dequant_elem(a, row - a.row_begin, e) * scale;
  for (uint32_t p = 0; p < peers.n; ++p)
    reinterpret_cast<float*>(reinterpret_cast<char*>(peers.base[p]) + h_byte_off)[size_t(m) * a.n_cols + e] = v;
  __threadfence_system();
}

}  // namespace

TL_EXPORT(llmi_debug_timeline_glue)

// ------------------------------------------------------------------ launchers

cudaError_t llmi_launch_embed(const EmbedArgs& a, const int32_t* token, float scale, float* h, cudaStream_t s,
                              uint32_t n_tok, const LLCtx* ll, uint32_t ll_off) {
  if (ll && ll->peers.n && n_tok != 1) return cudaErrorInvalidValue;
  return llmi_launch(embed_kernel, dim3((a.n_cols + 255) / 256, n_tok), dim3(256), 0, s, a, token, scale, h,
                     ll ? *ll : LLCtx(), ll_off);
}

// LLMI_NORM_CLUSTER=0 keeps the single-CTA kernel (A/B, re-read by every llmi_model_load); both produce the same bits.
static int g_norm_cluster = 1;
void llmi_glue_read_env() {
  const char* e = getenv("LLMI_NORM_CLUSTER");
  g_norm_cluster = (e && e[0] == '0') ? 0 : 1;
}
cudaError_t llmi_launch_norm_act(const NormArgs& a, cudaStream_t s) {
  const int threads = a.n >= 2048 ? 1024 : 512;
  const uint32_t n_tok = a.n_tok ? a.n_tok : 1;
  const bool kind_ok = a.act_kind != ACT_Q8_K || a.xn_out != nullptr;
  // the cluster pays from ~2 elements per logical thread on: at E = 1152 (one element per thread) its barriers cost more than
  // the single CTA's issue pressure (3.0 vs 3.6 us per launch on the step timeline), at E = 5376 it is 6.0 vs 3.8 us
  if (g_norm_cluster && kind_ok && a.n >= 2048 && a.n <= uint32_t(NORM_PER * threads)) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_tok * NORM_CL);
    cfg.blockDim = dim3(threads / NORM_CL);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = s;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = NORM_CL;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = g_llmi_pdl ? 2 : 1;
    return threads == 1024 ? cudaLaunchKernelEx(&cfg, norm_act_cluster_kernel<1024>, a)
                           : cudaLaunchKernelEx(&cfg, norm_act_cluster_kernel<512>, a);
  }
  return llmi_launch(norm_act_kernel, dim3(n_tok), dim3(threads), a.n * sizeof(float), s, a);
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

bool llmi_attention_batch_tc(uint32_t H, uint32_t HK, uint32_t D) {
  return llmi_gemv_prefill_fast() && D >= 64 && D <= 256 && HK && H % HK == 0 && 64 % (H / HK) == 0 && !getenv("LLMI_FAST_ATTN_NO_TC");
}

template <int D>
static cudaError_t attention_launch(const AttnArgs& a, uint32_t n_tok, cudaStream_t s) {
  const size_t smem = llmi_attention_smem(a.t_max, a.D);
  const uint32_t nbuf = attention_nbuf(a.t_max, a.D);
  if (n_tok <= 1 && !a.qbuf) return llmi_launch(attention_kernel<D, 0>, dim3(a.H), dim3(1024), smem, s, a, nbuf);
  if (!a.qbuf) return cudaErrorInvalidValue;
  cudaError_t e;
  if (llmi_gemv_prefill_fast() && D <= 256 && !getenv("LLMI_FAST_ATTN_NO_TC")) {
    const uint32_t warps = n_tok * (a.H + a.HK);
    e = llmi_launch(fast_attn_prologue_kernel<D>, dim3((warps + 7) / 8), dim3(256), 0, s, a, n_tok);
  } else {
    e = llmi_launch(attention_kernel<D, 1>, dim3(a.H, n_tok), dim3(1024), 0, s, a, nbuf);
  }
  if (e != cudaSuccess) return e;
  if constexpr (D <= 256)
  if (llmi_attention_batch_tc(a.H, a.HK, D)) {
    // throughput prefill, tensor-core form (grow-only f16 scratch for q and K)
    __half *qh = nullptr, *kh = nullptr;
    const size_t nq = size_t(n_tok) * a.H * D, nk = size_t(a.HK) * a.t_max * D;
    if ((e = llmi_stream_scratch(s, SCR_ATT_Q, nq * 2, (void**)&qh)) != cudaSuccess) return e;
    if ((e = llmi_stream_scratch(s, SCR_ATT_K, nk * 2, (void**)&kh)) != cudaSuccess) return e;
    const unsigned pb = unsigned(std::min<size_t>((nq + nk + 255) / 256, 148 * 16));
    if ((e = llmi_launch(fast_attn_prep_kernel, dim3(pb), dim3(256), 0, s, a, n_tok, qh, kh)) != cudaSuccess) return e;
    constexpr int LD = D + 8;
    const size_t fsm = size_t(3) * 64 * LD * 2 + size_t(4) * 16 * 68 * 4 + size_t(4) * 16 * 72 * 2;
    static bool optin = false;
    if (!optin) {
      if ((e = cudaFuncSetAttribute(fast_attention_tc_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fsm))) != cudaSuccess)
        return e;
      optin = true;
    }
    const uint32_t tpb = 64 / (a.H / a.HK);
    if ((e = llmi_launch(fast_attention_tc_kernel<D>, dim3(a.hk_count ? a.hk_count : a.HK, (n_tok + tpb - 1) / tpb), dim3(128), fsm, s, a, n_tok,
                         (const __half*)qh, (const __half*)kh)) != cudaSuccess)
      return e;
    if (a.act_kind == ACT_NONE) return cudaSuccess;
    return llmi_launch_act(a.out, a.H * a.D, a.act_kind, a.act_buf, s, n_tok, a.act_stride);
  }
  if (llmi_gemv_prefill_fast() && D >= 64 && D <= 256 && FA_WARPS % (a.H / a.HK) == 0) {  // CUDA-core form
    const uint32_t qb = FA_WARPS / (a.H / a.HK) * FA_QPW;  // tokens per CTA
    const size_t fsm = size_t(FA_TILE) * (D + 1) * 4 + 32 + size_t(FA_TILE) * D * 2 + size_t(FA_WARPS) * FA_QPW * D * 4;
    static bool optin = false;
    if (!optin) {
      if ((e = cudaFuncSetAttribute(fast_attention_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(fsm))) != cudaSuccess)
        return e;
      optin = true;
    }
    if ((e = llmi_launch(fast_attention_kernel<D>, dim3(a.HK, (n_tok + qb - 1) / qb), dim3(FA_WARPS * 32), fsm, s, a, n_tok,
                         qb)) != cudaSuccess)
      return e;
    if (a.act_kind == ACT_NONE) return cudaSuccess;
    return llmi_launch_act(a.out, a.H * a.D, a.act_kind, a.act_buf, s, n_tok, a.act_stride);
  }
  return llmi_launch(attention_kernel<D, 2>, dim3(a.hk_count ? a.hk_count * (a.H / a.HK) : a.H, n_tok), dim3(1024), smem, s, a, nbuf);
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
  if (ll_gate && (n_tok != 1 || !tag)) return cudaErrorInvalidValue;  // ll_up == nullptr: ll_gate carries gelu(gate)*up
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

cudaError_t llmi_launch_bx_exchange(const BxArgs& a, cudaStream_t s) {
  if (a.peers.n < 2 || a.n_seg > 3 || !a.counter) return cudaErrorInvalidValue;
  uint64_t items = 0;
  for (uint32_t g = 0; g < a.n_seg; ++g) items += uint64_t(a.n_tok) * (a.seg[g].cols / 4 + 1);
  const unsigned blocks = unsigned(std::max<uint64_t>(1, std::min<uint64_t>((items + 255) / 256, 148ull * 8)));
  return llmi_launch(bx_exchange_kernel, dim3(blocks), dim3(256), 0, s, a);
}

cudaError_t llmi_launch_embed_shard_batch(const EmbedArgs& a, const int32_t* token, float scale, const LLPeers& peers,
                                          uint64_t h_byte_off, uint32_t n_tok, cudaStream_t s) {
  return llmi_launch(embed_shard_batch_kernel, dim3((a.n_cols + 255) / 256, n_tok), dim3(256), 0, s, a, token, scale, peers,
                     h_byte_off);
}
