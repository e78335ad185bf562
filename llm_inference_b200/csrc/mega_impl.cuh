// Persistent decode kernel for sm_100a: ONE cooperative launch runs n_steps tokens of the whole device-resident
// forward (Model::forward, model.cpp:706-1048) — embedding, per layer the 7 mat-vecs with their norms, attention
// and GEGLU, then the logits mat-vec and the greedy argmax (main.cpp:172-221).
//
// Why: a decode token is a chain of ~8 dependent stages per layer.  As separate launches (model.cu run_step, kept
// as the reference path) every stage pays launch + ramp + drain, ~3 us each, and the glue stages run on one CTA
// while 147 SMs and all of HBM idle: gemma-3-1b spent 0.83 ms per token on 0.15 ms of HBM time.  Here one CTA per
// SM stays resident and walks the same program (mega.h):
//   * dataflow instead of barriers: a vector produced by one phase travels as flagged {value bits, tag} words
//     (launch.cuh) in an L2-resident exchange buffer; a consumer takes exactly the elements it needs and checks
//     their tags.  There is no grid-wide barrier and no launch between phases, and on a row-sharded model the
//     very same stores go to every rank's buffer over NVLink peer memory — the all-gather is the mat-vec's
//     epilogue.  So that 150k threads do not spin on the same L2 lines, every CTA bumps a per-exchange
//     arrival counter after its stores and ONE thread per consumer CTA sleeps on that counter first: a hint only —
//     correctness rests on the tags;
//   * every CTA keeps the residual stream h in shared memory and applies the (cheap, E-element) RMSNorm +
//     residual + quantizer itself as the PROLOGUE of the mat-vec that consumes it, with the arithmetic and the
//     reduction tree of norm_act_kernel, so the norms cost no stage of their own;
//   * GEGLU is the epilogue of the gate/up phase (a CTA owns the same slab of both matrices);
//   * the weights of a phase depend on nothing: each CTA requests the first part of its share into L2
//     (cp.async.bulk.prefetch.L2) BEFORE it waits for the phase's input and keeps requesting a fixed distance
//     ahead of its loads, so DRAM streams through the waits and the demand loads hit L2;
//   * rows are dealt to CTAs by slab and every row is summed in the canonical order of gemv_slab_kernel
//     (gemv_bodies.cuh): the bits do not depend on the grid, and equal the per-launch path's bits (tested).
// All per-CTA state (arguments, the current program entry, the matrices' planes) sits in file-scope shared
// variables: the helpers below take no pointers, which keeps the 64-register budget of a 1024-thread CTA for the
// weight fragments.
#pragma once

#include <cuda_fp16.h>

#include <algorithm>

#include "gemv_bodies.cuh"
#include "glue_device.cuh"
#include "mega.h"

#ifndef MEGA_HOT  // how the hot bodies are attached to the kernel: inlined (specialized TUs) or out of line (mega_any.cu)
#define MEGA_HOT __forceinline__
#endif

namespace {

constexpr int MEGA_WARPS = MEGA_THREADS / 32;
constexpr int NORM_PER = 6;                      // elements per thread of the norm stage (norm_act_kernel)
constexpr uint32_t PREFETCH_BYTES = 160 * 1024;  // of a CTA's share requested into L2 at phase start
constexpr uint32_t PREFETCH_AHEAD = 6;           // rounds (one item per warp) the rolling prefetch runs ahead of the loads

struct PhaseRun {  // per-CTA view of the current entry
  uint32_t my_count, SB, sub, in_tag, out_tag;
};

__shared__ __align__(16) unsigned char sA_raw[sizeof(MegaArgs)];  // (MegaArgs has member initializers: raw storage)
#define sA (*reinterpret_cast<MegaArgs*>(sA_raw))
__shared__ MegaPhase sP[2];  // sP[0]: the entry being run (a constant address keeps its users' registers free);
                             // sP[1]: the next entry, copied in from global memory while this one runs
__shared__ PhaseRun sR;
__shared__ float s_red[32];
__shared__ int32_t s_tok;
__shared__ uint32_t s_last;
extern __shared__ __align__(128) uint8_t smem[];

#ifdef LLMI_MEGA_TIMING  // dev only (tools/mega_timeline.py): %globaltimer stamps of two CTAs for the LAST step
__device__ unsigned long long g_mega_stamp[2][1024][16];
__device__ __forceinline__ void mega_stamp(uint32_t pc, int i) {
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1) && pc < 1024) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_mega_stamp[blockIdx.x ? 1 : 0][pc][i] = t;
  }
}

#define MEGA_STAMP(pc, i) mega_stamp(pc, i)
// cheap cycle stamps of CTA 0 / thread 0 inside the helpers (slot = 32 per entry)
__device__ long long g_mega_cyc[1024][32];
__shared__ uint32_t s_pc;  // the entry being run
#define MEGA_CYC(i) do { if (threadIdx.x == 0 && blockIdx.x == 0 && s_pc < 1024) g_mega_cyc[s_pc][i] = clock64(); } while (0)
#else
#define MEGA_STAMP(pc, i) do { } while (0)
#define MEGA_CYC(i) do { } while (0)
#endif

// One cache line into L2.  Per-lane addresses: a warp covers 32 lines (4 KB) with one instruction — unlike the bulk
// prefetch (cp.async.bulk.prefetch.L2), whose operands must be warp-uniform and which ptxas therefore wraps in a
// lane-serializing loop.
__device__ __forceinline__ void l2_prefetch_line(const void* p) {
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p) : "memory");
}
__device__ __forceinline__ uint2* my_buf() { return sA.peers.base[sA.rank]; }

template <int N>
__device__ __forceinline__ void ll_wait_n(const uint2* const (&p)[N], const bool (&ok)[N], uint32_t tag, uint32_t (&v)[N]) {
  ll_wait_many<N>(p, ok, tag, sA.err, v);
}

// flagged store into this rank's buffer (gpu scope: the readers are this GPU's CTAs) and, `all`, into every peer's
__device__ __forceinline__ void ll_store_gpu(uint2* p, uint32_t bits, uint32_t tag) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(tag) : "memory");
}
__device__ __forceinline__ void push(bool all, uint32_t off, uint32_t bits, uint32_t tag) {
  ll_store_gpu(my_buf() + off, bits, tag);
  if (all) {
    for (uint32_t p = 0; p < sA.peers.n; ++p)
      if (p != sA.rank) ll_store(sA.peers.base[p] + off, bits, tag);
  }
}

// ---- arrival counters (hints) -----------------------------------------------------------------------------------
// Slot `slot` of the counter region counts, monotonically, the CTAs that have finished storing into the exchange
// vector(s) the slot stands for.  Every CTA of every storing rank bumps it once per use (also CTAs without rows),
// so after use number u the counter reads u * arrivals.  A consumer sleeps on it before it touches the flagged
// words; if the count is late or lost, the tag checks still decide.
__device__ __forceinline__ uint32_t* cnt_ptr(uint2* base, uint32_t slot) {
  return reinterpret_cast<uint32_t*>(base + sA.off_cnt + slot * MEGA_CNT_STRIDE);
}
__device__ __forceinline__ void hint_bump(uint32_t slot, bool all) {  // one thread, after the CTA's stores
  asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(cnt_ptr(my_buf(), slot)) : "memory");
  if (all) {
    for (uint32_t p = 0; p < sA.peers.n; ++p)
      if (p != sA.rank)
        asm volatile("red.relaxed.sys.global.add.u32 [%0], 1;" ::"l"(cnt_ptr(sA.peers.base[p], slot)) : "memory");
  }
}
__device__ __noinline__ void hint_wait(uint32_t slot, uint32_t expected) {  // one thread
  const uint32_t* p = cnt_ptr(my_buf(), slot);
  for (uint32_t n = 0; n < 4096; ++n) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    if (int32_t(v - expected) >= 0) return;
    __nanosleep(20);
  }
}
// uses of a per-layer slot up to and including layer `l` of global step `g` (slots come in two copies by layer parity)
__device__ __forceinline__ uint32_t layer_uses(uint32_t g, uint32_t l) {
  const uint32_t p = l & 1u, n_p = (sA.L + 1u - p) / 2u;
  return g * n_p + l / 2u + 1u;
}

// ---- prologues: the phase's input vector becomes the quantized activation in shared memory -------------------

// Q8_0 quantizer of a vector in shared memory, QI blocks per warp at a time so that their shuffle chains overlap
// (a CTA has few warps and the stage is latency-bound); per block the arithmetic of warp_quantize_q8_0.
template <int QI>
__device__ __forceinline__ void quantize_q8_0_ilp(const float (&v)[QI], const uint32_t (&b)[QI], const bool (&ok)[QI],
                                                  uint32_t n, uint8_t* buf, int lane) {
  float amax[QI];
#pragma unroll
  for (int k = 0; k < QI; ++k) amax[k] = fabsf(v[k]);
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int k = 0; k < QI; ++k) amax[k] = fmaxf(amax[k], __shfl_xor_sync(0xffffffffu, amax[k], o));
  int q[QI], sum[QI];
  float d[QI];
#pragma unroll
  for (int k = 0; k < QI; ++k) {
    d[k] = __fdiv_rn(amax[k], 127.0f);
    const float id = d[k] != 0.0f ? __fdiv_rn(1.0f, d[k]) : 0.0f;
    q[k] = nearest_int_fma(v[k], id);
    sum[k] = q[k];
  }
#pragma unroll
  for (int o = 16; o; o >>= 1)
#pragma unroll
    for (int k = 0; k < QI; ++k) sum[k] += __shfl_xor_sync(0xffffffffu, sum[k], o);
#pragma unroll
  for (int k = 0; k < QI; ++k) {
    if (!ok[k]) continue;
    reinterpret_cast<int8_t*>(buf)[b[k] * 32 + lane] = (int8_t)q[k];
    if (lane == 0)
      reinterpret_cast<uint32_t*>(buf + n)[b[k]] =
          uint32_t(__half_as_ushort(__float2half_rn(d[k]))) | (uint32_t(uint16_t(int16_t(sum[k]))) << 16);
  }
}

// activation of kind `kind` from the fp32 vector xs[0..n) in shared memory (emit_act of glue_device.cuh, with the
// Q8_0 blocks interleaved)
template <uint32_t KM>
__device__ __forceinline__ void emit_act_mega(int kind, const float* xs, uint32_t n, uint8_t* buf) {
  if ((KM & (1u << ACT_Q8_0)) && kind == ACT_Q8_0) {
    constexpr int QI = 4;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t nblk = n / 32;
    for (uint32_t b0 = warp; b0 < nblk; b0 += QI * MEGA_WARPS) {
      float v[QI];
      uint32_t b[QI];
      bool ok[QI];
#pragma unroll
      for (int k = 0; k < QI; ++k) {
        b[k] = b0 + k * MEGA_WARPS;
        ok[k] = b[k] < nblk;
        v[k] = ok[k] ? xs[b[k] * 32 + lane] : 0.0f;
      }
      quantize_q8_0_ilp<QI>(v, b, ok, n, buf, lane);
    }
  } else {
    emit_act<KM>(kind, xs, n, buf);
  }
}

// MEGA_PRO_NORM / _FIRST: norm_act_kernel's arithmetic.  That kernel runs T = 512 or 1024 threads (by n), thread t
// summing the squares of elements t, t + T, ... in that order, then a shuffle tree per warp, then the T/32 warp sums
// left to right.  Here the vector sits in shared memory and thread t plays that kernel's threads t, t + MEGA_THREADS
// (NORM_PASS of them): same partial sums, same tree, same final order — the bits of the per-launch path — with a
// handful of live registers.  The norm weights were copied to shared memory (cp.async) before the CTA started to wait
// for its input.
constexpr int NORM_PASS = 1024 / MEGA_THREADS;
__device__ __forceinline__ float norm_sum_sq(const float* x, uint32_t n, uint32_t T) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int p = 0; p < NORM_PASS; ++p) {
    const uint32_t lt = threadIdx.x + p * MEGA_THREADS;  // the per-launch kernel's thread
    float ss = 0.0f;
    if (lt < T)
      for (uint32_t i = lt; i < n; i += T) ss += __fmul_rn(x[i], x[i]);
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0) s_red[warp + p * MEGA_WARPS] = ss;
  }
  __syncthreads();
  const int nw = int(T / 32);
  float4 r[8];  // all warp sums first (independent loads), then the left-to-right chain
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = reinterpret_cast<const float4*>(s_red)[k];
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (4 * k < nw) {
      s += r[k].x;
      s += r[k].y;
      s += r[k].z;
      s += r[k].w;
    }
  }
  __syncthreads();  // s_red is reused
  return s;
}

// norm weights of the entry -> shared memory, asynchronously (issued before the wait for the entry's input)
__device__ __forceinline__ void stage_norm_weights(const MegaPhase& P) {
  if (P.pro != MEGA_PRO_NORM && P.pro != MEGA_PRO_FIRST) return;
  const uint32_t n4 = sA.E / 4;  // E is a multiple of 32
  const uint32_t wp = smem_u32(smem + sA.sm_wp), wn = smem_u32(smem + sA.sm_wn);
  for (uint32_t i = threadIdx.x; i < n4; i += MEGA_THREADS) {
    if (P.w_post && P.pro == MEGA_PRO_NORM)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wp + i * 16), "l"(P.w_post + i * 4) : "memory");
    if (P.w_next) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(wn + i * 16), "l"(P.w_next + i * 4) : "memory");
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <uint32_t TM>
__device__ MEGA_HOT void norm_prologue(uint32_t pb) {
  const MegaPhase& P = sP[pb];
  float* h_s = reinterpret_cast<float*>(smem + sA.sm_h);
  float* xs = reinterpret_cast<float*>(smem + sA.sm_xs);
  const float* wps = reinterpret_cast<const float*>(smem + sA.sm_wp);
  const float* wns = reinterpret_cast<const float*>(smem + sA.sm_wn);
  const uint32_t n = sA.E, T = n >= 2048 ? 1024u : 512u;
  const bool post = P.pro == MEGA_PRO_NORM;
  MEGA_CYC(3);
  if (post) {  // h += (rms_scale(y) * y) * w_post (model.cpp:843-854, 915-924)
    const uint2* y = my_buf() + P.in_off;
    for (uint32_t i0 = threadIdx.x; i0 < n; i0 += 4 * MEGA_THREADS) {
      const uint2* p[4];
      bool ok[4];
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        ok[k] = i0 + k * MEGA_THREADS < n;
        p[k] = y + i0 + k * MEGA_THREADS;
      }
      ll_wait_n<4>(p, ok, sR.in_tag, v);
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (ok[k]) xs[i0 + k * MEGA_THREADS] = __uint_as_float(v[k]);
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    MEGA_CYC(4);
    const float sc = rms_scale(norm_sum_sq(xs, n, T), n, sA.eps);
    MEGA_CYC(6);
    for (uint32_t i = threadIdx.x; i < n; i += MEGA_THREADS) {
      const float yv = xs[i];
      const float add = P.w_post ? __fmul_rn(__fmul_rn(sc, yv), wps[i]) : yv;
      h_s[i] = __fadd_rn(h_s[i], add);
    }
    __syncthreads();
  } else {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }
  if (!P.w_next) return;
  MEGA_CYC(7);
  const float sc = rms_scale(norm_sum_sq(h_s, n, T), n, sA.eps);  // xn = (rms_scale(h) * h) * w (model.cpp:346-386)
  MEGA_CYC(8);
  for (uint32_t i = threadIdx.x; i < n; i += MEGA_THREADS) xs[i] = __fmul_rn(__fmul_rn(sc, h_s[i]), wns[i]);
  __syncthreads();
  MEGA_CYC(9);
  emit_act_mega<llmi_kind_mask(TM)>(int(P.act_kind), xs, n, smem + sA.sm_act);
  MEGA_CYC(10);
}

// MEGA_PRO_QUANT: flagged fp32 vector -> activation (the quantizers attention_kernel / geglu_act_kernel fuse)
template <uint32_t TM>
__device__ MEGA_HOT void quant_prologue(uint32_t pb) {
  constexpr uint32_t KM = llmi_kind_mask(TM);
  const MegaPhase& P = sP[pb];
  uint8_t* act = smem + sA.sm_act;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t n = P.K, in_tag = sR.in_tag;
  const uint2* x = my_buf() + P.in_off;
  if ((KM & (1u << ACT_Q8_0)) && P.act_kind == ACT_Q8_0) {
    constexpr int G = 4;  // blocks in flight per warp (loads and shuffle chains)
    const uint32_t nblk = n / 32;
    for (uint32_t b0 = warp; b0 < nblk; b0 += MEGA_WARPS * G) {
      const uint2* p[G];
      bool ok[G];
      uint32_t v[G], b[G];
      float f[G];
#pragma unroll
      for (int k = 0; k < G; ++k) {
        b[k] = b0 + k * MEGA_WARPS;
        ok[k] = b[k] < nblk;
        p[k] = x + b[k] * 32 + lane;
      }
      ll_wait_n<G>(p, ok, in_tag, v);
#pragma unroll
      for (int k = 0; k < G; ++k) f[k] = __uint_as_float(v[k]);
      quantize_q8_0_ilp<G>(f, b, ok, n, act, lane);
    }
  } else if ((KM & (1u << ACT_Q8_K)) && P.act_kind == ACT_Q8_K) {
    for (uint32_t sb = warp; sb < n / 256; sb += MEGA_WARPS) {
      const uint2* p[8];
      bool ok[8];
      uint32_t v[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        ok[k] = true;
        p[k] = x + sb * 256 + lane * 8 + k;
      }
      ll_wait_n<8>(p, ok, in_tag, v);
      float fv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) fv[k] = __uint_as_float(v[k]);
      warp_quantize_q8_k(fv, sb, n, act, lane);
    }
  } else if (KM & ((1u << ACT_F16) | (1u << ACT_F32))) {
    const uint32_t n_pad = P.act_kind == ACT_F16 ? ((n + 7) & ~7u) : ((n + 3) & ~3u);
    for (uint32_t i0 = threadIdx.x; i0 < n_pad; i0 += MEGA_THREADS * 4) {
      const uint2* p[4];
      bool ok[4];
      uint32_t v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t i = i0 + k * MEGA_THREADS;
        ok[k] = i < n;
        p[k] = x + i;
      }
      ll_wait_n<4>(p, ok, in_tag, v);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t i = i0 + k * MEGA_THREADS;
        if (i >= n_pad) continue;
        const float f = ok[k] ? __uint_as_float(v[k]) : 0.0f;
        if (P.act_kind == ACT_F16) reinterpret_cast<uint16_t*>(act)[i] = f2h(f);
        else reinterpret_cast<float*>(act)[i] = f;
      }
    }
  }
}

// ---- the mat-vec phase ------------------------------------------------------------------------------------------
// The phase's slabs (of all its matrices, in order; GEGLU: slab pairs) are dealt to the CTAs round-robin: CTA c
// owns virtual slabs c, c + G, c + 2G ...  It works through them in batches of SB (what the partials buffer
// holds): the batch's (slab, K-chunk) items go round-robin to the warps exactly as in gemv_slab_kernel, chunk
// partials land in shared memory, one barrier, then thread (slab, row) adds its row's partials left to right.

// virtual slab -> (matrix, slab)
__device__ __forceinline__ void resolve(const MegaPhase& P, uint32_t v, uint32_t& mi, uint32_t& s) {
  mi = 0;
  s = v;
  if (P.n_mats > 1 && s >= P.m[0].n_slabs) {
    s -= P.m[0].n_slabs;
    mi = 1;
    if (P.n_mats > 2 && s >= P.m[1].n_slabs) {
      s -= P.m[1].n_slabs;
      mi = 2;
    }
  }
}

__device__ __forceinline__ unsigned long long gemv_epilogue(uint32_t pb, uint32_t i0, uint32_t nsl, unsigned long long best) {
  const MegaPhase& P = sP[pb];
  const float* part = reinterpret_cast<const float*>(smem + sA.sm_part);
  const uint32_t J = P.J, out_tag = sR.out_tag;
  for (uint32_t idx = threadIdx.x; idx < nsl * LLMI_SLAB; idx += MEGA_THREADS) {
    const uint32_t sl = idx / LLMI_SLAB, rr = idx % LLMI_SLAB;
    const uint32_t v = blockIdx.x + (i0 + sl) * gridDim.x;
    if (P.epi == MEGA_EPI_GEGLU) {
      const float* pg = part + size_t(2 * sl) * J * LLMI_SLAB + rr;
      const float* pu = pg + size_t(J) * LLMI_SLAB;
      float g = pg[0], u = pu[0];
      for (uint32_t j = 1; j < J; ++j) {  // canonical order
        g += pg[j * LLMI_SLAB];
        u += pu[j * LLMI_SLAB];
      }
      const uint32_t row = v * LLMI_SLAB + rr;
      if (row < P.m[0].n_local) push(P.push_all, P.out_off[0] + P.m[0].row0 + row, __float_as_uint(geglu(g, u)), out_tag);
      continue;
    }
    const float* p = part + size_t(sl) * J * LLMI_SLAB + rr;
    float sum = p[0];
    for (uint32_t j = 1; j < J; ++j) sum += p[j * LLMI_SLAB];  // canonical order
    uint32_t mi, s;
    resolve(P, v, mi, s);
    const GemvArgs& a = P.m[mi];
    const uint32_t row = s * LLMI_SLAB + rr;
    if (row >= a.n_local) continue;
    if (P.epi == MEGA_EPI_FLAG) {
      push(P.push_all, P.out_off[mi] + a.row0 + row, __float_as_uint(sum), out_tag);
    } else {  // logits: final soft-cap (model.cpp:1036-1041), then greedy argmax (main.cpp:193) and / or the value
      if (sA.final_softcap > 0.0f) sum = __fmul_rn(sA.final_softcap, tanhf(__fdiv_rn(sum, sA.final_softcap)));
      if (sA.logits_mode == 2) {  // order-preserving float -> uint; ties go to the smaller row (std::max_element)
        uint32_t u = __float_as_uint(sum);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        const unsigned long long k = (uint64_t(u) << 32) | uint32_t(0xffffffffu - (a.row0 + row));
        best = k > best ? k : best;
      }
      sA.logits[a.row0 + row] = sum;
      if (sA.logits_mode == 1 && sA.world > 1) push(true, sA.off_logits + a.row0 + row, __float_as_uint(sum), out_tag);
    }
  }
  return best;
}

// (slab, chunk) item t of the batch starting at list position i0 -> matrix, slab, chunk
struct Item {
  uint32_t mi, s, j;
};
__device__ __forceinline__ Item item_of(const MegaPhase& P, uint32_t t, uint32_t i0, uint32_t J, uint32_t sub) {
  Item it;
  const uint32_t sl = t / J;
  it.j = t - sl * J;
  const uint32_t v = blockIdx.x + (i0 + sl / sub) * gridDim.x;
  if (sub == 2) {  // GEGLU: (gate slab v, up slab v)
    it.mi = sl & 1u;
    it.s = v;
  } else {
    resolve(P, v, it.mi, it.s);
  }
  return it;
}

// The item's pieces of the planes into L2, by the whole warp: chunk j of a slab is bytes [j * chunk, (j + 1) * chunk)
// of the slab's run in every plane (the last chunk may be short); lane l takes the l-th 128-byte line.
__device__ __forceinline__ void prefetch_item(const MegaPhase& P, const Item& it, int lane) {
  if (!sA.pf_mode) return;
  const GemvArgs& a = P.m[it.mi];
  const uint32_t bq = P.slab_q[it.mi], bd = P.slab_d[it.mi], bx = P.slab_x[it.mi];
  const uint32_t cq = P.chunk_q[it.mi], cd = P.chunk_d[it.mi], cx = P.chunk_x[it.mi];
  const uint32_t q0 = it.j * cq, d0 = it.j * cd, x0 = it.j * cx;
  const uint32_t nq = (min(cq, bq - q0) + 127) / 128, nd = bd ? (min(cd, bd - d0) + 127) / 128 : 0u,
                 nx = bx ? (min(cx, bx - x0) + 127) / 128 : 0u;
  for (uint32_t l = lane; l < nq + nd + nx; l += 32) {
    const uint8_t* p = l < nq ? a.q + size_t(it.s) * bq + q0 + l * 128
                     : l < nq + nd ? a.d + size_t(it.s) * bd + d0 + (l - nq) * 128
                                   : a.x + size_t(it.s) * bx + x0 + (l - nq - nd) * 128;
    l2_prefetch_line(p);
  }
}

template <class B>
__device__ MEGA_HOT void gemv_loop(uint32_t pb) {
  unsigned long long best = 0;
  constexpr int N = B::C;
  const MegaPhase& P = sP[pb];
  const uint8_t* act = smem + sA.sm_act;
  float* part = reinterpret_cast<float*>(smem + sA.sm_part);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane & 7, sub = lane >> 3;
  const uint32_t J = P.J, my_count = sR.my_count, SB = sR.SB, msub = sR.sub;
  for (uint32_t i0 = 0; i0 < my_count; i0 += SB) {
    const uint32_t nsl = min(SB, my_count - i0);
    const uint32_t n_items = nsl * msub * J;
    // Software pipeline over the warp's items t, t + W, t + 2W ...: the loads of the next item are in flight while
    // this one is folded (two fragment sets, ping-pong), and lane 0 requests the item PREFETCH_AHEAD rounds further
    // down into L2 (the next batch included).
    auto prefetch_ahead = [&](uint32_t t) {
      const uint32_t tp = t + PREFETCH_AHEAD * MEGA_WARPS;
      if (tp < n_items) {
        prefetch_item(P, item_of(P, tp, i0, J, msub), lane);
      } else if (i0 + SB < my_count && tp - n_items < min(SB, my_count - i0 - SB) * msub * J) {
        prefetch_item(P, item_of(P, tp - n_items, i0 + SB, J, msub), lane);
      }
    };
    FragSet<B, N> fa, fb;
    Item ia, ib;
    uint32_t t = warp;
    if (t < n_items) {
      ia = item_of(P, t, i0, J, msub);
      MEGA_CYC(13);
      load_item<B, N>(fa, P.m[ia.mi], ia.s, ia.j, r, sub);
      MEGA_CYC(21);
    }
#pragma unroll 1
    while (t < n_items) {
      const uint32_t t1 = t + MEGA_WARPS, t2 = t1 + MEGA_WARPS;
      if (t1 < n_items) {
        ib = item_of(P, t1, i0, J, msub);
        load_item<B, N>(fb, P.m[ib.mi], ib.s, ib.j, r, sub);
      }
      prefetch_ahead(t);
      if (t == uint32_t(warp)) MEGA_CYC(22);
      {
        const float v = compute_item<B, N>(fa, P.m[ia.mi], act, ia.j, sub);
        if (t == uint32_t(warp)) MEGA_CYC(23);
        if (lane < LLMI_SLAB) part[t * LLMI_SLAB + lane] = v;
      }
      if (t == uint32_t(warp)) MEGA_CYC(24);
      if (t1 >= n_items) break;
      if (t2 < n_items) {
        ia = item_of(P, t2, i0, J, msub);
        load_item<B, N>(fa, P.m[ia.mi], ia.s, ia.j, r, sub);
      }
      prefetch_ahead(t1);
      {
        const float v = compute_item<B, N>(fb, P.m[ib.mi], act, ib.j, sub);
        if (lane < LLMI_SLAB) part[t1 * LLMI_SLAB + lane] = v;
      }
      t = t2;
    }
    MEGA_CYC(14);
    __syncthreads();
    MEGA_CYC(15);
    best = gemv_epilogue(pb, i0, nsl, best);
    MEGA_CYC(16);
    __syncthreads();  // the partials are reused by the next batch
    MEGA_CYC(17);
  }
  if (P.epi == MEGA_EPI_LOGITS && sA.logits_mode == 2) {  // greedy argmax: warp keys -> the rank's key
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0 && best) atomicMax(sA.key, best);
  }
}

// The first PREFETCH_BYTES of this CTA's share of an entry's weights, requested into L2 line by line (thread t takes
// lines t, t + MEGA_THREADS ... of the concatenation of its first slabs' planes).
__device__ __forceinline__ void prefetch_share(const MegaPhase& P) {
  if (P.kind != MEGA_GEMV || !sA.pf_mode) return;
  const uint32_t my_count = P.v_total > blockIdx.x ? (P.v_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
  const uint32_t msub = P.epi == MEGA_EPI_GEGLU ? 2u : 1u;
  if (my_count == 0) return;
  const uint32_t lq = (P.slab_q[0] + 127) / 128, ld = (P.slab_d[0] + 127) / 128, lx = (P.slab_x[0] + 127) / 128;
  const uint32_t per = lq + ld + lx;  // lines of one slab (all matrices of an entry share K, so also their slab sizes)
  const uint32_t n_sl = min(my_count * msub, max(1u, PREFETCH_BYTES / 128 / max(per, 1u)));
  for (uint32_t t = threadIdx.x; t < n_sl * per; t += MEGA_THREADS) {
    const uint32_t sl = t / per, l = t - sl * per;
    const uint32_t v = blockIdx.x + (sl / msub) * gridDim.x;
    uint32_t mi, s;
    if (msub == 2) {
      mi = sl & 1u;
      s = v;
    } else {
      resolve(P, v, mi, s);
    }
    const GemvArgs& a = P.m[mi];
    const uint8_t* p = l < lq ? a.q + size_t(s) * P.slab_q[mi] + l * 128
                     : l < lq + ld ? a.d + size_t(s) * P.slab_d[mi] + (l - lq) * 128
                                   : a.x + size_t(s) * P.slab_x[mi] + (l - lq - ld) * 128;
    l2_prefetch_line(p);
  }
}

template <uint32_t TM>
__device__ __forceinline__ void gemv_phase(uint32_t pb, uint32_t tagbase, uint32_t g, uint32_t pc) {
  const MegaPhase& P = sP[pb];
  if (threadIdx.x == 0) {
    PhaseRun R;
    R.my_count = P.v_total > blockIdx.x ? (P.v_total - blockIdx.x + gridDim.x - 1) / gridDim.x : 0u;
    R.sub = P.epi == MEGA_EPI_GEGLU ? 2u : 1u;
    R.SB = max(1u, sA.part_floats / (LLMI_SLAB * P.J * R.sub));
    R.in_tag = tagbase + P.in_tag;
    R.out_tag = tagbase + P.out_tag;
    sR = R;
  }
  __syncthreads();
  MEGA_CYC(0);
  // the weights depend on nothing: ask for the NEXT entry's share now (its descriptor has landed in sP[1]), so that it
  // is in L2 a whole phase before its first load; entry 0's own share goes out here too
  if (pc == 0) prefetch_share(P);
  if (pc + 1 < sA.n_prog) prefetch_share(sP[1]);
  stage_norm_weights(P);
  MEGA_CYC(1);
  if (P.pro != MEGA_PRO_REUSE && P.pro != MEGA_PRO_FIRST) {
    // sleep on the producers' arrival counter first (one thread), then take the flagged words
    if (threadIdx.x == 0 && sA.hints) hint_wait(P.in_slot, layer_uses(g, P.in_layer) * P.in_arrivals);
    __syncthreads();
  }
  MEGA_CYC(2);
  if (P.pro == MEGA_PRO_QUANT) {
    // (a CTA without rows here still prepares the activation: a MEGA_PRO_REUSE entry may follow)
    quant_prologue<TM>(pb);
  } else if (P.pro != MEGA_PRO_REUSE) {
    norm_prologue<TM>(pb);
  }
  __syncthreads();
  MEGA_STAMP(pc, 1);
  MEGA_CYC(12);
  if (sR.my_count != 0) {
    // (only the formats of this instantiation exist in its code: see the note at decode_mega_kernel)
    const uint32_t ty = P.type;
    if ((TM & mega_type_bit(LLMI_Q4_0)) && ty == LLMI_Q4_0) gemv_loop<BodyQ4_0>(pb);
    if ((TM & mega_type_bit(LLMI_Q8_0)) && ty == LLMI_Q8_0) gemv_loop<BodyQ8_0>(pb);
    if ((TM & mega_type_bit(LLMI_Q5_0)) && ty == LLMI_Q5_0) gemv_loop<BodyQ5_0>(pb);
    if ((TM & mega_type_bit(LLMI_Q4_K)) && ty == LLMI_Q4_K) gemv_loop<BodyQ4_K>(pb);
    if ((TM & mega_type_bit(LLMI_Q6_K)) && ty == LLMI_Q6_K) gemv_loop<BodyQ6_K>(pb);
    if ((TM & mega_type_bit(LLMI_F16)) && ty == LLMI_F16) gemv_loop<BodyHalf<false>>(pb);
    if ((TM & mega_type_bit(LLMI_BF16)) && ty == LLMI_BF16) gemv_loop<BodyHalf<true>>(pb);
  }
  if (P.bump) {  // every CTA, rows or not: the consumers count arrivals (gemv_loop ends with a barrier)
    if (threadIdx.x == 0 && sA.hints) hint_bump(P.out_slot, P.push_all != 0);
  }
}

// ---- attention phase -------------------------------------------------------------------------------------------
template <int D>
__device__ MEGA_HOT void attention_head(uint32_t layer, uint32_t head, int pos, uint32_t tagbase) {
  const MegaAttn& T = sA.attn[layer];
  AttnArgs a;
  a.q = a.k = a.v = nullptr;
  a.wq_norm = T.q_norm;
  a.wk_norm = T.k_norm;
  a.kcache = T.kcache;
  a.vcache = T.vcache;
  a.H = sA.H; a.HK = sA.HK; a.D = sA.D; a.t_max = sA.t_max;
  a.eps = sA.eps;
  a.attn_scale = sA.attn_scale;
  a.rope_table = T.rope;
  a.pos = nullptr;
  a.softcap = sA.attn_softcap;
  a.out = nullptr;
  const uint2* mine = my_buf();
  a.ll_q = mine + T.off_q;
  a.ll_k = mine + T.off_k;
  a.ll_v = mine + T.off_v;
  AttnMega mg;
  mg.pos = pos;
  mg.in_tag = tagbase + T.in_tag;
  mg.out_tag = tagbase + T.out_tag;
  mg.out_off = T.off_out;
  mg.err = sA.err;
  if (sA.attn_push_all) {
    mg.out_peers = sA.peers;
  } else {
    mg.out_peers.n = 1;
    mg.out_peers.base[0] = my_buf();
  }
  attention_body<D, 0, true>(a, sA.attn_nbuf, head, 0u, smem + sA.sm_attn, &mg);
}

template <int DS>
__device__ __forceinline__ void attention_phase(uint32_t layer, int pos, uint32_t tagbase, uint32_t g) {
  const uint32_t slot_in = MEGA_CNT_LAYER + (layer & 1u) * MEGA_TAGS_PER_LAYER + 0u;   // q/k/v
  const uint32_t slot_out = MEGA_CNT_LAYER + (layer & 1u) * MEGA_TAGS_PER_LAYER + 1u;  // heads' outputs
  __syncthreads();          // sP[1] (the entry after the attention) has landed
  prefetch_share(sP[1]);    // every CTA, before its heads: the attn_output rows of a head's CTA are on the critical path
  bool first = true;
  for (uint32_t head = sA.head_begin + blockIdx.x; head < sA.head_end; head += gridDim.x) {
    if (first && sA.hints) {
      if (threadIdx.x == 0) hint_wait(slot_in, layer_uses(g, layer) * gridDim.x * (sA.attn_push_all ? 1u : sA.world));
      __syncthreads();
    }
    first = false;
    if constexpr (DS != 0) {
      attention_head<DS>(layer, head, pos, tagbase);
    } else {
      switch (sA.D) {
        case 64: attention_head<64>(layer, head, pos, tagbase); break;
        case 128: attention_head<128>(layer, head, pos, tagbase); break;
        case 256: attention_head<256>(layer, head, pos, tagbase); break;
        case 512: attention_head<512>(layer, head, pos, tagbase); break;
        default: break;
      }
    }
    __syncthreads();  // the attention scratch is the next phase's activation / partials
  }
  if (threadIdx.x == 0 && sA.hints) hint_bump(slot_out, sA.attn_push_all != 0);
}

// ---- the kernel -------------------------------------------------------------------------------------------------
constexpr uint32_t PHASE_WORDS = sizeof(MegaPhase) / 4;
static_assert(sizeof(MegaPhase) % 4 == 0 && PHASE_WORDS <= MEGA_THREADS, "MegaPhase is copied by words");
__device__ __forceinline__ void fetch_entry(uint32_t pc) {  // program entry pc -> sP[1] (global -> shared)
  if (pc < sA.n_prog && threadIdx.x < PHASE_WORDS)
    reinterpret_cast<uint32_t*>(&sP[1])[threadIdx.x] = reinterpret_cast<const uint32_t*>(sA.prog + pc)[threadIdx.x];
}
__device__ __forceinline__ void advance_entry() {  // sP[1] -> sP[0]; callers put barriers on both sides
  if (threadIdx.x < PHASE_WORDS) reinterpret_cast<uint32_t*>(&sP[0])[threadIdx.x] = reinterpret_cast<const uint32_t*>(&sP[1])[threadIdx.x];
}

// this step's token, then its embedding row * sqrt(E) (model.cpp:240-344, 710-712) into the CTA's copy of the
// residual stream
template <uint32_t TM>
__device__ __noinline__ void step_begin(uint32_t step) {
  float* h_s = reinterpret_cast<float*>(smem + sA.sm_h);
  const uint2* mine = my_buf();
  const uint32_t g = sA.epoch0 - 1u + step;
  const uint32_t tagbase = (sA.epoch0 + step) * sA.tag_mul;
  fetch_entry(0);
  if (threadIdx.x == 0) {
    int32_t t;
    if (sA.tokens) {
      t = sA.tokens[step];
    } else if (step == 0) {
      t = sA.first_token;
    } else {
      if (sA.hints) hint_wait(MEGA_CNT_TOK, sA.tok_uses0 + step);
      t = int32_t(ll_wait(mine + sA.off_tok, tagbase - sA.tag_mul + sA.tag_tok, sA.err));
    }
    s_tok = t;
  }
  __syncthreads();
  const uint32_t tok = uint32_t(s_tok);
  EmbedArgs e{sA.embd_type, sA.embd_nb, sA.E, sA.embd_q, sA.embd_d, sA.embd_x, sA.embd_row_begin, sA.embd_row_end};
  if (sA.world == 1) {
    for (uint32_t i = threadIdx.x; i < sA.E; i += MEGA_THREADS) h_s[i] = dequant_elem<TM>(e, tok, i) * sA.embed_scale;
  } else {  // the rank that holds the row sends it to everybody (its CTAs split the elements)
    const uint32_t tag = tagbase + sA.tag_h;
    if (tok >= sA.embd_row_begin && tok < sA.embd_row_end)
      for (uint32_t i = blockIdx.x * MEGA_THREADS + threadIdx.x; i < sA.E; i += gridDim.x * MEGA_THREADS)
        push(true, sA.off_h + i, __float_as_uint(dequant_elem<TM>(e, tok - sA.embd_row_begin, i) * sA.embed_scale), tag);
    __syncthreads();
    if (threadIdx.x == 0 && sA.hints) {
      hint_bump(MEGA_CNT_H, true);
      hint_wait(MEGA_CNT_H, (g + 1u) * gridDim.x * sA.world);
    }
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < sA.E; i += MEGA_THREADS) h_s[i] = ll_waitf(mine + sA.off_h + i, tag, sA.err);
  }
  __syncthreads();
}

// per layer: norm + q/k/v (model.cpp:754-803), attention (:388-550), attn_output (:557), post-attention norm +
// residual + ffn_norm + gate/up + GEGLU (:843-901), ffn_down (:909); last: final norms + logits (:983-1041)
template <uint32_t TM, int DS>
__device__ __forceinline__ void run_entry(uint32_t step, uint32_t pc) {
  constexpr uint32_t pb = 0;
  const uint32_t g = sA.epoch0 - 1u + step;  // steps this model has run before this one
  const uint32_t tagbase = (sA.epoch0 + step) * sA.tag_mul;
  advance_entry();      // (the previous entry ended with a barrier)
  __syncthreads();
  fetch_entry(pc + 1);  // lands in sP[1] while this entry runs
#ifdef LLMI_MEGA_TIMING
  if (threadIdx.x == 0) s_pc = pc;
#endif
  MEGA_STAMP(pc, 0);
  if (sP[pb].kind == MEGA_ATTN) {
    attention_phase<DS>(sP[pb].layer, sA.pos0 + int(step), tagbase, g);
  } else {
    const bool want_logits = sA.logits_mode == 2 || (sA.logits_mode == 1 && step + 1 == sA.n_steps);
    // (a prompt token has no logits entry to run: nothing reads the last residual update)
    if (sP[pb].epi != MEGA_EPI_LOGITS || want_logits) gemv_phase<TM>(pb, tagbase, g, pc);
  }
  __syncthreads();
  MEGA_STAMP(pc, 2);
}

// greedy argmax: the rank's key is complete when every CTA has passed here; the last one turns it into the token
__device__ __noinline__ void step_end(uint32_t step) {
  const uint2* mine = my_buf();
  const uint32_t tagbase = (sA.epoch0 + step) * sA.tag_mul;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t old = atomicAdd(sA.done_ctr, 1u);
    s_last = old == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (s_last && threadIdx.x == 0) {
    __threadfence();
    unsigned long long k = atomicExch(sA.key, 0ull);
    *reinterpret_cast<volatile uint32_t*>(sA.done_ctr) = 0u;
    if (sA.world > 1) {  // every rank's key covers its own rows: swap them (two flagged words per rank)
      const uint32_t tag = tagbase + sA.tag_key;
      for (uint32_t p = 0; p < sA.peers.n; ++p) {
        ll_store(sA.peers.base[p] + sA.off_key + 2 * sA.rank, uint32_t(k >> 32), tag);
        ll_store(sA.peers.base[p] + sA.off_key + 2 * sA.rank + 1, uint32_t(k), tag);
      }
      k = 0ull;
      for (uint32_t r = 0; r < sA.peers.n; ++r) {
        const unsigned long long hi = ll_wait(mine + sA.off_key + 2 * r, tag, sA.err);
        const unsigned long long lo = ll_wait(mine + sA.off_key + 2 * r + 1, tag, sA.err);
        const unsigned long long kk = (hi << 32) | lo;
        k = kk > k ? kk : k;
      }
    }
    const int32_t next = int32_t(0xffffffffu - uint32_t(k));
    if (sA.gen) sA.gen[step] = next;
    if (sA.d_tok) *sA.d_tok = next;
    __threadfence();  // key / counter resets before anybody can start the next step's argmax
    ll_store(my_buf() + sA.off_tok, uint32_t(next), tagbase + sA.tag_tok);
    if (sA.hints) hint_bump(MEGA_CNT_TOK, false);
  }
}

// TM = the weight formats (mega_type_bit) this instantiation carries code for, DS = its head size (0: any of
// 64/128/256/512).  ptxas allots registers to the functions of one call graph out of ONE budget: with all seven
// format loops and four attention bodies reachable from one kernel it spilled the freshly loaded weight
// fragments of every loop (LDG -> STL -> LDL, i.e. serialized loads).  Specialized instantiations inline their
// two or three hot bodies into the kernel and keep the fragments in registers; the all-formats instantiation
// (mega_any.cu, bodies out of line) stays as the fallback for anything else.
template <uint32_t TM, int DS>
__global__ void __launch_bounds__(MEGA_THREADS, 1) decode_mega_kernel(const __grid_constant__ MegaArgs A) {
  {
    constexpr uint32_t WORDS = sizeof(MegaArgs) / 4;
    static_assert(sizeof(MegaArgs) % 4 == 0 && WORDS <= MEGA_THREADS, "MegaArgs is copied by words");
    if (threadIdx.x < WORDS) reinterpret_cast<uint32_t*>(&sA)[threadIdx.x] = reinterpret_cast<const uint32_t*>(&A)[threadIdx.x];
  }
  __syncthreads();
  for (uint32_t step = 0; step < sA.n_steps; ++step) {
    step_begin<TM>(step);
    for (uint32_t pc = 0; pc < sA.n_prog; ++pc) run_entry<TM, DS>(step, pc);
    if (sA.logits_mode == 2) step_end(step);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && sA.d_pos) *sA.d_pos = sA.pos0 + int(sA.n_steps);
}


// ---- per-instantiation host glue ------------------------------------------------------------------------------
template <uint32_t TM, int DS>
cudaError_t mega_variant_init(size_t* smem_limit) {
  int dev = 0, optin = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev)) != cudaSuccess) return e;
  cudaFuncAttributes fa;
  if ((e = cudaFuncGetAttributes(&fa, decode_mega_kernel<TM, DS>)) != cudaSuccess) return e;
  *smem_limit = size_t(optin) - fa.sharedSizeBytes;
  return cudaFuncSetAttribute(decode_mega_kernel<TM, DS>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(*smem_limit));
}

template <uint32_t TM, int DS>
cudaError_t mega_variant_launch(const MegaArgs& a, uint32_t n_ctas, size_t smem_bytes, cudaStream_t s) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas);
  cfg.blockDim = dim3(MEGA_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;  // all CTAs co-resident: they wait for each other's flagged stores
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, decode_mega_kernel<TM, DS>, a);
}

#ifdef LLMI_MEGA_TIMING
cudaError_t mega_variant_stamps(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_mega_stamp, sizeof(unsigned long long) * 2 * 1024 * 16);
}
cudaError_t mega_variant_cycles(long long* out) { return cudaMemcpyFromSymbol(out, g_mega_cyc, sizeof(long long) * 1024 * 32); }
#endif

#define MEGA_VARIANT(TM, DS) MegaVariant{TM, DS, mega_variant_init<TM, DS>, mega_variant_launch<TM, DS>}

}  // namespace
