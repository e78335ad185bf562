// Persistent, bulk-copy-fed decode mat-vec for sm_100a — included by gemv.cu.
//
// gemv_slab_kernel gets its bytes in flight from occupancy: thousands of short CTAs, one work item (2 KB of weights)
// per warp per DRAM round trip, and a fixed cost of ~3-4 us per launch in ramp, activation hand-over and tail that a
// 15-65 MB matrix cannot amortise (0.48 / 0.69 of the HBM peak at the 4b / 27b gate shapes).  This kernel keeps the
// same work items, the same per-item arithmetic (Body::compute) and the same canonical summation order — it is
// bit-identical — but changes who moves the bytes:
//   * the grid is PERSISTENT: ctas_per_sm x SM-count CTAs, each owning a contiguous range of the launch's
//     (slab, K-chunk) items, split at ITEM granularity so every CTA gets the same number of bytes (+-1 item);
//   * every warp owns a private ring of D shared-memory slots filled by asynchronous 16-byte copies (cp.async.cg ->
//     LDGSTS: global -> shared without a register round trip, L1 bypassed), one commit group per item.  The warp
//     that consumes a slot refills it for its item D rounds ahead, so there is no producer warp and no cross-warp
//     hand-over in the loop; bytes in flight per SM = CTAs x W x D x item bytes (147 KB at 2 x 8 x 4 x 2304),
//     independent of registers.  (The bulk-copy engine was measured first and is the wrong tool at this grain: a
//     warp completes only ~3.9 cp.async.bulk per microsecond whatever their size below 16 KB — tools/bulk_bench.cu,
//     profiles/r02_bulk_bench.txt — so 2 KB items with two planes ran at 0.39 of the peak.)
//   * weights depend on nothing, so the rings are filled BEFORE griddepcontrol.wait: while the predecessor (a glue
//     kernel with a handful of CTAs) runs, the HBM pipe keeps streaming the first W x D items of every CTA;
//   * a slab whose K-chunks straddle two (or more) CTAs is summed by the CTA that holds its LAST chunk; the others
//     process those chunks FIRST and hand their chunk partials over as flagged 64-bit words {value, 1} in a small
//     scratch buffer (single-copy atomic: no fence).  The owner adds all J partials left to right — the canonical
//     order — and clears the words it consumed.  A CTA therefore only waits for CTAs with a smaller index, which
//     start first and wait for nobody behind them: no cooperative launch is needed for forward progress.  Publishing
//     happens after griddepcontrol.wait and consuming before the kernel ends, so launches of one stream never overlap
//     in the buffer.
#pragma once
// (included inside gemv.cu's anonymous namespace, like umma_prefill.cuh)

// ---- per-format staging geometry of one work item -------------------------------------------------------------
// CELLS = K-cells (32-blocks / super-blocks / 8-element groups) per item; QB / DB / XB = bytes of one cell of the
// q / d / x plane for the 8 rows of a slab.  Slot layout: [q cells][d cells][x cells].
template <class B>
struct RingFmt;
template <>
struct RingFmt<BodyQ4_0> {
  static constexpr uint32_t CELLS = 16, QB = 128, DB = 16, XB = 0;
};
template <>
struct RingFmt<BodyQ8_0> {
  static constexpr uint32_t CELLS = 16, QB = 256, DB = 16, XB = 0;
};
template <>
struct RingFmt<BodyQ5_0> {
  static constexpr uint32_t CELLS = 16, QB = 128, DB = 16, XB = 32;
};
template <>
struct RingFmt<BodyQ4_K> {
  static constexpr uint32_t CELLS = 2, QB = 1024, DB = 0, XB = 128;
};
template <>
struct RingFmt<BodyQ6_K> {
  static constexpr uint32_t CELLS = 2, QB = 1536, DB = 16, XB = 128;
};
template <bool BF>
struct RingFmt<BodyHalf<BF>> {
  static constexpr uint32_t CELLS = 16, QB = 128, DB = 0, XB = 0;
};
template <class B>
struct RingGeo {
  using F = RingFmt<B>;
  static constexpr uint32_t OFF_D = F::CELLS * F::QB, OFF_X = OFF_D + F::CELLS * F::DB;
  static constexpr uint32_t SLOT = (OFF_X + F::CELLS * F::XB + 127u) & ~127u;
};

// One K-unit of the item (the t-th of Body::C) out of the slot: the shared-memory twin of Body::load.
// nc = cells the item holds (a slab's last item may be short).
__device__ __forceinline__ void ring_lds(BodyQ4_0::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r, int sub) {
  const uint32_t c = 4 * t + sub;
  if (c < nc) {
    f.w = *reinterpret_cast<const uint4*>(slot + (c * 8 + r) * 16);
    f.d = *reinterpret_cast<const uint16_t*>(slot + RingGeo<BodyQ4_0>::OFF_D + (c * 8 + r) * 2);
  }
}
__device__ __forceinline__ void ring_lds(BodyQ8_0::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r, int sub) {
  const uint32_t c = 4 * t + sub;
  if (c < nc) {
    f.w0 = *reinterpret_cast<const uint4*>(slot + (c * 16 + r) * 16);
    f.w1 = *reinterpret_cast<const uint4*>(slot + (c * 16 + 8 + r) * 16);
    f.d = *reinterpret_cast<const uint16_t*>(slot + RingGeo<BodyQ8_0>::OFF_D + (c * 8 + r) * 2);
  }
}
__device__ __forceinline__ void ring_lds(BodyQ5_0::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r, int sub) {
  const uint32_t c = 4 * t + sub;
  if (c < nc) {
    f.w = *reinterpret_cast<const uint4*>(slot + (c * 8 + r) * 16);
    f.d = *reinterpret_cast<const uint16_t*>(slot + RingGeo<BodyQ5_0>::OFF_D + (c * 8 + r) * 2);
    f.qh = *reinterpret_cast<const uint32_t*>(slot + RingGeo<BodyQ5_0>::OFF_X + (c * 8 + r) * 4);
  }
}
__device__ __forceinline__ void ring_lds(BodyQ4_K::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r, int c4) {
  if (t < nc) {
    f.h = *reinterpret_cast<const uint4*>(slot + RingGeo<BodyQ4_K>::OFF_X + (t * 8 + r) * 16);
    const uint8_t* q = slot + ((t * 4 + c4) * 16 + r) * 16;
    f.qa = *reinterpret_cast<const uint4*>(q);
    f.qb = *reinterpret_cast<const uint4*>(q + 128);
  }
}
__device__ __forceinline__ void ring_lds(BodyQ6_K::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r, int sub) {
  if (t < nc) {
    const uint8_t* q = slot + ((t * 12 + sub) * 8 + r) * 16;
    f.qa = *reinterpret_cast<const uint4*>(q);
    f.qb = *reinterpret_cast<const uint4*>(q + 512);
    f.qh = *reinterpret_cast<const uint4*>(q + 1024);
    f.sc = *reinterpret_cast<const uint2*>(slot + RingGeo<BodyQ6_K>::OFF_X + ((t * 8 + r) * 2 + (sub >> 1)) * 8);
    f.d = *reinterpret_cast<const uint16_t*>(slot + RingGeo<BodyQ6_K>::OFF_D + (t * 8 + r) * 2);
  }
}
template <bool BF>
__device__ __forceinline__ void ring_lds(typename BodyHalf<BF>::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r,
                                         int sub) {
  const uint32_t c = 4 * t + sub;
  if (c < nc) f.w = *reinterpret_cast<const uint4*>(slot + (c * 8 + r) * 16);
}
template <class B>
struct RingLds {
  __device__ __forceinline__ static void run(typename B::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc, int r, int sub) {
    ring_lds(f, slot, t, nc, r, sub);
  }
};
template <bool BF>
struct RingLds<BodyHalf<BF>> {
  __device__ __forceinline__ static void run(typename BodyHalf<BF>::Frag& f, const uint8_t* slot, uint32_t t, uint32_t nc,
                                             int r, int sub) {
    ring_lds<BF>(f, slot, t, nc, r, sub);
  }
};

// Optional prologue of the ring kernel: the RMSNorm stage that produces its activation (norm_act_kernel's arithmetic:
// h' = h + (rms_scale(y) * y) * w_post; xn = (rms_scale(h') * h') * w; Q8_0(xn)), run by EVERY CTA on its own copy.
struct RingNorm {
  const float* y = nullptr;       // output of the previous mat-vec
  const float* w_post = nullptr;  // its post-norm weight (nullptr: h' = h + y)
  const float* h_in = nullptr;    // residual stream before the stage
  float* h_out = nullptr;         // ... after it: ANOTHER buffer (CTA 0 writes it while the other CTAs still read h_in)
  const float* w = nullptr;       // norm weight of the stage that feeds the mat-vec
  uint32_t n = 0, T = 0;          // elements; logical threads of norm_act_kernel (512 / 1024)
  double eps = 0.0;
};

struct RingBatch {
  GemvArgs a[GEMV_MAX_BATCH];          // same format, same activation => same K, nb, units, chunks
  uint32_t slab_end[GEMV_MAX_BATCH];   // exclusive prefix of slabs per matrix
  int n;
  uint32_t total;   // work items of the launch = (slabs of all matrices) x chunks
  uint32_t part_items;  // capacity of the chunk-partial array of a CTA (items)
  uint32_t late_fill;   // 1: do not touch HBM before griddepcontrol.wait (A/B knob)
  RingNorm norm;        // PRO instantiations only
  uint2* fix;       // [gridDim.x][chunks][8] flagged chunk partials of slabs split across CTAs; all zero between launches
  LLPeers peers;
  LLTag tag;
};

__device__ __forceinline__ void ring_fix_store(uint2* p, float v) {
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(1u) : "memory");
}
__device__ __forceinline__ float ring_fix_take(uint2* p) {
  uint32_t v, f;
  for (;;) {
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
    if (f) break;
    __nanosleep(20);
  }
  asm volatile("st.relaxed.gpu.global.v2.u32 [%0], {%1, %1};" ::"l"(p), "r"(0u) : "memory");
  return __uint_as_float(v);
}

// shared-memory accesses by 32-bit shared address (no generic-address arithmetic in the item loop)
__device__ __forceinline__ uint4 lds128(uint32_t a) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t a) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t lds16(uint32_t a) {
  uint16_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(a) : "memory");
  return v;
}
// `bytes` (a multiple of 16, <= MAXB) from global to shared by the 32 lanes of a warp, 16 bytes per lane per step;
// dst / src already include lane * 16
template <uint32_t MAXB, bool FULL>
__device__ __forceinline__ void ring_copy(uint32_t dst, const uint8_t* src, uint32_t bytes, int lane) {
#pragma unroll
  for (uint32_t i = 0; i < (MAXB + 511u) / 512u; ++i) {
    const uint32_t off = i * 512u;
    if (FULL ? (MAXB % 512u == 0 || off + 512u <= MAXB || off + uint32_t(lane) * 16u < MAXB) : (off + uint32_t(lane) * 16u < bytes))
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + off), "l"(src + off) : "memory");
  }
}

// ---- one work item out of a ring slot ---------------------------------------------------------------------------
// Generic form: the shared-memory twin of load_item + the format's own compute (Body::compute), any format.
template <class B, bool FULL>
struct RingItem {
  __device__ __forceinline__ static float run(const uint8_t* slot, uint32_t, const GemvArgs& a, const uint8_t* sm_act, uint32_t,
                                              uint32_t j, uint32_t nc, int lane) {
    constexpr int N = B::C;
    const int r = lane & 7, sub = lane >> 3;
    FragSet<B, N> f;
#pragma unroll
    for (int t = 0; t < N; ++t)
      if (j * N + t < a.units) RingLds<B>::run(f.f[t], slot, t, nc, r, sub);
    return compute_item<B, N>(f, a, sm_act, j, sub);
  }
};
// Q4_0, the headline format, hand-scheduled: BodyQ4_0::compute operation for operation in the floating-point part
// (dw * dx, one fma per block, blocks 4t + sub in order, xor-shuffle tree) and exact in the integer part, with the
// instruction count cut to what an issue-bound kernel can afford (the mat-vec runs at 2.5 of 4 issue slots per
// clock: profiles/r02_notes.md): high nibbles stay in place (w & 0xf0f0f0f0 as unsigned bytes: dp4a gives 16 x the
// block's high-half dot, an exact multiple of 16), shared memory is addressed by 32-bit shared addresses with the
// unit offsets as immediates, and full items carry no bounds predicates.
template <bool FULL>
struct RingItem<BodyQ4_0, FULL> {
  __device__ __forceinline__ static float run(const uint8_t*, uint32_t slot_s, const GemvArgs& a, const uint8_t*, uint32_t act_s,
                                              uint32_t j, uint32_t nc, int lane) {
    const int sub = lane >> 3;
    const uint32_t wq = slot_s + uint32_t(lane) * 16u, wd = slot_s + RingGeo<BodyQ4_0>::OFF_D + uint32_t(lane) * 2u;
    const uint32_t xq = act_s + (j * 16u + uint32_t(sub)) * 32u, xm = act_s + a.n_cols + (j * 16u + uint32_t(sub)) * 4u;
    float acc = 0.0f;
#pragma unroll
    for (uint32_t t = 0; t < 4; ++t) {
      if (FULL || 4u * t + uint32_t(sub) < nc) {
        const uint4 w = lds128(wq + t * 512u);
        const uint32_t dw = lds16(wd + t * 64u);
        const uint4 xa = lds128(xq + t * 128u), xb = lds128(xq + t * 128u + 16u);
        const uint32_t m = lds32(xm + t * 16u);
        int lo = __dp4a(int(w.x & 0x0f0f0f0fu), int(xa.x), 0);
        lo = __dp4a(int(w.y & 0x0f0f0f0fu), int(xa.y), lo);
        lo = __dp4a(int(w.z & 0x0f0f0f0fu), int(xa.z), lo);
        lo = __dp4a(int(w.w & 0x0f0f0f0fu), int(xa.w), lo);
        int hi = dp4a_us(w.x & 0xf0f0f0f0u, int(xb.x), 0);
        hi = dp4a_us(w.y & 0xf0f0f0f0u, int(xb.y), hi);
        hi = dp4a_us(w.z & 0xf0f0f0f0u, int(xb.z), hi);
        hi = dp4a_us(w.w & 0xf0f0f0f0u, int(xb.w), hi);
        const int dot = lo + (hi >> 4) - 8 * int(int16_t(m >> 16));
        acc = fmaf(h2f(uint16_t(dw)) * h2f(uint16_t(m & 0xffffu)), float(dot), acc);  // ops.cpp:380-395
      }
    }
    acc += __shfl_xor_sync(0xffffffffu, acc, 8);
    acc += __shfl_xor_sync(0xffffffffu, acc, 16);
    return acc;
  }
};

// (S, j) of a warp's item sequence, advanced by W items per step without a division; the matrix of the batch that
// holds slab S is tracked alongside (it changes at most twice per CTA)
struct RingCursor {
  uint32_t S, j;     // linear slab over the batch, K-chunk
  uint32_t s_begin, s_end;  // slab range of the current matrix
  int mi;
  __device__ __forceinline__ void init(const RingBatch& b, uint32_t g, uint32_t J) {
    S = g / J;
    j = g - S * J;
    mi = 0;
    s_begin = 0;
    s_end = b.slab_end[0];
    seek(b);
  }
  __device__ __forceinline__ void seek(const RingBatch& b) {
    while (S >= s_end && mi + 1 < b.n) {
      ++mi;
      s_begin = s_end;
      s_end = b.slab_end[mi];
    }
  }
  template <int W>
  __device__ __forceinline__ void advance(const RingBatch& b, uint32_t J) {
    j += W;
    while (j >= J) {
      j -= J;
      ++S;
    }
    if (S >= s_end) seek(b);
  }
};

constexpr int RING_NORM_PER = 6;  // elements per logical thread (norm_act_kernel's NORM_PER)
// norm_act_kernel's arithmetic for the T = 512 / 1024 LOGICAL threads of that kernel, played by the W * 32 physical
// threads of this CTA (thread t plays logical threads t, t + W * 32, ...): the same per-thread sums, the same
// xor-shuffle tree per logical warp, the T/32 warp sums left to right, the same quantizer per 32-element block —
// bit for bit the activation the separate kernel would have left in global memory, written to `act` (shared memory).
template <int W>
__device__ __forceinline__ void ring_norm_prologue(const RingNorm& nm, uint8_t* act, float* red1, float* red2) {
  constexpr int PT = W * 32, MAXP = 1024 / PT;
  const uint32_t n = nm.n, T = nm.T;
  const int NP = int(T) / PT, NW = int(T) / 32;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float yv[MAXP][RING_NORM_PER], hv[MAXP][RING_NORM_PER];
#pragma unroll
  for (int p = 0; p < MAXP; ++p)
#pragma unroll
    for (int k = 0; k < RING_NORM_PER; ++k) {
      const uint32_t i = threadIdx.x + uint32_t(p) * PT + uint32_t(k) * T;
      const bool ok = p < NP && i < n;
      hv[p][k] = ok ? nm.h_in[i] : 0.0f;
      yv[p][k] = ok ? nm.y[i] : 0.0f;
    }
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < RING_NORM_PER; ++k) ss += __fmul_rn(yv[p][k], yv[p][k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0 && p < NP) red1[warp + p * W] = ss;
  }
  __syncthreads();
  {
    float tot = 0.0f;
    for (int i = 0; i < NW; ++i) tot += red1[i];
    const float sc = rms_scale(tot, n, nm.eps);
#pragma unroll
    for (int p = 0; p < MAXP; ++p)
#pragma unroll
      for (int k = 0; k < RING_NORM_PER; ++k) {
        const uint32_t i = threadIdx.x + uint32_t(p) * PT + uint32_t(k) * T;
        const bool ok = p < NP && i < n;
        const float add = nm.w_post ? __fmul_rn(__fmul_rn(sc, yv[p][k]), ok ? nm.w_post[i] : 0.0f) : yv[p][k];
        hv[p][k] = __fadd_rn(hv[p][k], add);
        if (ok && blockIdx.x == 0) nm.h_out[i] = hv[p][k];
      }
  }
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    float ss = 0.0f;
#pragma unroll
    for (int k = 0; k < RING_NORM_PER; ++k) ss += __fmul_rn(hv[p][k], hv[p][k]);
#pragma unroll
    for (int o = 16; o; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    if (lane == 0 && p < NP) red2[warp + p * W] = ss;
  }
  __syncthreads();
  float tot = 0.0f;
  for (int i = 0; i < NW; ++i) tot += red2[i];
  const float sc = rms_scale(tot, n, nm.eps);
#pragma unroll
  for (int p = 0; p < MAXP; ++p) {
    if (p >= NP) break;
    float xv[RING_NORM_PER];
    bool live[RING_NORM_PER];
    uint32_t blk[RING_NORM_PER];
    const uint32_t lw = uint32_t(warp + p * W);  // logical warp: pass k holds block k * NW + lw
#pragma unroll
    for (int k = 0; k < RING_NORM_PER; ++k) {
      const uint32_t i = threadIdx.x + uint32_t(p) * PT + uint32_t(k) * T;
      xv[k] = __fmul_rn(__fmul_rn(sc, hv[p][k]), i < n ? nm.w[i] : 0.0f);
      live[k] = lw * 32u + uint32_t(k) * T < n;
      blk[k] = uint32_t(k) * uint32_t(NW) + lw;
    }
    warp_quantize_q8_0_multi<RING_NORM_PER>(xv, live, blk, n, act, lane);
  }
  __syncthreads();
}

template <class B, int W, int D, bool PUSH, bool PRO = false>
// (minimum CTAs per SM = 1024 threads' worth: caps the kernel at 64 registers so that two 16-warp or four 8-warp CTAs fit)
__global__ void __launch_bounds__(W * 32, 1024 / (W * 32)) gemv_ring_kernel(const RingBatch batch) {
  using G = RingGeo<B>;
  using F = RingFmt<B>;
  extern __shared__ __align__(128) uint8_t smem[];  // [W*D slots][activation][chunk partials of this CTA]
  __shared__ __align__(8) uint64_t bars[1];         // activation staging
  __shared__ float red1[PRO ? 32 : 1], red2[PRO ? 32 : 1];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();
  TL_ENTER(2);
  const GemvArgs& a0 = batch.a[0];
  const uint32_t J = a0.chunks, nb = a0.nb;
  const uint32_t smem_s = smem_u32(smem);
  const uint32_t ring_off = uint32_t(warp) * D * G::SLOT, act_off = uint32_t(W) * D * G::SLOT;
  uint8_t* sm_act = smem + act_off;
  float* part = reinterpret_cast<float*>(sm_act + ((a0.act_bytes + 127u) & ~127u));
  if (threadIdx.x == 0) mbar_init(&bars[0], 1);
  // This CTA's items [g0, g1).  A slab belongs to the CTA that holds its LAST chunk, so a CTA only ever waits for CTAs
  // with a SMALLER index — which the hardware starts first and which wait for nobody behind them: forward progress
  // does not depend on the whole grid being resident (the look-back argument of single-pass scans).  The chunks of a
  // trailing slab this CTA does not own go FIRST in its sequence and are handed over as soon as they are computed;
  // sequence position i -> item: i < n_tail ? tail0 + i : g0 + (i - n_tail).
  const uint32_t g0 = uint32_t(uint64_t(batch.total) * blockIdx.x / gridDim.x);
  const uint32_t g1 = uint32_t(uint64_t(batch.total) * (blockIdx.x + 1) / gridDim.x);
  const uint32_t n_my = g1 - g0;
  uint32_t n_tail = 0, tail0 = g1, owner = 0;
  if (n_my) {
    const uint32_t sb = (g1 - 1) / J, last = sb * J + J - 1;  // last chunk of the last slab touched
    if (last >= g1) {
      tail0 = max(g0, sb * J);
      n_tail = g1 - tail0;
      owner = uint32_t((uint64_t(last + 1) * gridDim.x - 1) / batch.total);  // the CTA whose range holds item `last`
    }
  }
  auto item_of = [&](uint32_t i) -> uint32_t { return i < n_tail ? tail0 + i : g0 + (i - n_tail); };
  __syncthreads();
  // all lanes: asynchronous copies of the item under cursor c (if any) into slot s; always one commit group
  auto fill = [&](const RingCursor& c, bool live, int s) {
    if (live) {
      const GemvArgs& a = batch.a[c.mi];
      const uint32_t c0 = c.j * F::CELLS, nc = min(F::CELLS, nb - c0);
      const size_t cell = size_t(c.S - c.s_begin) * nb + c0;
      const uint32_t dst = smem_s + ring_off + uint32_t(s) * G::SLOT + uint32_t(lane) * 16u;
      const uint32_t lo = uint32_t(lane) * 16u;
      if (nc == F::CELLS) {
        ring_copy<F::CELLS * F::QB, true>(dst, a.q + cell * F::QB + lo, 0, lane);
        if (F::DB) ring_copy<F::CELLS * F::DB, true>(dst + G::OFF_D, a.d + cell * F::DB + lo, 0, lane);
        if (F::XB) ring_copy<F::CELLS * F::XB, true>(dst + G::OFF_X, a.x + cell * F::XB + lo, 0, lane);
      } else {
        ring_copy<F::CELLS * F::QB, false>(dst, a.q + cell * F::QB + lo, nc * F::QB, lane);
        if (F::DB) ring_copy<F::CELLS * F::DB, false>(dst + G::OFF_D, a.d + cell * F::DB + lo, nc * F::DB, lane);
        if (F::XB) ring_copy<F::CELLS * F::XB, false>(dst + G::OFF_X, a.x + cell * F::XB + lo, nc * F::XB, lane);
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
  RingCursor cf, cc;  // fill cursor (runs D rounds ahead) and consume cursor
  uint32_t i_f = warp;
  if (uint32_t(warp) < n_my) {
    cf.init(batch, item_of(warp), J);
    cc = cf;
  }
  // sequence position i_old -> i_old + W: one step of the cursor, except across the jump from the tail items back to g0
  auto step = [&](RingCursor& c, uint32_t i_new) {
    if (i_new - W < n_tail && i_new >= n_tail) c.init(batch, item_of(i_new), J);
    else c.template advance<W>(batch, J);
  };
  if (batch.late_fill) pdl_wait();  // A/B knob (LLMI_RING_LATE=1): rings filled only once the predecessor has finished
#pragma unroll
  for (int s = 0; s < D; ++s) {
    fill(cf, i_f < n_my, s);
    i_f += W;
    if (i_f < n_my) step(cf, i_f);
  }
  // (An L2 prefetch of the items BEHIND the rings — cp.async.bulk.prefetch.L2 per item plane, issued here — was
  // measured and removed: the decode step got slower, 4.92 -> 5.13 / 5.42 ms at 32 / 96 items per CTA,
  // profiles/r02_ab_ring_prefetch.txt.)
  // the activation vector is the predecessor's output
  pdl_wait();
  TL_MARK(1);
  if constexpr (PRO) {
    ring_norm_prologue<W>(batch.norm, sm_act, red1, red2);
  } else {
    // (plain 16-byte loads by all threads + a barrier instead of the bulk copy were measured: slower — 27b gate 12.9 -> 13.0 us,
    // down 12.9 -> 13.6 us, step 4.39 -> 4.45 ms)
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bars[0], a0.act_bytes);
      bulk_g2s(sm_act, a0.act, a0.act_bytes, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
  }
  TL_MARK(3);
  const uint32_t act_s = smem_s + act_off;
  int s = 0;
#pragma unroll 1
  for (uint32_t i = warp; i < n_my; i += W) {
    const GemvArgs& a = batch.a[cc.mi];
    const uint32_t j = cc.j;
    const uint32_t nc = min(F::CELLS, nb - j * F::CELLS);
    const uint32_t slot_off = ring_off + uint32_t(s) * G::SLOT;
    asm volatile("cp.async.wait_group %0;" ::"n"(D - 1) : "memory");  // this lane's copies of item i have landed
    __syncwarp();                                                      // ... and every other lane's
    float v;
    if (nc == F::CELLS) v = RingItem<B, true>::run(smem + slot_off, smem_s + slot_off, a, sm_act, act_s, j, nc, lane);
    else v = RingItem<B, false>::run(smem + slot_off, smem_s + slot_off, a, sm_act, act_s, j, nc, lane);
    // (the shuffles of the item's reduction are past every lane's reads of the slot: it can be refilled)
    fill(cf, i_f < n_my, s);
    i_f += W;
    if (i_f < n_my) step(cf, i_f);
    if (lane < LLMI_SLAB) {
      part[(cc.S * J + j - g0) * LLMI_SLAB + lane] = v;
      if (i < n_tail) ring_fix_store(batch.fix + (size_t(owner) * J + j) * LLMI_SLAB + lane, v);
    }
    if (i + W < n_my) step(cc, i + W);
    s = s + 1 == D ? 0 : s + 1;
  }
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  TL_MARK(4);
  // the leading chunks of this CTA's first slab may have been computed by earlier CTAs (FIRST in their sequences): every
  // missing (chunk, row) word is fetched by its own thread — one round trip for all of them — into the partials array
  // behind this CTA's own items
  const uint32_t Sa = g0 / J, lead = g0 - Sa * J;  // first slab touched, its chunks computed elsewhere
  const bool owns = n_my > n_tail;                 // the items in front of tail0 end on a slab boundary
  if (owns && lead) {
    uint2* fp = batch.fix + size_t(blockIdx.x) * J * LLMI_SLAB;
    for (uint32_t idx = threadIdx.x; idx < lead * LLMI_SLAB; idx += W * 32) part[n_my * LLMI_SLAB + idx] = ring_fix_take(fp + idx);
  }
  __syncthreads();
  TL_MARK(5);
  // rows of the slabs this CTA owns: chunk partials left to right, the canonical order
  unsigned long long best = 0;
  const uint32_t tag = PUSH ? ll_tag(batch.tag) : 0u;
  if (owns) {
    const uint32_t n_own = (tail0 - 1) / J - Sa + 1;
    for (uint32_t idx = threadIdx.x; idx < n_own * LLMI_SLAB; idx += W * 32) {
      const uint32_t S = Sa + idx / LLMI_SLAB, rr = idx % LLMI_SLAB;
      const uint32_t nl = S == Sa ? lead : 0u;  // chunks of this slab that sit in the collected region
      const float* pl = part + size_t(n_my) * LLMI_SLAB + rr;
      const float* p = part + (ptrdiff_t(S * J) - ptrdiff_t(g0)) * LLMI_SLAB + rr;  // dereferenced from chunk nl on
      float sum = nl ? pl[0] : p[0];
      for (uint32_t j = 1; j < nl; ++j) sum += pl[j * LLMI_SLAB];
      uint32_t j = nl ? nl : 1u;
      for (; j + 8 <= J; j += 8) {  // eight independent loads, then the left-to-right chain (J is up to 64: the loads of a
        float v[8];                 // rolled loop would each wait for the previous add)
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = p[(j + u) * LLMI_SLAB];
#pragma unroll
        for (int u = 0; u < 8; ++u) sum += v[u];
      }
      for (; j < J; ++j) sum += p[j * LLMI_SLAB];
      int mi = 0;
      while (mi + 1 < batch.n && S >= batch.slab_end[mi]) ++mi;
      const GemvArgs& a = batch.a[mi];
      const uint32_t row = (S - (mi ? batch.slab_end[mi - 1] : 0u)) * LLMI_SLAB + rr;
      if (row < a.n_local) {
        if (a.argmax_key) {  // logits epilogue: see gemv_slab_kernel
          if (a.softcap > 0.0f) sum = __fmul_rn(a.softcap, tanhf(__fdiv_rn(sum, a.softcap)));
          uint32_t u = __float_as_uint(sum);
          u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
          const unsigned long long key = (uint64_t(u) << 32) | uint32_t(0xffffffffu - (a.row0 + row));
          best = key > best ? key : best;
        }
        if (PUSH) {
          for (uint32_t p2 = 0; p2 < batch.peers.n; ++p2)
            ll_store(batch.peers.base[p2] + a.ll_off + a.row0 + row, __float_as_uint(sum), tag);
        } else {
          a.out[row] = sum;
        }
      }
    }
  }
  if (a0.argmax_key) {
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0 && best) atomicMax(a0.argmax_key, best);
  }
  TL_MARK(2);
}
