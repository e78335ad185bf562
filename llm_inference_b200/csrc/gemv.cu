// Decode mat-vec kernels for sm_100a — the replacement for the compute_range
// bodies + thread-pool fan-out of ops.cpp:188-931.
//
// One kernel shape for every format:
//   * data unit = one SLAB of 8 output rows; lane = 8*sub + row, so one 128-bit
//     load per lane covers 4 K-units x 8 rows = 512 contiguous bytes of the
//     quant plane (weights are read exactly once, straight into registers,
//     L1 no-allocate; no shared-memory round trip for weights);
//   * work item = (slab, K-chunk): 512 elements of K for the block formats,
//     128 for F16/BF16.  A CTA of W warps owns S consecutive slabs and deals
//     their items round-robin to its warps — the reference's row partition
//     over threads (ops.cpp:439-448) becomes a (row-slab, K-chunk) partition
//     over the grid;
//   * canonical summation: a row's result is the left-to-right sum over its
//     K-chunks of (xor-shuffle tree over the 4 sub-lanes of (sequential sum
//     over the chunk's units)).  The order depends only on K and the format,
//     so results are deterministic and independent of W, S, the grid, the SM
//     count and of row sharding across GPUs (bit-identical shards);
//   * the quantized activation vector (K bytes + scales) is staged into shared
//     memory once per CTA with one bulk async copy (cp.async.bulk -> UBLKCP)
//     completing on an mbarrier, overlapped with the first weight loads;
//   * integer block dots are __dp4a, bit-exact with the reference's per-block
//     sums; the fp32 scale product and accumulation follow the reference's
//     formulas (summation ORDER differs: that is the documented 1e-5 bound).
#include <cuda.h>  // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint)
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <algorithm>
#include <map>
#include <mutex>
#include <string>
#include <type_traits>

#include "launch.cuh"
#include "llmi_internal.h"

#include "gemv_bodies.cuh"
#include "glue_device.cuh"  // geglu(), the warp quantizers

namespace {

// CTA = W warps over S consecutive slabs.  The S*J work items (slab, K-chunk)
// are dealt round-robin to the warps; each warp loads one item (C units = 4
// independent 128-bit loads per lane) per round trip, folds it, and drops the
// chunk partial (8 floats) in shared memory.  After one barrier, thread (s, r)
// adds the J chunk partials of its row in canonical order.  Bytes in flight
// come from occupancy (small register footprint, many resident CTAs), not from
// a deep per-warp pipeline: a warp has only 6 scoreboard slots, and a second
// generation of loads in flight makes ptxas share them, which turns the
// address setup of the next loads into waits on unrelated loads (measured 2x
// slower, profiles/r01_notes.md).
// Up to GEMV_MAX_BATCH matrices of the same format that consume the same
// activation vector (q/k/v, gate/up) go out as ONE grid: the CTA index selects
// the matrix.  The launch floor (~2.3 us) is paid once and the small matrices
// fill the SMs together.
constexpr int GEMV_MAX_BATCH = 3;
struct GemvBatch {
  GemvArgs a[GEMV_MAX_BATCH];
  uint32_t cta_end[GEMV_MAX_BATCH];  // exclusive prefix of CTAs per matrix
  uint32_t S[GEMV_MAX_BATCH];        // slabs per CTA
  int n;
  LLPeers peers;  // exchange buffers of all ranks (launch.cuh); used by matrices with ll_off != LL_NONE
  LLTag tag;
};

#ifdef LLMI_GEMV_TIMING  // dev only (tools/gemv_chain_bench.py): %globaltimer stamps of CTA 0 / thread 0 per launch
__device__ unsigned long long g_gemv_stamp[256][6];
__device__ unsigned int g_gemv_launch;
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define GEMV_STAMP(slot, i) do { if (blockIdx.x == 0 && threadIdx.x == 0) g_gemv_stamp[(slot) & 255][i] = gtime(); } while (0)
#else
#define GEMV_STAMP(slot, i) do { } while (0)
#endif
#ifdef LLMI_TIMELINE
#define TL_SLOT tl_slot
#else
#define TL_SLOT (-1)
#endif

// Called by every thread after its first weight loads are in flight: wait for
// the predecessor grid (PDL), let thread 0 start the bulk copy of the activation
// vector it produced, wait for the bytes to land.
__device__ __forceinline__ void stage_activation(const GemvArgs& a, uint8_t* sm_act, uint64_t* bar, unsigned slot = 0,
                                                 int tl_slot = -1) {
  pdl_wait();
  TL_MARK(1);
  GEMV_STAMP(slot, 1);
  if (threadIdx.x == 0) {
    mbar_expect_tx(bar, a.act_bytes);
    bulk_g2s(sm_act, a.act, a.act_bytes, bar);
  }
  mbar_wait(bar, 0);
  TL_MARK(3);
  GEMV_STAMP(slot, 2);
}

// PUSH: row-sharded model — the rows go, flagged, into every rank's exchange buffer instead of `out` (its own
// instantiation: the peer table and the tag would cost the single-GPU kernel registers, i.e. resident warps).
// (Two items in flight per warp — 8 instead of 4 128-bit loads per lane — were measured in round 2 and are slower on
// every shape but one: 27b gate 14.3 -> 16.0 us, 4b gate 4.7 -> 5.7 us, profiles/r02_sweep_unroll.txt.  Occupancy, not
// per-warp depth, is what feeds HBM here.)
template <class B, int W, bool PUSH>
__global__ void __launch_bounds__(W * 32) gemv_slab_kernel(const GemvBatch batch) {
  extern __shared__ __align__(128) uint8_t sm_act[];  // [act_bytes][S * chunks * 8 floats]
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane & 7, sub = lane >> 3;
  int mi = 0;
  while (mi + 1 < batch.n && blockIdx.x >= batch.cta_end[mi]) ++mi;
  const GemvArgs& a = batch.a[mi];
  const uint32_t S = batch.S[mi];
  const uint32_t cta = blockIdx.x - (mi ? batch.cta_end[mi - 1] : 0u);
  pdl_trigger();
  TL_ENTER(1);
#ifdef LLMI_GEMV_TIMING
  __shared__ unsigned slot_s;
  if (blockIdx.x == 0 && threadIdx.x == 0) slot_s = atomicAdd(&g_gemv_launch, 1u);
#define SLOT (slot_s)
#else
#define SLOT 0u
#endif
  GEMV_STAMP(SLOT, 0);
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  constexpr int N = B::C;
  const uint32_t J = a.chunks;
  const uint32_t slab0 = cta * S;
  const uint32_t n_sl = min(S, a.n_slabs - slab0);
  const uint32_t n_items = n_sl * J;
  float* part = reinterpret_cast<float*>(sm_act + a.act_bytes);
  bool waited = false;
#pragma unroll 1
  for (uint32_t t = warp; t < n_items; t += W) {
    const uint32_t sl = t / J, j = t - sl * J;
    FragSet<B, N> f;
    load_item<B, N>(f, a, slab0 + sl, j, r, sub);  // weights: independent of the predecessor kernel
    if (!waited) {
      stage_activation(a, sm_act, &bar, SLOT, TL_SLOT);
      waited = true;
    }
    const float v = compute_item<B, N>(f, a, sm_act, j, sub);
    if (lane < LLMI_SLAB) part[t * LLMI_SLAB + lane] = v;
  }
  if (!waited) stage_activation(a, sm_act, &bar, SLOT, TL_SLOT);  // never exit with the bulk copy into our smem in flight
  __syncthreads();
  TL_MARK(5);
  GEMV_STAMP(SLOT, 3);
  unsigned long long best = 0;
  const uint32_t tag = PUSH ? ll_tag(batch.tag) : 0u;  // every thread is past pdl_wait here
  for (uint32_t idx = threadIdx.x; idx < n_sl * LLMI_SLAB; idx += W * 32) {
    const uint32_t sl = idx / LLMI_SLAB, rr = idx % LLMI_SLAB;
    const float* p = part + (size_t)sl * J * LLMI_SLAB + rr;
    float sum = p[0];
    for (uint32_t j = 1; j < J; ++j) sum += p[j * LLMI_SLAB];  // canonical order
    const uint32_t row = (slab0 + sl) * LLMI_SLAB + rr;
    if (row < a.n_local) {
      if (a.argmax_key) {  // logits: final soft-cap (model.cpp:1036-1041) + greedy argmax (main.cpp:193)
        if (a.softcap > 0.0f) sum = __fmul_rn(a.softcap, tanhf(__fdiv_rn(sum, a.softcap)));
        // order-preserving float -> uint; ties go to the smaller row (std::max_element)
        uint32_t u = __float_as_uint(sum);
        u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
        const unsigned long long k = (uint64_t(u) << 32) | uint32_t(0xffffffffu - (a.row0 + row));
        best = k > best ? k : best;
      }
      if (PUSH) {  // the all-gather IS this store: one flagged 64-bit write into every rank's copy of the vector
        for (uint32_t p = 0; p < batch.peers.n; ++p)
          ll_store(batch.peers.base[p] + a.ll_off + a.row0 + row, __float_as_uint(sum), tag);
      } else {
        a.out[row] = sum;
      }
    }
  }
  if (a.argmax_key) {  // one atomicMax per warp that saw rows (max is order-independent: deterministic)
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
      best = other > best ? other : best;
    }
    if (lane == 0 && best) atomicMax(a.argmax_key, best);
  }
  GEMV_STAMP(SLOT, 4);
  TL_MARK(2);
#undef SLOT
}


// ------------------------------------------------------------------ L2 prefetch
// The first `lines` 128-byte lines of up to 6 plane ranges into L2.  Launched on a side stream while a glue kernel
// (attention, norm, GEGLU: a handful of CTAs, no HBM traffic to speak of) holds the main stream, so that the HBM
// pipe works through the gap and the next mat-vec finds the front of its weights in L2 (model.cu run_step).
struct PrefetchArgs {
  const uint8_t* p[6];
  uint32_t lines[6];
  int n;
};
__global__ void l2_prefetch_kernel(const PrefetchArgs a) {
  const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int i = 0; i < a.n; ++i)
    for (uint32_t l = tid; l < a.lines[i]; l += nth)
      asm volatile("prefetch.global.L2 [%0];" ::"l"(a.p[i] + size_t(l) * 128) : "memory");
}

// ---------------------------------------------------------------- gate/up + GEGLU
// ffn_gate and ffn_up (model.cpp:875, 877) consume the same vector and GEGLU (model.cpp:887-901) combines their rows
// pairwise, so a CTA that owns the SAME slabs of both matrices can finish the stage: it owns 4 consecutive slabs =
// 32 rows of gate and of up (items: 8 slabs x J chunks, dealt to the warps as in gemv_slab_kernel), sums every row
// in the canonical order, and one warp then computes hidden = gelu(gate) * up for the 32 rows and
//   * writes them as one Q8_0 block of the activation of ffn_down (32 rows = one block: the quantizer of
//     geglu_act_kernel, bit for bit), so the GEGLU launch and its round trip of two F-vectors disappear, or
//   * PUSH (row-sharded model): stores them flagged into every rank's exchange buffer — half the words the separate
//     gate and up vectors took.
template <class B, int W, bool PUSH>
__global__ void __launch_bounds__(W * 32) gemv_geglu_kernel(const GemvBatch batch, uint8_t* act_out, uint32_t act_n) {
  extern __shared__ __align__(128) uint8_t sm_act[];  // [act_bytes][2 matrices * 4 slabs * chunks * 8 floats]
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane & 7, sub = lane >> 3;
  const GemvArgs& ag = batch.a[0];  // gate; batch.a[1] = up (same shape, same format)
  pdl_trigger();
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  constexpr int N = B::C;
  const uint32_t J = ag.chunks;
  const uint32_t slab0 = blockIdx.x * 4;
  const uint32_t n_sl = min(4u, ag.n_slabs - slab0);
  const uint32_t n_items = 2 * n_sl * J;  // item t: matrix t / (n_sl * J), slab, chunk
  float* part = reinterpret_cast<float*>(sm_act + ag.act_bytes);
  bool waited = false;
#pragma unroll 1
  for (uint32_t t = warp; t < n_items; t += W) {
    const uint32_t mi = t / (n_sl * J), rem = t - mi * (n_sl * J), sl = rem / J, j = rem - sl * J;
    const GemvArgs& a = batch.a[mi];
    FragSet<B, N> f;
    load_item<B, N>(f, a, slab0 + sl, j, r, sub);  // weights: independent of the predecessor kernel
    if (!waited) {
      stage_activation(ag, sm_act, &bar);
      waited = true;
    }
    const float v = compute_item<B, N>(f, a, sm_act, j, sub);
    if (lane < LLMI_SLAB) part[t * LLMI_SLAB + lane] = v;
  }
  if (!waited) stage_activation(ag, sm_act, &bar);
  __syncthreads();
  if (warp != 0) return;
  const uint32_t sl = lane >> 3, rr = lane & 7;
  const uint32_t row = (slab0 + sl) * LLMI_SLAB + rr;
  float hidden = 0.0f;
  if (sl < n_sl) {
    const float* pg = part + size_t(sl) * J * LLMI_SLAB + rr;
    const float* pu = pg + size_t(n_sl) * J * LLMI_SLAB;
    float g = pg[0], u = pu[0];
    for (uint32_t j = 1; j < J; ++j) {  // canonical order
      g += pg[j * LLMI_SLAB];
      u += pu[j * LLMI_SLAB];
    }
    if (row < ag.n_local) hidden = geglu(g, u);
  }
  if (PUSH) {
    const uint32_t tag = ll_tag(batch.tag);
    if (sl < n_sl && row < ag.n_local)
      for (uint32_t p = 0; p < batch.peers.n; ++p)
        ll_store(batch.peers.base[p] + ag.ll_off + ag.row0 + row, __float_as_uint(hidden), tag);
  } else {
    warp_quantize_q8_0(hidden, blockIdx.x, act_n, act_out, lane);  // rows 32c .. 32c+31 = Q8_0 block c of ffn_down's input
  }
}

// ------------------------------------------------------ token-batched (prefill)
// The same work decomposition for M tokens at once: a warp loads the weights
// of its item ONCE and folds them against the activations of MT tokens staged
// in shared memory (one bulk copy per token tile: the M activation buffers are
// contiguous).  Per (token, row) the arithmetic and the summation order are
// exactly those of gemv_slab_kernel, so a prompt processed in batches gives the
// bits the token-by-token path gives.  Weights are re-read once per token tile
// (from L2 when a layer fits).
template <class B, int W>
__global__ void __launch_bounds__(W * 32) gemv_slab_tok_kernel(const GemvBatch batch, const uint32_t MT) {
  extern __shared__ __align__(128) uint8_t sm_act[];  // [MT][act_stride] then [MT][S * chunks * 8 floats]
  __shared__ __align__(8) uint64_t bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = lane & 7, sub = lane >> 3;
  int mi = 0;
  while (mi + 1 < batch.n && blockIdx.x >= batch.cta_end[mi]) ++mi;
  const GemvArgs& a = batch.a[mi];
  const uint32_t S = batch.S[mi];
  const uint32_t cta = blockIdx.x - (mi ? batch.cta_end[mi - 1] : 0u);
  pdl_trigger();
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  constexpr int N = B::C;
  const uint32_t J = a.chunks;
  const uint32_t slab0 = cta * S;
  const uint32_t n_sl = min(S, a.n_slabs - slab0);
  const uint32_t n_items = n_sl * J;
  float* part = reinterpret_cast<float*>(sm_act + size_t(MT) * a.act_stride);
  uint32_t parity = 0;
  pdl_wait();
  // blockIdx.y strides over the token tiles: small matrices get their parallelism from the tokens
  for (uint32_t m0 = blockIdx.y * MT; m0 < a.n_tok; m0 += gridDim.y * MT) {
    const uint32_t mt = min(MT, a.n_tok - m0);
    if (threadIdx.x == 0) {
      mbar_expect_tx(&bar, mt * a.act_stride);
      bulk_g2s(sm_act, a.act + size_t(m0) * a.act_stride, mt * a.act_stride, &bar);
    }
    bool waited = false;
#pragma unroll 1
    for (uint32_t t = warp; t < n_items; t += W) {
      const uint32_t sl = t / J, j = t - sl * J;
      FragSet<B, N> f;
      load_item<B, N>(f, a, slab0 + sl, j, r, sub);
      if (!waited) {
        mbar_wait(&bar, parity);
        waited = true;
      }
#pragma unroll 1
      for (uint32_t m = 0; m < mt; ++m) {
        const float v = compute_item<B, N>(f, a, sm_act + size_t(m) * a.act_stride, j, sub);
        if (lane < LLMI_SLAB) part[(size_t(m) * n_items + t) * LLMI_SLAB + lane] = v;
      }
    }
    if (!waited) mbar_wait(&bar, parity);
    parity ^= 1;
    __syncthreads();
    for (uint32_t idx = threadIdx.x; idx < mt * n_sl * LLMI_SLAB; idx += W * 32) {
      const uint32_t m = idx / (n_sl * LLMI_SLAB), rem = idx - m * (n_sl * LLMI_SLAB);
      const uint32_t sl = rem / LLMI_SLAB, rr = rem % LLMI_SLAB;
      const float* p = part + (size_t(m) * n_items + size_t(sl) * J) * LLMI_SLAB + rr;
      float sum = p[0];
      for (uint32_t j = 1; j < J; ++j) sum += p[j * LLMI_SLAB];  // canonical order
      const uint32_t row = (slab0 + sl) * LLMI_SLAB + rr;
      if (row < a.n_local) a.out[size_t(m0 + m) * a.out_stride + row] = sum;
    }
    __syncthreads();  // the tile and the partials are reused
  }
}

// ------------------------------------- token-per-lane kernel (Q4_0 / Q8_0 weights)
// For batches of >= 16 tokens the roles flip: a lane owns a TOKEN and keeps that
// token's activation block in registers, the weights of the warp's slab (8 rows)
// sit in shared memory, unpacked to int8 once per K-chunk, and are read by
// broadcast — 2 shared-memory wavefronts serve 32 tokens, and the nibble unpack
// is paid once per 32 tokens.  Per (row, token) the arithmetic is the one of
// BodyQ4_0 / BodyQ8_0 and the summation order is the canonical one: inside a
// K-chunk of 16 blocks four chains s[k] take the blocks b with b % 4 == k in
// order (the four sub-lanes of gemv_slab_kernel), chunk partial = (s0+s1)+(s2+s3)
// (its two xor shuffles), chunk partials added left to right.  Bit-identical to
// the one-token kernel (tests/test_model_gpu.py).
// CTA = W warps = W slabs sharing the activation tile of one group of 32 tokens;
// grid = (slab groups of all matrices of the batch, token groups).
constexpr int TL_ACT_ROW = 528;  // 512 int8 of a chunk + 16 pad: lanes 33 x 16 B apart -> conflict-free LDS.128
constexpr int TL_SC_ROW = 17;    // 16 scale words + 1 pad

template <bool IS_Q8>
struct TokLane {
  // stage blocks [b0, b0+nbk) of slab `slab` into wq (int8 [16][8][32]) and wd (f32(f16 d) [16][8])
  __device__ static void stage_weights(const GemvArgs& a, uint32_t slab, uint32_t b0, uint32_t nbk, uint8_t* wq,
                                       float* wd, int lane) {
    if (IS_Q8) {
      const uint4* src = reinterpret_cast<const uint4*>(a.q) + (size_t(slab) * a.nb + b0) * 16;
      for (uint32_t i = lane; i < nbk * 16; i += 32) {  // item i = (lb*2 + h)*8 + r
        const uint32_t lb = i >> 4, hh = (i >> 3) & 1, r = i & 7;
        *reinterpret_cast<uint4*>(wq + (lb * 8 + r) * 32 + hh * 16) = ldg_stream(src + i);
      }
    } else {
      const uint4* src = reinterpret_cast<const uint4*>(a.q) + (size_t(slab) * a.nb + b0) * 8;
      for (uint32_t i = lane; i < nbk * 8; i += 32) {  // item i = lb*8 + r: byte j = element j (low nibble) and j+16 (high)
        const uint4 w = ldg_stream(src + i);
        uint4 lo, hi;
        lo.x = w.x & 0x0f0f0f0fu; lo.y = w.y & 0x0f0f0f0fu; lo.z = w.z & 0x0f0f0f0fu; lo.w = w.w & 0x0f0f0f0fu;
        hi.x = (w.x >> 4) & 0x0f0f0f0fu; hi.y = (w.y >> 4) & 0x0f0f0f0fu;
        hi.z = (w.z >> 4) & 0x0f0f0f0fu; hi.w = (w.w >> 4) & 0x0f0f0f0fu;
        *reinterpret_cast<uint4*>(wq + i * 32) = lo;
        *reinterpret_cast<uint4*>(wq + i * 32 + 16) = hi;
      }
    }
    const uint16_t* dsrc = reinterpret_cast<const uint16_t*>(a.d) + (size_t(slab) * a.nb + b0) * 8;
    for (uint32_t i = lane; i < nbk * 8; i += 32) wd[i] = h2f(ldg_stream(dsrc + i));
  }
  // one block of one row folded into the chain accumulator (BodyQ4_0::compute / BodyQ8_0::compute)
  // m8 = -8 * (sum of the activation block's quants) for Q4_0 (sum (nib-8)*q = sum nib*q - 8*sum q), 0 for Q8_0
  __device__ static float fold(const uint4 w0, const uint4 w1, const float dw, const int4 xa, const int4 xb,
                               const float dx, const int m8, float acc) {
    int dp = m8;
    dp = __dp4a(int(w0.x), xa.x, dp);
    dp = __dp4a(int(w0.y), xa.y, dp);
    dp = __dp4a(int(w0.z), xa.z, dp);
    dp = __dp4a(int(w0.w), xa.w, dp);
    dp = __dp4a(int(w1.x), xb.x, dp);
    dp = __dp4a(int(w1.y), xb.y, dp);
    dp = __dp4a(int(w1.z), xb.z, dp);
    dp = __dp4a(int(w1.w), xb.w, dp);
    if (IS_Q8) return fmaf(float(dp) * dw, dx, acc);  // (int*dw)*dx, ops.cpp:820
    return fmaf(dw * dx, float(dp), acc);              // ops.cpp:380-395
  }
};

// One CTA = one K-chunk (blockIdx.z) x W slabs x 32 tokens (blockIdx.y): chunk partials are independent of
// each other, so the K dimension is a grid dimension too and short, wide matrices (ffn_down) still fill the
// GPU; toklane_reduce_kernel then adds the chunk partials of every (token, row) left to right.
template <bool IS_Q8, int W>
__global__ void __launch_bounds__(W * 32) gemm_toklane_kernel(const GemvBatch batch) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int mi = 0;
  while (mi + 1 < batch.n && blockIdx.x >= batch.cta_end[mi]) ++mi;
  const GemvArgs& a = batch.a[mi];
  const uint32_t cta = blockIdx.x - (mi ? batch.cta_end[mi - 1] : 0u);
  pdl_trigger();
  const uint32_t j = blockIdx.z, b0 = j * 16;
  if (b0 >= a.nb) return;  // matrices of one launch may differ in K (never for q/k/v, gate/up)
  const uint32_t nbk = min(16u, a.nb - b0);
  uint8_t* act_q = smem;                                                        // [32 tokens][TL_ACT_ROW]
  uint32_t* act_s = reinterpret_cast<uint32_t*>(smem + 32 * TL_ACT_ROW);        // [32 tokens][TL_SC_ROW]
  uint8_t* wq = smem + 32 * TL_ACT_ROW + 32 * TL_SC_ROW * 4 + size_t(warp) * (16 * 8 * 32 + 16 * 8 * 4);
  float* wd = reinterpret_cast<float*>(wq + 16 * 8 * 32);
  const uint32_t slab = cta * W + warp;
  const bool have_slab = slab < a.n_slabs;
  const uint32_t m0 = blockIdx.y * 32, n_here = min(32u, a.n_tok - m0);
  if (have_slab) TokLane<IS_Q8>::stage_weights(a, slab, b0, nbk, wq, wd, lane);  // independent of the predecessor
  pdl_wait();
  // activation tile of the chunk: token-major rows, every thread copies 16-byte pieces (coalesced per token)
  for (uint32_t p = threadIdx.x; p < n_here * 32; p += W * 32) {
    const uint32_t tk = p >> 5, piece = p & 31;
    if (piece < nbk * 2) {
      const uint4 v = *reinterpret_cast<const uint4*>(a.act + size_t(m0 + tk) * a.act_stride + size_t(b0) * 32 + piece * 16);
      *reinterpret_cast<uint4*>(act_q + tk * TL_ACT_ROW + piece * 16) = v;
    }
  }
  for (uint32_t p = threadIdx.x; p < n_here * 16; p += W * 32) {
    const uint32_t tk = p >> 4, lb = p & 15;
    if (lb < nbk)
      act_s[tk * TL_SC_ROW + lb] = reinterpret_cast<const uint32_t*>(a.act + size_t(m0 + tk) * a.act_stride + a.n_cols)[b0 + lb];
  }
  __syncthreads();
  if (!have_slab) return;
  float s[LLMI_SLAB][4];
#pragma unroll
  for (int r = 0; r < LLMI_SLAB; ++r)
#pragma unroll
    for (int k = 0; k < 4; ++k) s[r][k] = 0.0f;
  for (uint32_t lb4 = 0; lb4 < nbk; lb4 += 4) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t lb = lb4 + k;
      if (lb < nbk) {
        const int4 xa = *reinterpret_cast<const int4*>(act_q + lane * TL_ACT_ROW + lb * 32);
        const int4 xb = *reinterpret_cast<const int4*>(act_q + lane * TL_ACT_ROW + lb * 32 + 16);
        const uint32_t sw = act_s[lane * TL_SC_ROW + lb];
        const float dx = h2f(uint16_t(sw & 0xffffu));
        const int m8 = IS_Q8 ? 0 : -8 * int(int16_t(sw >> 16));
#pragma unroll
        for (int r = 0; r < LLMI_SLAB; ++r) {
          const uint4 w0 = *reinterpret_cast<const uint4*>(wq + (lb * 8 + r) * 32);
          const uint4 w1 = *reinterpret_cast<const uint4*>(wq + (lb * 8 + r) * 32 + 16);
          s[r][k] = TokLane<IS_Q8>::fold(w0, w1, wd[lb * 8 + r], xa, xb, dx, m8, s[r][k]);
        }
      }
    }
  }
  if (uint32_t(lane) < n_here) {
    float4* o = reinterpret_cast<float4*>(a.part + (size_t(j) * a.n_tok + m0 + lane) * (a.n_slabs * LLMI_SLAB) +
                                          size_t(slab) * LLMI_SLAB);
    o[0] = make_float4((s[0][0] + s[0][1]) + (s[0][2] + s[0][3]), (s[1][0] + s[1][1]) + (s[1][2] + s[1][3]),
                       (s[2][0] + s[2][1]) + (s[2][2] + s[2][3]), (s[3][0] + s[3][1]) + (s[3][2] + s[3][3]));
    o[1] = make_float4((s[4][0] + s[4][1]) + (s[4][2] + s[4][3]), (s[5][0] + s[5][1]) + (s[5][2] + s[5][3]),
                       (s[6][0] + s[6][1]) + (s[6][2] + s[6][3]), (s[7][0] + s[7][1]) + (s[7][2] + s[7][3]));
  }
}

// out[token][row] = p[0] + p[1] + ... + p[J-1] (left to right: the canonical order); blockIdx.y = matrix
__global__ void toklane_reduce_kernel(const GemvBatch batch) {
  pdl_trigger();
  pdl_wait();
  const GemvArgs& a = batch.a[blockIdx.y];
  const uint32_t rows_p = a.n_slabs * LLMI_SLAB, J = (a.nb + 15) / 16;
  const uint64_t total = uint64_t(a.n_tok) * a.n_local;
  for (uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; idx < total; idx += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t m = uint32_t(idx / a.n_local), row = uint32_t(idx % a.n_local);
    const float* p = a.part + size_t(m) * rows_p + row;
    float sum = p[0];
    for (uint32_t j = 1; j < J; ++j) sum += p[size_t(j) * a.n_tok * rows_p];
    a.out[size_t(m) * a.out_stride + row] = sum;
  }
}

#include "umma_prefill.cuh"
#include "gemv_ring.cuh"
#include "gemm_bf16.cuh"

// --------------------------------------------------------------- debug dump
// One thread per (local row, block): recomputes the integer block dot with the
// same device functions and planes as the GEMV.
__global__ void block_dots_kernel(const GemvArgs a, uint32_t type, int32_t* dots) {
  const uint64_t per_row = type == LLMI_Q6_K ? uint64_t(a.nb) * 2 : (type == LLMI_Q4_K ? uint64_t(a.nb) * 8 : a.nb);
  const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (idx >= uint64_t(a.n_local) * per_row) return;
  const uint32_t row = uint32_t(idx / per_row), u = uint32_t(idx % per_row);
  const uint32_t slab = row / LLMI_SLAB, r = row % LLMI_SLAB, nb = a.nb;
  const int4* xs = reinterpret_cast<const int4*>(a.act);
  if (type == LLMI_Q4_0) {
    const uint4 w = reinterpret_cast<const uint4*>(a.q)[((size_t)slab * nb + u) * 8 + r];
    const uint32_t m = reinterpret_cast<const uint32_t*>(a.act + a.n_cols)[u];
    dots[idx] = q4_0_block_dot(w, xs[2 * u], xs[2 * u + 1], int(int16_t(m >> 16)));
  } else if (type == LLMI_Q8_0) {
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + ((size_t)slab * nb + u) * 16 + r;
    dots[idx] = q8_0_block_dot(q[0], q[8], xs[2 * u], xs[2 * u + 1]);
  } else if (type == LLMI_Q4_K) {
    const uint32_t sb = u / 8, j = u % 8, c = j / 2;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (((size_t)slab * nb + sb) * 4 + c) * 16 + r;
    const int4* xp = xs + sb * 16 + c * 4;
    int lo, hi;
    q4_k_pair_dots(q[0], q[8], xp[0], xp[1], xp[2], xp[3], lo, hi);
    dots[idx] = (j & 1) ? hi : lo;
  } else if (type == LLMI_Q6_K) {
    const uint32_t sb = u / 2, n = u % 2;
    const int16_t* bs = reinterpret_cast<const int16_t*>(a.act + a.n_cols);
    int part = 0;
    for (int hh = 0; hh < 2; ++hh) {
      const int sub = 2 * n + hh;
      const uint4* q = reinterpret_cast<const uint4*>(a.q) + (((size_t)slab * nb + sb) * 12 + sub) * 8 + r;
      const uint2 sc = (reinterpret_cast<const uint2*>(a.x) + (((size_t)slab * nb + sb) * 8 + r) * 2)[n];
      const uint32_t g0 = sb * 16 + n * 8 + hh;
      part += q6_k_part(q[0], q[32], q[64], xs[g0], xs[g0 + 2], xs[g0 + 4], xs[g0 + 6], sbyte(sc, hh),
                        sbyte(sc, hh + 2), sbyte(sc, hh + 4), sbyte(sc, hh + 6), bs[g0], bs[g0 + 2], bs[g0 + 4],
                        bs[g0 + 6]);
    }
    dots[idx] = part;
  }
}

// ------------------------------------------------------------------ dispatch
constexpr int MAX_DYN_SMEM = 96 * 1024;  // fp32 activations of K=21504 need 86 KB

int g_sm_count = 148;

using Q4_0 = BodyQ4_0;
using Q8_0 = BodyQ8_0;
using Q5_0 = BodyQ5_0;
using Q4_K = BodyQ4_K;
using Q6_K = BodyQ6_K;
using F16 = BodyHalf<false>;
using BF16 = BodyHalf<true>;

// Benches can pin the CTA shape: g_warps in {0 (heuristic), 4, 8, 16},
// g_slabs_per_cta in {0 (heuristic), 1..}.  Results never depend on it.
int g_warps = 0, g_slabs_per_cta = 0;

// (W, S) for one matrix.  Heuristic from tools/gemv_sweep.py
// (profiles/r01_sweep_v4.jsonl): one slab per CTA and the fewest warps per CTA
// that still put ~24 warps on every SM (more, smaller CTAs beat fewer, larger
// ones); never more warps than the slab has chunks; the k-quants (bigger
// register footprint) stop at 8.  Huge-N / short-K matrices (logits) get 4
// slabs per CTA so that every warp streams several items and the activation
// staging is amortised.
template <class B>
void pick_shape(const GemvArgs& a, uint64_t slabs_in_launch, int& W, uint32_t& S) {
  const uint64_t n_items = uint64_t(a.n_slabs) * a.chunks;
  const bool kq = B::C == 2;
  W = 4;
  while (W < (kq ? 8 : 16) && slabs_in_launch * W < uint64_t(g_sm_count) * 24) W *= 2;
  if (a.chunks <= 4) W = 4;
  else if (a.chunks <= 8 && W > 8) W = 8;
  S = (a.chunks <= 10 && n_items >= uint64_t(g_sm_count) * 64 * 6) ? 4 : 1;
  if (g_warps) W = g_warps;
  if (g_slabs_per_cta) S = uint32_t(g_slabs_per_cta);
  while (S > 1 && a.act_bytes + size_t(S) * a.chunks * LLMI_SLAB * 4 > size_t(MAX_DYN_SMEM)) --S;
}


// ---- persistent bulk-copy-fed kernel (gemv_ring.cuh) --------------------------------------------------------
// g_ring_mode: 0 = heuristic (ring_wanted), 1 = never, 2 = wherever it fits.  g_ring_cps / g_ring_depth: CTAs per SM
// and ring slots per warp (0 = default).  Results never depend on any of them.
int g_ring_mode = 0, g_ring_cps = 0, g_ring_depth = 0, g_ring_w = 16;  // g_ring_w: warps per CTA, 8 or 16
constexpr uint32_t RING_MAX_CTAS = 148 * 4, RING_MAX_CHUNKS = 64, RING_MAX_PART_ITEMS = 768;
constexpr size_t RING_MAX_SMEM = 112 * 1024;

// Flagged scratch of the slabs split across CTAs: one buffer per stream (launches of one stream never overlap in
// it, gemv_ring.cuh), zero between launches.  Allocated on a stream's first launch — outside graph capture: every
// captured path is warmed up first.
std::mutex g_fix_mu;
std::map<cudaStream_t, uint2*> g_fix;
struct StreamScratch {
  void* p[SCR_COUNT] = {};
  size_t bytes[SCR_COUNT] = {};
};
std::mutex g_scr_mu;
std::map<cudaStream_t, StreamScratch> g_scr;
cudaError_t ring_fix_for(cudaStream_t s, uint2** out) {
  std::lock_guard<std::mutex> lk(g_fix_mu);
  auto it = g_fix.find(s);
  if (it != g_fix.end()) {
    *out = it->second;
    return cudaSuccess;
  }
  uint2* p = nullptr;
  const size_t bytes = size_t(RING_MAX_CTAS) * RING_MAX_CHUNKS * LLMI_SLAB * sizeof(uint2);
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) return e;
  if ((e = cudaMemset(p, 0, bytes)) != cudaSuccess) return e;
  g_fix[s] = p;
  *out = p;
  return cudaSuccess;
}

template <class B>
size_t ring_smem(int D, uint32_t act_bytes, uint32_t part_items) {
  return size_t(g_ring_w) * D * RingGeo<B>::SLOT + ((act_bytes + 127u) & ~127u) + size_t(part_items) * LLMI_SLAB * 4;
}

// Grid and ring depth for a launch of `total` items; false: the launch does not fit this kernel.
template <class B>
bool ring_plan(const GemvArgs* args, int n, uint32_t& ctas, int& D, uint32_t& total, size_t& smem) {
  uint64_t slabs = 0;
  for (int i = 0; i < n; ++i) {
    slabs += args[i].n_slabs;
    if (args[i].nb != args[0].nb || args[i].act != args[0].act) return false;  // one activation, one K
  }
  const uint64_t t = slabs * args[0].chunks;
  if (t == 0 || t > 0x7fffffffu || args[0].chunks > RING_MAX_CHUNKS) return false;
  total = uint32_t(t);
  const int cps = g_ring_cps ? g_ring_cps : 2;
  D = g_ring_depth ? g_ring_depth : 2;
  if (g_ring_w == 16 && D > 3) D = 3;  // instantiated: 8 warps x {2, 3, 4} slots, 16 warps x {2, 3}
  uint64_t c = std::min<uint64_t>(uint64_t(g_sm_count) * cps, RING_MAX_CTAS);
  c = std::min<uint64_t>(c, (t + g_ring_w - 1) / g_ring_w);  // no CTA with fewer items than warps
  ctas = uint32_t(std::max<uint64_t>(c, 1));
  const uint32_t per = uint32_t((t + ctas - 1) / ctas);
  if (per > RING_MAX_PART_ITEMS) return false;
  smem = ring_smem<B>(D, args[0].act_bytes, per + args[0].chunks);  // + the collected chunks of a split last slab
  return smem <= RING_MAX_SMEM;
}

// The heuristic of mode 0 (measured in the decode step, profiles/r02_ab_ring_v2.txt and r02_notes.md): the persistent
// kernel pays where its item body is the hand-scheduled one (Q4_0) and the launch streams enough bytes for the ring
// prefill and the even item split to outweigh its longer prologue and tail — gemma-3-27b: 4.63 -> 4.29 ms per token;
// it loses on gemma-3-1b's 0.15-9 MB launches (0.84 -> 0.93 ms) and, with the generic item body, on the k-quants.
template <class B>
bool ring_wanted(const GemvArgs* args, int n) {
  if (!std::is_same<B, BodyQ4_0>::value) return false;
  uint64_t bytes = 0;
  for (int i = 0; i < n; ++i) bytes += uint64_t(args[i].n_slabs) * args[i].nb * (RingFmt<B>::QB + RingFmt<B>::DB + RingFmt<B>::XB);
  return bytes >= (10ull << 20);
}

template <class B>
cudaError_t launch_ring(const GemvArgs* args, int n, cudaStream_t s, const GemvLL* ll, uint32_t ctas, int D, uint32_t total,
                        size_t smem, const RingNorm* norm = nullptr) {
  RingBatch b;
  if (norm) b.norm = *norm;
  b.n = n;
  b.total = total;
  b.part_items = uint32_t((uint64_t(total) + ctas - 1) / ctas);
  static const bool late = getenv("LLMI_RING_LATE") && getenv("LLMI_RING_LATE")[0] == '1';
  b.late_fill = late ? 1u : 0u;
  uint32_t slabs = 0;
  for (int i = 0; i < GEMV_MAX_BATCH; ++i) {
    b.a[i] = args[i < n ? i : 0];
    if (i < n) slabs += args[i].n_slabs;
    b.slab_end[i] = slabs;
  }
  if (ll) {
    b.peers = ll->peers;
    b.tag = ll->tag;
    for (int i = 0; i < n; ++i)
      if (b.a[i].ll_off == LL_NONE) return cudaErrorInvalidValue;
  }
  cudaError_t e = ring_fix_for(s, &b.fix);
  if (e != cudaSuccess) return e;
  const dim3 grid(ctas), block(g_ring_w * 32);
  if (norm) {  // the RMSNorm stage as the kernel's prologue: 16 warps, single GPU (gemv_ring.cuh ring_norm_prologue)
    if (g_ring_w != 16 || ll) return cudaErrorNotSupported;
    if (D == 2) return llmi_launch(gemv_ring_kernel<B, 16, 2, false, true>, grid, block, smem, s, b);
    if (D == 3) return llmi_launch(gemv_ring_kernel<B, 16, 3, false, true>, grid, block, smem, s, b);
    return cudaErrorNotSupported;
  }
#define LLMI_RING_CASE(WW, DD)                                                                               \
  if (g_ring_w == WW && D == DD)                                                                             \
    return ll ? llmi_launch(gemv_ring_kernel<B, WW, DD, true>, grid, block, smem, s, b)                      \
              : llmi_launch(gemv_ring_kernel<B, WW, DD, false>, grid, block, smem, s, b)
  LLMI_RING_CASE(8, 2);
  LLMI_RING_CASE(8, 3);
  LLMI_RING_CASE(8, 4);
  LLMI_RING_CASE(16, 2);
  LLMI_RING_CASE(16, 3);
#undef LLMI_RING_CASE
  return cudaErrorInvalidValue;
}

template <class B>
cudaError_t ring_optin() {
  cudaError_t e;
#define LLMI_RING_OPT(WW, DD, P)                                                                                \
  if ((e = cudaFuncSetAttribute(gemv_ring_kernel<B, WW, DD, P>, cudaFuncAttributeMaxDynamicSharedMemorySize,   \
                                int(RING_MAX_SMEM))) != cudaSuccess)                                            \
    return e
  LLMI_RING_OPT(8, 2, false); LLMI_RING_OPT(8, 3, false); LLMI_RING_OPT(8, 4, false);
  LLMI_RING_OPT(8, 2, true); LLMI_RING_OPT(8, 3, true); LLMI_RING_OPT(8, 4, true);
  LLMI_RING_OPT(16, 2, false); LLMI_RING_OPT(16, 3, false);
  LLMI_RING_OPT(16, 2, true); LLMI_RING_OPT(16, 3, true);
#undef LLMI_RING_OPT
  if (std::is_same<B, BodyQ4_0>::value || std::is_same<B, BodyQ8_0>::value) {  // the norm prologue emits Q8_0 activations
    if ((e = cudaFuncSetAttribute(gemv_ring_kernel<B, 16, 2, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(RING_MAX_SMEM))) != cudaSuccess) return e;
    if ((e = cudaFuncSetAttribute(gemv_ring_kernel<B, 16, 3, false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  int(RING_MAX_SMEM))) != cudaSuccess) return e;
  }
  return cudaSuccess;
}

template <class B>
cudaError_t launch_batch(const GemvArgs* args, int n, cudaStream_t s, const GemvLL* ll) {
  if (g_ring_mode != 1) {
    uint32_t rc = 0, rt = 0;
    int rd = 0;
    size_t rs = 0;
    if ((g_ring_mode == 2 || ring_wanted<B>(args, n)) && ring_plan<B>(args, n, rc, rd, rt, rs))
      return launch_ring<B>(args, n, s, ll, rc, rd, rt, rs);
  }
  GemvBatch b;
  b.n = n;
  if (ll) {
    b.peers = ll->peers;
    b.tag = ll->tag;
  }
  uint64_t slabs = 0;
  for (int i = 0; i < n; ++i) slabs += args[i].n_slabs;
  int W = 4;
  size_t smem = 0;
  uint32_t ctas = 0;
  for (int i = 0; i < n; ++i) {
    int wi;
    uint32_t si;
    pick_shape<B>(args[i], slabs, wi, si);
    if (wi > W) W = wi;
    b.a[i] = args[i];
    b.S[i] = si;
    ctas += (args[i].n_slabs + si - 1) / si;
    b.cta_end[i] = ctas;
    const size_t need = args[i].act_bytes + size_t(si) * args[i].chunks * LLMI_SLAB * 4;
    if (need > smem) smem = need;
  }
  for (int i = n; i < GEMV_MAX_BATCH; ++i) {
    b.a[i] = args[0];
    b.S[i] = 1;
    b.cta_end[i] = ctas;
  }
  if (ctas == 0) return cudaSuccess;
  if (ll) {  // every matrix of a sharded launch pushes (model.cu passes an offset per matrix)
    for (int i = 0; i < n; ++i)
      if (b.a[i].ll_off == LL_NONE) return cudaErrorInvalidValue;
    switch (W) {
      case 4: return llmi_launch(gemv_slab_kernel<B, 4, true>, dim3(ctas), dim3(128), smem, s, b);
      case 8: return llmi_launch(gemv_slab_kernel<B, 8, true>, dim3(ctas), dim3(256), smem, s, b);
      default: return llmi_launch(gemv_slab_kernel<B, 16, true>, dim3(ctas), dim3(512), smem, s, b);
    }
  }
  switch (W) {
    case 4: return llmi_launch(gemv_slab_kernel<B, 4, false>, dim3(ctas), dim3(128), smem, s, b);
    case 8: return llmi_launch(gemv_slab_kernel<B, 8, false>, dim3(ctas), dim3(256), smem, s, b);
    default: return llmi_launch(gemv_slab_kernel<B, 16, false>, dim3(ctas), dim3(512), smem, s, b);
  }
}

constexpr size_t TOK_TILE_BYTES = 64 * 1024;  // activations of one token tile in shared memory
constexpr int TOK_MAX_TILE = 8;

constexpr size_t tl_smem(int W) { return 32 * TL_ACT_ROW + 32 * TL_SC_ROW * 4 + size_t(W) * (16 * 8 * 32 + 16 * 8 * 4); }


template <bool IS_Q8>
cudaError_t launch_toklane(const GemvArgs* args, int n, cudaStream_t s) {
  constexpr int W = 8;
  GemvBatch b;
  b.n = n;
  uint32_t ctas = 0, jmax = 0;
  size_t need = 0;
  for (int i = 0; i < n; ++i) {
    const uint32_t J = (args[i].nb + 15) / 16;
    if (J > jmax) jmax = J;
    need += size_t(J) * args[i].n_tok * args[i].n_slabs * LLMI_SLAB;
  }
  float* g_part = nullptr;  // chunk partials of this launch (per-stream scratch)
  if (cudaError_t e = llmi_stream_scratch(s, SCR_PART, need * sizeof(float), (void**)&g_part)) return e;
  size_t off = 0;
  uint64_t outs = 0;
  for (int i = 0; i < n; ++i) {
    b.a[i] = args[i];
    b.a[i].part = g_part + off;
    off += size_t((args[i].nb + 15) / 16) * args[i].n_tok * args[i].n_slabs * LLMI_SLAB;
    b.S[i] = W;
    ctas += (args[i].n_slabs + W - 1) / W;
    b.cta_end[i] = ctas;
    outs = std::max<uint64_t>(outs, uint64_t(args[i].n_tok) * args[i].n_local);
  }
  for (int i = n; i < GEMV_MAX_BATCH; ++i) {
    b.a[i] = b.a[0];
    b.S[i] = W;
    b.cta_end[i] = ctas;
  }
  if (ctas == 0) return cudaSuccess;
  cudaError_t e = llmi_launch(gemm_toklane_kernel<IS_Q8, W>, dim3(ctas, (args[0].n_tok + 31) / 32, jmax), dim3(W * 32),
                              tl_smem(W), s, b);
  if (e != cudaSuccess) return e;
  const unsigned rb = unsigned(std::min<uint64_t>((outs + 255) / 256, uint64_t(g_sm_count) * 8));
  return llmi_launch(toklane_reduce_kernel, dim3(rb, n), dim3(256), 0, s, b);
}

// grow-only scratch of one tensor-core prefill launch: activations in UMMA operand order + fp32 scales
bool g_umma = true;  // LLMI_NO_UMMA=1: token batches stay on the dp4a token-per-lane kernel (A/B)
// Below this many tokens the dp4a token-per-lane kernel is at least as fast (measured, profiles/r01_notes.md: a CTA
// of the tensor-core kernel has ~3 us of fixed cost); LLMI_UMMA_MIN_TOKENS overrides (>= 32: one token tile).
uint32_t g_umma_min_tokens = 128;
uint32_t g_fast_tnf = 0;  // LLMI_FAST_TNF = 128 | 256 pins the token tile of the bf16 GEMM (0: by grid fill)

// The quant plane of a matrix as a 2-D tensor for TMA: [slab][the slab's K run] in 8-byte elements; box = 16 slabs x
// one stage (8 blocks).  The encoder lives in the driver library: fetched once through the runtime, no -lcuda.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static cudaError_t make_plane_map(CUtensorMap* tm, const GemvArgs& a, bool is_q8) {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr);
    if (e != cudaSuccess) return e;
    if (!p || qr != cudaDriverEntryPointSuccess) return cudaErrorNotSupported;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  const cuuint64_t per_blk = is_q8 ? 32 : 16;  // 8-byte elements per (slab, block): 8 rows x 32 / 16 bytes
  const cuuint64_t dims[2] = {cuuint64_t(a.nb) * per_blk, cuuint64_t(a.n_slabs)};
  const cuuint64_t strides[1] = {cuuint64_t(a.nb) * per_blk * 8};
  const cuuint32_t box[2] = {cuuint32_t(umma::SB * per_blk), cuuint32_t(umma::TM / LLMI_SLAB)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, const_cast<uint8_t*>(a.q), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

// Token batches of >= 32 tokens, Q4_0 / Q8_0 weights: exact int8 tensor-core path (umma_prefill.cuh).
template <bool IS_Q8>
cudaError_t launch_umma(const GemvArgs* args, int n, cudaStream_t s) {
  GemvBatch b;
  b.n = n;
  const uint32_t n_tok = args[0].n_tok, nb0 = args[0].nb;
  for (int i = 1; i < n; ++i)
    if (args[i].nb != nb0) return cudaErrorInvalidValue;  // one activation -> one K
  const uint32_t tiles_n = (n_tok + umma::TN - 1) / umma::TN, J = (nb0 + 15) / 16;
  size_t need = 0;
  for (int i = 0; i < n; ++i) need += size_t(J) * n_tok * args[i].n_slabs * LLMI_SLAB;
  const size_t bq_items = size_t(tiles_n) * nb0 * 64, bd_floats = size_t(tiles_n) * nb0 * umma::TN;
  float* g_part = nullptr;  // per-stream scratch: chunk partials, the batch's quants and scales in operand order
  uint4* g_bq = nullptr;
  float* g_bd = nullptr;
  cudaError_t se;
  if ((se = llmi_stream_scratch(s, SCR_PART, need * sizeof(float), (void**)&g_part)) != cudaSuccess) return se;
  if ((se = llmi_stream_scratch(s, SCR_BQ, bq_items * sizeof(uint4), (void**)&g_bq)) != cudaSuccess) return se;
  if ((se = llmi_stream_scratch(s, SCR_BD, bd_floats * sizeof(float), (void**)&g_bd)) != cudaSuccess) return se;
  size_t off = 0;
  uint32_t ctas = 0;
  uint64_t outs = 0;
  for (int i = 0; i < n; ++i) {
    b.a[i] = args[i];
    b.a[i].part = g_part + off;
    off += size_t(J) * n_tok * args[i].n_slabs * LLMI_SLAB;
    b.S[i] = umma::TM / LLMI_SLAB;
    ctas += (args[i].n_slabs + umma::TM / LLMI_SLAB - 1) / (umma::TM / LLMI_SLAB);
    b.cta_end[i] = ctas;
    outs = std::max<uint64_t>(outs, uint64_t(n_tok) * args[i].n_local);
  }
  for (int i = n; i < GEMV_MAX_BATCH; ++i) {
    b.a[i] = b.a[0];
    b.S[i] = umma::TM / LLMI_SLAB;
    b.cta_end[i] = ctas;
  }
  if (ctas == 0) return cudaSuccess;
  const unsigned pb = unsigned(std::min<uint64_t>((bq_items + 255) / 256, uint64_t(g_sm_count) * 8));
  cudaError_t e = llmi_launch(umma_pack_act_kernel, dim3(pb), dim3(256), 0, s, args[0].act, args[0].act_stride,
                              args[0].n_cols, nb0, n_tok, g_bq, g_bd);
  if (e != cudaSuccess) return e;
  // K-chunk groups across grid.z until the grid holds ~2 CTAs per SM (one CTA per SM is resident)
  uint32_t gz = 1;
  while (gz < J && uint64_t(ctas) * tiles_n * gz < uint64_t(g_sm_count) * 2) ++gz;
  const uint32_t nj = (J + gz - 1) / gz;
  gz = (J + nj - 1) / nj;
  alignas(64) CUtensorMap tms[GEMV_MAX_BATCH];
  for (int i = 0; i < GEMV_MAX_BATCH; ++i)
    if ((e = make_plane_map(&tms[i], b.a[i], IS_Q8)) != cudaSuccess) return e;
  if (gz == 1)  // every CTA walks all chunks of its tile: running totals in registers, rows stored once, no reduce launch
    return llmi_launch(gemm_umma_kernel<IS_Q8, true>, dim3(ctas, tiles_n, 1), dim3(umma::WARPS * 32),
                       umma::Cfg<IS_Q8>::SMEM_BYTES, s, b, (const uint4*)g_bq, (const float*)g_bd, nj, tms[0], tms[1], tms[2]);
  e = llmi_launch(gemm_umma_kernel<IS_Q8, false>, dim3(ctas, tiles_n, gz), dim3(umma::WARPS * 32),
                  umma::Cfg<IS_Q8>::SMEM_BYTES, s, b, (const uint4*)g_bq, (const float*)g_bd, nj, tms[0], tms[1], tms[2]);
  if (e != cudaSuccess) return e;
  const unsigned rb = unsigned(std::min<uint64_t>((outs + 255) / 256, uint64_t(g_sm_count) * 8));
  return llmi_launch(toklane_reduce_kernel, dim3(rb, n), dim3(256), 0, s, b);
}

// ---- throughput prefill (gemm_bf16.cuh): opt-in, not the parity path ---------------------------------------------
bool g_prefill_fast = false;  // LLMI_PREFILL=fast / llmi_set_prefill_mode(1)

template <class B>
constexpr uint32_t body_type() {
  return std::is_same<B, BodyQ4_0>::value   ? LLMI_Q4_0
         : std::is_same<B, BodyQ8_0>::value ? LLMI_Q8_0
         : std::is_same<B, BodyQ5_0>::value ? LLMI_Q5_0
         : std::is_same<B, BodyQ4_K>::value ? LLMI_Q4_K
         : std::is_same<B, BodyQ6_K>::value ? LLMI_Q6_K
         : std::is_same<B, BodyHalf<false>>::value ? LLMI_F16 : LLMI_BF16;
}

// Every matrix of the batch: dequantize -> (activations packed once) -> tcgen05 bf16 GEMM.
template <class B>
cudaError_t launch_fast(const GemvArgs* args, int n, cudaStream_t s, const float* geglu_gate = nullptr,
                        const float* geglu_up = nullptr, GemmPush* push = nullptr, const void* hid16 = nullptr) {
  const uint32_t type = body_type<B>(), K = args[0].n_cols, n_tok = args[0].n_tok;
  const uint32_t nkb = K / fastmm::KB;
  // token tile: 256 when the grid fills the GPU; 128 when the launch is short of CTAs (row shards of a sharded model:
  // 5376 / 8 rows x 2048 tokens is 48 CTAs of 256 tokens) and halving the tile shortens the last wave.  Estimated
  // time = waves x tile width; a 128-token CTA takes ~0.66 of a 256-token CTA's time, not half (measured: all-128 makes
  // the 27b prompt's GEMMs 25 % slower where the wave count says 5 % faster, profiles/r02_fast_tnf_ab.txt), so the
  // narrow tile must win by 30 %.  The result does not depend on the tile (bitwise, same file).
  uint32_t tnf = n_tok > 128 ? 256u : 128u;
  if (tnf == 256 && g_fast_tnf != 256) {
    uint64_t t256 = 0, t128 = 0;
    for (int i = 0; i < n; ++i) {
      const uint64_t tiles = (uint64_t(args[i].n_slabs) * LLMI_SLAB + fastmm::TM - 1) / fastmm::TM;
      const uint64_t c256 = tiles * ((n_tok + 255) / 256), c128 = tiles * ((n_tok + 127) / 128);
      t256 += (c256 + g_sm_count - 1) / g_sm_count * 256;
      t128 += (c128 + g_sm_count - 1) / g_sm_count * 128;
    }
    if (g_fast_tnf == 128 || t128 * 13 < t256 * 10) tnf = 128;
  }
  const uint32_t n_tt = (n_tok + tnf - 1) / tnf;
  size_t w_need = 0;
  for (int i = 0; i < n; ++i) {
    const size_t tiles = (size_t(args[i].n_slabs) * LLMI_SLAB + fastmm::TM - 1) / fastmm::TM;
    w_need = std::max(w_need, tiles * nkb * fastmm::A_BYTES);
  }
  const size_t x_need = size_t(n_tt) * nkb * tnf * fastmm::KB * 2;
  uint8_t *g_fast_w = nullptr, *g_fast_x = nullptr;  // per-stream scratch: dequantized weights / packed activations of one launch
  {
    cudaError_t e;
    if ((e = llmi_stream_scratch(s, SCR_FAST_W, w_need, (void**)&g_fast_w)) != cudaSuccess) return e;
    if ((e = llmi_stream_scratch(s, SCR_FAST_X, x_need, (void**)&g_fast_x)) != cudaSuccess) return e;
  }
  const int kind = llmi_act_kind_for(type);
  {
    const uint64_t items = uint64_t(n_tt) * nkb * tnf * 8;
    const unsigned blocks = unsigned(std::min<uint64_t>((items + 255) / 256, uint64_t(g_sm_count) * 16));
    cudaError_t e;
    if (hid16)  // the hidden batch as bf16 (row-sharded model: combined and rounded by the columns' owners)
      e = llmi_launch(fast_pack_act_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const uint8_t*>(hid16), K * 2u,
                      int(ACT_BF16_RAW), K, n_tok, tnf, n_tt, nkb, reinterpret_cast<uint4*>(g_fast_x));
    else if (geglu_gate && !geglu_up)  // `gate` already holds the hidden batch (row-sharded model: combined by the columns' owners)
      e = llmi_launch(fast_pack_act_kernel, dim3(blocks), dim3(256), 0, s, reinterpret_cast<const uint8_t*>(geglu_gate), K * 4u,
                      int(ACT_F32), K, n_tok, tnf, n_tt, nkb, reinterpret_cast<uint4*>(g_fast_x));
    else if (geglu_gate)  // the activation IS gelu(gate) * up of the two fp32 batches: no quantized detour
      e = llmi_launch(fast_geglu_pack_kernel, dim3(blocks), dim3(256), 0, s, geglu_gate, geglu_up, K, n_tok, tnf, n_tt, nkb,
                      reinterpret_cast<uint4*>(g_fast_x));
    else
      e = llmi_launch(fast_pack_act_kernel, dim3(blocks), dim3(256), 0, s, args[0].act, args[0].act_stride, kind, K, n_tok, tnf,
                      n_tt, nkb, reinterpret_cast<uint4*>(g_fast_x));
    if (e != cudaSuccess) return e;
  }
  for (int i = 0; i < n; ++i) {
    const uint32_t tiles = uint32_t((size_t(args[i].n_slabs) * LLMI_SLAB + fastmm::TM - 1) / fastmm::TM);
    if (tiles == 0) continue;
    const uint64_t items = uint64_t(tiles) * nkb * 1024;
    const unsigned blocks = unsigned(std::min<uint64_t>((items + 255) / 256, uint64_t(g_sm_count) * 16));
    cudaError_t e = llmi_launch(fast_dequant_kernel, dim3(blocks), dim3(256), 0, s, args[i], type, tiles, nkb,
                                reinterpret_cast<uint4*>(g_fast_w));
    if (e != cudaSuccess) return e;
    if (tnf == 256)
      e = llmi_launch(gemm_bf16_kernel<256>, dim3(n_tt, tiles), dim3(192), fastmm::Cfg<256>::SMEM, s, (const uint8_t*)g_fast_w,
                      (const uint8_t*)g_fast_x, args[i].out, args[i].out_stride, args[i].n_local, n_tok, nkb,
                      llmi_peer_out(push, args[i].out));
    else
      e = llmi_launch(gemm_bf16_kernel<128>, dim3(n_tt, tiles), dim3(192), fastmm::Cfg<128>::SMEM, s, (const uint8_t*)g_fast_w,
                      (const uint8_t*)g_fast_x, args[i].out, args[i].out_stride, args[i].n_local, n_tok, nkb,
                      llmi_peer_out(push, args[i].out));
    if (e != cudaSuccess) return e;
  }
  if (push) push->done = true;
  return cudaSuccess;
}

template <class B>
cudaError_t launch_tokens(const GemvArgs* args, int n, cudaStream_t s, GemmPush* push = nullptr) {
  if (push) push->done = false;  // only the tensor-core GEMM stores into the peers itself
  if (g_prefill_fast && args[0].n_tok >= 64 && args[0].n_cols % fastmm::KB == 0) {
    bool same = true;  // one activation, one K
    for (int i = 1; i < n; ++i) same = same && args[i].act == args[0].act && args[i].n_cols == args[0].n_cols;
    if (same) return launch_fast<B>(args, n, s, nullptr, nullptr, push);
  }
  if (g_umma && args[0].n_tok >= g_umma_min_tokens && args[0].nb >= uint32_t(umma::SB)) {  // K >= one stage (256)
    if (std::is_same<B, BodyQ4_0>::value) return launch_umma<false>(args, n, s);
    if (std::is_same<B, BodyQ8_0>::value) return launch_umma<true>(args, n, s);
  }
  if (args[0].n_tok >= 16) {
    if (std::is_same<B, BodyQ4_0>::value) return launch_toklane<false>(args, n, s);
    if (std::is_same<B, BodyQ8_0>::value) return launch_toklane<true>(args, n, s);
  }
  GemvBatch b;
  b.n = n;
  uint64_t slabs = 0;
  for (int i = 0; i < n; ++i) slabs += args[i].n_slabs;
  int W = 4;
  uint32_t ctas = 0, MT = TOK_MAX_TILE;
  for (int i = 0; i < n; ++i) {
    const uint32_t fit = uint32_t(TOK_TILE_BYTES / args[i].act_stride);
    if (fit < MT) MT = fit;
  }
  if (MT == 0) return cudaErrorInvalidValue;
  if (MT > args[0].n_tok) MT = args[0].n_tok;
  size_t smem = 0;
  for (int i = 0; i < n; ++i) {
    int wi;
    uint32_t si;
    pick_shape<B>(args[i], slabs, wi, si);
    if (wi > W) W = wi;
    b.a[i] = args[i];
    b.S[i] = si;
    ctas += (args[i].n_slabs + si - 1) / si;
    b.cta_end[i] = ctas;
    const size_t need = size_t(MT) * (args[i].act_stride + size_t(si) * args[i].chunks * LLMI_SLAB * 4);
    if (need > smem) smem = need;
  }
  for (int i = n; i < GEMV_MAX_BATCH; ++i) {
    b.a[i] = args[0];
    b.S[i] = 1;
    b.cta_end[i] = ctas;
  }
  if (ctas == 0) return cudaSuccess;
  if (smem > size_t(MAX_DYN_SMEM)) return cudaErrorInvalidValue;
  // token tiles across grid.y until the grid holds ~8 CTAs per SM (the weights of a tile's CTAs come from L2)
  const uint32_t tiles = (args[0].n_tok + MT - 1) / MT;
  uint32_t gy = 1;
  while (gy < tiles && uint64_t(ctas) * gy < uint64_t(g_sm_count) * 8) gy *= 2;
  if (gy > tiles) gy = tiles;
  switch (W) {
    case 4: return llmi_launch(gemv_slab_tok_kernel<B, 4>, dim3(ctas, gy), dim3(128), smem, s, b, MT);
    case 8: return llmi_launch(gemv_slab_tok_kernel<B, 8>, dim3(ctas, gy), dim3(256), smem, s, b, MT);
    default: return llmi_launch(gemv_slab_tok_kernel<B, 16>, dim3(ctas, gy), dim3(512), smem, s, b, MT);
  }
}

template <class B>
cudaError_t optin() {
  cudaError_t e;
  if ((e = cudaFuncSetAttribute(gemv_slab_kernel<B, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_kernel<B, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_kernel<B, 16, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_kernel<B, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_kernel<B, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_kernel<B, 16, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_tok_kernel<B, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_tok_kernel<B, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_slab_tok_kernel<B, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_geglu_kernel<B, 4, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_geglu_kernel<B, 8, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  if ((e = cudaFuncSetAttribute(gemv_geglu_kernel<B, 4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                MAX_DYN_SMEM)) != cudaSuccess) return e;
  return cudaFuncSetAttribute(gemv_geglu_kernel<B, 8, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_DYN_SMEM);
}

uint32_t units_of(const llmi_weight_s& w) {
  return (w.type == LLMI_Q4_K || w.type == LLMI_Q6_K) ? uint32_t(w.nb) : uint32_t((w.nb + 3) / 4);
}
uint32_t chunk_units_of(uint32_t type) { return (type == LLMI_Q4_K || type == LLMI_Q6_K) ? 2u : 4u; }

}  // namespace

int llmi_act_kind_for(uint32_t t) {
  switch (t) {
    case LLMI_Q4_0:
    case LLMI_Q8_0: return ACT_Q8_0;
    case LLMI_Q4_K:
    case LLMI_Q6_K: return ACT_Q8_K;
    case LLMI_F16: return ACT_F16;
    case LLMI_Q5_0:
    case LLMI_BF16: return ACT_F32;
    default: return ACT_NONE;
  }
}

// K-chunks (work items) per slab of a weight; also sizes the split-slab scratch.
uint32_t llmi_gemv_chunks(const llmi_weight_s& w) {
  const uint32_t u = units_of(w), c = chunk_units_of(w.type);
  return (u + c - 1) / c;
}

void llmi_gemv_set_ring(int mode, int ctas_per_sm, int depth, int warps) {
  g_ring_w = warps == 8 ? 8 : 16;
  g_ring_mode = mode;
  g_ring_cps = ctas_per_sm;
  g_ring_depth = depth;
}

void llmi_gemv_set_prefill_fast(int on) { g_prefill_fast = on != 0; }
int llmi_gemv_prefill_fast() { return g_prefill_fast ? 1 : 0; }

void llmi_gemv_set_shape(int warps, int slabs_per_cta) {
  g_warps = warps;
  g_slabs_per_cta = slabs_per_cta;
}

cudaError_t llmi_stream_scratch(cudaStream_t s, int slot, size_t bytes, void** out) {
  if (slot < 0 || slot >= SCR_COUNT || !out) return cudaErrorInvalidValue;
  std::lock_guard<std::mutex> lk(g_scr_mu);
  StreamScratch& sc = g_scr[s];
  if (bytes > sc.bytes[slot] || !sc.p[slot]) {
    cudaError_t e = cudaStreamSynchronize(s);  // launches in flight may still use the old buffer
    if (e != cudaSuccess) return e;
    if (sc.p[slot]) cudaFree(sc.p[slot]);
    sc.p[slot] = nullptr;
    sc.bytes[slot] = 0;
    const size_t want = bytes ? bytes : 16;
    if ((e = cudaMalloc(&sc.p[slot], want)) != cudaSuccess) return e;
    sc.bytes[slot] = want;
  }
  *out = sc.p[slot];
  return cudaSuccess;
}

void llmi_stream_scratch_release(cudaStream_t s) {  // the stream is about to be destroyed (llmi_model_free)
  std::lock_guard<std::mutex> lk(g_scr_mu);
  auto it = g_scr.find(s);
  if (it == g_scr.end()) return;
  for (void* p : it->second.p)
    if (p) cudaFree(p);
  g_scr.erase(it);
}

// llmi_shutdown: the grow-only scratch of the token-batched launches (chunk partials, packed activations).
void llmi_gemv_shutdown() {
  {
    std::lock_guard<std::mutex> lk(g_fix_mu);
    for (auto& kv : g_fix) cudaFree(kv.second);
    g_fix.clear();
  }
  std::lock_guard<std::mutex> lk(g_scr_mu);
  for (auto& kv : g_scr)
    for (void* p : kv.second.p)
      if (p) cudaFree(p);
  g_scr.clear();
}

// Bench / test knobs of the token-batched path, re-read by every llmi_model_load.
void llmi_gemv_read_env() {
  const char* e = getenv("LLMI_NO_UMMA");
  g_umma = !(e && e[0] == '1');
  e = getenv("LLMI_UMMA_MIN_TOKENS");
  g_umma_min_tokens = e ? uint32_t(std::max(32, atoi(e))) : 128u;
  e = getenv("LLMI_FAST_TNF");
  g_fast_tnf = e ? uint32_t(atoi(e)) : 0u;
  if (g_fast_tnf != 128 && g_fast_tnf != 256) g_fast_tnf = 0;
  e = getenv("LLMI_PREFILL");  // "fast" / anything else = exact; unset: whatever llmi_set_prefill_mode last said
  if (e) g_prefill_fast = std::string(e) == "fast";
}

cudaError_t llmi_gemv_init() {
  cudaError_t e0;
  llmi_gemv_read_env();
  if ((e0 = cudaFuncSetAttribute(gemm_umma_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(umma::Cfg<false>::SMEM_BYTES))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_umma_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(umma::Cfg<true>::SMEM_BYTES))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_umma_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(umma::Cfg<false>::SMEM_BYTES))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_umma_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(umma::Cfg<true>::SMEM_BYTES))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_bf16_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(fastmm::Cfg<128>::SMEM))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_bf16_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(fastmm::Cfg<256>::SMEM))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_toklane_kernel<false, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(tl_smem(8)))) != cudaSuccess) return e0;
  if ((e0 = cudaFuncSetAttribute(gemm_toklane_kernel<true, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 int(tl_smem(8)))) != cudaSuccess) return e0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  if ((e = ring_optin<Q4_0>()) != cudaSuccess) return e;
  if ((e = ring_optin<Q8_0>()) != cudaSuccess) return e;
  if ((e = ring_optin<Q5_0>()) != cudaSuccess) return e;
  if ((e = ring_optin<Q4_K>()) != cudaSuccess) return e;
  if ((e = ring_optin<Q6_K>()) != cudaSuccess) return e;
  if ((e = ring_optin<F16>()) != cudaSuccess) return e;
  if ((e = ring_optin<BF16>()) != cudaSuccess) return e;
  if ((e = optin<Q4_0>()) != cudaSuccess) return e;
  if ((e = optin<Q8_0>()) != cudaSuccess) return e;
  if ((e = optin<Q5_0>()) != cudaSuccess) return e;
  if ((e = optin<Q4_K>()) != cudaSuccess) return e;
  if ((e = optin<Q6_K>()) != cudaSuccess) return e;
  if ((e = optin<F16>()) != cudaSuccess) return e;
  return optin<BF16>();
}

static GemvArgs make_args(const llmi_weight_s& w, const llmi_act_s& a, float* out) {
  GemvArgs g;
  g.q = w.p_q;
  g.d = w.p_d;
  g.x = w.p_x;
  g.act = a.buf;
  g.act_bytes = uint32_t(act_bytes(a.kind, a.n));
  g.out = out ? out + w.row_begin : nullptr;
  g.n_local = uint32_t(w.n_local);
  g.n_slabs = uint32_t(w.n_slabs);
  g.nb = uint32_t(w.nb);
  g.n_cols = uint32_t(w.n_cols);
  g.units = units_of(w);
  g.chunks = llmi_gemv_chunks(w);
  g.argmax_key = nullptr;
  g.softcap = 0.0f;
  g.row0 = uint32_t(w.row_begin);
  g.n_tok = 1;
  g.act_stride = g.act_bytes;
  g.out_stride = 0;
  g.part = nullptr;
  g.ll_off = LL_NONE;
  return g;
}

// The static fields of a matrix for the persistent decode kernel (mega.cu): planes, geometry, first row.
GemvArgs llmi_gemv_args(const llmi_weight_s& w) {
  llmi_act_s none;
  return make_args(w, none, nullptr);
}

// One launch for up to GEMV_MAX_BATCH matrices of the SAME format consuming the
// same prepared activation (q/k/v, gate/up): the grid is the union of their CTAs.
// key / softcap: the optional fused epilogue of the logits mat-vec (llmi_launch_gemv_argmax).
static cudaError_t launch_gemv_batch_impl(const llmi_weight_s* const* ws, float* const* outs, int n, const llmi_act_s& a,
                                          cudaStream_t s, const GemvLL* ll, unsigned long long* g_argmax_key,
                                          float g_argmax_softcap) {
  if (n < 1 || n > GEMV_MAX_BATCH) return cudaErrorInvalidValue;
  GemvArgs args[GEMV_MAX_BATCH];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (ws[i]->type != ws[0]->type) return cudaErrorInvalidValue;
    if (ws[i]->n_slabs == 0) continue;
    args[m] = make_args(*ws[i], a, outs[i]);
    args[m].argmax_key = g_argmax_key;
    args[m].softcap = g_argmax_softcap;
    if (ll) args[m].ll_off = ll->off[i];
    if (args[m].act_bytes > (uint32_t)MAX_DYN_SMEM) return cudaErrorInvalidValue;
    ++m;
  }
  if (m == 0) return cudaSuccess;
  switch (ws[0]->type) {
    case LLMI_Q4_0: return launch_batch<Q4_0>(args, m, s, ll);
    case LLMI_Q8_0: return launch_batch<Q8_0>(args, m, s, ll);
    case LLMI_Q5_0: return launch_batch<Q5_0>(args, m, s, ll);
    case LLMI_Q4_K: return launch_batch<Q4_K>(args, m, s, ll);
    case LLMI_Q6_K: return launch_batch<Q6_K>(args, m, s, ll);
    case LLMI_F16: return launch_batch<F16>(args, m, s, ll);
    case LLMI_BF16: return launch_batch<BF16>(args, m, s, ll);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t llmi_launch_gemv_batch(const llmi_weight_s* const* ws, float* const* outs, int n, const llmi_act_s& a,
                                   cudaStream_t s, const GemvLL* ll) {
  return launch_gemv_batch_impl(ws, outs, n, a, s, ll, nullptr, 0.0f);
}

// Bytes [offset, offset + budget) of the concatenation of up to two matrices, taken as the same fraction window of
// each of a matrix' planes.  The CTAs of a mat-vec grid are scheduled in order — matrix by matrix, slab by slab, all
// planes of a slab together — so this is the order in which the mat-vec will read them.
cudaError_t llmi_launch_l2_prefetch(const llmi_weight_s* const* ws, int n, size_t offset_bytes, size_t budget_bytes,
                                    cudaStream_t s) {
  if (n < 1 || n > 2) return cudaErrorInvalidValue;
  if (budget_bytes == 0) return cudaSuccess;
  PrefetchArgs a;
  a.n = 0;
  uint64_t lines = 0;
  size_t cum = 0;
  for (int i = 0; i < n; ++i) {
    const llmi_weight_s& w = *ws[i];
    const size_t lo = std::max(offset_bytes, cum), hi = std::min(offset_bytes + budget_bytes, cum + w.bytes);
    cum += w.bytes;
    if (!w.base || hi <= lo) continue;
    const double f0 = double(lo - (cum - w.bytes)) / double(w.bytes), f1 = double(hi - (cum - w.bytes)) / double(w.bytes);
    const uint8_t* starts[3] = {w.p_q, w.p_d, w.p_x};
    const uint8_t* ends[3] = {w.p_d, w.p_x, w.base + w.bytes};  // planes are laid out q, d, x in one allocation
    for (int k = 0; k < 3; ++k) {
      const size_t plane = size_t(ends[k] - starts[k]);
      const size_t b0 = size_t(double(plane) * f0) & ~size_t(127), b1 = size_t(double(plane) * f1);
      if (b1 < b0 + 128) continue;
      a.p[a.n] = starts[k] + b0;
      a.lines[a.n] = uint32_t((b1 - b0) / 128);
      lines += a.lines[a.n];
      ++a.n;
    }
  }
  if (a.n == 0) return cudaSuccess;
  const unsigned blocks = unsigned(std::min<uint64_t>((lines + 255) / 256, uint64_t(g_sm_count) * 4));
  l2_prefetch_kernel<<<blocks, 256, 0, s>>>(a);
  return cudaGetLastError();
}

template <class B>
cudaError_t launch_geglu(const GemvArgs& g, const GemvArgs& u, uint8_t* act_out, uint32_t act_n, cudaStream_t s,
                         const GemvLL* ll) {
  GemvBatch b;
  b.n = 2;
  b.a[0] = g;
  b.a[1] = u;
  b.a[2] = g;
  if (ll) {
    b.peers = ll->peers;
    b.tag = ll->tag;
  }
  const uint32_t ctas = (g.n_slabs + 3) / 4;
  for (int i = 0; i < GEMV_MAX_BATCH; ++i) {
    b.S[i] = 4;
    b.cta_end[i] = ctas;
  }
  if (ctas == 0) return cudaSuccess;
  const size_t smem = g.act_bytes + size_t(8) * g.chunks * LLMI_SLAB * 4;
  if (smem > size_t(MAX_DYN_SMEM)) return cudaErrorInvalidValue;
  // 8 slabs x J items per CTA: 8 warps when a warp then still gets >= 2 items, else 4
  const bool w8 = uint64_t(8) * g.chunks >= 16;
  if (ll) {
    if (w8) return llmi_launch(gemv_geglu_kernel<B, 8, true>, dim3(ctas), dim3(256), smem, s, b, act_out, act_n);
    return llmi_launch(gemv_geglu_kernel<B, 4, true>, dim3(ctas), dim3(128), smem, s, b, act_out, act_n);
  }
  if (w8) return llmi_launch(gemv_geglu_kernel<B, 8, false>, dim3(ctas), dim3(256), smem, s, b, act_out, act_n);
  return llmi_launch(gemv_geglu_kernel<B, 4, false>, dim3(ctas), dim3(128), smem, s, b, act_out, act_n);
}

// ffn_gate + ffn_up + GEGLU (+ the Q8_0 quantizer of ffn_down's input) as ONE launch (gemv_geglu_kernel).  gate and up:
// same format, same shape, same row range.  Single GPU (ll == nullptr): act_out receives the ACT_Q8_0 activation of
// length gate.n_rows (a multiple of 32).  Row-sharded (ll): the hidden rows go, flagged, to ll->off[0] of every rank.
cudaError_t llmi_launch_gemv_geglu(const llmi_weight_s& gate, const llmi_weight_s& up, const llmi_act_s& a, uint8_t* act_out,
                                   cudaStream_t s, const GemvLL* ll) {
  if (gate.type != up.type || gate.n_cols != up.n_cols || gate.n_rows != up.n_rows || gate.row_begin != up.row_begin ||
      gate.row_end != up.row_end)
    return cudaErrorInvalidValue;
  if (!ll && (gate.n_rows % 32 || gate.row_begin != 0 || gate.row_end != gate.n_rows)) return cudaErrorInvalidValue;
  if (gate.n_slabs == 0) return cudaSuccess;
  GemvArgs g = make_args(gate, a, nullptr), u = make_args(up, a, nullptr);
  if (ll) g.ll_off = u.ll_off = ll->off[0];
  const uint32_t n = uint32_t(gate.n_rows);
  switch (gate.type) {
    case LLMI_Q4_0: return launch_geglu<Q4_0>(g, u, act_out, n, s, ll);
    case LLMI_Q8_0: return launch_geglu<Q8_0>(g, u, act_out, n, s, ll);
    case LLMI_Q5_0: return launch_geglu<Q5_0>(g, u, act_out, n, s, ll);
    case LLMI_Q4_K: return launch_geglu<Q4_K>(g, u, act_out, n, s, ll);
    case LLMI_Q6_K: return launch_geglu<Q6_K>(g, u, act_out, n, s, ll);
    case LLMI_F16: return launch_geglu<F16>(g, u, act_out, n, s, ll);
    case LLMI_BF16: return launch_geglu<BF16>(g, u, act_out, n, s, ll);
    default: return cudaErrorInvalidValue;
  }
}

// Token-batched form (prefill): `n_tok` activations of kind/length (act_kind, act_n), act_bytes() apart
// starting at act_base; matrix i writes token m's rows to outs[i] + m * out_strides[i].
cudaError_t llmi_launch_gemv_tokens(const llmi_weight_s* const* ws, float* const* outs, const uint32_t* out_strides,
                                    int n, int act_kind, uint64_t act_n, const uint8_t* act_base, uint32_t n_tok,
                                    cudaStream_t s, GemmPush* push) {
  if (push) push->done = false;
  if (n < 1 || n > GEMV_MAX_BATCH || n_tok == 0) return cudaErrorInvalidValue;
  llmi_act_s a;
  a.kind = act_kind;
  a.n = act_n;
  a.buf = const_cast<uint8_t*>(act_base);
  GemvArgs args[GEMV_MAX_BATCH];
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (ws[i]->type != ws[0]->type) return cudaErrorInvalidValue;
    if (ws[i]->n_slabs == 0) continue;
    args[m] = make_args(*ws[i], a, outs[i]);
    args[m].n_tok = n_tok;
    args[m].out_stride = out_strides[i];
    ++m;
  }
  if (m == 0) return cudaSuccess;
  switch (ws[0]->type) {
    case LLMI_Q4_0: return launch_tokens<Q4_0>(args, m, s, push);
    case LLMI_Q8_0: return launch_tokens<Q8_0>(args, m, s, push);
    case LLMI_Q5_0: return launch_tokens<Q5_0>(args, m, s, push);
    case LLMI_Q4_K: return launch_tokens<Q4_K>(args, m, s, push);
    case LLMI_Q6_K: return launch_tokens<Q6_K>(args, m, s, push);
    case LLMI_F16: return launch_tokens<F16>(args, m, s, push);
    case LLMI_BF16: return launch_tokens<BF16>(args, m, s, push);
    default: return cudaErrorInvalidValue;
  }
}

// The RMSNorm stage that feeds a batch of mat-vecs, fused into their launch as the ring kernel's prologue
// (gemv_ring.cuh): norm_act_kernel's stage — a 5 us hand-over + single-SM latency chain between two mat-vecs — becomes
// ~1.5 us of redundant work on every SM inside the consumer.  Bit-identical to norm_act_kernel + llmi_launch_gemv_batch.
// cudaErrorNotSupported: the fused form does not apply to this launch (the caller runs the two kernels instead).
cudaError_t llmi_launch_gemv_batch_norm(const llmi_weight_s* const* ws, float* const* outs, int n, const float* y,
                                        const float* w_post, const float* h_in, float* h_out, const float* w_norm, uint32_t n_cols,
                                        double eps, cudaStream_t s) {
  if (n < 1 || n > GEMV_MAX_BATCH || g_ring_mode == 1 || !y || h_in == h_out) return cudaErrorNotSupported;
  const uint32_t type = ws[0]->type;
  if (type != LLMI_Q4_0 && type != LLMI_Q8_0) return cudaErrorNotSupported;
  const uint32_t T = n_cols >= 2048 ? 1024u : 512u;
  if (n_cols % 32 || n_cols > T * RING_NORM_PER) return cudaErrorNotSupported;
  llmi_act_s a;
  a.kind = ACT_Q8_0;
  a.n = n_cols;
  GemvArgs args[GEMV_MAX_BATCH];
  for (int i = 0; i < n; ++i) {
    if (ws[i]->type != type || ws[i]->n_cols != n_cols || ws[i]->n_slabs == 0) return cudaErrorNotSupported;
    args[i] = make_args(*ws[i], a, outs[i]);
  }
  RingNorm nm;
  nm.y = y; nm.w_post = w_post; nm.h_in = h_in; nm.h_out = h_out; nm.w = w_norm; nm.n = n_cols; nm.T = T; nm.eps = eps;
  uint32_t rc = 0, rt = 0;
  int rd = 0;
  size_t rs = 0;
  if (type == LLMI_Q4_0) {
    if (!(g_ring_mode == 2 || ring_wanted<Q4_0>(args, n)) || !ring_plan<Q4_0>(args, n, rc, rd, rt, rs)) return cudaErrorNotSupported;
    return launch_ring<Q4_0>(args, n, s, nullptr, rc, rd, rt, rs, &nm);
  }
  if (!(g_ring_mode == 2 || ring_wanted<Q8_0>(args, n)) || !ring_plan<Q8_0>(args, n, rc, rd, rt, rs)) return cudaErrorNotSupported;
  return launch_ring<Q8_0>(args, n, s, nullptr, rc, rd, rt, rs, &nm);
}

// Throughput prefill only: out[token][row] = W . (gelu_tanh(gate[token]) * up[token]) for a token batch — ffn_down fed by
// the fp32 gate / up batches ([n_tok][n_cols] each) without the quantized GEGLU stage in between (gemm_bf16.cuh).
cudaError_t llmi_launch_fast_ffn_down(const llmi_weight_s& w, const float* gate, const float* up, float* out, uint32_t out_stride,
                                      uint32_t n_tok, cudaStream_t s, GemmPush* push, const void* hid16) {
  if (push) push->done = false;
  if (!g_prefill_fast || w.n_cols % fastmm::KB || w.n_slabs == 0) return cudaErrorInvalidValue;
  llmi_act_s none;
  GemvArgs g = make_args(w, none, out);
  g.n_tok = n_tok;
  g.out_stride = out_stride;
  switch (w.type) {
    case LLMI_Q4_0: return launch_fast<Q4_0>(&g, 1, s, gate, up, push, hid16);
    case LLMI_Q8_0: return launch_fast<Q8_0>(&g, 1, s, gate, up, push, hid16);
    case LLMI_Q5_0: return launch_fast<Q5_0>(&g, 1, s, gate, up, push, hid16);
    case LLMI_Q4_K: return launch_fast<Q4_K>(&g, 1, s, gate, up, push, hid16);
    case LLMI_Q6_K: return launch_fast<Q6_K>(&g, 1, s, gate, up, push, hid16);
    case LLMI_F16: return launch_fast<F16>(&g, 1, s, gate, up, push, hid16);
    case LLMI_BF16: return launch_fast<BF16>(&g, 1, s, gate, up, push, hid16);
    default: return cudaErrorInvalidValue;
  }
}

// Row-sharded token batch: GEGLU in place on this rank's columns of the gate batch (gemm_bf16.cuh geglu_cols_kernel).
cudaError_t llmi_launch_geglu_cols(float* gate, const float* up, uint32_t stride, uint32_t col0, uint32_t cols, uint32_t n_tok,
                                   bool fast, cudaStream_t s, const GemmPush* push, void* hid16) {
  if (cols == 0 || n_tok == 0) return cudaSuccess;
  const PeerOut po = llmi_peer_out(push, gate);
  if (((stride | col0 | cols) & 7u) == 0) {
    const uint64_t total = uint64_t(n_tok) * (cols / 8);
    const unsigned blocks = unsigned(std::min<uint64_t>((total + 255) / 256, uint64_t(g_sm_count) * 16));
    __nv_bfloat16* h16 = static_cast<__nv_bfloat16*>(hid16);
    const ptrdiff_t delta = hid16 ? reinterpret_cast<const char*>(hid16) - reinterpret_cast<const char*>(gate) : 0;
    if (fast && hid16)
      return llmi_launch(geglu_cols8_kernel<true, true>, dim3(blocks), dim3(256), 0, s, gate, up, stride, col0, cols, n_tok, po, h16, delta);
    if (hid16) return cudaErrorInvalidValue;  // bf16 hidden values exist in the throughput mode only
    return fast ? llmi_launch(geglu_cols8_kernel<true, false>, dim3(blocks), dim3(256), 0, s, gate, up, stride, col0, cols, n_tok, po, h16, delta)
                : llmi_launch(geglu_cols8_kernel<false, false>, dim3(blocks), dim3(256), 0, s, gate, up, stride, col0, cols, n_tok, po, h16, delta);
  }
  if (hid16) return cudaErrorInvalidValue;
  const uint64_t total = uint64_t(n_tok) * cols;
  const unsigned blocks = unsigned(std::min<uint64_t>((total + 255) / 256, uint64_t(g_sm_count) * 16));
  return fast ? llmi_launch(geglu_cols_kernel<true>, dim3(blocks), dim3(256), 0, s, gate, up, stride, col0, cols, n_tok, po)
              : llmi_launch(geglu_cols_kernel<false>, dim3(blocks), dim3(256), 0, s, gate, up, stride, col0, cols, n_tok, po);
}

cudaError_t llmi_launch_gemv(const llmi_weight_s& w, const llmi_act_s& a, float* out, cudaStream_t s) {
  const llmi_weight_s* ws[1] = {&w};
  float* outs[1] = {out};
  return llmi_launch_gemv_batch(ws, outs, 1, a, s);
}

// Logits mat-vec with the fused soft-cap + argmax epilogue: every CTA folds its
// rows into *key with one atomicMax per warp (key = ordered value bits << 32 |
// ~row, so the maximum is the first index of the largest logit).  The caller
// zeroes *key before and decodes it after (finish_token_kernel, glue.cu).
cudaError_t llmi_launch_gemv_argmax(const llmi_weight_s& w, const llmi_act_s& a, float* out,
                                    unsigned long long* key, float softcap, cudaStream_t s, const GemvLL* ll) {
  const llmi_weight_s* ws[1] = {&w};
  float* outs[1] = {out};
  return launch_gemv_batch_impl(ws, outs, 1, a, s, ll, key, softcap);
}

cudaError_t llmi_launch_block_dots(const llmi_weight_s& w, const llmi_act_s& a, int32_t* dots_dev, cudaStream_t s) {
  const GemvArgs g = make_args(w, a, nullptr);
  const uint64_t per_row = w.type == LLMI_Q6_K ? w.nb * 2 : (w.type == LLMI_Q4_K ? w.nb * 8 : w.nb);
  const uint64_t total = w.n_local * per_row;
  if (!total) return cudaSuccess;
  block_dots_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(g, w.type, dots_dev);
  return cudaGetLastError();
}

#ifdef LLMI_UMMA_TIMING
extern "C" int llmi_debug_umma_stamps(long long* out /*[10][160]*/) {
  return int(cudaMemcpyFromSymbol(out, g_umma_stamp, sizeof(long long) * 10 * 160));
}
#endif

TL_EXPORT(llmi_debug_timeline_gemv)

#ifdef LLMI_GEMV_TIMING
extern "C" int llmi_debug_gemv_stamps(unsigned long long* out /*[256][6]*/, unsigned* n_launches, int reset) {
  if (out) cudaMemcpyFromSymbol(out, g_gemv_stamp, sizeof(unsigned long long) * 256 * 6);
  if (n_launches) cudaMemcpyFromSymbol(n_launches, g_gemv_launch, 4);
  if (reset) {
    const unsigned z = 0;
    cudaMemcpyToSymbol(g_gemv_launch, &z, 4);
  }
  return 0;
}
#endif
