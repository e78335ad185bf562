// Throughput prefill (opt-in, LLMI_PREFILL=fast / llmi_set_prefill_mode(1)): the dequantize-to-bf16 tcgen05 GEMM that
// north_star sketches for M >= 16 — included by gemv.cu (inside its anonymous namespace, after umma_prefill.cuh).
//
// This is NOT the parity path.  The reference multiplies int8 activation quants by int4/int8 weight quants exactly and
// rounds once per 32-element block (ops.cpp:373-396); bf16 rounds every operand to 8 mantissa bits.  The exact
// tensor-core form (umma_prefill.cuh: kind::i8, one MMA per quant block, fp32 chain epilogue) stays the default and
// the gate; its per-block CUDA-core epilogue caps the tensor pipe at ~15 % by construction.  This mode trades the
// bits for the pipe: out[token][row] = sum_k bf16(w_deq[row][k]) * bf16(x_hat[token][k]) with fp32 accumulation in
// tensor memory, where w_deq is the reference's own dequantization of the weight (d * (nib - 8), d * sc * q - dmin * m,
// ...) and x_hat the value the reference's quantized activation stands for (d_x * q_x): every format the path
// supports, one kernel.  Tolerance (tests/test_gemv_gpu.py): |o - o_exact| <= 2e-2 * max|o_exact| per call; greedy
// tokens are reported against the exact path, not asserted identical.
//
// Three launches per matrix:
//   fast_dequant_kernel   weight planes (slab layout) -> bf16 in UMMA operand order: [row tile 128][K block 64] stages of
//                         16 KB = 8 K-core-columns x 16 row groups x (8 rows x 16 bytes) — K-major, no swizzle;
//   fast_pack_act_kernel  quantized activations of the token batch -> bf16 in the same order, token tiles of TNF;
//   gemm_bf16_kernel      CTA = 128 rows x TNF (128 / 256) tokens: warp 0 feeds a 4-stage ring with ONE bulk copy per
//                         operand per stage (both operands are contiguous in this order: no tensor map needed), warp 1
//                         owns tensor memory and issues tcgen05.mma.kind::f16 M128 x N(TNF) x K16, four per stage, then
//                         commits the stage back to the producer; warps 2-5 read the fp32 accumulators
//                         (tcgen05.ld 32x32b) and store out[token][row] — 32 lanes = 32 consecutive rows per store.
// The dequantized weights live in a scratch buffer for the duration of one call (2 bytes per weight: 231 MB for the
// 27b gate matrix); with >= 1024 tokens per batch the round trip is a few per cent of the GEMM.
#pragma once

namespace fastmm {

constexpr int TM = 128, KB = 64, NST = 4;
constexpr uint32_t A_BYTES = TM * KB * 2;  // 16 KB
template <int TNF>
struct Cfg {
  static constexpr uint32_t B_BYTES = TNF * KB * 2, STAGE = A_BYTES + B_BYTES;
  static constexpr size_t SMEM = size_t(NST) * STAGE + 1024;
  // cute::UMMA::InstrDescriptor: c_format F32 (1) @4, a/b format BF16 (1) @7/@10, K-major both, N>>3 @17, M>>4 @24
  static constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TNF >> 3) << 17) | (uint32_t(TM >> 4) << 24);
};

__device__ __forceinline__ uint32_t bf16x2(float lo, float hi) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(bf16x2(v[0], v[1]), bf16x2(v[2], v[3]), bf16x2(v[4], v[5]), bf16x2(v[6], v[7]));
}

// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, M128 x N x K16
__device__ __forceinline__ void mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

}  // namespace fastmm

// Elements 8*kg .. 8*kg+7 of local row `row` of a repacked matrix, dequantized as the reference does.
__device__ __forceinline__ void fast_dequant8(const GemvArgs& a, uint32_t type, uint32_t row, uint32_t kg, float (&v)[8]) {
  const uint32_t s = row >> 3, r = row & 7, nb = a.nb;
  if (type == LLMI_Q4_0) {  // ops.cpp:1005-1020: d * (nibble - 8); byte j = element j (low) and j + 16 (high)
    const uint32_t b = kg >> 2, g = kg & 3;
    const size_t cell = (size_t(s) * nb + b) * 8 + r;
    const uint2 w = *reinterpret_cast<const uint2*>(a.q + cell * 16 + (g & 1) * 8);
    const float d = h2f(reinterpret_cast<const uint16_t*>(a.d)[cell]);
    const uint32_t sh = g >= 2 ? 4u : 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t word = i < 4 ? w.x : w.y;
      v[i] = d * float(int((word >> (8 * (i & 3) + sh)) & 0xfu) - 8);
    }
  } else if (type == LLMI_Q8_0) {
    const uint32_t b = kg >> 2, g = kg & 3;
    const size_t su = size_t(s) * nb + b;
    const uint2 w = *reinterpret_cast<const uint2*>(a.q + ((su * 2 + (g >> 1)) * 8 + r) * 16 + (g & 1) * 8);
    const float d = h2f(reinterpret_cast<const uint16_t*>(a.d)[su * 8 + r]);
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = d * float(int(int8_t(((i < 4 ? w.x : w.y) >> (8 * (i & 3))) & 0xffu)));
  } else if (type == LLMI_Q4_K) {  // ops.cpp:1043-1062: d * sc * nibble - dmin * m per 32-element sub-block
    const uint32_t u = kg >> 5, e0 = (kg & 31) * 8, c = e0 >> 6, p0 = e0 & 63;
    const size_t su = size_t(s) * nb + u;
    const uint4 h = *reinterpret_cast<const uint4*>(a.x + (su * 8 + r) * 16);
    const uint32_t hh = (p0 & 31) >> 4, off = p0 & 15;
    const uint2 w = *reinterpret_cast<const uint2*>(a.q + (((su * 4 + c) * 2 + hh) * 8 + r) * 16 + off);
    int sc, mn;
    q4_k_scale_min(h, int(2 * c + (p0 >> 5)), sc, mn);
    const float d = h2f(uint16_t(h.x & 0xffffu)) * float(sc), m = h2f(uint16_t(h.x >> 16)) * float(mn);
    const uint32_t sh = p0 >= 32 ? 4u : 0u;
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = d * float(((i < 4 ? w.x : w.y) >> (8 * (i & 3) + sh)) & 0xfu) - m;
  } else if (type == LLMI_F16 || type == LLMI_BF16) {
    const uint4 w = *reinterpret_cast<const uint4*>(a.q + ((size_t(s) * nb + kg) * 8 + r) * 16);
    const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint16_t hbits = uint16_t(ws[i >> 1] >> (16 * (i & 1)));
      v[i] = type == LLMI_F16 ? h2f(hbits) : __uint_as_float(uint32_t(hbits) << 16);
    }
  } else {  // Q6_K, Q5_0: element by element through the embedding-row dequantizer (glue_device.cuh)
    EmbedArgs ea{type, nb, a.n_cols, a.q, a.d, a.x, 0, 0};
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = dequant_elem(ea, row, kg * 8 + i);
  }
}

// out: [row tile][K block][kc 8][row group 16][r 8] x 16 bytes; rows past n_local are zero
__global__ void fast_dequant_kernel(const GemvArgs a, uint32_t type, uint32_t n_tiles, uint32_t nkb, uint4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const uint64_t total = uint64_t(n_tiles) * nkb * 1024;
  for (uint64_t o = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; o < total; o += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t r = uint32_t(o & 7), rg = uint32_t(o >> 3) & 15, kc = uint32_t(o >> 7) & 7;
    const uint64_t tb = o >> 10;
    const uint32_t kb = uint32_t(tb % nkb), rt = uint32_t(tb / nkb);
    const uint32_t row = rt * fastmm::TM + rg * 8 + r;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (row < a.n_local) fast_dequant8(a, type, row, kb * 8 + kc, v);
    out[o] = fastmm::pack8(v);
  }
}

// Activations of the token batch (the kind the exact path prepared) -> what they stand for, in bf16, operand order:
// [token tile TNF][K block][kc 8][token group TNF/8][t 8] x 16 bytes; tokens past n_tok are zero
__global__ void fast_pack_act_kernel(const uint8_t* __restrict__ act, uint32_t act_stride, int kind, uint32_t n_cols, uint32_t n_tok,
                                     uint32_t tnf, uint32_t n_ttiles, uint32_t nkb, uint4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const uint32_t per_stage = tnf * 8;  // 16-byte items
  const uint64_t total = uint64_t(n_ttiles) * nkb * per_stage;
  for (uint64_t o = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; o < total; o += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t in = uint32_t(o % per_stage);
    const uint64_t tb = o / per_stage;
    const uint32_t kb = uint32_t(tb % nkb), tt = uint32_t(tb / nkb);
    const uint32_t t = in & 7, tg = (in >> 3) % (tnf / 8), kc = in / tnf;
    const uint32_t tok = tt * tnf + tg * 8 + t, k0 = (kb * 8 + kc) * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tok < n_tok) {
      const uint8_t* p = act + size_t(tok) * act_stride;
      if (kind == ACT_Q8_0 || kind == ACT_Q8_K) {
        const uint2 q = *reinterpret_cast<const uint2*>(p + k0);
        const float d = kind == ACT_Q8_0 ? h2f(uint16_t(reinterpret_cast<const uint32_t*>(p + n_cols)[k0 >> 5] & 0xffffu))
                                         : reinterpret_cast<const float*>(p + n_cols + n_cols / 8)[k0 >> 8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = d * float(int(int8_t(((i < 4 ? q.x : q.y) >> (8 * (i & 3))) & 0xffu)));
      } else if (kind == ACT_F16) {
        const uint4 w = *reinterpret_cast<const uint4*>(p + size_t(k0) * 2);
        const uint32_t ws[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = h2f(uint16_t(ws[i >> 1] >> (16 * (i & 1))));
      } else if (kind == ACT_BF16_RAW) {  // already bf16 (a sharded model's hidden batch): re-ordered, not re-rounded
        out[o] = *reinterpret_cast<const uint4*>(p + size_t(k0) * 2);
        continue;
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = reinterpret_cast<const float*>(p)[k0 + i];
      }
    }
    out[o] = fastmm::pack8(v);
  }
}

// The throughput mode's GEGLU (model.cpp:887-901 with a fast tanh).  Every operation rounds on its own (no contraction),
// so the value does not depend on the kernel it is inlined into: a row-sharded model computes it on the rank that
// owns the column (geglu_cols_kernel) and must get the bits of the fused pack below.
__device__ __forceinline__ float fast_geglu(float x, float u) {
  const float x3 = __fmul_rn(__fmul_rn(__fmul_rn(0.044715f, x), x), x);
  const float inner = __fmul_rn(0.7978845608028654f, __fadd_rn(x, x3));
  const float th = __fsub_rn(1.0f, __fdiv_rn(2.0f, __fadd_rn(__expf(__fmul_rn(2.0f, inner)), 1.0f)));  // tanh; exp overflow -> 1, underflow -> -1
  return __fmul_rn(__fmul_rn(__fmul_rn(0.5f, x), __fadd_rn(1.0f, th)), u);
}

// hidden = gelu_tanh(gate) * up of a token batch, straight into the down-projection's bf16 operand order — in this mode
// the GEGLU stage neither quantizes (geglu_act_kernel: 11 % of a fast prompt, most of it the Q8_0 quantizer's shuffles
// and IEEE divisions) nor takes the detour through fast_pack_act_kernel.  The reference's formula (model.cpp:887-901).
__global__ void fast_geglu_pack_kernel(const float* __restrict__ gate, const float* __restrict__ up, uint32_t n_cols, uint32_t n_tok,
                                       uint32_t tnf, uint32_t n_ttiles, uint32_t nkb, uint4* __restrict__ out) {
  pdl_trigger();
  pdl_wait();
  const uint32_t per_stage = tnf * 8;
  const uint64_t total = uint64_t(n_ttiles) * nkb * per_stage;
  for (uint64_t o = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; o < total; o += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t in = uint32_t(o % per_stage);
    const uint64_t tb = o / per_stage;
    const uint32_t kb = uint32_t(tb % nkb), tt = uint32_t(tb / nkb);
    const uint32_t t = in & 7, tg = (in >> 3) % (tnf / 8), kc = in / tnf;
    const uint32_t tok = tt * tnf + tg * 8 + t, k0 = (kb * 8 + kc) * 8;
    float v[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (tok < n_tok) {
      const float4* g4 = reinterpret_cast<const float4*>(gate + size_t(tok) * n_cols + k0);
      const float4* u4 = reinterpret_cast<const float4*>(up + size_t(tok) * n_cols + k0);
      const float4 ga = g4[0], gb = g4[1], ua = u4[0], ub = u4[1];
      const float gs[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w}, us[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) v[i] = fast_geglu(gs[i], us[i]);
    }
    out[o] = fastmm::pack8(v);
  }
}

// Row-sharded token batch: gate[token][c] = GEGLU(gate[token][c], up[token][c]) in place for this rank's columns
// [col0, col0 + cols) — the rank that computed a gate / up column pair combines it, so only the hidden column travels
// (half the exchange) and the GEGLU arithmetic is split over the ranks.  FAST: the throughput mode's formula, else the
// reference's operation order (glue_device.cuh geglu()).
template <bool FAST>
__global__ void geglu_cols_kernel(float* __restrict__ gate, const float* __restrict__ up, uint32_t stride, uint32_t col0,
                                  uint32_t cols, uint32_t n_tok, PeerOut po) {
  pdl_trigger();
  pdl_wait();
  const uint64_t total = uint64_t(n_tok) * cols;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const size_t e = size_t(i / cols) * stride + col0 + uint32_t(i % cols);
    const float hv = FAST ? fast_geglu(gate[e], up[e]) : geglu(gate[e], up[e]);
    gate[e] = hv;
    for (uint32_t j = 0; j < po.n; ++j) po.p[j][e] = hv;  // (po.n > 0: the hidden columns go to every peer right here)
  }
  if (po.n) __threadfence_system();
}
// The same, eight columns per thread (col0, cols and stride multiples of 8: 16-byte stores over NVLink).  H16: the
// hidden values leave as bf16 into `hid16` ([token][stride] halves, local and peers) instead — the throughput mode's
// ffn_down rounds its activation to bf16 anyway (fast_pack_act_kernel), so rounding before the exchange gives the same
// operand bits and halves the bytes again.
template <bool FAST, bool H16>
__global__ void geglu_cols8_kernel(float* __restrict__ gate, const float* __restrict__ up, uint32_t stride, uint32_t col0,
                                   uint32_t cols, uint32_t n_tok, PeerOut po, __nv_bfloat16* __restrict__ hid16,
                                   ptrdiff_t hid16_delta /* bytes from the gate batch to the hid16 batch, same on every rank */) {
  pdl_trigger();
  pdl_wait();
  const uint32_t per = cols / 8;
  const uint64_t total = uint64_t(n_tok) * per;
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < total; i += uint64_t(gridDim.x) * blockDim.x) {
    const size_t e = size_t(i / per) * stride + col0 + uint32_t(i % per) * 8;
    const float4 ga = *reinterpret_cast<const float4*>(gate + e), gb = *reinterpret_cast<const float4*>(gate + e + 4);
    const float4 ua = *reinterpret_cast<const float4*>(up + e), ub = *reinterpret_cast<const float4*>(up + e + 4);
    const float gs[8] = {ga.x, ga.y, ga.z, ga.w, gb.x, gb.y, gb.z, gb.w}, us[8] = {ua.x, ua.y, ua.z, ua.w, ub.x, ub.y, ub.z, ub.w};
    float v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = FAST ? fast_geglu(gs[k], us[k]) : geglu(gs[k], us[k]);
    if (H16) {
      const uint4 w = fastmm::pack8(v);
      *reinterpret_cast<uint4*>(hid16 + e) = w;
      for (uint32_t j = 0; j < po.n; ++j)
        *reinterpret_cast<uint4*>(reinterpret_cast<char*>(po.p[j]) + hid16_delta + e * 2) = w;
    } else {
      const float4 va = make_float4(v[0], v[1], v[2], v[3]), vb = make_float4(v[4], v[5], v[6], v[7]);
      *reinterpret_cast<float4*>(gate + e) = va;
      *reinterpret_cast<float4*>(gate + e + 4) = vb;
      for (uint32_t j = 0; j < po.n; ++j) {
        *reinterpret_cast<float4*>(po.p[j] + e) = va;
        *reinterpret_cast<float4*>(po.p[j] + e + 4) = vb;
      }
    }
  }
  if (po.n) __threadfence_system();  // (as in the GEMM epilogue: visible on the peers before the barrier kernel's flags)
}

// grid = (token tiles, row tiles) — the token tiles of one row tile run side by side, so its dequantized weights are
// read from DRAM once and from L2 by the others (with the tiles the other way round the 21504 x 5376 matrix was read
// once per token tile: 945 MB of DRAM reads for 231 MB of weights, profiles/r02_ncu_gemm_bf16_gate27b.csv);
// A / Bt: the operand-order buffers of the two kernels above
template <int TNF>
__global__ void __launch_bounds__(192, 1)
gemm_bf16_kernel(const uint8_t* __restrict__ A, const uint8_t* __restrict__ Bt, float* __restrict__ out, uint32_t out_stride,
                 uint32_t n_rows, uint32_t n_tok, uint32_t nkb, PeerOut po) {
  using namespace fastmm;
  using C = Cfg<TNF>;
  extern __shared__ __align__(1024) uint8_t fsm_raw[];
  __shared__ __align__(8) uint64_t full[NST], empty[NST], acc_full;
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rt = blockIdx.y, tt = blockIdx.x;
  pdl_trigger();
  uint8_t* smem = fsm_raw + ((1024u - (smem_u32(fsm_raw) & 1023u)) & 1023u);
  if (threadIdx.x == 0) {
    for (int s = 0; s < NST; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    mbar_init(&acc_full, 1);
  }
  if (warp == 1) {  // tensor memory: TNF fp32 columns x 128 lanes = the tile's accumulators
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(uint32_t(TNF))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  umma::tc_fence_before();
  __syncthreads();
  umma::tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();  // both operand buffers come from the predecessor kernels
  if (warp == 0) {
    if (lane == 0) {
      const uint8_t* a_src = A + size_t(rt) * nkb * A_BYTES;
      const uint8_t* b_src = Bt + size_t(tt) * nkb * C::B_BYTES;
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        const uint32_t s = kb % NST;
        if (kb >= NST) mbar_wait(&empty[s], ((kb / NST) & 1) ^ 1);
        mbar_expect_tx(&full[s], C::STAGE);
        bulk_g2s(smem + size_t(s) * C::STAGE, a_src + size_t(kb) * A_BYTES, A_BYTES, &full[s]);
        bulk_g2s(smem + size_t(s) * C::STAGE + A_BYTES, b_src + size_t(kb) * C::B_BYTES, C::B_BYTES, &full[s]);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      // K-major, no swizzle: core matrix (kc, group) of a stage sits at (kc * groups + group) * 128 bytes, so the K
      // stride between core columns (lbo) is groups * 128 and the stride between 8-row groups (sbo) 128
      constexpr uint32_t A_LBO = (TM / 8) * 128, B_LBO = (TNF / 8) * 128;
      const uint64_t a_hi = (uint64_t(128 >> 4) << 32) | (uint64_t(A_LBO >> 4) << 16) | (1ull << 46);
      const uint64_t b_hi = (uint64_t(128 >> 4) << 32) | (uint64_t(B_LBO >> 4) << 16) | (1ull << 46);
      const uint32_t a_lo0 = (smem_u32(smem) & 0x3ffffu) >> 4, b_lo0 = (smem_u32(smem + A_BYTES) & 0x3ffffu) >> 4;
      for (uint32_t kb = 0; kb < nkb; ++kb) {
        const uint32_t s = kb % NST;
        mbar_wait(&full[s], (kb / NST) & 1);
        umma::tc_fence_after();
        const uint64_t ad = a_hi | uint64_t(a_lo0 + s * (C::STAGE >> 4));
        const uint64_t bd = b_hi | uint64_t(b_lo0 + s * (C::STAGE >> 4));
#pragma unroll
        for (uint32_t k4 = 0; k4 < KB / 16; ++k4)  // one MMA = K 16 = two core columns
          mma_bf16(tmem_base, ad + uint64_t(k4 * ((2 * A_LBO) >> 4)), bd + uint64_t(k4 * ((2 * B_LBO) >> 4)), C::IDESC, (kb | k4) ? 1u : 0u);
        umma::tc_commit(&empty[s]);  // arrives when these MMAs have read the stage
      }
      umma::tc_commit(&acc_full);
    }
  } else {
    // epilogue: warp w reads TMEM lanes 32 * (w % 4) .. + 31 (the hardware's lane-quarter rule); thread = row
    const uint32_t q = warp & 3;
    const uint32_t row = rt * TM + q * 32 + lane;
    mbar_wait(&acc_full, 0);
    umma::tc_fence_after();
#pragma unroll 1
    for (uint32_t c0 = 0; c0 < uint32_t(TNF); c0 += 16) {
      int v[16];
      umma::tmem_ld8_issue(tmem_base + ((q * 32u) << 16) + c0, v);
      umma::tmem_wait_ld(v);
      const uint32_t tok0 = tt * TNF + c0;
      if (row < n_rows) {
#pragma unroll
        for (int k = 0; k < 16; ++k)
          if (tok0 + k < n_tok) out[size_t(tok0 + k) * out_stride + row] = __int_as_float(v[k]);
        // row-sharded model: the same 128-byte row segments into every peer's batch over NVLink (posted stores: the
        // all-gather rides under the other CTAs' MMAs; a barrier kernel follows the launch)
        for (uint32_t j = 0; j < po.n; ++j) {
          float* dst = po.p[j];
#pragma unroll
          for (int k = 0; k < 16; ++k)
            if (tok0 + k < n_tok) dst[size_t(tok0 + k) * out_stride + row] = __int_as_float(v[k]);
        }
      }
    }
    // the peers' copies are system-scope visible before this grid can be seen as complete: the barrier kernel that
    // follows raises the flags the peers' consumers wait for
    if (po.n) __threadfence_system();
  }
  umma::tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(TNF)) : "memory");
}
