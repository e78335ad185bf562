// Prefill on the 5th-generation tensor cores, EXACT: tcgen05.mma kind::i8 for Q4_0 / Q8_0 weights.
// Included by gemv.cu (uses its GemvArgs / GemvBatch and the part[] / toklane_reduce_kernel contract).
//
// The reference has no batched matmul (its forward loops tokens around mat_vec_mul, model.cpp:714-960), and its
// arithmetic is: exact integer dot of one 32-element block (int8 activation quants x int4/int8 weight quants),
// ONE fp32 rounding per block (scale product, fma), blocks accumulated in a fixed order.  A dequantize-to-bf16
// GEMM cannot reproduce that; an int8 MMA with K = 32 can: one tcgen05.mma per quant block gives the int32
// block dots of a 128-row x 32-token tile in tensor memory, bit-identical to the dp4a dots, and the fp32 part
// (float(dot), dw*dx, fma into the chain of the canonical summation order) is the epilogue.  Results are
// bit-identical to the one-token kernel; the tensor pipe is ~25 % busy by construction (the per-block fp32
// epilogue on the CUDA cores is the bound — 4 ops per output per block vs ~14 for the dp4a kernel).
//
// CTA = 128 weight rows (16 slabs) x 32 tokens x a range of K-chunks, 12 warps, warp-specialized:
//   warp 0      producer: bulk async copies (UBLKCP) of the stage's weights (one contiguous run per slab: the
//               Q8_0 slab layout IS the UMMA K-major core-matrix layout, 8 rows x 16 bytes), of the token tile's
//               activation quants (pre-packed into core-matrix order by umma_pack_act_kernel) and fp32 scales
//   warp 1      tensor-memory allocation; one lane issues tcgen05.mma (M128 N32 K32, no accumulate) per block
//               into a ring of 4 TMEM stages and commits to mbarriers
//   warps 2-3   Q4_0 only: unpack the nibbles of a stage to signed int8 core matrices (generic -> async proxy fence)
//   warps 4-11  epilogue: tcgen05.ld the block dots (thread = row, 16 token columns), four chain accumulators
//               per (row, token) = the four sub-lanes of gemv_slab_kernel, chunk partial (s0+s1)+(s2+s3) stored
//               to part[chunk][token][row]; toklane_reduce_kernel adds the chunk partials left to right.
// Shared-memory ring: 3 stages of 8 blocks (A 32 KB int8 + B 8 KB + scales 1 KB, + 16 KB packed nibbles for Q4_0).
#pragma once

namespace umma {

constexpr int TM = 128, TN = 32;          // tile: weight rows x tokens
constexpr int SB = 8;                     // quant blocks per shared-memory stage
constexpr int NSTAGE = 3, NTMEM = 4;      // smem ring depth, TMEM ring depth (x TN columns)
constexpr int WARPS = 12, EPI_WARP0 = 4;
constexpr uint32_t A_BYTES = TM * SB * 32, B_BYTES = TN * SB * 32, D_BYTES = TN * SB * 4, P_BYTES = TM * SB * 16;
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + D_BYTES + P_BYTES;  // 58368, a multiple of 1024
constexpr size_t SMEM_BYTES = size_t(NSTAGE) * STAGE_BYTES + 1024;

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// tcgen05.commit: the mbarrier gets one arrival when every MMA this thread issued so far has completed
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
// K-major, no swizzle: core matrix = 8 rows x 16 bytes (128 contiguous bytes); lbo = byte distance between the two
// 16-byte K halves of a 32-byte block row, sbo = byte distance between 8-row groups (cute::UMMA::SmemDescriptor)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (1ull << 46);
}
// D[tmem] = A[smem] * B[smem]^T, int8 x int8 -> int32, M128 N32 K32, overwrite (no accumulate)
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(0u), "r"(0u)
      : "memory");
}
// cute::UMMA::InstrDescriptor: c_format S32 (2) @4, a/b format INT8 (1) @7/@10, K-major both, N>>3 @17, M>>4 @24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TN >> 3) << 17) | (uint32_t(TM >> 4) << 24);

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, int (&v)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      "tcgen05.wait::ld.sync.aligned;\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
// exact int -> float for |x| < 2^22 on the full-rate pipes (I2F is a quarter-rate conversion)
__device__ __forceinline__ float int_to_float(int x) { return __int_as_float(0x4B400000 + x) - 12582912.0f; }

}  // namespace umma

// Activations of a token batch -> UMMA operand order.  bq: [token tile of 32][block][token/8][K half][token%8][16 B]
// (1 KB per (tile, block): four K-major core-matrix pairs), bd: [tile][block][32] fp32 scales.  Tokens past n_tok: 0.
__global__ void umma_pack_act_kernel(const uint8_t* __restrict__ act, uint32_t act_stride, uint32_t n_cols, uint32_t nb,
                                     uint32_t n_tok, uint4* __restrict__ bq, float* __restrict__ bd) {
  pdl_trigger();
  pdl_wait();
  const uint32_t tiles = (n_tok + umma::TN - 1) / umma::TN;
  const uint64_t items = uint64_t(tiles) * nb * 64;  // 16-byte items
  for (uint64_t i = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x; i < items; i += uint64_t(gridDim.x) * blockDim.x) {
    const uint32_t nn = uint32_t(i & 7), h = uint32_t(i >> 3) & 1, n8 = uint32_t(i >> 4) & 3;
    const uint64_t tb = i >> 6;
    const uint32_t b = uint32_t(tb % nb), tile = uint32_t(tb / nb), tok = tile * umma::TN + n8 * 8 + nn;
    uint4 v = make_uint4(0, 0, 0, 0);
    if (tok < n_tok) v = *reinterpret_cast<const uint4*>(act + size_t(tok) * act_stride + size_t(b) * 32 + h * 16);
    bq[i] = v;
    if (h == 0) {
      float d = 0.0f;
      if (tok < n_tok)
        d = h2f(uint16_t(reinterpret_cast<const uint32_t*>(act + size_t(tok) * act_stride + n_cols)[b] & 0xffffu));
      bd[tb * umma::TN + n8 * 8 + nn] = d;
    }
  }
}

// grid = (row tiles of all matrices of the batch, token tiles, K-chunk groups of `nj` chunks)
template <bool IS_Q8>
__global__ void __launch_bounds__(umma::WARPS * 32, 1)
gemm_umma_kernel(const GemvBatch batch, const uint4* __restrict__ bq, const float* __restrict__ bd, uint32_t nj) {
  using namespace umma;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[NSTAGE], pk_full[NSTAGE], empty[NSTAGE], tfull[NTMEM], tempty[NTMEM];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  int mi = 0;
  while (mi + 1 < batch.n && blockIdx.x >= batch.cta_end[mi]) ++mi;
  const GemvArgs& a = batch.a[mi];
  const uint32_t tile_m = blockIdx.x - (mi ? batch.cta_end[mi - 1] : 0u), tile_n = blockIdx.y;
  pdl_trigger();
  const uint32_t nb = a.nb, J = (nb + 15) / 16;
  const uint32_t j0 = blockIdx.z * nj, j1 = min(J, j0 + nj);
  const bool active = j0 < J;  // matrices of one launch may differ in K
  const uint32_t b_begin = j0 * 16, b_end = active ? min(nb, j1 * 16) : b_begin;
  const uint32_t n_blk = b_end - b_begin, n_st = (n_blk + SB - 1) / SB;
  const uint32_t slab0 = tile_m * (TM / 8), n_sl = min(uint32_t(TM / 8), a.n_slabs - slab0);
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  auto stA = [&](uint32_t s) { return smem + size_t(s) * STAGE_BYTES; };                      // [slab 16][blk 8][h 2][r 8][16]
  auto stB = [&](uint32_t s) { return smem + size_t(s) * STAGE_BYTES + A_BYTES; };            // [blk 8][n/8 4][h 2][n%8 8][16]
  auto stD = [&](uint32_t s) { return reinterpret_cast<float*>(smem + size_t(s) * STAGE_BYTES + A_BYTES + B_BYTES); };  // [blk 8][32]
  auto stP = [&](uint32_t s) { return smem + size_t(s) * STAGE_BYTES + A_BYTES + B_BYTES + D_BYTES; };  // [slab 16][blk 8][r 8][16] nibbles

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], IS_Q8 ? 1 : 1 + 64);   // producer's expect_tx (+ the 64 unpack threads)
      mbar_init(&pk_full[s], 1);
      mbar_init(&empty[s], 1 + (WARPS - EPI_WARP0));  // MMA commit + the epilogue warps (done with the scales)
    }
    for (int t = 0; t < NTMEM; ++t) {
      mbar_init(&tfull[t], 1);
      mbar_init(&tempty[t], WARPS - EPI_WARP0);
    }
  }
  if (warp == 1) {  // tensor memory: NTMEM stages of TN int32 columns (128 lanes = the tile's rows)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)),
                 "r"(uint32_t(NTMEM * TN))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  pdl_wait();  // the packed activations come from the predecessor kernel

  if (warp == 0) {
    // ------------------------------------------------------------------ producer
    if (lane == 0) {
      for (uint32_t st = 0; st < n_st; ++st) {
        const uint32_t s = st % NSTAGE, ph = (st / NSTAGE) & 1;
        if (st >= NSTAGE) mbar_wait(&empty[s], ph ^ 1);
        const uint32_t b0 = b_begin + st * SB, nbs = min(uint32_t(SB), b_end - b0);
        const uint32_t w_run = nbs * (IS_Q8 ? 256u : 128u);  // bytes of one slab's blocks [b0, b0+nbs): contiguous
        const uint32_t tb = (tile_n * nb + b0);
        if (IS_Q8) {
          mbar_expect_tx(&full[s], n_sl * w_run + nbs * (1024u + 128u));
          for (uint32_t sl = 0; sl < n_sl; ++sl)
            bulk_g2s(stA(s) + sl * (SB * 256), a.q + (size_t(slab0 + sl) * nb + b0) * 256, w_run, &full[s]);
        } else {
          mbar_expect_tx(&pk_full[s], n_sl * w_run);
          for (uint32_t sl = 0; sl < n_sl; ++sl)
            bulk_g2s(stP(s) + sl * (SB * 128), a.q + (size_t(slab0 + sl) * nb + b0) * 128, w_run, &pk_full[s]);
          mbar_expect_tx(&full[s], nbs * (1024u + 128u));
        }
        bulk_g2s(stB(s), bq + size_t(tb) * 64, nbs * 1024u, &full[s]);
        bulk_g2s(stD(s), bd + size_t(tb) * TN, nbs * 128u, &full[s]);
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    for (uint32_t i = 0; i < n_blk; ++i) {
      const uint32_t st = i / SB, ib = i % SB, s = st % NSTAGE, t = i % NTMEM;
      if (ib == 0) mbar_wait(&full[s], (st / NSTAGE) & 1);
      if (i >= NTMEM) mbar_wait(&tempty[t], ((i / NTMEM) & 1) ^ 1);
      tc_fence_after();
      if (lane == 0) {
        const uint64_t ad = smem_desc(smem_u32(stA(s)) + ib * 256, 128, SB * 256);
        const uint64_t bdsc = smem_desc(smem_u32(stB(s)) + ib * 1024, 128, 256);
        mma_i8(tmem_base + t * TN, ad, bdsc, IDESC);
        tc_commit(&tfull[t]);
        if (ib == SB - 1 || i == n_blk - 1) tc_commit(&empty[s]);
      }
      __syncwarp();
    }
  } else if (warp < EPI_WARP0) {
    // ------------------------------------------- Q4_0: nibbles -> signed int8 core matrices
    if (!IS_Q8) {
      const uint32_t ut = threadIdx.x - 64;  // 0..63
      for (uint32_t st = 0; st < n_st; ++st) {
        const uint32_t s = st % NSTAGE, ph = (st / NSTAGE) & 1;
        const uint32_t nbs = min(uint32_t(SB), b_end - (b_begin + st * SB));
        if (st >= NSTAGE) mbar_wait(&empty[s], ph ^ 1);  // the MMAs that read this stage's A are done
        mbar_wait(&pk_full[s], ph);
        for (uint32_t it = ut; it < n_sl * nbs * 8; it += 64) {  // item = (slab, blk, r): 16 nibble bytes
          const uint32_t r = it & 7, blk = (it >> 3) % nbs, sl = (it >> 3) / nbs;
          const uint4 w = *reinterpret_cast<const uint4*>(stP(s) + sl * (SB * 128) + blk * 128 + r * 16);
          // byte j: element j (low nibble) and element j+16 (high nibble); value = nibble - 8 as a signed byte:
          // flip bit 3 (4-bit two's complement of nibble-8), then sign-extend the 4-bit field to 8 bits
          auto sx = [](uint32_t n4) {
            const uint32_t t = n4 ^ 0x08080808u;
            return t | (((t >> 3) & 0x01010101u) * 0xf0u);
          };
          uint4 lo, hi;
          lo.x = sx(w.x & 0x0f0f0f0fu); lo.y = sx(w.y & 0x0f0f0f0fu); lo.z = sx(w.z & 0x0f0f0f0fu); lo.w = sx(w.w & 0x0f0f0f0fu);
          hi.x = sx((w.x >> 4) & 0x0f0f0f0fu); hi.y = sx((w.y >> 4) & 0x0f0f0f0fu);
          hi.z = sx((w.z >> 4) & 0x0f0f0f0fu); hi.w = sx((w.w >> 4) & 0x0f0f0f0fu);
          uint8_t* dst = stA(s) + sl * (SB * 256) + blk * 256 + r * 16;
          *reinterpret_cast<uint4*>(dst) = lo;
          *reinterpret_cast<uint4*>(dst + 128) = hi;
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> visible to the MMA
        mbar_arrive(&full[s]);
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue
    const uint32_t q = warp & 3, g = (warp - EPI_WARP0) >> 2;  // TMEM lane quarter, column group of 16 tokens
    const uint32_t row = tile_m * TM + q * 32 + lane, rows_p = a.n_slabs * LLMI_SLAB;
    const bool row_ok = row < rows_p;
    const uint16_t* dsrc = reinterpret_cast<const uint16_t*>(a.d) + (size_t(row >> 3) * nb) * 8 + (row & 7);
    const uint32_t tok0 = tile_n * TN + g * 16;
    float acc[4][16];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < 16; ++k) acc[c][k] = 0.0f;
    for (uint32_t st = 0; st < n_st; ++st) {
      const uint32_t s = st % NSTAGE, b0 = b_begin + st * SB;
      float dw[SB];
#pragma unroll
      for (int ib = 0; ib < SB; ++ib)  // this row's block scales of the stage (global, read-only path)
        dw[ib] = (row_ok && b0 + ib < b_end) ? h2f(ldg_stream(dsrc + size_t(b0 + ib) * 8)) : 0.0f;
      mbar_wait(&full[s], (st / NSTAGE) & 1);  // the stage's activation scales are in shared memory
      const float* dxs = stD(s) + g * 16;
#pragma unroll
      for (int ib = 0; ib < SB; ++ib) {
        const uint32_t i = st * SB + ib;
        if (i < n_blk) {
          const uint32_t t = i % NTMEM;
          mbar_wait(&tfull[t], (i / NTMEM) & 1);
          tc_fence_after();
          int v[16];
          tmem_ld16(tmem_base + ((q * 32u) << 16) + t * TN + g * 16, v);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&tempty[t]);
          float dx[16];
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {
            const float4 d4 = *reinterpret_cast<const float4*>(dxs + ib * TN + k4 * 4);
            dx[k4 * 4] = d4.x; dx[k4 * 4 + 1] = d4.y; dx[k4 * 4 + 2] = d4.z; dx[k4 * 4 + 3] = d4.w;
          }
#pragma unroll
          for (int k = 0; k < 16; ++k) {
            const float fd = int_to_float(v[k]);
            if (IS_Q8) acc[ib & 3][k] = fmaf(__fmul_rn(fd, dw[ib]), dx[k], acc[ib & 3][k]);  // (int*dw)*dx, ops.cpp:820
            else acc[ib & 3][k] = fmaf(__fmul_rn(dw[ib], dx[k]), fd, acc[ib & 3][k]);         // ops.cpp:380-395
          }
          const uint32_t b = b0 + ib;
          if ((b & 15) == 15 || b == nb - 1) {  // end of a K-chunk: (s0+s1)+(s2+s3) -> part[chunk][token][row]
            const uint32_t j = b >> 4;
#pragma unroll
            for (int k = 0; k < 16; ++k) {
              const float p = (acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k]);
              if (row_ok && tok0 + k < a.n_tok) a.part[(size_t(j) * a.n_tok + tok0 + k) * rows_p + row] = p;
              acc[0][k] = acc[1][k] = acc[2][k] = acc[3][k] = 0.0f;
            }
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);  // done with the stage's scales
    }
  }
  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(NTMEM * TN))
                 : "memory");
}
