// Prefill on the 5th-generation tensor cores, EXACT: tcgen05.mma kind::i8 for Q4_0 / Q8_0 weights.
// Included by gemv.cu (uses its GemvArgs / GemvBatch and the part[] / toklane_reduce_kernel contract).
//
// The reference has no batched matmul (its forward loops tokens around mat_vec_mul, model.cpp:714-960), and its
// arithmetic is: exact integer dot of one 32-element block (int8 activation quants x int4/int8 weight quants),
// ONE fp32 rounding per block (scale product, fma), blocks accumulated in a fixed order.  A dequantize-to-bf16
// GEMM cannot reproduce that; an int8 MMA with K = 32 can: one tcgen05.mma per quant block gives the int32
// block dots of a 128-row x 32-token tile in tensor memory, bit-identical to the dp4a dots, and the fp32 part
// (float(dot), dw*dx, fma into the chain of the canonical summation order) is the epilogue.  Results are
// bit-identical to the one-token kernel.  The per-block fp32 epilogue on the CUDA cores is the bound by
// construction: 3 ops per output per block (vs ~14 for the dp4a kernel), i.e. the tensor pipe can be ~15 % busy at
// best; measured 2.4 % (DESIGN.md 4.6 says why and what is next).
//
// CTA = 128 weight rows (16 slabs) x 32 tokens x a range of K-chunks, 12 warps, warp-specialized:
//   warp 0      producer: ONE 2-D TMA tensor copy per stage for the weights (quant plane = tensor [slab][K run], box =
//               16 slabs x 8 blocks; the Q8_0 slab layout IS the UMMA K-major core-matrix layout, 8 rows x 16 bytes),
//               one bulk copy each for the token tile's activation quants (pre-packed into core-matrix order by
//               umma_pack_act_kernel) and fp32 scales; its idle lanes prefetch the weight lines 8 stages ahead into L2
//   warp 1      tensor-memory allocation; one thread issues tcgen05.mma (M128 N32 K32, accumulating onto the "armed"
//               cells, see the epilogue) per block into a ring of 8 TMEM slots and commits to mbarriers
//   warps 2-3   Q4_0 only: unpack the nibbles of a stage to signed int8 core matrices (generic -> async proxy fence)
//   warps 4-11  epilogue: tcgen05.ld the block dots (thread = row, 16 token columns), four chain accumulators
//               per (row, token) = the four sub-lanes of gemv_slab_kernel, chunk partial (s0+s1)+(s2+s3) stored
//               to part[chunk][token][row]; toklane_reduce_kernel adds the chunk partials left to right.
// Shared memory: 5 (Q8_0) / 3 (Q4_0) operand stages of 8 blocks (A 32 KB int8 + B 8 KB, + 16 KB packed nibbles for
// Q4_0) released by the MMA commit alone, and a deeper ring of 8 scale stages (1 KB) the epilogue releases.
#pragma once

namespace umma {

constexpr int TM = 128, TN = 32;          // tile: weight rows x tokens
constexpr int SB = 8;                     // quant blocks per shared-memory stage
constexpr int NTMEM = 8;                  // TMEM ring depth (x TN columns) = blocks per stage: ring slot = block in stage
constexpr int NDS = 8;                    // ring of activation-scale stages (1 KB each), deeper than the operand ring
static_assert(NTMEM == SB, "the TMEM slot of a block is its index in the stage (static barrier addresses)");
constexpr int WARPS = 12, EPI_WARP0 = 4;   // 8 epilogue warps: 4 TMEM lane quarters x 2 groups of EC token columns
constexpr int EC = 16;                    // token columns per epilogue thread (measured: 8 warps x 16 columns beat 16 x 8)
constexpr uint32_t A_BYTES = TM * SB * 32, B_BYTES = TN * SB * 32, D_BYTES = TN * SB * 4, P_BYTES = TM * SB * 16;
// The operand ring is what hides the refill latency (18 bulk copies per stage, ~3 us from issue to landed): a stage
// is released by the MMA commit alone — the scales the epilogue still needs live in their own, deeper ring — and
// the ring is as deep as shared memory allows: 5 stages of A+B for Q8_0, 3 of A+B+packed nibbles for Q4_0.
template <bool IS_Q8>
struct Cfg {
  static constexpr int NSTAGE = IS_Q8 ? 5 : 3;
  static constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES + (IS_Q8 ? 0u : P_BYTES);
  static constexpr size_t SMEM_BYTES = size_t(NSTAGE) * STAGE_BYTES + size_t(NDS) * D_BYTES + 1024;
};

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
// D[tmem] += A[smem] * B[smem]^T, int8 x int8 -> int32, M128 N32 K32 (D is pre-armed, see the epilogue)
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(1u), "r"(0u)
      : "memory");
}
// cute::UMMA::InstrDescriptor: c_format S32 (2) @4, a/b format INT8 (1) @7/@10, K-major both, N>>3 @17, M>>4 @24
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(TN >> 3) << 17) | (uint32_t(TM >> 4) << 24);

// 2-D tiled TMA load (coordinates in elements of the tensor map: c0 innermost), completion on the mbarrier
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tm, uint32_t c0, uint32_t c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, int (&v)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
               : "r"(taddr)
               : "memory");
}
// the registers of the pending load are operands, so nothing that reads them can move above the wait
__device__ __forceinline__ void tmem_wait_ld(int (&v)[8]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7])
               :
               : "memory");
}
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_ld8_issue(uint32_t taddr, int (&v)[16]) {  // 16-column overload
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld(int (&v)[16]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(v[0]), "+r"(v[1]), "+r"(v[2]), "+r"(v[3]), "+r"(v[4]), "+r"(v[5]), "+r"(v[6]), "+r"(v[7]), "+r"(v[8]),
                 "+r"(v[9]), "+r"(v[10]), "+r"(v[11]), "+r"(v[12]), "+r"(v[13]), "+r"(v[14]), "+r"(v[15])
               :
               : "memory");
}
constexpr uint32_t MAGIC = 0x4B400000u;  // bits of 12582912.0f = 1.5 * 2^23: bits(MAGIC + x) = 12582912 + x for |x| < 2^22
// this thread's row, 8 columns <- MAGIC (mg: eight registers holding MAGIC, kept live by the caller — the store
// wants a register vector and re-materializing it costs eight moves per block)
__device__ __forceinline__ void tmem_arm8(uint32_t taddr, const uint32_t (&mg)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(taddr), "r"(mg[0]),
               "r"(mg[1]), "r"(mg[2]), "r"(mg[3]), "r"(mg[4]), "r"(mg[5]), "r"(mg[6]), "r"(mg[7])
               : "memory");
}
// this thread's row, EC columns <- MAGIC
__device__ __forceinline__ void tmem_arm(uint32_t taddr, const uint32_t (&mg)[8]) {
#pragma unroll
  for (int c = 0; c < EC; c += 8) tmem_arm8(taddr + c, mg);
}

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

#ifdef LLMI_UMMA_TIMING  // dev only (tools/umma_timeline.py): clock64 stamps of CTA (0,0,0), one row per role
__device__ long long g_umma_stamp[10][160];
#define UMMA_STAMP(role, idx)                                                                                   \
  do {                                                                                                          \
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && (idx) < 160) g_umma_stamp[role][idx] = clock64(); \
  } while (0)
#else
#define UMMA_STAMP(role, idx) do { } while (0)
#endif

// grid = (row tiles of all matrices of the batch, token tiles, K-chunk groups of `nj` chunks)
// DIRECT: the CTA walks ALL K-chunks of its tile in order (grid.z == 1, true for every large matrix): the chunk
// partials are added left to right in registers — the canonical order — and each (token, row) is stored once.  The
// partials never travel through memory (they were 5.5x the call's algorithmic traffic, profiles/r01_ncu_umma_q8_final.csv)
// and the reduce launch disappears.  !DIRECT: K-chunk groups across grid.z, partials to part[], toklane_reduce_kernel.
template <bool IS_Q8, bool DIRECT>
__global__ void __launch_bounds__(umma::WARPS * 32, 1)
gemm_umma_kernel(const GemvBatch batch, const uint4* __restrict__ bq, const float* __restrict__ bd, uint32_t nj,
                 const __grid_constant__ CUtensorMap tm0, const __grid_constant__ CUtensorMap tm1,
                 const __grid_constant__ CUtensorMap tm2) {
  using namespace umma;
  constexpr int NSTAGE = Cfg<IS_Q8>::NSTAGE;
  constexpr uint32_t STAGE_BYTES = Cfg<IS_Q8>::STAGE_BYTES;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t full[NSTAGE], pk_full[NSTAGE], empty[NSTAGE], dfull[NDS], dempty[NDS], tfull[NTMEM],
      tempty[NTMEM];
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
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  auto stA = [&](uint32_t s) { return smem + size_t(s) * STAGE_BYTES; };                      // [slab 16][blk 8][h 2][r 8][16]
  auto stB = [&](uint32_t s) { return smem + size_t(s) * STAGE_BYTES + A_BYTES; };            // [blk 8][n/8 4][h 2][n%8 8][16]
  auto stP = [&](uint32_t s) { return smem + size_t(s) * STAGE_BYTES + A_BYTES + B_BYTES; };  // [slab 16][blk 8][r 8][16] nibbles
  auto stD = [&](uint32_t d) { return reinterpret_cast<float*>(smem + size_t(NSTAGE) * STAGE_BYTES + size_t(d) * D_BYTES); };  // [blk 8][32]

  if (threadIdx.x == 0) {
    for (int s = 0; s < NSTAGE; ++s) {
      mbar_init(&full[s], IS_Q8 ? 1 : 1 + 64);   // producer's expect_tx (+ the 64 unpack threads)
      mbar_init(&pk_full[s], 1);
      mbar_init(&empty[s], 1);                    // MMA commit: the stage's operands have been read
    }
    for (int d = 0; d < NDS; ++d) {
      mbar_init(&dfull[d], 1);
      mbar_init(&dempty[d], WARPS - EPI_WARP0);  // the epilogue warps are done with the stage's scales
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
    // One TMA tensor copy per stage for the weights: the quant plane is a 2-D tensor [slab][K run] and the stage is
    // the box 16 slabs x (8 blocks' bytes), which lands as [slab][blk][...] — the operand layout.  (16 separate
    // bulk copies of 2 KB ran at ~6 B/clk per SM: the copy engine works through them one after the other; the
    // clock64 timeline of tools/umma_timeline.py showed a stage landing every ~7000 cycles whatever else changed.)
    // The other 31 lanes pull the weight lines of the stage PF_AHEAD stages further down into L2 through the
    // load/store unit (prefetch.global.L2), so the copy engine's requests hit L2.  (Measured: Q4_0 32 -> 37 TMAC/s;
    // the cadence of stage arrivals in tools/umma_timeline.py — one per ~6000 cycles — did not change, see DESIGN.md.)
    constexpr uint32_t PF_AHEAD = NSTAGE + 3, BLK_BYTES = IS_Q8 ? 256u : 128u;
    const CUtensorMap* tm = mi == 0 ? &tm0 : (mi == 1 ? &tm1 : &tm2);
    auto prefetch_stage = [&](uint32_t st) {  // lanes 1..31
      if (st >= n_st) return;
      const uint32_t b0 = b_begin + st * SB, nbs = min(uint32_t(SB), b_end - b0);
      const uint32_t lines_per_slab = nbs * (BLK_BYTES / 128u), n_lines = n_sl * lines_per_slab;
      for (uint32_t ln = lane - 1; ln < n_lines; ln += 31) {
        const uint32_t sl = ln / lines_per_slab, o = ln % lines_per_slab;
        const uint8_t* p = a.q + (size_t(slab0 + sl) * nb + b0) * BLK_BYTES + size_t(o) * 128u;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
      }
    };
    if (lane != 0)
      for (uint32_t st = 0; st < min(PF_AHEAD, n_st); ++st) prefetch_stage(st);
    for (uint32_t st = 0; st < n_st; ++st) {
      if (lane == 0) {
        const uint32_t s = st % NSTAGE, ph = (st / NSTAGE) & 1;
        const uint32_t ds = st % NDS;
        if (st >= NSTAGE) mbar_wait(&empty[s], ph ^ 1);
        if (st >= NDS) mbar_wait(&dempty[ds], ((st / NDS) & 1) ^ 1);
        UMMA_STAMP(0, st);
        const uint32_t b0 = b_begin + st * SB, nbs = min(uint32_t(SB), b_end - b0);
        const uint32_t tb = (tile_n * nb + b0);
        // the box is always whole (rows / blocks past the end arrive as zeros and are never used)
        if (IS_Q8) {
          mbar_expect_tx(&full[s], A_BYTES + nbs * 1024u);
          tma_load_2d(stA(s), tm, b0 * 32, slab0, &full[s]);
        } else {
          mbar_expect_tx(&pk_full[s], P_BYTES);
          tma_load_2d(stP(s), tm, b0 * 16, slab0, &pk_full[s]);
          mbar_expect_tx(&full[s], nbs * 1024u);
        }
        bulk_g2s(stB(s), bq + size_t(tb) * 64, nbs * 1024u, &full[s]);
        mbar_expect_tx(&dfull[ds], nbs * 128u);
        bulk_g2s(stD(ds), bd + size_t(tb) * TN, nbs * 128u, &dfull[ds]);
      } else {
        prefetch_stage(st + PF_AHEAD);
      }
      __syncwarp();  // the prefetching lanes keep pace with the ring
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer
    // ONE thread runs this loop (tcgen05.mma / tcgen05.commit are single-thread instructions): the loop is a
    // serial chain of uniform-datapath instructions, so it is kept to a wait, a descriptor add, the MMA and a
    // commit per block — with the whole warp looping, electing and re-converging it was the kernel's bound
    // (~580 cycles per block in the ncu source view, profiles/r01_notes.md).
    if (lane == 0) {
      const uint64_t a_hi = (uint64_t((SB * 256) >> 4) << 32) | (uint64_t(128 >> 4) << 16) | (1ull << 46);
      const uint64_t b_hi = (uint64_t(256 >> 4) << 32) | (uint64_t(128 >> 4) << 16) | (1ull << 46);
      const uint32_t a_lo0 = (smem_u32(stA(0)) & 0x3ffffu) >> 4, b_lo0 = (smem_u32(stB(0)) & 0x3ffffu) >> 4;
      uint32_t s = 0, sph = 0;
      for (uint32_t st = 0; st < n_st; ++st) {
        mbar_wait(&full[s], sph);
        UMMA_STAMP(1, st);
        tc_fence_after();
        const uint32_t nbs = min(uint32_t(SB), n_blk - st * SB);
        const uint64_t ad0 = a_hi | uint64_t(a_lo0 + s * (STAGE_BYTES >> 4));
        const uint64_t bd0 = b_hi | uint64_t(b_lo0 + s * (STAGE_BYTES >> 4));
#pragma unroll
        for (int ib = 0; ib < SB; ++ib) {
          if (uint32_t(ib) < nbs) {
            mbar_wait(&tempty[ib], st & 1);  // armed (completion 0 = the initial arming) and drained
            tc_fence_after();
            mma_i8(tmem_base + ib * TN, ad0 + uint64_t(ib * (256 >> 4)), bd0 + uint64_t(ib * (1024 >> 4)), IDESC);
            tc_commit(&tfull[ib]);
            UMMA_STAMP(2, st * SB + ib);
          }
        }
        tc_commit(&empty[s]);  // arrives when the stage's MMAs have read their operands
        if (++s == NSTAGE) {
          s = 0;
          sph ^= 1;
        }
      }
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
    // Every TMEM stage is kept "armed" with the integer 0x4B400000 in every cell and the MMA ACCUMULATES onto it:
    // the bits of cell + dot are the float 12582912 + dot, so float(dot) is one FADD (exact for |dot| < 2^22)
    // instead of an integer add + FADD or a quarter-rate I2F.  The thread that read a cell re-arms it.
    const uint32_t q = warp & 3, g = (warp - EPI_WARP0) >> 2;  // TMEM lane quarter, group of EC token columns
    const uint32_t row = tile_m * TM + q * 32 + lane, rows_p = a.n_slabs * LLMI_SLAB;
    const bool row_ok = row < rows_p;
    const uint16_t* dsrc = reinterpret_cast<const uint16_t*>(a.d) + (size_t(row >> 3) * nb) * 8 + (row & 7);
    const uint32_t tok0 = tile_n * TN + g * EC;
    const uint32_t tcol = tmem_base + ((q * 32u) << 16) + g * EC;
    uint32_t mg[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) asm volatile("mov.u32 %0, %1;" : "=r"(mg[k]) : "r"(MAGIC));  // opaque: stays in registers
#pragma unroll
    for (int t = 0; t < NTMEM; ++t) tmem_arm(tcol + t * TN, mg);
    tmem_wait_st();
    tc_fence_before();
    __syncwarp();
    if (lane == 0)
      for (int t = 0; t < NTMEM; ++t) mbar_arrive(&tempty[t]);
    float acc[4][EC];
    float total[DIRECT ? EC : 1];  // DIRECT: running sum of the chunk partials of (row, token column k)
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int k = 0; k < EC; ++k) acc[c][k] = 0.0f;
    // The loop body is ONE group of four blocks (one block per chain), not unrolled further: the fully unrolled
    // stage (8 blocks, each with its own copy of the chunk-end stores) was ~5000 instructions = 80 KB of straight-line
    // code that no instruction cache level kept, and the whole epilogue ran at ~5 cycles per instruction
    // (profiles/r01_notes.md).  Per group: this row's four block scales were fetched one group ahead.
    uint16_t dwn[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) dwn[k] = (row_ok && b_begin + k < b_end) ? ldg_stream(dsrc + size_t(b_begin + k) * 8) : uint16_t(0);
    int va[EC], vb[EC];  // block dots of the even / odd block in flight
    if (n_blk) {
      mbar_wait(&tfull[0], 0);
      tc_fence_after();
      tmem_ld8_issue(tcol, va);
    }
    const uint32_t n_grp = (n_blk + 3) / 4;
#pragma unroll 1
    for (uint32_t gi = 0; gi < n_grp; ++gi) {
      const uint32_t st = gi >> 1, half = gi & 1, ds = st % NDS, bg = b_begin + gi * 4;  // bg: first block of the group
      float dw[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) dw[k] = h2f(dwn[k]);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const uint32_t bn = bg + 4 + k;
        dwn[k] = (row_ok && bn < b_end) ? ldg_stream(dsrc + size_t(bn) * 8) : uint16_t(0);
      }
      if (half == 0) {
        if (warp == EPI_WARP0 && lane == 0) UMMA_STAMP(8, st);
        mbar_wait(&dfull[ds], (st / NDS) & 1);
        if (warp == EPI_WARP0 && lane == 0) UMMA_STAMP(9, st);
      }
      const uint32_t dxs = smem_u32(stD(ds)) + g * (EC * 4) + half * (4 * TN * 4);
#pragma unroll
      for (int k4b = 0; k4b < 4; ++k4b) {  // block of the group = chain index
        const uint32_t i = gi * 4 + k4b;
        if (i < n_blk) {
          int(&v)[EC] = (k4b & 1) ? vb : va;
          int(&vn)[EC] = (k4b & 1) ? va : vb;
          const uint32_t t = half * 4 + k4b, tn = (t + 1) % NTMEM;  // TMEM slots of this block and the next
          tmem_wait_ld(v);  // v = block i
          if (warp == EPI_WARP0 && lane == 0) UMMA_STAMP(3, i);
          if (i + 1 < n_blk) {  // next block's dots fly while this block is folded
            mbar_wait(&tfull[tn], (tn == 0 ? st + 1 : st) & 1);
            tc_fence_after();
            if (warp == EPI_WARP0 && lane == 0) UMMA_STAMP(4, i);
            tmem_ld8_issue(tcol + tn * TN, vn);
          }
          // release the PREVIOUS block's slot now: its re-arming store was issued an iteration ago, so the wait is
          // free — waiting right after the store put the tensor-memory store latency into every warp's per-block
          // chain (each warp walks every block, so that chain, not the instruction count, set the pace)
          if (i > 0) {
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tempty[(t + NTMEM - 1) % NTMEM]);
          }
          tmem_arm(tcol + t * TN, mg);
          float dx[EC];
#pragma unroll
          for (int k4 = 0; k4 < EC / 4; ++k4) {
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                         : "=f"(dx[k4 * 4]), "=f"(dx[k4 * 4 + 1]), "=f"(dx[k4 * 4 + 2]), "=f"(dx[k4 * 4 + 3])
                         : "r"(dxs + k4b * (TN * 4) + k4 * 16));
          }
#pragma unroll
          for (int k = 0; k < EC; ++k) {
            const float fd = __int_as_float(v[k]) - 12582912.0f;
            if (IS_Q8) acc[k4b][k] = fmaf(__fmul_rn(fd, dw[k4b]), dx[k], acc[k4b][k]);  // (int*dw)*dx, ops.cpp:820
            else acc[k4b][k] = fmaf(__fmul_rn(dw[k4b], dx[k]), fd, acc[k4b][k]);         // ops.cpp:380-395
          }
          if (warp == EPI_WARP0 && lane == 0) UMMA_STAMP(5, i);
        }
      }
      // K-chunks are 16 blocks and groups 4, so a chunk ends with a group: the one holding block 15 (mod 16) or the
      // last block of the row.  (s0+s1)+(s2+s3) -> part[chunk][token][row]
      const uint32_t b_last = min(bg + 3, b_end - 1);
      if ((b_last & 15) == 15 || b_last == nb - 1) {
        const uint32_t j = b_last >> 4;
#pragma unroll
        for (int k = 0; k < EC; ++k) {
          const float p = (acc[0][k] + acc[1][k]) + (acc[2][k] + acc[3][k]);
          if constexpr (DIRECT) {
            total[k] = j ? total[k] + p : p;  // p0, then + p1, + p2 ...: the order of toklane_reduce_kernel
            if (b_last == nb - 1 && row < a.n_local && tok0 + k < a.n_tok)
              a.out[size_t(tok0 + k) * a.out_stride + row] = total[k];
          } else {
            if (row_ok && tok0 + k < a.n_tok) a.part[(size_t(j) * a.n_tok + tok0 + k) * rows_p + row] = p;
          }
          acc[0][k] = acc[1][k] = acc[2][k] = acc[3][k] = 0.0f;
        }
      }
      if (half == 1 || gi == n_grp - 1) {
        __syncwarp();
        if (lane == 0) mbar_arrive(&dempty[ds]);  // done with the stage's scales
      }
    }
    tmem_wait_st();  // the last re-arming store (nobody consumes it) must have landed before the dealloc
  }
  // ------------------------------------------------------------------ teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(uint32_t(NTMEM * TN))
                 : "memory");
}
