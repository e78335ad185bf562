// Dev tool (round-2 groundwork for the tensor-core prefill kernel): isolated latencies of the
// pieces whose chain sets that kernel's pace.  Build + run on the B200:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o /tmp/umma_micro tools/umma_micro.cu && /tmp/umma_micro
// One CTA on one SM unless said otherwise; all numbers are SM cycles (clock64).
//   A  tcgen05.ld 32x32b.x16 + wait::ld, back to back, 1 / 4 / 8 warps doing it concurrently
//   B  tcgen05.st x8 x2 + wait::st
//   C  tcgen05.mma kind::i8 M128 N32 K32 issued back to back by one thread (cycles per MMA until the last commit
//      is observed) and D  one MMA -> commit -> first successful mbarrier poll (round-trip latency)
//   E  mbarrier arrive by one warp -> observed by another (the hand-over cost between the kernel's roles)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t done;
  asm volatile(
      "{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n"
      : "=r"(done)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return done != 0;
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return uint64_t((addr & 0x3ffffu) >> 4) | (uint64_t(lbo >> 4) << 16) | (uint64_t(sbo >> 4) << 32) | (1ull << 46);
}
__device__ __forceinline__ void mma_i8(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc), "r"(0u)
      : "memory");
}
constexpr uint32_t IDESC = (2u << 4) | (1u << 7) | (1u << 10) | (uint32_t(32 >> 3) << 17) | (uint32_t(128 >> 4) << 24);

__global__ void __launch_bounds__(384, 1) micro(long long* out, int iters) {
  extern __shared__ __align__(1024) uint8_t smem[];  // A: 4 KB core matrices, B: 1 KB
  __shared__ __align__(8) uint64_t bar[4];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 8192; i += blockDim.x) smem[i] = uint8_t(i * 7);
  if (threadIdx.x == 0)
    for (int i = 0; i < 4; ++i) mbar_init(&bar[i], 1);
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(256u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tb = tmem_base_s, tcol = tb + ((uint32_t(warp & 3) * 32u) << 16);

  // ---- A: TMEM load latency with 1, 4, 8 warps loading concurrently (warps 0..n-1; lane quarter = warp % 4)
  for (int n_w = 1, slot = 0; n_w <= 8; n_w *= (n_w == 1 ? 4 : 2), ++slot) {
    __syncthreads();
    if (warp < n_w) {
      int v[16];
      const long long t0 = clock64();
      for (int i = 0; i < iters; ++i) {
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n"
            "tcgen05.wait::ld.sync.aligned;\n"
            : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
              "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
            : "r"(tcol + (i & 7) * 16)
            : "memory");
      }
      const long long t1 = clock64();
      if (threadIdx.x == 0) out[slot] = (t1 - t0) / iters;
      if (v[0] == 0x7fffffff) out[63] = v[3];  // keep the loads alive
    }
  }
  // ---- B: TMEM store (2 x x8) + wait::st
  __syncthreads();
  if (warp == 0) {
    const uint32_t m = 0x4B400000u;
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      asm volatile(
          "tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %1, %1, %1, %1, %1, %1, %1};\n"
          "tcgen05.st.sync.aligned.32x32b.x8.b32 [%2], {%1, %1, %1, %1, %1, %1, %1, %1};\n"
          "tcgen05.wait::st.sync.aligned;\n" ::"r"(tcol + (i & 7) * 16),
          "r"(m), "r"(tcol + (i & 7) * 16 + 8)
          : "memory");
    }
    const long long t1 = clock64();
    if (lane == 0) out[3] = (t1 - t0) / iters;
  }
  // ---- C: MMA issue cadence (one thread, `iters` MMAs, one commit at the end) and D: single MMA round trip
  __syncthreads();
  if (threadIdx.x == 32) {
    const uint64_t ad = smem_desc(smem_u32(smem), 128, 256), bd = smem_desc(smem_u32(smem) + 4096, 128, 256);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) mma_i8(tb + (i & 3) * 32, ad, bd, IDESC, 0);
    tc_commit(&bar[0]);
    const long long t_issue = clock64();
    while (!mbar_try(&bar[0], 0)) {}
    long long t1 = clock64();
    out[4] = (t_issue - t0) / iters;  // issue cost per MMA
    out[5] = (t1 - t0) / iters;       // throughput per MMA incl. drain
    t0 = clock64();
    mma_i8(tb, ad, bd, IDESC, 0);
    tc_commit(&bar[1]);
    while (!mbar_try(&bar[1], 0)) {}
    t1 = clock64();
    out[6] = t1 - t0;  // one MMA: issue -> commit observed
    // per-MMA commit (what the kernel does): cycles per (MMA + commit) when each is committed to its own barrier
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      mma_i8(tb + (i & 3) * 32, ad, bd, IDESC, 0);
      tc_commit(&bar[2]);
      while (!mbar_try(&bar[2], i & 1)) {}
    }
    t1 = clock64();
    out[7] = (t1 - t0) / iters;  // fully serialized MMA -> commit -> observe
  }
  // ---- E: mbarrier hand-over between two warps (ping-pong), cycles per one-way hand-over
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
  }
  __syncthreads();
  if (warp == 4 || warp == 5) {
    const long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      if (warp == 4) {
        if (lane == 0) mbar_arrive(&bar[0]);
        while (!mbar_try(&bar[1], i & 1)) {}
      } else {
        while (!mbar_try(&bar[0], i & 1)) {}
        if (lane == 0) mbar_arrive(&bar[1]);
      }
      __syncwarp();
    }
    const long long t1 = clock64();
    if (threadIdx.x == 128) out[8] = (t1 - t0) / (2 * iters);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tb), "r"(256u) : "memory");
}

int main() {
  long long* d;
  cudaMalloc(&d, 64 * sizeof(long long));
  cudaMemset(d, 0, 64 * sizeof(long long));
  cudaFuncSetAttribute(micro, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384);
  micro<<<1, 384, 16384>>>(d, 2000);
  const cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    return 1;
  }
  long long h[64];
  cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
  printf("A tcgen05.ld x16 + wait::ld, cycles per load: 1 warp %lld, 4 warps %lld, 8 warps %lld\n", h[0], h[1], h[2]);
  printf("B tcgen05.st 2 x x8 + wait::st: %lld\n", h[3]);
  printf("C tcgen05.mma i8 M128 N32 K32: issue %lld cycles each, %lld incl. drain; D single MMA -> commit observed %lld; "
         "serialized MMA+commit+observe %lld\n", h[4], h[5], h[6], h[7]);
  printf("E mbarrier hand-over between two warps: %lld cycles one way\n", h[8]);
  return 0;
}
