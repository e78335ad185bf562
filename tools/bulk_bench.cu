// Dev tool: how fast does one SM pull HBM through cp.async.bulk (UBLKCP), by copy size, copies in flight and issuers?
// Each CTA streams its own contiguous range of a buffer (> L2) through a shared-memory ring; nothing is computed.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bulk_bench tools/bulk_bench.cu && /tmp/bulk_bench
// Columns: copy bytes, slots per issuer, issuers (warps) per CTA, CTAs per SM -> GB/s.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "W:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DN;\n"
      "bra W;\n"
      "DN:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// every warp is an issuer with its own ring of `depth` slots of `bytes`; a CTA covers [cta*per, (cta+1)*per)
__global__ void bulk_kernel(const uint8_t* src, size_t per_cta, uint32_t bytes, int depth, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ __align__(8) uint64_t bars[64];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  if (threadIdx.x < W * depth) mbar_init(&bars[threadIdx.x], 1);
  __syncthreads();
  uint8_t* ring = smem + size_t(warp) * depth * bytes;
  uint64_t* wbar = bars + warp * depth;
  const uint8_t* base = src + size_t(blockIdx.x) * per_cta;
  const uint32_t n = uint32_t(per_cta / bytes);  // copies of this CTA; warp w takes w, w+W, ...
  if (lane == 0)
    for (int s = 0; s < depth; ++s) {
      const uint32_t i = warp + s * W;
      if (i < n) {
        mbar_expect_tx(&wbar[s], bytes);
        bulk_g2s(ring + size_t(s) * bytes, base + size_t(i) * bytes, bytes, &wbar[s]);
      }
    }
  unsigned acc = 0;
  uint32_t k = 0;
  for (uint32_t i = warp; i < n; i += W, ++k) {
    const int s = int(k % depth);
    mbar_wait(&wbar[s], (k / depth) & 1);
    acc += *reinterpret_cast<const unsigned*>(ring + size_t(s) * bytes + lane * 4);
    __syncwarp();
    const uint32_t i2 = i + depth * W;
    if (lane == 0 && i2 < n) {
      mbar_expect_tx(&wbar[s], bytes);
      bulk_g2s(ring + size_t(s) * bytes, base + size_t(i2) * bytes, bytes, &wbar[s]);
    }
  }
  if (acc == 0x12345678u) *sink = acc;
}

// reference: plain 128-bit loads, grid-stride, 4 loads in flight per thread
__global__ void ldg_kernel(const uint4* src, size_t n16, unsigned* sink) {
  unsigned acc = 0;
  const size_t stride = size_t(gridDim.x) * blockDim.x;
  size_t i = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    uint4 a, b, c, d;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(src + i));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(src + i + stride));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(c.x), "=r"(c.y), "=r"(c.z), "=r"(c.w) : "l"(src + i + 2 * stride));
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w) : "l"(src + i + 3 * stride));
    acc += a.x ^ b.y ^ c.z ^ d.w;
  }
  if (acc == 0x12345678u) *sink = acc;
}

// the same ring filled by 16-byte cp.async (LDGSTS) issued by all 32 lanes, one commit group per slot
template <int DEPTH>
__global__ void ldgsts_kernel(const uint8_t* src, size_t per_cta, uint32_t bytes, unsigned* sink) {
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, W = blockDim.x >> 5;
  uint8_t* ring = smem + size_t(warp) * DEPTH * bytes;
  const uint8_t* base = src + size_t(blockIdx.x) * per_cta;
  const uint32_t n = uint32_t(per_cta / bytes);
  auto fill = [&](uint32_t i, int s) {
    if (i < n)
      for (uint32_t off = lane * 16; off < bytes; off += 512)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(ring + size_t(s) * bytes + off)),
                     "l"(base + size_t(i) * bytes + off)
                     : "memory");
    asm volatile("cp.async.commit_group;" ::: "memory");
  };
#pragma unroll
  for (int s = 0; s < DEPTH; ++s) fill(warp + s * W, s);
  unsigned acc = 0;
  uint32_t k = 0;
  for (uint32_t i = warp; i < n; i += W, ++k) {
    const int s = int(k % DEPTH);
    asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    __syncwarp();
    acc += *reinterpret_cast<const unsigned*>(ring + size_t(s) * bytes + lane * 4);
    __syncwarp();
    fill(i + DEPTH * W, s);
  }
  if (acc == 0x12345678u) *sink = acc;
}
template <int DEPTH>
void run_ldgsts(const uint8_t* buf, size_t total, int sms, unsigned* sink, cudaEvent_t e0, cudaEvent_t e1) {
  cudaFuncSetAttribute(ldgsts_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (uint32_t bytes : {2048u, 2304u, 4352u})
    for (int w : {4, 8, 16})
      for (int c : {1, 2}) {
        const size_t smem = size_t(w) * DEPTH * bytes;
        if (smem * c > 200 * 1024) continue;
        const int ctas = sms * c;
        size_t per = total / ctas / (size_t(bytes) * w) * (size_t(bytes) * w);
        ldgsts_kernel<DEPTH><<<ctas, w * 32, smem>>>(buf, per, bytes, sink);
        cudaEventRecord(e0);
        ldgsts_kernel<DEPTH><<<ctas, w * 32, smem>>>(buf, per, bytes, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        printf("ldgsts %6u %6d %6d %4d %10.0f %8.0f\n", bytes, DEPTH, w, c, smem * c / 1024.0, double(per) * ctas / ms * 1e-6);
      }
}

int main() {
  const size_t total = size_t(1) << 30;  // 1 GiB > L2
  uint8_t* buf;
  unsigned* sink;
  cudaMalloc(&buf, total);
  cudaMalloc(&sink, 4);
  cudaMemset(buf, 1, total);
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  cudaFuncSetAttribute(bulk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  {
    ldg_kernel<<<sms * 8, 256>>>((const uint4*)buf, total / 16, sink);
    cudaEventRecord(e0);
    ldg_kernel<<<sms * 8, 256>>>((const uint4*)buf, total / 16, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("ldg128 x4 in flight, %d CTAs x 256: %.0f GB/s\n", sms * 8, total / ms * 1e-6);
  }
  run_ldgsts<2>(buf, total, sms, sink, e0, e1);
  run_ldgsts<3>(buf, total, sms, sink, e0, e1);
  run_ldgsts<4>(buf, total, sms, sink, e0, e1);
  if (getenv("BULK_ONLY_LDGSTS")) return 0;
  const uint32_t sizes[] = {2048, 4096, 8192, 16384, 32768};
  const int warps[] = {1, 2, 4, 8};
  const int cps[] = {1, 2};
  printf("%8s %6s %6s %4s %10s %8s\n", "bytes", "depth", "warps", "cps", "inflightKB", "GB/s");
  for (uint32_t bytes : sizes)
    for (int w : warps)
      for (int c : cps)
        for (int depth : {2, 4, 8}) {
          const size_t smem = size_t(w) * depth * bytes;
          if (smem * c > 200 * 1024 || smem > 200 * 1024 || w * depth > 64) continue;
          const int ctas = sms * c;
          size_t per = total / ctas / (size_t(bytes) * w) * (size_t(bytes) * w);
          bulk_kernel<<<ctas, w * 32, smem>>>(buf, per, bytes, depth, sink);
          cudaEventRecord(e0);
          bulk_kernel<<<ctas, w * 32, smem>>>(buf, per, bytes, depth, sink);
          cudaEventRecord(e1);
          cudaEventSynchronize(e1);
          float ms;
          cudaEventElapsedTime(&ms, e0, e1);
          if (cudaGetLastError() != cudaSuccess) {
            printf("launch failed\n");
            continue;
          }
          printf("%8u %6d %6d %4d %10.0f %8.0f\n", bytes, depth, w, c, smem * c / 1024.0, double(per) * ctas / ms * 1e-6);
        }
  return 0;
}
