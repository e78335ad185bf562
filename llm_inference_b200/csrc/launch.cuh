// Kernel launch helper with programmatic dependent launch (PDL).
//
// A decode token is a chain of ~200 short dependent kernels; each costs ~2.3 us
// of launch + ramp when serialized.  With the programmatic-stream-serialization
// attribute the next kernel of the stream (or of the captured graph) is allowed
// to start while its predecessor is still running.  Every kernel therefore
//   * calls pdl_trigger() first thing (lets its own successor be scheduled), and
//   * calls pdl_wait() before it reads or writes anything a predecessor touches
//     (griddepcontrol.wait: returns when all prerequisite grids have completed
//     and their memory is visible).
// What runs before pdl_wait() overlaps the predecessor's tail: barrier setup
// and — in the mat-vec kernel — the first weight loads, which depend on nothing.
// Without the attribute (LLMI_NO_PDL=1, or a launch after a memcpy) both
// instructions are no-ops.
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <utility>

extern bool g_llmi_pdl;

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier + 1-D bulk async copy (cp.async.bulk -> UBLKCP), shared by the mat-vec and attention kernels
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// global -> shared, `bytes` a multiple of 16, both addresses 16-byte aligned; completion counted on the mbarrier
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
      "LLMI_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra LLMI_DONE;\n"
      "bra LLMI_WAIT;\n"
      "LLMI_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

template <typename... KArgs, typename... Args>
cudaError_t llmi_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_llmi_pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(std::forward<Args>(args))...);
}
