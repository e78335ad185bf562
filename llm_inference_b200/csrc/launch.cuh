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

#include "llmi_internal.h"  // LLTag, LLPeers

extern bool g_llmi_pdl;

__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---- mbarrier + 1-D bulk async copy (cp.async.bulk -> UBLKCP), shared by the mat-vec and attention kernels
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_inval(uint64_t* bar) {
  asm volatile("mbarrier.inval.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
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


// ---- dev only (-DLLMI_TIMELINE, tools/step_timeline.py): %globaltimer stamps of CTA 0 / thread 0 of every launch:
// entry, griddepcontrol.wait returned, last statement.  One record buffer per translation unit (no relocatable device
// code in this build); the host merges them by time.
#ifdef LLMI_TIMELINE
struct TlRec {
  unsigned long long t[10];  // 0 entry, 1 wait returned, 2 last statement, 3.. kernel-specific phase boundaries
  unsigned kind, ctas;
};
static __device__ TlRec g_tl[8192];
static __device__ unsigned g_tl_n;
__device__ __forceinline__ unsigned long long tl_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ int tl_enter(unsigned kind) {
  if (blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) return -1;
  const unsigned s = atomicAdd(&g_tl_n, 1u) & 8191u;
  g_tl[s].kind = kind;
  g_tl[s].ctas = gridDim.x * gridDim.y * gridDim.z;
  g_tl[s].t[0] = tl_now();
  for (int i = 1; i < 10; ++i) g_tl[s].t[i] = 0;
  return int(s);
}
__device__ __forceinline__ void tl_mark(int slot, int i) {
  if (slot >= 0) g_tl[slot].t[i] = tl_now();
}
#define TL_ENTER(kind) const int tl_slot = tl_enter(kind)
#define TL_MARK(i) tl_mark(tl_slot, i)
#define TL_EXPORT(name)                                                                        \
  extern "C" int name(void* out, unsigned* n, int reset) {                                    \
    if (out) cudaMemcpyFromSymbol(out, g_tl, sizeof(g_tl));                                    \
    if (n) cudaMemcpyFromSymbol(n, g_tl_n, 4);                                                 \
    if (reset) {                                                                               \
      const unsigned z = 0;                                                                    \
      cudaMemcpyToSymbol(g_tl_n, &z, 4);                                                       \
    }                                                                                          \
    return 0;                                                                                  \
  }
#else
#define TL_ENTER(kind) do { } while (0)
#define TL_MARK(i) do { } while (0)
#define TL_EXPORT(name)
#endif

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

// ---------------------------------------------------------------------------
// Flagged 8-byte exchange between the GPUs of a row-sharded model.
// A vector that one rank produces and every rank consumes travels as {value bits,
// tag} pairs: the producing mat-vec writes each of its rows with ONE 64-bit
// store straight into the consumer's buffer on every peer (NVLink peer memory),
// and the consuming kernel spins on the element until the tag matches.  A
// naturally aligned 64-bit access is single-copy atomic, so data and flag
// arrive together: no fence, no ticket, no collective kernel — the transfer is
// the mat-vec's own epilogue and its latency is one NVLink store.
// tag = *epoch * mul + add: `epoch` is a device counter every rank bumps once
// per token step (the ranks run the same steps), `add` identifies the exchange
// inside the step, so a buffer reused every layer never shows a stale match.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t ll_tag(const LLTag& t) {
  if (!t.epoch) return t.add;  // tag given by value (persistent decode kernel: the host counts the steps)
  uint32_t e;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(e) : "l"(t.epoch) : "memory");
  return e * t.mul + t.add;
}
__device__ __forceinline__ void ll_store(uint2* p, uint32_t bits, uint32_t tag) {
  asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(bits), "r"(tag) : "memory");
}
// Spins until element *p carries `tag`; gives up after ~4 s (a peer that never
// arrives must not hang the GPU) and reports through *err.  The spin is out of line: it is the slow path, and the
// persistent decode kernel has dozens of wait sites whose code must stay small (instruction cache).
static __device__ __noinline__ uint32_t ll_wait_slow(const uint2* p, uint32_t tag, uint32_t* err) {
  uint32_t v, f;
  unsigned long long limit_ns = 4000000000ull;  // err[1]: the limit in milliseconds (LLMI_EXCHANGE_TIMEOUT_MS at load)
  if (err) {  // an earlier wait already timed out: the run is lost, drain quickly
    uint32_t dead, ms;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(dead) : "l"(err) : "memory");
    if (dead) return 0;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(ms) : "l"(err + 1) : "memory");
    if (ms) limit_ns = (unsigned long long)ms * 1000000ull;
  }
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (uint32_t spins = 1;; ++spins) {
    asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
    if (f == tag) return v;
    if ((spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if (t1 - t0 > limit_ns) {
        if (err) *err = 1;
        return 0;
      }
    }
    __nanosleep(32);  // thousands of threads may be polling the same L2 lines: leave room for the producer's store
  }
}
// (the reader always polls its OWN GPU's buffer: a gpu-scope strong load, served by L2, is enough — also for words a
// peer wrote over NVLink, which land in this GPU's L2 — and, unlike ld.volatile, several of them pipeline)
__device__ __forceinline__ uint32_t ll_wait(const uint2* p, uint32_t tag, uint32_t* err) {
  uint32_t v, f;
  asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v), "=r"(f) : "l"(p) : "memory");
  if (f == tag) return v;
  return ll_wait_slow(p, tag, err);
}
// N flagged words at once: every load is issued before the first tag is examined (one L2 round trip when the data
// is there); stragglers take the spinning wait.
template <int N>
__device__ __forceinline__ void ll_wait_many(const uint2* const (&p)[N], const bool (&ok)[N], uint32_t tag, uint32_t* err,
                                             uint32_t (&v)[N]) {
  uint32_t f[N];
#pragma unroll
  for (int k = 0; k < N; ++k) {
    v[k] = 0;
    f[k] = tag;
    if (ok[k]) asm volatile("ld.relaxed.gpu.global.v2.u32 {%0, %1}, [%2];" : "=r"(v[k]), "=r"(f[k]) : "l"(p[k]) : "memory");
  }
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (f[k] != tag) v[k] = ll_wait_slow(p[k], tag, err);
}
__device__ __forceinline__ float ll_waitf(const uint2* p, uint32_t tag, uint32_t* err) {
  return __uint_as_float(ll_wait(p, tag, err));
}
