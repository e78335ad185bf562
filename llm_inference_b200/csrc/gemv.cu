// Decode mat-vec kernels for sm_100a — the replacement for the compute_range
// bodies + thread-pool fan-out of ops.cpp:188-931.
//
// Shape of every kernel:
//   * work unit = one SLAB of 8 output rows; lane = 8*sub + row, so one 128-bit
//     load per lane covers 4 K-units x 8 rows = 512 contiguous bytes of the
//     quant plane (weights are read exactly once, straight into registers,
//     L1 no-allocate; no shared-memory round trip for weights);
//   * KSPLIT warps cooperate on a slab, interleaved along K (the reference's
//     row partition, ops.cpp:439-448, becomes grid partitioning; K-split adds
//     parallelism for the short/wide matrices and is reduced in a fixed order,
//     so results are deterministic and independent of the grid);
//   * the quantized activation vector (K bytes + scales) is staged into shared
//     memory once per CTA with one bulk async copy (cp.async.bulk -> UBLKCP)
//     completing on an mbarrier, overlapped with the first weight loads;
//   * integer block dots are __dp4a, bit-exact with the reference's per-block
//     sums; the fp32 scale product and accumulation follow the reference's
//     formulas (summation ORDER differs: that is the documented 1e-5 bound).
#include <cuda_fp16.h>

#include "llmi_internal.h"

namespace {

// ------------------------------------------------------------------ helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ uint2 ldg_stream(const uint2* p) {
  uint2 r;
  asm("ld.global.nc.L1::no_allocate.v2.u32 {%0,%1}, [%2];" : "=r"(r.x), "=r"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ uint32_t ldg_stream(const uint32_t* p) {
  uint32_t r;
  asm("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ uint16_t ldg_stream(const uint16_t* p) {
  uint16_t r;
  asm("ld.global.nc.L1::no_allocate.u16 %0, [%1];" : "=h"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 1-D bulk async copy global -> shared, completion counted on the mbarrier.
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

struct GemvArgs {
  const uint8_t* q;
  const uint8_t* d;
  const uint8_t* x;
  const uint8_t* act;
  uint32_t act_bytes;
  float* out;  // already offset to this handle's first row
  uint32_t n_local, n_slabs, nb, n_cols;
};

// ------------------------------------------------ integer block dot products
// (shared by the GEMV kernels and the debug dump, so the dumped integers are
// the ones the GEMV consumed)

// Q4_0 block (ops.cpp:373-396): byte j of w holds element j (low nibble) and
// element j+16 (high nibble); xa = q8 elements 0..15, xb = 16..31.
__device__ __forceinline__ int q4_0_block_dot(const uint4 w, const int4 xa, const int4 xb, const int xsum) {
  int dp = 0;
  dp = __dp4a(int(w.x & 0x0f0f0f0fu), xa.x, dp);
  dp = __dp4a(int(w.y & 0x0f0f0f0fu), xa.y, dp);
  dp = __dp4a(int(w.z & 0x0f0f0f0fu), xa.z, dp);
  dp = __dp4a(int(w.w & 0x0f0f0f0fu), xa.w, dp);
  dp = __dp4a(int((w.x >> 4) & 0x0f0f0f0fu), xb.x, dp);
  dp = __dp4a(int((w.y >> 4) & 0x0f0f0f0fu), xb.y, dp);
  dp = __dp4a(int((w.z >> 4) & 0x0f0f0f0fu), xb.z, dp);
  dp = __dp4a(int((w.w >> 4) & 0x0f0f0f0fu), xb.w, dp);
  return dp - 8 * xsum;  // sum (nib-8)*q == sum nib*q - 8*sum q
}

// Q8_0 block (ops.cpp:816-819)
__device__ __forceinline__ int q8_0_block_dot(const uint4 w0, const uint4 w1, const int4 xa, const int4 xb) {
  int dp = 0;
  dp = __dp4a(int(w0.x), xa.x, dp);
  dp = __dp4a(int(w0.y), xa.y, dp);
  dp = __dp4a(int(w0.z), xa.z, dp);
  dp = __dp4a(int(w0.w), xa.w, dp);
  dp = __dp4a(int(w1.x), xb.x, dp);
  dp = __dp4a(int(w1.y), xb.y, dp);
  dp = __dp4a(int(w1.z), xb.z, dp);
  dp = __dp4a(int(w1.w), xb.w, dp);
  return dp;
}

__device__ __forceinline__ int dp16(const uint4 w, const int4 x) {
  int dp = __dp4a(int(w.x), x.x, 0);
  dp = __dp4a(int(w.y), x.y, dp);
  dp = __dp4a(int(w.z), x.z, dp);
  return __dp4a(int(w.w), x.w, dp);
}
__device__ __forceinline__ uint4 and4(const uint4 a, const uint32_t m) {
  return make_uint4(a.x & m, a.y & m, a.z & m, a.w & m);
}
__device__ __forceinline__ uint4 shr4(const uint4 a, const int s) {
  return make_uint4(a.x >> s, a.y >> s, a.z >> s, a.w >> s);
}
__device__ __forceinline__ uint4 shl4(const uint4 a, const int s) {
  return make_uint4(a.x << s, a.y << s, a.z << s, a.w << s);
}
__device__ __forceinline__ uint4 or4(const uint4 a, const uint4 b) {
  return make_uint4(a.x | b.x, a.y | b.y, a.z | b.z, a.w | b.w);
}

// Q4_K 64-element pair c of a super-block (ops.cpp:662-688): qa/qb = qs bytes
// 32c..32c+15 / +16..+31; low nibbles pair with q8[64c..64c+31], high nibbles
// with q8[64c+32..64c+63].  Nibbles are unsigned.
__device__ __forceinline__ void q4_k_pair_dots(const uint4 qa, const uint4 qb, const int4 x0, const int4 x1,
                                               const int4 x2, const int4 x3, int& sum_lo, int& sum_hi) {
  sum_lo = dp16(and4(qa, 0x0f0f0f0fu), x0) + dp16(and4(qb, 0x0f0f0f0fu), x1);
  sum_hi = dp16(and4(shr4(qa, 4), 0x0f0f0f0fu), x2) + dp16(and4(shr4(qb, 4), 0x0f0f0f0fu), x3);
}

// Q6_K (ops.cpp:744-767), the 64 elements {l, l+32, l+64, l+96 : l in
// [16hh,16hh+16)} of 128-half n: qa = ql[l], qb = ql[l+32], qh = qh[l];
// s0..s6 = int8 scales sc[hh], sc[hh+2], sc[hh+4], sc[hh+6]; b0..b6 the q8
// group sums of the four 16-element groups.  (q-32)*x summed = dp(q,x) - 32*bsum.
__device__ __forceinline__ int q6_k_part(const uint4 qa, const uint4 qb, const uint4 qh, const int4 x0, const int4 x1,
                                         const int4 x2, const int4 x3, const int s0, const int s2, const int s4,
                                         const int s6, const int b0, const int b2, const int b4, const int b6) {
  const uint4 l1 = or4(and4(qa, 0x0f0f0f0fu), and4(shl4(qh, 4), 0x30303030u));
  const uint4 l2 = or4(and4(qb, 0x0f0f0f0fu), and4(shl4(qh, 2), 0x30303030u));
  const uint4 l3 = or4(and4(shr4(qa, 4), 0x0f0f0f0fu), and4(qh, 0x30303030u));
  const uint4 l4 = or4(and4(shr4(qb, 4), 0x0f0f0f0fu), and4(shr4(qh, 2), 0x30303030u));
  return s0 * (dp16(l1, x0) - 32 * b0) + s2 * (dp16(l2, x1) - 32 * b2) + s4 * (dp16(l3, x2) - 32 * b4) +
         s6 * (dp16(l4, x3) - 32 * b6);
}

// 6-bit scale / min of sub-block j from the 12 packed bytes (ops.cpp:633-641);
// the bytes are words y,z,w of the 16-byte header item {d, dmin, scales[12]}.
__device__ __forceinline__ uint32_t hdr_byte(const uint4 h, const int i) {
  const uint32_t w = i < 4 ? h.y : (i < 8 ? h.z : h.w);
  return (w >> (8 * (i & 3))) & 0xffu;
}
__device__ __forceinline__ void q4_k_scale_min(const uint4 h, const int j, int& sc, int& mn) {
  if (j < 4) {
    sc = int(hdr_byte(h, j) & 63u);
    mn = int(hdr_byte(h, j + 4) & 63u);
  } else {
    sc = int((hdr_byte(h, j + 4) & 0x0fu) | ((hdr_byte(h, j - 4) >> 6) << 4));
    mn = int((hdr_byte(h, j + 4) >> 4) | ((hdr_byte(h, j) >> 6) << 4));
  }
}

__device__ __forceinline__ int sbyte(const uint2 v, const int i) {  // signed byte i (0..7) of 8 bytes
  const uint32_t w = i < 4 ? v.x : v.y;
  return int(int8_t((w >> (8 * (i & 3))) & 0xffu));
}

// ------------------------------------------------------------ format bodies
// Each body returns the lane's partial sum over the K-units it owns:
//   units u = kw + ks*j, lane's sub-position = lane>>3, row in slab = lane&7.

template <int UNROLL>
struct BodyQ4_0 {
  __device__ static float run(const GemvArgs& a, const uint8_t* sm, uint64_t* bar, uint32_t slab, int kw, int ks,
                              int lane) {
    const int r = lane & 7, sub = lane >> 3;
    const uint32_t nb = a.nb;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (size_t)slab * nb * 8 + r;
    const uint16_t* d = reinterpret_cast<const uint16_t*>(a.d) + (size_t)slab * nb * 8 + r;
    const int4* xs = reinterpret_cast<const int4*>(sm);
    const uint32_t* meta = reinterpret_cast<const uint32_t*>(sm + a.n_cols);
    float acc = 0.0f;
    bool waited = false;
    const uint32_t step = 4u * ks;
    for (uint32_t b0 = 4u * kw + sub; b0 < nb; b0 += step * UNROLL) {
      uint4 w[UNROLL];
      uint16_t dw[UNROLL];
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t b = b0 + i * step;
        if (b < nb) {
          w[i] = ldg_stream(q + (size_t)b * 8);
          dw[i] = ldg_stream(d + (size_t)b * 8);
        }
      }
      if (!waited) {
        mbar_wait(bar, 0);
        waited = true;
      }
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t b = b0 + i * step;
        if (b < nb) {
          const int4 xa = xs[2 * b], xb = xs[2 * b + 1];
          const uint32_t m = meta[b];
          const int dot = q4_0_block_dot(w[i], xa, xb, int(int16_t(m >> 16)));
          acc = fmaf(h2f(dw[i]) * h2f(uint16_t(m & 0xffffu)), float(dot), acc);  // ops.cpp:380-395
        }
      }
    }
    if (!waited) mbar_wait(bar, 0);
    return acc;
  }
};

template <int UNROLL>
struct BodyQ8_0 {
  __device__ static float run(const GemvArgs& a, const uint8_t* sm, uint64_t* bar, uint32_t slab, int kw, int ks,
                              int lane) {
    const int r = lane & 7, sub = lane >> 3;
    const uint32_t nb = a.nb;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (size_t)slab * nb * 16 + r;
    const uint16_t* d = reinterpret_cast<const uint16_t*>(a.d) + (size_t)slab * nb * 8 + r;
    const int4* xs = reinterpret_cast<const int4*>(sm);
    const uint32_t* meta = reinterpret_cast<const uint32_t*>(sm + a.n_cols);
    float acc = 0.0f;
    bool waited = false;
    const uint32_t step = 4u * ks;
    for (uint32_t b0 = 4u * kw + sub; b0 < nb; b0 += step * UNROLL) {
      uint4 w0[UNROLL], w1[UNROLL];
      uint16_t dw[UNROLL];
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t b = b0 + i * step;
        if (b < nb) {
          w0[i] = ldg_stream(q + (size_t)b * 16);
          w1[i] = ldg_stream(q + (size_t)b * 16 + 8);
          dw[i] = ldg_stream(d + (size_t)b * 8);
        }
      }
      if (!waited) {
        mbar_wait(bar, 0);
        waited = true;
      }
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t b = b0 + i * step;
        if (b < nb) {
          const int dot = q8_0_block_dot(w0[i], w1[i], xs[2 * b], xs[2 * b + 1]);
          const float dx = h2f(uint16_t(meta[b] & 0xffffu));
          acc = fmaf(float(dot) * h2f(dw[i]), dx, acc);  // (int*dw)*dx, ops.cpp:820
        }
      }
    }
    if (!waited) mbar_wait(bar, 0);
    return acc;
  }
};

// Q5_0 keeps fp32 activations (ops.cpp:856-878): no integer dot.
template <int UNROLL>
struct BodyQ5_0 {
  __device__ static float run(const GemvArgs& a, const uint8_t* sm, uint64_t* bar, uint32_t slab, int kw, int ks,
                              int lane) {
    const int r = lane & 7, sub = lane >> 3;
    const uint32_t nb = a.nb;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (size_t)slab * nb * 8 + r;
    const uint32_t* qhp = reinterpret_cast<const uint32_t*>(a.x) + (size_t)slab * nb * 8 + r;
    const uint16_t* d = reinterpret_cast<const uint16_t*>(a.d) + (size_t)slab * nb * 8 + r;
    const float4* xs = reinterpret_cast<const float4*>(sm);
    float acc0 = 0.0f, acc1 = 0.0f;
    bool waited = false;
    const uint32_t step = 4u * ks;
    for (uint32_t b0 = 4u * kw + sub; b0 < nb; b0 += step * UNROLL) {
      uint4 w[UNROLL];
      uint32_t qh[UNROLL];
      uint16_t dw[UNROLL];
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t b = b0 + i * step;
        if (b < nb) {
          w[i] = ldg_stream(q + (size_t)b * 8);
          qh[i] = ldg_stream(qhp + (size_t)b * 8);
          dw[i] = ldg_stream(d + (size_t)b * 8);
        }
      }
      if (!waited) {
        mbar_wait(bar, 0);
        waited = true;
      }
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t b = b0 + i * step;
        if (b < nb) {
          const float dv = h2f(dw[i]);
          const uint32_t ws[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const float4 xl = xs[b * 8 + g], xh = xs[b * 8 + 4 + g];
            const float xlv[4] = {xl.x, xl.y, xl.z, xl.w}, xhv[4] = {xh.x, xh.y, xh.z, xh.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const int idx = 4 * g + e;
              const uint32_t byte = (ws[g] >> (8 * e)) & 0xffu;
              const int q0 = int((byte & 0x0fu) | (((qh[i] >> idx) & 1u) << 4));
              const int q1 = int((byte >> 4) | (((qh[i] >> (idx + 16)) & 1u) << 4));
              acc0 = fmaf(dv * float(q0 - 16), xlv[e], acc0);  // ops.cpp:873
              acc1 = fmaf(dv * float(q1 - 16), xhv[e], acc1);  // ops.cpp:874
            }
          }
        }
      }
    }
    if (!waited) mbar_wait(bar, 0);
    return acc0 + acc1;
  }
};

// K-quants: one super-block of 256 per warp iteration; sub = 64-element chunk.
template <int UNROLL>
struct BodyQ4_K {
  __device__ static float run(const GemvArgs& a, const uint8_t* sm, uint64_t* bar, uint32_t slab, int kw, int ks,
                              int lane) {
    const int r = lane & 7, c = lane >> 3;
    const uint32_t nb = a.nb;
    const uint4* hdr = reinterpret_cast<const uint4*>(a.x) + (size_t)slab * nb * 8 + r;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + ((size_t)slab * nb * 4 + c) * 16 + r;
    const int4* xs = reinterpret_cast<const int4*>(sm);
    const uint2* bs = reinterpret_cast<const uint2*>(sm + a.n_cols);
    const float* xd = reinterpret_cast<const float*>(sm + a.n_cols + a.n_cols / 8);
    float acc = 0.0f;
    bool waited = false;
    for (uint32_t s0 = kw; s0 < nb; s0 += ks * UNROLL) {
      uint4 h[UNROLL], qa[UNROLL], qb[UNROLL];
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t sb = s0 + i * ks;
        if (sb < nb) {
          h[i] = ldg_stream(hdr + (size_t)sb * 8);
          qa[i] = ldg_stream(q + (size_t)sb * 64);
          qb[i] = ldg_stream(q + (size_t)sb * 64 + 8);
        }
      }
      if (!waited) {
        mbar_wait(bar, 0);
        waited = true;
      }
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t sb = s0 + i * ks;
        if (sb < nb) {
          const int4* xp = xs + sb * 16 + c * 4;
          int sum_lo, sum_hi;
          q4_k_pair_dots(qa[i], qb[i], xp[0], xp[1], xp[2], xp[3], sum_lo, sum_hi);
          const uint2 b4 = bs[sb * 4 + c];  // bsums 4c..4c+3 of this super-block
          const int bs_lo = int(int16_t(b4.x & 0xffffu)) + int(int16_t(b4.x >> 16));
          const int bs_hi = int(int16_t(b4.y & 0xffffu)) + int(int16_t(b4.y >> 16));
          int sc1, m1, sc2, m2;
          q4_k_scale_min(h[i], 2 * c, sc1, m1);
          q4_k_scale_min(h[i], 2 * c + 1, sc2, m2);
          const float dx = xd[sb];
          const float dd = h2f(uint16_t(h[i].x & 0xffffu)) * dx;  // ops.cpp:654
          const float mm = h2f(uint16_t(h[i].x >> 16)) * dx;      // ops.cpp:655
          acc += fmaf(dd * float(sc1), float(sum_lo), -((mm * float(m1)) * float(bs_lo)));  // ops.cpp:671
          acc += fmaf(dd * float(sc2), float(sum_hi), -((mm * float(m2)) * float(bs_hi)));  // ops.cpp:682
        }
      }
    }
    if (!waited) mbar_wait(bar, 0);
    return acc;
  }
};

template <int UNROLL>
struct BodyQ6_K {
  __device__ static float run(const GemvArgs& a, const uint8_t* sm, uint64_t* bar, uint32_t slab, int kw, int ks,
                              int lane) {
    const int r = lane & 7, sub = lane >> 3, n = sub >> 1, hh = sub & 1;
    const uint32_t nb = a.nb;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + ((size_t)slab * nb * 12 + sub) * 8 + r;
    const uint2* scp = reinterpret_cast<const uint2*>(a.x) + ((size_t)slab * nb * 8 + r) * 2 + n;
    const uint16_t* d = reinterpret_cast<const uint16_t*>(a.d) + (size_t)slab * nb * 8 + r;
    const int4* xs = reinterpret_cast<const int4*>(sm);
    const int16_t* bs = reinterpret_cast<const int16_t*>(sm + a.n_cols);
    const float* xd = reinterpret_cast<const float*>(sm + a.n_cols + a.n_cols / 8);
    float acc = 0.0f;
    bool waited = false;
    for (uint32_t s0 = kw; s0 < nb; s0 += ks * UNROLL) {
      uint4 qa[UNROLL], qb[UNROLL], qh[UNROLL];
      uint2 sc[UNROLL];
      uint16_t dw[UNROLL];
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t sb = s0 + i * ks;
        if (sb < nb) {
          qa[i] = ldg_stream(q + (size_t)sb * 96);
          qb[i] = ldg_stream(q + (size_t)sb * 96 + 32);
          qh[i] = ldg_stream(q + (size_t)sb * 96 + 64);
          sc[i] = ldg_stream(scp + (size_t)sb * 16);
          dw[i] = ldg_stream(d + (size_t)sb * 8);
        }
      }
      if (!waited) {
        mbar_wait(bar, 0);
        waited = true;
      }
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t sb = s0 + i * ks;
        if (sb < nb) {
          const uint32_t g0 = sb * 16 + n * 8 + hh;  // 16-element group of x0
          const int part = q6_k_part(qa[i], qb[i], qh[i], xs[g0], xs[g0 + 2], xs[g0 + 4], xs[g0 + 6], sbyte(sc[i], hh),
                                     sbyte(sc[i], hh + 2), sbyte(sc[i], hh + 4), sbyte(sc[i], hh + 6), bs[g0],
                                     bs[g0 + 2], bs[g0 + 4], bs[g0 + 6]);
          acc = fmaf(h2f(dw[i]) * xd[sb], float(part), acc);  // ops.cpp:738,762
        }
      }
    }
    if (!waited) mbar_wait(bar, 0);
    return acc;
  }
};

// F16 (ops.cpp:541-586): x rounded to f16 first, products exact in fp32.
// BF16 (ops.cpp:908-916): x stays fp32.  unit = 4 chunks of 8 elements.
template <int UNROLL, bool IS_BF16>
struct BodyHalf {
  __device__ static float run(const GemvArgs& a, const uint8_t* sm, uint64_t* bar, uint32_t slab, int kw, int ks,
                              int lane) {
    const int r = lane & 7, sub = lane >> 3;
    const uint32_t nb = a.nb;  // chunks of 8
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (size_t)slab * nb * 8 + r;
    float acc0 = 0.0f, acc1 = 0.0f, acc2 = 0.0f, acc3 = 0.0f;
    bool waited = false;
    const uint32_t step = 4u * ks;
    for (uint32_t c0 = 4u * kw + sub; c0 < nb; c0 += step * UNROLL) {
      uint4 w[UNROLL];
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t c = c0 + i * step;
        if (c < nb) w[i] = ldg_stream(q + (size_t)c * 8);
      }
      if (!waited) {
        mbar_wait(bar, 0);
        waited = true;
      }
#pragma unroll
      for (int i = 0; i < UNROLL; ++i) {
        const uint32_t c = c0 + i * step;
        if (c < nb) {
          if (IS_BF16) {
            const float4 xa = reinterpret_cast<const float4*>(sm)[2 * c];
            const float4 xb = reinterpret_cast<const float4*>(sm)[2 * c + 1];
            acc0 = fmaf(__uint_as_float(w[i].x << 16), xa.x, acc0);
            acc1 = fmaf(__uint_as_float(w[i].x & 0xffff0000u), xa.y, acc1);
            acc2 = fmaf(__uint_as_float(w[i].y << 16), xa.z, acc2);
            acc3 = fmaf(__uint_as_float(w[i].y & 0xffff0000u), xa.w, acc3);
            acc0 = fmaf(__uint_as_float(w[i].z << 16), xb.x, acc0);
            acc1 = fmaf(__uint_as_float(w[i].z & 0xffff0000u), xb.y, acc1);
            acc2 = fmaf(__uint_as_float(w[i].w << 16), xb.z, acc2);
            acc3 = fmaf(__uint_as_float(w[i].w & 0xffff0000u), xb.w, acc3);
          } else {
            const uint4 xv = reinterpret_cast<const uint4*>(sm)[c];
            const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&w[i].x));
            const float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&w[i].y));
            const float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&w[i].z));
            const float2 w3 = __half22float2(*reinterpret_cast<const __half2*>(&w[i].w));
            const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&xv.x));
            const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&xv.y));
            const float2 x2 = __half22float2(*reinterpret_cast<const __half2*>(&xv.z));
            const float2 x3 = __half22float2(*reinterpret_cast<const __half2*>(&xv.w));
            acc0 = fmaf(w0.x, x0.x, acc0);
            acc1 = fmaf(w0.y, x0.y, acc1);
            acc2 = fmaf(w1.x, x1.x, acc2);
            acc3 = fmaf(w1.y, x1.y, acc3);
            acc0 = fmaf(w2.x, x2.x, acc0);
            acc1 = fmaf(w2.y, x2.y, acc1);
            acc2 = fmaf(w3.x, x3.x, acc2);
            acc3 = fmaf(w3.y, x3.y, acc3);
          }
        }
      }
    }
    if (!waited) mbar_wait(bar, 0);
    return (acc0 + acc1) + (acc2 + acc3);
  }
};

// ------------------------------------------------------------ kernel skeleton
template <int KSPLIT>
struct Cfg {
  static constexpr int WARPS = KSPLIT >= 4 ? KSPLIT : 4;
  static constexpr int SLABS_PER_CTA = WARPS / KSPLIT;
};

template <class Body, int KSPLIT>
__global__ void __launch_bounds__(Cfg<KSPLIT>::WARPS * 32) gemv_slab_kernel(const GemvArgs a) {
  extern __shared__ __align__(128) uint8_t sm_act[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ float red[Cfg<KSPLIT>::WARPS][LLMI_SLAB];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) mbar_init(&bar, 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    mbar_expect_tx(&bar, a.act_bytes);
    bulk_g2s(sm_act, a.act, a.act_bytes, &bar);
  }
  const uint32_t slab = blockIdx.x * Cfg<KSPLIT>::SLABS_PER_CTA + warp / KSPLIT;
  const int kw = warp % KSPLIT;
  float acc = 0.0f;
  if (slab < a.n_slabs) {
    acc = Body::run(a, sm_act, &bar, slab, kw, KSPLIT, lane);
  } else {
    mbar_wait(&bar, 0);  // never exit with the bulk copy into our smem in flight
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  const uint32_t row = slab * LLMI_SLAB + lane;
  if (KSPLIT == 1) {
    if (lane < LLMI_SLAB && slab < a.n_slabs && row < a.n_local) a.out[row] = acc;
  } else {
    if (lane < LLMI_SLAB) red[warp][lane] = acc;
    __syncthreads();
    if (kw == 0 && lane < LLMI_SLAB && slab < a.n_slabs && row < a.n_local) {
      float s = red[warp][lane];
#pragma unroll
      for (int i = 1; i < KSPLIT; ++i) s += red[warp + i][lane];  // fixed order: deterministic
      a.out[row] = s;
    }
  }
}

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

template <class Body, int KSPLIT>
cudaError_t launch_one(const GemvArgs& a, cudaStream_t s) {
  using C = Cfg<KSPLIT>;
  const uint32_t grid = (a.n_slabs + C::SLABS_PER_CTA - 1) / C::SLABS_PER_CTA;
  gemv_slab_kernel<Body, KSPLIT><<<grid, C::WARPS * 32, a.act_bytes, s>>>(a);
  return cudaGetLastError();
}

template <class Body>
cudaError_t launch_ks(const GemvArgs& a, int ks, cudaStream_t s) {
  switch (ks) {
    case 1: return launch_one<Body, 1>(a, s);
    case 2: return launch_one<Body, 2>(a, s);
    case 4: return launch_one<Body, 4>(a, s);
    case 8: return launch_one<Body, 8>(a, s);
    default: return launch_one<Body, 16>(a, s);
  }
}

template <class Body, int KSPLIT>
cudaError_t optin_one() {
  return cudaFuncSetAttribute(gemv_slab_kernel<Body, KSPLIT>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                              MAX_DYN_SMEM);
}
template <class Body>
cudaError_t optin_all() {
  cudaError_t e;
  if ((e = optin_one<Body, 1>()) != cudaSuccess) return e;
  if ((e = optin_one<Body, 2>()) != cudaSuccess) return e;
  if ((e = optin_one<Body, 4>()) != cudaSuccess) return e;
  if ((e = optin_one<Body, 8>()) != cudaSuccess) return e;
  return optin_one<Body, 16>();
}

using Q4_0 = BodyQ4_0<4>;
using Q8_0 = BodyQ8_0<2>;
using Q5_0 = BodyQ5_0<2>;
using Q4_K = BodyQ4_K<2>;
using Q6_K = BodyQ6_K<2>;
using F16 = BodyHalf<4, false>;
using BF16 = BodyHalf<4, true>;

int g_sm_count = 148;

// K-split heuristic: enough warps to cover the chip (~16 per SM) while every
// warp still owns >= 2 K-units.
int pick_ksplit(const llmi_weight_s& w) {
  const uint64_t units = (w.type == LLMI_Q4_K || w.type == LLMI_Q6_K) ? w.nb : (w.nb + 3) / 4;
  const uint64_t want = uint64_t(g_sm_count) * 16;
  int ks = 1;
  while (ks < 16 && w.n_slabs * ks < want && units / (ks * 2) >= 2) ks *= 2;
  return ks;
}

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

cudaError_t llmi_gemv_init() {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  e = cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
  if (e != cudaSuccess) return e;
  if ((e = optin_all<Q4_0>()) != cudaSuccess) return e;
  if ((e = optin_all<Q8_0>()) != cudaSuccess) return e;
  if ((e = optin_all<Q5_0>()) != cudaSuccess) return e;
  if ((e = optin_all<Q4_K>()) != cudaSuccess) return e;
  if ((e = optin_all<Q6_K>()) != cudaSuccess) return e;
  if ((e = optin_all<F16>()) != cudaSuccess) return e;
  return optin_all<BF16>();
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
  return g;
}

cudaError_t llmi_launch_gemv(const llmi_weight_s& w, const llmi_act_s& a, float* out, int ksplit_override,
                             cudaStream_t s) {
  if (w.n_slabs == 0) return cudaSuccess;
  const GemvArgs g = make_args(w, a, out);
  if (g.act_bytes > (uint32_t)MAX_DYN_SMEM) return cudaErrorInvalidValue;
  const int ks = ksplit_override > 0 ? ksplit_override : pick_ksplit(w);
  switch (w.type) {
    case LLMI_Q4_0: return launch_ks<Q4_0>(g, ks, s);
    case LLMI_Q8_0: return launch_ks<Q8_0>(g, ks, s);
    case LLMI_Q5_0: return launch_ks<Q5_0>(g, ks, s);
    case LLMI_Q4_K: return launch_ks<Q4_K>(g, ks, s);
    case LLMI_Q6_K: return launch_ks<Q6_K>(g, ks, s);
    case LLMI_F16: return launch_ks<F16>(g, ks, s);
    case LLMI_BF16: return launch_ks<BF16>(g, ks, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t llmi_launch_block_dots(const llmi_weight_s& w, const llmi_act_s& a, int32_t* dots_dev, cudaStream_t s) {
  const GemvArgs g = make_args(w, a, nullptr);
  const uint64_t per_row = w.type == LLMI_Q6_K ? w.nb * 2 : (w.type == LLMI_Q4_K ? w.nb * 8 : w.nb);
  const uint64_t total = w.n_local * per_row;
  if (!total) return cudaSuccess;
  block_dots_kernel<<<unsigned((total + 255) / 256), 256, 0, s>>>(g, w.type, dots_dev);
  return cudaGetLastError();
}
