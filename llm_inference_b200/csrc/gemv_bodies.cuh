// Format bodies of the mat-vec kernels: the per-format weight
// fragments, their global loads and the fold of one K-unit into a lane's partial
// sum.  Shared by the one-launch-per-op kernels (gemv.cu) and the persistent
// decode kernel (mega.cu) so that both produce the same bits.
#pragma once

#include <cuda_fp16.h>
#include <stdint.h>

#include "llmi_internal.h"

namespace {

// ------------------------------------------------------------------ helpers
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
#ifndef LLMI_H2F_DEFINED
#define LLMI_H2F_DEFINED
__device__ __forceinline__ float h2f(uint16_t h) { return __half2float(__ushort_as_half(h)); }
#endif


// ------------------------------------------------ integer block dot products
// (shared by the GEMV kernels and the debug dump, so the dumped integers are
// the ones the GEMV consumed)

// Q4_0 block (ops.cpp:373-396): byte j of w holds element j (low nibble) and
// element j+16 (high nibble); xa = q8 elements 0..15, xb = 16..31.
// The high nibbles stay in place: (w & 0xf0f0f0f0) as UNSIGNED bytes dotted with the signed q8 bytes is 16 x the
// high-half dot, an exact multiple of 16 (no shift per word; the mat-vec kernels are issue-bound).
__device__ __forceinline__ int dp4a_us(uint32_t a, int b, int c) {  // unsigned x signed: no such __dp4a overload
  int d;
  asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
  return d;
}
__device__ __forceinline__ int q4_0_block_dot(const uint4 w, const int4 xa, const int4 xb, const int xsum) {
  int lo = __dp4a(int(w.x & 0x0f0f0f0fu), xa.x, 0);
  lo = __dp4a(int(w.y & 0x0f0f0f0fu), xa.y, lo);
  lo = __dp4a(int(w.z & 0x0f0f0f0fu), xa.z, lo);
  lo = __dp4a(int(w.w & 0x0f0f0f0fu), xa.w, lo);
  int hi = dp4a_us(w.x & 0xf0f0f0f0u, xb.x, 0);
  hi = dp4a_us(w.y & 0xf0f0f0f0u, xb.y, hi);
  hi = dp4a_us(w.z & 0xf0f0f0f0u, xb.z, hi);
  hi = dp4a_us(w.w & 0xf0f0f0f0u, xb.w, hi);
  return lo + (hi >> 4) - 8 * xsum;  // sum (nib-8)*q == sum nib*q - 8*sum q
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
// A Body describes one format: Frag = the registers one lane loads for one
// K-unit, load() issues the global loads, compute() folds the unit into the
// lane's partial sum.  C = units per K-chunk (512 elements; 128 for F16/BF16).
// lane = 8*sub + r: r = row in slab, sub = position inside the unit.

struct BodyQ4_0 {
  static constexpr int C = 4;
  struct Frag {
    uint4 w;
    uint16_t d;
  };
  __device__ __forceinline__ static void load(Frag& f, const GemvArgs& a, uint32_t slab, uint32_t u, int r, int sub) {
    const uint32_t b = 4 * u + sub;
    if (b < a.nb) {
      const size_t i = ((size_t)slab * a.nb + b) * 8 + r;
      f.w = ldg_stream(reinterpret_cast<const uint4*>(a.q) + i);
      f.d = ldg_stream(reinterpret_cast<const uint16_t*>(a.d) + i);
    }
  }
  __device__ __forceinline__ static float compute(const Frag& f, const GemvArgs& a, const uint8_t* sm, uint32_t u, int sub,
                                  float acc) {
    const uint32_t b = 4 * u + sub;
    if (b < a.nb) {
      const int4* xs = reinterpret_cast<const int4*>(sm);
      const uint32_t m = reinterpret_cast<const uint32_t*>(sm + a.n_cols)[b];
      const int dot = q4_0_block_dot(f.w, xs[2 * b], xs[2 * b + 1], int(int16_t(m >> 16)));
      acc = fmaf(h2f(f.d) * h2f(uint16_t(m & 0xffffu)), float(dot), acc);  // ops.cpp:380-395
    }
    return acc;
  }
};

struct BodyQ8_0 {
  static constexpr int C = 4;
  struct Frag {
    uint4 w0, w1;
    uint16_t d;
  };
  __device__ __forceinline__ static void load(Frag& f, const GemvArgs& a, uint32_t slab, uint32_t u, int r, int sub) {
    const uint32_t b = 4 * u + sub;
    if (b < a.nb) {
      const size_t i = ((size_t)slab * a.nb + b) * 8 + r;
      const uint4* q = reinterpret_cast<const uint4*>(a.q) + ((size_t)slab * a.nb + b) * 16 + r;
      f.w0 = ldg_stream(q);
      f.w1 = ldg_stream(q + 8);
      f.d = ldg_stream(reinterpret_cast<const uint16_t*>(a.d) + i);
    }
  }
  __device__ __forceinline__ static float compute(const Frag& f, const GemvArgs& a, const uint8_t* sm, uint32_t u, int sub,
                                  float acc) {
    const uint32_t b = 4 * u + sub;
    if (b < a.nb) {
      const int4* xs = reinterpret_cast<const int4*>(sm);
      const int dot = q8_0_block_dot(f.w0, f.w1, xs[2 * b], xs[2 * b + 1]);
      const float dx = h2f(uint16_t(reinterpret_cast<const uint32_t*>(sm + a.n_cols)[b] & 0xffffu));
      acc = fmaf(float(dot) * h2f(f.d), dx, acc);  // (int*dw)*dx, ops.cpp:820
    }
    return acc;
  }
};

// Q5_0 keeps fp32 activations (ops.cpp:856-878): no integer dot.
struct BodyQ5_0 {
  static constexpr int C = 4;
  struct Frag {
    uint4 w;
    uint32_t qh;
    uint16_t d;
  };
  __device__ __forceinline__ static void load(Frag& f, const GemvArgs& a, uint32_t slab, uint32_t u, int r, int sub) {
    const uint32_t b = 4 * u + sub;
    if (b < a.nb) {
      const size_t i = ((size_t)slab * a.nb + b) * 8 + r;
      f.w = ldg_stream(reinterpret_cast<const uint4*>(a.q) + i);
      f.qh = ldg_stream(reinterpret_cast<const uint32_t*>(a.x) + i);
      f.d = ldg_stream(reinterpret_cast<const uint16_t*>(a.d) + i);
    }
  }
  __device__ __forceinline__ static float compute(const Frag& f, const GemvArgs& a, const uint8_t* sm, uint32_t u, int sub,
                                  float acc) {
    const uint32_t b = 4 * u + sub;
    if (b < a.nb) {
      const float4* xs = reinterpret_cast<const float4*>(sm) + (size_t)b * 8;
      const float dv = h2f(f.d);
      const uint32_t ws[4] = {f.w.x, f.w.y, f.w.z, f.w.w};
      float acc1 = 0.0f;
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const float4 xl = xs[g], xh = xs[4 + g];
        const float xlv[4] = {xl.x, xl.y, xl.z, xl.w}, xhv[4] = {xh.x, xh.y, xh.z, xh.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int idx = 4 * g + e;
          const uint32_t byte = (ws[g] >> (8 * e)) & 0xffu;
          const int q0 = int((byte & 0x0fu) | (((f.qh >> idx) & 1u) << 4));
          const int q1 = int((byte >> 4) | (((f.qh >> (idx + 16)) & 1u) << 4));
          acc = fmaf(dv * float(q0 - 16), xlv[e], acc);    // ops.cpp:873
          acc1 = fmaf(dv * float(q1 - 16), xhv[e], acc1);  // ops.cpp:874
        }
      }
      acc += acc1;
    }
    return acc;
  }
};

// K-quants: unit = one super-block of 256; sub = its 64-element quarter.
struct BodyQ4_K {
  static constexpr int C = 2;
  struct Frag {
    uint4 h, qa, qb;
  };
  __device__ __forceinline__ static void load(Frag& f, const GemvArgs& a, uint32_t slab, uint32_t u, int r, int c) {
    const size_t su = (size_t)slab * a.nb + u;
    f.h = ldg_stream(reinterpret_cast<const uint4*>(a.x) + su * 8 + r);
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (su * 4 + c) * 16 + r;
    f.qa = ldg_stream(q);
    f.qb = ldg_stream(q + 8);
  }
  __device__ __forceinline__ static float compute(const Frag& f, const GemvArgs& a, const uint8_t* sm, uint32_t sb, int c,
                                  float acc) {
    const int4* xp = reinterpret_cast<const int4*>(sm) + sb * 16 + c * 4;
    int sum_lo, sum_hi;
    q4_k_pair_dots(f.qa, f.qb, xp[0], xp[1], xp[2], xp[3], sum_lo, sum_hi);
    const uint2 b4 = reinterpret_cast<const uint2*>(sm + a.n_cols)[sb * 4 + c];  // bsums 4c..4c+3
    const int bs_lo = int(int16_t(b4.x & 0xffffu)) + int(int16_t(b4.x >> 16));
    const int bs_hi = int(int16_t(b4.y & 0xffffu)) + int(int16_t(b4.y >> 16));
    int sc1, m1, sc2, m2;
    q4_k_scale_min(f.h, 2 * c, sc1, m1);
    q4_k_scale_min(f.h, 2 * c + 1, sc2, m2);
    const float dx = reinterpret_cast<const float*>(sm + a.n_cols + a.n_cols / 8)[sb];
    const float dd = h2f(uint16_t(f.h.x & 0xffffu)) * dx;  // ops.cpp:654
    const float mm = h2f(uint16_t(f.h.x >> 16)) * dx;      // ops.cpp:655
    acc += fmaf(dd * float(sc1), float(sum_lo), -((mm * float(m1)) * float(bs_lo)));  // ops.cpp:671
    acc += fmaf(dd * float(sc2), float(sum_hi), -((mm * float(m2)) * float(bs_hi)));  // ops.cpp:682
    return acc;
  }
};

struct BodyQ6_K {
  static constexpr int C = 2;
  struct Frag {
    uint4 qa, qb, qh;
    uint2 sc;
    uint16_t d;
  };
  __device__ __forceinline__ static void load(Frag& f, const GemvArgs& a, uint32_t slab, uint32_t u, int r, int sub) {
    const size_t su = (size_t)slab * a.nb + u;
    const uint4* q = reinterpret_cast<const uint4*>(a.q) + (su * 12 + sub) * 8 + r;
    f.qa = ldg_stream(q);
    f.qb = ldg_stream(q + 32);
    f.qh = ldg_stream(q + 64);
    f.sc = ldg_stream(reinterpret_cast<const uint2*>(a.x) + (su * 8 + r) * 2 + (sub >> 1));
    f.d = ldg_stream(reinterpret_cast<const uint16_t*>(a.d) + su * 8 + r);
  }
  __device__ __forceinline__ static float compute(const Frag& f, const GemvArgs& a, const uint8_t* sm, uint32_t sb, int sub,
                                  float acc) {
    const int n = sub >> 1, hh = sub & 1;
    const int4* xs = reinterpret_cast<const int4*>(sm);
    const int16_t* bs = reinterpret_cast<const int16_t*>(sm + a.n_cols);
    const uint32_t g0 = sb * 16 + n * 8 + hh;  // 16-element group of x0
    const int part = q6_k_part(f.qa, f.qb, f.qh, xs[g0], xs[g0 + 2], xs[g0 + 4], xs[g0 + 6], sbyte(f.sc, hh),
                               sbyte(f.sc, hh + 2), sbyte(f.sc, hh + 4), sbyte(f.sc, hh + 6), bs[g0], bs[g0 + 2],
                               bs[g0 + 4], bs[g0 + 6]);
    const float dx = reinterpret_cast<const float*>(sm + a.n_cols + a.n_cols / 8)[sb];
    return fmaf(h2f(f.d) * dx, float(part), acc);  // ops.cpp:738,762
  }
};

// F16 (ops.cpp:541-586): x rounded to f16 first, products exact in fp32.
// BF16 (ops.cpp:908-916): x stays fp32.  unit = 4 chunks of 8 elements.
template <bool IS_BF16>
struct BodyHalf {
  static constexpr int C = 4;  // 4 units x 32 elements
  struct Frag {
    uint4 w;
  };
  __device__ __forceinline__ static void load(Frag& f, const GemvArgs& a, uint32_t slab, uint32_t u, int r, int sub) {
    const uint32_t c = 4 * u + sub;
    if (c < a.nb) f.w = ldg_stream(reinterpret_cast<const uint4*>(a.q) + ((size_t)slab * a.nb + c) * 8 + r);
  }
  __device__ __forceinline__ static float compute(const Frag& f, const GemvArgs& a, const uint8_t* sm, uint32_t u, int sub,
                                  float acc) {
    const uint32_t c = 4 * u + sub;
    if (c < a.nb) {
      float a0 = 0.0f, a1 = 0.0f;
      if (IS_BF16) {
        const float4 xa = reinterpret_cast<const float4*>(sm)[2 * c];
        const float4 xb = reinterpret_cast<const float4*>(sm)[2 * c + 1];
        a0 = fmaf(__uint_as_float(f.w.x << 16), xa.x, a0);
        a1 = fmaf(__uint_as_float(f.w.x & 0xffff0000u), xa.y, a1);
        a0 = fmaf(__uint_as_float(f.w.y << 16), xa.z, a0);
        a1 = fmaf(__uint_as_float(f.w.y & 0xffff0000u), xa.w, a1);
        a0 = fmaf(__uint_as_float(f.w.z << 16), xb.x, a0);
        a1 = fmaf(__uint_as_float(f.w.z & 0xffff0000u), xb.y, a1);
        a0 = fmaf(__uint_as_float(f.w.w << 16), xb.z, a0);
        a1 = fmaf(__uint_as_float(f.w.w & 0xffff0000u), xb.w, a1);
      } else {
        const uint4 xv = reinterpret_cast<const uint4*>(sm)[c];
        const float2 w0 = __half22float2(*reinterpret_cast<const __half2*>(&f.w.x));
        const float2 w1 = __half22float2(*reinterpret_cast<const __half2*>(&f.w.y));
        const float2 w2 = __half22float2(*reinterpret_cast<const __half2*>(&f.w.z));
        const float2 w3 = __half22float2(*reinterpret_cast<const __half2*>(&f.w.w));
        const float2 x0 = __half22float2(*reinterpret_cast<const __half2*>(&xv.x));
        const float2 x1 = __half22float2(*reinterpret_cast<const __half2*>(&xv.y));
        const float2 x2 = __half22float2(*reinterpret_cast<const __half2*>(&xv.z));
        const float2 x3 = __half22float2(*reinterpret_cast<const __half2*>(&xv.w));
        a0 = fmaf(w0.x, x0.x, a0);
        a1 = fmaf(w0.y, x0.y, a1);
        a0 = fmaf(w1.x, x1.x, a0);
        a1 = fmaf(w1.y, x1.y, a1);
        a0 = fmaf(w2.x, x2.x, a0);
        a1 = fmaf(w2.y, x2.y, a1);
        a0 = fmaf(w3.x, x3.x, a0);
        a1 = fmaf(w3.y, x3.y, a1);
      }
      acc += a0 + a1;
    }
    return acc;
  }
};

// ------------------------------------------------------------ kernel skeleton
template <class B, int N>
struct FragSet {
  typename B::Frag f[N];
};

template <class B, int N>
__device__ __forceinline__ void load_item(FragSet<B, N>& fs, const GemvArgs& a, uint32_t slab, uint32_t j, int r,
                                          int sub) {
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const uint32_t u = j * N + t;
    if (u < a.units) B::load(fs.f[t], a, slab, u, r, sub);
  }
}

template <class B, int N>
__device__ __forceinline__ float compute_item(const FragSet<B, N>& fs, const GemvArgs& a, const uint8_t* sm,
                                              uint32_t j, int sub) {
  float acc = 0.0f;
#pragma unroll
  for (int t = 0; t < N; ++t) {
    const uint32_t u = j * N + t;
    if (u < a.units) acc = B::compute(fs.f[t], a, sm, u, sub, acc);
  }
  acc += __shfl_xor_sync(0xffffffffu, acc, 8);
  acc += __shfl_xor_sync(0xffffffffu, acc, 16);
  return acc;  // lanes 0..7: chunk partial of rows 0..7
}

}  // namespace
