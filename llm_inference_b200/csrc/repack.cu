// Weight upload: reference/GGUF block layout -> slab layout planes.
//
// The reference reads weights straight out of the GGUF mmap, block by block
// (ops.cpp:206,221,374-377: 18-byte Q4_0 records at 2-byte alignment at best).
// On the GPU every weight byte is read exactly once per mat-vec, so the layout
// is chosen for the load unit: quants and scales are split into separate
// planes of 16-byte items, and the items of the 8 rows of a slab are
// interleaved so a warp's 128-bit loads are contiguous (512 B per request).
//
// Item order of each plane (s = slab, u = K-unit, r = row in slab):
//   Q4_0  q[(s*nb+u)*8+r]            = 16 nibble bytes of block u (ops.cpp:377)
//         d[(s*nb+u)*8+r]            = f16 scale
//   Q8_0  q[((s*nb+u)*2+h)*8+r]      = int8 quants 16h..16h+15 of block u
//         d[(s*nb+u)*8+r]
//   Q5_0  q[(s*nb+u)*8+r] = qs, x[(s*nb+u)*8+r] = qh word, d[...] (ops.h:25-31)
//   Q4_K  x[(s*nb+u)*8+r]            = {f16 d, f16 dmin, scales[12]} (ops.h:11-16)
//         q[(((s*nb+u)*4+c)*2+h)*8+r] = qs bytes 32c+16h .. +15
//   Q6_K  q[(((s*nb+u)*3+k)*4+sub)*8+r], sub = 2n+hh (ops.h:18-23):
//            k=0: ql[64n+16hh..], k=1: ql[64n+32+16hh..], k=2: qh[32n+16hh..]
//         x[(s*nb+u)*8+r] = int8 scales[16],  d[(s*nb+u)*8+r] = f16 d
//   F16/BF16 q[(s*nb+u)*8+r]         = elements 8u..8u+7 of the row (0-padded)
#include "llmi_internal.h"

namespace {

__device__ __forceinline__ void copy16(uint8_t* dst, const uint8_t* src) {
  uint32_t w[4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
    w[i] = uint32_t(src[4 * i]) | (uint32_t(src[4 * i + 1]) << 8) | (uint32_t(src[4 * i + 2]) << 16) |
           (uint32_t(src[4 * i + 3]) << 24);
  *reinterpret_cast<uint4*>(dst) = make_uint4(w[0], w[1], w[2], w[3]);
}
__device__ __forceinline__ void fill16(uint8_t* dst, uint32_t v) {
  *reinterpret_cast<uint4*>(dst) = make_uint4(v, v, v, v);
}
__device__ __forceinline__ uint16_t ld16(const uint8_t* p) { return uint16_t(p[0]) | (uint16_t(p[1]) << 8); }

struct RepackArgs {
  uint32_t type;
  uint64_t n_local, n_slabs, nb, n_cols;
  uint8_t *q, *d, *x;
};

// cells [cell0, cell0 + n_cells) of the planes = a range of whole slabs; `raw` is the address local row 0 WOULD have
// (the caller subtracts the rows in front of the staged chunk)
__global__ void repack_kernel(RepackArgs a, const uint8_t* __restrict__ raw, uint64_t cell0, uint64_t n_cells) {
  const uint64_t idx = blockIdx.x * uint64_t(blockDim.x) + threadIdx.x;
  if (idx >= n_cells) return;
  const uint64_t cell = cell0 + idx;
  const int r = int(cell % LLMI_SLAB);
  const uint64_t su = cell / LLMI_SLAB;  // s*nb + u
  const uint64_t u = su % a.nb, s = su / a.nb;
  const uint64_t row = s * LLMI_SLAB + r;
  const bool live = row < a.n_local;
  uint16_t* dplane = reinterpret_cast<uint16_t*>(a.d);
  switch (a.type) {
    case LLMI_Q4_0: {
      const uint8_t* b = raw + (row * a.nb + u) * 18;
      if (live) {
        copy16(a.q + cell * 16, b + 2);
        dplane[cell] = ld16(b);
      } else {
        fill16(a.q + cell * 16, 0x88888888u);
        dplane[cell] = 0;
      }
    } break;
    case LLMI_Q8_0: {
      const uint8_t* b = raw + (row * a.nb + u) * 34;
      for (int h = 0; h < 2; ++h) {
        uint8_t* dst = a.q + ((su * 2 + h) * LLMI_SLAB + r) * 16;
        if (live) copy16(dst, b + 2 + 16 * h); else fill16(dst, 0);
      }
      dplane[cell] = live ? ld16(b) : uint16_t(0);
    } break;
    case LLMI_Q5_0: {
      const uint8_t* b = raw + (row * a.nb + u) * 22;
      uint32_t* xh = reinterpret_cast<uint32_t*>(a.x);
      if (live) {
        copy16(a.q + cell * 16, b + 6);
        xh[cell] = uint32_t(b[2]) | (uint32_t(b[3]) << 8) | (uint32_t(b[4]) << 16) | (uint32_t(b[5]) << 24);
        dplane[cell] = ld16(b);
      } else {
        fill16(a.q + cell * 16, 0);
        xh[cell] = 0;
        dplane[cell] = 0;
      }
    } break;
    case LLMI_Q4_K: {
      const uint8_t* b = raw + (row * a.nb + u) * 144;
      if (live) copy16(a.x + cell * 16, b); else fill16(a.x + cell * 16, 0);
      for (int c = 0; c < 4; ++c)
        for (int h = 0; h < 2; ++h) {
          uint8_t* dst = a.q + (((su * 4 + c) * 2 + h) * LLMI_SLAB + r) * 16;
          if (live) copy16(dst, b + 16 + 32 * c + 16 * h); else fill16(dst, 0);
        }
    } break;
    case LLMI_Q6_K: {
      const uint8_t* b = raw + (row * a.nb + u) * 210;
      for (int k = 0; k < 3; ++k)
        for (int sub = 0; sub < 4; ++sub) {
          const int n = sub >> 1, hh = sub & 1;
          const int off = k == 0 ? 64 * n + 16 * hh : k == 1 ? 64 * n + 32 + 16 * hh : 128 + 32 * n + 16 * hh;
          uint8_t* dst = a.q + (((su * 3 + k) * 4 + sub) * LLMI_SLAB + r) * 16;
          if (live) copy16(dst, b + off); else fill16(dst, 0);
        }
      if (live) copy16(a.x + cell * 16, b + 192); else fill16(a.x + cell * 16, 0);
      dplane[cell] = live ? ld16(b + 208) : uint16_t(0);
    } break;
    case LLMI_F16:
    case LLMI_BF16: {
      const uint8_t* rowp = raw + row * a.n_cols * 2;
      uint16_t v[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const uint64_t e = u * 8 + i;
        v[i] = (live && e < a.n_cols) ? ld16(rowp + 2 * e) : uint16_t(0);
      }
      *reinterpret_cast<uint4*>(a.q + cell * 16) =
          make_uint4(v[0] | (uint32_t(v[1]) << 16), v[2] | (uint32_t(v[3]) << 16), v[4] | (uint32_t(v[5]) << 16),
                     v[6] | (uint32_t(v[7]) << 16));
    } break;
    default: break;
  }
}

}  // namespace

// Computes nb / n_slabs and lays the planes out in one allocation (256-byte
// aligned each).  Plane pointers are set as OFFSETS from a null base; the
// caller rebases them after cudaMalloc.
size_t llmi_plan_planes(llmi_weight_s& w) {
  const uint64_t K = w.n_cols;
  w.n_local = w.row_end - w.row_begin;
  w.n_slabs = (w.n_local + LLMI_SLAB - 1) / LLMI_SLAB;
  size_t q = 0, d = 0, x = 0;
  switch (w.type) {
    case LLMI_Q4_0: w.nb = K / 32; q = 16; d = 2; break;
    case LLMI_Q8_0: w.nb = K / 32; q = 32; d = 2; break;
    case LLMI_Q5_0: w.nb = K / 32; q = 16; d = 2; x = 4; break;
    case LLMI_Q4_K: w.nb = K / 256; q = 128; x = 16; break;
    case LLMI_Q6_K: w.nb = K / 256; q = 192; d = 2; x = 16; break;
    case LLMI_F16:
    case LLMI_BF16: w.nb = (K + 7) / 8; q = 16; break;
    default: return 0;
  }
  const size_t cells = size_t(w.n_slabs) * w.nb * LLMI_SLAB;
  size_t off = 0;
  w.p_q = reinterpret_cast<uint8_t*>(off);
  off = round_up(off + cells * q, 256);
  w.p_d = reinterpret_cast<uint8_t*>(off);
  off = round_up(off + cells * d, 256);
  w.p_x = reinterpret_cast<uint8_t*>(off);
  off = round_up(off + cells * x, 256);
  return off;
}

cudaError_t llmi_launch_repack(const llmi_weight_s& w, const uint8_t* raw_dev, cudaStream_t s) {
  return llmi_launch_repack_slabs(w, raw_dev, 0, w.n_slabs, s);
}

// Slabs [slab0, slab0 + n_sl) only; raw_chunk holds the raw rows of exactly those slabs (upload pipeline, capi.cu).
cudaError_t llmi_launch_repack_slabs(const llmi_weight_s& w, const uint8_t* raw_chunk, uint64_t slab0, uint64_t n_sl,
                                     cudaStream_t s) {
  RepackArgs a{w.type, w.n_local, w.n_slabs, w.nb, w.n_cols, w.p_q, w.p_d, w.p_x};
  const uint64_t cells = n_sl * w.nb * LLMI_SLAB;
  if (cells == 0) return cudaSuccess;
  const uint64_t row_bytes = llmi_row_bytes(w.type, w.n_cols);
  const uint8_t* vraw = reinterpret_cast<const uint8_t*>(reinterpret_cast<uintptr_t>(raw_chunk) -
                                                         uintptr_t(slab0) * LLMI_SLAB * row_bytes);
  const int threads = 256;
  const uint64_t blocks = (cells + threads - 1) / threads;
  repack_kernel<<<dim3((unsigned)blocks), threads, 0, s>>>(a, vraw, slab0 * w.nb * LLMI_SLAB, cells);
  return cudaGetLastError();
}
