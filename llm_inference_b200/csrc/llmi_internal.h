// Internal declarations shared by the translation units of libllmi_cuda.so.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <string>

#include "llmi_cuda.h"

// ---------------------------------------------------------------------------
// Slab layout.  Output rows are grouped into slabs of 8.  Inside
// a slab every plane is stored "unit-major, row-minor" in 16-byte items, so that
// the 32 lanes of a warp (lane = 8*sub + row) read 512 contiguous bytes per
// 128-bit load, for any K.  Rows past the end of the matrix are padded with
// zero-weight blocks.
// ---------------------------------------------------------------------------
constexpr int LLMI_SLAB = 8;

struct llmi_weight_s {
  uint32_t type = 0;
  uint64_t n_cols = 0, n_rows = 0;       // K, N of the full matrix
  uint64_t row_begin = 0, row_end = 0;   // rows held by this handle
  uint64_t n_local = 0, n_slabs = 0;
  uint64_t nb = 0;        // K-units per row: 32-blocks, 256-super-blocks or 8-halves chunks
  uint8_t* base = nullptr;  // one allocation, planes below point into it
  size_t bytes = 0;
  // format-specific planes (see repack.cu for the exact item order)
  uint8_t* p_q = nullptr;    // quants (16-byte items)
  uint8_t* p_d = nullptr;    // f16 block scales
  uint8_t* p_x = nullptr;    // Q5_0: qh words; Q4_K: header items; Q6_K: int8 scales
};

// kinds of prepared activation
enum : int { ACT_NONE = 0, ACT_Q8_0 = 1, ACT_Q8_K = 2, ACT_F16 = 3, ACT_F32 = 4 };
constexpr int ACT_BF16_RAW = 5;  // throughput prefill only: a [token][K] bf16 batch handed to fast_pack_act_kernel as it is

// One bit per supported weight format / per activation kind: kernels that are compiled for a subset of the formats
// (mega_impl.cuh) carry a mask of these as a template parameter.
#ifdef __CUDACC__
#define LLMI_HD __host__ __device__
#else
#define LLMI_HD
#endif
LLMI_HD constexpr uint32_t llmi_type_bit(uint32_t ggml_type) {
  return ggml_type == LLMI_Q4_0 ? 1u : ggml_type == LLMI_Q8_0 ? 2u : ggml_type == LLMI_Q5_0 ? 4u : ggml_type == LLMI_Q4_K ? 8u
       : ggml_type == LLMI_Q6_K ? 16u : ggml_type == LLMI_F16 ? 32u : ggml_type == LLMI_BF16 ? 64u : 0x80000000u;
}
constexpr uint32_t LLMI_ALL_TYPES = 0x7fu;
// activation kinds the formats of a type mask consume (bit = 1 << kind)
LLMI_HD constexpr uint32_t llmi_kind_mask(uint32_t type_mask) {
  return ((type_mask & (llmi_type_bit(LLMI_Q4_0) | llmi_type_bit(LLMI_Q8_0))) ? (1u << ACT_Q8_0) : 0u) |
         ((type_mask & (llmi_type_bit(LLMI_Q4_K) | llmi_type_bit(LLMI_Q6_K))) ? (1u << ACT_Q8_K) : 0u) |
         ((type_mask & llmi_type_bit(LLMI_F16)) ? (1u << ACT_F16) : 0u) |
         ((type_mask & (llmi_type_bit(LLMI_Q5_0) | llmi_type_bit(LLMI_BF16))) ? (1u << ACT_F32) : 0u);
}

// Device layout of a prepared activation (one contiguous, 16-byte-multiple
// buffer so a single bulk async copy stages it into shared memory):
//   ACT_Q8_0: [K int8 quants][K/32 x {f16 d, int16 sum-of-quants}]
//   ACT_Q8_K: [K int8 quants][K/16 int16 bsums][K/256 fp32 d]
//   ACT_F16 : [K f16]            (x rounded to f16)
//   ACT_F32 : [K fp32]
struct llmi_act_s {
  uint64_t max_cols = 0;
  uint64_t n = 0;
  int kind = ACT_NONE;
  uint8_t* buf = nullptr;
  size_t buf_bytes = 0;
  cudaStream_t last_stream = nullptr;
};

inline size_t round_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

inline size_t act_bytes(int kind, uint64_t n) {
  switch (kind) {
    case ACT_Q8_0: return round_up(n + 4 * (n / 32), 16);
    case ACT_Q8_K: return round_up(n + 2 * (n / 16) + 4 * (n / 256), 16);
    case ACT_F16: return round_up(2 * n, 16);
    case ACT_F32: return round_up(4 * n, 16);
    default: return 0;
  }
}

// ---- flagged exchange of a row-sharded model (protocol: launch.cuh) ----------
constexpr int LLMI_MAX_WORLD = 8;
constexpr uint32_t LL_NONE = 0xffffffffu;

struct LLTag {
  const uint32_t* epoch = nullptr;
  uint32_t mul = 0, add = 0;
  uint32_t* err = nullptr;  // set to 1 by a consumer that gave up waiting (a peer died): the host checks it
};
struct LLPeers {
  uint2* base[LLMI_MAX_WORLD] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t n = 0;  // 0: not sharded
};

// Grow-only device scratch of the token-batched launches, one set PER STREAM: two models prefilling on different
// streams never share a partials / operand buffer.  Returns a buffer of at least `bytes` (a larger request drains the
// stream, frees and reallocates).  Everything is freed by llmi_shutdown.
enum : int { SCR_PART = 0, SCR_BQ, SCR_BD, SCR_FAST_W, SCR_FAST_X, SCR_ATT_Q, SCR_ATT_K, SCR_COUNT };
cudaError_t llmi_stream_scratch(cudaStream_t s, int slot, size_t bytes, void** out);
void llmi_stream_scratch_release(cudaStream_t s);

// Token batches of a row-sharded model (throughput mode): a GEMM whose output batch lives in the exchange allocation
// stores every element into the same place of every peer's allocation from its own epilogue (the all-gather overlaps
// the math tile by tile); `done` reports whether the launch that ran could do so (else the caller copies afterwards).
struct GemmPush {
  LLPeers peers;
  uint32_t rank = 0;
  bool done = false;
  const float *skip_begin = nullptr, *skip_end = nullptr;  // an output batch inside this range stays local
};
struct PeerOut {  // the peers' images of an output pointer
  float* p[LLMI_MAX_WORLD - 1] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t n = 0;
};
inline PeerOut llmi_peer_out(const GemmPush* g, const float* mine) {
  PeerOut po;
  if (!g || (mine >= g->skip_begin && mine < g->skip_end)) return po;
  const ptrdiff_t off = reinterpret_cast<const char*>(mine) - reinterpret_cast<const char*>(g->peers.base[g->rank]);
  for (uint32_t r = 0; r < g->peers.n; ++r)
    if (r != g->rank) po.p[po.n++] = reinterpret_cast<float*>(reinterpret_cast<char*>(g->peers.base[r]) + off);
  return po;
}

// What a mat-vec launch needs to push its rows to every rank: the ranks' exchange buffers, the tag of this
// exchange, and per matrix of the batch the element offset of its output vector inside the buffer.
struct GemvLL {
  LLPeers peers;
  LLTag tag;
  uint32_t off[3] = {LL_NONE, LL_NONE, LL_NONE};
};

// One weight matrix (or row shard) as the mat-vec kernels see it (gemv_bodies.cuh); filled by gemv.cu per launch
// and by model.cu once per matrix for the persistent decode kernel (mega.cu).
struct GemvArgs {
  const uint8_t* q;
  const uint8_t* d;
  const uint8_t* x;
  const uint8_t* act;
  uint32_t act_bytes;
  float* out;  // already offset to this handle's first row
  uint32_t n_local, n_slabs, nb, n_cols;
  uint32_t units;   // K-units per slab row (4 blocks of 32 / one super-block / 4 chunks of 8)
  uint32_t chunks;  // K-chunks (work items) per slab = ceil(units / Body::C)
  // optional fused epilogue of the logits mat-vec: soft-cap + running argmax key
  unsigned long long* argmax_key;
  float softcap;
  uint32_t row0;  // global index of this handle's first row
  // token-batched launches (prefill): n_tok activation buffers act_stride bytes apart, outputs out_stride floats apart
  uint32_t n_tok, act_stride, out_stride;
  float* part;  // token-per-lane kernel: chunk partials [chunk][token][8 * n_slabs] of this matrix
  // row-sharded model: element offset of this matrix' output vector inside the exchange buffer of every rank
  // (LL_NONE: plain local store to `out`)
  uint32_t ll_off;
};

// error plumbing (capi.cu)
void llmi_set_error(const std::string& msg);
int llmi_fail(int code, const std::string& msg);
int llmi_cuda_fail(cudaError_t e, const char* what);

#define LLMI_CUDA_TRY(expr)                                  \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return llmi_cuda_fail(_e, #expr); \
  } while (0)

// launchers implemented in the .cu files -------------------------------------
// repack.cu: raw (reference layout, rows [0,n_local)) -> planes of w
cudaError_t llmi_launch_repack(const llmi_weight_s& w, const uint8_t* raw_dev, cudaStream_t s);
cudaError_t llmi_launch_repack_slabs(const llmi_weight_s& w, const uint8_t* raw_chunk, uint64_t slab0, uint64_t n_sl,
                                     cudaStream_t s);
// capi.cu: llmi_weight_upload without the final wait (model load uploads hundreds of matrices back to back through
// the staging pipeline and waits once, llmi_upload_wait)
int llmi_weight_upload_async(const void* host_blocks, uint32_t ggml_type, uint64_t n_cols, uint64_t n_rows,
                             uint64_t row_begin, uint64_t row_end, llmi_weight_t* out);
int llmi_upload_wait();
size_t llmi_plan_planes(llmi_weight_s& w);  // fills nb/n_slabs, returns total bytes, sets plane offsets relative to 0

// quantize.cu
cudaError_t llmi_launch_quantize_q8_0(const float* x, uint64_t n, uint8_t* buf, cudaStream_t s);
cudaError_t llmi_launch_quantize_q8_k(const float* x, uint64_t n, uint8_t* buf, cudaStream_t s);
cudaError_t llmi_launch_round_f16(const float* x, uint64_t n, uint8_t* buf, cudaStream_t s);
cudaError_t llmi_launch_export_q8_0(const uint8_t* buf, uint64_t n, uint8_t* out34, cudaStream_t s);
cudaError_t llmi_launch_export_q8_k(const uint8_t* buf, uint64_t n, uint8_t* out292, cudaStream_t s);

// gemv.cu
void llmi_gemv_shutdown();  // frees the scratch of the token-batched launches
void llmi_gemv_read_env();  // LLMI_NO_UMMA, LLMI_UMMA_MIN_TOKENS
cudaError_t llmi_gemv_init();  // opt-in dynamic shared memory for every instantiation
cudaError_t llmi_launch_gemv(const llmi_weight_s& w, const llmi_act_s& a, float* out, cudaStream_t s);
// token-batched mat-vec (prefill): n_tok activations act_bytes(kind, n) apart, outputs out_strides[i] floats apart
cudaError_t llmi_launch_gemv_tokens(const llmi_weight_s* const* ws, float* const* outs, const uint32_t* out_strides,
                                    int n, int act_kind, uint64_t act_n, const uint8_t* act_base, uint32_t n_tok,
                                    cudaStream_t s, GemmPush* push = nullptr);
cudaError_t llmi_launch_gemv_batch(const llmi_weight_s* const* ws, float* const* outs, int n, const llmi_act_s& a,
                                   cudaStream_t s, const GemvLL* ll = nullptr);  // same format, same activation, n <= 3
// front of the weights of up to two matrices into L2 (side stream, during a glue kernel): gemv.cu l2_prefetch_kernel
cudaError_t llmi_launch_l2_prefetch(const llmi_weight_s* const* ws, int n, size_t offset_bytes, size_t budget_bytes,
                                    cudaStream_t s);
// gate + up + GEGLU (+ Q8_0 quantizer of the result) in one launch: gemv.cu gemv_geglu_kernel
cudaError_t llmi_launch_gemv_geglu(const llmi_weight_s& gate, const llmi_weight_s& up, const llmi_act_s& a, uint8_t* act_out,
                                   cudaStream_t s, const GemvLL* ll = nullptr);
uint32_t llmi_gemv_chunks(const llmi_weight_s& w);
void llmi_gemv_set_shape(int warps, int slabs_per_cta);  // 0 = heuristic
// persistent bulk-copy-fed kernel (gemv_ring.cuh): mode 0 heuristic / 1 never / 2 wherever it fits; CTAs per SM and
// ring slots per warp (0 = default)
void llmi_gemv_set_ring(int mode, int ctas_per_sm, int depth, int warps);
// norm_act_kernel's stage as the prologue of the mat-vec launch it feeds (ring kernel only): cudaErrorNotSupported when
// the fused form does not apply — the caller then launches llmi_launch_norm_act + llmi_launch_gemv_batch
cudaError_t llmi_launch_gemv_batch_norm(const llmi_weight_s* const* ws, float* const* outs, int n, const float* y,
                                        const float* w_post, const float* h_in, float* h_out, const float* w_norm, uint32_t n_cols,
                                        double eps, cudaStream_t s);
// throughput prefill (gemm_bf16.cuh): token batches of >= 64 go through the dequantize-to-bf16 tcgen05 GEMM (not bit-exact)
void llmi_gemv_set_prefill_fast(int on);
cudaError_t llmi_launch_geglu_cols(float* gate, const float* up, uint32_t stride, uint32_t col0, uint32_t cols, uint32_t n_tok,
                                   bool fast, cudaStream_t s, const GemmPush* push = nullptr,
                                   void* hid16 = nullptr);
// (up == nullptr: `gate` already holds the hidden batch)
cudaError_t llmi_launch_fast_ffn_down(const llmi_weight_s& w, const float* gate, const float* up, float* out, uint32_t out_stride,
                                      uint32_t n_tok, cudaStream_t s, GemmPush* push = nullptr,
                                      const void* hid16 = nullptr);  // fast mode: ffn_down fed by gelu(gate) * up directly
int llmi_gemv_prefill_fast();
cudaError_t llmi_launch_gemv_argmax(const llmi_weight_s& w, const llmi_act_s& a, float* out, unsigned long long* key,
                                    float softcap, cudaStream_t s, const GemvLL* ll = nullptr);
cudaError_t llmi_launch_block_dots(const llmi_weight_s& w, const llmi_act_s& a, int32_t* dots_dev, cudaStream_t s);
int llmi_act_kind_for(uint32_t ggml_type);
