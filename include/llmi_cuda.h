/* llmi_cuda.h — C ABI of libllmi_cuda.so, the B200 (sm_100a) implementation of
 * llm_inference's quantized mat-vec hot path.
 *
 * The reference (corywalker/llm_inference) has no FFI: its operator boundary is
 * the set of free functions in ops.h, called from model.cpp.  This header is
 * the C boundary underneath a drop-in for those functions
 * (llm_inference_b200/host/ops_cuda.cpp defines the very same C++ symbols and
 * forwards here).  Every entry point cites the reference interface it replaces.
 *
 * Conventions: plain pointers and sizes only; every function returns an int
 * status (LLMI_OK == 0) and never throws; llmi_last_error() gives the message
 * of the last failure on the calling thread.  One host thread drives a device
 * (the reference's ops are not re-entrant either: ops.cpp:18-19).  "dev"
 * pointers are CUDA device pointers; streams are cudaStream_t passed as void*.
 * All device-tier calls are stream-ordered and CUDA-graph capturable.
 *
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * fails with LLMI_ERR_CUDA.
 */
#ifndef LLMI_CUDA_H
#define LLMI_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LLMI_ABI_VERSION 1

enum llmi_status {
  LLMI_OK = 0,
  LLMI_ERR_ARG = 1,         /* null pointer, bad range, K not a block multiple */
  LLMI_ERR_TYPE = 2,        /* "mat_vec_mul: unsupported tensor type N" (ops.cpp:953) */
  LLMI_ERR_SIZE = 3,        /* "...: input vector size mismatch" (ops.cpp:197 etc.) */
  LLMI_ERR_CUDA = 4,        /* CUDA runtime/driver failure, or no device */
  LLMI_ERR_STATE = 5        /* llmi_init not called, activation not prepared ... */
};

/* ggml tensor type ids as stored in GGUF (gguf.h:30-46) */
enum llmi_type {
  LLMI_F32 = 0, LLMI_F16 = 1, LLMI_Q4_0 = 2, LLMI_Q5_0 = 6, LLMI_Q8_0 = 8,
  LLMI_Q4_K = 12, LLMI_Q6_K = 14, LLMI_BF16 = 30
};

typedef struct llmi_weight_s* llmi_weight_t; /* one uploaded, repacked matrix */
typedef struct llmi_act_s* llmi_act_t;       /* one quantized activation vector */
typedef void* llmi_stream_t;                 /* cudaStream_t */

/* ---- lifecycle -------------------------------------------------------- */

/* Replaces init_ops(int n_threads) (ops.h:38, ops.cpp:21-24): instead of a
 * thread pool, selects the CUDA device and creates the library context.  Must
 * be called before anything else; idempotent for the same device. */
int llmi_init(int device);
int llmi_shutdown(void);
const char* llmi_last_error(void);
int llmi_abi_version(void);
/* number of SMs of the active device (grid sizing is exposed for benches) */
int llmi_sm_count(void);

/* ---- weights ---------------------------------------------------------- */

/* Replaces GGUFFile::get_tensor_data + per-call row pointer math
 * (gguf.cpp:354-356, ops.cpp:206,221): uploads rows [row_begin,row_end) of a
 * K=n_cols x N=n_rows matrix stored in the reference/GGUF block layout
 * (row r at host_blocks + r*row_bytes, dense) ONCE and repacks it on the device
 * into 16-byte-aligned quant / scale planes ("slab layout").
 * host_blocks may be unaligned (GGUF tensor offsets are not, model_test.cpp:377).
 * The full range [0,n_rows) is the single-GPU case; a sub-range is one rank's
 * shard of an output-row-sharded matrix (the multi-GPU image of the thread
 * partition at ops.cpp:439-448). */
int llmi_weight_upload(const void* host_blocks, uint32_t ggml_type,
                       uint64_t n_cols, uint64_t n_rows, uint64_t row_begin,
                       uint64_t row_end, llmi_weight_t* out);
int llmi_weight_free(llmi_weight_t w);
int llmi_weight_dims(llmi_weight_t w, uint32_t* ggml_type, uint64_t* n_cols,
                     uint64_t* n_rows, uint64_t* row_begin, uint64_t* row_end);
uint64_t llmi_weight_device_bytes(llmi_weight_t w);
/* bytes of one weight row in the reference layout (0 = unsupported / bad K) */
uint64_t llmi_row_bytes(uint32_t ggml_type, uint64_t n_cols);

/* Registry keyed by the host pointer of the blocks, for callers that only hold
 * (TensorInfo, GGUFFile) like ops.h's mat_vec_mul: first use uploads, later
 * uses hit the cache.  Lifetime = until llmi_registry_clear()/llmi_shutdown(). */
int llmi_registry_get(const void* host_blocks, uint32_t ggml_type,
                      uint64_t n_cols, uint64_t n_rows, llmi_weight_t* out);
int llmi_registry_clear(void);

/* ---- activations ------------------------------------------------------ */

int llmi_act_create(uint64_t max_cols, llmi_act_t* out);
int llmi_act_free(llmi_act_t a);

/* quantize_row_q8_0 (ops.h:94, ops.cpp:116-139) on the device: x_dev[n] fp32
 * -> int8 quants + f16 scales (bit-exact with the reference). */
int llmi_quantize_q8_0(const float* x_dev, uint64_t n, llmi_act_t a, llmi_stream_t s);
/* quantize_row_q8_k (ops.h:104, ops.cpp:142-178): int8 quants, fp32 scale per
 * 256, int16 sums per 16 (bit-exact with the reference). */
int llmi_quantize_q8_k(const float* x_dev, uint64_t n, llmi_act_t a, llmi_stream_t s);
/* x -> f16 rounding used by mat_vec_mul_fp16 (ops.cpp:542-551). */
int llmi_round_f16(const float* x_dev, uint64_t n, llmi_act_t a, llmi_stream_t s);
/* fp32 pass-through for the formats that do not quantize x (Q5_0 ops.cpp:856-878,
 * BF16 ops.cpp:908-916). */
int llmi_stage_f32(const float* x_dev, uint64_t n, llmi_act_t a, llmi_stream_t s);
/* Picks whichever of the four the weight format consumes. */
int llmi_act_prepare(llmi_weight_t w, const float* x_dev, llmi_act_t a, llmi_stream_t s);

/* Copy the quantized activation back in the reference's record layout, for
 * bit-exact checks: n/32 x BlockQ8_0 (34 B, ops.h:89-92) or n/256 x block_q8_K
 * (292 B, ops.h:98-102).  Synchronises the stream the act was produced on. */
int llmi_act_export_q8_0(llmi_act_t a, void* host_blocks);
int llmi_act_export_q8_k(llmi_act_t a, void* host_blocks);

/* ---- mat-vec, device tier (the timed path) ---------------------------- */

/* o[row_begin..row_end) = W[rows] . act   for any supported format; replaces
 * the compute_range bodies + pool fan-out of mat_vec_mul_q4_0/_q8_0/_q4_k/_q6_k/
 * _q5_0/_bf16/_fp16 (ops.cpp:188-931).  out_dev is the FULL n_rows-long fp32
 * vector; only this handle's row range is written. */
int llmi_gemv(llmi_weight_t w, llmi_act_t a, float* out_dev, llmi_stream_t s);
/* Up to 3 matrices of the SAME format that consume the same prepared activation
 * (q/k/v, gate/up: model.cpp:754,784,803 and 875,877) as ONE grid: the launch
 * floor is paid once and the small matrices fill the SMs together.  Results are
 * identical to n separate llmi_gemv calls. */
int llmi_gemv_batch(const llmi_weight_t* ws, float* const* outs_dev, int n, llmi_act_t a, llmi_stream_t s);
/* llmi_act_prepare + llmi_gemv: exactly one reference mat_vec_mul call. */
int llmi_mat_vec_mul_dev(llmi_weight_t w, const float* x_dev, llmi_act_t a,
                         float* out_dev, llmi_stream_t s);

/* Tuning knob for benches: pins the CTA shape of the mat-vec kernel — warps per
 * CTA (4, 8 or 16) and consecutive 8-row slabs per CTA; 0 = heuristic.  Results
 * never depend on it (canonical summation order). */
int llmi_set_gemv_shape(int warps, int slabs_per_cta);
/* Second tuning knob: the persistent, bulk-copy-fed form of the same mat-vec (one CTA range of (slab, K-chunk)
 * items per CTA, per-warp shared-memory rings filled by cp.async.bulk before the predecessor kernel has finished).
 * mode 0 = heuristic (large Q4_0 launches), 1 = never, 2 = wherever it fits; ctas_per_sm in 1..4, depth (ring slots
 * per warp) in 2..4 and warps per CTA (8 or 16), 0 = default (2 x 16 warps x 2 slots).  Bit-identical to the other
 * form (same items, same canonical order).  Env: LLMI_GEMV_RING=mode,cps,depth,warps. */
int llmi_set_gemv_ring(int mode, int ctas_per_sm, int depth, int warps);

/* Token-batched mat-vec (prefill; the M >= 16 entry SURVEY §8b calls llmi_gemm_prefill): x_dev is
 * [n_tokens][n_cols] fp32, out_dev [n_tokens][n_rows] fp32 (a row-shard handle fills its own rows).  The
 * reference has no batched matmul — its forward loops the tokens around mat_vec_mul (model.cpp:714-960) —
 * so the contract is "n_tokens calls of mat_vec_mul": every (token, row) is bit-identical to
 * llmi_mat_vec_mul_dev.  Three kernels by token count: a token loop around the one-token decomposition (any
 * format; weights read once per 8 tokens), from 16 tokens a dp4a kernel with the token on the lane (Q4_0 / Q8_0),
 * from 128 tokens the exact int8 tensor-core kernel (Q4_0 / Q8_0; tcgen05.mma kind::i8, one MMA per quant block,
 * the per-block fp32 scale-and-accumulate as its epilogue). */
int llmi_gemm_tokens(llmi_weight_t w, const float* x_dev, uint32_t n_tokens, float* out_dev, llmi_stream_t stream);

/* Prefill mode of llmi_gemm_tokens and of the model's prompt path.  0 (default) = exact: every (token, row) bit-identical
 * to llmi_mat_vec_mul_dev, the parity gate.  1 = fast: batches of >= 64 tokens go through a dequantize-to-bf16
 * tcgen05.mma (kind::f16, fp32 accumulation in tensor memory) GEMM — every format, tensor-core throughput, NOT
 * bit-exact: |o - o_exact| <= 2e-2 * max|o_exact| per call (tests state the bar); greedy-token agreement with the exact
 * path is reported by bench.py.  Env: LLMI_PREFILL=fast (read by llmi_init and every llmi_model_load). */
int llmi_set_prefill_mode(int mode);

/* Per-block integer dot products (must be bit-exact with the reference):
 * Q4_0/Q8_0: rows*(K/32) int32; Q4_K: rows*(K/32); Q6_K: rows*(K/128).
 * dots_host is indexed [local_row][block]. */
int llmi_debug_block_dots(llmi_weight_t w, llmi_act_t a, int32_t* dots_host);

/* ---- mat-vec, host-vector tier (what the ops.h drop-in calls) ---------- */

/* mat_vec_mul(o, w_tensor, gguf_file, x) (ops.h:53): H2D x, quantize, GEMV,
 * D2H o, synchronous like the reference (returns after the join, ops.cpp:450).
 * n_x must equal K (else LLMI_ERR_SIZE); n_o must equal N. */
int llmi_host_mat_vec_mul(llmi_weight_t w, const float* x, uint64_t n_x,
                          float* o, uint64_t n_o);
/* quantize_row_q8_0 / quantize_row_q8_k with host vectors (ops.h:94,104);
 * y receives reference-layout records. */
int llmi_host_quantize_row_q8_0(const float* x, uint64_t n, void* y);
int llmi_host_quantize_row_q8_k(const float* x, uint64_t n, void* y);

/* ---- device-resident forward ------------------------------ */

typedef struct llmi_model_s* llmi_model_t;

/* Model(GGUFFile&) (model.h:78, model.cpp:13-56) on the device: parses the GGUF
 * image (borrowed only during the call), uploads + repacks every matrix once,
 * allocates activations and an fp16 KV cache of max_positions (0 = 4096).
 * Architecture "gemma3" only; anything else returns LLMI_ERR_TYPE (use the
 * ops.h drop-in with the reference's model.cpp for those). */
int llmi_model_load(const void* gguf_image, uint64_t size, uint32_t max_positions, llmi_model_t* out);
/* Row-sharded model over `world` GPUs of one node, one PROCESS per GPU (llmi_init binds a process to one device and a CUDA IPC handle
 * cannot be opened by the process that created it): rank
 * `rank` uploads a contiguous, slab-aligned range of the output rows of EVERY matrix — the thread partition of
 * ops.cpp:439-448 lifted to devices — and keeps activations, norms, attention and the KV cache replicated.  The
 * vectors the mat-vecs produce are exchanged by the mat-vec kernels themselves: each output row is written,
 * together with a tag, by one 64-bit store into an exchange buffer on every rank over NVLink peer memory, and
 * the consuming kernel spins on the tag — no collective call, no extra launch.  A row is computed
 * start to finish on one device in the canonical order, so logits and tokens are bit-identical to world = 1.
 * Wiring: every rank calls llmi_model_comm_handle, the host all-gathers the 64-byte handles (torch.distributed,
 * MPI, ...), every rank calls llmi_model_comm_connect with the world x 64 bytes in rank order.  All ranks must
 * then make the same forward / decode calls.  world = 1 is llmi_model_load.
 * Prompts (n_tokens > 1) of a sharded model go through the token-batched kernels like the single-GPU model's: every
 * rank computes its rows of every token of a batch and copies that column block of the [token][row] batch into every
 * peer's buffer (plain stores over peer memory + a flag barrier, four exchanges per layer); bit-identical to the
 * single-GPU batch in both prefill modes.  LLMI_NO_SHARD_PREFILL=1 at load: token by token through the tagged exchange. */
int llmi_model_load_shard(const void* gguf_image, uint64_t size, uint32_t max_positions, int world, int rank,
                          llmi_model_t* out);
/* The row range [*row_begin, *row_end) of an n_rows matrix that rank `rank` of `world` holds: contiguous,
 * aligned to the 8-row slab of the device layout, as even as possible (trailing ranks may be empty when the
 * matrix has fewer slabs than ranks).  Pure host arithmetic (no device needed); llmi_model_load_shard uses it
 * for every matrix, a host that shards handles itself passes the result to llmi_weight_upload. */
int llmi_shard_range(uint64_t n_rows, int world, int rank, uint64_t* row_begin, uint64_t* row_end);
int llmi_model_comm_handle(llmi_model_t m, void* handle64);
int llmi_model_comm_connect(llmi_model_t m, const void* handles);
/* 1 if a kernel of this rank gave up (after ~4 s) waiting for exchanged rows (a peer died, stalled, or ran other
 * steps).  The flag is sticky: llmi_model_forward / llmi_model_decode_greedy return LLMI_ERR_STATE while it is set. */
int llmi_model_comm_error(llmi_model_t m);
/* Clears the flag (and the argmax scratch) after the host has dealt with the failure; every rank calls it, with all
 * ranks' streams drained, before the next step. */
int llmi_model_comm_reset(llmi_model_t m);
/* Unmaps the peers' exchange buffers (drains this rank's stream first).  Tear-down order of a sharded model: every
 * rank disconnects, the host synchronizes the ranks (barrier), then every rank calls llmi_model_free — a rank must not
 * free a buffer a peer still has mapped.  The model cannot run again afterwards. */
int llmi_model_comm_disconnect(llmi_model_t m);
int llmi_model_free(llmi_model_t m);
/* dims[8] = {n_layer, n_embd, n_ff, n_head, n_head_kv, head_dim, vocab, max_positions} */
int llmi_model_info(llmi_model_t m, uint32_t* dims, uint64_t* weight_bytes);
/* Model::forward(tokens, pos) (model.h:91, model.cpp:706-1048): host token ids
 * in, host logits of the last token out; activations and KV stay on the device.
 * A prompt (n_tokens > 1) goes through each layer in batches of up to 256 tokens
 * (env LLMI_PREFILL_BATCH) — the reference's layer-major loop with the tokens
 * inside — with bit-identical results to feeding the tokens one by one
 * (LLMI_NO_PREFILL=1 forces that).  Synchronous. */
int llmi_model_forward(llmi_model_t m, const int32_t* tokens, int n_tokens, int pos, float* logits_host);
/* CUDA-event time and kernel launches of the last llmi_model_forward (either pointer may be NULL) */
int llmi_model_last_forward_stats(llmi_model_t m, float* ms_device, int* launches);
/* Greedy generation loop of main.cpp:172-221 entirely on the device (argmax of
 * step i feeds step i+1; one CUDA graph launch per token): consumes first_token
 * at position pos, returns n_steps token ids and the CUDA-event time. */
int llmi_model_decode_greedy(llmi_model_t m, int32_t first_token, int pos, int n_steps, int32_t* out_tokens,
                             float* ms_device);
int llmi_model_last_logits(llmi_model_t m, float* logits_host);
/* kernels launched per decode token on the per-launch path; 1 (per llmi_model_decode_greedy / one-token
 * llmi_model_forward call) on the persistent-kernel path (for bench.py's gpu_launches) */
int llmi_model_launches_per_step(llmi_model_t m);
/* 1: this model decodes with the persistent kernel (one cooperative launch per call, csrc/mega_impl.cuh; opt-in
 * with LLMI_DECODE=mega at load time); 0: one CUDA graph of per-stage kernels per token. */
int llmi_model_decode_path(llmi_model_t m);

/* ---- device memory helpers (benches / tests; plain cudaMalloc wrappers) - */
int llmi_dev_alloc(uint64_t bytes, void** out_dev);
int llmi_dev_free(void* dev);
int llmi_h2d(void* dst_dev, const void* src_host, uint64_t bytes);
int llmi_d2h(void* dst_host, const void* src_dev, uint64_t bytes);
int llmi_device_sync(void);

#ifdef __cplusplus
}
#endif
#endif /* LLMI_CUDA_H */
