// Persistent decode kernel (mega.cu): descriptors built once per model by model.cu.
//
// One cooperative launch runs n_steps decode tokens of the whole device-resident forward
// (Model::forward, model.cpp:706-1048).  Every CTA walks the same list of phases; vectors travel between phases
// as flagged {value, tag} words in the exchange buffer (launch.cuh) — of this rank, and of every rank of a
// row-sharded model — so there is neither a kernel boundary nor a grid barrier between the stages of a layer.
#pragma once

#include <cuda_fp16.h>

#include "llmi_internal.h"

// how a mat-vec phase obtains its input vector
enum : uint32_t {
  MEGA_PRO_QUANT = 0,  // flagged fp32 vector -> activation quantizer (attn_output, ffn_down)
  MEGA_PRO_NORM = 1,   // flagged y: h += rmsnorm(y)*w_post (residual), xn = rmsnorm(h)*w_next -> quantizer
  MEGA_PRO_FIRST = 2,  // layer 0: xn = rmsnorm(h)*w_next -> quantizer (h comes from the embedding stage)
  MEGA_PRO_REUSE = 3,  // the activation the previous entry prepared (same vector, matrices of another format)
};
enum : uint32_t { MEGA_GEMV = 0, MEGA_ATTN = 1 };
// what happens to a row's sum
enum : uint32_t {
  MEGA_EPI_FLAG = 0,    // flagged store into the exchange buffer(s)
  MEGA_EPI_GEGLU = 1,   // matrices are (gate, up): hidden = gelu(gate) * up (model.cpp:887-901), flagged store
  MEGA_EPI_LOGITS = 2,  // final soft-cap, then plain store / flagged store / running argmax key
};

constexpr int MEGA_MAX_MATS = 3;

// One entry of the program every CTA walks per token: a mat-vec phase (prologue + matrices of ONE format + epilogue)
// or the attention stage of a layer.
struct MegaPhase {
  uint32_t kind, layer;       // MEGA_GEMV / MEGA_ATTN (then only `layer` is used: index into MegaArgs::attn)
  GemvArgs m[MEGA_MAX_MATS];  // q/d/x planes, nb, n_cols, units, chunks, n_local, n_slabs, row0
  uint32_t slab_q[MEGA_MAX_MATS], slab_d[MEGA_MAX_MATS], slab_x[MEGA_MAX_MATS];  // bytes of one slab per plane
  uint32_t chunk_q[MEGA_MAX_MATS], chunk_d[MEGA_MAX_MATS], chunk_x[MEGA_MAX_MATS];  // ... of one K-chunk of a slab
  uint32_t n_mats, type, act_kind, K, J;
  uint32_t v_total;  // slabs of all matrices (GEGLU: slab PAIRS) — the unit dealt to the CTAs
  uint32_t pro, epi;
  uint32_t in_off, in_tag;                    // flagged input vector (element offset, tag index inside the step)
  uint32_t out_off[MEGA_MAX_MATS], out_tag;   // flagged output vector per matrix
  uint32_t push_all;                          // 1: every rank's buffer, 0: this rank's only
  // arrival-counter hints (mega.cu): the slot the input's producers bump, the layer whose use of it is awaited, the
  // bumps per use; the slot this entry bumps (only the last entry of a stage does)
  uint32_t in_slot, in_layer, in_arrivals, out_slot, bump;
  const float *w_post, *w_next;               // norm weights of MEGA_PRO_NORM / _FIRST (w_next may be null: no output)
};

struct MegaAttn {
  const float *q_norm, *k_norm;
  uint32_t* kcache;
  __half* vcache;
  const float2* rope;
  uint32_t off_q, off_k, off_v, in_tag;
  uint32_t off_out, out_tag;
};

struct MegaArgs {
  uint32_t L, E, F, H, HK, D, V, t_max;
  double eps;
  float attn_scale, attn_softcap, final_softcap, embed_scale;
  const MegaPhase* prog;  // per layer: qkv (one entry per format), attention, attn_output, gate/up, down; then logits
  uint32_t n_prog;
  const MegaAttn* attn;   // [L]
  // token embedding (row gather + dequantization)
  uint32_t embd_type, embd_row_begin, embd_row_end;
  uint64_t embd_nb;
  const uint8_t *embd_q, *embd_d, *embd_x;
  // exchange
  LLPeers peers;  // n = world; base[rank] is this rank's buffer
  uint32_t rank, world;
  uint32_t* err;
  uint32_t tag_mul, epoch0;
  uint32_t off_h, off_key, off_tok, off_logits, off_cnt;
  uint32_t hints;      // 1: consumers sleep on the arrival counters before they take flagged words
  uint32_t pf_mode;    // 1: L2 prefetch of the weights ahead of their loads (0: off, for A/B measurements)
  uint32_t tok_uses0;  // greedy steps this model ran before this launch (the token counter's value)
  uint32_t tag_h, tag_key, tag_tok, tag_logits;
  uint32_t head_begin, head_end;  // query heads this rank runs
  uint32_t attn_push_all;         // head-sharded attention: a head's outputs go to every rank
  // step control
  const int32_t* tokens;  // optional: the token of every step (prompts); null: first_token, then the argmax
  int32_t first_token;
  int pos0;
  uint32_t n_steps;
  uint32_t logits_mode;  // 0: none, 1: plain logits of the last step, 2: greedy argmax every step
  float* logits;
  unsigned long long* key;
  uint32_t* done_ctr;
  int32_t *gen, *d_tok, *d_pos;
  // shared-memory layout (byte offsets into the dynamic segment)
  uint32_t sm_h, sm_xs, sm_wp, sm_wn, sm_act, sm_part, sm_attn;
  uint32_t part_floats, attn_nbuf;
};

constexpr uint32_t MEGA_TAGS_PER_LAYER = 5;  // qkv, attention, attn_output, hidden, ffn_down
// arrival counters: one per 128-byte line of the exchange buffer (16 elements apart) — embedding, token, then the
// five per-layer exchanges in two copies (layer parity)
constexpr uint32_t MEGA_CNT_STRIDE = 16, MEGA_CNT_H = 0, MEGA_CNT_TOK = 1, MEGA_CNT_LAYER = 2;
constexpr uint32_t MEGA_CNT_SLOTS = MEGA_CNT_LAYER + 2 * MEGA_TAGS_PER_LAYER;
constexpr int MEGA_THREADS = 512;  // one CTA per SM with a 128-register budget per thread (at 1024 threads / 64 registers the weight fragments spill)

// One instantiation of the kernel (mega_impl.cuh): the weight formats it carries code for, its head size (0: any).
struct MegaVariant {
  uint32_t type_mask;
  int head_dim;
  cudaError_t (*init)(size_t* smem_limit);
  cudaError_t (*launch)(const MegaArgs& a, uint32_t n_ctas, size_t smem, cudaStream_t s);
};
__host__ __device__ constexpr uint32_t mega_type_bit(uint32_t ggml_type) { return llmi_type_bit(ggml_type); }

cudaError_t llmi_mega_init();                      // opt-in shared memory of every instantiation, cooperative launch
uint32_t llmi_mega_max_ctas();                     // co-resident CTAs of the kernel (one per SM)
// the instantiation for a model's formats and head size (-1: none) and the dynamic shared memory it may use
int llmi_mega_select(uint32_t type_mask, uint32_t head_dim, size_t* smem_limit);
uint32_t llmi_mega_attention_nbuf(uint32_t t_max, uint32_t D, size_t avail, size_t* bytes);
cudaError_t llmi_launch_mega(int variant, const MegaArgs& a, uint32_t n_ctas, size_t smem, cudaStream_t s);
GemvArgs llmi_gemv_args(const llmi_weight_s& w);   // gemv.cu
