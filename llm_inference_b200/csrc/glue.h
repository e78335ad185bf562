// Launchers of the device-resident glue kernels (glue.cu).
#pragma once

#include <cuda_fp16.h>

#include "llmi_internal.h"

struct EmbedArgs {
  uint32_t type;
  uint64_t nb;
  uint32_t n_cols;
  const uint8_t *q, *d, *x;
  uint32_t row_begin, row_end;  // rows of token_embd this rank holds (a row-sharded model: the owner of the token's row
                                // pushes it to every rank through the flagged exchange buffer)
};

inline EmbedArgs make_embed_args(const llmi_weight_s& w) {
  return EmbedArgs{w.type, w.nb, uint32_t(w.n_cols), w.p_q, w.p_d, w.p_x, uint32_t(w.row_begin), uint32_t(w.row_end)};
}

// Row-sharded model: where a glue kernel finds / sends exchanged vectors (launch.cuh).  peers.n == 0: single GPU.
struct LLCtx {
  LLPeers peers;
  LLTag tag;
  uint32_t rank = 0;
};

struct NormArgs {
  const float* y = nullptr;       // optional: output of the previous mat-vec (post-norm + residual stage)
  const float* w_post = nullptr;  // its norm weight
  float* h = nullptr;             // residual stream (updated in place when y != nullptr)
  const float* w = nullptr;       // optional: norm weight of the next stage
  uint32_t n = 0;
  double eps = 0.0;
  float* xn_out = nullptr;        // optional fp32 copy of the normalized vector
  int act_kind = ACT_NONE;
  uint8_t* act_buf = nullptr;
  int32_t* pos_inc = nullptr;     // optional device counter to bump (end of a token step) by the number of tokens
  uint32_t pos_inc_by = 0;        // 0: the number of tokens of this launch (a launch over a token slice passes the batch size)
  uint32_t n_tok = 1;             // prefill batch: one CTA per token, vectors n apart, activations act_stride apart
  uint32_t act_stride = 0;
  // row-sharded model: y arrives through the flagged exchange buffer instead (tag = ll_tag); the last norm of a
  // token step bumps the exchange epoch
  const uint2* ll_y = nullptr;
  LLTag ll_tag;
  uint32_t* epoch_inc = nullptr;
};

struct AttnArgs {
  const float *q, *k, *v;          // raw outputs of the q/k/v mat-vecs: [H*D], [HK*D], [HK*D]
  const float *wq_norm, *wk_norm;  // [D]
  uint32_t* kcache;                // this layer's keys [HK][t_max][D]: high word of double(f16 value) per element
  __half* vcache;                  // this layer's values [HK][t_max][D] (f16); row `pos` of every head is appended
  uint32_t H, HK, D, t_max;
  double eps;
  float attn_scale;
  const float2* rope_table;  // [t_max][D/2] (cos, sin) for this layer's rope base
  const int32_t* pos;
  float softcap;
  float* out;  // [H*D]
  int act_kind = ACT_NONE;  // fused quantizer for the attn_output mat-vec (ACT_NONE: caller quantizes)
  uint8_t* act_buf = nullptr;
  // prefill batch (grid.y = token): q/k/v/out are H*D resp. HK*D apart, activations act_stride apart, *pos is
  // the position of token 0; qbuf [n_tok][H][D] carries the rotated q between the two kernels
  uint32_t* qbuf = nullptr;
  uint32_t act_stride = 0;
  // row-sharded model: q/k/v arrive through the flagged exchange buffer instead
  const uint2 *ll_q = nullptr, *ll_k = nullptr, *ll_v = nullptr;
  LLTag ll_tag;
  // token batch of a row-sharded model: the main kernel (the tensor-core kernel of the throughput mode, MODE 2 of the
  // exact one) runs the KV heads [hk_begin, hk_begin + hk_count) only and writes their query heads' columns of `out`;
  // the K/V part of the prologue (norm, RoPE, append for every KV head) stays replicated.  hk_count == 0: every head.
  uint32_t hk_begin = 0, hk_count = 0;
};

cudaError_t llmi_launch_embed(const EmbedArgs& a, const int32_t* token, float scale, float* h, cudaStream_t s,
                              uint32_t n_tok = 1, const LLCtx* ll = nullptr, uint32_t ll_off = 0);
cudaError_t llmi_launch_norm_act(const NormArgs& a, cudaStream_t s);
void llmi_glue_read_env();  // LLMI_NORM_CLUSTER (re-read by every llmi_model_load)
cudaError_t llmi_launch_act(const float* x, uint32_t n, int kind, uint8_t* buf, cudaStream_t s, uint32_t n_tok = 1,
                            uint32_t act_stride = 0);
cudaError_t llmi_launch_rope_table(float2* table, uint32_t t_max, uint32_t D, float base, float scale, cudaStream_t s);
size_t llmi_attention_smem(uint32_t t_max, uint32_t D);
cudaError_t llmi_attention_init(uint32_t t_max, uint32_t D);
cudaError_t llmi_launch_attention(const AttnArgs& a, cudaStream_t s, uint32_t n_tok = 1);
bool llmi_attention_batch_tc(uint32_t H, uint32_t HK, uint32_t D);  // a token batch runs the tensor-core attention kernel
cudaError_t llmi_launch_geglu_act(const float* gate, const float* up, uint32_t n, int kind, uint8_t* buf,
                                  float* hidden_out, cudaStream_t s, uint32_t n_tok = 1, uint32_t act_stride = 0,
                                  const uint2* ll_gate = nullptr, const uint2* ll_up = nullptr,
                                  const LLTag* tag = nullptr);
// ll != nullptr: every rank's running key travels to every rank (two flagged words at ll_off + 2*rank) and the
// token is the maximum over the ranks
cudaError_t llmi_launch_finish_token(unsigned long long* key, int32_t* cur_tok, int32_t* gen, int32_t* gen_count,
                                     cudaStream_t s, const LLCtx* ll = nullptr, uint32_t ll_off = 0);
// exchanged vector -> plain floats (+ optional soft-cap): the host-facing logits of a row-sharded model
cudaError_t llmi_launch_ll_unpack(const uint2* ll, const LLTag& tag, float* out, uint32_t n, float softcap,
                                  cudaStream_t s);
// Token batches of a row-sharded model (glue.cu bx_exchange_kernel): this rank's columns of up to three
// [n_tok][stride] fp32 batches go to every peer, then a barrier over the ranks.
struct BxSeg {
  uint64_t byte_off = 0;                 // of the batch buffer inside every rank's exchange allocation
  uint32_t stride = 0, col0 = 0, cols = 0;  // floats per token; this rank's columns [col0, col0 + cols)
};
struct BxArgs {
  LLPeers peers;
  uint32_t rank = 0, n_seg = 0, n_tok = 0;
  BxSeg seg[3];
  uint32_t flag_off = 0;        // element offset of the barrier flags (one per rank) in the exchange buffer
  uint32_t seq = 0;             // number of this exchange (the same on every rank)
  uint32_t* counter = nullptr;  // CTA ticket (device word, zero between launches)
  uint32_t* err = nullptr;      // as LLTag::err
};
cudaError_t llmi_launch_bx_exchange(const BxArgs& a, cudaStream_t s);
cudaError_t llmi_launch_embed_shard_batch(const EmbedArgs& a, const int32_t* token, float scale, const LLPeers& peers,
                                          uint64_t h_byte_off, uint32_t n_tok, cudaStream_t s);
cudaError_t llmi_launch_softcap(float* logits, uint32_t n, float softcap, cudaStream_t s);
