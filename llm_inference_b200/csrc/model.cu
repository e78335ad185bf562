// Device-resident Gemma-3 forward: weights uploaded once,
// activations and the fp16 KV cache resident on the device between ops, one
// decode token = one CUDA graph.  Follows Model::forward (model.cpp:706-1048)
// call for call: the 7 mat-vecs per layer + the logits mat-vec are llmi_gemv
// launches on the same repacked weights as the ops.h drop-in; everything
// between them is glue.cu.  Prefill is the reference's per-token loop
// (model.cpp:752-756 etc.): n tokens = n decode-shaped steps without logits.
//
// Only the "gemma3" architecture is handled here (gemma4's per-layer
// embeddings / shared KV / V-norm stay on the drop-in path through the
// reference's own model.cpp).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "glue.h"
#include "gguf_reader.h"
#include "mega.h"

namespace {

struct ActSet {  // one activation buffer per kind, created on demand
  llmi_act_t a[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};

struct LayerW {
  llmi_weight_t q = nullptr, k = nullptr, v = nullptr, o = nullptr, gate = nullptr, up = nullptr, down = nullptr;
  float *attn_norm = nullptr, *ffn_norm = nullptr, *post_attn_norm = nullptr, *post_ffw_norm = nullptr;
  float *q_norm = nullptr, *k_norm = nullptr;
  bool swa = false;
};

}  // namespace

struct llmi_model_s {
  uint32_t L = 0, E = 0, F = 0, H = 0, HK = 0, D = 0, V = 0, t_max = 0;
  double eps = 0;
  float rope_base = 0, rope_scale = 1.0f, attn_scale = 0, attn_softcap = 0, final_softcap = 0;
  std::vector<LayerW> layers;
  llmi_weight_t embd = nullptr;
  float* out_norm = nullptr;
  std::vector<void*> owned;  // plain device allocations to free
  float *h = nullptr, *xn = nullptr, *q = nullptr, *k = nullptr, *v = nullptr, *q_rot = nullptr, *attn = nullptr,
        *attn_out = nullptr, *gate = nullptr, *up = nullptr, *ffn_out = nullptr, *logits = nullptr;
  ActSet act_E, act_HD, act_F;
  uint32_t* kcache = nullptr;  // [L][HK][t_max][D] keys as double high words (glue.cu attention_kernel)
  __half* vcache = nullptr;    // [L][HK][t_max][D] values, f16
  int32_t *d_tok = nullptr, *d_pos = nullptr, *d_gen = nullptr, *d_gen_count = nullptr, *d_toks = nullptr;
  unsigned long long* d_key = nullptr;  // running argmax key of the logits mat-vec epilogue
  float2 *rope_swa = nullptr, *rope_global = nullptr;  // [t_max][D/2] (cos, sin) per rope base
  uint32_t toks_cap = 0, gen_cap = 0;
  cudaStream_t stream = nullptr;
  cudaGraphExec_t decode_graph = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float* logits_pinned = nullptr;
  uint64_t weight_bytes = 0;
  int launches_per_step = 0;
  // prefill: up to `batch` prompt tokens go through a layer together (run_batch).  The fp32 vectors above are
  // allocated `batch` deep; quantized activations of a batch live in bact_* (batch x act_bytes, lazily per kind)
  uint32_t batch = 1;
  uint8_t *bact_E[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}, *bact_HD[5] = {nullptr, nullptr, nullptr, nullptr, nullptr},
          *bact_F[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  uint32_t* qbuf = nullptr;  // [batch][H][D] rotated q between the two prefill attention kernels
  // L2 prefetch of the next mat-vec's weights on a side stream while a glue kernel holds the main stream (run_step):
  // megabytes requested during attention / the post-attention norm / GEGLU / the post-ffw norm (LLMI_PF_MB, 0 = off)
  cudaStream_t pf_stream = nullptr;
  std::vector<cudaEvent_t> pf_events;
  size_t pf_next = 0;
  bool pf_used = false;
  size_t pf_mb[4] = {48, 24, 24, 24};
  bool prefill_ok = true;
  float* h2 = nullptr;     // second residual buffer: a norm stage fused into a mat-vec launch reads one and writes the other
  bool fuse_norm = false;  // LLMI_FUSE_NORM=1: norm stages as prologues of the ring mat-vec launches they feed.  Bit-identical, two
                           // launches fewer per layer, but measured SLOWER (27b: 4.39 -> 5.30 ms per token): every CTA redoes
                           // the quantizer-heavy stage, 4x the single-CTA kernel's work per SM (profiles/r02_notes.md)
  bool fuse_geglu = false;  // LLMI_FUSE_GEGLU=1: gate, up and GEGLU as one launch (gemv_geglu_kernel).  Bit-identical, one
                            // launch fewer per layer, but measured SLOWER (1b 0.828 vs 0.765 ms/token, 27b 5.25 vs 5.02): a
                            // CTA that owns 32 rows of both matrices walks 3-11 items per warp one after the other,
                            // the separate grids put every item on its own warp (profiles/r02_notes.md)
  int prefill_launches = 0;
  // row-sharded model: every matrix holds the slab-aligned row range of `rank`; the vectors the
  // mat-vecs produce travel through `comm`, one flagged-exchange buffer per rank with the same layout everywhere
  // (element offsets off_*), written by the mat-vec epilogues of ALL ranks over peer memory (launch.cuh)
  int world = 1, rank = 0;
  uint2* comm = nullptr;
  size_t comm_elems = 0;
  LLCtx ll;                       // peers (filled by llmi_model_comm_connect), tag template, rank
  uint32_t *d_epoch = nullptr, *d_llerr = nullptr;
  uint32_t off_h = 0, off_q = 0, off_k = 0, off_v = 0, off_ao = 0, off_gate = 0, off_up = 0, off_fo = 0, off_key = 0,
           off_logits = 0;
  std::vector<void*> ipc_opened;  // peer mappings to close
  // token batches of a sharded model (run_batch): the fp32 batch buffers h, q, k, v, attn_out, gate, up, ffn_out and the
  // logits live INSIDE the exchange allocation (same byte offset on every rank), so a rank's columns can be copied
  // straight into its peers' buffers (glue.cu bx_exchange_kernel); off_bar: the barrier flags, bx_seq: exchanges so far
  uint32_t off_bar = 0, bx_seq = 0;
  void* hid16 = nullptr;  // [batch][F] bf16: the hidden batch of the throughput mode, rounded by the columns' owners
  uint32_t* d_bx_counter = nullptr;
  uint64_t bx_off(const void* p) const { return uint64_t(static_cast<const char*>(p) - reinterpret_cast<const char*>(comm)); }
  // persistent decode kernel (mega.cu): one cooperative launch per decode call.  Its exchange
  // region follows the per-launch path's inside `comm` (one allocation, one IPC handle).
  bool use_mega = false;
  MegaArgs mega;             // everything but the step-control fields
  std::vector<uint32_t> mega_off;  // MegaOffsets (element offsets of the per-layer vectors, two copies each)
  MegaPhase* d_prog = nullptr;
  MegaAttn* d_mattn = nullptr;
  uint32_t* d_done = nullptr;
  uint32_t mega_epoch = 1;   // steps run so far + 1 (every rank runs the same steps)
  uint32_t mega_tok_uses = 0;  // greedy steps run so far
  uint32_t mega_ctas = 0;
  size_t mega_smem = 0;
  int mega_variant = -1;     // the kernel instantiation for this model's formats and head size (mega.cu)
  int mega_launches = 0;     // kernel launches of the last decode / forward call
  bool sharded() const { return world > 1; }
  LLTag tag(uint32_t idx) const {  // idx: 1 + 4*layer + {0 qkv, 1 attn_out, 2 gate/up, 3 ffn_out}; 4L+1 embed, +2 keys, +3 logits
    LLTag t = ll.tag;
    t.add = idx;
    return t;
  }
};

namespace {

#define M_TRY(expr)                                          \
  do {                                                       \
    cudaError_t _e = (expr);                                 \
    if (_e != cudaSuccess) return llmi_cuda_fail(_e, #expr); \
  } while (0)
#define M_RC(expr)            \
  do {                        \
    int _rc = (expr);         \
    if (_rc != LLMI_OK) return _rc; \
  } while (0)

int dev_alloc(llmi_model_s* m, void** p, size_t bytes) {
  M_TRY(cudaMalloc(p, bytes ? bytes : 16));
  m->owned.push_back(*p);
  return LLMI_OK;
}

int upload_f32(llmi_model_s* m, const llmi::GgufTensor* t, uint64_t n, float** out, const char* name) {
  if (!t) return llmi_fail(LLMI_ERR_ARG, std::string("llmi_model_load: missing tensor ") + name);
  if (t->type != LLMI_F32 || t->n_elements() < n)
    return llmi_fail(LLMI_ERR_TYPE, std::string("llmi_model_load: ") + name + " must be F32");
  M_RC(dev_alloc(m, (void**)out, n * 4));
  M_TRY(cudaMemcpy(*out, t->data, n * 4, cudaMemcpyHostToDevice));
  return LLMI_OK;
}

int upload_matrix(llmi_model_s* m, const llmi::GgufTensor* t, uint64_t k, uint64_t n, llmi_weight_t* out,
                  const char* name) {
  if (!t) return llmi_fail(LLMI_ERR_ARG, std::string("llmi_model_load: missing tensor ") + name);
  if (t->shape.size() != 2 || t->shape[0] != k || t->shape[1] < n)
    return llmi_fail(LLMI_ERR_SIZE, std::string("llmi_model_load: unexpected shape of ") + name);
  // this rank's rows: contiguous, slab-aligned, as even as possible (shard.row_ranges; the thread partition of
  // ops.cpp:439-448 lifted to devices)
  uint64_t rb = 0, re = n;
  M_RC(llmi_shard_range(n, m->world, m->rank, &rb, &re));
  M_RC(llmi_weight_upload_async(t->data, t->type, k, n, rb, re, out));  // one wait for all of them: load_impl
  m->weight_bytes += llmi_row_bytes(t->type, k) * (re - rb);
  return LLMI_OK;
}

llmi_act_t get_act(ActSet& s, int kind, uint64_t n) {
  if (!s.a[kind]) {
    if (llmi_act_create(n, &s.a[kind]) != LLMI_OK) return nullptr;
  }
  s.a[kind]->kind = kind;
  s.a[kind]->n = n;
  return s.a[kind];
}

// Emits every activation kind the consumers of one fp32 vector need.  `first`
// was already produced by the fused kernel; any other kind is quantized from
// the fp32 copy.
int extra_acts(llmi_model_s* m, ActSet& set, const float* x, uint64_t n, int first,
               std::initializer_list<llmi_weight_t> consumers) {
  bool done[5] = {false, false, false, false, false};
  done[first] = true;
  for (llmi_weight_t w : consumers) {
    const int kind = llmi_act_kind_for(w->type);
    if (done[kind]) continue;
    llmi_act_t a = get_act(set, kind, n);
    if (!a) return LLMI_ERR_CUDA;
    M_TRY(llmi_launch_act(x, uint32_t(n), kind, a->buf, m->stream));
    m->launches_per_step++;
    done[kind] = true;
  }
  return LLMI_OK;
}

int gemv(llmi_model_s* m, llmi_weight_t w, ActSet& set, float* out) {
  llmi_act_t a = set.a[llmi_act_kind_for(w->type)];
  M_TRY(llmi_launch_gemv(*w, *a, out, m->stream));
  m->launches_per_step++;
  return LLMI_OK;
}

// Mat-vecs that consume the same activation vector: matrices of the same format
// go out as one grid (llmi_launch_gemv_batch), in the order given.
int gemv_group(llmi_model_s* m, std::initializer_list<llmi_weight_t> ws, std::initializer_list<float*> outs,
               ActSet& set, const GemvLL* ll = nullptr) {
  std::vector<llmi_weight_t> w(ws);
  std::vector<float*> o(outs);
  std::vector<bool> done(w.size(), false);
  for (size_t i = 0; i < w.size(); ++i) {
    if (done[i]) continue;
    const llmi_weight_s* bw[3];
    float* bo[3];
    GemvLL sub;  // the exchange offsets follow their matrices into the per-format launches
    if (ll) sub = *ll;
    int n = 0;
    for (size_t j = i; j < w.size() && n < 3; ++j)
      if (!done[j] && w[j]->type == w[i]->type) {
        bw[n] = w[j];
        bo[n] = o[j];
        if (ll) sub.off[n] = ll->off[j];
        done[j] = true;
        ++n;
      }
    llmi_act_t a = set.a[llmi_act_kind_for(w[i]->type)];
    M_TRY(llmi_launch_gemv_batch(bw, bo, n, *a, m->stream, ll ? &sub : nullptr));
    m->launches_per_step++;
  }
  return LLMI_OK;
}

// One token through all layers.  tok: device pointer to the token id.
int run_step(llmi_model_s* m, const int32_t* tok, bool want_logits, bool want_argmax) {
  cudaStream_t s = m->stream;
  m->launches_per_step = 0;
  const uint32_t E = m->E, F = m->F, HD = m->H * m->D;
  const bool sh = m->sharded();
  if (sh && m->ll.peers.n != uint32_t(m->world))
    return llmi_fail(LLMI_ERR_STATE, "row-sharded model: llmi_model_comm_connect has not been called");
  LLCtx ll_embed = m->ll;
  ll_embed.tag = m->tag(4 * m->L + 1);
  M_TRY(llmi_launch_embed(make_embed_args(*m->embd), tok, std::sqrt(float(E)), m->h, s, 1, sh ? &ll_embed : nullptr,
                          m->off_h));  // model.cpp:710-712
  m->launches_per_step++;
  // the residual stream lives in h_cur; a norm stage that runs as the prologue of the mat-vec it feeds
  // (llmi_launch_gemv_batch_norm) reads h_cur and leaves the updated stream in h_alt
  float *h_cur = m->h, *h_alt = m->h2;
  bool qkv_done = false;  // this layer's q/k/v launch already went out with the previous layer's last norm in front
  auto try_fused = [&](std::initializer_list<llmi_weight_t> ws, std::initializer_list<float*> outs, const float* y,
                       const float* w_post, const float* w_norm, bool* fused) -> int {
    *fused = false;
    if (sh || !m->fuse_norm || !h_alt) return LLMI_OK;
    const llmi_weight_s* bw[3];
    float* bo[3];
    int n = 0;
    for (llmi_weight_t w : ws) bw[n++] = w;
    n = 0;
    for (float* o : outs) bo[n++] = o;
    const cudaError_t e = llmi_launch_gemv_batch_norm(bw, bo, n, y, w_post, h_cur, h_alt, w_norm, E, m->eps, s);
    if (e == cudaErrorNotSupported) return LLMI_OK;
    M_TRY(e);
    std::swap(h_cur, h_alt);
    m->launches_per_step++;
    *fused = true;
    return LLMI_OK;
  };
  GemvLL gl;  // exchange context of the mat-vec launches of a sharded model
  if (sh) gl.peers = m->ll.peers;
  auto push = [&](uint32_t idx, std::initializer_list<uint32_t> offs) -> const GemvLL* {
    if (!sh) return nullptr;
    gl.tag = m->tag(idx);
    int i = 0;
    for (uint32_t o : offs) gl.off[i++] = o;
    return &gl;
  };
  // side-stream L2 prefetch of bytes [off, off + mb MB) of `ws`, ordered after everything launched on `s` so far
  auto prefetch = [&](std::initializer_list<llmi_weight_t> ws, size_t off, size_t mb) -> int {
    if (!m->pf_stream || mb == 0) return LLMI_OK;
    const llmi_weight_s* v[2];
    int n = 0;
    for (llmi_weight_t w : ws)
      if (n < 2) v[n++] = w;
    cudaEvent_t ev = m->pf_events[m->pf_next++ % m->pf_events.size()];
    M_TRY(cudaEventRecord(ev, s));
    M_TRY(cudaStreamWaitEvent(m->pf_stream, ev, 0));
    M_TRY(llmi_launch_l2_prefetch(v, n, off, mb << 20, m->pf_stream));
    m->pf_used = true;
    return LLMI_OK;
  };
  m->pf_used = false;
  for (uint32_t l = 0; l < m->L; ++l) {
    LayerW& w = m->layers[l];
    const int kq = llmi_act_kind_for(w.q->type);
    if (l == 0) {  // attn_norm of layer 0 (later layers: fused into the previous layer's last kernel)
      NormArgs na;
      na.h = h_cur; na.w = w.attn_norm; na.n = E; na.eps = m->eps; na.xn_out = m->xn;
      na.act_kind = kq; na.act_buf = get_act(m->act_E, kq, E)->buf;
      M_TRY(llmi_launch_norm_act(na, s));
      m->launches_per_step++;
    }
    if (!qkv_done) {
      M_RC(extra_acts(m, m->act_E, m->xn, E, kq, {w.k, w.v}));
      // model.cpp:754 (Q), 784 (K), 803 (V): one grid per format group
      M_RC(gemv_group(m, {w.q, w.k, w.v}, {m->q, m->k, m->v}, m->act_E, push(1 + 4 * l, {m->off_q, m->off_k, m->off_v})));
    }
    qkv_done = false;
    AttnArgs aa;  // q/k norm, RoPE, KV append and attention in one kernel
    if (sh) {
      aa.ll_q = m->comm + m->off_q; aa.ll_k = m->comm + m->off_k; aa.ll_v = m->comm + m->off_v;
      aa.ll_tag = m->tag(1 + 4 * l);
    }
    aa.q = m->q; aa.k = m->k; aa.v = m->v; aa.wq_norm = w.q_norm; aa.wk_norm = w.k_norm;
    aa.kcache = m->kcache + size_t(l) * m->t_max * m->HK * m->D;
    aa.vcache = m->vcache + size_t(l) * m->t_max * m->HK * m->D;
    aa.H = m->H; aa.HK = m->HK; aa.D = m->D; aa.t_max = m->t_max; aa.eps = m->eps;
    aa.rope_table = w.swa ? m->rope_swa : m->rope_global;  // base 10000 on SWA layers (model.cpp:732)
    aa.attn_scale = m->attn_scale; aa.pos = m->d_pos;
    aa.softcap = m->attn_softcap; aa.out = m->attn;
    const int ko = llmi_act_kind_for(w.o->type);
    uint8_t* ko_buf = get_act(m->act_HD, ko, HD)->buf;
    const bool fuse_act = ko != ACT_Q8_K || m->D % 256 == 0;  // a head holds whole quantization blocks
    if (fuse_act) {
      aa.act_kind = ko;
      aa.act_buf = ko_buf;
    }
    // while attention runs (H CTAs, no HBM traffic to speak of): attn_output, then the front of gate/up
    M_RC(prefetch({w.o}, 0, m->pf_mb[0]));
    if ((w.o->bytes >> 20) < m->pf_mb[0]) M_RC(prefetch({w.gate, w.up}, 0, m->pf_mb[0] - (w.o->bytes >> 20)));
    M_TRY(llmi_launch_attention(aa, s));
    m->launches_per_step++;
    if (!fuse_act) {
      M_TRY(llmi_launch_act(m->attn, HD, ko, ko_buf, s));
      m->launches_per_step++;
    }
    M_RC(gemv_group(m, {w.o}, {m->attn_out}, m->act_HD, push(2 + 4 * l, {m->off_ao})));  // model.cpp:557
    bool gate_up_done = false;  // post-attention norm + residual + ffn_norm as the prologue of the gate / up launch
    if (!m->fuse_geglu && w.gate->type == w.up->type)
      M_RC(try_fused({w.gate, w.up}, {m->gate, m->up}, m->attn_out, w.post_attn_norm, w.ffn_norm, &gate_up_done));
    if (!gate_up_done) {
      const int kg = llmi_act_kind_for(w.gate->type);
      NormArgs na;  // post-attention norm + residual, then ffn_norm (model.cpp:843-858)
      na.y = m->attn_out; na.w_post = w.post_attn_norm; na.h = h_cur; na.w = w.ffn_norm; na.n = E; na.eps = m->eps;
      if (sh) { na.y = nullptr; na.ll_y = m->comm + m->off_ao; na.ll_tag = m->tag(2 + 4 * l); }
      na.xn_out = m->xn; na.act_kind = kg; na.act_buf = get_act(m->act_E, kg, E)->buf;
      {  // while the single-CTA norm runs: gate/up, continuing behind what the attention gap asked for
        const size_t done = (w.o->bytes >> 20) < m->pf_mb[0] ? m->pf_mb[0] - (w.o->bytes >> 20) : 0;
        M_RC(prefetch({w.gate, w.up}, done << 20, m->pf_mb[1]));
      }
      M_TRY(llmi_launch_norm_act(na, s));
      m->launches_per_step++;
      M_RC(extra_acts(m, m->act_E, m->xn, E, kg, {w.up}));
    }
    const int kd = llmi_act_kind_for(w.down->type);
    // gate, up and GEGLU as one launch (gemv_geglu_kernel): single GPU when ffn_down takes Q8_0 activations (the CTA
    // that owns 32 rows of gate and up writes their Q8_0 block); row-sharded always (the hidden rows travel instead
    // of the gate and the up rows: half the exchange, and the quantizer below only gathers them)
    const bool fuse_geglu = m->fuse_geglu && w.gate->type == w.up->type && (sh || (kd == ACT_Q8_0 && F % 32 == 0));
    if (fuse_geglu) {
      llmi_act_t ag = m->act_E.a[llmi_act_kind_for(w.gate->type)];
      uint8_t* act_f = get_act(m->act_F, kd, F)->buf;
      M_TRY(llmi_launch_gemv_geglu(*w.gate, *w.up, *ag, act_f, s, push(3 + 4 * l, {m->off_gate})));  // model.cpp:875-901
      m->launches_per_step++;
      if (sh) {
        const LLTag tg = m->tag(3 + 4 * l);
        M_TRY(llmi_launch_geglu_act(nullptr, nullptr, F, kd, act_f, nullptr, s, 1, 0, m->comm + m->off_gate, nullptr, &tg));
        m->launches_per_step++;
      }
    } else {
      if (!gate_up_done)
        M_RC(gemv_group(m, {w.gate, w.up}, {m->gate, m->up}, m->act_E,
                        push(3 + 4 * l, {m->off_gate, m->off_up})));  // model.cpp:875, 877
      const LLTag tg = m->tag(3 + 4 * l);
      M_RC(prefetch({w.down}, 0, m->pf_mb[2]));  // while GEGLU runs
      M_TRY(llmi_launch_geglu_act(m->gate, m->up, F, kd, get_act(m->act_F, kd, F)->buf, nullptr, s, 1, 0,
                                  sh ? m->comm + m->off_gate : nullptr, sh ? m->comm + m->off_up : nullptr,
                                  sh ? &tg : nullptr));
      m->launches_per_step++;
    }
    M_RC(gemv_group(m, {w.down}, {m->ffn_out}, m->act_F, push(4 + 4 * l, {m->off_fo})));  // model.cpp:909
    if (l + 1 < m->L) {  // post-ffw norm + residual + the next layer's attn_norm as the prologue of its q / k / v launch
      LayerW& nx = m->layers[l + 1];
      if (nx.q->type == nx.k->type && nx.q->type == nx.v->type)
        M_RC(try_fused({nx.q, nx.k, nx.v}, {m->q, m->k, m->v}, m->ffn_out, w.post_ffw_norm, nx.attn_norm, &qkv_done));
    }
    if (!qkv_done) {
      NormArgs na;  // post-ffw norm + residual (model.cpp:915-924), then the next norm
      na.y = m->ffn_out; na.w_post = w.post_ffw_norm; na.h = h_cur; na.n = E; na.eps = m->eps; na.xn_out = m->xn;
      if (sh) { na.y = nullptr; na.ll_y = m->comm + m->off_fo; na.ll_tag = m->tag(4 + 4 * l); }
      if (l + 1 < m->L) {
        const int kn = llmi_act_kind_for(m->layers[l + 1].q->type);
        na.w = m->layers[l + 1].attn_norm; na.act_kind = kn; na.act_buf = get_act(m->act_E, kn, E)->buf;
      } else {
        na.pos_inc = m->d_pos;  // the token is done
        if (sh) na.epoch_inc = m->d_epoch;
        if (want_logits) {      // final RMSNorm (model.cpp:983-986)
          const int kl = llmi_act_kind_for(m->embd->type);
          na.w = m->out_norm; na.act_kind = kl; na.act_buf = get_act(m->act_E, kl, E)->buf;
        }
      }
      // while the single-CTA norm runs: the next layer's q/k/v (or the front of the logits matrix)
      if (l + 1 < m->L) {
        M_RC(prefetch({m->layers[l + 1].q, m->layers[l + 1].k}, 0, m->pf_mb[3]));
      } else if (want_logits) {
        M_RC(prefetch({m->embd}, 0, m->pf_mb[3]));
      }
      M_TRY(llmi_launch_norm_act(na, s));
      m->launches_per_step++;
    }
  }
  if (m->pf_used) {  // join the side stream (a captured graph must end in one stream)
    cudaEvent_t ev = m->pf_events[m->pf_next++ % m->pf_events.size()];
    M_TRY(cudaEventRecord(ev, m->pf_stream));
    M_TRY(cudaStreamWaitEvent(s, ev, 0));
  }
  if (want_logits) {
    if (want_argmax) {  // logits mat-vec (model.cpp:1000 / 1027) with the soft-cap + argmax epilogue
      llmi_act_t la = m->act_E.a[llmi_act_kind_for(m->embd->type)];
      LLCtx ll_key = m->ll;
      ll_key.tag = m->tag(4 * m->L + 2);
      // sharded: every rank keeps the logits of its own rows only and the ranks swap their argmax keys
      M_TRY(llmi_launch_gemv_argmax(*m->embd, *la, m->logits, m->d_key, m->final_softcap, s));
      M_TRY(llmi_launch_finish_token(m->d_key, m->d_tok, m->d_gen, m->d_gen_count, s, sh ? &ll_key : nullptr,
                                     m->off_key));
      m->launches_per_step += 2;
    } else if (sh) {  // host-facing logits: rows of every rank through the exchange buffer, then plain (+ soft-cap)
      M_RC(gemv_group(m, {m->embd}, {m->logits}, m->act_E, push(4 * m->L + 3, {m->off_logits})));
      M_TRY(llmi_launch_ll_unpack(m->comm + m->off_logits, m->tag(4 * m->L + 3), m->logits, m->V, m->final_softcap, s));
      m->launches_per_step++;
    } else {
      M_RC(gemv(m, m->embd, m->act_E, m->logits));
      if (m->final_softcap > 0.0f) {  // model.cpp:1036-1041
        M_TRY(llmi_launch_softcap(m->logits, m->V, m->final_softcap, s));
        m->launches_per_step++;
      }
    }
  }
  return LLMI_OK;
}

// Quantized activations of a prefill batch: `batch` buffers act_bytes(kind, n) apart.
uint8_t* get_bact(llmi_model_s* m, uint8_t** set, int kind, uint64_t n) {
  if (!set[kind]) {
    if (dev_alloc(m, (void**)&set[kind], size_t(m->batch) * act_bytes(kind, n)) != LLMI_OK) return nullptr;
  }
  return set[kind];
}

// Mat-vecs of a batch that consume the same activations: one grid per format group.
int gemv_tokens_group(llmi_model_s* m, std::initializer_list<llmi_weight_t> ws, std::initializer_list<float*> outs,
                      std::initializer_list<uint32_t> strides, uint8_t** set, uint64_t n, uint32_t n_tok,
                      GemmPush* push = nullptr) {
  bool all_pushed = push != nullptr;
  std::vector<llmi_weight_t> w(ws);
  std::vector<float*> o(outs);
  std::vector<uint32_t> st(strides);
  std::vector<bool> done(w.size(), false);
  for (size_t i = 0; i < w.size(); ++i) {
    if (done[i]) continue;
    const llmi_weight_s* bw[3];
    float* bo[3];
    uint32_t bs[3];
    int k = 0;
    for (size_t j = i; j < w.size() && k < 3; ++j)
      if (!done[j] && w[j]->type == w[i]->type) {
        bw[k] = w[j];
        bo[k] = o[j];
        bs[k] = st[j];
        done[j] = true;
        ++k;
      }
    const int kind = llmi_act_kind_for(w[i]->type);
    M_TRY(llmi_launch_gemv_tokens(bw, bo, bs, k, kind, n, set[kind], n_tok, m->stream, push));
    if (push) all_pushed = all_pushed && push->done;
    m->prefill_launches++;
  }
  if (push) push->done = all_pushed;
  return LLMI_OK;
}

// Row-sharded token batch: this rank's columns of the given [n_tok][stride] fp32 batches (the rows of the matrices
// that produced them) go to every peer, then a barrier (glue.cu bx_exchange_kernel).  No buffers: the barrier alone.
struct BxBuf {
  const float* buf;
  uint32_t stride;
  llmi_weight_t w;      // this rank's columns = the rows of w it holds, or (w == nullptr) [col0, col0 + cols)
  uint32_t col0 = 0, cols = 0;
};
int bx_exchange(llmi_model_s* m, std::initializer_list<BxBuf> bufs, uint32_t n_tok) {
  BxArgs a;
  a.peers = m->ll.peers;
  a.rank = uint32_t(m->rank);
  a.n_tok = n_tok;
  a.flag_off = m->off_bar;
  a.seq = ++m->bx_seq;
  a.counter = m->d_bx_counter;
  a.err = m->d_llerr;
  for (const BxBuf& b : bufs) {
    BxSeg& sg = a.seg[a.n_seg++];
    sg.byte_off = m->bx_off(b.buf);
    sg.stride = b.stride;
    sg.col0 = b.w ? uint32_t(b.w->row_begin) : b.col0;
    sg.cols = b.w ? uint32_t(b.w->row_end - b.w->row_begin) : b.cols;
  }
  M_TRY(llmi_launch_bx_exchange(a, m->stream));
  m->prefill_launches++;
  return LLMI_OK;
}

// n_tok <= batch prompt tokens through all layers together (the loop nest of the reference's forward is
// layer-major with the tokens inside, model.cpp:714-960).  Per token the arithmetic is the one run_step
// does — same kernels with a token index — so the KV cache and the logits are bit-identical to feeding the
// tokens one by one; what changes is that every weight matrix is read once per token tile (8 to 32 tokens,
// gemv.cu launch_tokens) instead of once per token and a layer costs ~9 launches per batch instead of 8 per token.  Requires prefill_ok (every
// consumer of a vector takes the same activation kind).  Logits (of the last token) only if want_logits.
int run_batch(llmi_model_s* m, const int32_t* toks, uint32_t n_tok, bool want_logits) {
  cudaStream_t s = m->stream;
  const uint32_t E = m->E, F = m->F, HD = m->H * m->D, KD = m->HK * m->D;
  // Row-sharded model: every token-batched mat-vec computes this rank's rows of every token (the
  // columns [row_begin, row_end) of the [n_tok][rows] batch) and the column blocks are all-gathered over NVLink peer
  // memory — by bx_exchange_kernel, or by the producer's own stores followed by the bare barrier.  GEGLU runs on the
  // columns' owner, the main attention kernel with the rank's own heads, long batches' norm stages on a token slice
  // per rank; the K/V prologue, the KV cache and the quantizers are replicated as in run_step.  Every value is still
  // computed start to finish on one device, so the batch is bit-identical to the single-GPU batch of the same mode.
  const bool sh = m->sharded();
  if (sh && m->ll.peers.n != uint32_t(m->world))
    return llmi_fail(LLMI_ERR_STATE, "row-sharded model: llmi_model_comm_connect has not been called");
  // throughput mode: the GEMM epilogues (and the GEGLU kernel) store into every peer's batch themselves, so an
  // exchange shrinks to its barrier; the exact kernels leave the copy to bx_exchange_kernel
  GemmPush gp;
  gp.peers = m->ll.peers;
  gp.rank = uint32_t(m->rank);
  GemmPush* push = sh ? &gp : nullptr;
  if (const char* e = getenv("LLMI_NO_GEMM_PUSH")) if (e[0] == '1') push = nullptr;
  auto pushed = [&]() { return push && push->done; };
  // sharded: the main attention kernel runs this rank's KV heads only (1 / world of the attention arithmetic), the
  // heads' output columns travel and the quantizer runs on the complete batch.  Throughput mode (q_local): the
  // prologue skips the other ranks' query heads too, so q never leaves the rank that computed it.
  bool own_heads = sh && m->HK % uint32_t(m->world) == 0;
  if (own_heads) {  // (the q rows this rank computes must be exactly its heads')
    const uint64_t per = uint64_t(HD) / uint32_t(m->world);
    for (const LayerW& lw : m->layers)
      own_heads = own_heads && lw.q->row_begin == per * uint32_t(m->rank) && lw.q->row_end == per * (uint32_t(m->rank) + 1);
  }
  if (const char* e = getenv("LLMI_NO_HEAD_SHARD")) own_heads = own_heads && !(e[0] == '1');
  // (throughput mode: only the tensor-core attention kernel takes a head range, its CUDA-core fallback does not)
  if (llmi_gemv_prefill_fast() && !llmi_attention_batch_tc(m->H, m->HK, m->D)) own_heads = false;
  const bool q_local = own_heads && llmi_attention_batch_tc(m->H, m->HK, m->D);
  if (q_local) {
    gp.skip_begin = m->q;
    gp.skip_end = m->q + size_t(m->batch) * HD;
  }
  // sharded: the norm stages (post-norm + residual + next norm + quantizer: one CTA or cluster per token) run on a
  // contiguous token slice per rank — the residual batch stays token-sliced, nobody else reads it — and the slices of
  // the quantized activation (about one byte per element) are all-gathered for the mat-vecs that follow
  const uint32_t tok_per = sh ? (n_tok + uint32_t(m->world) - 1) / uint32_t(m->world) : n_tok;
  const uint32_t t0 = sh ? std::min(n_tok, tok_per * uint32_t(m->rank)) : 0, t1 = std::min(n_tok, t0 + tok_per);
  bool seq_par = sh && n_tok >= 8u * uint32_t(m->world) && tok_per * uint32_t(m->world - 1) < n_tok;
  {
    // a norm stage costs ~48 ns per token, the extra exchange (copy of the slice + flag barrier) ~12 us: the slice pays
    // from ~250 tokens taken off a rank (measured at N = 2: 2048-token batches 118.2 -> 116.3 ms, 256-token batches of
    // the exact mode 1410 -> 1425 ms).  LLMI_SEQ_NORM_MIN_TOKENS overrides the threshold (tests: 1; 0 = never).
    uint32_t min_off = 512;
    if (const char* e = getenv("LLMI_SEQ_NORM_MIN_TOKENS")) min_off = uint32_t(std::max(0, atoi(e)));
    seq_par = seq_par && min_off > 0 && n_tok - tok_per >= min_off;  // (the same decision on every rank)
  }
  auto in_comm = [&](const void* p) {
    const char* c = static_cast<const char*>(p);
    return c >= reinterpret_cast<const char*>(m->comm) && c < reinterpret_cast<const char*>(m->comm + m->comm_elems);
  };
  auto norm_stage = [&](NormArgs na) -> int {
    if (na.act_kind != ACT_Q8_K) na.xn_out = nullptr;  // nobody reads the fp32 copy of a batch (Q8_K: the cluster kernel's staging)
    const bool has_act = na.act_buf && na.act_kind != ACT_NONE;
    if (!seq_par || (has_act && !in_comm(na.act_buf))) {
      M_TRY(llmi_launch_norm_act(na, s));
      m->prefill_launches++;
      return LLMI_OK;
    }
    NormArgs sl = na;
    if (sl.y) sl.y += size_t(t0) * E;
    sl.h += size_t(t0) * E;
    if (sl.xn_out) sl.xn_out += size_t(t0) * E;
    if (sl.act_buf) sl.act_buf += size_t(t0) * sl.act_stride;
    sl.n_tok = t1 - t0;
    sl.pos_inc_by = n_tok;
    M_TRY(llmi_launch_norm_act(sl, s));
    m->prefill_launches++;
    if (has_act)  // the slice is one contiguous byte range of the activation batch
      M_RC(bx_exchange(m, {{reinterpret_cast<const float*>(na.act_buf), 0, nullptr, t0 * (na.act_stride / 4), (t1 - t0) * (na.act_stride / 4)}}, 1));
    return LLMI_OK;
  };
  if (sh) {
    // every rank has left its previous call (nobody still reads the residual batch), then the owners of the tokens'
    // embedding rows write them into every rank's batch
    M_RC(bx_exchange(m, {}, n_tok));
    M_TRY(llmi_launch_embed_shard_batch(make_embed_args(*m->embd), toks, std::sqrt(float(E)), m->ll.peers, m->bx_off(m->h),
                                        n_tok, s));
    m->prefill_launches++;
    M_RC(bx_exchange(m, {}, n_tok));
  } else {
    M_TRY(llmi_launch_embed(make_embed_args(*m->embd), toks, std::sqrt(float(E)), m->h, s, n_tok));
    m->prefill_launches++;
  }
  for (uint32_t l = 0; l < m->L; ++l) {
    LayerW& w = m->layers[l];
    const int kq = llmi_act_kind_for(w.q->type);
    if (l == 0) {
      NormArgs na;
      na.h = m->h; na.w = w.attn_norm; na.n = E; na.eps = m->eps; na.xn_out = m->xn;
      na.act_kind = kq; na.act_buf = get_bact(m, m->bact_E, kq, E); na.n_tok = n_tok;
      na.act_stride = uint32_t(act_bytes(kq, E));
      M_RC(norm_stage(na));
    }
    M_RC(gemv_tokens_group(m, {w.q, w.k, w.v}, {m->q, m->k, m->v}, {HD, KD, KD}, m->bact_E, E, n_tok, push));
    if (sh && pushed()) M_RC(bx_exchange(m, {}, n_tok));
    else if (sh && q_local) M_RC(bx_exchange(m, {{m->k, KD, w.k}, {m->v, KD, w.v}}, n_tok));
    else if (sh) M_RC(bx_exchange(m, {{m->q, HD, w.q}, {m->k, KD, w.k}, {m->v, KD, w.v}}, n_tok));
    AttnArgs aa;
    aa.q = m->q; aa.k = m->k; aa.v = m->v; aa.wq_norm = w.q_norm; aa.wk_norm = w.k_norm;
    aa.kcache = m->kcache + size_t(l) * m->t_max * m->HK * m->D;
    aa.vcache = m->vcache + size_t(l) * m->t_max * m->HK * m->D;
    aa.H = m->H; aa.HK = m->HK; aa.D = m->D; aa.t_max = m->t_max; aa.eps = m->eps;
    aa.rope_table = w.swa ? m->rope_swa : m->rope_global;
    aa.attn_scale = m->attn_scale; aa.pos = m->d_pos;
    aa.softcap = m->attn_softcap; aa.out = m->attn;
    const int ko = llmi_act_kind_for(w.o->type);
    aa.act_kind = ko; aa.act_buf = get_bact(m, m->bact_HD, ko, HD); aa.act_stride = uint32_t(act_bytes(ko, HD));
    aa.qbuf = m->qbuf;
    if (own_heads) {
      aa.hk_count = m->HK / uint32_t(m->world);
      aa.hk_begin = aa.hk_count * uint32_t(m->rank);
      aa.act_kind = ACT_NONE;
    }
    M_TRY(llmi_launch_attention(aa, s, n_tok));
    m->prefill_launches += 2;
    if (own_heads) {
      const uint32_t per = aa.hk_count * (m->H / m->HK) * m->D;
      M_RC(bx_exchange(m, {{m->attn, HD, nullptr, per * uint32_t(m->rank), per}}, n_tok));
      M_TRY(llmi_launch_act(m->attn, HD, ko, aa.act_buf, s, n_tok, aa.act_stride));
      m->prefill_launches++;
    }
    M_RC(gemv_tokens_group(m, {w.o}, {m->attn_out}, {E}, m->bact_HD, HD, n_tok, push));
    if (sh && pushed()) M_RC(bx_exchange(m, {}, n_tok));
    else if (sh) M_RC(bx_exchange(m, {{m->attn_out, E, w.o}}, n_tok));
    {
      const int kg = llmi_act_kind_for(w.gate->type);
      NormArgs na;
      na.y = m->attn_out; na.w_post = w.post_attn_norm; na.h = m->h; na.w = w.ffn_norm; na.n = E; na.eps = m->eps;
      na.xn_out = m->xn; na.act_kind = kg; na.act_buf = get_bact(m, m->bact_E, kg, E); na.n_tok = n_tok;
      na.act_stride = uint32_t(act_bytes(kg, E));
      M_RC(norm_stage(na));
    }
    M_RC(gemv_tokens_group(m, {w.gate, w.up}, {m->gate, m->up}, {F, F}, m->bact_E, E, n_tok));
    const int kd = llmi_act_kind_for(w.down->type);
    const bool fast_down = llmi_gemv_prefill_fast() && n_tok >= 64 && F % 64 == 0;
    // sharded: gate and up hold the same columns on a rank (same shape, same partition), so the rank combines them
    // and only the hidden columns travel — half the exchange, and the GEGLU arithmetic is split over the ranks
    const bool own_geglu = sh && w.gate->row_begin == w.up->row_begin && w.gate->row_end == w.up->row_end;
    // (throughput mode, stores into the peers enabled: the hidden values travel as the bf16 operand they become)
    void* hid16 = own_geglu && fast_down && push ? m->hid16 : nullptr;
    if (own_geglu) {
      M_TRY(llmi_launch_geglu_cols(m->gate, m->up, F, uint32_t(w.gate->row_begin), uint32_t(w.gate->row_end - w.gate->row_begin),
                                   n_tok, fast_down, s, push, hid16));
      m->prefill_launches++;
      if (push) M_RC(bx_exchange(m, {}, n_tok));
      else M_RC(bx_exchange(m, {{m->gate, F, w.gate}}, n_tok));
    } else if (sh) {
      M_RC(bx_exchange(m, {{m->gate, F, w.gate}, {m->up, F, w.up}}, n_tok));
    }
    if (push) push->done = false;
    if (fast_down) {  // throughput mode: GEGLU straight into ffn_down's operand
      M_TRY(llmi_launch_fast_ffn_down(*w.down, m->gate, own_geglu ? nullptr : m->up, m->ffn_out, E, n_tok, s, push, hid16));
      m->prefill_launches += 3;
    } else {
      if (own_geglu) {  // the hidden batch is complete: the quantizer alone
        M_TRY(llmi_launch_act(m->gate, F, kd, get_bact(m, m->bact_F, kd, F), s, n_tok, uint32_t(act_bytes(kd, F))));
      } else {
        M_TRY(llmi_launch_geglu_act(m->gate, m->up, F, kd, get_bact(m, m->bact_F, kd, F), nullptr, s, n_tok,
                                    uint32_t(act_bytes(kd, F))));
      }
      m->prefill_launches++;
      M_RC(gemv_tokens_group(m, {w.down}, {m->ffn_out}, {E}, m->bact_F, F, n_tok, push));
    }
    if (sh && pushed()) M_RC(bx_exchange(m, {}, n_tok));
    else if (sh) M_RC(bx_exchange(m, {{m->ffn_out, E, w.down}}, n_tok));
    {
      NormArgs na;
      na.y = m->ffn_out; na.w_post = w.post_ffw_norm; na.h = m->h; na.n = E; na.eps = m->eps; na.xn_out = m->xn;
      na.n_tok = n_tok;
      if (l + 1 < m->L) {
        const int kn = llmi_act_kind_for(m->layers[l + 1].q->type);
        na.w = m->layers[l + 1].attn_norm; na.act_kind = kn; na.act_buf = get_bact(m, m->bact_E, kn, E);
        na.act_stride = uint32_t(act_bytes(kn, E));
      } else {
        na.pos_inc = m->d_pos;  // the batch is done: += n_tok
        if (want_logits) {
          const int kl = llmi_act_kind_for(m->embd->type);
          na.w = m->out_norm; na.act_kind = kl; na.act_buf = get_bact(m, m->bact_E, kl, E);
          na.act_stride = uint32_t(act_bytes(kl, E));
        }
      }
      M_RC(norm_stage(na));
    }
  }
  if (want_logits) {  // logits of the last token of the batch: the ordinary one-token mat-vec
    const int kl = llmi_act_kind_for(m->embd->type);
    llmi_act_s la;
    la.kind = kl;
    la.n = E;
    la.buf = m->bact_E[kl] + size_t(n_tok - 1) * act_bytes(kl, E);
    M_TRY(llmi_launch_gemv(*m->embd, la, m->logits, s));
    m->prefill_launches++;
    if (sh) M_RC(bx_exchange(m, {{m->logits, m->V, m->embd}}, 1));
    if (m->final_softcap > 0.0f) {
      M_TRY(llmi_launch_softcap(m->logits, m->V, m->final_softcap, s));
      m->prefill_launches++;
    }
  }
  return LLMI_OK;
}


// ---- persistent decode kernel: plan + descriptors (mega.h) ----------------------------------------------------
struct MegaOffsets {
  uint32_t q[2], k[2], v[2], attn[2], ao[2], hid[2], fo[2];
};

uint32_t slab_bytes(uint32_t type, uint64_t nb, int plane) {  // bytes of one slab of a plane (repack.cu llmi_plan_planes)
  uint32_t q = 0, d = 0, x = 0;
  switch (type) {
    case LLMI_Q4_0: q = 16; d = 2; break;
    case LLMI_Q8_0: q = 32; d = 2; break;
    case LLMI_Q5_0: q = 16; d = 2; x = 4; break;
    case LLMI_Q4_K: q = 128; x = 16; break;
    case LLMI_Q6_K: q = 192; d = 2; x = 16; break;
    case LLMI_F16:
    case LLMI_BF16: q = 16; break;
    default: break;
  }
  return uint32_t(nb) * LLMI_SLAB * (plane == 0 ? q : (plane == 1 ? d : x));
}

// Decides whether this model runs on the persistent kernel and reserves its region of the exchange buffer
// (element offsets from `off` on).  Conditions: what norm_act_kernel supports (E <= 6144), one activation kind per
// consumer group, gate and up of one format, shared memory for the activations + the attention scratch.
int mega_plan(llmi_model_s* m, uint32_t& off) {
  const uint32_t E = m->E, F = m->F, HD = m->H * m->D, KD = m->HK * m->D;
  MegaArgs& a = m->mega;
  a = MegaArgs();
  // LLMI_DECODE = mega | legacy.  Default: the per-launch path; the persistent kernel is bit-identical but (round 2
  // measurements, profiles/r02_notes.md) not yet faster on one GPU
  m->use_mega = false;
  if (const char* e = getenv("LLMI_DECODE")) m->use_mega = std::string(e) == "mega";
  if (E > 6144 || llmi_mega_max_ctas() == 0) m->use_mega = false;
  for (const LayerW& w : m->layers) {
    const int kq = llmi_act_kind_for(w.q->type);
    if (llmi_act_kind_for(w.k->type) != kq || llmi_act_kind_for(w.v->type) != kq) m->use_mega = false;
    if (w.gate->type != w.up->type) m->use_mega = false;
    if (w.gate->n_slabs != w.up->n_slabs) m->use_mega = false;
  }
  if (!m->use_mega) return LLMI_OK;
  auto take = [&](uint32_t n) { const uint32_t o = off; off += (n + 1u) & ~1u; return o; };
  a.off_h = take(E);
  a.off_key = take(2 * LLMI_MAX_WORLD);
  a.off_tok = take(2);
  a.off_logits = m->sharded() ? take(m->V) : 0;
  off = (off + MEGA_CNT_STRIDE - 1) / MEGA_CNT_STRIDE * MEGA_CNT_STRIDE;  // counters on their own 128-byte lines
  a.off_cnt = take(MEGA_CNT_SLOTS * MEGA_CNT_STRIDE);
  // per-layer vectors twice (layer parity): a CTA that lags a whole layer behind still finds its inputs intact
  static_assert(sizeof(MegaOffsets) == 14 * 4, "MegaOffsets");
  MegaOffsets o;
  for (int p = 0; p < 2; ++p) {
    o.q[p] = take(HD); o.k[p] = take(KD); o.v[p] = take(KD); o.attn[p] = take(HD); o.ao[p] = take(E);
    o.hid[p] = take(F); o.fo[p] = take(E);
  }
  m->mega_off.assign(reinterpret_cast<uint32_t*>(&o), reinterpret_cast<uint32_t*>(&o) + 14);
  return LLMI_OK;
}

int mega_fill_mat(MegaPhase& P, int i, llmi_weight_t w, uint32_t out_off) {
  P.m[i] = llmi_gemv_args(*w);
  P.slab_q[i] = slab_bytes(w->type, w->nb, 0);
  P.slab_d[i] = slab_bytes(w->type, w->nb, 1);
  P.slab_x[i] = slab_bytes(w->type, w->nb, 2);
  // a K-chunk (work item) is 16 blocks of 32, 2 super-blocks of 256 or 16 groups of 8 halves (gemv_bodies.cuh Body::C)
  const uint64_t per_chunk = (w->type == LLMI_Q4_K || w->type == LLMI_Q6_K) ? 2 : 16;
  P.chunk_q[i] = slab_bytes(w->type, per_chunk, 0);
  P.chunk_d[i] = slab_bytes(w->type, per_chunk, 1);
  P.chunk_x[i] = slab_bytes(w->type, per_chunk, 2);
  P.out_off[i] = out_off;
  return LLMI_OK;
}

// The program of one token (mega.h) and the kernel's static arguments.  Called at the end of load_impl.
int mega_build(llmi_model_s* m) {
  if (!m->use_mega) return LLMI_OK;
  const uint32_t E = m->E, F = m->F, HD = m->H * m->D, L = m->L;
  MegaArgs& a = m->mega;
  MegaOffsets o;
  memcpy(&o, m->mega_off.data(), sizeof(o));
  a.L = L; a.E = E; a.F = F; a.H = m->H; a.HK = m->HK; a.D = m->D; a.V = m->V; a.t_max = m->t_max;
  a.eps = m->eps;
  a.attn_scale = m->attn_scale; a.attn_softcap = m->attn_softcap; a.final_softcap = m->final_softcap;
  a.embed_scale = std::sqrt(float(E));
  a.embd_type = m->embd->type; a.embd_nb = m->embd->nb;
  a.embd_row_begin = uint32_t(m->embd->row_begin); a.embd_row_end = uint32_t(m->embd->row_end);
  a.embd_q = m->embd->p_q; a.embd_d = m->embd->p_d; a.embd_x = m->embd->p_x;
  a.rank = uint32_t(m->rank); a.world = uint32_t(m->world);
  a.err = m->d_llerr;
  a.tag_mul = MEGA_TAGS_PER_LAYER * L + 8;
  a.tag_h = MEGA_TAGS_PER_LAYER * L + 1; a.tag_key = a.tag_h + 1; a.tag_tok = a.tag_h + 2; a.tag_logits = a.tag_h + 3;
  // head-sharded attention: the q/k/v rows of this rank are exactly the rows of its own heads, so q/k/v never
  // leave the rank and only the heads' outputs travel; otherwise every rank runs every head
  bool head_sharded = m->sharded() && m->H % m->world == 0 && m->HK % m->world == 0;
  if (head_sharded) {
    uint64_t rb = 0, re = 0;
    M_RC(llmi_shard_range(HD, m->world, m->rank, &rb, &re));
    const uint64_t per = HD / m->world;
    if (rb != per * m->rank || re != rb + per) head_sharded = false;
    M_RC(llmi_shard_range(m->HK * m->D, m->world, m->rank, &rb, &re));
    const uint64_t perk = uint64_t(m->HK) * m->D / m->world;
    if (rb != perk * m->rank || re != rb + perk) head_sharded = false;
  }
  if (const char* e = getenv("LLMI_NO_HEAD_SHARD")) head_sharded = head_sharded && !(e[0] == '1');
  a.attn_push_all = head_sharded ? 1u : 0u;
  a.head_begin = head_sharded ? m->rank * (m->H / m->world) : 0u;
  a.head_end = head_sharded ? a.head_begin + m->H / m->world : m->H;
  a.logits = m->logits; a.key = m->d_key; a.done_ctr = m->d_done; a.d_tok = m->d_tok; a.d_pos = m->d_pos;

  m->mega_ctas = llmi_mega_max_ctas();  // the same on every rank: the arrival counts depend on it
  if (const char* e = getenv("LLMI_MEGA_CTAS")) m->mega_ctas = std::min<uint32_t>(m->mega_ctas, uint32_t(std::max(1, atoi(e))));
  a.hints = 1;
  if (const char* e = getenv("LLMI_MEGA_NO_HINTS")) a.hints = e[0] == '1' ? 0u : 1u;
  a.pf_mode = 1;
  if (const char* e = getenv("LLMI_MEGA_PF")) a.pf_mode = uint32_t(atoi(e));
  std::vector<MegaPhase> prog;
  std::vector<MegaAttn> attn(L);
  size_t act_max = 0;
  uint32_t jmax = 1;
  auto tag_of = [&](uint32_t l, uint32_t i) { return 1 + MEGA_TAGS_PER_LAYER * l + i; };  // i: qkv, attn, ao, hid, fo
  // one mat-vec entry: the matrices of `ws` that share the format of ws[first]
  // exchange i (qkv, attn, ao, hid, fo) of layer l: its counter slot; bumps per use of an exchange
  auto slot_of = [&](uint32_t l, uint32_t i) { return MEGA_CNT_LAYER + (l & 1u) * MEGA_TAGS_PER_LAYER + i; };
  auto arrivals = [&](bool all) { return m->mega_ctas * (all ? uint32_t(m->world) : 1u); };
  auto add_gemv = [&](std::vector<llmi_weight_t> ws, std::vector<uint32_t> offs, uint32_t K, uint32_t pro, uint32_t epi,
                      uint32_t in_off, uint32_t in_tag, uint32_t out_tag, bool push_all, const float* w_post,
                      const float* w_next, uint32_t in_slot, uint32_t in_layer, uint32_t in_arr, uint32_t out_slot) -> int {
    std::vector<bool> done(ws.size(), false);
    bool first = true;
    const size_t first_entry = prog.size();
    for (size_t i = 0; i < ws.size(); ++i) {
      if (done[i]) continue;
      MegaPhase P;
      memset(&P, 0, sizeof(P));
      P.kind = MEGA_GEMV;
      P.type = ws[i]->type;
      P.act_kind = uint32_t(llmi_act_kind_for(ws[i]->type));
      P.K = K;
      P.J = llmi_gemv_chunks(*ws[i]);
      P.pro = first ? pro : uint32_t(MEGA_PRO_REUSE);
      P.epi = epi;
      P.in_off = in_off; P.in_tag = in_tag; P.out_tag = out_tag; P.push_all = push_all ? 1u : 0u;
      P.w_post = w_post; P.w_next = w_next;
      P.in_slot = in_slot; P.in_layer = in_layer; P.in_arrivals = in_arr; P.out_slot = out_slot;
      int n = 0;
      uint32_t v_total = 0;
      for (size_t j = i; j < ws.size(); ++j)
        if (!done[j] && ws[j]->type == ws[i]->type && (epi == MEGA_EPI_GEGLU || n < MEGA_MAX_MATS)) {
          if (ws[j]->n_cols != K) return llmi_fail(LLMI_ERR_SIZE, "llmi_model_load: matrices of one stage differ in K");
          mega_fill_mat(P, n, ws[j], offs[j]);
          v_total += uint32_t(ws[j]->n_slabs);
          done[j] = true;
          ++n;
        }
      P.n_mats = uint32_t(n);
      P.v_total = epi == MEGA_EPI_GEGLU ? uint32_t(ws[i]->n_slabs) : v_total;
      act_max = std::max(act_max, act_bytes(int(P.act_kind), K));
      jmax = std::max(jmax, P.J);
      prog.push_back(P);
      first = false;
    }
    if (prog.size() > first_entry && epi != MEGA_EPI_LOGITS) prog.back().bump = 1;
    return LLMI_OK;
  };
  for (uint32_t l = 0; l < L; ++l) {
    const LayerW& w = m->layers[l];
    const int p = int(l & 1), pp = int((l + 1) & 1);  // this layer's copies, the previous layer's
    M_RC(add_gemv({w.q, w.k, w.v}, {o.q[p], o.k[p], o.v[p]}, E, l == 0 ? uint32_t(MEGA_PRO_FIRST) : uint32_t(MEGA_PRO_NORM),
                  MEGA_EPI_FLAG, l ? o.fo[pp] : 0u, l ? tag_of(l - 1, 4) : 0u, tag_of(l, 0), !head_sharded,
                  l ? m->layers[l - 1].post_ffw_norm : nullptr, w.attn_norm, l ? slot_of(l - 1, 4) : 0u, l ? l - 1 : 0u,
                  arrivals(true), slot_of(l, 0)));
    MegaPhase T;
    memset(&T, 0, sizeof(T));
    T.kind = MEGA_ATTN;
    T.layer = l;
    prog.push_back(T);
    MegaAttn& t = attn[l];
    t.q_norm = w.q_norm; t.k_norm = w.k_norm;
    t.kcache = m->kcache + size_t(l) * m->t_max * m->HK * m->D;
    t.vcache = m->vcache + size_t(l) * m->t_max * m->HK * m->D;
    t.rope = w.swa ? m->rope_swa : m->rope_global;
    t.off_q = o.q[p]; t.off_k = o.k[p]; t.off_v = o.v[p]; t.in_tag = tag_of(l, 0);
    t.off_out = o.attn[p]; t.out_tag = tag_of(l, 1);
    M_RC(add_gemv({w.o}, {o.ao[p]}, HD, MEGA_PRO_QUANT, MEGA_EPI_FLAG, o.attn[p], tag_of(l, 1), tag_of(l, 2), true,
                  nullptr, nullptr, slot_of(l, 1), l, arrivals(head_sharded), slot_of(l, 2)));
    M_RC(add_gemv({w.gate, w.up}, {o.hid[p], o.hid[p]}, E, MEGA_PRO_NORM, MEGA_EPI_GEGLU, o.ao[p], tag_of(l, 2),
                  tag_of(l, 3), true, w.post_attn_norm, w.ffn_norm, slot_of(l, 2), l, arrivals(true), slot_of(l, 3)));
    M_RC(add_gemv({w.down}, {o.fo[p]}, F, MEGA_PRO_QUANT, MEGA_EPI_FLAG, o.hid[p], tag_of(l, 3), tag_of(l, 4), true,
                  nullptr, nullptr, slot_of(l, 3), l, arrivals(true), slot_of(l, 4)));
  }
  M_RC(add_gemv({m->embd}, {a.off_logits}, E, MEGA_PRO_NORM, MEGA_EPI_LOGITS, o.fo[(L - 1) & 1], tag_of(L - 1, 4),
                a.tag_logits, true, m->layers[L - 1].post_ffw_norm, m->out_norm, slot_of(L - 1, 4), L - 1, arrivals(true),
                0u));
  // shared memory: [h][xs][activation][chunk partials], the attention scratch over everything but h
  auto up128 = [](size_t v) { return (v + 127) / 128 * 128; };
  uint32_t type_mask = 0;
  for (const MegaPhase& P : prog)
    if (P.kind == MEGA_GEMV) type_mask |= mega_type_bit(P.type);
  size_t limit = 0;
  m->mega_variant = llmi_mega_select(type_mask, m->D, &limit);
  if (m->mega_variant < 0) {
    m->use_mega = false;
    return LLMI_OK;
  }
  a.sm_h = 0;
  a.sm_xs = uint32_t(up128(size_t(E) * 4));
  a.sm_wp = a.sm_xs + uint32_t(up128(size_t(E) * 4));  // norm weights of the running entry (post-norm, next norm)
  a.sm_wn = a.sm_wp + uint32_t(up128(size_t(E) * 4));
  a.sm_act = a.sm_wn + uint32_t(up128(size_t(E) * 4));
  a.sm_part = a.sm_act + uint32_t(up128(act_max));
  a.sm_attn = a.sm_xs;
  size_t part_bytes = 32 * 1024;
  const size_t part_min = size_t(LLMI_SLAB) * jmax * 2 * 4;
  if (part_bytes < part_min) part_bytes = up128(part_min);
  size_t attn_bytes = 0;
  // K/V tiles of the attention ring: what fits, but no more than 3 — the rest of the 228 KB stays L1 (the kernel's
  // few spill slots and the descriptors live there)
  size_t attn_cap = 3;
  if (const char* e = getenv("LLMI_MEGA_NBUF")) attn_cap = size_t(std::max(2, atoi(e)));
  a.attn_nbuf = limit > a.sm_attn ? llmi_mega_attention_nbuf(m->t_max, m->D, limit - a.sm_attn, &attn_bytes) : 0;
  if (a.attn_nbuf > attn_cap) {
    attn_bytes -= (a.attn_nbuf - attn_cap) * 32768;
    a.attn_nbuf = uint32_t(attn_cap);
  }
  if (a.attn_nbuf == 0 || a.sm_part + part_bytes > limit) {
    m->use_mega = false;  // context or K too large for one CTA's shared memory: the per-launch path runs it
    return LLMI_OK;
  }
  a.part_floats = uint32_t(part_bytes / 4);
  m->mega_smem = std::max<size_t>(a.sm_part + part_bytes, a.sm_attn + attn_bytes);
  a.n_prog = uint32_t(prog.size());
  M_RC(dev_alloc(m, (void**)&m->d_prog, prog.size() * sizeof(MegaPhase)));
  M_RC(dev_alloc(m, (void**)&m->d_mattn, attn.size() * sizeof(MegaAttn)));
  M_TRY(cudaMemcpy(m->d_prog, prog.data(), prog.size() * sizeof(MegaPhase), cudaMemcpyHostToDevice));
  M_TRY(cudaMemcpy(m->d_mattn, attn.data(), attn.size() * sizeof(MegaAttn), cudaMemcpyHostToDevice));
  a.prog = m->d_prog;
  a.attn = m->d_mattn;
  return LLMI_OK;
}

// n_steps tokens in ONE launch.  tokens_dev: the token of every step (prompts) or null (first_token, then the
// argmax of the previous step).  logits_mode as in MegaArgs.
int mega_run(llmi_model_s* m, const int32_t* tokens_dev, int32_t first_token, int pos, uint32_t n_steps,
             uint32_t logits_mode) {
  if (m->sharded() && m->mega.peers.n != uint32_t(m->world))
    return llmi_fail(LLMI_ERR_STATE, "row-sharded model: llmi_model_comm_connect has not been called");
  MegaArgs a = m->mega;
  a.tokens = tokens_dev;
  a.first_token = first_token;
  a.pos0 = pos;
  a.n_steps = n_steps;
  a.logits_mode = logits_mode;
  a.gen = logits_mode == 2 ? m->d_gen : nullptr;
  a.epoch0 = m->mega_epoch;
  a.tok_uses0 = m->mega_tok_uses;
  m->mega_epoch += n_steps;
  if (logits_mode == 2) m->mega_tok_uses += n_steps;
  M_TRY(llmi_launch_mega(m->mega_variant, a, m->mega_ctas, m->mega_smem, m->stream));
  m->mega_launches++;
  if (m->sharded() && logits_mode == 1) {  // host-facing logits: the rows of every rank, soft-capped by their producers
    LLTag t;
    t.add = (a.epoch0 + n_steps - 1) * a.tag_mul + a.tag_logits;
    t.err = m->d_llerr;
    M_TRY(llmi_launch_ll_unpack(m->comm + a.off_logits, t, m->logits, m->V, 0.0f, m->stream));
    m->mega_launches++;
  }
  return LLMI_OK;
}

// A consumer that gave up waiting (launch.cuh ll_wait) leaves a sticky flag: the results of that call are garbage.
int check_exchange(llmi_model_s* m, const char* who) {
  uint32_t e = 0;
  M_TRY(cudaMemcpy(&e, m->d_llerr, 4, cudaMemcpyDeviceToHost));
  if (e)
    return llmi_fail(LLMI_ERR_STATE, std::string(who) + ": a kernel gave up waiting for exchanged rows (a peer rank died, "
                                                        "stalled for seconds, or ran different steps); call "
                                                        "llmi_model_comm_reset on every rank to continue");
  return LLMI_OK;
}

double kv_f(const llmi::GgufImage& g, const std::string& key, double dflt, bool* found = nullptr) {
  const llmi::GgufValue* v = g.find(key);
  if (found) *found = v != nullptr;
  if (!v) return dflt;
  return (v->type == 6 || v->type == 12) ? v->f : double(v->u);
}

int load_impl(llmi_model_s* m, const uint8_t* image, uint64_t size, uint32_t t_max, int world, int rank) {
  m->world = world;
  m->rank = rank;
  llmi::GgufImage g(image, size);
  const llmi::GgufValue* arch = g.find("general.architecture");
  if (!arch) return llmi_fail(LLMI_ERR_ARG, "Failed to find metadata key: general.architecture");
  const std::string a = arch->s;
  if (a != "gemma3")
    return llmi_fail(LLMI_ERR_TYPE, "llmi_model_load: only the gemma3 architecture runs device-resident (got '" + a +
                                        "'); use the ops.h drop-in with the reference model.cpp");
  for (const char* k : {".block_count", ".embedding_length", ".feed_forward_length", ".attention.head_count",
                        ".attention.head_count_kv", ".attention.layer_norm_rms_epsilon", ".rope.freq_base"})
    if (!g.find(a + k)) return llmi_fail(LLMI_ERR_ARG, "Failed to find metadata key: " + a + k);  // model.cpp:63-67
  m->L = uint32_t(kv_f(g, a + ".block_count", 0));
  m->E = uint32_t(kv_f(g, a + ".embedding_length", 0));
  m->F = uint32_t(kv_f(g, a + ".feed_forward_length", 0));
  m->H = uint32_t(kv_f(g, a + ".attention.head_count", 0));
  m->HK = uint32_t(kv_f(g, a + ".attention.head_count_kv", 0));
  m->eps = double(float(kv_f(g, a + ".attention.layer_norm_rms_epsilon", 0)));  // stored f32, widened (model.cpp:84)
  m->rope_base = float(kv_f(g, a + ".rope.freq_base", 0));
  m->rope_scale = 1.0f;  // model.cpp:92
  m->D = uint32_t(kv_f(g, a + ".attention.key_length", double(m->E / (m->H ? m->H : 1))));  // model.cpp:93-99
  const uint32_t Dv = uint32_t(kv_f(g, a + ".attention.value_length", double(m->D)));
  if (Dv != m->D || g.find(a + ".attention.key_length_swa") || g.find(a + ".attention.value_length_swa"))
    return llmi_fail(LLMI_ERR_TYPE, "llmi_model_load: per-layer / asymmetric head sizes are not supported");
  if (m->D % 64 || m->D > 512 || m->H == 0 || m->HK == 0 || m->H % m->HK)
    return llmi_fail(LLMI_ERR_SIZE, "llmi_model_load: head_dim must be a multiple of 64 and n_head a multiple of n_head_kv");
  if (kv_f(g, a + ".attention.max_alibi_bias", 0.0) > 0.0)
    return llmi_fail(LLMI_ERR_TYPE, "llmi_model_load: ALiBi is not supported");
  if (m->E > 6144 || m->E % 32)
    return llmi_fail(LLMI_ERR_SIZE, "llmi_model_load: embedding_length must be a multiple of 32 and at most 6144 (the "
                                    "norm stage holds 6 elements per thread of a 1024-thread CTA)");
  m->attn_scale = 1.0f / std::sqrt(float(m->D));  // model.cpp:117
  m->attn_softcap = float(kv_f(g, a + ".attention.logit_softcapping", 0.0));
  m->final_softcap = float(kv_f(g, a + ".attention.final_logit_softcapping", 0.0));  // key as in model.cpp:142
  std::vector<bool> swa_meta;
  if (const llmi::GgufValue* sw = g.find(a + ".attention.sliding_window_pattern"))
    if (sw->type == 9)
      for (const auto& e : sw->arr) swa_meta.push_back(e.u != 0);
  m->t_max = t_max;
  const uint32_t E = m->E, F = m->F, HD = m->H * m->D, KD = m->HK * m->D;

  const llmi::GgufTensor* te = g.tensor("token_embd.weight");
  if (!te || te->shape.size() != 2) return llmi_fail(LLMI_ERR_ARG, "llmi_model_load: missing tensor token_embd.weight");
  if (te->type != LLMI_F16 && te->type != LLMI_Q6_K && te->type != LLMI_Q8_0 && te->type != LLMI_Q5_0)
    return llmi_fail(LLMI_ERR_TYPE, "Error: embed_tokens: Unsupported token embedding tensor type: " +
                                        std::to_string(te->type));  // model.cpp:324-330
  m->V = uint32_t(te->shape[1]);
  M_RC(upload_matrix(m, te, E, m->V, &m->embd, "token_embd.weight"));
  M_RC(upload_f32(m, g.tensor("output_norm.weight"), E, &m->out_norm, "output_norm.weight"));
  m->layers.resize(m->L);
  for (uint32_t l = 0; l < m->L; ++l) {
    LayerW& w = m->layers[l];
    const std::string p = "blk." + std::to_string(l) + ".";
    auto T = [&](const char* n) { return g.tensor(p + n + ".weight"); };
    auto T2 = [&](const char* n, const char* alt) {
      const llmi::GgufTensor* t = g.tensor(p + n + ".weight");
      return t ? t : g.tensor(p + alt + ".weight");
    };
    w.swa = l < swa_meta.size() ? bool(swa_meta[l]) : (l % 6 < 5);  // model.cpp:723-729
    M_RC(upload_f32(m, T("attn_norm"), E, &w.attn_norm, "attn_norm"));
    M_RC(upload_f32(m, T("ffn_norm"), E, &w.ffn_norm, "ffn_norm"));
    M_RC(upload_f32(m, T("attn_q_norm"), m->D, &w.q_norm, "attn_q_norm"));
    M_RC(upload_f32(m, T("attn_k_norm"), m->D, &w.k_norm, "attn_k_norm"));
    if (const llmi::GgufTensor* t = T2("post_attention_norm", "attn_post_norm"))
      M_RC(upload_f32(m, t, E, &w.post_attn_norm, "post_attention_norm"));
    if (const llmi::GgufTensor* t = T2("post_ffw_norm", "ffn_post_norm"))
      M_RC(upload_f32(m, t, E, &w.post_ffw_norm, "post_ffw_norm"));
    M_RC(upload_matrix(m, T("attn_q"), E, HD, &w.q, "attn_q"));
    M_RC(upload_matrix(m, T("attn_k"), E, KD, &w.k, "attn_k"));
    M_RC(upload_matrix(m, T("attn_v"), E, KD, &w.v, "attn_v"));
    M_RC(upload_matrix(m, T("attn_output"), HD, E, &w.o, "attn_output"));
    M_RC(upload_matrix(m, T("ffn_gate"), E, F, &w.gate, "ffn_gate"));
    M_RC(upload_matrix(m, T("ffn_up"), E, F, &w.up, "ffn_up"));
    M_RC(upload_matrix(m, T("ffn_down"), F, E, &w.down, "ffn_down"));
  }
  M_RC(llmi_upload_wait());  // every matrix has gone through the staging pipeline and its repack kernel
  const size_t mx = std::max<size_t>({E, F, HD});
  m->batch = 256;  // >= 128 tokens per batch go through the tensor-core mat-vec (gemv.cu launch_tokens)
  // throughput prefill (gemm_bf16.cuh) dequantizes a matrix once per batch: larger batches amortize it
  if (llmi_gemv_prefill_fast()) m->batch = 2048;
  if (const char* e = getenv("LLMI_PREFILL_BATCH")) m->batch = uint32_t(std::max(1, atoi(e)));
  if (m->batch > t_max) m->batch = t_max;
  const size_t B = m->batch;
  if (!m->sharded()) {  // (sharded: these are regions of the exchange allocation, below)
    M_RC(dev_alloc(m, (void**)&m->h, B * E * 4));
    M_RC(dev_alloc(m, (void**)&m->q, B * HD * 4));
    M_RC(dev_alloc(m, (void**)&m->k, B * KD * 4));
    M_RC(dev_alloc(m, (void**)&m->v, B * KD * 4));
    M_RC(dev_alloc(m, (void**)&m->attn_out, B * E * 4));
    M_RC(dev_alloc(m, (void**)&m->gate, B * F * 4));
    M_RC(dev_alloc(m, (void**)&m->up, B * F * 4));
    M_RC(dev_alloc(m, (void**)&m->ffn_out, B * E * 4));
    M_RC(dev_alloc(m, (void**)&m->logits, size_t(m->V) * 4));
    M_RC(dev_alloc(m, (void**)&m->attn, B * HD * 4));
  }
  M_RC(dev_alloc(m, (void**)&m->h2, E * 4));
  M_RC(dev_alloc(m, (void**)&m->xn, B * mx * 4));
  M_RC(dev_alloc(m, (void**)&m->q_rot, HD * 4));
  M_RC(dev_alloc(m, (void**)&m->qbuf, B * HD * 4));
  // a batch needs every consumer of a vector to take the same activation kind (true for the uniform and the
  // Q4_K_M layouts; otherwise prompts go token by token) and attention's fused quantizer to apply
  for (const LayerW& w : m->layers) {
    const int kq = llmi_act_kind_for(w.q->type), kg = llmi_act_kind_for(w.gate->type), ko = llmi_act_kind_for(w.o->type);
    if (llmi_act_kind_for(w.k->type) != kq || llmi_act_kind_for(w.v->type) != kq || llmi_act_kind_for(w.up->type) != kg)
      m->prefill_ok = false;
    if (ko == ACT_Q8_K && m->D % 256 != 0) m->prefill_ok = false;
  }
  if (m->batch < 2) m->prefill_ok = false;
  if (const char* e = getenv("LLMI_NO_PREFILL")) m->prefill_ok = m->prefill_ok && !(e[0] == '1');
  if (const char* e = getenv("LLMI_FUSE_GEGLU")) m->fuse_geglu = e[0] == '1';
  if (const char* e = getenv("LLMI_FUSE_NORM")) m->fuse_norm = e[0] == '1';
  {
    // one exchange buffer, same layout on every rank (element = {value bits, tag}): the region of the per-launch
    // path (sharded models only), then the region of the persistent decode kernel
    uint32_t off = 0;
    auto take = [&](uint32_t n) { const uint32_t o = off; off += (n + 1u) & ~1u; return o; };
    if (m->sharded()) {
      m->off_h = take(E); m->off_q = take(HD); m->off_k = take(KD); m->off_v = take(KD); m->off_ao = take(E);
      m->off_gate = take(F); m->off_up = take(F); m->off_fo = take(E); m->off_key = take(2 * LLMI_MAX_WORLD);
      m->off_logits = take(m->V);
      m->off_bar = take(LLMI_MAX_WORLD);
      if (const char* e = getenv("LLMI_NO_SHARD_PREFILL")) m->prefill_ok = m->prefill_ok && !(e[0] == '1');
    }
    M_RC(mega_plan(m, off));
    // the fp32 batch buffers of a sharded model: regions of the same allocation, 256-byte aligned
    uint64_t reg[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}, bact_reg[5] = {0, 0, 0, 0, 0};
    if (m->sharded()) {
      const size_t bytes[11] = {B * E * 4, B * HD * 4, B * KD * 4, B * KD * 4, B * E * 4, B * F * 4, B * F * 4, B * E * 4,
                                size_t(m->V) * 4, B * HD * 4, B * F * 2};
      uint64_t o64 = off;
      for (int i = 0; i < 11; ++i) {
        o64 = (o64 + 31) / 32 * 32;
        reg[i] = o64;
        o64 += (bytes[i] + 7) / 8;
      }
      // quantized activations of the residual width: the norm stages of a batch run on a token slice per rank and
      // the slices are all-gathered (run_batch), so these batches live in the exchange allocation too
      bool used[5] = {false, false, false, false, false};
      used[llmi_act_kind_for(m->embd->type)] = true;
      for (const LayerW& w : m->layers) {
        used[llmi_act_kind_for(w.q->type)] = true;
        used[llmi_act_kind_for(w.gate->type)] = true;
      }
      for (int k = 1; k < 5; ++k)
        if (used[k]) {
          o64 = (o64 + 31) / 32 * 32;
          bact_reg[k] = o64;
          o64 += (B * act_bytes(k, E) + 7) / 8;
        }
      if (o64 > 0xffffffffull) return llmi_fail(LLMI_ERR_SIZE, "llmi_model_load_shard: exchange allocation too large");
      off = uint32_t(o64);
    }
    m->comm_elems = off;
    M_TRY(cudaMalloc((void**)&m->comm, size_t(off) * sizeof(uint2)));  // not in `owned`: exported through CUDA IPC
    M_TRY(cudaMemset(m->comm, 0, size_t(off) * sizeof(uint2)));
    if (m->sharded()) {
      float** dst[10] = {&m->h, &m->q, &m->k, &m->v, &m->attn_out, &m->gate, &m->up, &m->ffn_out, &m->logits, &m->attn};
      for (int i = 0; i < 10; ++i) *dst[i] = reinterpret_cast<float*>(m->comm + reg[i]);
      m->hid16 = m->comm + reg[10];
      for (int k = 1; k < 5; ++k)
        if (bact_reg[k]) m->bact_E[k] = reinterpret_cast<uint8_t*>(m->comm + bact_reg[k]);
      M_RC(dev_alloc(m, (void**)&m->d_bx_counter, 16));
      M_TRY(cudaMemset(m->d_bx_counter, 0, 16));
    }
    M_RC(dev_alloc(m, (void**)&m->d_epoch, 16));
    M_RC(dev_alloc(m, (void**)&m->d_llerr, 16));
    M_RC(dev_alloc(m, (void**)&m->d_done, 16));
    const uint32_t one = 1;
    M_TRY(cudaMemcpy(m->d_epoch, &one, 4, cudaMemcpyHostToDevice));
    M_TRY(cudaMemset(m->d_llerr, 0, 16));
    M_TRY(cudaMemset(m->d_done, 0, 16));
    {  // word 1 of the error flag: how long a consumer waits for exchanged rows before it gives up (launch.cuh)
      uint32_t ms = 4000;
      if (const char* e = getenv("LLMI_EXCHANGE_TIMEOUT_MS")) ms = uint32_t(std::max(1, atoi(e)));
      M_TRY(cudaMemcpy(m->d_llerr + 1, &ms, 4, cudaMemcpyHostToDevice));
    }
    m->ll.rank = uint32_t(rank);
    m->ll.tag.epoch = m->d_epoch;
    m->ll.tag.mul = 4 * m->L + 4;
    m->ll.tag.err = m->d_llerr;
    if (!m->sharded()) {  // a single GPU is a world of one: the kernel's "peers" are its own buffer
      m->mega.peers.n = 1;
      m->mega.peers.base[0] = m->comm;
    }
    M_TRY(cudaDeviceSynchronize());
  }
  const size_t kv_elems = size_t(m->L) * t_max * KD;
  M_RC(dev_alloc(m, (void**)&m->kcache, kv_elems * 4));
  M_RC(dev_alloc(m, (void**)&m->vcache, kv_elems * 2));
  M_RC(dev_alloc(m, (void**)&m->d_tok, 16));
  M_RC(dev_alloc(m, (void**)&m->d_pos, 16));
  M_RC(dev_alloc(m, (void**)&m->d_gen_count, 16));
  M_RC(dev_alloc(m, (void**)&m->d_key, 16));
  M_TRY(cudaMemset(m->d_key, 0, 16));
  m->gen_cap = t_max + 1;
  M_RC(dev_alloc(m, (void**)&m->d_gen, size_t(m->gen_cap) * 4));
  m->toks_cap = t_max;
  M_RC(dev_alloc(m, (void**)&m->d_toks, size_t(m->toks_cap) * 4));
  M_TRY(cudaMemset(m->d_pos, 0, 16));
  M_TRY(cudaMemset(m->d_gen_count, 0, 16));
  M_TRY(cudaMallocHost((void**)&m->logits_pinned, size_t(m->V) * 4));
  M_TRY(cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking));
  {
    bool on = false;  // LLMI_PF_MB="attn,norm1,geglu,norm2" megabytes (e.g. 48,24,24,24) turns the side-stream prefetch on
    if (const char* e = getenv("LLMI_PF_MB")) {
      unsigned a0 = 0, a1 = 0, a2 = 0, a3 = 0;
      const int k = sscanf(e, "%u,%u,%u,%u", &a0, &a1, &a2, &a3);
      if (k >= 1) {
        m->pf_mb[0] = a0; m->pf_mb[1] = k > 1 ? a1 : a0; m->pf_mb[2] = k > 2 ? a2 : a0; m->pf_mb[3] = k > 3 ? a3 : a0;
        on = m->pf_mb[0] + m->pf_mb[1] + m->pf_mb[2] + m->pf_mb[3] > 0;
      }
    }
    if (on) {
      M_TRY(cudaStreamCreateWithFlags(&m->pf_stream, cudaStreamNonBlocking));
      m->pf_events.resize(16);
      for (auto& ev : m->pf_events) M_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    }
  }
  M_TRY(cudaEventCreate(&m->ev0));
  M_TRY(cudaEventCreate(&m->ev1));
  if (llmi_attention_smem(t_max, m->D) == 0)
    return llmi_fail(LLMI_ERR_SIZE, "llmi_model_load: max_positions " + std::to_string(t_max) + " is too large: the attention "
                                    "stage keeps 17 bytes per position in shared memory next to its K/V tiles (limit "
                                    "about 9500 positions)");
  M_TRY(llmi_attention_init(t_max, m->D));
  M_RC(dev_alloc(m, (void**)&m->rope_swa, size_t(t_max) * (m->D / 2) * sizeof(float2)));
  M_RC(dev_alloc(m, (void**)&m->rope_global, size_t(t_max) * (m->D / 2) * sizeof(float2)));
  M_TRY(llmi_launch_rope_table(m->rope_swa, t_max, m->D, 10000.0f, m->rope_scale, m->stream));
  M_TRY(llmi_launch_rope_table(m->rope_global, t_max, m->D, m->rope_base, m->rope_scale, m->stream));
  M_TRY(cudaStreamSynchronize(m->stream));
  M_RC(mega_build(m));
  return LLMI_OK;
}

int ensure_decode_graph(llmi_model_s* m, int pos) {
  if (m->decode_graph) return LLMI_OK;
  // warm-up run creates every lazily-allocated activation buffer outside capture
  const int32_t zero[4] = {0, 0, 0, 0};
  int32_t saved_pos = 0, saved_cnt = 0;
  M_TRY(cudaMemcpy(&saved_pos, m->d_pos, 4, cudaMemcpyDeviceToHost));
  M_TRY(cudaMemcpy(&saved_cnt, m->d_gen_count, 4, cudaMemcpyDeviceToHost));
  M_TRY(cudaMemcpy(m->d_tok, zero, 4, cudaMemcpyHostToDevice));
  // warm up at the caller's (validated) position: its K/V row is the one the first real step rewrites anyway, whereas
  // the device counter may sit at t_max after a forward that filled the cache (row t_max is out of bounds)
  M_TRY(cudaMemcpy(m->d_pos, &pos, 4, cudaMemcpyHostToDevice));
  M_RC(run_step(m, m->d_tok, true, true));
  M_TRY(cudaStreamSynchronize(m->stream));
  M_TRY(cudaMemcpy(m->d_pos, &saved_pos, 4, cudaMemcpyHostToDevice));
  M_TRY(cudaMemcpy(m->d_gen_count, &saved_cnt, 4, cudaMemcpyHostToDevice));
  cudaGraph_t graph = nullptr;
  M_TRY(cudaStreamBeginCapture(m->stream, cudaStreamCaptureModeThreadLocal));
  const int rc = run_step(m, m->d_tok, true, true);
  cudaError_t e = cudaStreamEndCapture(m->stream, &graph);
  if (rc != LLMI_OK) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  M_TRY(e);
  e = cudaGraphInstantiate(&m->decode_graph, graph, 0);
  cudaGraphDestroy(graph);
  M_TRY(e);
  return LLMI_OK;
}

}  // namespace

extern "C" {

int llmi_shard_range(uint64_t n_rows, int world, int rank, uint64_t* row_begin, uint64_t* row_end) {
  if (!row_begin || !row_end) return llmi_fail(LLMI_ERR_ARG, "llmi_shard_range: null pointer");
  if (world < 1 || rank < 0 || rank >= world) return llmi_fail(LLMI_ERR_ARG, "llmi_shard_range: need 0 <= rank < world");
  const uint64_t units = (n_rows + LLMI_SLAB - 1) / LLMI_SLAB, W = uint64_t(world), r = uint64_t(rank);
  const uint64_t first = r * (units / W) + std::min(r, units % W), cnt = units / W + (r < units % W ? 1 : 0);
  *row_begin = std::min(n_rows, first * LLMI_SLAB);
  *row_end = std::min(n_rows, (first + cnt) * LLMI_SLAB);
  return LLMI_OK;
}

int llmi_model_load(const void* gguf_image, uint64_t size, uint32_t max_positions, llmi_model_t* out) {
  return llmi_model_load_shard(gguf_image, size, max_positions, 1, 0, out);
}

int llmi_model_load_shard(const void* gguf_image, uint64_t size, uint32_t max_positions, int world, int rank,
                          llmi_model_t* out) {
  if (!gguf_image || !out) return llmi_fail(LLMI_ERR_ARG, "llmi_model_load: null pointer");
  if (llmi_sm_count() == 0) return llmi_fail(LLMI_ERR_STATE, "llmi_init() has not been called");
  if (world < 1 || world > LLMI_MAX_WORLD || rank < 0 || rank >= world)
    return llmi_fail(LLMI_ERR_ARG, "llmi_model_load_shard: need 1 <= world <= 8 and 0 <= rank < world");
  if (max_positions == 0) max_positions = 4096;
  llmi_gemv_read_env();
  llmi_glue_read_env();
  std::unique_ptr<llmi_model_s> m(new llmi_model_s());
  int rc;
  try {
    rc = load_impl(m.get(), static_cast<const uint8_t*>(gguf_image), size, max_positions, world, rank);
  } catch (const std::exception& e) {
    rc = llmi_fail(LLMI_ERR_ARG, e.what());
  }
  if (rc != LLMI_OK) {
    llmi_model_free(m.release());
    return rc;
  }
  *out = m.release();
  return LLMI_OK;
}

int llmi_model_free(llmi_model_t m) {
  if (!m) return LLMI_OK;
  if (m->stream) cudaStreamSynchronize(m->stream);
  if (m->decode_graph) cudaGraphExecDestroy(m->decode_graph);
  for (auto& w : m->layers)
    for (llmi_weight_t h : {w.q, w.k, w.v, w.o, w.gate, w.up, w.down}) llmi_weight_free(h);
  llmi_weight_free(m->embd);
  for (ActSet* s : {&m->act_E, &m->act_HD, &m->act_F})
    for (llmi_act_t a : s->a) llmi_act_free(a);
  for (void* p : m->owned) cudaFree(p);
  for (void* p : m->ipc_opened) cudaIpcCloseMemHandle(p);
  if (m->comm) cudaFree(m->comm);
  if (m->logits_pinned) cudaFreeHost(m->logits_pinned);
  for (cudaEvent_t ev : m->pf_events) cudaEventDestroy(ev);
  if (m->pf_stream) cudaStreamDestroy(m->pf_stream);
  if (m->ev0) cudaEventDestroy(m->ev0);
  if (m->ev1) cudaEventDestroy(m->ev1);
  if (m->stream) {
    llmi_stream_scratch_release(m->stream);
    cudaStreamDestroy(m->stream);
  }
  delete m;
  return LLMI_OK;
}

// ---- row-sharded model: wiring the ranks' exchange buffers together ------------------------------------------
// One process per GPU: every rank exports its buffer as a CUDA IPC handle (64 bytes), the host plumbing
// (torch.distributed, MPI, a pipe ...) all-gathers the handles, and every rank maps its peers' buffers.  After
// that the ranks never call a collective again: the mat-vec epilogues store into the mapped buffers.
int llmi_model_comm_handle(llmi_model_t m, void* handle64) {
  if (!m || !handle64) return llmi_fail(LLMI_ERR_ARG, "llmi_model_comm_handle: null pointer");
  if (!m->sharded()) return llmi_fail(LLMI_ERR_STATE, "llmi_model_comm_handle: the model is not sharded");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
  cudaIpcMemHandle_t h;
  M_TRY(cudaIpcGetMemHandle(&h, m->comm));
  memcpy(handle64, &h, sizeof(h));
  return LLMI_OK;
}

int llmi_model_comm_connect(llmi_model_t m, const void* handles /* world x 64 bytes, rank order */) {
  if (!m || !handles) return llmi_fail(LLMI_ERR_ARG, "llmi_model_comm_connect: null pointer");
  if (!m->sharded()) return llmi_fail(LLMI_ERR_STATE, "llmi_model_comm_connect: the model is not sharded");
  if (m->ll.peers.n) return llmi_fail(LLMI_ERR_STATE, "llmi_model_comm_connect: already connected");
  for (int r = 0; r < m->world; ++r) {
    if (r == m->rank) {
      m->ll.peers.base[r] = m->comm;
      continue;
    }
    cudaIpcMemHandle_t h;
    memcpy(&h, static_cast<const uint8_t*>(handles) + size_t(r) * sizeof(h), sizeof(h));
    void* p = nullptr;
    M_TRY(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    m->ipc_opened.push_back(p);
    m->ll.peers.base[r] = static_cast<uint2*>(p);
  }
  m->ll.peers.n = uint32_t(m->world);
  m->mega.peers = m->ll.peers;
  return LLMI_OK;
}

// After a failed exchange (llmi_model_comm_error != 0; forward / decode_greedy returned LLMI_ERR_STATE): clears the
// sticky flag.  Every rank must call it, after all ranks have drained their streams, before the next step.
int llmi_model_comm_reset(llmi_model_t m) {
  if (!m) return llmi_fail(LLMI_ERR_ARG, "llmi_model_comm_reset: null model");
  M_TRY(cudaStreamSynchronize(m->stream));
  M_TRY(cudaMemset(m->d_llerr, 0, 4));  // (word 1 holds the time limit)
  M_TRY(cudaMemset(m->d_done, 0, 16));
  M_TRY(cudaMemset(m->d_key, 0, 16));
  if (m->use_mega) {  // the arrival counters restart with the step count (identically on every rank)
    M_TRY(cudaMemset(m->comm + m->mega.off_cnt, 0, size_t(MEGA_CNT_SLOTS) * MEGA_CNT_STRIDE * sizeof(uint2)));
    m->mega_epoch = 1;
    m->mega_tok_uses = 0;
  }
  return LLMI_OK;
}

// Unmaps the peers' buffers: every rank disconnects, the host barriers, then the ranks free their models (a buffer
// must not be freed while a peer still has it mapped).
int llmi_model_comm_disconnect(llmi_model_t m) {
  if (!m) return llmi_fail(LLMI_ERR_ARG, "llmi_model_comm_disconnect: null model");
  if (m->stream) M_TRY(cudaStreamSynchronize(m->stream));
  if (m->decode_graph) {  // the captured kernels hold the peers' addresses
    cudaGraphExecDestroy(m->decode_graph);
    m->decode_graph = nullptr;
  }
  for (void* p : m->ipc_opened) cudaIpcCloseMemHandle(p);
  m->ipc_opened.clear();
  m->ll.peers.n = 0;
  m->mega.peers.n = 0;
  return LLMI_OK;
}

// 1 when a kernel gave up waiting for a peer's rows (the peer died or the ranks ran different steps).
int llmi_model_comm_error(llmi_model_t m) {
  if (!m || !m->d_llerr) return 0;
  uint32_t e = 0;
  if (cudaMemcpy(&e, m->d_llerr, 4, cudaMemcpyDeviceToHost) != cudaSuccess) return 1;
  return int(e);
}

int llmi_model_info(llmi_model_t m, uint32_t* dims /*[8]: L,E,F,H,HK,D,V,t_max*/, uint64_t* weight_bytes) {
  if (!m) return llmi_fail(LLMI_ERR_ARG, "llmi_model_info: null model");
  if (dims) {
    const uint32_t v[8] = {m->L, m->E, m->F, m->H, m->HK, m->D, m->V, m->t_max};
    for (int i = 0; i < 8; ++i) dims[i] = v[i];
  }
  if (weight_bytes) *weight_bytes = m->weight_bytes;
  return LLMI_OK;
}

// Model::forward(tokens, pos) (model.h:91, model.cpp:706-1048): host token ids
// in, host logits of the LAST token out.  Synchronous.
int llmi_model_forward(llmi_model_t m, const int32_t* tokens, int n_tokens, int pos, float* logits_host) {
  if (!m || !tokens || !logits_host) return llmi_fail(LLMI_ERR_ARG, "llmi_model_forward: null pointer");
  if (n_tokens <= 0) return llmi_fail(LLMI_ERR_ARG, "llmi_model_forward: no tokens");
  if (pos < 0 || uint32_t(pos) + uint32_t(n_tokens) > m->t_max || uint32_t(n_tokens) > m->toks_cap)
    return llmi_fail(LLMI_ERR_SIZE, "llmi_model_forward: position exceeds the KV cache capacity");
  for (int i = 0; i < n_tokens; ++i)
    if (tokens[i] < 0 || uint32_t(tokens[i]) >= m->V) return llmi_fail(LLMI_ERR_ARG, "llmi_model_forward: bad token id");
  cudaStream_t s = m->stream;
  M_TRY(cudaMemcpyAsync(m->d_toks, tokens, size_t(n_tokens) * 4, cudaMemcpyHostToDevice, s));
  M_TRY(cudaMemcpyAsync(m->d_pos, &pos, 4, cudaMemcpyHostToDevice, s));
  m->prefill_launches = 0;
  m->mega_launches = 0;
  M_TRY(cudaEventRecord(m->ev0, s));
  if (m->use_mega && (n_tokens == 1 || !m->prefill_ok)) {
    // one token (a decode step through the reference-facing call), or a prompt of a model without the batched
    // path (row-sharded): all tokens in one launch of the persistent kernel, logits of the last one
    M_RC(mega_run(m, m->d_toks, 0, pos, uint32_t(n_tokens), 1));
    m->prefill_launches = m->mega_launches;
  } else if (m->prefill_ok && n_tokens > 1) {
    for (int t = 0; t < n_tokens; t += int(m->batch)) {
      const int nb = std::min(int(m->batch), n_tokens - t);
      M_RC(run_batch(m, m->d_toks + t, uint32_t(nb), t + nb == n_tokens));
    }
  } else {
    for (int t = 0; t < n_tokens; ++t) {
      M_RC(run_step(m, m->d_toks + t, t == n_tokens - 1, false));
      m->prefill_launches += m->launches_per_step;
    }
  }
  M_TRY(cudaEventRecord(m->ev1, s));
  M_TRY(cudaMemcpyAsync(m->logits_pinned, m->logits, size_t(m->V) * 4, cudaMemcpyDeviceToHost, s));
  M_TRY(cudaStreamSynchronize(s));
  M_RC(check_exchange(m, "llmi_model_forward"));
  memcpy(logits_host, m->logits_pinned, size_t(m->V) * 4);
  return LLMI_OK;
}

// Greedy decode entirely on the device: the argmax of step i is the token of
// step i+1 (main.cpp:172-221 without the printing).  first_token is consumed at
// position pos; out_tokens receives n_steps generated ids; ms_device (optional)
// the CUDA-event time of the n_steps graph launches.
int llmi_model_decode_greedy(llmi_model_t m, int32_t first_token, int pos, int n_steps, int32_t* out_tokens,
                             float* ms_device) {
  if (!m || !out_tokens) return llmi_fail(LLMI_ERR_ARG, "llmi_model_decode_greedy: null pointer");
  if (n_steps <= 0 || pos < 0 || uint32_t(pos) + uint32_t(n_steps) > m->t_max || uint32_t(n_steps) > m->gen_cap)
    return llmi_fail(LLMI_ERR_SIZE, "llmi_model_decode_greedy: position exceeds the KV cache capacity");
  if (first_token < 0 || uint32_t(first_token) >= m->V)
    return llmi_fail(LLMI_ERR_ARG, "llmi_model_decode_greedy: bad token id");
  cudaStream_t s = m->stream;
  if (m->use_mega) {  // all n_steps tokens in one launch of the persistent kernel
    m->mega_launches = 0;
    M_TRY(cudaEventRecord(m->ev0, s));
    M_RC(mega_run(m, nullptr, first_token, pos, uint32_t(n_steps), 2));
    M_TRY(cudaEventRecord(m->ev1, s));
    M_TRY(cudaMemcpyAsync(out_tokens, m->d_gen, size_t(n_steps) * 4, cudaMemcpyDeviceToHost, s));
    M_TRY(cudaStreamSynchronize(s));
    if (ms_device) M_TRY(cudaEventElapsedTime(ms_device, m->ev0, m->ev1));
    m->launches_per_step = 1;
    return check_exchange(m, "llmi_model_decode_greedy");
  }
  M_RC(ensure_decode_graph(m, pos));
  const int32_t zero = 0;
  M_TRY(cudaMemcpyAsync(m->d_tok, &first_token, 4, cudaMemcpyHostToDevice, s));
  M_TRY(cudaMemcpyAsync(m->d_pos, &pos, 4, cudaMemcpyHostToDevice, s));
  M_TRY(cudaMemcpyAsync(m->d_gen_count, &zero, 4, cudaMemcpyHostToDevice, s));
  M_TRY(cudaEventRecord(m->ev0, s));
  for (int i = 0; i < n_steps; ++i) M_TRY(cudaGraphLaunch(m->decode_graph, s));
  M_TRY(cudaEventRecord(m->ev1, s));
  M_TRY(cudaMemcpyAsync(out_tokens, m->d_gen, size_t(n_steps) * 4, cudaMemcpyDeviceToHost, s));
  M_TRY(cudaStreamSynchronize(s));
  if (ms_device) M_TRY(cudaEventElapsedTime(ms_device, m->ev0, m->ev1));
  return check_exchange(m, "llmi_model_decode_greedy");
}

// Logits of the last executed step (device -> host), e.g. after decode_greedy.
int llmi_model_last_logits(llmi_model_t m, float* logits_host) {
  if (!m || !logits_host) return llmi_fail(LLMI_ERR_ARG, "llmi_model_last_logits: null pointer");
  if (m->sharded())
    return llmi_fail(LLMI_ERR_STATE, "llmi_model_last_logits: a sharded greedy step keeps only this rank's rows (the "
                                     "ranks swap argmax keys); call llmi_model_forward for full logits");
  M_TRY(cudaMemcpyAsync(m->logits_pinned, m->logits, size_t(m->V) * 4, cudaMemcpyDeviceToHost, m->stream));
  M_TRY(cudaStreamSynchronize(m->stream));
  memcpy(logits_host, m->logits_pinned, size_t(m->V) * 4);
  return LLMI_OK;
}

int llmi_model_launches_per_step(llmi_model_t m) { return m ? m->launches_per_step : 0; }
int llmi_model_decode_path(llmi_model_t m) { return m && m->use_mega ? 1 : 0; }

int llmi_model_last_forward_stats(llmi_model_t m, float* ms_device, int* launches) {
  if (!m) return llmi_fail(LLMI_ERR_ARG, "llmi_model_last_forward_stats: null model");
  if (ms_device) M_TRY(cudaEventElapsedTime(ms_device, m->ev0, m->ev1));
  if (launches) *launches = m->prefill_launches;
  return LLMI_OK;
}

}  // extern "C"
