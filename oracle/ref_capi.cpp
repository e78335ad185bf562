// TEST INFRASTRUCTURE — not product code.
//
// extern "C" harness around the UNMODIFIED reference sources, which are compiled
// where they lie under /root/reference by oracle/Makefile (outputs only into
// oracle/_ref/).  Nothing from the reference is copied here: this file only
// *calls* the reference's public functions (ops.h, gguf.h, model.h) so that
// Python tests (ctypes) can use the reference's own CPU code as the parity
// oracle and as the CPU baseline (`cpu_baseline.kind == "reference"`).
//
// The same source is also compiled against the CUDA drop-in (host/ops_cuda.cpp)
// to produce oracle/_ref/libdropin.so: there, the very same calls
// (mat_vec_mul, Model::forward -> model.cpp:557,754,784,803,875,877,909,1000)
// run on the B200 path through the unchanged ops.h signatures.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the libraries built from this file.
#include <cstdint>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "gguf.h"
#include "model.h"
#include "ops.h"

bool verbose_g = false;  // common.h:7 — every embedding binary must define it

namespace {

thread_local std::string g_err;

// Minimal single-tensor GGUF image, following the layout the reference parser
// expects (gguf.cpp:281-303: header, tensor infos, data section aligned to 32).
// Same construction idea as ops_test.cpp:96-136 (create_minimal_gguf).
struct RefTensor {
  std::vector<uint8_t> image;
  std::unique_ptr<GGUFFile> file;
  const TensorInfo* info = nullptr;
  uint64_t n_cols = 0, n_rows = 0;
  std::vector<float> x, o;
};

template <typename T>
void put(std::vector<uint8_t>& b, const T& v) {
  const uint8_t* p = reinterpret_cast<const uint8_t*>(&v);
  b.insert(b.end(), p, p + sizeof(T));
}

struct RefModel {
  std::vector<uint8_t> image;  // owned copy when requested
  std::unique_ptr<GGUFFile> file;
  std::unique_ptr<Model> model;
};

}  // namespace

extern "C" {

const char* ref_last_error() { return g_err.c_str(); }

void ref_init_ops(int n_threads) { init_ops(n_threads); }  // ops.h:38

float ref_f16_to_f32(uint16_t h) { return f16_to_f32(h); }    // gguf.h:124
uint16_t ref_f32_to_f16(float f) { return f32_to_f16(f); }    // gguf.h:125
float ref_bf16_to_f32(uint16_t h) { return bf16_to_f32(h); }  // gguf.h:126

void ref_f16_to_f32_n(const uint16_t* h, float* f, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) f[i] = f16_to_f32(h[i]);
}
void ref_f32_to_f16_n(const float* f, uint16_t* h, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) h[i] = f32_to_f16(f[i]);
}

// ops.h:94 — out receives n/32 BlockQ8_0 records (34 bytes each).
int ref_quantize_row_q8_0(const float* x, uint64_t n, uint8_t* out) {
  try {
    std::vector<float> xv(x, x + n);
    std::vector<BlockQ8_0> y;
    quantize_row_q8_0(xv, y, n);
    static_assert(sizeof(BlockQ8_0) == 34, "BlockQ8_0 layout");
    memcpy(out, y.data(), y.size() * sizeof(BlockQ8_0));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ops.h:104 — out receives n/256 block_q8_K records (292 bytes each).
int ref_quantize_row_q8_k(const float* x, uint64_t n, uint8_t* out) {
  try {
    std::vector<float> xv(x, x + n);
    std::vector<block_q8_K> y;
    quantize_row_q8_k(xv, y, n);
    static_assert(sizeof(block_q8_K) == 292, "block_q8_K layout");
    memcpy(out, y.data(), y.size() * sizeof(block_q8_K));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Wrap raw weight blocks as a one-tensor in-memory GGUF so the reference's
// mat_vec_mul(o, TensorInfo, GGUFFile, x) can be called on them unchanged.
void* ref_tensor_create(uint32_t ggml_type, const uint8_t* blocks,
                        uint64_t n_bytes, uint64_t n_cols, uint64_t n_rows) {
  try {
    auto t = std::make_unique<RefTensor>();
    std::vector<uint8_t>& b = t->image;
    GGUFHeader h = {GGUF_MAGIC, GGUF_VERSION, 1, 0};
    put(b, h);
    const std::string name = "w";
    put<uint64_t>(b, name.size());
    b.insert(b.end(), name.begin(), name.end());
    put<uint32_t>(b, 2);
    put<uint64_t>(b, n_cols);  // shape[0] = K (ops.cpp:193-194)
    put<uint64_t>(b, n_rows);  // shape[1] = N
    put<uint32_t>(b, ggml_type);
    put<uint64_t>(b, 0);
    size_t start = (b.size() + 31) & ~size_t(31);  // gguf.cpp:301-303
    b.resize(start + n_bytes);
    memcpy(b.data() + start, blocks, n_bytes);
    t->file = std::make_unique<GGUFFile>(b.data(), b.size());
    t->info = &t->file->get_tensor_infos()[0];
    t->n_cols = n_cols;
    t->n_rows = n_rows;
    return t.release();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}

void ref_tensor_free(void* p) { delete static_cast<RefTensor*>(p); }

// ops.h:53 dispatcher (ops.cpp:933-956). Returns 1 + message on throw.
int ref_tensor_mat_vec_mul(void* p, const float* x, uint64_t n_x, float* o,
                           uint64_t n_o) {
  RefTensor* t = static_cast<RefTensor*>(p);
  try {
    t->x.assign(x, x + n_x);
    t->o.assign(n_o, 0.0f);
    mat_vec_mul(t->o, *t->info, *t->file, t->x);
    if (t->o.size() != n_o) {
      g_err = "ref_tensor_mat_vec_mul: output size " +
              std::to_string(t->o.size()) + " != " + std::to_string(n_o);
      return 2;
    }
    memcpy(o, t->o.data(), n_o * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Timed loop for the CPU baseline: `iters` calls on the same x, no copies
// inside the loop other than what the reference itself does.
int ref_tensor_mat_vec_mul_loop(void* p, const float* x, uint64_t n_x,
                                int iters) {
  RefTensor* t = static_cast<RefTensor*>(p);
  try {
    t->x.assign(x, x + n_x);
    t->o.assign(t->n_rows, 0.0f);
    for (int i = 0; i < iters; ++i) mat_vec_mul(t->o, *t->info, *t->file, t->x);
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ops.h:41
int ref_mat_vec_mul_fp16(const uint16_t* w, uint64_t n_w, const float* x,
                         uint64_t n_x, uint64_t n_rows, uint64_t n_cols,
                         float* o) {
  try {
    std::vector<uint16_t> wv(w, w + n_w);
    std::vector<float> xv(x, x + n_x), ov;
    mat_vec_mul_fp16(ov, wv, xv, n_rows, n_cols);
    memcpy(o, ov.data(), ov.size() * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// Persistent F16 matrix for timing (avoids the vector copy per call).
struct RefF16 {
  std::vector<uint16_t> w;
  std::vector<float> x, o;
  uint64_t n_rows, n_cols;
};
void* ref_f16_create(const uint16_t* w, uint64_t n_rows, uint64_t n_cols) {
  auto* t = new RefF16();
  t->w.assign(w, w + n_rows * n_cols);
  t->n_rows = n_rows;
  t->n_cols = n_cols;
  return t;
}
void ref_f16_free(void* p) { delete static_cast<RefF16*>(p); }
int ref_f16_mat_vec_mul_loop(void* p, const float* x, float* o, int iters) {
  RefF16* t = static_cast<RefF16*>(p);
  try {
    t->x.assign(x, x + t->n_cols);
    for (int i = 0; i < iters; ++i)
      mat_vec_mul_fp16(t->o, t->w, t->x, t->n_rows, t->n_cols);
    if (o) memcpy(o, t->o.data(), t->o.size() * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ops.h:72-79 row dequantizers (used by Model::embed_tokens, model.cpp:288-322).
int ref_dequantize_row(uint32_t ggml_type, const uint8_t* blocks,
                       uint64_t n_cols, float* out) {
  try {
    std::vector<float> o;
    switch (ggml_type) {
      case 12: dequantize_q4_k_row(o, blocks, n_cols); break;
      case 14: dequantize_q6_k_row(o, blocks, n_cols); break;
      case 8: dequantize_q8_0_row(o, blocks, n_cols); break;
      case 6: dequantize_q5_0_row(o, blocks, n_cols); break;
      default:
        g_err = "ref_dequantize_row: unsupported type";
        return 2;
    }
    memcpy(out, o.data(), o.size() * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

// ops.h:82-86 — the small non-matmul ops, for glue-kernel parity.
void ref_rms_norm(const float* x, uint64_t n, double eps, float* o) {
  std::vector<float> xv(x, x + n), ov(n);
  rms_norm(ov, xv, eps);
  memcpy(o, ov.data(), n * sizeof(float));
}
void ref_softmax(float* x, uint64_t n) {
  std::vector<float> xv(x, x + n);
  softmax(xv);
  memcpy(x, xv.data(), n * sizeof(float));
}
// tensor is [n_tokens][n_heads][head_dim] flattened.
void ref_rope(float* t, uint32_t n_tokens, uint32_t n_heads, uint32_t head_dim,
              int n_rot, float freq_base, float freq_scale, int pos) {
  tensor_3 v(n_tokens, tensor_2(n_heads, tensor_1(head_dim)));
  for (uint32_t a = 0; a < n_tokens; ++a)
    for (uint32_t b = 0; b < n_heads; ++b)
      for (uint32_t c = 0; c < head_dim; ++c)
        v[a][b][c] = t[(a * n_heads + b) * head_dim + c];
  rope(v, n_rot, freq_base, freq_scale, pos);
  for (uint32_t a = 0; a < n_tokens; ++a)
    for (uint32_t b = 0; b < n_heads; ++b)
      for (uint32_t c = 0; c < head_dim; ++c)
        t[(a * n_heads + b) * head_dim + c] = v[a][b][c];
}

// Model (model.h:72-117). `copy` != 0 keeps a private copy of the image so the
// caller may free its buffer (GGUFFile borrows the bytes, gguf.cpp:265-270).
void* ref_model_create(const uint8_t* gguf, uint64_t size, int copy) {
  try {
    auto m = std::make_unique<RefModel>();
    const uint8_t* base = gguf;
    if (copy) {
      m->image.assign(gguf, gguf + size);
      base = m->image.data();
    }
    m->file = std::make_unique<GGUFFile>(base, size);
    m->model = std::make_unique<Model>(*m->file);
    return m.release();
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void ref_model_free(void* p) { delete static_cast<RefModel*>(p); }

// Model::forward(tokens, pos) -> logits of the last token (model.cpp:706-1048).
int ref_model_forward(void* p, const int32_t* tokens, int n_tokens, int pos,
                      float* logits, uint64_t n_logits) {
  RefModel* m = static_cast<RefModel*>(p);
  try {
    std::vector<int> tk(tokens, tokens + n_tokens);
    auto r = m->model->forward(tk, pos);
    if (r.size() != 1 || r[0].size() != n_logits) {
      g_err = "ref_model_forward: unexpected logits shape";
      return 2;
    }
    memcpy(logits, r[0].data(), n_logits * sizeof(float));
    return 0;
  } catch (const std::exception& e) {
    g_err = e.what();
    return 1;
  }
}

uint64_t ref_model_vocab(void* p) {
  return static_cast<RefModel*>(p)->model->token_embd_weight()->shape[1];
}

}  // extern "C"
