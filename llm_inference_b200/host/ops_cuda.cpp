// ops_cuda.cpp — drop-in replacement for the matmul half of the reference's
// ops.cpp.  It defines the SAME C++ symbols ops.h declares for the quantized
// mat-vec path, with the same signatures, ownership and error behaviour, and
// forwards to the C ABI of libllmi_cuda.so (include/llmi_cuda.h).  model.cpp
// and the reference's *_test.cpp files compile and link against it unmodified
// (oracle/Makefile `dropin` target; INTEGRATION.md shows the Bazel stanza).
//
// Compiled against the reference's own headers (-I<reference>), never a copy.
//
// Replaced (ops.h line -> reference body):
//   init_ops              :38  ops.cpp:21-24    thread pool -> CUDA context
//   mat_vec_mul           :53  ops.cpp:933-956  dispatcher, same if/else order
//   mat_vec_mul_q4_0      :56  ops.cpp:188-451
//   mat_vec_mul_q4_k      :60  ops.cpp:614-706
//   mat_vec_mul_q6_k      :63  ops.cpp:708-785
//   mat_vec_mul_q8_0      :65  ops.cpp:787-838
//   mat_vec_mul_q5_0      :67  ops.cpp:840-893
//   mat_vec_mul_bf16      :69  ops.cpp:895-931
//   mat_vec_mul_fp16      :41  ops.cpp:455-612
//   quantize_row_q8_0     :94  ops.cpp:116-139
//   quantize_row_q8_k     :104 ops.cpp:142-178
// Not replaced (stay the reference's CPU bodies): rms_norm, softmax, rope,
// scale, vec_scale_f16, vec_mad_f16, dequantize_*_row.
//
// Contract kept from the reference: `o` is caller-owned and resize()d to
// n_rows (ops.cpp:200); `x` is borrowed; weights are borrowed from the GGUF
// image for its lifetime (here: uploaded + repacked on first use, keyed by the
// host pointer get_tensor_data returns); calls are synchronous; errors are
// std::runtime_error with the reference's strings; there is no CPU fallback —
// a missing GPU surfaces as a runtime_error from init_ops().
#include <cstdlib>
#include <iostream>
#include <stdexcept>
#include <string>
#include <vector>

#include "gguf.h"
#include "llmi_cuda.h"
#include "ops.h"

namespace {

void must(int rc) {
  if (rc != LLMI_OK) throw std::runtime_error(llmi_last_error());
}

// One reference mat_vec_mul_<fmt> call: size check, resize, registry lookup
// (upload + repack on first use), host-vector mat-vec.
void run_typed(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
               const std::vector<float>& x, const char* fn, uint32_t ggml_type) {
  const size_t n_rows = w_tensor.shape[1];  // ops.cpp:193
  const size_t n_cols = w_tensor.shape[0];  // ops.cpp:194
  if (x.size() != n_cols) {
    throw std::runtime_error(std::string(fn) + ": input vector size mismatch");
  }
  o.resize(n_rows);
  const uint8_t* w_data = gguf_file.get_tensor_data(w_tensor);  // ops.cpp:206
  llmi_weight_t w = nullptr;
  must(llmi_registry_get(w_data, ggml_type, n_cols, n_rows, &w));
  must(llmi_host_mat_vec_mul(w, x.data(), x.size(), o.data(), o.size()));
}

}  // namespace

void init_ops(int n_threads) {
  (void)n_threads;  // row partitioning is grid partitioning on the GPU
  int device = 0;
  if (const char* e = std::getenv("LLMI_DEVICE")) device = std::atoi(e);
  must(llmi_init(device));
}

void mat_vec_mul_q4_0(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                      const std::vector<float>& x) {
  run_typed(o, w_tensor, gguf_file, x, "mat_vec_mul_q4_0", LLMI_Q4_0);
}

void mat_vec_mul_q4_k(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                      const std::vector<float>& x) {
  run_typed(o, w_tensor, gguf_file, x, "mat_vec_mul_q4_k", LLMI_Q4_K);
}

void mat_vec_mul_q6_k(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                      const std::vector<float>& x) {
  run_typed(o, w_tensor, gguf_file, x, "mat_vec_mul_q6_k", LLMI_Q6_K);
}

void mat_vec_mul_q8_0(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                      const std::vector<float>& x) {
  run_typed(o, w_tensor, gguf_file, x, "mat_vec_mul_q8_0", LLMI_Q8_0);
}

void mat_vec_mul_q5_0(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                      const std::vector<float>& x) {
  run_typed(o, w_tensor, gguf_file, x, "mat_vec_mul_q5_0", LLMI_Q5_0);
}

void mat_vec_mul_bf16(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                      const std::vector<float>& x) {
  run_typed(o, w_tensor, gguf_file, x, "mat_vec_mul_bf16", LLMI_BF16);
}

void mat_vec_mul(std::vector<float>& o, const TensorInfo& w_tensor, const GGUFFile& gguf_file,
                 const std::vector<float>& x) {
  if (x.size() != w_tensor.shape[0]) {  // only logged here, the typed op throws (ops.cpp:935-939)
    std::cerr << "mat_vec_mul size mismatch: tensor: " << w_tensor.name
              << " w_tensor.shape[0]=" << w_tensor.shape[0] << " x.size()=" << x.size() << std::endl;
  }
  switch (static_cast<GGUFTensorType>(w_tensor.tensor_type)) {
    case GGUFTensorType::Q4_0: mat_vec_mul_q4_0(o, w_tensor, gguf_file, x); break;
    case GGUFTensorType::Q4_K: mat_vec_mul_q4_k(o, w_tensor, gguf_file, x); break;
    case GGUFTensorType::Q6_K: mat_vec_mul_q6_k(o, w_tensor, gguf_file, x); break;
    case GGUFTensorType::Q8_0: mat_vec_mul_q8_0(o, w_tensor, gguf_file, x); break;
    case GGUFTensorType::Q5_0: mat_vec_mul_q5_0(o, w_tensor, gguf_file, x); break;
    case GGUFTensorType::BF16: mat_vec_mul_bf16(o, w_tensor, gguf_file, x); break;
    default:
      throw std::runtime_error("mat_vec_mul: unsupported tensor type " + std::to_string(w_tensor.tensor_type));
  }
}

void mat_vec_mul_fp16(std::vector<float>& o, const std::vector<uint16_t>& w, const std::vector<float>& x,
                      size_t n_rows, size_t n_cols) {
  if (x.size() != n_cols) {
    throw std::runtime_error("mat_vec_mul_fp16: input vector size mismatch");
  }
  if (w.size() != n_rows * n_cols) {
    throw std::runtime_error("mat_vec_mul_fp16: weight matrix size mismatch");
  }
  o.resize(n_rows);
  llmi_weight_t h = nullptr;
  // keyed by the vector's storage: Model keeps token_embd_weight_f16_ alive (model.cpp:46-55)
  must(llmi_registry_get(w.data(), LLMI_F16, n_cols, n_rows, &h));
  must(llmi_host_mat_vec_mul(h, x.data(), x.size(), o.data(), o.size()));
}

void quantize_row_q8_0(const std::vector<float>& x, std::vector<BlockQ8_0>& y, size_t size) {
  static_assert(sizeof(BlockQ8_0) == 34, "BlockQ8_0 must match the 34-byte record of the C ABI");
  y.resize(size / 32);
  if (size / 32 == 0) return;
  must(llmi_host_quantize_row_q8_0(x.data(), size, y.data()));
}

void quantize_row_q8_k(const std::vector<float>& x, std::vector<block_q8_K>& y, size_t size) {
  static_assert(sizeof(block_q8_K) == 292, "block_q8_K must match the 292-byte record of the C ABI");
  y.resize(size / QK_K);
  if (size / QK_K == 0) return;
  must(llmi_host_quantize_row_q8_k(x.data(), size, y.data()));
}
