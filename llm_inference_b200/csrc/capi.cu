// C ABI of libllmi_cuda.so (include/llmi_cuda.h): context, weights, activations,
// the host-vector tier used by the ops.h drop-in, and small device helpers.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "glue.h"
#include "launch.cuh"
#include "mega.h"
#include "llmi_internal.h"

// Programmatic dependent launch for every kernel of the library (launch.cuh);
// LLMI_NO_PDL=1 turns it off for A/B measurements.
bool g_llmi_pdl = true;

namespace {

thread_local std::string t_err;

struct Context {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;  // stream of the host-vector tier
  // host-tier scratch, grown on demand
  float* x_dev = nullptr;
  size_t x_cap = 0;
  float* o_dev = nullptr;
  size_t o_cap = 0;
  llmi_act_t act = nullptr;
  uint8_t* scratch = nullptr;
  size_t scratch_cap = 0;
  uint8_t* tok_act = nullptr;  // quantized activations of a token batch (llmi_gemm_tokens)
  size_t tok_act_cap = 0;
  // registry of the ops.h drop-in: repacked weights keyed by (host pointer, type, K, N), each with a fingerprint of
  // the host bytes it was built from (see llmi_registry_get)
  struct RegEntry {
    llmi_weight_t w = nullptr;
    uint64_t fp = 0;
  };
  std::map<std::tuple<const void*, uint32_t, uint64_t, uint64_t>, RegEntry> registry;
  // upload pipeline: raw GGUF rows -> pinned staging -> device staging -> repack kernel, two slots
  // deep, on its own stream; allocated on first use, reused by every upload of the process
  cudaStream_t up_stream = nullptr;
  uint8_t* up_pinned[2] = {nullptr, nullptr};
  uint8_t* up_dev[2] = {nullptr, nullptr};
  cudaEvent_t up_done[2] = {nullptr, nullptr};
  size_t up_cap = 0;
  int up_next = 0;
};
Context g;
std::mutex g_mu;

bool supported(uint32_t t) {
  switch (t) {
    case LLMI_Q4_0: case LLMI_Q8_0: case LLMI_Q5_0: case LLMI_Q4_K: case LLMI_Q6_K: case LLMI_F16: case LLMI_BF16:
      return true;
    default: return false;
  }
}

int ensure_cap(void** p, size_t* cap, size_t need) {
  if (*cap >= need) return LLMI_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  const size_t want = round_up(need, 1 << 16);
  LLMI_CUDA_TRY(cudaMalloc(p, want));
  *cap = want;
  return LLMI_OK;
}

void upload_pipeline_free();  // defined with the upload pipeline below

#define LLMI_NEED_INIT() \
  if (!g.ready) return llmi_fail(LLMI_ERR_STATE, "llmi_init() has not been called")

}  // namespace

void llmi_set_error(const std::string& msg) { t_err = msg; }
int llmi_fail(int code, const std::string& msg) {
  t_err = msg;
  return code;
}
int llmi_cuda_fail(cudaError_t e, const char* what) {
  t_err = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return LLMI_ERR_CUDA;
}

extern "C" {

const char* llmi_last_error(void) { return t_err.c_str(); }
int llmi_abi_version(void) { return LLMI_ABI_VERSION; }
int llmi_sm_count(void) { return g.sm_count; }

uint64_t llmi_row_bytes(uint32_t t, uint64_t k) {
  switch (t) {
    case LLMI_Q4_0: return k % 32 ? 0 : k / 32 * 18;
    case LLMI_Q8_0: return k % 32 ? 0 : k / 32 * 34;
    case LLMI_Q5_0: return k % 32 ? 0 : k / 32 * 22;
    case LLMI_Q4_K: return k % 256 ? 0 : k / 256 * 144;
    case LLMI_Q6_K: return k % 256 ? 0 : k / 256 * 210;
    case LLMI_F16:
    case LLMI_BF16: return k * 2;
    default: return 0;
  }
}

int llmi_init(int device) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (g.ready && g.device == device) return LLMI_OK;
  if (g.ready) return llmi_fail(LLMI_ERR_STATE, "llmi_init: already initialised on another device");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) {
    (void)cudaGetLastError();
    return llmi_fail(LLMI_ERR_CUDA,
                     "llmi_init: no CUDA device available (this library has no CPU fallback)");
  }
  if (device < 0 || device >= n) return llmi_fail(LLMI_ERR_ARG, "llmi_init: bad device index");
  LLMI_CUDA_TRY(cudaSetDevice(device));
  cudaDeviceProp prop;
  LLMI_CUDA_TRY(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    return llmi_fail(LLMI_ERR_CUDA, std::string("llmi_init: device '") + prop.name +
                                        "' is not sm_100 (this library is built for B200 only)");
  }
  g.sm_count = prop.multiProcessorCount;
  if (const char* e = getenv("LLMI_NO_PDL")) g_llmi_pdl = !(e[0] == '1');
  if (const char* e = getenv("LLMI_GEMV_RING")) {  // "mode[,ctas_per_sm[,depth[,warps]]]" (llmi_set_gemv_ring), A/B runs
    int mode = 0, cps = 0, depth = 0, warps = 0;
    sscanf(e, "%d,%d,%d,%d", &mode, &cps, &depth, &warps);
    if (mode >= 0 && mode <= 2 && cps >= 0 && cps <= 4 && (depth == 0 || (depth >= 2 && depth <= 4)))
      llmi_gemv_set_ring(mode, cps, depth, warps);
  }
  LLMI_CUDA_TRY(cudaStreamCreateWithFlags(&g.stream, cudaStreamNonBlocking));
  LLMI_CUDA_TRY(llmi_gemv_init());
  LLMI_CUDA_TRY(llmi_mega_init());
  g.device = device;
  g.ready = true;
  return LLMI_OK;
}

int llmi_shutdown(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  if (!g.ready) return LLMI_OK;
  cudaDeviceSynchronize();
  for (auto& kv : g.registry) llmi_weight_free(kv.second.w);
  g.registry.clear();
  upload_pipeline_free();
  if (g.tok_act) cudaFree(g.tok_act);
  if (g.act) llmi_act_free(g.act);
  if (g.x_dev) cudaFree(g.x_dev);
  if (g.o_dev) cudaFree(g.o_dev);
  if (g.scratch) cudaFree(g.scratch);
  llmi_gemv_shutdown();
  if (g.stream) cudaStreamDestroy(g.stream);
  g = Context();
  return LLMI_OK;
}

// ------------------------------------------------------------------ weights

}  // extern "C"

namespace {

constexpr size_t UPLOAD_SLOT_BYTES = size_t(64) << 20;

int upload_pipeline_init(size_t need) {
  const size_t cap = std::max(UPLOAD_SLOT_BYTES, round_up(need, size_t(1) << 20));
  if (g.up_cap >= cap) return LLMI_OK;
  if (g.up_stream) LLMI_CUDA_TRY(cudaStreamSynchronize(g.up_stream));
  for (int i = 0; i < 2; ++i) {
    if (g.up_pinned[i]) cudaFreeHost(g.up_pinned[i]);
    if (g.up_dev[i]) cudaFree(g.up_dev[i]);
    g.up_pinned[i] = g.up_dev[i] = nullptr;
  }
  g.up_cap = 0;
  if (!g.up_stream) LLMI_CUDA_TRY(cudaStreamCreateWithFlags(&g.up_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; ++i) {
    LLMI_CUDA_TRY(cudaMallocHost(reinterpret_cast<void**>(&g.up_pinned[i]), cap));
    LLMI_CUDA_TRY(cudaMalloc(reinterpret_cast<void**>(&g.up_dev[i]), cap));
    if (!g.up_done[i]) LLMI_CUDA_TRY(cudaEventCreateWithFlags(&g.up_done[i], cudaEventDisableTiming));
  }
  g.up_cap = cap;
  return LLMI_OK;
}

// host copy into the pinned slot on a few threads: one thread moves ~4 GB/s out of a page-cache mapping (page faults
// included), which would make the CPU, not PCIe, the bound of a 17 GB load
void parallel_copy(uint8_t* dst, const uint8_t* src, size_t bytes) {
  constexpr size_t MIN_PER_THREAD = size_t(4) << 20;
  unsigned n = std::min<unsigned>(8, std::max(1u, std::thread::hardware_concurrency() / 2));
  n = unsigned(std::min<size_t>(n, std::max<size_t>(1, bytes / MIN_PER_THREAD)));
  if (n <= 1) {
    memcpy(dst, src, bytes);
    return;
  }
  std::vector<std::thread> th;
  const size_t per = (bytes / n + 4095) & ~size_t(4095);
  for (unsigned i = 0; i < n; ++i) {
    const size_t b0 = std::min(bytes, per * i), b1 = std::min(bytes, per * (i + 1));
    if (b1 > b0) th.emplace_back([=] { memcpy(dst + b0, src + b0, b1 - b0); });
  }
  for (auto& t : th) t.join();
}

void upload_pipeline_free() {
  if (g.up_stream) cudaStreamSynchronize(g.up_stream);
  for (int i = 0; i < 2; ++i) {
    if (g.up_pinned[i]) cudaFreeHost(g.up_pinned[i]);
    if (g.up_dev[i]) cudaFree(g.up_dev[i]);
    if (g.up_done[i]) cudaEventDestroy(g.up_done[i]);
  }
  if (g.up_stream) cudaStreamDestroy(g.up_stream);
}

}  // namespace

// GGUFFile::get_tensor_data (gguf.cpp:354-356) becomes: stream this handle's raw rows through the staging pipeline
// in chunks of whole slabs and repack each chunk on the device.  While the copy engine moves chunk i and the repack
// kernel rewrites it, the host fills chunk i + 1 into the other pinned slot.  No allocation, no stream
// synchronization per tensor: the planes are valid once the upload stream has drained (llmi_upload_wait).
int llmi_weight_upload_async(const void* host_blocks, uint32_t ggml_type, uint64_t n_cols, uint64_t n_rows,
                             uint64_t row_begin, uint64_t row_end, llmi_weight_t* out) {
  LLMI_NEED_INIT();
  if (!host_blocks || !out) return llmi_fail(LLMI_ERR_ARG, "llmi_weight_upload: null pointer");
  if (!supported(ggml_type))
    return llmi_fail(LLMI_ERR_TYPE, "mat_vec_mul: unsupported tensor type " + std::to_string(ggml_type));
  const uint64_t rb = llmi_row_bytes(ggml_type, n_cols);
  if (rb == 0 || n_cols == 0)
    return llmi_fail(LLMI_ERR_ARG, "llmi_weight_upload: n_cols is not a multiple of the block size");
  if (row_begin > row_end || row_end > n_rows) return llmi_fail(LLMI_ERR_ARG, "llmi_weight_upload: bad row range");
  llmi_weight_s* w = new llmi_weight_s();
  w->type = ggml_type;
  w->n_cols = n_cols;
  w->n_rows = n_rows;
  w->row_begin = row_begin;
  w->row_end = row_end;
  const size_t total = llmi_plan_planes(*w);
  w->bytes = total;
  if (w->n_local == 0) {
    *out = w;
    return LLMI_OK;
  }
  cudaError_t e = cudaMalloc(&w->base, total);
  if (e != cudaSuccess) {
    delete w;
    return llmi_cuda_fail(e, "cudaMalloc(weight planes)");
  }
  w->p_q = w->base + reinterpret_cast<size_t>(w->p_q);
  w->p_d = w->base + reinterpret_cast<size_t>(w->p_d);
  w->p_x = w->base + reinterpret_cast<size_t>(w->p_x);
  std::lock_guard<std::mutex> lk(g_mu);
  int rc = upload_pipeline_init(size_t(rb) * LLMI_SLAB);
  const uint8_t* src = static_cast<const uint8_t*>(host_blocks) + size_t(row_begin) * rb;
  const uint64_t slabs_per_chunk = std::max<uint64_t>(1, g.up_cap / (rb * LLMI_SLAB));
  for (uint64_t s0 = 0; rc == LLMI_OK && s0 < w->n_slabs; s0 += slabs_per_chunk) {
    const uint64_t n_sl = std::min<uint64_t>(slabs_per_chunk, w->n_slabs - s0);
    const uint64_t r0 = s0 * LLMI_SLAB, r1 = std::min<uint64_t>(w->n_local, (s0 + n_sl) * LLMI_SLAB);
    const size_t bytes = size_t(r1 - r0) * rb;
    const int slot = g.up_next;
    g.up_next ^= 1;
    e = cudaEventSynchronize(g.up_done[slot]);  // the slot's previous chunk has been copied AND repacked
    if (e == cudaSuccess) {
      parallel_copy(g.up_pinned[slot], src + size_t(r0) * rb, bytes);
      e = cudaMemcpyAsync(g.up_dev[slot], g.up_pinned[slot], bytes, cudaMemcpyHostToDevice, g.up_stream);
    }
    if (e == cudaSuccess) e = llmi_launch_repack_slabs(*w, g.up_dev[slot], s0, n_sl, g.up_stream);
    if (e == cudaSuccess) e = cudaEventRecord(g.up_done[slot], g.up_stream);
    if (e != cudaSuccess) rc = llmi_cuda_fail(e, "weight upload/repack");
  }
  if (rc != LLMI_OK) {
    cudaStreamSynchronize(g.up_stream);
    cudaFree(w->base);
    delete w;
    return rc;
  }
  *out = w;
  return LLMI_OK;
}

int llmi_upload_wait() {
  if (g.up_stream) LLMI_CUDA_TRY(cudaStreamSynchronize(g.up_stream));
  return LLMI_OK;
}

extern "C" {

int llmi_weight_upload(const void* host_blocks, uint32_t ggml_type, uint64_t n_cols, uint64_t n_rows,
                       uint64_t row_begin, uint64_t row_end, llmi_weight_t* out) {
  const int rc = llmi_weight_upload_async(host_blocks, ggml_type, n_cols, n_rows, row_begin, row_end, out);
  return rc == LLMI_OK ? llmi_upload_wait() : rc;
}

int llmi_weight_free(llmi_weight_t w) {
  if (!w) return LLMI_OK;
  if (w->base) cudaFree(w->base);
  delete w;
  return LLMI_OK;
}

int llmi_weight_dims(llmi_weight_t w, uint32_t* t, uint64_t* k, uint64_t* n, uint64_t* rb, uint64_t* re) {
  if (!w) return llmi_fail(LLMI_ERR_ARG, "llmi_weight_dims: null handle");
  if (t) *t = w->type;
  if (k) *k = w->n_cols;
  if (n) *n = w->n_rows;
  if (rb) *rb = w->row_begin;
  if (re) *re = w->row_end;
  return LLMI_OK;
}

uint64_t llmi_weight_device_bytes(llmi_weight_t w) { return w ? w->bytes : 0; }

namespace {
// Cheap content fingerprint of a host tensor: length, first and last KB and 64 samples in between (FNV-1a).  The
// registry is keyed by host ADDRESS (what an unmodified model.cpp hands to mat_vec_mul); if a GGUF image or a
// std::vector is freed and another one of the same shape lands at the same address, the address matches but the
// bytes do not — the stale device copy must not be used.
uint64_t fingerprint(const uint8_t* p, uint64_t bytes) {
  uint64_t h = 1469598103934665603ull ^ bytes;
  auto eat = [&](const uint8_t* q, uint64_t n) {
    for (uint64_t i = 0; i < n; ++i) h = (h ^ q[i]) * 1099511628211ull;
  };
  const uint64_t edge = std::min<uint64_t>(bytes, 1024);
  eat(p, edge);
  if (bytes > edge) eat(p + bytes - edge, edge);
  if (bytes > 4096)
    for (uint64_t i = 1; i <= 64; ++i) eat(p + (bytes / 65) * i, 16);
  return h;
}
}  // namespace

int llmi_registry_get(const void* host_blocks, uint32_t t, uint64_t k, uint64_t n, llmi_weight_t* out) {
  LLMI_NEED_INIT();
  if (!out || !host_blocks) return llmi_fail(LLMI_ERR_ARG, "llmi_registry_get: null pointer");
  const uint64_t bytes = llmi_row_bytes(t, k) * n;
  const uint64_t fp = fingerprint(static_cast<const uint8_t*>(host_blocks), bytes);
  auto key = std::make_tuple(host_blocks, t, k, n);
  {
    std::lock_guard<std::mutex> lk(g_mu);
    auto it = g.registry.find(key);
    if (it != g.registry.end()) {
      if (it->second.fp == fp) {
        *out = it->second.w;
        return LLMI_OK;
      }
      llmi_weight_free(it->second.w);  // same address, other bytes: the host buffer was recycled
      g.registry.erase(it);
    }
  }
  llmi_weight_t w = nullptr;
  const int rc = llmi_weight_upload(host_blocks, t, k, n, 0, n, &w);
  if (rc != LLMI_OK) return rc;
  std::lock_guard<std::mutex> lk(g_mu);
  g.registry[key] = Context::RegEntry{w, fp};
  *out = w;
  return LLMI_OK;
}

int llmi_registry_clear(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (auto& kv : g.registry) llmi_weight_free(kv.second.w);
  g.registry.clear();
  return LLMI_OK;
}

// -------------------------------------------------------------- activations

int llmi_act_create(uint64_t max_cols, llmi_act_t* out) {
  LLMI_NEED_INIT();
  if (!out || max_cols == 0) return llmi_fail(LLMI_ERR_ARG, "llmi_act_create: bad argument");
  llmi_act_s* a = new llmi_act_s();
  a->max_cols = max_cols;
  a->buf_bytes = round_up(4 * max_cols + 64, 256);  // fp32 staging is the largest kind
  cudaError_t e = cudaMalloc(&a->buf, a->buf_bytes);
  if (e != cudaSuccess) {
    delete a;
    return llmi_cuda_fail(e, "cudaMalloc(activation)");
  }
  *out = a;
  return LLMI_OK;
}

int llmi_act_free(llmi_act_t a) {
  if (!a) return LLMI_OK;
  if (a->buf) cudaFree(a->buf);
  delete a;
  return LLMI_OK;
}

static int act_check(llmi_act_t a, const float* x, uint64_t n, uint64_t multiple, const char* who) {
  if (!a || !x) return llmi_fail(LLMI_ERR_ARG, std::string(who) + ": null pointer");
  if (n == 0 || n > a->max_cols) return llmi_fail(LLMI_ERR_SIZE, std::string(who) + ": n exceeds activation capacity");
  if (n % multiple) return llmi_fail(LLMI_ERR_ARG, std::string(who) + ": n is not a multiple of the block size");
  return LLMI_OK;
}

int llmi_quantize_q8_0(const float* x, uint64_t n, llmi_act_t a, llmi_stream_t s) {
  LLMI_NEED_INIT();
  if (int rc = act_check(a, x, n, 32, "llmi_quantize_q8_0")) return rc;
  LLMI_CUDA_TRY(llmi_launch_quantize_q8_0(x, n, a->buf, (cudaStream_t)s));
  a->kind = ACT_Q8_0;
  a->n = n;
  a->last_stream = (cudaStream_t)s;
  return LLMI_OK;
}

int llmi_quantize_q8_k(const float* x, uint64_t n, llmi_act_t a, llmi_stream_t s) {
  LLMI_NEED_INIT();
  if (int rc = act_check(a, x, n, 256, "llmi_quantize_q8_k")) return rc;
  LLMI_CUDA_TRY(llmi_launch_quantize_q8_k(x, n, a->buf, (cudaStream_t)s));
  a->kind = ACT_Q8_K;
  a->n = n;
  a->last_stream = (cudaStream_t)s;
  return LLMI_OK;
}

int llmi_round_f16(const float* x, uint64_t n, llmi_act_t a, llmi_stream_t s) {
  LLMI_NEED_INIT();
  if (int rc = act_check(a, x, n, 1, "llmi_round_f16")) return rc;
  LLMI_CUDA_TRY(llmi_launch_round_f16(x, n, a->buf, (cudaStream_t)s));
  a->kind = ACT_F16;
  a->n = n;
  a->last_stream = (cudaStream_t)s;
  return LLMI_OK;
}

int llmi_stage_f32(const float* x, uint64_t n, llmi_act_t a, llmi_stream_t s) {
  LLMI_NEED_INIT();
  if (int rc = act_check(a, x, n, 1, "llmi_stage_f32")) return rc;
  const size_t padded = act_bytes(ACT_F32, n);
  if (padded > 4 * n) LLMI_CUDA_TRY(cudaMemsetAsync(a->buf + 4 * n, 0, padded - 4 * n, (cudaStream_t)s));
  LLMI_CUDA_TRY(cudaMemcpyAsync(a->buf, x, 4 * n, cudaMemcpyDeviceToDevice, (cudaStream_t)s));
  a->kind = ACT_F32;
  a->n = n;
  a->last_stream = (cudaStream_t)s;
  return LLMI_OK;
}

int llmi_act_prepare(llmi_weight_t w, const float* x, llmi_act_t a, llmi_stream_t s) {
  if (!w) return llmi_fail(LLMI_ERR_ARG, "llmi_act_prepare: null weight");
  switch (llmi_act_kind_for(w->type)) {
    case ACT_Q8_0: return llmi_quantize_q8_0(x, w->n_cols, a, s);
    case ACT_Q8_K: return llmi_quantize_q8_k(x, w->n_cols, a, s);
    case ACT_F16: return llmi_round_f16(x, w->n_cols, a, s);
    case ACT_F32: return llmi_stage_f32(x, w->n_cols, a, s);
    default: return llmi_fail(LLMI_ERR_TYPE, "mat_vec_mul: unsupported tensor type " + std::to_string(w->type));
  }
}

static int act_export(llmi_act_t a, void* host, int kind, size_t rec, size_t per, const char* who) {
  LLMI_NEED_INIT();
  if (!a || !host) return llmi_fail(LLMI_ERR_ARG, std::string(who) + ": null pointer");
  if (a->kind != kind) return llmi_fail(LLMI_ERR_STATE, std::string(who) + ": activation holds another format");
  const size_t bytes = a->n / per * rec;
  if (int rc = ensure_cap((void**)&g.scratch, &g.scratch_cap, bytes)) return rc;
  cudaStream_t s = a->last_stream;
  if (kind == ACT_Q8_0) LLMI_CUDA_TRY(llmi_launch_export_q8_0(a->buf, a->n, g.scratch, s));
  else LLMI_CUDA_TRY(llmi_launch_export_q8_k(a->buf, a->n, g.scratch, s));
  LLMI_CUDA_TRY(cudaMemcpyAsync(host, g.scratch, bytes, cudaMemcpyDeviceToHost, s));
  LLMI_CUDA_TRY(cudaStreamSynchronize(s));
  return LLMI_OK;
}

int llmi_act_export_q8_0(llmi_act_t a, void* host) { return act_export(a, host, ACT_Q8_0, 34, 32, "llmi_act_export_q8_0"); }
int llmi_act_export_q8_k(llmi_act_t a, void* host) { return act_export(a, host, ACT_Q8_K, 292, 256, "llmi_act_export_q8_k"); }

// ------------------------------------------------------------------- mat-vec

int llmi_set_gemv_shape(int warps, int slabs_per_cta) {
  if ((warps != 0 && warps != 4 && warps != 8 && warps != 16) || slabs_per_cta < 0 || slabs_per_cta > 64)
    return llmi_fail(LLMI_ERR_ARG, "llmi_set_gemv_shape: warps in {0,4,8,16}, slabs_per_cta in [0,64]");
  llmi_gemv_set_shape(warps, slabs_per_cta);
  return LLMI_OK;
}

int llmi_set_gemv_ring(int mode, int ctas_per_sm, int depth, int warps) {
  if (mode < 0 || mode > 2 || ctas_per_sm < 0 || ctas_per_sm > 4 || (depth != 0 && (depth < 2 || depth > 4)) ||
      (warps != 0 && warps != 8 && warps != 16))
    return llmi_fail(LLMI_ERR_ARG, "llmi_set_gemv_ring: mode in {0,1,2}, ctas_per_sm in [0,4], depth in {0,2,3,4}, warps in {0,8,16}");
  llmi_gemv_set_ring(mode, ctas_per_sm, depth, warps);
  return LLMI_OK;
}

int llmi_set_prefill_mode(int mode) {
  if (mode != 0 && mode != 1) return llmi_fail(LLMI_ERR_ARG, "llmi_set_prefill_mode: 0 = exact (default), 1 = fast (bf16 tensor cores)");
  llmi_gemv_set_prefill_fast(mode);
  return LLMI_OK;
}

int llmi_gemv(llmi_weight_t w, llmi_act_t a, float* out, llmi_stream_t s) {
  LLMI_NEED_INIT();
  if (!w || !a || !out) return llmi_fail(LLMI_ERR_ARG, "llmi_gemv: null pointer");
  if (a->kind != llmi_act_kind_for(w->type))
    return llmi_fail(LLMI_ERR_STATE, "llmi_gemv: activation was not prepared for this weight format");
  if (a->n != w->n_cols) return llmi_fail(LLMI_ERR_SIZE, "mat_vec_mul: input vector size mismatch");
  LLMI_CUDA_TRY(llmi_launch_gemv(*w, *a, out, (cudaStream_t)s));
  return LLMI_OK;
}

int llmi_gemv_batch(const llmi_weight_t* ws, float* const* outs, int n, llmi_act_t a, llmi_stream_t s) {
  LLMI_NEED_INIT();
  if (!ws || !outs || !a || n < 1 || n > 3) return llmi_fail(LLMI_ERR_ARG, "llmi_gemv_batch: need 1..3 matrices");
  const llmi_weight_s* w[3];
  for (int i = 0; i < n; ++i) {
    if (!ws[i] || !outs[i]) return llmi_fail(LLMI_ERR_ARG, "llmi_gemv_batch: null pointer");
    if (ws[i]->type != ws[0]->type) return llmi_fail(LLMI_ERR_TYPE, "llmi_gemv_batch: matrices must share one format");
    if (a->kind != llmi_act_kind_for(ws[i]->type))
      return llmi_fail(LLMI_ERR_STATE, "llmi_gemv_batch: activation was not prepared for this weight format");
    if (a->n != ws[i]->n_cols) return llmi_fail(LLMI_ERR_SIZE, "mat_vec_mul: input vector size mismatch");
    w[i] = ws[i];
  }
  LLMI_CUDA_TRY(llmi_launch_gemv_batch(w, outs, n, *a, (cudaStream_t)s));
  return LLMI_OK;
}

int llmi_mat_vec_mul_dev(llmi_weight_t w, const float* x, llmi_act_t a, float* out, llmi_stream_t s) {
  if (int rc = llmi_act_prepare(w, x, a, s)) return rc;
  return llmi_gemv(w, a, out, s);
}

// Token-batched mat-vec (prefill), device tier: out[m][rows of w] = W * x[m] for m < n_tokens.  The activations
// are quantized as the format requires (one kernel over all tokens) into a library-owned buffer, then the
// token-batched kernels run (gemv.cu).  Bit-identical to n_tokens calls of llmi_mat_vec_mul_dev.
int llmi_gemm_tokens(llmi_weight_t w, const float* x, uint32_t n_tokens, float* out, llmi_stream_t stream) {
  LLMI_NEED_INIT();
  if (!w || !x || !out) return llmi_fail(LLMI_ERR_ARG, "llmi_gemm_tokens: null pointer");
  if (n_tokens == 0 || w->n_local == 0) return LLMI_OK;
  const int kind = llmi_act_kind_for(w->type);
  if (kind == ACT_NONE) return llmi_fail(LLMI_ERR_TYPE, "llmi_gemm_tokens: unsupported tensor type");
  if (kind == ACT_Q8_K && w->n_cols % 256 != 0)
    return llmi_fail(LLMI_ERR_SIZE, "mat_vec_mul: n_cols must be a multiple of 256");
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);  // NULL: the default stream, like the other device-tier calls
  const size_t stride = act_bytes(kind, w->n_cols);
  if (stride * n_tokens > g.tok_act_cap) LLMI_CUDA_TRY(cudaStreamSynchronize(s));
  if (int rc = ensure_cap((void**)&g.tok_act, &g.tok_act_cap, stride * n_tokens)) return rc;
  LLMI_CUDA_TRY(llmi_launch_act(x, uint32_t(w->n_cols), kind, g.tok_act, s, n_tokens, uint32_t(stride)));
  const llmi_weight_s* ws[1] = {w};
  float* outs[1] = {out};
  const uint32_t strides[1] = {uint32_t(w->n_rows)};
  LLMI_CUDA_TRY(llmi_launch_gemv_tokens(ws, outs, strides, 1, kind, w->n_cols, g.tok_act, n_tokens, s));
  return LLMI_OK;
}

int llmi_debug_block_dots(llmi_weight_t w, llmi_act_t a, int32_t* dots_host) {
  LLMI_NEED_INIT();
  if (!w || !a || !dots_host) return llmi_fail(LLMI_ERR_ARG, "llmi_debug_block_dots: null pointer");
  if (w->type != LLMI_Q4_0 && w->type != LLMI_Q8_0 && w->type != LLMI_Q4_K && w->type != LLMI_Q6_K)
    return llmi_fail(LLMI_ERR_TYPE, "llmi_debug_block_dots: format has no integer block dots");
  if (a->kind != llmi_act_kind_for(w->type) || a->n != w->n_cols)
    return llmi_fail(LLMI_ERR_STATE, "llmi_debug_block_dots: activation not prepared for this weight");
  const uint64_t per_row = w->type == LLMI_Q6_K ? w->nb * 2 : (w->type == LLMI_Q4_K ? w->nb * 8 : w->nb);
  const size_t bytes = size_t(w->n_local) * per_row * 4;
  if (int rc = ensure_cap((void**)&g.scratch, &g.scratch_cap, bytes)) return rc;
  cudaStream_t s = a->last_stream;
  LLMI_CUDA_TRY(llmi_launch_block_dots(*w, *a, reinterpret_cast<int32_t*>(g.scratch), s));
  LLMI_CUDA_TRY(cudaMemcpyAsync(dots_host, g.scratch, bytes, cudaMemcpyDeviceToHost, s));
  LLMI_CUDA_TRY(cudaStreamSynchronize(s));
  return LLMI_OK;
}

// ---------------------------------------------------------- host-vector tier

int llmi_host_mat_vec_mul(llmi_weight_t w, const float* x, uint64_t n_x, float* o, uint64_t n_o) {
  LLMI_NEED_INIT();
  if (!w || !x || !o) return llmi_fail(LLMI_ERR_ARG, "llmi_host_mat_vec_mul: null pointer");
  if (n_x != w->n_cols) return llmi_fail(LLMI_ERR_SIZE, "mat_vec_mul: input vector size mismatch");
  if (n_o != w->n_rows) return llmi_fail(LLMI_ERR_SIZE, "mat_vec_mul: output vector size mismatch");
  if (int rc = ensure_cap((void**)&g.x_dev, &g.x_cap, 4 * n_x)) return rc;
  if (int rc = ensure_cap((void**)&g.o_dev, &g.o_cap, 4 * n_o)) return rc;
  if (!g.act || g.act->max_cols < n_x) {
    if (g.act) llmi_act_free(g.act);
    g.act = nullptr;
    if (int rc = llmi_act_create(n_x, &g.act)) return rc;
  }
  LLMI_CUDA_TRY(cudaMemcpyAsync(g.x_dev, x, 4 * n_x, cudaMemcpyHostToDevice, g.stream));
  if (int rc = llmi_mat_vec_mul_dev(w, g.x_dev, g.act, g.o_dev, g.stream)) return rc;
  LLMI_CUDA_TRY(cudaMemcpyAsync(o + w->row_begin, g.o_dev + w->row_begin, 4 * w->n_local, cudaMemcpyDeviceToHost,
                                g.stream));
  LLMI_CUDA_TRY(cudaStreamSynchronize(g.stream));
  return LLMI_OK;
}

static int host_quantize(const float* x, uint64_t n, void* y, bool k_quant) {
  LLMI_NEED_INIT();
  if (!x || !y) return llmi_fail(LLMI_ERR_ARG, "quantize_row: null pointer");
  if (n == 0) return LLMI_OK;
  if (int rc = ensure_cap((void**)&g.x_dev, &g.x_cap, 4 * n)) return rc;
  if (!g.act || g.act->max_cols < n) {
    if (g.act) llmi_act_free(g.act);
    g.act = nullptr;
    if (int rc = llmi_act_create(n, &g.act)) return rc;
  }
  LLMI_CUDA_TRY(cudaMemcpyAsync(g.x_dev, x, 4 * n, cudaMemcpyHostToDevice, g.stream));
  if (k_quant) {
    if (int rc = llmi_quantize_q8_k(g.x_dev, n, g.act, g.stream)) return rc;
    return llmi_act_export_q8_k(g.act, y);
  }
  if (int rc = llmi_quantize_q8_0(g.x_dev, n, g.act, g.stream)) return rc;
  return llmi_act_export_q8_0(g.act, y);
}

int llmi_host_quantize_row_q8_0(const float* x, uint64_t n, void* y) { return host_quantize(x, n, y, false); }
int llmi_host_quantize_row_q8_k(const float* x, uint64_t n, void* y) { return host_quantize(x, n, y, true); }

// ------------------------------------------------------------ device helpers

int llmi_dev_alloc(uint64_t bytes, void** out) {
  LLMI_NEED_INIT();
  if (!out) return llmi_fail(LLMI_ERR_ARG, "llmi_dev_alloc: null out");
  LLMI_CUDA_TRY(cudaMalloc(out, bytes ? bytes : 16));
  return LLMI_OK;
}
int llmi_dev_free(void* p) {
  if (p) LLMI_CUDA_TRY(cudaFree(p));
  return LLMI_OK;
}
int llmi_h2d(void* dst, const void* src, uint64_t bytes) {
  LLMI_NEED_INIT();
  LLMI_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyHostToDevice));
  return LLMI_OK;
}
int llmi_d2h(void* dst, const void* src, uint64_t bytes) {
  LLMI_NEED_INIT();
  LLMI_CUDA_TRY(cudaMemcpy(dst, src, bytes, cudaMemcpyDeviceToHost));
  return LLMI_OK;
}
int llmi_device_sync(void) {
  LLMI_NEED_INIT();
  LLMI_CUDA_TRY(cudaDeviceSynchronize());
  return LLMI_OK;
}

}  // extern "C"
