// Persistent decode kernel: the table of instantiations (mega_impl.cuh) and the choice among them.
#include "mega.h"

const MegaVariant* llmi_mega_variants_q4(int* n);   // mega_q4.cu : Q4_0 layers + F16 / Q4_0 logits
const MegaVariant* llmi_mega_variants_q8(int* n);   // mega_q8.cu : Q8_0
const MegaVariant* llmi_mega_variants_kq(int* n);   // mega_kq.cu : Q4_K / Q6_K
const MegaVariant* llmi_mega_variants_any(int* n);  // mega_any.cu: every format, every head size (fallback)

namespace {
struct Entry {
  MegaVariant v;
  size_t smem = 0;
};
Entry g_variants[32];
int g_n_variants = 0;
uint32_t g_mega_ctas = 0;
}  // namespace

cudaError_t llmi_mega_init() {
  int dev = 0, sms = 0, coop = 0;
  cudaError_t e;
  if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return e;
  if ((e = cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev)) != cudaSuccess) return e;
  if (!coop) return cudaErrorNotSupported;
  g_n_variants = 0;
  for (auto get : {llmi_mega_variants_q4, llmi_mega_variants_q8, llmi_mega_variants_kq, llmi_mega_variants_any}) {
    int n = 0;
    const MegaVariant* v = get(&n);
    for (int i = 0; i < n && g_n_variants < 32; ++i) {
      Entry& en = g_variants[g_n_variants];
      en.v = v[i];
      if ((e = v[i].init(&en.smem)) != cudaSuccess) return e;
      ++g_n_variants;
    }
  }
  g_mega_ctas = uint32_t(sms);
  return cudaSuccess;
}

uint32_t llmi_mega_max_ctas() { return g_mega_ctas; }

// The first (most specialized) instantiation that carries every format of `type_mask` and the head size.
int llmi_mega_select(uint32_t type_mask, uint32_t head_dim, size_t* smem_limit) {
  for (int i = 0; i < g_n_variants; ++i) {
    const MegaVariant& v = g_variants[i].v;
    if ((type_mask & ~v.type_mask) == 0 && (v.head_dim == 0 || uint32_t(v.head_dim) == head_dim)) {
      if (smem_limit) *smem_limit = g_variants[i].smem;
      return i;
    }
  }
  return -1;
}

// K/V tiles (2..6) of 32 KB that fit into `avail` bytes next to the per-position arrays; 0 = t_max too large.
uint32_t llmi_mega_attention_nbuf(uint32_t t_max, uint32_t D, size_t avail, size_t* bytes) {
  constexpr size_t TILE = 32768, MAX_BUF = 6;  // ATT_TILE_BYTES, ATT_MAX_BUF (glue_device.cuh)
  const size_t tp = (t_max + 15) & ~15u;
  const size_t fixed = tp * (8 + 4 + 4 + 1) + size_t(D) * (4 + 4 + 4 + 2) + 128;
  if (fixed + 2 * TILE > avail) return 0;
  size_t n = (avail - fixed) / TILE;
  if (n > MAX_BUF) n = MAX_BUF;
  if (bytes) *bytes = fixed + n * TILE;
  return uint32_t(n);
}

cudaError_t llmi_launch_mega(int variant, const MegaArgs& a, uint32_t n_ctas, size_t smem, cudaStream_t s) {
  if (variant < 0 || variant >= g_n_variants) return cudaErrorInvalidValue;
  if (n_ctas == 0 || n_ctas > g_mega_ctas || smem > g_variants[variant].smem) return cudaErrorInvalidValue;
  return g_variants[variant].v.launch(a, n_ctas, smem, s);
}
