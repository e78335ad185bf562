// Dev tool: times attention_kernel alone and prints CTA-0 cycle stamps per phase.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -DLLMI_ATTN_TIMING \
//        -I include -I llm_inference_b200/csrc tools/attn_bench.cu -o /tmp/attn_bench && /tmp/attn_bench 100
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
bool g_llmi_pdl = false;
#include "../llm_inference_b200/csrc/glue.cu"
int main(int argc, char** argv) {
  const int T = argc > 1 ? atoi(argv[1]) : 100;
  const uint32_t H = 4, HK = 1, D = 256, t_max = argc > 2 ? atoi(argv[2]) : 256;
  float *q, *k, *v, *wq, *wk, *out;
  uint32_t* kc;
  __half* vc;
  float2* rope;
  int32_t* pos;
  uint8_t* act;
  cudaMalloc(&q, H * D * 4); cudaMalloc(&k, HK * D * 4); cudaMalloc(&v, HK * D * 4);
  cudaMalloc(&wq, D * 4); cudaMalloc(&wk, D * 4); cudaMalloc(&out, H * D * 4);
  cudaMalloc(&kc, size_t(t_max) * HK * D * 4); cudaMalloc(&vc, size_t(t_max) * HK * D * 2);
  cudaMalloc(&rope, size_t(t_max) * (D / 2) * 8); cudaMalloc(&pos, 4); cudaMalloc(&act, H * D * 2);
  std::vector<float> h(H * D, 0.5f);
  cudaMemcpy(q, h.data(), H * D * 4, cudaMemcpyHostToDevice); cudaMemcpy(k, h.data(), HK * D * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(v, h.data(), HK * D * 4, cudaMemcpyHostToDevice); cudaMemcpy(wq, h.data(), D * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(wk, h.data(), D * 4, cudaMemcpyHostToDevice);
  std::vector<__half> hk(size_t(t_max) * HK * D);
  for (size_t i = 0; i < hk.size(); ++i) hk[i] = __float2half(float((i * 7919) % 97) / 97.0f - 0.5f);
  std::vector<uint32_t> hk32(hk.size());
  for (size_t i = 0; i < hk.size(); ++i) { double d = double(__half2float(hk[i])); unsigned long long b; memcpy(&b, &d, 8); hk32[i] = uint32_t(b >> 32); }
  cudaMemcpy(kc, hk32.data(), hk32.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(vc, hk.data(), hk.size() * 2, cudaMemcpyHostToDevice);
  const int p = T - 1;
  cudaMemcpy(pos, &p, 4, cudaMemcpyHostToDevice);
  llmi_launch_rope_table(rope, t_max, D, 10000.0f, 1.0f, 0);
  llmi_attention_init(t_max, D);
  AttnArgs a{};
  a.q = q; a.k = k; a.v = v; a.wq_norm = wq; a.wk_norm = wk; a.kcache = kc; a.vcache = vc; a.H = H; a.HK = HK; a.D = D;
  a.t_max = t_max; a.eps = 1e-6; a.attn_scale = 0.0625f; a.rope_table = rope; a.pos = pos; a.softcap = 0; a.out = out;
  a.act_kind = ACT_Q8_0; a.act_buf = act;
  for (int i = 0; i < 5; ++i) llmi_launch_attention(a, 0);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 200; ++i) llmi_launch_attention(a, 0);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  long long st[16] = {0};
#ifdef LLMI_ATTN_TIMING
  cudaMemcpyFromSymbol(st, g_attn_stamp, sizeof st);
#endif
  printf("T=%d: %.2f us per launch (back to back); err=%s\n", T, ms * 1000 / 200, cudaGetErrorString(cudaGetLastError()));
  const char* names[] = {"pdl_wait, tile issue, input loads", "norms+rope+kv write", "phase1 (K tiles, scores)",
                         "phase2a/2b", "phase2c+phase3 (V tiles)", "output"};
  for (int i = 0; i < 6; ++i) printf("  %-34s %6lld cycles\n", names[i], st[i + 1] - st[i]);
  printf("  2c scan (thread 1023) %lld cycles; V tile 0: chunks %lld cycles, tile_done %lld, wait-for-tile after stamp4 %lld\n", st[9] - st[8], st[11] - st[10], st[12] - st[11], st[10] - st[4]);
  return 0;
}
