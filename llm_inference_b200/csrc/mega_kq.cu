// Persistent decode kernel, instantiations for the k-quants (Q4_K_M layout: Q4_K / Q6_K): gemma-3-4b.
#include "mega_impl.cuh"

const MegaVariant* llmi_mega_variants_kq(int* n) {
  static const MegaVariant v[] = {MEGA_VARIANT(mega_type_bit(LLMI_Q4_K) | mega_type_bit(LLMI_Q6_K), 128), MEGA_VARIANT(mega_type_bit(LLMI_Q4_K) | mega_type_bit(LLMI_Q6_K), 256)};
  *n = int(sizeof(v) / sizeof(v[0]));
  return v;
}

#ifdef LLMI_MEGA_TIMING  // dev only (tools/mega_timeline.py): the stamps of this file's instantiations
extern "C" int llmi_debug_mega_stamps_kq(unsigned long long* out /*[2][1024][16]*/) { return int(mega_variant_stamps(out)); }
extern "C" int llmi_debug_mega_cycles_kq(long long* out /*[1024][32]*/) { return int(mega_variant_cycles(out)); }
#endif
