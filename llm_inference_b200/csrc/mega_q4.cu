// Persistent decode kernel, instantiations for Q4_0 layers with F16 (or Q4_0) embeddings / logits: gemma-3-1b / 27b Q4_0.
#include "mega_impl.cuh"

const MegaVariant* llmi_mega_variants_q4(int* n) {
  static const MegaVariant v[] = {MEGA_VARIANT(mega_type_bit(LLMI_Q4_0) | mega_type_bit(LLMI_F16), 128), MEGA_VARIANT(mega_type_bit(LLMI_Q4_0) | mega_type_bit(LLMI_F16), 256)};
  *n = int(sizeof(v) / sizeof(v[0]));
  return v;
}

#ifdef LLMI_MEGA_TIMING  // dev only (tools/mega_timeline.py): the stamps of this file's instantiations
extern "C" int llmi_debug_mega_stamps_q4(unsigned long long* out /*[2][1024][16]*/) { return int(mega_variant_stamps(out)); }
extern "C" int llmi_debug_mega_cycles_q4(long long* out /*[1024][32]*/) { return int(mega_variant_cycles(out)); }
#endif
