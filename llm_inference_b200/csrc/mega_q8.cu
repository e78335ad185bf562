// Persistent decode kernel, instantiations for all-Q8_0 models: gemma-3-12b Q8_0.
#include "mega_impl.cuh"

const MegaVariant* llmi_mega_variants_q8(int* n) {
  static const MegaVariant v[] = {MEGA_VARIANT(mega_type_bit(LLMI_Q8_0), 128), MEGA_VARIANT(mega_type_bit(LLMI_Q8_0), 256)};
  *n = int(sizeof(v) / sizeof(v[0]));
  return v;
}

#ifdef LLMI_MEGA_TIMING  // dev only (tools/mega_timeline.py): the stamps of this file's instantiations
extern "C" int llmi_debug_mega_stamps_q8(unsigned long long* out /*[2][1024][16]*/) { return int(mega_variant_stamps(out)); }
extern "C" int llmi_debug_mega_cycles_q8(long long* out /*[1024][32]*/) { return int(mega_variant_cycles(out)); }
#endif
