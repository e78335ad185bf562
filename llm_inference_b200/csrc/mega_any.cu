// Persistent decode kernel, the all-formats / any-head-size instantiation (fallback; hot bodies out of line).
#define MEGA_HOT __noinline__
#include "mega_impl.cuh"

const MegaVariant* llmi_mega_variants_any(int* n) {
  static const MegaVariant v[] = {MEGA_VARIANT(0x7fu, 0)};
  *n = int(sizeof(v) / sizeof(v[0]));
  return v;
}

#ifdef LLMI_MEGA_TIMING  // dev only (tools/mega_timeline.py): the stamps of this file's instantiations
extern "C" int llmi_debug_mega_stamps_any(unsigned long long* out /*[2][1024][16]*/) { return int(mega_variant_stamps(out)); }
extern "C" int llmi_debug_mega_cycles_any(long long* out /*[1024][32]*/) { return int(mega_variant_cycles(out)); }
#endif
