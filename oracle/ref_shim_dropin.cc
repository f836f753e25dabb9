// oracle/ref_shim_dropin.cc -- TEST INFRASTRUCTURE ONLY: the drop-in check.
//
// Built into oracle/_ref/libgzref_dropin.so, SEPARATE from libgzref.so: the reference checker
// (libgzref.so = reference sources + forwarding shim) must not depend on the product it checks, so the
// one test that deliberately couples the two -- the unmodified reference Processor driving the product's
// comparator through the C ABI -- lives in its own library, which links libgzref.so (for
// guetzli::ProcessJpegData / EncodeRGBToJpeg), integration/gzb_comparator.o and libgzb200.so.
#include <cstdint>
#include <cstring>
#include <vector>

#include "guetzli/jpeg_data.h"
#include "guetzli/jpeg_data_encoder.h"
#include "guetzli/processor.h"
#include "guetzli/stats.h"
#include "gzb_comparator.h"   // integration/: the reference-side adaptor over libgzb200.so

extern "C" {

// The UNMODIFIED reference Processor (guetzli::ProcessJpegData, processor.cc:931-1027) driven by
// the B200 comparator adaptor of integration/gzb_comparator.{h,cc}: the drop-in test. MODE_CPU
// control flow, so the per-block CompareBlock virtual is used (one GPU launch per call).
long ref_process_rgb_b200_params(const uint8_t* rgb, int w, int h, float target, int try_420, int force_420,
                                 int device, uint8_t* out, long cap, int* iters);
long ref_process_rgb_b200(const uint8_t* rgb, int w, int h, float target, int device, uint8_t* out,
                          long cap, int* iters) {
  return ref_process_rgb_b200_params(rgb, w, h, target, 0, 0, device, out, cap, iters);
}
long ref_process_rgb_b200_params(const uint8_t* rgb, int w, int h, float target, int try_420, int force_420,
                                 int device, uint8_t* out, long cap, int* iters) {
  guetzli::Params params;
  params.butteraugli_target = target;
  params.try_420 = try_420 != 0;
  params.force_420 = force_420 != 0;
  guetzli::ProcessStats stats;
  std::vector<uint8_t> v(rgb, rgb + size_t(3) * w * h);
  guetzli::JPEGData jpg;
  if (!guetzli::EncodeRGBToJpeg(v, w, h, &jpg)) return -1;
  guetzli::GuetzliOutput o;
  {
    guetzli::B200ButteraugliComparator cmp(w, h, &v, target, &stats, device);
    if (!guetzli::ProcessJpegData(params, jpg, &cmp, &o, &stats)) return -1;
  }
  if (static_cast<long>(o.jpeg_data.size()) <= cap) memcpy(out, o.jpeg_data.data(), o.jpeg_data.size());
  if (iters) *iters = stats.counters[guetzli::kNumItersCnt];
  return static_cast<long>(o.jpeg_data.size());
}

}  // extern "C"
