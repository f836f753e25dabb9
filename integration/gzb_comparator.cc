// integration/gzb_comparator.cc -- see gzb_comparator.h.
#include "gzb_comparator.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "guetzli/debug_print.h"

namespace guetzli {

B200ButteraugliComparator::B200ButteraugliComparator(int width, int height, const std::vector<uint8_t>* rgb,
                                                     float target_distance, ProcessStats* stats, int device)
    : width_(width), height_(height), ctx_(nullptr), stats_(stats), distance_(0.0f), block_x_(0), block_y_(0),
      factor_x_(1), factor_y_(1) {
  if (gzb_create(device, width, height, rgb->data(), target_distance, &ctx_) != GZB_OK) Die("gzb_create");
}

B200ButteraugliComparator::~B200ButteraugliComparator() { gzb_destroy(ctx_); }

// The reference logs accelerator errors and carries on (clguetzli/ocu.h:14); a wrong JPEG is worse
// than no JPEG, and there is no CPU fallback to hide behind: abort.
void B200ButteraugliComparator::Die(const char* what) const {
  fprintf(stderr, "gzb200: %s failed: %s\n", what, gzb_last_error(ctx_));
  abort();
}

namespace {
bool Is444(const OutputImage& img) {
  for (int c = 0; c < 3; ++c)
    if (img.component(c).factor_x() != 1 || img.component(c).factor_y() != 1) return false;
  return true;
}
bool Is420(const OutputImage& img) {
  return img.component(0).factor_x() == 1 && img.component(0).factor_y() == 1 &&
         img.component(1).factor_x() == 2 && img.component(1).factor_y() == 2 &&
         img.component(2).factor_x() == 2 && img.component(2).factor_y() == 2;
}
// The luma plane of a 4:2:0 image in gzb's MCU-padded layout (2*ceil(w/16) blocks per row); the
// padding blocks only matter to the JPEG writer and stay zero here.
std::vector<coeff_t> PaddedLuma(const OutputImage& img) {
  const OutputImageComponent& y = img.component(0);
  const int bw = y.width_in_blocks(), bh = y.height_in_blocks();
  const int cbw = 2 * img.component(1).width_in_blocks(), cbh = 2 * img.component(1).height_in_blocks();
  std::vector<coeff_t> out(static_cast<size_t>(cbw) * cbh * kDCTBlockSize, 0);
  for (int by = 0; by < bh; ++by)
    memcpy(&out[static_cast<size_t>(by) * cbw * kDCTBlockSize], y.coeffs() + static_cast<size_t>(by) * bw * kDCTBlockSize,
           static_cast<size_t>(bw) * kDCTBlockSize * sizeof(coeff_t));
  return out;
}
}  // namespace

// OutputImage keeps dequantised coefficients block-major per component: exactly gzb's layout (for a
// 4:2:0 image the luma rows are re-pitched to whole MCUs).
void B200ButteraugliComparator::PushImage(const OutputImage& img) const {
  if (Is444(img)) {
    if (gzb_set_sampling(ctx_, 1) != GZB_OK) Die("gzb_set_sampling");
    if (gzb_set_coeffs(ctx_, img.component(0).coeffs(), img.component(1).coeffs(), img.component(2).coeffs()) != GZB_OK)
      Die("gzb_set_coeffs");
  } else if (Is420(img)) {
    if (gzb_set_sampling(ctx_, 2) != GZB_OK) Die("gzb_set_sampling");
    const std::vector<coeff_t> y = PaddedLuma(img);
    if (gzb_set_coeffs(ctx_, y.data(), img.component(1).coeffs(), img.component(2).coeffs()) != GZB_OK)
      Die("gzb_set_coeffs");
  } else {
    Die("chroma sampling other than 4:4:4 / 4:2:0");
  }
}

void B200ButteraugliComparator::Compare(const OutputImage& img) {
  PushImage(img);
  if (gzb_compare(ctx_, &distance_) != GZB_OK) Die("gzb_compare");
  GUETZLI_LOG(stats_, " BA[100.00%%] D[%6.4f]", distance_);
}

void B200ButteraugliComparator::StartBlockComparisons() {
  if (gzb_start_block_comparisons(ctx_) != GZB_OK) Die("gzb_start_block_comparisons");
}
void B200ButteraugliComparator::FinishBlockComparisons() { gzb_finish_block_comparisons(ctx_); }

void B200ButteraugliComparator::SwitchBlock(int block_x, int block_y, int factor_x, int factor_y) {
  block_x_ = block_x;
  block_y_ = block_y;
  factor_x_ = factor_x;
  factor_y_ = factor_y;
}

// Per-call block comparison (one tiny launch per call). Correct but latency-bound: the batched
// ComputeBlockZeroingOrder below is what a --cuda build should call.
double B200ButteraugliComparator::CompareBlock(const OutputImage& img, int off_x, int off_y,
                                               const coeff_t* candidate_block, const int comp_mask) const {
  double err = 0.0;
  if (factor_x_ == 1 && factor_y_ == 1 && comp_mask == 7 && Is444(img)) {
    if (gzb_compare_block(ctx_, block_x_, block_y_, candidate_block, &err) != GZB_OK) Die("gzb_compare_block");
    return err;
  }
  // Any other sampling / component mask: the window is rendered by the caller's OutputImage
  // (guetzli/butteraugli_comparator.cc:117-126) and compared on the device.
  const int bx = block_x_ * factor_x_ + off_x, by = block_y_ * factor_y_ + off_y;
  const std::vector<uint8_t> rgb = img.ToSRGB(8 * bx, 8 * by, 8, 8);
  if (gzb_compare_block_srgb(ctx_, bx, by, rgb.data(), &err) != GZB_OK) Die("gzb_compare_block_srgb");
  return err;
}

double B200ButteraugliComparator::ScoreOutputSize(int size) const { return gzb_score_output_size(ctx_, size); }
bool B200ButteraugliComparator::DistanceOK(double target_mul) const { return gzb_distance_ok(ctx_, target_mul) != 0; }
float B200ButteraugliComparator::distmap_aggregate() const { return distance_; }
float B200ButteraugliComparator::BlockErrorLimit() const { return gzb_block_error_limit(ctx_); }

const std::vector<float> B200ButteraugliComparator::distmap() const {
  std::vector<float> d(static_cast<size_t>(width_) * height_);
  if (gzb_get_distmap(ctx_, d.data()) != GZB_OK) Die("gzb_get_distmap");
  return d;
}

void B200ButteraugliComparator::ComputeBlockErrorAdjustmentWeights(int direction, int max_block_dist,
                                                                   double target_mul, int factor_x, int factor_y,
                                                                   const std::vector<float>& distmap,
                                                                   std::vector<float>* block_weight) {
  if (factor_x != factor_y || (factor_x != 1 && factor_x != 2)) Die("ComputeBlockErrorAdjustmentWeights: sampling factors");
  // The caller's vector is all zeros on entry (guetzli/processor.cc:776); gzb overwrites it.
  if (gzb_compute_block_error_adjustment_weights_f(ctx_, direction, max_block_dist, target_mul, factor_x, distmap.data(),
                                                   block_weight->data()) != GZB_OK)
    Die("gzb_compute_block_error_adjustment_weights");
}

bool B200ButteraugliComparator::ComputeBlockZeroingOrder(const JPEGData& jpg, const OutputImage& img, int comp_mask,
                                                         std::vector<gzb_coeff_data>* output_order) {
  const int last_c = (comp_mask & 4) ? 2 : (comp_mask & 2) ? 1 : 0;   // processor.cc:565-572
  const size_t nblocks = static_cast<size_t>(img.component(last_c).width_in_blocks()) * img.component(last_c).height_in_blocks();
  output_order->assign(nblocks * 192, gzb_coeff_data{0, 0.0f});
  // SaveToJpegData drops all-zero chroma planes (guetzli/output_image.cc:588): a grey jpg has ONE component.
  // Its chroma coefficients are zeros of the layout the context expects.
  if (jpg.components.empty() || (comp_mask >> jpg.components.size()) != 0) return false;   // processor.cc:735
  const std::vector<coeff_t>* planes[3];
  std::vector<coeff_t> zeros;
  for (int c = 0; c < 3; ++c) {
    if (static_cast<size_t>(c) < jpg.components.size()) {
      planes[c] = &jpg.components[c].coeffs;
    } else {
      const size_t want = static_cast<size_t>(img.component(c).width_in_blocks()) * img.component(c).height_in_blocks() * 64;
      if (zeros.size() < want) zeros.assign(want, 0);
      planes[c] = &zeros;
    }
  }
  if (Is444(img)) {
    if (gzb_set_jpeg_coeffs(ctx_, planes[0]->data(), planes[1]->data(), planes[2]->data()) != GZB_OK) return false;
  } else {
    if (gzb_set_jpeg_coeffs_420(ctx_, planes[0]->data(), planes[1]->data(), planes[2]->data()) != GZB_OK) return false;
  }
  PushImage(img);
  return gzb_compute_block_zeroing_order(ctx_, comp_mask, output_order->data()) == GZB_OK;
}

}  // namespace guetzli
