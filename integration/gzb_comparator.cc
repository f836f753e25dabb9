// integration/gzb_comparator.cc -- see gzb_comparator.h.
#include "gzb_comparator.h"

#include <cassert>
#include <cstdio>
#include <cstdlib>

#include "guetzli/debug_print.h"

namespace guetzli {

B200ButteraugliComparator::B200ButteraugliComparator(int width, int height, const std::vector<uint8_t>* rgb,
                                                     float target_distance, ProcessStats* stats, int device)
    : width_(width), height_(height), ctx_(nullptr), stats_(stats), distance_(0.0f), block_x_(0), block_y_(0) {
  if (gzb_create(device, width, height, rgb->data(), target_distance, &ctx_) != GZB_OK) Die("gzb_create");
}

B200ButteraugliComparator::~B200ButteraugliComparator() { gzb_destroy(ctx_); }

// The reference logs accelerator errors and carries on (clguetzli/ocu.h:14); a wrong JPEG is worse
// than no JPEG, and there is no CPU fallback to hide behind: abort.
void B200ButteraugliComparator::Die(const char* what) const {
  fprintf(stderr, "gzb200: %s failed: %s\n", what, gzb_last_error(ctx_));
  abort();
}

// OutputImage keeps dequantised coefficients block-major per component: exactly gzb's layout.
void B200ButteraugliComparator::PushImage(const OutputImage& img) const {
  if (img.component(0).factor_x() != 1 || img.component(1).factor_x() != 1 || img.component(2).factor_x() != 1)
    Die("4:2:0 candidate (not supported by gzb200 yet)");
  if (gzb_set_coeffs(ctx_, img.component(0).coeffs(), img.component(1).coeffs(), img.component(2).coeffs()) != GZB_OK)
    Die("gzb_set_coeffs");
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
  if (factor_x != 1 || factor_y != 1) Die("SwitchBlock with subsampling factors");
  block_x_ = block_x;
  block_y_ = block_y;
}

// Per-call block comparison (one tiny launch per call). Correct but latency-bound: the batched
// ComputeBlockZeroingOrder below is what a --cuda build should call.
double B200ButteraugliComparator::CompareBlock(const OutputImage& img, int off_x, int off_y,
                                               const coeff_t* candidate_block, const int comp_mask) const {
  (void)img; (void)off_x; (void)off_y; (void)comp_mask;
  double err = 0.0;
  if (gzb_compare_block(ctx_, block_x_, block_y_, candidate_block, &err) != GZB_OK) Die("gzb_compare_block");
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
  if (factor_x != 1 || factor_y != 1) Die("ComputeBlockErrorAdjustmentWeights with subsampling factors");
  // The caller's vector is all zeros on entry (guetzli/processor.cc:776); gzb overwrites it.
  if (gzb_compute_block_error_adjustment_weights(ctx_, direction, max_block_dist, target_mul, distmap.data(),
                                                 block_weight->data()) != GZB_OK)
    Die("gzb_compute_block_error_adjustment_weights");
}

bool B200ButteraugliComparator::ComputeBlockZeroingOrder(const JPEGData& jpg, const OutputImage& img, int comp_mask,
                                                         std::vector<gzb_coeff_data>* output_order) {
  const size_t nblocks = static_cast<size_t>(img.component(0).width_in_blocks()) * img.component(0).height_in_blocks();
  output_order->assign(nblocks * 192, gzb_coeff_data{0, 0.0f});
  if (gzb_set_jpeg_coeffs(ctx_, jpg.components[0].coeffs.data(), jpg.components[1].coeffs.data(),
                          jpg.components[2].coeffs.data()) != GZB_OK) return false;
  PushImage(img);
  return gzb_compute_block_zeroing_order(ctx_, comp_mask, output_order->data()) == GZB_OK;
}

}  // namespace guetzli
