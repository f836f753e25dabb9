// integration/gzb_comparator.h -- the reference-side binding: a guetzli::Comparator that forwards
// to libgzb200.so (include/gzb200.h). Compiled AGAINST THE REFERENCE'S HEADERS; nothing in the
// reference is modified. Inject it through the public
//   guetzli::ProcessJpegData(params, jpg, Comparator*, GuetzliOutput*, ProcessStats*)
// (guetzli/processor.h:53-55) in place of ButteraugliComparator / ButteraugliComparatorEx
// (guetzli/processor.cc:1170-1181).
#pragma once
#include <string>
#include <vector>

#include "guetzli/comparator.h"
#include "guetzli/output_image.h"
#include "guetzli/stats.h"
#include "gzb200.h"

namespace guetzli {

class B200ButteraugliComparator : public Comparator {
 public:
  // Same arguments as ButteraugliComparator (guetzli/butteraugli_comparator.h:35-37) + device.
  B200ButteraugliComparator(int width, int height, const std::vector<uint8_t>* rgb,
                            float target_distance, ProcessStats* stats, int device = 0);
  ~B200ButteraugliComparator() override;

  void Compare(const OutputImage& img) override;
  void StartBlockComparisons() override;
  void FinishBlockComparisons() override;
  void SwitchBlock(int block_x, int block_y, int factor_x, int factor_y) override;
  double CompareBlock(const OutputImage& img, int off_x, int off_y, const coeff_t* candidate_block,
                      const int comp_mask) const override;
  double ScoreOutputSize(int size) const override;
  bool DistanceOK(double target_mul) const override;
  const std::vector<float> distmap() const override;
  float distmap_aggregate() const override;
  float BlockErrorLimit() const override;
  void ComputeBlockErrorAdjustmentWeights(int direction, int max_block_dist, double target_mul,
                                          int factor_x, int factor_y,
                                          const std::vector<float>& distmap,
                                          std::vector<float>* block_weight) override;

  // The batched call the reference's own GPU modes use instead of the per-block loop
  // (cuComputeBlockZeroingOrder, guetzli/processor.cc:618-632): fills `output_order`
  // (num_blocks * 192 CoeffData, zero-filled) for the coefficients of `img` against `jpg`.
  bool ComputeBlockZeroingOrder(const JPEGData& jpg, const OutputImage& img, int comp_mask,
                                std::vector<gzb_coeff_data>* output_order);

  gzb_ctx* context() const { return ctx_; }

 private:
  void Die(const char* what) const;
  void PushImage(const OutputImage& img) const;
  int width_, height_;
  gzb_ctx* ctx_;
  ProcessStats* stats_;
  float distance_;
  int block_x_, block_y_, factor_x_, factor_y_;
};

}  // namespace guetzli
