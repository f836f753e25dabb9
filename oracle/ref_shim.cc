// oracle/ref_shim.cc -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// C-ABI shim around the UNMODIFIED reference (yyamamoto79/guetzli-cuda-opencl)
// compiled from its own sources where they lie under /root/reference (see
// oracle/Makefile). Nothing here restates the algorithm: every function below
// forwards to the reference's own C++ entry points with g_mathMode = MODE_CPU.
// The resulting oracle/_ref/libgzref.so is the checker used by tests/, by
// __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference
// arms. The product (guetzli-cuda-opencl_b200/) never links or loads it.
//
// processor.cc is pulled in by #include so that the anonymous-namespace
// `Processor` class (ComputeBlockZeroingOrder etc., guetzli/processor.cc:52-92)
// is reachable; `private` is opened only for this translation unit.

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>
#include <string>
#include <utility>
#include <vector>

#define private public
#define protected public
#include "guetzli/processor.cc"
#undef private
#undef protected

#include "guetzli/entropy_encode.h"
#include "guetzli/gamma_correct.h"
#include "guetzli/idct.h"
#include "guetzli/fdct.h"
#include "guetzli/dct_double.h"
#include "guetzli/quality.h"
#include "guetzli/score.h"
#include "guetzli/color_transform.h"
#include "guetzli/butteraugli_comparator.h"
#include "clguetzli/clbutter_comparator.h"

using guetzli::coeff_t;
typedef std::vector<std::vector<float> > Planes;

namespace guetzli {
// Defined (non-static) in guetzli/butteraugli_comparator.cc:31 but not declared in a header.
std::vector<std::vector<float> > ComputeOpsinDynamicsImage(const int width, const int height,
                                                           const std::vector<uint8_t>& rgb);
}

namespace {

Planes ToPlanes(const float* p, size_t n) {
  Planes v(3);
  for (int c = 0; c < 3; ++c) v[c].assign(p + c * n, p + (c + 1) * n);
  return v;
}
void FromPlanes(const Planes& v, float* p, size_t n) {
  for (int c = 0; c < 3; ++c) memcpy(p + c * n, v[c].data(), n * sizeof(float));
}

// Opens the protected stage methods of the butteraugli comparator.
struct StageComparator : public butteraugli::clButteraugliComparator {
  StageComparator(size_t w, size_t h) : clButteraugliComparator(w, h, 3) {}
};

struct Session {
  int w, h;
  std::vector<uint8_t> rgb;
  guetzli::JPEGData jpg;          // q=1 coefficients (EncodeRGBToJpeg)
  guetzli::OutputImage img;
  guetzli::ProcessStats stats;
  guetzli::Params params;
  std::unique_ptr<guetzli::ButteraugliComparator> cmp;
  guetzli::Processor proc;
  guetzli::GuetzliOutput out;
  Session(int w_, int h_) : w(w_), h(h_), img(w_, h_) {}
};

}  // namespace

extern "C" {

// g_mathMode of the reference (clguetzli/clguetzli.h): 0 = MODE_CPU. Only the speed-only CUDA comparator
// build (oracle/Makefile: refcuda) ever sets another mode.
void ref_set_math_mode(int mode) { g_mathMode = static_cast<MATH_MODE>(mode); }
int ref_get_math_mode(void) { return static_cast<int>(g_mathMode); }

// ---------------------------------------------------------------- scalars
void ref_srgb8_to_linear_table(double* out256) {
  memcpy(out256, guetzli::Srgb8ToLinearTable(), 256 * sizeof(double));
}
double ref_butteraugli_score_for_quality(double q) {
  return guetzli::ButteraugliScoreForQuality(q);
}
double ref_score_jpeg(double dist, int size, double target) {
  return guetzli::ScoreJPEG(dist, size, target);
}
void ref_color_tables(int* cr_r, int* cb_b, int* cr_g, int* cb_g, uint8_t* range_limit_1024) {
  memcpy(cr_r, guetzli::kCrToRedTable, 256 * sizeof(int));
  memcpy(cb_b, guetzli::kCbToBlueTable, 256 * sizeof(int));
  memcpy(cr_g, guetzli::kCrToGreenTable, 256 * sizeof(int));
  memcpy(cb_g, guetzli::kCbToGreenTable, 256 * sizeof(int));
  memcpy(range_limit_1024, guetzli::kRangeLimitLut, 1024);
}
void ref_ycbcr_to_rgb(uint8_t* pixels, int n) {
  for (int i = 0; i < n; ++i) guetzli::ColorTransformYCbCrToRGB(pixels + 3 * i);
}

// CreateHuffmanTree (guetzli/entropy_encode.cc:68-143) on a 257-entry histogram, limit 16.
void ref_create_huffman_tree(const uint32_t* counts257, uint8_t* depth257) {
  std::vector<guetzli::HuffmanTree> tree(2 * 257 + 1);
  memset(depth257, 0, 257);
  guetzli::CreateHuffmanTree(counts257, 257, 16, &tree[0], depth257);
}

// ---------------------------------------------------------------- integer transforms
void ref_idct(const int16_t* block, uint8_t* out) { guetzli::ComputeBlockIDCT(block, out); }
void ref_fdct(int16_t* block) { guetzli::ComputeBlockDCT(block); }
void ref_dct_double(double* block) { guetzli::ComputeBlockDCTDouble(block); }
void ref_idct_double(double* block) { guetzli::ComputeBlockIDCTDouble(block); }
int ref_quantize(int coeff, int q) { return guetzli::Quantize(static_cast<coeff_t>(coeff), q); }

// ---------------------------------------------------------------- butteraugli stages
void ref_blur(float* plane, int w, int h, double sigma, double border_ratio) {
  butteraugli::Blur(w, h, plane, sigma, border_ratio);
}
// Horizontal convolution + transpose (butteraugli.cc:68-98); result has
// ceil(w/xstep) * h floats.
void ref_convolution(int w, int h, int xstep, int len, int offset, const float* mult,
                     const float* inp, double border_ratio, float* result) {
  butteraugli::_Convolution(w, h, xstep, len, offset, mult, inp, border_ratio, result);
}
void ref_opsin_dynamics_image(float* planes, int w, int h) {
  Planes v = ToPlanes(planes, size_t(w) * h);
  butteraugli::OpsinDynamicsImage(w, h, v);
  FromPlanes(v, planes, size_t(w) * h);
}
void ref_mask_high_intensity_change(const float* c0, const float* c1, int w, int h,
                                    float* out0, float* out1) {
  size_t n = size_t(w) * h;
  Planes a = ToPlanes(c0, n), b = ToPlanes(c1, n), o0 = a, o1 = b;
  butteraugli::MaskHighIntensityChange(w, h, a, b, o0, o1);
  FromPlanes(o0, out0, n);
  FromPlanes(o1, out1, n);
}
void ref_block_diff(const double* b0, const double* b1, double* dc, double* ac, double* edge_dc) {
  double x0[192], x1[192];
  memcpy(x0, b0, sizeof(x0));
  memcpy(x1, b1, sizeof(x1));
  dc[0] = dc[1] = dc[2] = ac[0] = ac[1] = ac[2] = edge_dc[0] = edge_dc[1] = edge_dc[2] = 0.0;
  butteraugli::ButteraugliBlockDiff(x0, x1, dc, ac, edge_dc);
}
void ref_diff_precompute(const float* xyb0, const float* xyb1, int w, int h, float* out) {
  size_t n = size_t(w) * h;
  Planes a = ToPlanes(xyb0, n), b = ToPlanes(xyb1, n), m;
  butteraugli::_DiffPrecompute(a, b, w, h, &m);
  FromPlanes(m, out, n);
}
void ref_average5x5(float* plane, int w, int h) {
  std::vector<float> v(plane, plane + size_t(w) * h);
  butteraugli::_Average5x5(w, h, &v);
  memcpy(plane, v.data(), v.size() * sizeof(float));
}
void ref_min_square_val(float* plane, int w, int h, int square, int offset) {
  butteraugli::_MinSquareVal(square, offset, w, h, plane);
}
void ref_mask(const float* xyb0, const float* xyb1, int w, int h, float* mask, float* mask_dc) {
  size_t n = size_t(w) * h;
  Planes a = ToPlanes(xyb0, n), b = ToPlanes(xyb1, n), m, mdc;
  butteraugli::Mask(a, b, w, h, &m, &mdc);
  FromPlanes(m, mask, n);
  FromPlanes(mdc, mask_dc, n);
}
void ref_calculate_diffmap(float* res_in_full_out, int w, int h) {
  // in: res map (ceil(w/3)*ceil(h/3) floats used) ; out: w*h floats.
  size_t rn = size_t((w + 2) / 3) * ((h + 2) / 3);
  std::vector<float> v(res_in_full_out, res_in_full_out + rn);
  butteraugli::_CalculateDiffmap(w, h, 3, &v);
  memcpy(res_in_full_out, v.data(), size_t(w) * h * sizeof(float));
}

// Full DiffmapOpsinDynamicsImage with every intermediate exposed (any out
// pointer may be NULL). Mirrors butteraugli.cc:1046-1079 call for call.
void ref_diffmap_stages(const float* xyb0_in, const float* xyb1_in, int w, int h,
                        float* mhic0, float* mhic1, float* edge_map, float* block_dc,
                        float* block_ac_pre, float* block_ac, float* mask, float* mask_dc,
                        float* combined, float* diffmap) {
  size_t n = size_t(w) * h;
  StageComparator sc(w, h);
  Planes xyb0 = ToPlanes(xyb0_in, n), xyb1 = ToPlanes(xyb1_in, n);
  {
    Planes c0 = xyb0, c1 = xyb1;
    butteraugli::MaskHighIntensityChange(w, h, c0, c1, xyb0, xyb1);
  }
  if (mhic0) FromPlanes(xyb0, mhic0, n);
  if (mhic1) FromPlanes(xyb1, mhic1, n);
  size_t rn = sc.res_xsize_ * sc.res_ysize_;
  std::vector<float> edm(3 * rn), dc(3 * rn), ac(3 * rn);
  sc.EdgeDetectorMap(xyb0, xyb1, &edm);
  if (edge_map) memcpy(edge_map, edm.data(), edm.size() * sizeof(float));
  sc.BlockDiffMap(xyb0, xyb1, &dc, &ac);
  if (block_dc) memcpy(block_dc, dc.data(), dc.size() * sizeof(float));
  if (block_ac_pre) memcpy(block_ac_pre, ac.data(), ac.size() * sizeof(float));
  sc.EdgeDetectorLowFreq(xyb0, xyb1, &ac);
  if (block_ac) memcpy(block_ac, ac.data(), ac.size() * sizeof(float));
  Planes m, mdc;
  butteraugli::Mask(xyb0, xyb1, w, h, &m, &mdc);
  if (mask) FromPlanes(m, mask, n);
  if (mask_dc) FromPlanes(mdc, mask_dc, n);
  std::vector<float> result;
  sc.CombineChannels(m, mdc, dc, ac, edm, &result);
  if (combined) memcpy(combined, result.data(), rn * sizeof(float));
  butteraugli::_CalculateDiffmap(w, h, 3, &result);
  if (diffmap) memcpy(diffmap, result.data(), n * sizeof(float));
}

void ref_diffmap(const float* xyb0_in, const float* xyb1_in, int w, int h, float* diffmap) {
  size_t n = size_t(w) * h;
  butteraugli::clButteraugliComparator c(w, h, 3);
  Planes xyb0 = ToPlanes(xyb0_in, n), xyb1 = ToPlanes(xyb1_in, n);
  std::vector<float> result;
  c.DiffmapOpsinDynamicsImage(xyb0, xyb1, result);
  memcpy(diffmap, result.data(), n * sizeof(float));
}

// sRGB8 interleaved -> XYB planes (guetzli/butteraugli_comparator.cc:31-46).
void ref_compute_opsin_dynamics_image(const uint8_t* rgb, int w, int h, float* out) {
  std::vector<uint8_t> v(rgb, rgb + size_t(3) * w * h);
  Planes p = guetzli::ComputeOpsinDynamicsImage(w, h, v);
  FromPlanes(p, out, size_t(w) * h);
}

// ---------------------------------------------------------------- encoder session
void* ref_session_new(const uint8_t* rgb, int w, int h, float target) {
  Session* s = new Session(w, h);
  s->rgb.assign(rgb, rgb + size_t(3) * w * h);
  if (!guetzli::EncodeRGBToJpeg(s->rgb, w, h, &s->jpg)) { delete s; return nullptr; }
  s->params.butteraugli_target = target;
  s->cmp.reset(new guetzli::ButteraugliComparator(w, h, &s->rgb, target, &s->stats));
  s->img.CopyFromJpegData(s->jpg);
  s->proc.params_ = s->params;
  s->proc.comparator_ = s->cmp.get();
  s->proc.final_output_ = &s->out;
  s->proc.stats_ = &s->stats;
  s->out.score = -1;
  return s;
}
void ref_session_free(void* p) { delete static_cast<Session*>(p); }
void ref_session_dims(void* p, int* bw, int* bh) {
  Session* s = static_cast<Session*>(p);
  *bw = s->img.component(0).width_in_blocks();
  *bh = s->img.component(0).height_in_blocks();
}
void ref_session_jpg_coeffs(void* p, int c, int16_t* out) {
  Session* s = static_cast<Session*>(p);
  memcpy(out, s->jpg.components[c].coeffs.data(), s->jpg.components[c].coeffs.size() * sizeof(int16_t));
}
void ref_session_reset(void* p) {
  Session* s = static_cast<Session*>(p);
  s->img.CopyFromJpegData(s->jpg);
}
void ref_session_apply_quant(void* p, const int* q192) {
  Session* s = static_cast<Session*>(p);
  int q[3][guetzli::kDCTBlockSize];
  memcpy(q, q192, sizeof(q));
  s->img.ApplyGlobalQuantization(q);
}
void ref_session_get_coeffs(void* p, int c, int16_t* out) {
  Session* s = static_cast<Session*>(p);
  const guetzli::OutputImageComponent& comp = s->img.component(c);
  memcpy(out, comp.coeffs(),
         size_t(comp.width_in_blocks()) * comp.height_in_blocks() * 64 * sizeof(int16_t));
}
void ref_session_set_coeffs(void* p, int c, const int16_t* in) {
  Session* s = static_cast<Session*>(p);
  guetzli::OutputImageComponent& comp = s->img.component(c);
  for (int by = 0, i = 0; by < comp.height_in_blocks(); ++by)
    for (int bx = 0; bx < comp.width_in_blocks(); ++bx, ++i)
      comp.SetCoeffBlock(bx, by, in + size_t(i) * 64);
}
void ref_session_to_srgb(void* p, uint8_t* out) {
  Session* s = static_cast<Session*>(p);
  std::vector<uint8_t> v = s->img.ToSRGB();
  memcpy(out, v.data(), v.size());
}
void ref_session_to_linear(void* p, float* out) {
  Session* s = static_cast<Session*>(p);
  Planes rgb(3, std::vector<float>(size_t(s->w) * s->h));
  s->img.ToLinearRGB(&rgb);
  FromPlanes(rgb, out, size_t(s->w) * s->h);
}
float ref_session_compare(void* p, float* distmap_out) {
  Session* s = static_cast<Session*>(p);
  s->cmp->Compare(s->img);
  if (distmap_out) {
    std::vector<float> d = s->cmp->distmap();
    memcpy(distmap_out, d.data(), d.size() * sizeof(float));
  }
  return s->cmp->distmap_aggregate();
}
int ref_session_jpeg_size(void* p) {
  Session* s = static_cast<Session*>(p);
  guetzli::JPEGData jpg_out = s->jpg;
  s->img.SaveToJpegData(&jpg_out);
  std::string enc;
  s->proc.OutputJpeg(jpg_out, &enc);
  return static_cast<int>(enc.size());
}
// Serialises the current candidate image to a JPEG byte stream.
int ref_session_write_jpeg(void* p, uint8_t* out, int cap) {
  Session* s = static_cast<Session*>(p);
  guetzli::JPEGData jpg_out = s->jpg;
  s->img.SaveToJpegData(&jpg_out);
  std::string enc;
  s->proc.OutputJpeg(jpg_out, &enc);
  if (static_cast<int>(enc.size()) <= cap) memcpy(out, enc.data(), enc.size());
  return static_cast<int>(enc.size());
}
// mask_xyz_ as computed by StartBlockComparisons (3*w*h floats).
void ref_session_start_block_comparisons(void* p, float* mask_xyz_out) {
  Session* s = static_cast<Session*>(p);
  s->cmp->StartBlockComparisons();
  if (mask_xyz_out) FromPlanes(s->cmp->mask_xyz_, mask_xyz_out, size_t(s->w) * s->h);
}
void ref_session_finish_block_comparisons(void* p) {
  static_cast<Session*>(p)->cmp->FinishBlockComparisons();
}
// per_block_pregamma_ of SwitchBlock(bx,by,1,1): 192 floats.
void ref_session_switch_block(void* p, int bx, int by, float* pregamma192) {
  Session* s = static_cast<Session*>(p);
  s->cmp->SwitchBlock(bx, by, 1, 1);
  if (pregamma192)
    for (int c = 0; c < 3; ++c)
      memcpy(pregamma192 + 64 * c, s->cmp->per_block_pregamma_[0][c].data(), 64 * sizeof(float));
}
// SetCoeffBlock x3 with `candidate`, CompareBlock, restore. Requires
// start_block_comparisons + switch_block(bx,by).
double ref_session_compare_block(void* p, int bx, int by, const int16_t* candidate192) {
  Session* s = static_cast<Session*>(p);
  int16_t saved[192];
  for (int c = 0; c < 3; ++c) {
    s->img.component(c).GetCoeffBlock(bx, by, saved + 64 * c);
    s->img.component(c).SetCoeffBlock(bx, by, candidate192 + 64 * c);
  }
  double r = s->cmp->CompareBlock(s->img, 0, 0, candidate192, 7);
  for (int c = 0; c < 3; ++c) s->img.component(c).SetCoeffBlock(bx, by, saved + 64 * c);
  return r;
}
// The MODE_CPU branch of SelectFrequencyMasking (processor.cc:638-672) for
// 4:4:4: out has nblocks*192 {int idx; float err} records, zero-filled.
// Requires start_block_comparisons. block_begin/block_end select a range.
// Overwrites the session's q=1 input coefficients (jpg.components[c].coeffs) -- for tests that need
// specific coefficient values (e.g. two equal zeroing-order keys in one block).
void ref_session_set_jpg_coeffs(void* p, const coeff_t* c0, const coeff_t* c1, const coeff_t* c2) {
  Session* s = static_cast<Session*>(p);
  const coeff_t* src[3] = {c0, c1, c2};
  for (int c = 0; c < 3; ++c)
    memcpy(s->jpg.components[c].coeffs.data(), src[c], s->jpg.components[c].coeffs.size() * sizeof(coeff_t));
}
void ref_session_zeroing_order(void* p, int comp_mask, int block_begin, int block_end,
                               guetzli::CoeffData* out) {
  Session* s = static_cast<Session*>(p);
  const int bw = s->img.component(0).width_in_blocks();
  for (int block_ix = block_begin; block_ix < block_end; ++block_ix) {
    const int block_x = block_ix % bw, block_y = block_ix / bw;
    coeff_t block[192] = {0};
    coeff_t orig_block[192] = {0};
    for (int c = 0; c < 3; ++c) {
      if (comp_mask & (1 << c)) {
        s->img.component(c).GetCoeffBlock(block_x, block_y, &block[c * 64]);
        const guetzli::JPEGComponent& comp = s->jpg.components[c];
        int jpg_block_ix = block_y * comp.width_in_blocks + block_x;
        memcpy(&orig_block[c * 64], &comp.coeffs[jpg_block_ix * 64], 64 * sizeof(coeff_t));
      }
    }
    std::vector<guetzli::CoeffData> order;
    s->proc.ComputeBlockZeroingOrder(block, orig_block, block_x, block_y, 1, 1,
                                     static_cast<uint8_t>(comp_mask), &s->img, &order);
    guetzli::CoeffData* q = out + size_t(block_ix - block_begin) * 192;
    memset(q, 0, 192 * sizeof(guetzli::CoeffData));
    for (size_t i = 0; i < order.size(); ++i) q[i] = order[i];
  }
}
void ref_session_block_weights(void* p, int direction, int max_block_dist, double target_mul,
                               const float* distmap, float* block_weight_inout) {
  Session* s = static_cast<Session*>(p);
  const int bw = s->img.component(0).width_in_blocks();
  const int bh = s->img.component(0).height_in_blocks();
  std::vector<float> d(distmap, distmap + size_t(s->w) * s->h);
  std::vector<float> wv(block_weight_inout, block_weight_inout + size_t(bw) * bh);
  s->cmp->ComputeBlockErrorAdjustmentWeights(direction, max_block_dist, target_mul, 1, 1, d, &wv);
  memcpy(block_weight_inout, wv.data(), wv.size() * sizeof(float));
}

// ---------------------------------------------------------------- 4:2:0 path
// Processor::DownsampleImage + OutputImage::SaveToJpegData, the head of the downsample pass of
// ProcessJpegData (processor.cc:990-997): the session's image and its jpg become YUV420.
void ref_session_downsample(void* p) {
  Session* s = static_cast<Session*>(p);
  s->img.CopyFromJpegData(s->jpg);
  s->proc.DownsampleImage(&s->img);
  s->img.SaveToJpegData(&s->jpg);
}
// Block geometry of component c: OutputImage layout, JPEGData (MCU-padded) layout, sampling factor.
void ref_session_comp_dims(void* p, int c, int* img_bw, int* img_bh, int* jpg_bw, int* jpg_bh, int* factor) {
  Session* s = static_cast<Session*>(p);
  *img_bw = s->img.component(c).width_in_blocks();
  *img_bh = s->img.component(c).height_in_blocks();
  *jpg_bw = s->jpg.components[c].width_in_blocks;
  *jpg_bh = s->jpg.components[c].height_in_blocks;
  *factor = s->img.component(c).factor_x();
}
// The MODE_CPU branch of SelectFrequencyMasking (processor.cc:638-672) with the sampling factor of
// the last component of comp_mask (1 -> 8x8 luma blocks, 6 -> 16x16 macro-blocks of a 4:2:0 image).
void ref_session_zeroing_order_f(void* p, int comp_mask, int block_begin, int block_end,
                                 guetzli::CoeffData* out) {
  Session* s = static_cast<Session*>(p);
  const int last_c = guetzli::Log2FloorNonZero(comp_mask);
  const int factor_x = s->img.component(last_c).factor_x();
  const int factor_y = s->img.component(last_c).factor_y();
  const int block_width = (s->w + 8 * factor_x - 1) / (8 * factor_x);
  for (int block_ix = block_begin; block_ix < block_end; ++block_ix) {
    const int block_x = block_ix % block_width, block_y = block_ix / block_width;
    coeff_t block[192] = {0};
    coeff_t orig_block[192] = {0};
    for (int c = 0; c < 3; ++c) {
      if (comp_mask & (1 << c)) {
        s->img.component(c).GetCoeffBlock(block_x, block_y, &block[c * 64]);
        const guetzli::JPEGComponent& comp = s->jpg.components[c];
        int jpg_block_ix = block_y * comp.width_in_blocks + block_x;
        memcpy(&orig_block[c * 64], &comp.coeffs[jpg_block_ix * 64], 64 * sizeof(coeff_t));
      }
    }
    std::vector<guetzli::CoeffData> order;
    s->proc.ComputeBlockZeroingOrder(block, orig_block, block_x, block_y, factor_x, factor_y,
                                     static_cast<uint8_t>(comp_mask), &s->img, &order);
    guetzli::CoeffData* q = out + size_t(block_ix - block_begin) * 192;
    memset(q, 0, 192 * sizeof(guetzli::CoeffData));
    for (size_t i = 0; i < order.size(); ++i) q[i] = order[i];
  }
}
void ref_session_block_weights_f(void* p, int direction, int max_block_dist, double target_mul, int factor,
                                 const float* distmap, float* block_weight_inout, int nblocks) {
  Session* s = static_cast<Session*>(p);
  std::vector<float> d(distmap, distmap + size_t(s->w) * s->h);
  std::vector<float> wv(block_weight_inout, block_weight_inout + nblocks);
  s->cmp->ComputeBlockErrorAdjustmentWeights(direction, max_block_dist, target_mul, factor, factor, d, &wv);
  memcpy(block_weight_inout, wv.data(), wv.size() * sizeof(float));
}
// guetzli::Process with Params::try_420 / force_420 (processor.h:34-42).
long ref_process_rgb_params(const uint8_t* rgb, int w, int h, float target, int try_420, int force_420,
                            uint8_t* out, long cap, char* trace, long trace_cap, int* iters) {
  guetzli::Params params;
  params.butteraugli_target = target;
  params.try_420 = try_420 != 0;
  params.force_420 = force_420 != 0;
  guetzli::ProcessStats stats;
  std::string dbg;
  if (trace) stats.debug_output = &dbg;
  std::vector<uint8_t> v(rgb, rgb + size_t(3) * w * h);
  std::string jpg;
  if (!guetzli::Process(params, &stats, v, w, h, &jpg)) return -1;
  if (static_cast<long>(jpg.size()) <= cap) memcpy(out, jpg.data(), jpg.size());
  if (trace && trace_cap > 0) {
    size_t n = std::min<size_t>(dbg.size(), trace_cap - 1);
    memcpy(trace, dbg.data(), n);
    trace[n] = 0;
  }
  if (iters) *iters = stats.counters[guetzli::kNumItersCnt];
  return static_cast<long>(jpg.size());
}

// ---------------------------------------------------------------- whole encoder
// guetzli::Process(params, stats, rgb, w, h, &out) (processor.cc:1157-1185).
// Returns the JPEG size (copied to `out` if it fits `cap`), or -1 on failure.
// `trace`, if non-NULL, receives the GUETZLI_LOG verbose trace (NUL-terminated).
long ref_process_rgb(const uint8_t* rgb, int w, int h, float target, uint8_t* out, long cap,
                     char* trace, long trace_cap, int* iters) {
  guetzli::Params params;
  params.butteraugli_target = target;
  guetzli::ProcessStats stats;
  std::string dbg;
  if (trace) stats.debug_output = &dbg;
  std::vector<uint8_t> v(rgb, rgb + size_t(3) * w * h);
  std::string jpg;
  if (!guetzli::Process(params, &stats, v, w, h, &jpg)) return -1;
  if (static_cast<long>(jpg.size()) <= cap) memcpy(out, jpg.data(), jpg.size());
  if (trace && trace_cap > 0) {
    size_t n = std::min<size_t>(dbg.size(), trace_cap - 1);
    memcpy(trace, dbg.data(), n);
    trace[n] = 0;
  }
  if (iters) *iters = stats.counters[guetzli::kNumItersCnt];
  return static_cast<long>(jpg.size());
}

}  // extern "C"
