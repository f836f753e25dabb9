// gzb200.cu -- context, plans and the C ABI (include/gzb200.h) over the sm_100a kernels.
// Single translation unit: nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <ctime>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/gzb200.h"
#include "gzb_kernels.cuh"
#include "gzb_zeroing.cuh"
#include "gzb_huffman.cuh"
#include "gzb_yuv420.cuh"
#include "gzb_backend.cuh"

namespace gzb {

// ---------------------------------------------------------------------------------------------
// Host-side constants: blur kernels (butteraugli.cc:100-112), LUTs.
// ---------------------------------------------------------------------------------------------
enum BlurKind { kB11 = 0, kB15, kB0586, kB04, kB14, kB9657, kB14264, kB4533, kB8851, kBUser, kNumBlurKinds };
static const double kSigmas[kNumBlurKinds] = {1.1, 1.5, 0.586, 0.4, 14.0, 9.65781083553,
                                              14.2644604355, 4.53358927369, 8.8510880283, 0.0};
struct HostKernel { int r, step; float taps[kMaxTaps]; };

static bool make_host_kernel(double sigma, HostKernel* k) {
  const double scaler = -1.0 / (2 * sigma * sigma);
  k->r = std::max<int>(1, static_cast<int>(2.25 * std::fabs(sigma)));
  if (2 * k->r + 1 > kMaxTaps) return false;
  for (int i = -k->r; i <= k->r; ++i) k->taps[i + k->r] = static_cast<float>(std::exp(scaler * i * i));
  k->step = std::max(1, static_cast<int>(sigma / 3));
  return true;
}

// scale = 1 / interpolated in-range tap weight (butteraugli.cc:76-89) at position `pos` of a
// line of `size` samples.
static double border_scale(const HostKernel& k, int pos, int size, double border_ratio) {
  double full = 0.0;
  for (int j = 0; j <= 2 * k.r; ++j) full += k.taps[j];
  const int lo = std::max(0, pos - k.r);
  const int hi = std::min(size, pos + k.r + 1) - 1;
  double w = 0.0;
  for (int j = lo; j <= hi; ++j) w += k.taps[j - pos + k.r];
  w = (1.0 - border_ratio) * w + border_ratio * full;
  return 1.0 / w;
}

static HostKernel g_hk[kNumBlurKinds];
static std::mutex g_init_mu;
static bool g_dev_ready[64];

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      char buf_[512];                                                                   \
      snprintf(buf_, sizeof(buf_), "%s:%d %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
      throw std::string(buf_);                                                          \
    }                                                                                   \
  } while (0)

static void init_device_tables(int device) {
  std::lock_guard<std::mutex> lock(g_init_mu);
  if (device < 64 && g_dev_ready[device]) return;
  for (int k = 0; k < kBUser; ++k) make_host_kernel(kSigmas[k], &g_hk[k]);
  DeviceTables* t = new DeviceTables;
  // 21-entry ramps built by repeated addition (butteraugli.cc:200-247)
  const double first[3] = {11.38708334481672, 1.4103373714040413, 5.2511644570349185};
  const double inc[3] = {14.550189611520716, 0.7084088867024, 5.2511644570349185};
  for (int l = 0; l < 3; ++l) {
    t->lut21[l][0] = 0.0;
    t->lut21[l][1] = first[l];
    for (int i = 2; i < 21; ++i) t->lut21[l][i] = t->lut21[l][i - 1] + inc[l];
  }
  // MakeMask (butteraugli.cc:1242-1254): {extmul, extoff, offset, scaler, mul}
  static const double MP[6][5] = {
      {0.975741017749, -4.25328244168, 0.454909521427, 0.0738288224836, 20.8029176447},
      {0.373995618954, 1.5307267433, 0.911952641929, 1.1731667845, 16.2447033988},
      {0.61582234137, -4.25376118646, 1.05105070921, 0.47434643535, 31.1444967089},
      {1.79116943438, -3.86797479189, 0.670960225853, 0.486575865525, 20.4563479139},
      {0.212223514236, -3.65647120524, 1.73396799447, 0.170392660501, 21.6566724788},
      {0.349376011816, -0.894711072781, 0.901647926679, 0.380086095024, 18.0373825149}};
  for (int m = 0; m < 6; ++m)
    for (int i = 0; i < 512; ++i) {
      const double c = MP[m][4] / ((0.01 * MP[m][3] * i) + MP[m][2]);
      double v = 1.0 + MP[m][0] * (c + MP[m][1]);
      v *= v;
      t->mask_lut[m][i] = v;
    }
  // Srgb8ToLinearTable (guetzli/gamma_correct.cc:23-33), stored as the float the reference casts to
  for (int i = 0; i < 256; ++i) {
    const double v = i < 11 ? i / 12.92 : 255.0 * std::pow(((i / 255.0) + 0.055) / 1.055, 2.4);
    t->srgb_lin[i] = static_cast<float>(v);
  }
  CK(cudaMemcpyToSymbol(g_tab, t, sizeof(DeviceTables)));
  delete t;
  float taps[12][kMaxTaps];
  memset(taps, 0, sizeof(taps));
  for (int k = 0; k < kBUser; ++k) memcpy(taps[k], g_hk[k].taps, sizeof(float) * (2 * g_hk[k].r + 1));
  CK(cudaMemcpyToSymbol(c_taps, taps, sizeof(taps)));
  float csf[192], bias[192];
  for (int i = 0; i < 192; ++i) { csf[i] = kGzbZeroModel[i].csf; bias[i] = kGzbZeroModel[i].bias; }
  CK(cudaMemcpyToSymbol(c_zero_csf, csf, sizeof(csf)));
  CK(cudaMemcpyToSymbol(c_zero_bias, bias, sizeof(bias)));
  double s8[8];
  for (int x = 0; x < 8; ++x) s8[x] = border_scale(g_hk[kB11], x, 8, 0.0);
  CK(cudaMemcpyToSymbol(c_scale8, s8, sizeof(s8)));
  CK(cudaMemcpyToSymbol(c_taps11, g_hk[kB11].taps, 5 * sizeof(float)));
  {  // kDCTMatrix of guetzli/dct_double.cc:28-45: 0.5*alpha(u)*cos((2x+1)u*pi/16) with 10 decimals
    double m[64];
    for (int u = 0; u < 8; ++u)
      for (int x = 0; x < 8; ++x) {
        const double v = 0.5 * (u == 0 ? std::sqrt(0.5) : 1.0) * std::cos((2 * x + 1) * u * M_PI / 16);
        char buf[64];
        snprintf(buf, sizeof(buf), "%.10f", v);
        m[8 * u + x] = strtod(buf, nullptr);
      }
    CK(cudaMemcpyToSymbol(c_dct_matrix, m, sizeof(m)));
  }
  if (device < 64) g_dev_ready[device] = true;
}

// ---------------------------------------------------------------------------------------------
// Blur plans
// ---------------------------------------------------------------------------------------------
struct BlurPlan {
  BlurGeom g{};
  double* d_sx = nullptr;
  double* d_sy = nullptr;
  bool own = false;            // tables allocated with their own cudaMalloc (stage entry points)
  CUtensorMap tmap;            // TMA descriptor of the H pass's input planes (when has_tmap)
  bool has_tmap = false;
  int box_w = 0;
  CUtensorMap tmap_v;          // TMA descriptor of the V pass's input: the H-pass scratch at tmap_v_base
  bool has_tmap_v = false;
  int box_rows = 0;
  const float* tmap_v_base = nullptr;
  std::vector<double> hx, hy;  // host copies of the scale tables
  void build(const HostKernel& hk, int kind, int in_w, int in_h, int in_pitch, int x0, int sx, int nx,
             int y0, int sy, int ny, double border_ratio, int ups) {
    g.in_w = in_w; g.in_h = in_h; g.in_pitch = in_pitch; g.r = hk.r; g.kind = kind;
    g.x0 = x0; g.sx = sx; g.nx = nx; g.y0 = y0; g.sy = sy; g.ny = ny;
    g.tmp_pitch = round_up(std::max(nx, 1), 32); g.ups = ups;
    g.oxn = std::max(1, std::min(kBhOx, (kBhMaxSpan - 2 * hk.r - 1) / sx + 1));
    g.oyn = std::max(1, std::min(kBvOy, (kBvMaxRows - 2 * hk.r - 1) / sy + 1));
    hx.assign(std::max(nx, 1), 0.0);
    hy.assign(std::max(ny, 1), 0.0);
    for (int i = 0; i < nx; ++i) hx[i] = border_scale(hk, x0 + i * sx, in_w, border_ratio);
    for (int i = 0; i < ny; ++i) hy[i] = border_scale(hk, y0 + i * sy, in_h, border_ratio);
  }
  // The plain decimated blur of the reference: lattice (0, step).
  void build_decimated(const HostKernel& hk, int kind, int in_w, int in_h, int in_pitch,
                       double border_ratio, int ups) {
    build(hk, kind, in_w, in_h, in_pitch, 0, hk.step, (in_w + hk.step - 1) / hk.step, 0, hk.step,
          (in_h + hk.step - 1) / hk.step, border_ratio, ups);
  }
  void upload(cudaStream_t st) {  // d_sx / d_sy already point into the context slab
    CK(cudaMemcpyAsync(d_sx, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_sy, hy.data(), hy.size() * sizeof(double), cudaMemcpyHostToDevice, st));
  }
  void upload_own() {
    own = true;
    CK(cudaMalloc(&d_sx, hx.size() * sizeof(double)));
    CK(cudaMalloc(&d_sy, hy.size() * sizeof(double)));
    CK(cudaMemcpy(d_sx, hx.data(), hx.size() * sizeof(double), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(d_sy, hy.data(), hy.size() * sizeof(double), cudaMemcpyHostToDevice));
  }
  void release() {
    if (own) {
      if (d_sx) cudaFree(d_sx);
      if (d_sy) cudaFree(d_sy);
    }
    d_sx = d_sy = nullptr;
  }
  size_t tmp_floats() const { return static_cast<size_t>(g.tmp_pitch) * g.in_h; }
  size_t out_floats() const { return static_cast<size_t>(g.tmp_pitch) * std::max(g.ny, 1); }
};

// TMA descriptor for the H pass of a plan whose input is `planes` float planes of in_pitch x in_h (plane stride
// `plane_stride` floats) at `in`: a 3-D tiled map (x, y, plane) with a box of box_w x kBhRows x 1 and zero
// fill outside [0, in_w) x [0, in_h). cuTensorMapEncodeTiled is looked up through the runtime, so the library
// does not link against libcuda. Returns false (the scalar loader is used) if anything is not as required.
typedef CUresult (*TensorMapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                      const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                      CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static bool make_blur_tmap(BlurPlan* pl, const float* in, size_t plane_stride, int planes) {
  static const bool disabled = getenv("GZB_NO_TMA") != nullptr;
  pl->has_tmap = false;
  if (disabled || pl->g.ups != 1 || pl->g.nx <= 0) return false;
  static TensorMapEncodeFn encode = [] {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess) fn = nullptr;
    cudaGetLastError();
    return reinterpret_cast<TensorMapEncodeFn>(fn);
  }();
  if (!encode) return false;
  const int span = (pl->g.oxn - 1) * pl->g.sx + 2 * pl->g.r + 1;
  const int box_w = (span + 3 + 3) & ~3;   // the box starts at a 16-byte boundary up to 3 floats left of the span
  if (box_w > kBhTmaMaxBox || (reinterpret_cast<uintptr_t>(in) & 15) || (pl->g.in_pitch & 3) || (plane_stride & 3)) return false;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(pl->g.in_w), static_cast<cuuint64_t>(pl->g.in_h), static_cast<cuuint64_t>(planes)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(pl->g.in_pitch) * sizeof(float),
                                 static_cast<cuuint64_t>(std::max<size_t>(plane_stride, static_cast<size_t>(pl->g.in_pitch) * pl->g.in_h)) * sizeof(float)};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_w), static_cast<cuuint32_t>(kBhRows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (encode(&pl->tmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(in), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  pl->box_w = box_w;
  pl->has_tmap = true;
  return true;
}
static bool make_blur_tmap_v(BlurPlan* pl, const float* tmp, int planes) {
  static const bool disabled = getenv("GZB_NO_TMA") != nullptr;
  pl->has_tmap_v = false;
  if (disabled || pl->g.nx <= 0 || pl->g.ny <= 0) return false;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess ||
      q != cudaDriverEntryPointSuccess || !fn) { cudaGetLastError(); return false; }
  const int rows = (pl->g.oyn - 1) * pl->g.sy + 2 * pl->g.r + 1;
  if (rows > kBvMaxRows || (reinterpret_cast<uintptr_t>(tmp) & 15) || (pl->g.tmp_pitch & 3)) return false;
  const cuuint64_t gdim[3] = {static_cast<cuuint64_t>(pl->g.nx), static_cast<cuuint64_t>(pl->g.in_h), static_cast<cuuint64_t>(planes)};
  const cuuint64_t gstride[2] = {static_cast<cuuint64_t>(pl->g.tmp_pitch) * sizeof(float),
                                 static_cast<cuuint64_t>(pl->tmp_floats()) * sizeof(float)};
  const cuuint32_t box[3] = {32, static_cast<cuuint32_t>(rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  if (reinterpret_cast<TensorMapEncodeFn>(fn)(&pl->tmap_v, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(tmp), gdim, gstride, box,
                                              estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) return false;
  pl->box_rows = rows;
  pl->tmap_v_base = tmp;
  pl->has_tmap_v = true;
  return true;
}

// The same for the V pass: its input is the H-pass scratch (nx columns at pitch tmp_pitch, in_h rows, `planes`
// planes of tmp_floats() each) at `tmp`; box = 32 columns x rows x 1.
static bool make_blur_tmap_v(BlurPlan* pl, const float* tmp, int planes);

// ---------------------------------------------------------------------------------------------
// Device slabs: one cudaMalloc per context, sub-allocated; released slabs are cached per device so
// that back-to-back encodes do not pay the driver's map/unmap cost (the reference's cumem_pool
// plays this role, clguetzli/cumem_pool.cpp:28-112).
// ---------------------------------------------------------------------------------------------
struct Slab {
  void* base = nullptr; size_t cap = 0; void* pinned = nullptr; void* pinned2 = nullptr; size_t pinned2_cap = 0;
  void* be = nullptr; size_t be_cap = 0;            // back-end arrays (gzb_backend.cuh), sized by the candidate count
  void* be_pinned = nullptr; size_t be_pinned_cap = 0;
};
static void slab_free(Slab& s) {
  cudaFree(s.base); cudaFreeHost(s.pinned);
  if (s.pinned2) cudaFreeHost(s.pinned2);
  if (s.be) cudaFree(s.be);
  if (s.be_pinned) cudaFreeHost(s.be_pinned);
  s = Slab();
}
static std::mutex g_slab_mu;
static std::vector<Slab> g_slab_cache[64];

static Slab slab_acquire(int device, size_t bytes) {
  {
    std::lock_guard<std::mutex> lock(g_slab_mu);
    std::vector<Slab>& v = g_slab_cache[device & 63];
    int best = -1;
    for (size_t i = 0; i < v.size(); ++i)
      if (v[i].cap >= bytes && (best < 0 || v[i].cap < v[best].cap)) best = static_cast<int>(i);
    if (best >= 0 && v[best].cap <= bytes + bytes / 2 + (64u << 20)) {
      Slab s = v[best];
      v.erase(v.begin() + best);
      return s;
    }
    for (Slab& s : v) slab_free(s);  // wrong sizes: drop them
    v.clear();
  }
  Slab s;
  s.cap = bytes;
  CK(cudaMalloc(&s.base, bytes));
  CK(cudaMallocHost(&s.pinned, 4096));
  return s;
}
static void slab_release(int device, Slab s) {
  if (!s.base) return;
  std::lock_guard<std::mutex> lock(g_slab_mu);
  std::vector<Slab>& v = g_slab_cache[device & 63];
  if (v.size() >= 4) { slab_free(s); return; }
  v.push_back(s);
}

}  // namespace gzb

using namespace gzb;

// ---------------------------------------------------------------------------------------------
// Per-kernel event profiling (off by default)
// ---------------------------------------------------------------------------------------------
enum KClass { KC_IDCT = 0, KC_OPSIN, KC_MHIC, KC_BLUR_H, KC_BLUR_V, KC_EDGE_MAP, KC_BLOCK_DIFF, KC_LOWFREQ,
              KC_MASK_FRONT, KC_COMBINE, KC_DIFFMAP_FINAL, KC_ZEROING, KC_BLOCK_MASK, KC_WEIGHTS, KC_MISC, KC_HUFFMAN, KC_COUNT };
static const char* const kClassNames[KC_COUNT] = {
    "k_coeffs_to_rgb8", "k_opsin_dynamics", "k_mask_high_intensity_change", "k_blur_h", "k_blur_v",
    "k_edge_detector_map", "k_block_diff_map", "k_edge_lowfreq", "k_mask_front", "k_combine",
    "k_diffmap_final", "k_zeroing_order", "k_block_mask_scale", "k_block_weights", "misc", "k_huffman"};
struct Prof {
  bool on = false;
  struct Pend { int k; cudaEvent_t a, b; };
  std::vector<cudaEvent_t> free_ev;
  std::vector<Pend> pend;
  double ms[KC_COUNT] = {0};
  unsigned long long n[KC_COUNT] = {0};
  cudaEvent_t get() {
    if (!free_ev.empty()) { cudaEvent_t e = free_ev.back(); free_ev.pop_back(); return e; }
    cudaEvent_t e;
    cudaEventCreate(&e);
    return e;
  }
};

// ---------------------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------------------
struct gzb_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t stream2 = nullptr;   // entropy coding of the candidate, concurrent with its Compare
  cudaStream_t stream_b = nullptr, stream_l = nullptr;   // BlockDiffMap / EdgeDetectorLowFreq branches of a Compare
  cudaEvent_t ev_fork = nullptr, ev_bdm = nullptr, ev_lf = nullptr;
  cudaEvent_t ev_cand = nullptr;    // the candidate (coefficients, quantiser, samples) is complete on the main stream
  bool concurrent = false;          // the branches' blur scratch regions fit side by side in d_tmp
  size_t tmp_main_off = 0;          // floats: where the main stream's blur scratch starts in d_tmp
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int W = 0, H = 0, P = 0, HP = 0, bw = 0, bh = 0, nblocks = 0, rxs = 0, rys = 0, sqp = 0;
  size_t ps = 0;       // floats per full-res plane (P * HP)
  size_t cs = 0;       // coefficients per component slot of d_orig / d_coef
  // YUV 4:2:0 (after gzb_downsample_420): chroma has mcw x mch blocks, luma is MCU-padded to cbw0 x cbh0
  bool mode420 = false;
  int mcw = 0, mch = 0, cbw0 = 0, cbh0 = 0;
  uint8_t* d_ycc = nullptr;    // [3] u8 planes: Y samples, Cb / Cr half-resolution samples (4:2:0 only)
  uint8_t* d_cup = nullptr;    // [2] u8 planes: upsampled Cb / Cr of the candidate (4:2:0 only)
  float target = 0.f;
  float distance = 0.f;
  float last_ms = 0.f;
  unsigned zeroing_tie_blocks = 0;   // blocks of the last zeroing search whose candidate keys tied (sorted the library's way)
  unsigned long long launches = 0, h2d_bytes = 0, d2h_bytes = 0;
  Slab slab;
  struct Req { void** pp; size_t bytes; };
  std::vector<Req> reqs;
  bool upd_own = false;
  bool packed_valid = false;   // packed zeroing candidates still sit in d_tmp
  int packed_mask = 0, packed_b0 = 0, packed_b1 = 0;
  Prof prof;
  bool have_orig_coeffs = false, have_coeffs = false, block_cmp = false, have_distmap = false;
  bool compare_pending = false;
  cudaGraphExec_t cmp_graph_exec = nullptr;   // the Compare pipeline as a graph (gzb_compare_begin)
  unsigned long long cmp_graph_launches = 0;
  bool graph_failed = false;
  int sm_count = 148;
  std::string err;

  uint8_t* d_rgb0 = nullptr;   // original, planar u8 [3][HP][P]
  uint8_t* d_rgb1 = nullptr;   // candidate
  uint8_t* d_stage_u8 = nullptr;  // interleaved staging (w*h*3)
  int16_t* d_orig = nullptr;   // [3][nblocks*64] q=1 coefficients
  int16_t* d_coef = nullptr;   // [3][nblocks*64] candidate
  float* d_xyb0 = nullptr;     // [3] planes
  float* d_xyb1 = nullptr;     // [3]
  float* d_mh = nullptr;       // [6] planes: MaskHighIntensityChange outputs (img0 x3, img1 x3)
  float* d_bl = nullptr;       // [6] planes: edge-detector blurs; reused as mask front [3]
  float* d_tmp = nullptr;      // [6] planes: H-pass scratch
  float* d_lf = nullptr;       // [6] decimated sigma-14 maps
  float* d_ms[3] = {nullptr, nullptr, nullptr};  // blurred mask lattices
  float* d_msb2 = nullptr;     // channel-2 mask lattice for block comparisons
  float* d_edm = nullptr, *d_dc = nullptr, *d_ac = nullptr;  // res maps x3 floats
  float* d_sq = nullptr;       // res map, pitch sqp
  float* d_dsmall = nullptr;   // decimated diffmap blur
  float* d_diffmap = nullptr;  // [HP][P]
  float* d_bmax = nullptr, *d_weight = nullptr, *d_mask_scale = nullptr, *d_block_err = nullptr;
  float* d_pregamma = nullptr;
  unsigned char* d_flags = nullptr;
  unsigned int* d_scalars = nullptr;  // [0] distance bits, [1] work counter, [2] grey flag, [3] tie counter, [4] block-change record valid
  uint8_t* d_blk_changed = nullptr;   // one byte per 8x8 block: samples changed since the last Compare (BlockChanges)
  bool blk_tracking = false;          // every change since the last Compare is recorded in d_blk_changed
  int* d_q = nullptr;          // 192 ints
  uint8_t* d_upd = nullptr;    // sparse-update staging: [cap] int32 | [cap] int16 | [cap] u8
  size_t upd_cap = 0;
  gzb_coeff_data* d_order = nullptr;
  float* h_pinned = nullptr;   // small pinned staging
  size_t lf_stride = 0;

  BlurPlan p_ops, p_ed[3], p_lf, p_mk[3], p_mkb2, p_dm;

  // ---- incremental Compare (see DirtyMask in gzb_kernels.cuh) ----
  uint8_t* d_dirty = nullptr;        // [DS_COUNT][mtw * mth] per-stage dirty tiles
  uint8_t* h_dirty = nullptr;        // pinned staging of the same
  int mtw = 0, mth = 0;
  bool dirty_is_all = false;         // d_dirty currently holds all ones
  unsigned int* d_ctamax = nullptr;  // per-CTA maxima of k_diffmap_final (persist between Compares)
  int n_ctamax = 0;
  float* d_lft = nullptr;            // EdgeDetectorLowFreq term, res map x3 floats
  bool inter_valid = false;          // the buffers of the last Compare describe the candidate before `changed`
  bool changed_overflow = false;     // too many changes since the last Compare to track
  std::vector<int4> changed;         // pixel rectangles (x0, y0, x1, y1) whose samples changed since the last Compare
  unsigned long long incremental_compares = 0, fine_bdm_compares = 0;

  // ---- SelectFrequencyBackEnd on the device (gzb_backend.cuh); pointers into slab.be ----
  struct Backend {
    bool active = false;
    int comp_mask = 0, factor = 1, num_blocks = 0, total = 0;
    BeGeom geom{};
    BeState* st = nullptr;
    BeEntry* small = nullptr;          // directly behind *st: one copy brings both to the host
    int* cand_off = nullptr; uint8_t* cand_idx = nullptr; float* cand_err = nullptr;
    int* last_index = nullptr; float* max_err = nullptr;
    int* counts = nullptr; int* offsets = nullptr; unsigned* pcount = nullptr; int* chunk_sums = nullptr;
    BeEntry* order = nullptr; unsigned* lpos = nullptr; unsigned* rpos = nullptr; unsigned* tcl = nullptr; unsigned* tcr = nullptr;
    int* req_blocks = nullptr; BeBlockState* req_out = nullptr;
    unsigned int* hist = nullptr;      // [48 + 768]
    unsigned* flag = nullptr;
    int select_grid = 0;
    BeRange h_stack[kBeStack];         // host copy of the pending ranges after the last select
    int h_top = 0;
    unsigned long long n = 0;          // entries of the current order
    unsigned long long selects = 0, levels = 0, host_ranges = 0;
  } be;
};

enum DirtyStage { DS_OPS = 0, DS_MHIC, DS_EB, DS_EDM, DS_BDM, DS_LFH, DS_LFV, DS_LOW, DS_MKH0, DS_MKV0, DS_MKH1, DS_MKV1,
                  DS_MKH2, DS_MKV2, DS_COMB, DS_DMH, DS_DMV, DS_FIN, DS_COUNT };

static thread_local std::string g_create_err;

namespace {

// Launch wrapper: counts the launch and, when profiling, brackets it with an event pair.
#define KLAUNCH_S(c, strm, kclass, ...)                                \
  do {                                                                 \
    cudaEvent_t ea_ = nullptr;                                         \
    if ((c)->prof.on) { ea_ = (c)->prof.get(); cudaEventRecord(ea_, (strm)); } \
    __VA_ARGS__;                                                       \
    if ((c)->prof.on) { cudaEvent_t eb_ = (c)->prof.get(); cudaEventRecord(eb_, (strm)); \
                        (c)->prof.pend.push_back({kclass, ea_, eb_}); } \
    (c)->launches += 1;                                                \
  } while (0)
#define KLAUNCH(c, kclass, ...) KLAUNCH_S(c, (c)->stream, kclass, __VA_ARGS__)

inline DirtyMask dmask(const gzb_ctx* c, int stage) {
  DirtyMask d;
  d.m = c->d_dirty ? c->d_dirty + static_cast<size_t>(stage) * c->mtw * c->mth : nullptr;
  d.tw = c->mtw; d.th = c->mth;
  return d;
}
const DirtyMask kAllDirty = {nullptr, 0, 0};

void prof_resolve(gzb_ctx* c) {
  for (auto& p : c->prof.pend) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, p.a, p.b) == cudaSuccess) { c->prof.ms[p.k] += ms; c->prof.n[p.k] += 1; }
    c->prof.free_ev.push_back(p.a);
    c->prof.free_ev.push_back(p.b);
  }
  c->prof.pend.clear();
}

template <typename T>
void dmalloc(T** p, size_t n) { CK(cudaMalloc(reinterpret_cast<void**>(p), std::max<size_t>(n, 1) * sizeof(T))); }

void run_blur(gzb_ctx* c, const BlurPlan& pl, const float* in, size_t in_stride, int planes,
              float* out, size_t out_stride, int out_pitch, cudaStream_t st = nullptr, float* tmp = nullptr,
              DirtyMask mh = kAllDirty, DirtyMask mv = kAllDirty) {
  const BlurGeom& g = pl.g;
  c->packed_valid = false;
  if (g.nx <= 0 || g.ny <= 0) return;
  if (!st) st = c->stream;
  if (!tmp) tmp = c->d_tmp + c->tmp_main_off;
  dim3 blk(32, 8);
  dim3 gh((g.nx + g.oxn - 1) / g.oxn, (g.in_h + kBhRows - 1) / kBhRows, planes);
  const size_t tstride = pl.tmp_floats();
  if (pl.has_tmap) KLAUNCH_S(c, st, KC_BLUR_H, k_blur_h_tma<<<gh, blk, 0, st>>>(pl.tmap, g, pl.box_w, pl.d_sx, tmp, tstride, mh));
  else if (g.ups == 1) KLAUNCH_S(c, st, KC_BLUR_H, k_blur_h<1><<<gh, blk, 0, st>>>(in, in_stride, g, pl.d_sx, tmp, tstride, mh));
  else KLAUNCH_S(c, st, KC_BLUR_H, k_blur_h<3><<<gh, blk, 0, st>>>(in, in_stride, g, pl.d_sx, tmp, tstride, mh));
  dim3 gv((g.nx + 31) / 32, (g.ny + g.oyn - 1) / g.oyn, planes);
  if (pl.has_tmap_v && pl.tmap_v_base == tmp)
    KLAUNCH_S(c, st, KC_BLUR_V, k_blur_v_tma<<<gv, blk, 0, st>>>(pl.tmap_v, g, pl.box_rows, pl.d_sy, out, out_stride, out_pitch, mv));
  else
    KLAUNCH_S(c, st, KC_BLUR_V, k_blur_v<<<gv, blk, 0, st>>>(tmp, tstride, g, pl.d_sy, out, out_stride, out_pitch, mv));
}

// Block geometry of component k's coefficient array in the context's current sampling mode.
inline int comp_bw(const gzb_ctx* c, int k) { return !c->mode420 ? c->bw : (k == 0 ? c->cbw0 : c->mcw); }
inline int comp_bh(const gzb_ctx* c, int k) { return !c->mode420 ? c->bh : (k == 0 ? c->cbh0 : c->mch); }
inline size_t comp_blocks(const gzb_ctx* c, int k) { return static_cast<size_t>(comp_bw(c, k)) * comp_bh(c, k); }

void render_candidate(gzb_ctx* c, int op) {
  c->packed_valid = false;
  const size_t cs = c->cs;
  const size_t us = static_cast<size_t>(c->P) * c->HP;
  if (c->mode420) {
    // per component: (quantise | scale) + IDCT -> Y samples / half-resolution chroma samples, then
    // the fancy upsampling + colour conversion of OutputImage::ToSRGB
    for (int k = 0; k < 3; ++k) {
      const int nb = static_cast<int>(comp_blocks(c, k));
      const int grid = (nb + 31) / 32;
      const int16_t* src = ((op == kCoeffScale || op == kCoeffQuantizeSrc) ? c->d_orig : c->d_coef) + k * cs;
      int16_t* dst = c->d_coef + k * cs;
      uint8_t* plane = c->d_ycc + k * us;
      if (op == kCoeffKeep)
        KLAUNCH(c, KC_IDCT, k420_idct_comp<kCoeffKeep><<<grid, 256, 0, c->stream>>>(src, dst, c->d_q + 64 * k, comp_bw(c, k), nb, c->P, plane));
      else if (op == kCoeffQuantize)
        KLAUNCH(c, KC_IDCT, k420_idct_comp<kCoeffQuantize><<<grid, 256, 0, c->stream>>>(src, dst, c->d_q + 64 * k, comp_bw(c, k), nb, c->P, plane));
      else if (op == kCoeffQuantizeSrc)
        KLAUNCH(c, KC_IDCT, k420_idct_comp<kCoeffQuantizeSrc><<<grid, 256, 0, c->stream>>>(src, dst, c->d_q + 64 * k, comp_bw(c, k), nb, c->P, plane));
      else
        KLAUNCH(c, KC_IDCT, k420_idct_comp<kCoeffScale><<<grid, 256, 0, c->stream>>>(src, dst, c->d_q + 64 * k, comp_bw(c, k), nb, c->P, plane));
    }
    dim3 grd((c->W + 1023) / 1024, c->H);
    KLAUNCH(c, KC_IDCT, k420_render<<<grd, 256, 0, c->stream>>>(c->d_ycc, us, c->W, c->H, c->P, c->d_rgb1, c->d_cup));
    CK(cudaEventRecord(c->ev_cand, c->stream));
    return;
  }
  const int grid = (c->nblocks + 31) / 32;
  if (op == kCoeffKeep)
    KLAUNCH(c, KC_IDCT, k_coeffs_to_rgb8<kCoeffKeep><<<grid, 256, 0, c->stream>>>(c->d_coef, c->d_coef, cs, c->d_q, c->bw, c->nblocks, c->P, c->d_rgb1, us));
  else if (op == kCoeffQuantize)
    KLAUNCH(c, KC_IDCT, k_coeffs_to_rgb8<kCoeffQuantize><<<grid, 256, 0, c->stream>>>(c->d_coef, c->d_coef, cs, c->d_q, c->bw, c->nblocks, c->P, c->d_rgb1, us));
  else if (op == kCoeffQuantizeSrc)
    KLAUNCH(c, KC_IDCT, k_coeffs_to_rgb8<kCoeffQuantizeSrc><<<grid, 256, 0, c->stream>>>(c->d_orig, c->d_coef, cs, c->d_q, c->bw, c->nblocks, c->P, c->d_rgb1, us));
  else
    KLAUNCH(c, KC_IDCT, k_coeffs_to_rgb8<kCoeffScale><<<grid, 256, 0, c->stream>>>(c->d_orig, c->d_coef, cs, c->d_q, c->bw, c->nblocks, c->P, c->d_rgb1, us));
  CK(cudaEventRecord(c->ev_cand, c->stream));
}

void opsin_from_u8(gzb_ctx* c, const uint8_t* planes, float* xyb, bool masked = false) {
  dim3 blk(32, 8), grd((c->W + 31) / 32, (c->H + 31) / 32);
  KLAUNCH(c, KC_OPSIN, k_opsin_dynamics<<<grd, blk, 0, c->stream>>>(planes, static_cast<size_t>(c->P) * c->HP, c->W, c->H, c->P,
                                               c->p_ops.d_sx, c->p_ops.d_sy, xyb, c->ps, masked ? dmask(c, DS_OPS) : kAllDirty));
}

MaskSample mask_sample(gzb_ctx* c, bool for_blocks) {
  MaskSample ms;
  for (int k = 0; k < 3; ++k) {
    const BlurPlan& pl = (k == 2 && for_blocks) ? c->p_mkb2 : c->p_mk[k];
    ms.m[k] = (k == 2 && for_blocks) ? c->d_msb2 : c->d_ms[k];
    ms.pitch[k] = pl.g.tmp_pitch;
    ms.x0[k] = pl.g.x0; ms.sx[k] = pl.g.sx; ms.y0[k] = pl.g.y0; ms.sy[k] = pl.g.sy;
  }
  return ms;
}

// Mask front + blurs for (a, b) image pair (planes with stride ps); for_blocks selects the
// channel-2 lattice.
void run_mask(gzb_ctx* c, const float* a, const float* b, bool for_blocks) {
  // the block-comparison mask (of the original against itself) is always computed in full
  const bool masked = !for_blocks;
  dim3 blk(32, 8), grd((c->W + 31) / 32, (c->H + 31) / 32, 3);
  KLAUNCH(c, KC_MASK_FRONT, k_mask_front<<<grd, blk, 0, c->stream>>>(a, b, c->ps, c->W, c->H, c->P, c->d_bl,
                                                                   masked ? dmask(c, DS_EB) : kAllDirty));
  for (int k = 0; k < 3; ++k) {
    const BlurPlan& pl = (k == 2 && for_blocks) ? c->p_mkb2 : c->p_mk[k];
    float* out = (k == 2 && for_blocks) ? c->d_msb2 : c->d_ms[k];
    run_blur(c, pl, c->d_bl + k * c->ps, 0, 1, out, 0, pl.g.tmp_pitch, nullptr, nullptr,
             masked ? dmask(c, DS_MKH0 + 2 * k) : kAllDirty, masked ? dmask(c, DS_MKV0 + 2 * k) : kAllDirty);
  }
}

// DiffmapOpsinDynamicsImage (butteraugli.cc:1046-1079) on d_xyb0 / d_xyb1 -> d_diffmap, distance.
void run_diffmap(gzb_ctx* c, const float* xyb0, const float* xyb1) {
  const int W = c->W, H = c->H, P = c->P;
  dim3 blk(32, 8), gpx((W + 31) / 32, (H + 7) / 8);
  float* m0 = c->d_mh;
  float* m1 = c->d_mh + 3 * c->ps;
  KLAUNCH(c, KC_MHIC, k_mask_high_intensity_change<<<gpx, blk, 0, c->stream>>>(xyb0, xyb1, c->ps, W, H, P, m0, m1, dmask(c, DS_MHIC)));
  if (c->concurrent) CK(cudaEventRecord(c->ev_fork, c->stream));
  // EdgeDetectorMap: the six small-sigma blurs (3 channels x 2 images) in one fused H+V launch
  {
    SmallBlur3 sb;
    bool fused = true;
    for (int k = 0; k < 3; ++k) {
      const BlurGeom& g = c->p_ed[k].g;
      sb.r[k] = g.r; sb.kind[k] = g.kind; sb.sx[k] = c->p_ed[k].d_sx; sb.sy[k] = c->p_ed[k].d_sy;
      fused = fused && g.r <= kSbR && g.sx == 1 && g.sy == 1 && g.x0 == 0 && g.y0 == 0 && g.ups == 1;
    }
    if (fused) {
      dim3 gsb((W + kSbT - 1) / kSbT, (H + kSbT - 1) / kSbT, 6);
      KLAUNCH(c, KC_BLUR_H, k_blur_small_hv<<<gsb, blk, 0, c->stream>>>(c->d_mh, c->d_bl, c->ps, W, H, P, sb, dmask(c, DS_EB)));
    } else {
      for (int k = 0; k < 3; ++k)   // (never taken with the reference's sigmas; unmasked, so Compares are then always full)
        run_blur(c, c->p_ed[k], c->d_mh + k * c->ps, 3 * c->ps, 2, c->d_bl + k * c->ps, 3 * c->ps, P);
    }
  }
  dim3 gres((c->rxs + 31) / 32, (c->rys + 7) / 8);
  KLAUNCH(c, KC_EDGE_MAP, k_edge_detector_map<<<gres, blk, 0, c->stream>>>(c->d_bl, c->d_bl + 3 * c->ps, c->ps, W, H, P, c->rxs, c->d_edm, dmask(c, DS_EDM)));
  // BlockDiffMap and EdgeDetectorLowFreq do not depend on the EdgeDetectorMap / Mask chain: they run
  // on two side streams (forked after MaskHighIntensityChange, joined before CombineChannels) so
  // that at small image sizes the short kernels of the three branches overlap.
  const int ncx = (W - 4 + 2) / 3, ncy = (H - 4 + 2) / 3;
  const int cells = ncx * ncy;
  const int ctas = std::min((cells + kBdmWarps - 1) / kBdmWarps, c->sm_count * 16);
  cudaStream_t sb = c->concurrent ? c->stream_b : c->stream;
  cudaStream_t sl = c->concurrent ? c->stream_l : c->stream;
  if (c->concurrent) {
    CK(cudaStreamWaitEvent(sb, c->ev_fork, 0));
    CK(cudaStreamWaitEvent(sl, c->ev_fork, 0));
  }
  // (block_diff_ac cells outside the kernel's domain were zeroed once, at context creation)
  const BlockChanges bc{c->d_blk_changed, c->d_scalars + 4, c->bw, c->bh};
  static const int strip_env = getenv("GZB_BDM_STRIP") ? atoi(getenv("GZB_BDM_STRIP")) : 1;
  if (strip_env) {
    const int strips = ncx * ((ncy + kBsCells - 1) / kBsCells);
    // (128 threads, 7 CTAs per SM: the best of 128x7/8, 160x5/6, 192x5, 256x4 on a B200 -- profiles/r2_bdm_strip.md)
    const int sctas = std::min(strips, c->sm_count * kBsMinCtas * 4);
    KLAUNCH_S(c, sb, KC_BLOCK_DIFF, k_block_diff_strip<<<sctas, kBsThreads, sizeof(BsSmem), sb>>>(m0, m1, c->ps, W, H, P, c->rxs, ncx, ncy, c->d_ac, dmask(c, DS_BDM), bc));
  } else {
    KLAUNCH_S(c, sb, KC_BLOCK_DIFF, k_block_diff_map<<<ctas, 32 * kBdmWarps, 0, sb>>>(m0, m1, c->ps, W, H, P, c->rxs, ncx, ncy, c->d_dc, c->d_ac, dmask(c, DS_BDM), bc));
  }
  // (the row-tiled DC kernel is 0.03 ms faster in a FULL 12 MPix Compare, but after flips -- 27 of the 46 Compares of
  // an encode, 5-18 % of the cells -- nearly every one of its CTAs still has a cell to do and loads all its rows,
  // where the lane-per-cell kernel only touches the needed cells: 154 against 112 us per Compare averaged over an
  // encode. The choice is baked into the Compare's CUDA graph, so the per-cell kernel stays the default.)
  static const int dc_rows_env = getenv("GZB_BDC_ROWS") ? atoi(getenv("GZB_BDC_ROWS")) : 0;
  if (dc_rows_env) {
    KLAUNCH_S(c, sb, KC_BLOCK_DIFF, k_block_dc_rows<<<dim3((ncx + kBdcCells - 1) / kBdcCells, ncy), kBdcThreads, 0, sb>>>(m0, m1, c->ps, W, H, P, c->rxs, ncx, ncy, c->d_dc, dmask(c, DS_BDM), bc));
  } else {
    KLAUNCH_S(c, sb, KC_BLOCK_DIFF, k_block_dc<<<(cells + 127) / 128, 128, 0, sb>>>(m0, m1, c->ps, W, H, P, c->rxs, ncx, ncy, c->d_dc, dmask(c, DS_BDM), bc));
  }
  if (c->concurrent) CK(cudaEventRecord(c->ev_bdm, sb));
  // EdgeDetectorLowFreq (its blur scratch is the first part of d_tmp, the main stream's the rest)
  run_blur(c, c->p_lf, c->d_mh, c->ps, 6, c->d_lf, c->lf_stride, c->p_lf.g.tmp_pitch, sl, c->d_tmp, dmask(c, DS_LFH), dmask(c, DS_LFV));
  if (c->concurrent) CK(cudaStreamWaitEvent(sl, c->ev_bdm, 0));   // joins the BlockDiffMap branch
  KLAUNCH_S(c, sl, KC_LOWFREQ, k_edge_lowfreq<<<gres, blk, 0, sl>>>(c->d_lf, c->d_lf + 3 * c->lf_stride, c->lf_stride, c->p_lf.g.tmp_pitch,
                                             c->p_lf.g.sx, W, H, c->rxs, c->d_lft, dmask(c, DS_LOW)));
  if (c->concurrent) CK(cudaEventRecord(c->ev_lf, sl));
  // Mask + combine
  run_mask(c, m0, m1, false);
  if (c->concurrent) CK(cudaStreamWaitEvent(c->stream, c->ev_lf, 0));   // ev_lf follows ev_bdm
  // the BlockDiffMap kernels have read the record of changed blocks: start a new one
  CK(cudaMemsetAsync(c->d_blk_changed, 0, c->nblocks, c->stream));
  KLAUNCH(c, KC_COMBINE, k_combine<<<gres, blk, 0, c->stream>>>(mask_sample(c, false), c->d_dc, c->d_ac, c->d_lft, c->d_edm, W, H, c->rxs, c->rys, c->sqp, c->d_sq,
                                                               dmask(c, DS_COMB)));
  // CalculateDiffmap
  run_blur(c, c->p_dm, c->d_sq, 0, 1, c->d_dsmall, 0, c->p_dm.g.tmp_pitch, nullptr, nullptr, dmask(c, DS_DMH), dmask(c, DS_DMV));
  CK(cudaMemsetAsync(c->d_scalars, 0, sizeof(unsigned int), c->stream));
  KLAUNCH(c, KC_DIFFMAP_FINAL, k_diffmap_final<<<gpx, blk, 0, c->stream>>>(c->d_sq, c->sqp, c->d_dsmall, c->p_dm.g.tmp_pitch, c->p_dm.g.sx, W, H, P,
                                             c->d_diffmap, c->d_ctamax, dmask(c, DS_FIN)));
  KLAUNCH(c, KC_DIFFMAP_FINAL, k_max_u32<<<std::min(64, (c->n_ctamax + 255) / 256), 256, 0, c->stream>>>(c->d_ctamax, c->n_ctamax, c->d_scalars));
}

void free_ctx(gzb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  if (c->upd_own && c->d_upd) cudaFree(c->d_upd);
  slab_release(c->device, c->slab);
  for (auto& pnd : c->prof.pend) { cudaEventDestroy(pnd.a); cudaEventDestroy(pnd.b); }
  for (cudaEvent_t ev : c->prof.free_ev) cudaEventDestroy(ev);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->cmp_graph_exec) cudaGraphExecDestroy(c->cmp_graph_exec);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->stream2) cudaStreamDestroy(c->stream2);
  if (c->stream_b) cudaStreamDestroy(c->stream_b);
  if (c->stream_l) cudaStreamDestroy(c->stream_l);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_bdm) cudaEventDestroy(c->ev_bdm);
  if (c->ev_lf) cudaEventDestroy(c->ev_lf);
  if (c->ev_cand) cudaEventDestroy(c->ev_cand);
  delete c;
}

template <typename T>
void need(gzb_ctx* c, T** pp, size_t n) {
  c->reqs.push_back({reinterpret_cast<void**>(pp), (std::max<size_t>(n, 1) * sizeof(T) + 255) / 256 * 256});
}
void commit_slab(gzb_ctx* c) {
  size_t total = 0;
  for (auto& r : c->reqs) total += r.bytes;
  c->slab = slab_acquire(c->device, total);
  char* p = static_cast<char*>(c->slab.base);
  for (auto& r : c->reqs) { *r.pp = p; p += r.bytes; }
  c->reqs.clear();
}

// Allocates everything for a W x H image (no original yet).
gzb_ctx* alloc_ctx(int device, int W, int H, float target) {
  gzb_ctx* c = new gzb_ctx;
  try {
    c->device = device;
    CK(cudaSetDevice(device));
    init_device_tables(device);
    CK(cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream2, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream_b, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&c->stream_l, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_bdm, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_lf, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&c->ev_cand, cudaEventDisableTiming));
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    c->W = W; c->H = H; c->target = target;
    c->P = round_up(W, 32);
    c->bw = (W + 7) / 8; c->bh = (H + 7) / 8; c->nblocks = c->bw * c->bh;
    c->HP = round_up(H, 16);   // whole 4:2:0 MCUs
    c->ps = static_cast<size_t>(c->P) * c->HP;
    c->mcw = (W + 15) / 16; c->mch = (H + 15) / 16; c->cbw0 = 2 * c->mcw; c->cbh0 = 2 * c->mch;
    c->cs = static_cast<size_t>(c->cbw0) * c->cbh0 * 64;
    c->rxs = (W + 2) / 3; c->rys = (H + 2) / 3; c->sqp = round_up(c->rxs, 32);
    const size_t us = c->ps;
    need(c, &c->d_rgb0, 3 * us); need(c, &c->d_rgb1, 3 * us); need(c, &c->d_stage_u8, static_cast<size_t>(3) * W * H);
    const size_t cs = c->cs;
    need(c, &c->d_orig, 3 * cs); need(c, &c->d_coef, 3 * cs);
    need(c, &c->d_ycc, 3 * us); need(c, &c->d_cup, 2 * us);
    need(c, &c->d_xyb0, 3 * c->ps); need(c, &c->d_xyb1, 3 * c->ps);
    need(c, &c->d_mh, 6 * c->ps); need(c, &c->d_bl, 6 * c->ps); need(c, &c->d_tmp, 6 * c->ps);
    // plans
    c->p_ops.build_decimated(g_hk[kB11], kB11, W, H, c->P, 0.0, 1);
    const int edk[3] = {kB15, kB0586, kB04};
    for (int k = 0; k < 3; ++k) c->p_ed[k].build_decimated(g_hk[edk[k]], edk[k], W, H, c->P, 0.0, 1);
    c->p_lf.build_decimated(g_hk[kB14], kB14, W, H, c->P, 0.0, 1);
    c->p_mk[0].build_decimated(g_hk[kB9657], kB9657, W, H, c->P, 0.0, 1);
    c->p_mk[1].build_decimated(g_hk[kB14264], kB14264, W, H, c->P, 0.0, 1);
    {  // channel 2 (step 1): only the pixels CombineChannels samples, (3rx+3, 3ry+3)
      const int nx = std::max(0, (W - 5 + 2) / 3), ny = std::max(0, (H - 5 + 2) / 3);
      c->p_mk[2].build(g_hk[kB4533], kB4533, W, H, c->P, 3, 3, nx, 3, 3, ny, 0.0, 1);
      c->p_mkb2.build(g_hk[kB4533], kB4533, W, H, c->P, 0, 8, c->bw, 0, 8, c->bh, 0.0, 1);
    }
    c->p_dm.build_decimated(g_hk[kB8851], kB8851, W - 5, H - 5, c->sqp, 0.03027655136, 3);
    c->lf_stride = c->p_lf.out_floats();
    need(c, &c->d_lf, 6 * c->lf_stride);
    for (int k = 0; k < 3; ++k) need(c, &c->d_ms[k], c->p_mk[k].out_floats());
    need(c, &c->d_msb2, c->p_mkb2.out_floats());
    const size_t rn = static_cast<size_t>(c->rxs) * c->rys;
    need(c, &c->d_edm, 3 * rn); need(c, &c->d_dc, 3 * rn); need(c, &c->d_ac, 3 * rn); need(c, &c->d_lft, 3 * rn);
    c->mtw = (W + 31) / 32; c->mth = (H + 31) / 32;
    need(c, &c->d_dirty, static_cast<size_t>(DS_COUNT) * c->mtw * c->mth);
    c->n_ctamax = ((W + 31) / 32) * ((H + 7) / 8);
    need(c, &c->d_ctamax, c->n_ctamax);
    need(c, &c->d_sq, static_cast<size_t>(c->sqp) * c->rys);
    need(c, &c->d_dsmall, c->p_dm.out_floats());
    need(c, &c->d_diffmap, c->ps);
    need(c, &c->d_bmax, std::max(c->nblocks, 256)); need(c, &c->d_weight, c->nblocks);  // d_bmax doubles as 193 LPT bins
    need(c, &c->d_mask_scale, static_cast<size_t>(3) * c->nblocks);
    need(c, &c->d_block_err, c->nblocks);
    need(c, &c->d_pregamma, static_cast<size_t>(192) * c->nblocks);
    need(c, &c->d_flags, c->nblocks);
    need(c, &c->d_scalars, 8);
    need(c, &c->d_blk_changed, c->nblocks);
    need(c, &c->d_q, 192);
    need(c, &c->d_order, static_cast<size_t>(192) * c->nblocks);
    need(c, &c->d_upd, static_cast<size_t>(c->nblocks) * 8 * 8);
    c->upd_cap = static_cast<size_t>(c->nblocks) * 8;
    BlurPlan* plans[] = {&c->p_ops, &c->p_ed[0], &c->p_ed[1], &c->p_ed[2], &c->p_lf, &c->p_mk[0], &c->p_mk[1],
                         &c->p_mk[2], &c->p_mkb2, &c->p_dm};
    for (BlurPlan* pl : plans) { need(c, &pl->d_sx, pl->hx.size()); need(c, &pl->d_sy, pl->hy.size()); }
    commit_slab(c);
    CK(cudaMemsetAsync(c->d_scalars, 0, 8 * sizeof(unsigned int), c->stream));   // (the block-change record starts invalid)
    // (function attributes are per device; the calls are cheap)
    CK(cudaFuncSetAttribute(k_block_diff_strip, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(BsSmem))));
    CK(cudaFuncSetAttribute(k_block_diff_strip, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    for (BlurPlan* pl : plans) pl->upload(c->stream);
    // TMA descriptors for the H passes whose input planes are fixed: the sigma-14 blur of the six
    // MaskHighIntensityChange planes and the three mask blurs (+ the block-comparison lattice)
    make_blur_tmap(&c->p_lf, c->d_mh, c->ps, 6);
    for (int k = 0; k < 3; ++k) make_blur_tmap(&c->p_mk[k], c->d_bl + k * c->ps, c->ps, 1);
    make_blur_tmap(&c->p_mkb2, c->d_bl + 2 * c->ps, c->ps, 1);
    c->h_pinned = static_cast<float*>(c->slab.pinned);
    CK(cudaMemsetAsync(c->d_rgb0, 0, 3 * us, c->stream));
    CK(cudaMemsetAsync(c->d_rgb1, 0, 3 * us, c->stream));
    CK(cudaMemsetAsync(c->d_edm, 0, 3 * rn * sizeof(float), c->stream));
    CK(cudaMemsetAsync(c->d_dc, 0, 3 * rn * sizeof(float), c->stream));
    CK(cudaMemsetAsync(c->d_ac, 0, 3 * rn * sizeof(float), c->stream));
    CK(cudaMemsetAsync(c->d_lft, 0, 3 * rn * sizeof(float), c->stream));
    CK(cudaMemsetAsync(c->d_ctamax, 0, c->n_ctamax * sizeof(unsigned int), c->stream));
    CK(cudaMemsetAsync(c->d_dirty, 1, static_cast<size_t>(DS_COUNT) * c->mtw * c->mth, c->stream));
    c->dirty_is_all = true;

    // the blur scratch must hold the widest H-pass output of any plan
    const BlurPlan* all[] = {&c->p_ed[0], &c->p_ed[1], &c->p_ed[2], &c->p_lf, &c->p_mk[0], &c->p_mk[1],
                             &c->p_mk[2], &c->p_mkb2, &c->p_dm};
    size_t main_need = 0;
    for (const BlurPlan* pl : all) {
      if (pl->tmp_floats() > c->ps) throw std::string("internal: blur scratch too small");
      const bool two_planes = pl == &c->p_ed[0] || pl == &c->p_ed[1] || pl == &c->p_ed[2];  // unfused fallback
      if (pl != &c->p_lf) main_need = std::max(main_need, pl->tmp_floats() * (two_planes ? 2 : 1));
    }
    // EdgeDetectorLowFreq's scratch (6 planes of its decimated width) in front, the main stream's behind
    const size_t lf_need = (6 * c->p_lf.tmp_floats() + 63) & ~size_t(63);
    c->concurrent = lf_need + main_need <= 6 * c->ps;
    c->tmp_main_off = c->concurrent ? lf_need : 0;
    // TMA descriptors for the V passes: their input is the H-pass scratch, whose place is now known
    make_blur_tmap_v(&c->p_lf, c->d_tmp, 6);
    for (int k = 0; k < 3; ++k) make_blur_tmap_v(&c->p_mk[k], c->d_tmp + c->tmp_main_off, 1);
    make_blur_tmap_v(&c->p_mkb2, c->d_tmp + c->tmp_main_off, 1);
    make_blur_tmap_v(&c->p_dm, c->d_tmp + c->tmp_main_off, 1);
    CK(cudaStreamSynchronize(c->stream));
    return c;
  } catch (const std::string& e) {
    g_create_err = e;
    free_ctx(c);
    return nullptr;
  }
}

void upload_original(gzb_ctx* c, const uint8_t* rgb) {
  const size_t n = static_cast<size_t>(3) * c->W * c->H;
  CK(cudaMemcpyAsync(c->d_stage_u8, rgb, n, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (n);
  dim3 grd((c->W + 255) / 256, c->H);
  KLAUNCH(c, KC_MISC, k_deinterleave_rgb<<<grd, 256, 0, c->stream>>>(c->d_stage_u8, c->W, c->H, c->P, c->d_rgb0, c->ps));
  opsin_from_u8(c, c->d_rgb0, c->d_xyb0);
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
}

// ---- incremental Compare: per-stage dirty-tile masks from the changed pixel rectangles --------------
// A mask marks the 32x32-pixel tiles in which an output of the stage may change. Radii follow the
// supports of the stages (butteraugli.cc): opsin blur 2, MaskHighIntensityChange 1, EdgeDetectorMap
// blurs <= 3 and its +-3 differences on the (clamped) 8x8 block of a res cell, the 8x8 windows of
// BlockDiffMap, the sigma-14 lattice (radius 31, step 4) and the +8 / -6 offsets of EdgeDetectorLowFreq,
// the mask front (-2 .. +5) and its three blurs, CombineChannels' sample at (+3, +3), the diffmap blur
// (radius 19 on the 3x-replicated res map, step 2) and the (x - 2) / 3 indexing of the final map. All are
// rounded up generously; a superfluous dirty tile only costs time.
// Buffers that do NOT survive a Compare (the H-pass scratch of every blur; planes 0..2 of d_bl, which
// hold the EdgeDetectorMap blurs and then the mask front) are refreshed over the whole footprint of their
// dirty consumers: those masks are dilations of the consumers' tile sets.
struct MaskBuilder {
  int tw, th, W, H;
  uint8_t* base;
  uint8_t* m(int stage) const { return base + static_cast<size_t>(stage) * tw * th; }
  void mark(int stage, int x0, int y0, int x1, int y1) const {
    if (x1 < 0 || y1 < 0 || x0 >= W || y0 >= H) return;
    const int tx0 = std::max(x0, 0) >> 5, ty0 = std::max(y0, 0) >> 5;
    const int tx1 = std::min(x1, W - 1) >> 5, ty1 = std::min(y1, H - 1) >> 5;
    uint8_t* p = m(stage);
    for (int ty = ty0; ty <= ty1; ++ty) memset(p + ty * tw + tx0, 1, tx1 - tx0 + 1);
  }
  // dst |= tiles within (fx, fy) pixels of a dirty tile of src
  void dilate(int src, int dst, int fx, int fy) const {
    const uint8_t* sp = m(src);
    for (int ty = 0; ty < th; ++ty)
      for (int tx = 0; tx < tw; ++tx)
        if (sp[ty * tw + tx]) mark(dst, 32 * tx - fx, 32 * ty - fy, 32 * tx + 31 + fx, 32 * ty + 31 + fy);
  }
  size_t count(int stage) const {
    size_t n = 0;
    const uint8_t* p = m(stage);
    for (int i = 0; i < tw * th; ++i) n += p[i];
    return n;
  }
};

// Fills h_dirty for the rectangles in c->changed; returns the fraction of tiles the last stage touches.
double build_dirty_masks(gzb_ctx* c) {
  MaskBuilder b{c->mtw, c->mth, c->W, c->H, c->h_dirty};
  memset(c->h_dirty, 0, static_cast<size_t>(DS_COUNT) * c->mtw * c->mth);
  const int r_lf = c->p_lf.g.r, s_lf = c->p_lf.g.sx, r_dm = c->p_dm.g.r, s_dm = c->p_dm.g.sx;
  for (const int4& q : c->changed) {
    auto grow = [&](int stage, int r) { b.mark(stage, q.x - r, q.y - r, q.z + r, q.w + r); };
    grow(DS_OPS, 2);
    grow(DS_MHIC, 3);
    grow(DS_EDM, 20);
    grow(DS_BDM, 20);
    grow(DS_EB, 9);
    grow(DS_LFV, 3 + r_lf + s_lf + 1);
    for (int k = 0; k < 3; ++k) grow(DS_MKV0 + 2 * k, 9 + c->p_mk[k].g.r + c->p_mk[k].g.sx);
  }
  b.dilate(DS_EDM, DS_EB, 11, 11);
  b.dilate(DS_LFV, DS_LFH, 0, r_lf);
  b.dilate(DS_LFV, DS_LOW, 16, 16);
  for (int k = 0; k < 3; ++k) b.dilate(DS_MKV0 + 2 * k, DS_MKH0 + 2 * k, 0, c->p_mk[k].g.r);
  b.dilate(DS_EDM, DS_COMB, 0, 0);
  b.dilate(DS_BDM, DS_COMB, 0, 0);
  b.dilate(DS_LOW, DS_COMB, 8, 1);
  for (int k = 0; k < 3; ++k) b.dilate(DS_MKV0 + 2 * k, DS_COMB, c->p_mk[k].g.sx + 4, c->p_mk[k].g.sx + 4);
  b.dilate(DS_COMB, DS_DMV, 3 + r_dm + s_dm + 4, 3 + r_dm + s_dm + 4);
  b.dilate(DS_DMV, DS_DMH, 0, r_dm + 2);
  b.dilate(DS_DMV, DS_FIN, 4 + s_dm, 4 + s_dm);
  b.dilate(DS_COMB, DS_FIN, 6, 6);
  return static_cast<double>(b.count(DS_FIN)) / (static_cast<double>(c->mtw) * c->mth);
}

// Everything a full-image operation on the candidate (or anything that overwrites buffers a Compare
// leaves behind) must do: the next Compare recomputes every tile.
void invalidate_compare_state(gzb_ctx* c) {
  c->inter_valid = false;
  c->changed.clear();
  c->changed_overflow = false;
}

// Chooses between a full and an incremental Compare and makes d_dirty say so (on the main stream).
void prepare_dirty_masks(gzb_ctx* c) {
  static const bool disabled = getenv("GZB_NO_INCREMENTAL") != nullptr;
  const size_t total = static_cast<size_t>(DS_COUNT) * c->mtw * c->mth;
  bool fused = true;   // the unfused EdgeDetectorMap blur fallback of run_diffmap is not masked
  for (int k = 0; k < 3; ++k) {
    const BlurGeom& g = c->p_ed[k].g;
    fused = fused && g.r <= kSbR && g.sx == 1 && g.sy == 1 && g.x0 == 0 && g.y0 == 0 && g.ups == 1;
  }
  bool incremental = !disabled && fused && c->inter_valid && !c->changed_overflow && !c->changed.empty();
  if (incremental) {
    if (c->slab.pinned2_cap < total) {
      if (c->slab.pinned2) cudaFreeHost(c->slab.pinned2);
      c->slab.pinned2 = nullptr;
      c->slab.pinned2_cap = 0;
      CK(cudaMallocHost(&c->slab.pinned2, total));
      c->slab.pinned2_cap = total;
    }
    c->h_dirty = static_cast<uint8_t*>(c->slab.pinned2);
    incremental = build_dirty_masks(c) < 0.6;
  }
  if (incremental) {
    CK(cudaMemcpyAsync(c->d_dirty, c->h_dirty, total, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += total;
    c->dirty_is_all = false;
    ++c->incremental_compares;
  } else if (!c->dirty_is_all) {
    CK(cudaMemsetAsync(c->d_dirty, 1, total, c->stream));
    c->dirty_is_all = true;
  }
}

int fail(gzb_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg; else g_create_err = msg;
  return code;
}

#define GZB_TRY(ctx) \
  if (!(ctx)) return GZB_ERR_BAD_ARG; \
  try { CK(cudaSetDevice((ctx)->device));
#define GZB_END(ctx) \
  } catch (const std::string& e) { return fail((ctx), GZB_ERR_CUDA, e); } \
  return GZB_OK;

void sync_check(gzb_ctx* c) {
  CK(cudaStreamSynchronize(c->stream));
  CK(cudaGetLastError());
  if (!c->prof.pend.empty()) prof_resolve(c);
}
// The side stream of the entropy coder. Pending profile events are resolved by the next sync_check
// (main-stream events may still be in flight here).
void sync_check2(gzb_ctx* c) {
  CK(cudaStreamSynchronize(c->stream2));
  CK(cudaGetLastError());
}

}  // namespace

// =============================================================================================
// C ABI
// =============================================================================================
extern "C" {

const char* gzb_version(void) { return "gzb200 0.1 (sm_100a)"; }
const char* gzb_last_error(const gzb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }
int gzb_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int gzb_create(int device, int width, int height, const uint8_t* rgb, float target_distance, gzb_ctx** out) {
  if (!out || !rgb) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_create: null argument");
  *out = nullptr;
  if (width < 32 || height < 32) return fail(nullptr, GZB_ERR_TOO_SMALL, "gzb_create: image smaller than 32x32");
  if (width >= (1 << 16) || height >= (1 << 16)) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_create: image too large");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0 || device < 0 || device >= n)
    return fail(nullptr, GZB_ERR_CUDA, std::string("gzb_create: no usable CUDA device (") +
                                           (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") + ")");
  gzb_ctx* c = alloc_ctx(device, width, height, target_distance);
  if (!c) return GZB_ERR_CUDA;
  try {
    upload_original(c, rgb);
  } catch (const std::string& msg) {
    g_create_err = msg;
    free_ctx(c);
    return GZB_ERR_CUDA;
  }
  *out = c;
  return GZB_OK;
}

void gzb_destroy(gzb_ctx* ctx) { free_ctx(ctx); }

int gzb_set_jpeg_coeffs(gzb_ctx* c, const int16_t* c0, const int16_t* c1, const int16_t* c2) {
  GZB_TRY(c)
  const size_t cs = c->cs;
  const int16_t* src[3] = {c0, c1, c2};
  c->mode420 = false;   // the input of the RGB front end is 4:4:4
  c->have_coeffs = false;
  invalidate_compare_state(c);
  const size_t nb2 = static_cast<size_t>(c->nblocks) * 64 * 2;
  for (int k = 0; k < 3; ++k) { CK(cudaMemcpyAsync(c->d_orig + k * cs, src[k], nb2, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += nb2; }
  sync_check(c);
  c->have_orig_coeffs = true;
  c->packed_valid = false;
  GZB_END(c)
}

int gzb_rgb_to_jpeg_coeffs_device(gzb_ctx* c) {
  GZB_TRY(c)
  c->mode420 = false;
  c->have_coeffs = false;
  invalidate_compare_state(c);
  KLAUNCH(c, KC_MISC, k_rgb_to_coeffs<<<(c->nblocks + 31) / 32, 256, 0, c->stream>>>(c->d_rgb0, c->ps, c->W, c->H, c->P, c->bw, c->nblocks,
                                                                                c->d_orig, c->cs));
  sync_check(c);
  c->have_orig_coeffs = true;
  c->packed_valid = false;
  GZB_END(c)
}

int gzb_copy_from_jpeg(gzb_ctx* c, const int* quant192) {
  GZB_TRY(c)
  if (!c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_copy_from_jpeg: gzb_set_jpeg_coeffs not called");
  CK(cudaMemcpyAsync(c->d_q, quant192, 192 * sizeof(int), cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (192 * sizeof(int));
  invalidate_compare_state(c);
  render_candidate(c, kCoeffScale);
  sync_check(c);
  c->have_coeffs = true;
  GZB_END(c)
}

int gzb_quantize_from_jpeg(gzb_ctx* c, const int* q192) {
  GZB_TRY(c)
  if (!c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_quantize_from_jpeg: gzb_set_jpeg_coeffs not called");
  for (int i = 0; i < 192; ++i) if (q192[i] <= 0) return fail(c, GZB_ERR_BAD_ARG, "quantiser must be positive");
  CK(cudaMemcpyAsync(c->d_q, q192, 192 * sizeof(int), cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (192 * sizeof(int));
  invalidate_compare_state(c);
  render_candidate(c, kCoeffQuantizeSrc);   // no host sync: later calls are ordered by the stream / ev_cand
  c->have_coeffs = true;
  GZB_END(c)
}

int gzb_apply_global_quantization(gzb_ctx* c, const int* q192) {
  GZB_TRY(c)
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_apply_global_quantization: no candidate coefficients");
  for (int i = 0; i < 192; ++i) if (q192[i] <= 0) return fail(c, GZB_ERR_BAD_ARG, "quantiser must be positive");
  CK(cudaMemcpyAsync(c->d_q, q192, 192 * sizeof(int), cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (192 * sizeof(int));
  invalidate_compare_state(c);
  render_candidate(c, kCoeffQuantize);
  sync_check(c);
  GZB_END(c)
}

int gzb_set_coeffs(gzb_ctx* c, const int16_t* c0, const int16_t* c1, const int16_t* c2) {
  GZB_TRY(c)
  const size_t cs = c->cs;
  const int16_t* src[3] = {c0, c1, c2};
  for (int k = 0; k < 3; ++k) {
    const size_t n2 = comp_blocks(c, k) * 64 * 2;
    CK(cudaMemcpyAsync(c->d_coef + k * cs, src[k], n2, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n2;
  }
  invalidate_compare_state(c);
  render_candidate(c, kCoeffKeep);
  sync_check(c);
  c->have_coeffs = true;
  GZB_END(c)
}

int gzb_get_coeffs(gzb_ctx* c, int16_t* c0, int16_t* c1, int16_t* c2) {
  GZB_TRY(c)
  const size_t cs = c->cs;
  int16_t* dst[3] = {c0, c1, c2};
  for (int k = 0; k < 3; ++k) {
    const size_t n2 = comp_blocks(c, k) * 64 * 2;
    CK(cudaMemcpyAsync(dst[k], c->d_coef + k * cs, n2, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += n2;
  }
  sync_check(c);
  GZB_END(c)
}

static double dbg_now_ms() {
  timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
int gzb_update_coeffs(gzb_ctx* c, const int32_t* block_ix, const uint8_t* idx, const int16_t* val, size_t n) {
  GZB_TRY(c)
  static const bool dbg = getenv("GZB_DEBUG") != nullptr;
  const double d0 = dbg_now_ms();
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_update_coeffs: no candidate coefficients");
  {
    const int lim[3] = {static_cast<int>(comp_blocks(c, 0)), static_cast<int>(comp_blocks(c, 1)), static_cast<int>(comp_blocks(c, 2))};
    for (size_t i = 0; i < n; ++i)
      if (idx[i] >= 192 || block_ix[i] < 0 || block_ix[i] >= lim[idx[i] >> 6]) return fail(c, GZB_ERR_BAD_ARG, "gzb_update_coeffs: index out of range");
  }
  if (c->inter_valid && !c->changed_overflow) {
    // pixel rectangles whose samples change: an 8x8 block, or the 16x16 area (+1: fancy upsampling) of a
    // sub-sampled chroma block. Large batches are not tracked -- the next Compare is then a full one.
    if (n + c->changed.size() > 256) {
      c->changed_overflow = true;
      c->changed.clear();
    } else {
      for (size_t i = 0; i < n; ++i) {
        const int comp = idx[i] >> 6, bwc = comp_bw(c, comp);
        const int bx = block_ix[i] % bwc, by = block_ix[i] / bwc;
        int4 q;
        if (c->mode420 && comp > 0) q = make_int4(16 * bx - 1, 16 * by - 1, 16 * bx + 16, 16 * by + 16);
        else q = make_int4(8 * bx, 8 * by, 8 * bx + 7, 8 * by + 7);
        bool dup = false;
        for (const int4& o : c->changed) dup = dup || (o.x == q.x && o.y == q.y && o.z == q.z && o.w == q.w);
        if (!dup) c->changed.push_back(q);
      }
    }
  }
  if (n > 0) {
    // staging: the blur scratch (6 planes, >= 24 bytes per pixel) holds 8 bytes per record for up to
    // three records per pixel -- more than there are AC coefficients; larger batches get their own
    c->packed_valid = false;
    uint8_t* stage = reinterpret_cast<uint8_t*>(c->d_tmp);
    size_t cap = 6 * c->ps * sizeof(float) / 8;
    if (n > cap) {
      if (n > c->upd_cap) {
        if (c->upd_own && c->d_upd) cudaFree(c->d_upd);
        c->upd_cap = n + n / 2 + 1024;
        dmalloc(&c->d_upd, c->upd_cap * 8);
        c->upd_own = true;
      }
      stage = c->d_upd;
      cap = c->upd_cap;
    }
    const double d1 = dbg_now_ms();
    // records are applied in order (later writes to the same coefficient win): one thread walks
    // duplicates, so pack {block, idx, val} and let the kernel resolve by record index
    CK(cudaMemcpyAsync(stage, block_ix, n * 4, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (n * 4);
    CK(cudaMemcpyAsync(stage + cap * 4, val, n * 2, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (n * 2);
    CK(cudaMemcpyAsync(stage + cap * 6, idx, n, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (n);
    const size_t cs = c->cs;
    KLAUNCH(c, KC_MISC, k_scatter_coeffs<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(
        reinterpret_cast<const int*>(stage), reinterpret_cast<const int16_t*>(stage + cap * 4),
        stage + cap * 6, n, cs, c->d_coef, c->mode420 ? nullptr : c->d_blk_changed));
    if (c->mode420) c->blk_tracking = false;   // (component block != image block: not recorded)
    if (dbg) { sync_check(c); fprintf(stderr, "update n=%zu validate %.3f ms, h2d+scatter %.3f ms\n", n, d1 - d0, dbg_now_ms() - d1); }
  }
  const double d2 = dbg_now_ms();
  render_candidate(c, kCoeffKeep);   // no host sync: later calls are ordered by the stream / ev_cand
  if (dbg) fprintf(stderr, "update render %.3f ms\n", dbg_now_ms() - d2);
  GZB_END(c)
}

int gzb_clear_distmap(gzb_ctx* c) {
  GZB_TRY(c)
  CK(cudaMemsetAsync(c->d_diffmap, 0, c->ps * sizeof(float), c->stream));
  invalidate_compare_state(c);
  c->have_distmap = true;
  GZB_END(c)
}

int gzb_to_srgb(gzb_ctx* c, uint8_t* rgb_out) {
  GZB_TRY(c)
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_to_srgb: no candidate coefficients");
  std::vector<uint8_t> pl(3 * static_cast<size_t>(c->W) * c->H);
  const size_t us = c->ps;
  for (int k = 0; k < 3; ++k)
    CK(cudaMemcpy2DAsync(pl.data() + static_cast<size_t>(k) * c->W * c->H, c->W, c->d_rgb1 + k * us, c->P, c->W, c->H,
                         cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<unsigned long long>(c->W) * (c->H));
  sync_check(c);
  const size_t n = static_cast<size_t>(c->W) * c->H;
  for (size_t i = 0; i < n; ++i) {
    rgb_out[3 * i] = pl[i];
    rgb_out[3 * i + 1] = pl[n + i];
    rgb_out[3 * i + 2] = pl[2 * n + i];
  }
  GZB_END(c)
}

int gzb_compare_begin(gzb_ctx* c) {
  GZB_TRY(c)
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_compare: no candidate coefficients");
  CK(cudaEventRecord(c->ev0, c->stream));
  // BlockDiffMap cells are recomputed only around the blocks flipped since the last Compare when that record is
  // complete and the buffers of the last Compare still describe the candidate before those flips.
  static const bool no_fine = getenv("GZB_NO_FINE_BDM") != nullptr;
  const bool fine_bdm = !no_fine && c->inter_valid && c->blk_tracking;
  CK(cudaMemsetAsync(c->d_scalars + 4, fine_bdm ? 1 : 0, sizeof(unsigned int), c->stream));
  if (fine_bdm) ++c->fine_bdm_compares;
  // Full or incremental: the per-stage dirty masks decide which tiles the kernels below recompute.
  prepare_dirty_masks(c);
  // Every Compare of a context runs the same ~25 launches on the same buffers (three streams, fork
  // and join): captured once into a CUDA graph and replayed, which removes the per-launch gaps that
  // dominate at small image sizes. Per-kernel profiling needs the individual launches.
  if (c->prof.on || c->graph_failed) {
    opsin_from_u8(c, c->d_rgb1, c->d_xyb1, true);
    run_diffmap(c, c->d_xyb0, c->d_xyb1);
  } else {
    if (!c->cmp_graph_exec) {
      const unsigned long long before = c->launches;
      const bool pv = c->packed_valid;
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed);
      if (e == cudaSuccess) {
        try {
          opsin_from_u8(c, c->d_rgb1, c->d_xyb1, true);
          run_diffmap(c, c->d_xyb0, c->d_xyb1);
        } catch (const std::string&) { e = cudaErrorUnknown; }
        const cudaError_t e2 = cudaStreamEndCapture(c->stream, &graph);
        if (e == cudaSuccess) e = e2;
      }
      if (e == cudaSuccess && graph) e = cudaGraphInstantiate(&c->cmp_graph_exec, graph, 0);
      if (graph) cudaGraphDestroy(graph);
      c->cmp_graph_launches = c->launches - before;
      c->launches = before;
      c->packed_valid = pv;
      if (e != cudaSuccess || !c->cmp_graph_exec) {   // no graph on this driver: plain launches from now on
        cudaGetLastError();
        c->cmp_graph_exec = nullptr;
        c->graph_failed = true;
      }
    }
    if (c->cmp_graph_exec) {
      c->packed_valid = false;
      CK(cudaGraphLaunch(c->cmp_graph_exec, c->stream));
      c->launches += c->cmp_graph_launches;
    } else {
      opsin_from_u8(c, c->d_rgb1, c->d_xyb1, true);
      run_diffmap(c, c->d_xyb0, c->d_xyb1);
    }
  }
  // from here on the context's buffers describe this candidate
  c->blk_tracking = true;   // (the Compare has cleared the record)
  c->inter_valid = true;
  c->changed.clear();
  c->changed_overflow = false;
  CK(cudaEventRecord(c->ev1, c->stream));
  CK(cudaMemcpyAsync(c->h_pinned, c->d_scalars, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (sizeof(unsigned int));
  c->compare_pending = true;
  GZB_END(c)
}

int gzb_compare_end(gzb_ctx* c, float* distance) {
  GZB_TRY(c)
  if (!c->compare_pending) return fail(c, GZB_ERR_STATE, "gzb_compare_end: no Compare in flight");
  c->compare_pending = false;
  sync_check(c);
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  memcpy(&c->distance, c->h_pinned, sizeof(float));
  c->have_distmap = true;
  if (distance) *distance = c->distance;
  GZB_END(c)
}

int gzb_compare(gzb_ctx* c, float* distance) {
  const int rc = gzb_compare_begin(c);
  return rc != GZB_OK ? rc : gzb_compare_end(c, distance);
}

int gzb_get_distmap(gzb_ctx* c, float* out) {
  GZB_TRY(c)
  if (!c->have_distmap) return fail(c, GZB_ERR_STATE, "gzb_get_distmap: no Compare yet");
  CK(cudaMemcpy2DAsync(out, c->W * sizeof(float), c->d_diffmap, c->P * sizeof(float), c->W * sizeof(float), c->H,
                       cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H));
  sync_check(c);
  GZB_END(c)
}

int gzb_distance_ok(const gzb_ctx* c, double target_mul) {
  // float distance_ <= double target_mul * float target_distance_ (butteraugli_comparator.h:52-54)
  return c && static_cast<double>(c->distance) <= target_mul * static_cast<double>(c->target) ? 1 : 0;
}

double gzb_score_output_size(const gzb_ctx* c, int size) {
  // ScoreJPEG (guetzli/score.cc:23-41)
  if (!c) return -1.0;
  const double diff = static_cast<double>(c->distance) - static_cast<double>(c->target);
  if (diff <= 0.0) return size;
  const double e = 50 * diff;
  if (e > 10) return 1e30 * std::exp(10.0) * diff + size;
  return std::exp(e) * size;
}

float gzb_block_error_limit(const gzb_ctx* c) { return c ? c->target : 0.f; }

int gzb_image_size(const gzb_ctx* c, int* width, int* height) {
  if (!c || !width || !height) return GZB_ERR_BAD_ARG;
  *width = c->W;
  *height = c->H;
  return GZB_OK;
}

int gzb_start_block_comparisons(gzb_ctx* c) {
  GZB_TRY(c)
  CK(cudaEventRecord(c->ev0, c->stream));
  invalidate_compare_state(c);   // the mask of the original overwrites the mask buffers of the last Compare
  run_mask(c, c->d_xyb0, c->d_xyb0, true);
  KLAUNCH(c, KC_BLOCK_MASK, k_block_mask_scale<<<(c->nblocks + 255) / 256, 256, 0, c->stream>>>(mask_sample(c, true), c->bw, c->bh, c->d_mask_scale));
  CK(cudaEventRecord(c->ev1, c->stream));
  sync_check(c);
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  c->block_cmp = true;
  GZB_END(c)
}

int gzb_finish_block_comparisons(gzb_ctx* c) {
  if (!c) return GZB_ERR_BAD_ARG;
  c->block_cmp = false;
  return GZB_OK;
}

// Number of units of the zeroing search for comp_mask: 8x8 blocks, or 16x16 macro-blocks when the
// last selected component is sub-sampled (processor.cc:566-572).
static int zeroing_units(const gzb_ctx* c, int comp_mask) {
  return (c->mode420 && (comp_mask & 6)) ? c->mcw * c->mch : c->nblocks;
}

// Where gzb_compute_block_zeroing_candidates_range packs its lists: d_tmp (6 planes of scratch).
static void packed_ptrs(gzb_ctx* c, int** d_counts, int** d_offsets, float** d_err, uint8_t** d_idx) {
  *d_counts = reinterpret_cast<int*>(c->d_tmp);
  *d_offsets = *d_counts + c->nblocks + 32;
  *d_err = reinterpret_cast<float*>(*d_offsets + c->nblocks + 32);
  *d_idx = reinterpret_cast<uint8_t*>(*d_err + static_cast<size_t>(192) * c->nblocks);
}

static int run_zeroing(gzb_ctx* c, int comp_mask, int mode, int b0 = 0, int b1 = -1) {
  const size_t cs = c->cs;
  const bool mb = c->mode420 && (comp_mask & 6);
  const int units = zeroing_units(c, comp_mask);
  if (b1 < 0) b1 = units;
  CK(cudaMemsetAsync(c->d_scalars + 1, 0, sizeof(unsigned int), c->stream));
  CK(cudaMemsetAsync(c->d_scalars + 3, 0, sizeof(unsigned int), c->stream));   // blocks whose sort keys tie
  if (mode == 0)
    CK(cudaMemsetAsync(c->d_order + 192 * static_cast<size_t>(b0), 0,
                       sizeof(gzb_coeff_data) * 192 * static_cast<size_t>(b1 - b0), c->stream));
  CK(cudaEventRecord(c->ev0, c->stream));
  const int pass_bw = mb ? c->mcw : c->bw;                       // units per row
  const int coef_bw = mb ? c->mcw : comp_bw(c, 0);               // coefficient blocks per row of the searched plane
  const int* lpt = nullptr;
  if (mode == 0) {
    // buffers that are idle during the zeroing search: the error array of the single-CompareBlock
    // modes (block order), the weight flags (per-block cost), the block-max array (193 bins)
    int* order = reinterpret_cast<int*>(c->d_block_err);
    unsigned int* bins = reinterpret_cast<unsigned int*>(c->d_bmax);
    const int g = (b1 - b0 + 255) / 256;
    CK(cudaMemsetAsync(bins, 0, 193 * sizeof(unsigned int), c->stream));
    KLAUNCH(c, KC_MISC, k_zero_block_cost<<<g, 256, 0, c->stream>>>(c->d_coef, cs, comp_mask, b0, b1, pass_bw, coef_bw, c->d_flags, bins));
    KLAUNCH(c, KC_MISC, k_zero_cost_scan<<<1, 32, 0, c->stream>>>(bins));
    KLAUNCH(c, KC_MISC, k_zero_lpt_scatter<<<g, 256, 0, c->stream>>>(c->d_flags, bins, b0, b1, order));
    lpt = order;
  }
  if (mb) {
    const int ctas = std::max(1, std::min(b1 - b0, c->sm_count * 5));
    KLAUNCH(c, KC_ZEROING, k_zeroing_order_mb<<<ctas, 128, 0, c->stream>>>(
        c->d_orig, c->d_coef, cs, c->d_rgb0, c->d_ycc, c->ps, c->P, c->W, c->H, c->bw, c->mcw, b1, c->d_mask_scale,
        c->target, 3, b0, reinterpret_cast<CoeffDataDev*>(c->d_order), c->d_scalars + 1, lpt, c->d_scalars + 3));
  } else {
    static const int per_sm = getenv("GZB_ZERO_CTAS_PER_SM") ? atoi(getenv("GZB_ZERO_CTAS_PER_SM")) : 6;   // tuning probe (7 fit; the pipes saturate at 5-6)
    const int ctas = std::max(1, std::min((b1 - b0 + kZeroWarps - 1) / kZeroWarps, c->sm_count * std::max(1, per_sm)));
    // Too few blocks to fill the GPU with one warp each (4 warps x 7 CTAs per SM): three warps per block, one per
    // look-ahead trial (8 teams per SM are resident). GZB_ZERO_TEAM=0/1 forces the choice.
    static const int team_env = getenv("GZB_ZERO_TEAM") ? atoi(getenv("GZB_ZERO_TEAM")) : -1;
    const bool team = mode == 0 && (team_env >= 0 ? team_env != 0 : (b1 - b0) <= 2 * c->sm_count * 8);
    if (team) {
      const int tctas = std::max(1, std::min((b1 - b0 + 1) / 2, c->sm_count * 4));
      KLAUNCH(c, KC_ZEROING, k_zeroing_order<3><<<tctas, 32 * kZeroTeamWarps, 0, c->stream>>>(
          c->d_orig, c->d_coef, cs, c->d_rgb0, c->ps, c->P, c->W, c->H, c->bw, b1, c->d_mask_scale, comp_mask,
          c->target, 3, mode, 0, b0, reinterpret_cast<CoeffDataDev*>(c->d_order), c->d_block_err, c->d_pregamma, c->d_scalars + 1, lpt,
          coef_bw, c->mode420 ? c->d_cup : nullptr, c->d_scalars + 3));
    } else {
      KLAUNCH(c, KC_ZEROING, k_zeroing_order<1><<<ctas, 32 * kZeroWarps, 0, c->stream>>>(
          c->d_orig, c->d_coef, cs, c->d_rgb0, c->ps, c->P, c->W, c->H, c->bw, b1, c->d_mask_scale, comp_mask,
          c->target, 3, mode, 0, b0, reinterpret_cast<CoeffDataDev*>(c->d_order), c->d_block_err, c->d_pregamma, c->d_scalars + 1, lpt,
          coef_bw, c->mode420 ? c->d_cup : nullptr, c->d_scalars + 3));
    }
  }
  CK(cudaEventRecord(c->ev1, c->stream));
  return 0;
}

int gzb_get_block_lists(gzb_ctx* c, float* mask_scale_out, float* opsin_blocks_out) {
  GZB_TRY(c)
  if (!c->block_cmp) return fail(c, GZB_ERR_STATE, "gzb_get_block_lists: StartBlockComparisons not called");
  if (c->mode420) return fail(c, GZB_ERR_UNSUPPORTED, "gzb_get_block_lists: 4:4:4 candidates only");
  if (opsin_blocks_out) {
    if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_get_block_lists: no candidate coefficients");
    run_zeroing(c, 7, 1);
    CK(cudaMemcpyAsync(opsin_blocks_out, c->d_pregamma, sizeof(float) * 192 * static_cast<size_t>(c->nblocks), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (sizeof(float) * 192 * static_cast<size_t>(c->nblocks));
  }
  if (mask_scale_out)
    CK(cudaMemcpyAsync(mask_scale_out, c->d_mask_scale, sizeof(float) * 3 * static_cast<size_t>(c->nblocks), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (sizeof(float) * 3 * static_cast<size_t>(c->nblocks));
  sync_check(c);
  GZB_END(c)
}

int gzb_compare_blocks(gzb_ctx* c, float* err_out) {
  GZB_TRY(c)
  if (!c->block_cmp) return fail(c, GZB_ERR_STATE, "gzb_compare_blocks: StartBlockComparisons not called");
  if (c->mode420) return fail(c, GZB_ERR_UNSUPPORTED, "gzb_compare_blocks: 4:4:4 candidates only");
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_compare_blocks: no candidate coefficients");
  run_zeroing(c, 7, 1);
  CK(cudaMemcpyAsync(err_out, c->d_block_err, sizeof(float) * c->nblocks, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (sizeof(float) * c->nblocks);
  sync_check(c);
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  GZB_END(c)
}

int gzb_compare_block(gzb_ctx* c, int block_x, int block_y, const int16_t* candidate192, double* err) {
  GZB_TRY(c)
  if (!c->block_cmp) return fail(c, GZB_ERR_STATE, "gzb_compare_block: StartBlockComparisons not called");
  if (c->mode420) return fail(c, GZB_ERR_UNSUPPORTED, "gzb_compare_block: 4:4:4 candidates only");
  if (!candidate192 || !err || block_x < 0 || block_x >= c->bw || block_y < 0 || block_y >= c->bh)
    return fail(c, GZB_ERR_BAD_ARG, "gzb_compare_block: bad argument");
  int16_t* d_cand = reinterpret_cast<int16_t*>(c->d_upd);  // 384 bytes of the update staging area
  CK(cudaMemcpyAsync(d_cand, candidate192, 192 * sizeof(int16_t), cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += 384;
  CK(cudaMemsetAsync(c->d_scalars + 1, 0, sizeof(unsigned int), c->stream));
  KLAUNCH(c, KC_ZEROING, k_zeroing_order<1><<<1, 32 * kZeroWarps, 0, c->stream>>>(
      d_cand, d_cand, 64, c->d_rgb0, c->ps, c->P, c->W, c->H, c->bw, 1, c->d_mask_scale, 7, c->target, 3, 2,
      block_y * c->bw + block_x, 0, nullptr, c->d_block_err, nullptr, c->d_scalars + 1, nullptr, c->bw, nullptr, nullptr));
  CK(cudaMemcpyAsync(c->h_pinned, c->d_block_err, sizeof(float), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 4;
  sync_check(c);
  *err = static_cast<double>(c->h_pinned[0]);
  GZB_END(c)
}

int gzb_compare_block_srgb(gzb_ctx* c, int block_x, int block_y, const uint8_t* rgb192, double* err) {
  GZB_TRY(c)
  if (!c->block_cmp) return fail(c, GZB_ERR_STATE, "gzb_compare_block_srgb: StartBlockComparisons not called");
  if (!rgb192 || !err || block_x < 0 || block_x >= c->bw || block_y < 0 || block_y >= c->bh)
    return fail(c, GZB_ERR_BAD_ARG, "gzb_compare_block_srgb: bad argument");
  uint8_t* d_win = c->d_upd;  // 192 bytes of the update staging area
  CK(cudaMemcpyAsync(d_win, rgb192, 192, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += 192;
  KLAUNCH(c, KC_ZEROING, k_compare_block_rgb<<<1, 32, 0, c->stream>>>(d_win, c->d_rgb0, c->ps, c->P, c->W, c->H, c->bw, block_x, block_y,
                                                                     c->d_mask_scale, c->d_block_err));
  CK(cudaMemcpyAsync(c->h_pinned, c->d_block_err, sizeof(float), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 4;
  sync_check(c);
  *err = static_cast<double>(c->h_pinned[0]);
  GZB_END(c)
}

int gzb_set_sampling(gzb_ctx* c, int chroma_factor) {
  if (!c) return GZB_ERR_BAD_ARG;
  if (chroma_factor != 1 && chroma_factor != 2) return fail(c, GZB_ERR_BAD_ARG, "gzb_set_sampling: factor must be 1 or 2");
  const bool m = chroma_factor == 2;
  if (m != c->mode420) { c->mode420 = m; c->have_coeffs = false; c->have_orig_coeffs = false; c->packed_valid = false; invalidate_compare_state(c); }
  return GZB_OK;
}

int gzb_set_jpeg_coeffs_420(gzb_ctx* c, const int16_t* c0, const int16_t* c1, const int16_t* c2) {
  GZB_TRY(c)
  if (!c0 || !c1 || !c2) return fail(c, GZB_ERR_BAD_ARG, "gzb_set_jpeg_coeffs_420: null argument");
  if (!c->mode420) { c->mode420 = true; c->have_coeffs = false; invalidate_compare_state(c); }
  const int16_t* src[3] = {c0, c1, c2};
  for (int k = 0; k < 3; ++k) {
    const size_t n2 = comp_blocks(c, k) * 64 * 2;
    CK(cudaMemcpyAsync(c->d_orig + k * c->cs, src[k], n2, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n2;
  }
  sync_check(c);
  c->have_orig_coeffs = true;
  c->packed_valid = false;
  GZB_END(c)
}

int gzb_compute_block_zeroing_order(gzb_ctx* c, int comp_mask, gzb_coeff_data* out) {
  GZB_TRY(c)
  if (!c->block_cmp) return fail(c, GZB_ERR_STATE, "gzb_compute_block_zeroing_order: StartBlockComparisons not called");
  if (!c->have_coeffs || !c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_compute_block_zeroing_order: coefficients missing");
  if (comp_mask < 1 || comp_mask > 7) return fail(c, GZB_ERR_BAD_ARG, "comp_mask must be in 1..7");
  if (c->mode420 && comp_mask != 1 && comp_mask != 6) return fail(c, GZB_ERR_BAD_ARG, "4:2:0: comp_mask must be 1 (luma) or 6 (chroma)");
  run_zeroing(c, comp_mask, 0);
  const size_t nrec = static_cast<size_t>(192) * zeroing_units(c, comp_mask);
  CK(cudaMemcpyAsync(out, c->d_order, sizeof(gzb_coeff_data) * nrec, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += sizeof(gzb_coeff_data) * nrec;
  unsigned* h_tie = reinterpret_cast<unsigned*>(c->h_pinned) + 1001;   // bytes 4004.. of the pinned page
  CK(cudaMemcpyAsync(h_tie, c->d_scalars + 3, 4, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 4;
  sync_check(c);
  c->zeroing_tie_blocks = *h_tie;
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  GZB_END(c)
}

int gzb_compute_block_zeroing_candidates_range(gzb_ctx* c, int comp_mask, int block_begin, int block_end, int* offsets,
                                               uint8_t* cand_idx, float* cand_err, size_t cap, size_t* n_out) {
  GZB_TRY(c)
  if (!c->block_cmp) return fail(c, GZB_ERR_STATE, "gzb_compute_block_zeroing_candidates: StartBlockComparisons not called");
  if (!c->have_coeffs || !c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_compute_block_zeroing_candidates: coefficients missing");
  if (comp_mask < 1 || comp_mask > 7 || !offsets || !n_out || block_begin < 0 || block_end > zeroing_units(c, comp_mask) || block_begin > block_end)
    return fail(c, GZB_ERR_BAD_ARG, "gzb_compute_block_zeroing_candidates: bad argument");
  if (c->mode420 && comp_mask != 1 && comp_mask != 6) return fail(c, GZB_ERR_BAD_ARG, "4:2:0: comp_mask must be 1 (luma) or 6 (chroma)");
  const int nloc = block_end - block_begin;
  if (nloc == 0) { offsets[0] = 0; *n_out = 0; return GZB_OK; }
  // d_tmp (6 planes of scratch) holds counts | offsets | packed err | packed idx
  int* d_counts; int* d_offsets; float* d_err; uint8_t* d_idx;
  packed_ptrs(c, &d_counts, &d_offsets, &d_err, &d_idx);
  if (!(c->packed_valid && c->packed_mask == comp_mask && c->packed_b0 == block_begin && c->packed_b1 == block_end)) {
    run_zeroing(c, comp_mask, 0, block_begin, block_end);  // a repeated call only re-fetches
    const CoeffRec* recs = reinterpret_cast<const CoeffRec*>(c->d_order) + static_cast<size_t>(block_begin) * 192;
    const int g = (nloc * 32 + 255) / 256;
    KLAUNCH(c, KC_MISC, k_count_candidates<<<g, 256, 0, c->stream>>>(recs, nloc, c->target, d_counts));
    KLAUNCH(c, KC_MISC, k_scan_counts<<<1, 1024, 0, c->stream>>>(d_counts, nloc, d_offsets));
    KLAUNCH(c, KC_MISC, k_pack_candidates<<<g, 256, 0, c->stream>>>(recs, nloc, c->target, d_offsets, d_idx, d_err));
    c->packed_valid = true;
    c->packed_mask = comp_mask;
    c->packed_b0 = block_begin;
    c->packed_b1 = block_end;
  }
  CK(cudaMemcpyAsync(offsets, d_offsets, (static_cast<size_t>(nloc) + 1) * sizeof(int), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<size_t>(nloc) + 1) * sizeof(int);
  unsigned* h_tie = reinterpret_cast<unsigned*>(c->h_pinned) + 1001;   // bytes 4004.. of the pinned page
  CK(cudaMemcpyAsync(h_tie, c->d_scalars + 3, 4, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 4;
  sync_check(c);
  c->zeroing_tie_blocks = *h_tie;
  CK(cudaEventElapsedTime(&c->last_ms, c->ev0, c->ev1));
  const size_t n = static_cast<size_t>(offsets[nloc]);
  *n_out = n;
  if (n > 0 && cand_idx && cap >= n) {   // cand_err == NULL: the errors stay on the device (gzb_be_begin)
    CK(cudaMemcpyAsync(cand_idx, d_idx, n, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += n;
    if (cand_err) { CK(cudaMemcpyAsync(cand_err, d_err, n * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += n * sizeof(float); }
    sync_check(c);
  }
  GZB_END(c)
}

int gzb_compute_block_zeroing_candidates(gzb_ctx* c, int comp_mask, int* offsets, uint8_t* cand_idx, float* cand_err,
                                         size_t cap, size_t* n_out) {
  if (!c) return GZB_ERR_BAD_ARG;
  return gzb_compute_block_zeroing_candidates_range(c, comp_mask, 0, zeroing_units(c, comp_mask), offsets, cand_idx, cand_err, cap, n_out);
}

// ---- entropy-coded segment on the device (gzb_huffman.cuh) -----------------------------------
// Scratch: the zeroing-order record array (1536 bytes per block), idle outside the zeroing search.
namespace {
struct HuffScratch {
  unsigned int* unit_bits;
  unsigned int* cta_bits;
  unsigned long long* cta_off;
  unsigned int* hist;            // [3][16] dc | [3][256] ac
  unsigned long long* out2;
  HuffDeviceTables* tables;
  unsigned int* words;
  size_t stream_cap;             // bytes
};
HuffScratch huff_scratch(gzb_ctx* c) {
  HuffScratch h;
  uint8_t* base = reinterpret_cast<uint8_t*>(c->d_order);
  const size_t total = sizeof(gzb_coeff_data) * 192 * static_cast<size_t>(c->nblocks);
  const size_t nunits = 3 * (c->cs / 64);   // upper bound for every layout
  size_t o = 0;
  auto take = [&](size_t bytes) { uint8_t* p = base + o; o += (bytes + 255) & ~size_t(255); return p; };
  const size_t nctas = (nunits + kHuffThreads - 1) / kHuffThreads;
  h.cta_off = reinterpret_cast<unsigned long long*>(take((nctas + 1) * 8));
  h.cta_bits = reinterpret_cast<unsigned int*>(take(nctas * 4));
  h.unit_bits = reinterpret_cast<unsigned int*>(take(nunits * 4));
  h.hist = reinterpret_cast<unsigned int*>(take((48 + 768) * 4));
  h.out2 = reinterpret_cast<unsigned long long*>(take(16));
  h.tables = reinterpret_cast<HuffDeviceTables*>(take(sizeof(HuffDeviceTables)));
  h.words = reinterpret_cast<unsigned int*>(base + o);
  h.stream_cap = total > o + 64 ? total - o - 64 : 0;
  return h;
}
// Scan layout of the resident candidate written with `ncomp` components.
HuffLayout huff_layout(const gzb_ctx* c, int ncomp, long long* nunits) {
  HuffLayout L;
  L.ncomp = ncomp; L.mcw = c->mcw; L.bw = c->bw; L.cbw0 = c->cbw0;
  if (!c->mode420) { L.mode = 0; *nunits = static_cast<long long>(ncomp) * c->nblocks; }
  else if (ncomp == 3) { L.mode = 1; *nunits = 6ll * c->mcw * c->mch; }
  else { L.mode = 2; *nunits = c->nblocks; }
  return L;
}
}  // namespace

int gzb_candidate_symbol_histograms(gzb_ctx* c, const int* q192, uint32_t* dc_hist48, uint32_t* ac_hist768) {
  return gzb_candidate_symbol_histograms_n(c, q192, 3, dc_hist48, ac_hist768);
}

int gzb_candidate_symbol_histograms_n(gzb_ctx* c, const int* q192, int ncomp, uint32_t* dc_hist48, uint32_t* ac_hist768) {
  GZB_TRY(c)
  if (ncomp != 1 && ncomp != 3) return fail(c, GZB_ERR_BAD_ARG, "gzb_candidate_symbol_histograms: ncomp must be 1 or 3");
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_candidate_symbol_histograms: no candidate coefficients");
  if (!dc_hist48 || !ac_hist768) return fail(c, GZB_ERR_BAD_ARG, "gzb_candidate_symbol_histograms: null argument");
  // the candidate may still be rendering on the main stream (render_candidate reads d_q too): wait first
  CK(cudaStreamWaitEvent(c->stream2, c->ev_cand, 0));
  if (q192) { CK(cudaMemcpyAsync(c->d_q, q192, 192 * sizeof(int), cudaMemcpyHostToDevice, c->stream2)); c->h2d_bytes += 192 * sizeof(int); }
  const HuffScratch h = huff_scratch(c);
  long long nunits = 0;
  const HuffLayout L = huff_layout(c, ncomp, &nunits);
  CK(cudaMemsetAsync(h.hist, 0, (48 + 768) * 4, c->stream2));
  KLAUNCH_S(c, c->stream2, KC_HUFFMAN, k_huff_histogram<<<static_cast<unsigned>((nunits + kHuffThreads - 1) / kHuffThreads), kHuffThreads, 0, c->stream2>>>(
      c->d_coef, c->cs, c->d_q, L, nunits, h.hist, h.hist + 48));
  uint32_t* hh = reinterpret_cast<uint32_t*>(c->h_pinned) + 128;  // bytes 512.. of the pinned page (Compare owns the start)
  CK(cudaMemcpyAsync(hh, h.hist, (48 + 768) * 4, cudaMemcpyDeviceToHost, c->stream2)); c->d2h_bytes += (48 + 768) * 4;
  sync_check2(c);
  memcpy(dc_hist48, hh, 48 * 4);
  memcpy(ac_hist768, hh + 48, 768 * 4);
  GZB_END(c)
}

int gzb_candidate_entropy_code(gzb_ctx* c, int ncomp, const uint16_t* dc_code, const uint8_t* dc_len,
                               const uint16_t* ac_code, const uint8_t* ac_len, uint64_t* scan_bytes, uint64_t* ff_bytes) {
  return gzb_candidate_entropy_code_sized(c, ncomp, dc_code, dc_len, ac_code, ac_len, ~0ull, scan_bytes, ff_bytes);
}

int gzb_candidate_entropy_code_sized(gzb_ctx* c, int ncomp, const uint16_t* dc_code, const uint8_t* dc_len,
                                     const uint16_t* ac_code, const uint8_t* ac_len, uint64_t total_bits_known,
                                     uint64_t* scan_bytes, uint64_t* ff_bytes) {
  GZB_TRY(c)
  if (!c->have_coeffs) return fail(c, GZB_ERR_STATE, "gzb_candidate_entropy_code: no candidate coefficients");
  if ((ncomp != 1 && ncomp != 3) || !dc_code || !dc_len || !ac_code || !ac_len || !scan_bytes || !ff_bytes)
    return fail(c, GZB_ERR_BAD_ARG, "gzb_candidate_entropy_code: bad argument");
  const HuffScratch h = huff_scratch(c);
  HuffDeviceTables t;
  memcpy(t.dc_code, dc_code, sizeof(t.dc_code));
  memcpy(t.dc_len, dc_len, sizeof(t.dc_len));
  memcpy(t.ac_code, ac_code, sizeof(t.ac_code));
  memcpy(t.ac_len, ac_len, sizeof(t.ac_len));
  CK(cudaStreamWaitEvent(c->stream2, c->ev_cand, 0));   // the candidate may still be rendering on the main stream
  CK(cudaMemcpyAsync(h.tables, &t, sizeof(t), cudaMemcpyHostToDevice, c->stream2)); c->h2d_bytes += sizeof(t);
  long long nunits = 0;
  const HuffLayout L = huff_layout(c, ncomp, &nunits);
  const unsigned grid = static_cast<unsigned>((nunits + kHuffThreads - 1) / kHuffThreads);
  const size_t cs = c->cs;
  KLAUNCH_S(c, c->stream2, KC_HUFFMAN, k_huff_code<false><<<grid, kHuffThreads, 0, c->stream2>>>(c->d_coef, cs, c->d_q, L, nunits, h.tables,
                                                                                  h.unit_bits, h.cta_bits, nullptr, nullptr));
  KLAUNCH_S(c, c->stream2, KC_HUFFMAN, k_scan_u64<<<1, 1024, 0, c->stream2>>>(h.cta_bits, static_cast<long long>(grid), h.cta_off));
  unsigned long long* hp = reinterpret_cast<unsigned long long*>(c->h_pinned) + 32;  // bytes 256..
  unsigned long long total_bits = total_bits_known;
  if (total_bits_known == ~0ull) {   // the caller does not know the size: read the total back before the emission pass
    CK(cudaMemcpyAsync(hp, h.cta_off + grid, 8, cudaMemcpyDeviceToHost, c->stream2)); c->d2h_bytes += 8;
    sync_check2(c);
    total_bits = hp[0];
  }
  const size_t nbytes = static_cast<size_t>((total_bits + 7) / 8);
  if (nbytes + 8 > h.stream_cap) return fail(c, GZB_ERR_UNSUPPORTED, "gzb_candidate_entropy_code: scan larger than the device buffer");
  CK(cudaMemsetAsync(h.words, 0, (nbytes + 7) & ~size_t(3), c->stream2));
  CK(cudaMemsetAsync(h.out2, 0, 16, c->stream2));
  KLAUNCH_S(c, c->stream2, KC_HUFFMAN, k_huff_code<true><<<grid, kHuffThreads, 0, c->stream2>>>(c->d_coef, cs, c->d_q, L, nunits, h.tables,
                                                                                 h.unit_bits, nullptr, h.cta_off, h.words));
  const unsigned fgrid = static_cast<unsigned>(std::max<size_t>(1, std::min<size_t>(static_cast<size_t>(c->sm_count) * 8, (nbytes + 4095) / 4096)));
  KLAUNCH_S(c, c->stream2, KC_HUFFMAN, k_huff_finish<<<fgrid, 256, 0, c->stream2>>>(reinterpret_cast<unsigned char*>(h.words), h.cta_off + grid, h.out2));
  CK(cudaMemcpyAsync(hp, h.out2, 16, cudaMemcpyDeviceToHost, c->stream2)); c->d2h_bytes += 16;
  sync_check2(c);
  if (hp[0] != nbytes) return fail(c, GZB_ERR_STATE, "gzb_candidate_entropy_code: the size given does not match the candidate");
  *scan_bytes = hp[0];
  *ff_bytes = hp[1];
  GZB_END(c)
}

int gzb_candidate_fetch_scan(gzb_ctx* c, uint8_t* out, uint64_t nbytes) {
  GZB_TRY(c)
  const HuffScratch h = huff_scratch(c);
  if (!out || nbytes > h.stream_cap) return fail(c, GZB_ERR_BAD_ARG, "gzb_candidate_fetch_scan: bad argument");
  CK(cudaMemcpyAsync(out, h.words, nbytes, cudaMemcpyDeviceToHost, c->stream2)); c->d2h_bytes += nbytes;
  sync_check2(c);
  GZB_END(c)
}

int gzb_dct_double(int device, double* blocks, size_t nblocks, int inverse) {
  if (!blocks) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_dct_double: null argument");
  if (nblocks == 0) return GZB_OK;
  try {
    CK(cudaSetDevice(device));
    init_device_tables(device);
    double* d = nullptr;
    dmalloc(&d, nblocks * 64);
    CK(cudaMemcpy(d, blocks, nblocks * 64 * sizeof(double), cudaMemcpyHostToDevice));
    k_dct_double<<<static_cast<unsigned>((nblocks + 3) / 4), 256>>>(d, nblocks, inverse);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    CK(cudaMemcpy(blocks, d, nblocks * 64 * sizeof(double), cudaMemcpyDeviceToHost));
    cudaFree(d);
  } catch (const std::string& e) { return fail(nullptr, GZB_ERR_CUDA, e); }
  return GZB_OK;
}

// ComputeBlockErrorAdjustmentWeights (butteraugli_comparator.cc:160-222) into c->d_weight; returns the block count.
static int launch_block_weights(gzb_ctx* c, const float* dm, int direction, int max_block_dist, double target_mul, int factor) {
  const double target = static_cast<double>(c->target) * target_mul;
  const int bs = 8 * factor, pbw = (c->W + bs - 1) / bs, pbh = (c->H + bs - 1) / bs, nb = pbw * pbh;
  const int g = (nb + 255) / 256;
  KLAUNCH(c, KC_WEIGHTS, k_block_max<<<g, 256, 0, c->stream>>>(dm, c->P, c->W, c->H, pbw, pbh, bs, c->d_bmax));
  KLAUNCH(c, KC_WEIGHTS, k_block_flags<<<g, 256, 0, c->stream>>>(c->d_bmax, pbw, pbh, direction, max_block_dist, target, c->d_flags));
  KLAUNCH(c, KC_WEIGHTS, k_block_weights<<<g, 256, 0, c->stream>>>(c->d_flags, pbw, pbh, direction, max_block_dist, c->d_weight));
  return nb;
}

int gzb_compute_block_error_adjustment_weights_f(gzb_ctx* c, int direction, int max_block_dist, double target_mul,
                                                 int factor, const float* distmap, float* block_weight) {
  GZB_TRY(c)
  if (factor != 1 && factor != 2) return fail(c, GZB_ERR_BAD_ARG, "gzb_compute_block_error_adjustment_weights: factor must be 1 or 2");
  const float* dm = c->d_diffmap;
  if (distmap) {  // a caller-provided map is staged in the blur scratch; the resident map is kept
    CK(cudaMemcpy2DAsync(c->d_tmp, c->P * sizeof(float), distmap, c->W * sizeof(float), c->W * sizeof(float), c->H,
                         cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H);
    dm = c->d_tmp;
    c->packed_valid = false;
  } else if (!c->have_distmap) {
    return fail(c, GZB_ERR_STATE, "gzb_compute_block_error_adjustment_weights: no distance map");
  }
  const int nb = launch_block_weights(c, dm, direction, max_block_dist, target_mul, factor);
  CK(cudaMemcpyAsync(block_weight, c->d_weight, sizeof(float) * nb, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (sizeof(float) * nb);
  sync_check(c);
  GZB_END(c)
}

int gzb_compute_block_error_adjustment_weights(gzb_ctx* c, int direction, int max_block_dist, double target_mul,
                                               const float* distmap, float* block_weight) {
  return gzb_compute_block_error_adjustment_weights_f(c, direction, max_block_dist, target_mul, 1, distmap, block_weight);
}

// ---- SelectFrequencyBackEnd on the device (gzb_backend.cuh) --------------------------------------
namespace {
const size_t kBePinState = 0;                      // BeState + small entries
const size_t kBePinHist = 96 << 10;                // 816 counters + scalars
const size_t kBePinGather = 104 << 10;             // gathered block states
const size_t kBePinBytes = kBePinGather + sizeof(BeBlockState) * kBeSmallMax;
static_assert(sizeof(BeState) + sizeof(BeEntry) * kBeSmallMax <= kBePinHist, "pinned layout");
static_assert(kBeSmallMax == GZB_BE_MAX_ENTRIES && kBeMaxRet == GZB_BE_MAX_RANGES, "include/gzb200.h");

// (Re)allocates the back-end arrays for num_blocks units and `total` candidates / order entries.
void be_reserve(gzb_ctx* c, int num_blocks, size_t total) {
  gzb_ctx::Backend& B = c->be;
  const size_t nb = static_cast<size_t>(num_blocks), T = std::max<size_t>(total, 1);
  const size_t ntiles = T / kBeTile + 2;
  size_t o = 0;
  auto take = [&](size_t bytes) { const size_t at = o; o += (bytes + 255) & ~size_t(255); return at; };
  const size_t o_state = take(sizeof(BeState) + sizeof(BeEntry) * kBeSmallMax);
  const size_t o_hist = take((48 + 768 + 8) * 4);
  const size_t o_off = take((nb + 1) * 4), o_li = take(nb * 4), o_me = take(nb * 4), o_cnt = take(nb * 4);
  const size_t o_offs = take((nb + 1) * 4), o_pc = take(nb * 4), o_cs = take(1024 * 4);
  const size_t o_req = take(kBeSmallMax * 4), o_out = take(sizeof(BeBlockState) * kBeSmallMax);
  const size_t o_tcl = take(ntiles * 4), o_tcr = take(ntiles * 4);
  const size_t o_idx = take(T), o_err = take(T * 4), o_order = take(T * 8), o_lp = take(T * 4), o_rp = take(T * 4);
  if (c->slab.be_cap < o) {
    CK(cudaStreamSynchronize(c->stream));
    if (c->slab.be) cudaFree(c->slab.be);
    c->slab.be = nullptr;
    c->slab.be_cap = 0;
    const size_t want = o + o / 4;   // headroom: the next image of a batch rarely has exactly as many candidates
    CK(cudaMalloc(&c->slab.be, want));
    c->slab.be_cap = want;
  }
  if (c->slab.be_pinned_cap < kBePinBytes) {
    if (c->slab.be_pinned) cudaFreeHost(c->slab.be_pinned);
    c->slab.be_pinned = nullptr;
    c->slab.be_pinned_cap = 0;
    CK(cudaMallocHost(&c->slab.be_pinned, kBePinBytes));
    c->slab.be_pinned_cap = kBePinBytes;
  }
  // (per device; the call is cheap)
  CK(cudaFuncSetAttribute(k_be_local, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kBeLocalSmemBytes)));
  char* base = static_cast<char*>(c->slab.be);
  B.st = reinterpret_cast<BeState*>(base + o_state);
  B.small = reinterpret_cast<BeEntry*>(base + o_state + sizeof(BeState));
  B.hist = reinterpret_cast<unsigned int*>(base + o_hist);
  B.flag = B.hist + 48 + 768;
  B.cand_off = reinterpret_cast<int*>(base + o_off);
  B.last_index = reinterpret_cast<int*>(base + o_li);
  B.max_err = reinterpret_cast<float*>(base + o_me);
  B.counts = reinterpret_cast<int*>(base + o_cnt);
  B.offsets = reinterpret_cast<int*>(base + o_offs);
  B.pcount = reinterpret_cast<unsigned*>(base + o_pc);
  B.chunk_sums = reinterpret_cast<int*>(base + o_cs);
  B.req_blocks = reinterpret_cast<int*>(base + o_req);
  B.req_out = reinterpret_cast<BeBlockState*>(base + o_out);
  B.tcl = reinterpret_cast<unsigned*>(base + o_tcl);
  B.tcr = reinterpret_cast<unsigned*>(base + o_tcr);
  B.cand_idx = reinterpret_cast<uint8_t*>(base + o_idx);
  B.cand_err = reinterpret_cast<float*>(base + o_err);
  B.order = reinterpret_cast<BeEntry*>(base + o_order);
  B.lpos = reinterpret_cast<unsigned*>(base + o_lp);
  B.rpos = reinterpret_cast<unsigned*>(base + o_rp);
  if (B.select_grid == 0) {
    static const int want = getenv("GZB_BE_CTAS_PER_SM") ? atoi(getenv("GZB_BE_CTAS_PER_SM")) : 2;
    B.select_grid = c->sm_count * std::max(1, std::min(8, want));
  }
}
char* be_pinned(gzb_ctx* c) { return static_cast<char*>(c->slab.be_pinned); }
BeCands be_cands(const gzb_ctx* c) { return BeCands{c->be.cand_off, c->be.cand_idx, c->be.cand_err, c->be.total}; }
}  // namespace

namespace {
// the part of gzb_be_begin before / after the candidate lists are in place
void be_begin_setup(gzb_ctx* c, int comp_mask, int units, size_t total) {
  gzb_ctx::Backend& B = c->be;
  be_reserve(c, units, total);
  B.comp_mask = comp_mask;
  B.factor = (c->mode420 && (comp_mask & 6)) ? 2 : 1;
  B.num_blocks = units;
  B.total = static_cast<int>(total);
  const int bs = 8 * B.factor;
  B.geom.num_blocks = units;
  B.geom.pass_bw = (c->W + bs - 1) / bs;
  B.geom.coef_bw = comp_bw(c, (comp_mask & 1) ? 0 : 1);
  B.geom.cs = c->cs;
  B.geom.factor = B.factor;
  B.geom.bw = c->bw;
  B.geom.bh = c->bh;
  B.geom.blk_changed = c->d_blk_changed;
}
void be_begin_finish(gzb_ctx* c, int units) {
  gzb_ctx::Backend& B = c->be;
  CK(cudaMemsetAsync(B.st, 0, sizeof(BeState), c->stream));
  CK(cudaMemsetAsync(B.last_index, 0, static_cast<size_t>(units) * 4, c->stream));
  CK(cudaMemsetAsync(B.max_err, 0, static_cast<size_t>(units) * 4, c->stream));
  CK(cudaMemsetAsync(B.pcount, 0, static_cast<size_t>(units) * 4, c->stream));
  sync_check(c);   // the host arrays may be released by the caller
  B.active = true;
  B.n = 0;
}
}  // namespace

int gzb_be_begin(gzb_ctx* c, int comp_mask, const int* offsets, const uint8_t* cand_idx, const float* cand_err, size_t total) {
  GZB_TRY(c)
  if (!c->have_coeffs || !c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_be_begin: coefficients missing");
  if (comp_mask < 1 || comp_mask > 7 || (c->mode420 && comp_mask != 1 && comp_mask != 6)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_begin: bad comp_mask");
  if (total >= (size_t(1) << 31)) return fail(c, GZB_ERR_UNSUPPORTED, "gzb_be_begin: too many candidates");
  gzb_ctx::Backend& B = c->be;
  const int units = zeroing_units(c, comp_mask);
  const bool resident = offsets == nullptr;
  if (resident && !(c->packed_valid && c->packed_mask == comp_mask && c->packed_b0 == 0 && c->packed_b1 == units))
    return fail(c, GZB_ERR_STATE, "gzb_be_begin: no resident candidate lists for this comp_mask");
  if (!resident && total > 0 && (!cand_idx || !cand_err)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_begin: null candidate arrays");
  if (!resident && offsets[units] != static_cast<int>(total)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_begin: offsets[num_blocks] != total");
  be_begin_setup(c, comp_mask, units, total);
  if (resident) {
    int* d_counts; int* d_offsets; float* d_err; uint8_t* d_idx;
    packed_ptrs(c, &d_counts, &d_offsets, &d_err, &d_idx);
    CK(cudaMemcpyAsync(B.cand_off, d_offsets, (static_cast<size_t>(units) + 1) * 4, cudaMemcpyDeviceToDevice, c->stream));
    if (total) {
      CK(cudaMemcpyAsync(B.cand_idx, d_idx, total, cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaMemcpyAsync(B.cand_err, d_err, total * 4, cudaMemcpyDeviceToDevice, c->stream));
    }
  } else {
    CK(cudaMemcpyAsync(B.cand_off, offsets, (static_cast<size_t>(units) + 1) * 4, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (static_cast<size_t>(units) + 1) * 4;
    if (total) {
      CK(cudaMemcpyAsync(B.cand_idx, cand_idx, total, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += total;
      CK(cudaMemcpyAsync(B.cand_err, cand_err, total * 4, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += total * 4;
    }
  }
  be_begin_finish(c, units);
  GZB_END(c)
}

int gzb_be_begin_gathered(gzb_ctx* c, int comp_mask, int world, int rank, const int* global_offsets, const uint64_t* counts,
                          int (*allgather_device)(void*, const void*, size_t, void*), void* user, int begin, uint8_t* cand_idx_out) {
  GZB_TRY(c)
  if (!c->have_coeffs || !c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_be_begin_gathered: coefficients missing");
  if (comp_mask < 1 || comp_mask > 7 || (c->mode420 && comp_mask != 1 && comp_mask != 6)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_begin_gathered: bad comp_mask");
  if (world < 1 || rank < 0 || rank >= world || !counts || !allgather_device || (begin && !global_offsets))
    return fail(c, GZB_ERR_BAD_ARG, "gzb_be_begin_gathered: bad argument");
  const int units = zeroing_units(c, comp_mask);
  const int b0 = static_cast<int>(static_cast<int64_t>(units) * rank / world), b1 = static_cast<int>(static_cast<int64_t>(units) * (rank + 1) / world);
  if (!(c->packed_valid && c->packed_mask == comp_mask && c->packed_b0 == b0 && c->packed_b1 == b1))
    return fail(c, GZB_ERR_STATE, "gzb_be_begin_gathered: no resident candidate lists for this rank's block range");
  size_t total = 0, maxn = 0;
  std::vector<size_t> base(static_cast<size_t>(world) + 1, 0);
  for (int r = 0; r < world; ++r) {
    maxn = std::max<size_t>(maxn, counts[r]);
    base[r + 1] = base[r] + counts[r];
  }
  total = base[world];
  if (total >= (size_t(1) << 31)) return fail(c, GZB_ERR_UNSUPPORTED, "gzb_be_begin_gathered: too many candidates");
  if (begin && global_offsets[units] != static_cast<int>(total)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_begin_gathered: offsets[num_blocks] != total");
  // one record per rank: [errors (4 * maxn) | coefficient indices (maxn)], padded to the longest list
  const size_t rec = (5 * maxn + 255) & ~size_t(255);
  char* tmp = nullptr;
  if (rec) CK(cudaMalloc(&tmp, rec * (static_cast<size_t>(world) + 1)));
  struct Free { char* p; ~Free() { if (p) cudaFree(p); } } guard{tmp};
  char* d_send = tmp;
  char* d_recv = tmp ? tmp + rec : nullptr;
  const size_t n = counts[rank];
  if (rec) {
    int* d_counts; int* d_offsets; float* d_err; uint8_t* d_idx;
    packed_ptrs(c, &d_counts, &d_offsets, &d_err, &d_idx);
    if (n) {
      CK(cudaMemcpyAsync(d_send, d_err, n * 4, cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaMemcpyAsync(d_send + 4 * maxn, d_idx, n, cudaMemcpyDeviceToDevice, c->stream));
    }
    sync_check(c);
    if (allgather_device(user, d_send, rec, d_recv) != 0) return fail(c, GZB_ERR_CUDA, "gzb_be_begin_gathered: the device all-gather failed");
  }
  if (begin) {
    gzb_ctx::Backend& B = c->be;
    be_begin_setup(c, comp_mask, units, total);
    CK(cudaMemcpyAsync(B.cand_off, global_offsets, (static_cast<size_t>(units) + 1) * 4, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (static_cast<size_t>(units) + 1) * 4;
    for (int r = 0; r < world; ++r) {
      const size_t nr = counts[r];
      if (!nr) continue;
      const char* src = d_recv + static_cast<size_t>(r) * rec;
      CK(cudaMemcpyAsync(B.cand_err + base[r], src, nr * 4, cudaMemcpyDeviceToDevice, c->stream));
      CK(cudaMemcpyAsync(B.cand_idx + base[r], src + 4 * maxn, nr, cudaMemcpyDeviceToDevice, c->stream));
    }
    if (cand_idx_out && total) { CK(cudaMemcpyAsync(cand_idx_out, B.cand_idx, total, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += total; }
    be_begin_finish(c, units);
  }
  GZB_END(c)
}

int gzb_test_memcpy_d2d(void* dst, const void* src, size_t nbytes) {
  return cudaMemcpy(dst, src, nbytes, cudaMemcpyDeviceToDevice) == cudaSuccess ? GZB_OK : GZB_ERR_CUDA;
}

int gzb_be_build_order(gzb_ctx* c, int direction, double target_mul, float below_limit, uint64_t* n, int* blocks_to_change,
                       uint64_t* below, int* rblock_out) {
  GZB_TRY(c)
  gzb_ctx::Backend& B = c->be;
  if (!B.active) return fail(c, GZB_ERR_STATE, "gzb_be_build_order: gzb_be_begin not called");
  if (!c->have_distmap) return fail(c, GZB_ERR_STATE, "gzb_be_build_order: no distance map");
  if (!n || !blocks_to_change || !below || (direction != 1 && direction != -1)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_build_order: bad argument");
  const int nb = B.num_blocks;
  const BeCands cands = be_cands(c);
  unsigned* hs = reinterpret_cast<unsigned*>(be_pinned(c) + kBePinState);
  int rblock = 1;
  for (; rblock <= 4; ++rblock) {
    launch_block_weights(c, c->d_diffmap, direction, rblock, target_mul, B.factor);
    CK(cudaMemsetAsync(B.st, 0, 16, c->stream));   // n, blocks_to_change, below, changed_blocks
    KLAUNCH(c, KC_MISC, k_be_count<<<(nb + 255) / 256, 256, 0, c->stream>>>(cands, c->d_weight, B.last_index, nb, direction, B.counts, B.st));
    {
      const int nchunks = (nb + kScanChunk - 1) / kScanChunk;
      if (nchunks <= 1024) {
        KLAUNCH(c, KC_MISC, k_scan_chunk_sums<<<nchunks, 256, 0, c->stream>>>(B.counts, nb, B.chunk_sums));
        KLAUNCH(c, KC_MISC, k_scan_chunk_offsets<<<1, 1024, 0, c->stream>>>(B.chunk_sums, nchunks, B.offsets, nb));
        KLAUNCH(c, KC_MISC, k_scan_chunk_apply<<<nchunks, 256, 0, c->stream>>>(B.counts, nb, B.chunk_sums, B.offsets));
      } else {
        KLAUNCH(c, KC_MISC, k_scan_counts<<<1, 1024, 0, c->stream>>>(B.counts, nb, B.offsets));
      }
    }
    KLAUNCH(c, KC_MISC, k_be_set_n<<<1, 1, 0, c->stream>>>(B.offsets, nb, B.st));
    KLAUNCH(c, KC_MISC, k_be_fill<<<(nb * 32 + 255) / 256, 256, 0, c->stream>>>(cands, c->d_weight, B.last_index, B.max_err, B.counts, B.offsets, nb,
                                                                                direction, below_limit, B.order, B.st));
    KLAUNCH(c, KC_MISC, k_be_sort_begin<<<1, 1, 0, c->stream>>>(B.st, -1));
    CK(cudaMemcpyAsync(hs, B.st, 16, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 16;
    sync_check(c);
    if (hs[0] > 0) break;
  }
  B.n = hs[0];
  B.h_top = 0;   // one pending range [0, n): gzb_be_select falls back to n
  *n = hs[0];
  *blocks_to_change = static_cast<int>(hs[1]);
  *below = hs[2];
  if (rblock_out) *rblock_out = std::min(rblock, 4);
  c->packed_valid = false;
  GZB_END(c)
}

int gzb_be_select_ranges(gzb_ctx* c, uint64_t p_set, int small_max, uint64_t want_end, int* status, int* nranges,
                         gzb_be_range* ranges, gzb_order_entry* entries_out) {
  GZB_TRY(c)
  gzb_ctx::Backend& B = c->be;
  if (!B.active) return fail(c, GZB_ERR_STATE, "gzb_be_select: gzb_be_begin not called");
  if (!status || !nranges || !ranges || small_max < 16 || small_max > kBeSmallMax) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_select: bad argument");
  const unsigned pset = static_cast<unsigned>(std::min<uint64_t>(p_set, B.n));
  KLAUNCH(c, KC_MISC, k_be_select_begin<<<1, 1, 0, c->stream>>>(B.order, B.st, pset, static_cast<unsigned>(small_max),
                                                                static_cast<unsigned>(std::min<uint64_t>(want_end, B.n))));
  // (one range is all a caller without want_end gets: the copy stays short)
  const size_t bytes = sizeof(BeState) + sizeof(BeEntry) * static_cast<size_t>(want_end > p_set ? kBeSmallMax : small_max);
  BeState* hst = reinterpret_cast<BeState*>(be_pinned(c) + kBePinState);
  // Length of the range the sort will work on first, from the host's copy of the pending ranges: a long range
  // needs about log2(length / kBeLocalMax) grid-level partitions before one CTA can take over. The kernels of a
  // level return at once when no long range is in flight, so a wrong guess only costs empty launches or one
  // more round trip.
  size_t len = B.n;
  for (int i = B.h_top - 1; i >= 0; --i) {
    if (B.h_stack[i].last <= pset) continue;
    len = B.h_stack[i].last - B.h_stack[i].first;
    break;
  }
  const int G = B.select_grid;
  for (;;) {
    int levels = 0;
    for (size_t l = len; l > kBeLocalMax; l >>= 1) ++levels;
    if (levels > 0) levels += 2;
    for (int l = 0; l < levels; ++l) {
      KLAUNCH(c, KC_MISC, k_be_tiles_count<<<G, kBeThreads, 0, c->stream>>>(B.order, B.tcl, B.tcr, B.st));
      KLAUNCH(c, KC_MISC, k_be_tiles_lists<<<G, kBeThreads, 0, c->stream>>>(B.order, B.lpos, B.rpos, B.tcl, B.tcr, B.st));
      KLAUNCH(c, KC_MISC, k_be_swap<<<G, kBeThreads, 0, c->stream>>>(B.order, B.lpos, B.rpos, B.st));
    }
    KLAUNCH(c, KC_MISC, k_be_local<<<1, kBeLocalThreads, kBeLocalSmemBytes, c->stream>>>(B.order, B.st, B.small));
    CK(cudaMemcpyAsync(hst, B.st, bytes, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += bytes;
    sync_check(c);
    if (hst->status != BE_RUNNING || hst->nret > 0) break;
    len = std::max<size_t>(static_cast<size_t>(hst->last - hst->first), 4 * static_cast<size_t>(kBeLocalMax));   // a long range is still in flight
  }
  ++B.selects;
  B.levels += hst->levels;
  B.h_top = std::max(0, std::min(hst->top, kBeStack));
  memcpy(B.h_stack, hst->stack, sizeof(BeRange) * B.h_top);
  *status = hst->status;
  *nranges = 0;
  const int nret = static_cast<int>(std::min<unsigned>(hst->nret, kBeMaxRet));
  if (nret > 0) {
    // short ranges, consecutive, their entries back to back behind the state (a range the sort stopped at
    // afterwards for another reason is reported by the next call)
    *status = BE_SMALL;
    *nranges = nret;
    for (int i = 0; i < nret; ++i) ranges[i] = gzb_be_range{hst->ret[i].first, hst->ret[i].last, hst->ret[i].depth, 0};
    if (entries_out) memcpy(entries_out, reinterpret_cast<const char*>(hst) + sizeof(BeState), sizeof(BeEntry) * hst->ret_total);
  } else if (hst->status == BE_EMPTY || hst->top <= 0) {
    *status = BE_EMPTY;
  } else {   // BE_HEAP: the caller fetches the range itself
    const BeRange r = hst->stack[hst->top - 1];
    *nranges = 1;
    ranges[0] = gzb_be_range{r.first, r.last, r.depth, 0};
  }
  GZB_END(c)
}

int gzb_be_select(gzb_ctx* c, uint64_t p_set, int small_max, int* status, uint64_t* first, uint64_t* last, int* depth,
                  gzb_order_entry* entries_out) {
  if (!c) return GZB_ERR_BAD_ARG;
  if (!first || !last || !depth) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_select: bad argument");
  int nranges = 0;
  gzb_be_range r[GZB_BE_MAX_RANGES];
  const int rc = gzb_be_select_ranges(c, p_set, small_max, 0, status, &nranges, r, entries_out);
  if (rc != GZB_OK) return rc;
  if (nranges > 0) { *first = r[0].first; *last = r[0].last; *depth = r[0].depth; }
  else { *first = *last = c->be.n; *depth = 0; }
  return GZB_OK;
}

int gzb_be_fetch_order(gzb_ctx* c, uint64_t first, gzb_order_entry* out, size_t n) {
  GZB_TRY(c)
  if (!c->be.active || !out || first + n > c->be.n) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_fetch_order: bad argument");
  CK(cudaMemcpyAsync(out, c->be.order + first, n * 8, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += n * 8;
  sync_check(c);
  GZB_END(c)
}

int gzb_be_store_order(gzb_ctx* c, uint64_t first, const gzb_order_entry* in, size_t n) {
  GZB_TRY(c)
  if (!c->be.active || !in || first + n > c->be.n) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_store_order: bad argument");
  // (pageable source: the copy is staged before the call returns; the rest is ordered by the stream)
  CK(cudaMemcpyAsync(c->be.order + first, in, n * 8, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n * 8;
  GZB_END(c)
}

namespace {
// an "up" step never reads the requantised values: its records are only the head of a BeBlockState
size_t be_state_stride(int direction) { return direction < 0 ? sizeof(BeBlockState) : offsetof(BeBlockState, requant); }
void be_launch_gather(gzb_ctx* c, const int* blocks, int nreq, int direction) {
  gzb_ctx::Backend& B = c->be;
  const size_t stride = be_state_stride(direction);
  CK(cudaMemcpyAsync(B.req_blocks, blocks, static_cast<size_t>(nreq) * 4, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += static_cast<size_t>(nreq) * 4;
  KLAUNCH(c, KC_MISC, k_be_gather<<<(nreq * 32 + 255) / 256, 256, 0, c->stream>>>(B.geom, B.req_blocks, nreq, c->d_coef, c->d_orig, c->d_q, B.last_index,
                                                                                B.pcount, B.comp_mask, direction < 0 ? 1 : 0,
                                                                                reinterpret_cast<char*>(B.req_out), stride));
  CK(cudaMemcpyAsync(be_pinned(c) + kBePinGather, B.req_out, stride * static_cast<size_t>(nreq), cudaMemcpyDeviceToHost, c->stream));
  c->d2h_bytes += stride * static_cast<size_t>(nreq);
}
void be_copy_states(gzb_ctx* c, gzb_be_block_state* out, int nreq, int direction) {
  const char* src = be_pinned(c) + kBePinGather;
  if (direction < 0) { memcpy(out, src, sizeof(BeBlockState) * static_cast<size_t>(nreq)); return; }
  const size_t stride = be_state_stride(direction);
  for (int i = 0; i < nreq; ++i) memcpy(reinterpret_cast<char*>(out + i), src + stride * static_cast<size_t>(i), stride);
}
}  // namespace

int gzb_be_apply_prefix(gzb_ctx* c, uint64_t p, int direction, int hist_ncomp, uint32_t* ac_hist768, int* changed_blocks,
                        const int* blocks, int nreq, gzb_be_block_state* states_out) {
  GZB_TRY(c)
  static_assert(sizeof(gzb_be_block_state) == sizeof(BeBlockState), "gzb_be_block_state layout");
  gzb_ctx::Backend& B = c->be;
  if (!B.active) return fail(c, GZB_ERR_STATE, "gzb_be_apply_prefix: gzb_be_begin not called");
  if (p > B.n || nreq < 0 || nreq > kBeSmallMax || (nreq > 0 && (!blocks || !states_out)) || (hist_ncomp != 1 && hist_ncomp != 3))
    return fail(c, GZB_ERR_BAD_ARG, "gzb_be_apply_prefix: bad argument");
  for (int i = 0; i < nreq; ++i) if (blocks[i] < 0 || blocks[i] >= B.num_blocks) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_apply_prefix: block out of range");
  unsigned* hh = reinterpret_cast<unsigned*>(be_pinned(c) + kBePinHist);
  CK(cudaMemsetAsync(&B.st->changed_blocks, 0, 4, c->stream));
  if (p > 0) {
    const unsigned pp = static_cast<unsigned>(p);
    KLAUNCH(c, KC_MISC, k_be_prefix_count<<<(pp + 255) / 256, 256, 0, c->stream>>>(B.order, pp, B.pcount));
    KLAUNCH(c, KC_MISC, k_be_apply_prefix<<<(B.num_blocks + 255) / 256, 256, 0, c->stream>>>(be_cands(c), B.geom, B.pcount, direction, c->d_orig, c->d_q,
                                                                                            c->d_coef, B.last_index, B.st));
    // the samples of most blocks change: the next Compare is a full one
    c->changed_overflow = true;
    c->changed.clear();
  }
  if (ac_hist768) {
    long long nunits = 0;
    const HuffLayout L = huff_layout(c, hist_ncomp, &nunits);
    CK(cudaMemsetAsync(B.hist, 0, (48 + 768) * 4, c->stream));
    KLAUNCH(c, KC_HUFFMAN, k_huff_histogram<<<static_cast<unsigned>((nunits + kHuffThreads - 1) / kHuffThreads), kHuffThreads, 0, c->stream>>>(
        c->d_coef, c->cs, c->d_q, L, nunits, B.hist, B.hist + 48));
    CK(cudaMemcpyAsync(hh, B.hist, (48 + 768) * 4, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (48 + 768) * 4;
  }
  CK(cudaMemcpyAsync(hh + 48 + 768, &B.st->changed_blocks, 4, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 4;
  if (nreq > 0) be_launch_gather(c, blocks, nreq, direction);
  sync_check(c);
  if (ac_hist768) memcpy(ac_hist768, hh + 48, 768 * 4);
  if (changed_blocks) *changed_blocks = static_cast<int>(hh[48 + 768]);
  if (nreq > 0) be_copy_states(c, states_out, nreq, direction);
  GZB_END(c)
}

int gzb_be_gather(gzb_ctx* c, const int* blocks, int nreq, int direction, gzb_be_block_state* states_out) {
  GZB_TRY(c)
  gzb_ctx::Backend& B = c->be;
  if (!B.active) return fail(c, GZB_ERR_STATE, "gzb_be_gather: gzb_be_begin not called");
  if (nreq < 0 || nreq > kBeSmallMax || (nreq > 0 && (!blocks || !states_out))) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_gather: bad argument");
  for (int i = 0; i < nreq; ++i) if (blocks[i] < 0 || blocks[i] >= B.num_blocks) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_gather: block out of range");
  if (nreq == 0) return GZB_OK;
  be_launch_gather(c, blocks, nreq, direction);
  sync_check(c);
  be_copy_states(c, states_out, nreq, direction);
  GZB_END(c)
}

int gzb_be_finish_iteration(gzb_ctx* c, const int32_t* blocks, const uint8_t* cidx, const int16_t* val, size_t n, int direction,
                            float val_threshold) {
  GZB_TRY(c)
  gzb_ctx::Backend& B = c->be;
  if (!B.active) return fail(c, GZB_ERR_STATE, "gzb_be_finish_iteration: gzb_be_begin not called");
  if (n > 0 && (!blocks || !cidx || !val)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_finish_iteration: null argument");
  for (size_t i = 0; i < n; ++i)
    if (blocks[i] < 0 || blocks[i] >= B.num_blocks || cidx[i] >= 192 || !(B.comp_mask >> (cidx[i] >> 6) & 1))
      return fail(c, GZB_ERR_BAD_ARG, "gzb_be_finish_iteration: flip out of range");
  if (c->inter_valid && !c->changed_overflow) {
    // pixel rectangles whose samples change (see gzb_update_coeffs)
    if (n + c->changed.size() > 256) {
      c->changed_overflow = true;
      c->changed.clear();
    } else {
      for (size_t i = 0; i < n; ++i) {
        const int comp = cidx[i] >> 6;
        const int bx = blocks[i] % B.geom.pass_bw, by = blocks[i] / B.geom.pass_bw;
        int4 q;
        if (c->mode420 && comp > 0) q = make_int4(16 * bx - 1, 16 * by - 1, 16 * bx + 16, 16 * by + 16);
        else q = make_int4(8 * bx, 8 * by, 8 * bx + 7, 8 * by + 7);
        bool dup = false;
        for (const int4& o : c->changed) dup = dup || (o.x == q.x && o.y == q.y && o.z == q.z && o.w == q.w);
        if (!dup) c->changed.push_back(q);
      }
    }
  }
  if (n > 0) {
    if (n > c->upd_cap) {
      if (c->upd_own && c->d_upd) cudaFree(c->d_upd);
      c->upd_cap = n + n / 2 + 1024;
      dmalloc(&c->d_upd, c->upd_cap * 8);
      c->upd_own = true;
    }
    uint8_t* stage = c->d_upd;
    const size_t cap = c->upd_cap;
    CK(cudaMemcpyAsync(stage, blocks, n * 4, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n * 4;
    CK(cudaMemcpyAsync(stage + cap * 4, val, n * 2, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n * 2;
    CK(cudaMemcpyAsync(stage + cap * 6, cidx, n, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += n;
    KLAUNCH(c, KC_MISC, k_be_apply_walk<<<static_cast<unsigned>((n + 255) / 256), 256, 0, c->stream>>>(
        B.geom, reinterpret_cast<const int*>(stage), stage + cap * 6, reinterpret_cast<const int16_t*>(stage + cap * 4), static_cast<int>(n),
        direction, c->d_coef, B.last_index));
  }
  KLAUNCH(c, KC_MISC, k_be_update_max_err<<<(B.num_blocks + 255) / 256, 256, 0, c->stream>>>(c->d_weight, val_threshold, direction, B.num_blocks, B.max_err));
  CK(cudaMemsetAsync(B.pcount, 0, static_cast<size_t>(B.num_blocks) * 4, c->stream));
  c->packed_valid = false;
  render_candidate(c, kCoeffKeep);   // no host sync: later calls are ordered by the stream / ev_cand
  GZB_END(c)
}

int gzb_input_is_gray(gzb_ctx* c, int* gray) {
  GZB_TRY(c)
  if (!gray) return fail(c, GZB_ERR_BAD_ARG, "gzb_input_is_gray: null argument");
  if (!c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_input_is_gray: no input coefficients");
  unsigned* d_flag = c->d_scalars + 2;
  CK(cudaMemsetAsync(d_flag, 0, 4, c->stream));
  for (int k = 1; k < 3; ++k) {
    const size_t n = comp_blocks(c, k) * 64;
    KLAUNCH(c, KC_MISC, k_any_nonzero<<<static_cast<unsigned>((n / 8 + 256) / 256), 256, 0, c->stream>>>(c->d_orig + k * c->cs, n, d_flag));
  }
  unsigned* h = reinterpret_cast<unsigned*>(c->h_pinned) + 1000;   // bytes 4000.. of the pinned page
  CK(cudaMemcpyAsync(h, d_flag, 4, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += 4;
  sync_check(c);
  *gray = h[0] ? 0 : 1;
  GZB_END(c)
}

int gzb_be_stats(const gzb_ctx* c, unsigned long long* selects, unsigned long long* levels) {
  if (!c) return GZB_ERR_BAD_ARG;
  if (selects) *selects = c->be.selects;
  if (levels) *levels = c->be.levels;
  return GZB_OK;
}

// Test hook: makes `entries` the order of the context (no candidate lists behind it).
int gzb_be_test_load_order_depth(gzb_ctx* c, const gzb_order_entry* entries, size_t n, int depth);
int gzb_be_test_load_order(gzb_ctx* c, const gzb_order_entry* entries, size_t n) {
  return gzb_be_test_load_order_depth(c, entries, n, -1);
}
// depth >= 0: the sort starts with this depth budget instead of 2 * log2(n) (reaches the heap-sort fallback)
int gzb_be_test_load_order_depth(gzb_ctx* c, const gzb_order_entry* entries, size_t n, int depth) {
  GZB_TRY(c)
  if (!entries || n == 0 || n >= (size_t(1) << 31)) return fail(c, GZB_ERR_BAD_ARG, "gzb_be_test_load_order: bad argument");
  gzb_ctx::Backend& B = c->be;
  be_reserve(c, c->nblocks, n);
  B.active = true;
  B.num_blocks = c->nblocks;
  B.total = 0;
  B.n = n;
  B.h_top = 0;
  CK(cudaMemsetAsync(B.st, 0, sizeof(BeState), c->stream));
  CK(cudaMemcpyAsync(B.order, entries, n * 8, cudaMemcpyHostToDevice, c->stream));
  const unsigned nn = static_cast<unsigned>(n);
  CK(cudaMemcpyAsync(&B.st->n, &nn, 4, cudaMemcpyHostToDevice, c->stream));
  KLAUNCH(c, KC_MISC, k_be_sort_begin<<<1, 1, 0, c->stream>>>(B.st, depth));
  sync_check(c);
  GZB_END(c)
}

// Test hook (host arithmetic): the fused-multiply-add quotient of gamma_rational (gzb_device_math.cuh) against
// IEEE division for EVERY float argument in [0, 1024]; returns the number of mismatches.
unsigned long long gzb_test_gamma_division(void) {
  unsigned long long bad = 0;
  for (uint32_t bits = 0; bits <= 0x44800000u; ++bits) {
    float xf;
    memcpy(&xf, &bits, 4);
    const double x = static_cast<double>(xf) - 0.770000000000000;
    volatile double c = kGammaRange;
    const double want = x / c;
    const double got = div_by_gamma_range(x);
    bad += memcmp(&want, &got, 8) != 0 ? 1 : 0;
  }
  return bad;
}

// Non-FMA FP64 peak of the device in Gflop/s (DADD + DMUL issued back to back on all SMs): the
// denominator for the FP64 rate of the search kernels, which are built with -fmad=false.
int gzb_measure_fp64_peak(int device, double* gflops) {
  if (!gflops) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_measure_fp64_peak: null argument");
  try {
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    const int ctas = sms * 8, iters = 8192;
    double* d = nullptr;
    dmalloc(&d, static_cast<size_t>(ctas) * 256);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {   // first repetition warms up (clocks, instruction cache)
      CK(cudaEventRecord(e0, nullptr));
      k_fp64_peak<<<ctas, 256>>>(d, iters, 1.0000001, 1e-9);
      CK(cudaEventRecord(e1, nullptr));
      CK(cudaEventSynchronize(e1));
      CK(cudaGetLastError());
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      const double flops = static_cast<double>(ctas) * 256 * iters * 16.0;
      if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    *gflops = best;
  } catch (const std::string& e) { return fail(nullptr, GZB_ERR_CUDA, e); }
  return GZB_OK;
}

// ---- YUV 4:2:0 ---------------------------------------------------------------------------------
int gzb_downsample_420(gzb_ctx* c) {
  GZB_TRY(c)
  if (!c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_downsample_420: gzb_set_jpeg_coeffs not called");
  if (c->mode420) return fail(c, GZB_ERR_STATE, "gzb_downsample_420: the image is already 4:2:0");
  const int W = c->W, H = c->H, P = c->P;
  const size_t cs = c->cs, ps = c->ps;
  c->packed_valid = false;
  // scratch (idle outside a Compare): yuv planes | normalised planes | conv temps | u8 maps
  float* yuv = c->d_mh;            // [3]
  float* nrm = c->d_mh + 3 * ps;   // [3]
  float* ts = c->d_bl;             // sharpen H pass
  float* tb = c->d_bl + ps;        // blur H pass
  float* yuv2 = c->d_bl + 2 * ps;  // [3] output of a PreProcessChannel call
  uint8_t* maps = reinterpret_cast<uint8_t*>(c->d_tmp);   // 6 byte maps of P*HP (d_tmp holds 24 bytes per pixel)
  uint8_t *m_dark = maps, *m_red = maps + ps, *m_a = maps + 2 * ps, *m_b = maps + 3 * ps, *m_sharp = maps + 4 * ps, *m_blur = maps + 5 * ps;
  // ToFloatPixels of the three 4:4:4 components (output_image.cc:553-556)
  for (int k = 0; k < 3; ++k)
    KLAUNCH(c, KC_MISC, k420_to_float<<<(c->nblocks + 3) / 4, 256, 0, c->stream>>>(c->d_orig + k * cs, c->bw, c->nblocks, W, H, P, yuv + k * ps));
  // PreProcessChannel(w, h, 2, 1.3f, 0.5f, ...) then (w, h, 1, ...) (output_image.cc:558-561)
  PreProcTaps tp;
  {
    auto normal = [](double x, double sigma) { return std::exp(-x * x / (2 * sigma * sigma)) * 0.3989422804014327 / sigma; };
    const double sig_sharpen = static_cast<double>(1.3f), sig_blur = 1.3;
    double ks[5], kb[5], ss = 0, sb = 0;
    for (int i = 0; i < 5; ++i) { ks[i] = normal(1.0 * i - 2, sig_sharpen); kb[i] = normal(1.0 * i - 2, sig_blur); }
    for (int i = 0; i < 5; ++i) { ss += ks[i]; sb += kb[i]; }
    for (int i = 0; i < 5; ++i) { tp.sharpen[i] = static_cast<float>(ks[i]); tp.blur[i] = static_cast<float>(kb[i]); }
    tp.sharpen_mul = static_cast<float>(1.0 / ss);
    tp.blur_mul = static_cast<float>(1.0 / sb);
  }
  dim3 gpx((W + 255) / 256, H);
  float* in = yuv;
  float* out = yuv2;
  for (int channel = 2; channel >= 1; --channel) {
    KLAUNCH(c, KC_MISC, k_pp_norm<<<gpx, 256, 0, c->stream>>>(in, ps, W, H, P, channel, nrm, m_dark, m_red));
    // Erode x3 on darkmap: dark -> a -> b -> dark
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_dark, m_a, W, H, P, 0));
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_a, m_b, W, H, P, 0));
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_b, m_dark, W, H, P, 0));
    // Dilate x3 on redmap
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_red, m_a, W, H, P, 1));
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_a, m_b, W, H, P, 1));
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_b, m_red, W, H, P, 1));
    KLAUNCH(c, KC_MISC, k_pp_blurmap<<<gpx, 256, 0, c->stream>>>(nrm, ps, W, H, P, channel, m_dark, m_red, m_sharp, m_a));
    // Erode x2 on blurmap: a -> b -> blur
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_a, m_b, W, H, P, 0));
    KLAUNCH(c, KC_MISC, k_pp_morph<<<gpx, 256, 0, c->stream>>>(m_b, m_blur, W, H, P, 0));
    KLAUNCH(c, KC_MISC, k_pp_conv_h<<<gpx, 256, 0, c->stream>>>(nrm + channel * ps, W, H, P, tp, ts, tb));
    KLAUNCH(c, KC_MISC, k_pp_final<<<gpx, 256, 0, c->stream>>>(nrm, ps, W, H, P, channel, tp, 0.5f, ts, tb, m_sharp, m_blur, out));
    std::swap(in, out);
  }
  // `in` now holds the pre-processed planes. SetDownsampledCoefficients for Cb and Cr (563-570);
  // SaveToJpegData: luma re-laid out MCU-padded (608-632), staged through the candidate array.
  const int nmb = c->mcw * c->mch;
  CK(cudaMemcpyAsync(c->d_coef, c->d_orig, static_cast<size_t>(c->nblocks) * 128, cudaMemcpyDeviceToDevice, c->stream));
  KLAUNCH(c, KC_MISC, k420_pad_luma<<<(c->cbw0 * c->cbh0 * 8 + 255) / 256, 256, 0, c->stream>>>(c->d_coef, c->bw, c->bh, c->cbw0, c->cbh0, c->d_orig));
  for (int k = 1; k < 3; ++k)
    KLAUNCH(c, KC_MISC, k420_downsample_chroma<<<(nmb + 3) / 4, 256, 0, c->stream>>>(in + k * ps, W, H, P, c->mcw, nmb, c->d_orig + k * cs));
  sync_check(c);
  c->mode420 = true;
  c->have_coeffs = false;
  invalidate_compare_state(c);   // the scratch planes used above are the Compare's
  GZB_END(c)
}

int gzb_component_dims(const gzb_ctx* c, int comp, int* blocks_w, int* blocks_h, int* factor) {
  if (!c || comp < 0 || comp > 2) return GZB_ERR_BAD_ARG;
  if (blocks_w) *blocks_w = comp_bw(c, comp);
  if (blocks_h) *blocks_h = comp_bh(c, comp);
  if (factor) *factor = (c->mode420 && comp > 0) ? 2 : 1;
  return GZB_OK;
}

int gzb_get_jpeg_coeffs(gzb_ctx* c, int16_t* c0, int16_t* c1, int16_t* c2) {
  GZB_TRY(c)
  if (!c->have_orig_coeffs) return fail(c, GZB_ERR_STATE, "gzb_get_jpeg_coeffs: no input coefficients");
  int16_t* dst[3] = {c0, c1, c2};
  for (int k = 0; k < 3; ++k) {
    if (!dst[k]) continue;
    const size_t n2 = comp_blocks(c, k) * 64 * 2;
    CK(cudaMemcpyAsync(dst[k], c->d_orig + k * c->cs, n2, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += n2;
  }
  sync_check(c);
  GZB_END(c)
}

int gzb_debug_fetch(gzb_ctx* c, const char* name, float* out, size_t cap, size_t* n_out) {
  GZB_TRY(c)
  const std::string s(name ? name : "");
  const size_t n = static_cast<size_t>(c->W) * c->H, rn = static_cast<size_t>(c->rxs) * c->rys;
  const float* planes = nullptr; int nplanes = 0;
  const float* flat = nullptr; size_t flat_n = 0;
  if (s == "xyb0") { planes = c->d_xyb0; nplanes = 3; }
  else if (s == "xyb1") { planes = c->d_xyb1; nplanes = 3; }
  else if (s == "mhic0") { planes = c->d_mh; nplanes = 3; }
  else if (s == "mhic1") { planes = c->d_mh + 3 * c->ps; nplanes = 3; }
  else if (s == "mask_front") { planes = c->d_bl; nplanes = 3; }
  else if (s == "diffmap") { planes = c->d_diffmap; nplanes = 1; }
  else if (s == "edge_map") { flat = c->d_edm; flat_n = 3 * rn; }
  else if (s == "block_dc") { flat = c->d_dc; flat_n = 3 * rn; }
  else if (s == "block_ac") { flat = c->d_ac; flat_n = 3 * rn; }
  else if (s == "combined_sqrt") { planes = nullptr; }
  else return fail(c, GZB_ERR_BAD_ARG, "gzb_debug_fetch: unknown name " + s);
  if (planes) {
    const size_t total = n * nplanes;
    if (n_out) *n_out = total;
    if (cap >= total)
      for (int k = 0; k < nplanes; ++k)
        CK(cudaMemcpy2DAsync(out + k * n, c->W * sizeof(float), planes + k * c->ps, c->P * sizeof(float),
                             c->W * sizeof(float), c->H, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H));
  } else if (flat) {
    if (n_out) *n_out = flat_n;
    if (cap >= flat_n) CK(cudaMemcpyAsync(out, flat, flat_n * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (flat_n * sizeof(float));
    if (cap >= flat_n && s == "block_ac") {   // the reference's block_diff_ac includes the EdgeDetectorLowFreq term
      std::vector<float> lft(flat_n);
      CK(cudaMemcpyAsync(lft.data(), c->d_lft, flat_n * sizeof(float), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (flat_n * sizeof(float));
      CK(cudaStreamSynchronize(c->stream));
      for (size_t i = 0; i < flat_n; ++i) out[i] = out[i] + lft[i];
    }
  } else {  // combined_sqrt: res map with pitch
    if (n_out) *n_out = rn;
    if (cap >= rn)
      CK(cudaMemcpy2DAsync(out, c->rxs * sizeof(float), c->d_sq, c->sqp * sizeof(float), c->rxs * sizeof(float), c->rys,
                           cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<unsigned long long>(c->rxs * sizeof(float)) * (c->rys));
  }
  sync_check(c);
  GZB_END(c)
}

int gzb_profile_enable(gzb_ctx* c, int on) {
  if (!c) return GZB_ERR_BAD_ARG;
  c->prof.on = on != 0;
  return GZB_OK;
}
int gzb_profile_count(void) { return KC_COUNT; }
const char* gzb_profile_name(int i) { return i >= 0 && i < KC_COUNT ? kClassNames[i] : ""; }
int gzb_profile_get(gzb_ctx* c, int i, double* ms, unsigned long long* launches) {
  if (!c || i < 0 || i >= KC_COUNT) return GZB_ERR_BAD_ARG;
  if (ms) *ms = c->prof.ms[i];
  if (launches) *launches = c->prof.n[i];
  return GZB_OK;
}
int gzb_profile_reset(gzb_ctx* c) {
  if (!c) return GZB_ERR_BAD_ARG;
  for (int i = 0; i < KC_COUNT; ++i) { c->prof.ms[i] = 0; c->prof.n[i] = 0; }
  return GZB_OK;
}
int gzb_get_transfer_bytes(const gzb_ctx* c, unsigned long long* h2d, unsigned long long* d2h) {
  if (!c) return GZB_ERR_BAD_ARG;
  if (h2d) *h2d = c->h2d_bytes;
  if (d2h) *d2h = c->d2h_bytes;
  return GZB_OK;
}
unsigned gzb_last_zeroing_tie_blocks(const gzb_ctx* c) { return c ? c->zeroing_tie_blocks : 0; }
float gzb_last_device_ms(const gzb_ctx* c) { return c ? c->last_ms : 0.f; }
unsigned long long gzb_launch_count(const gzb_ctx* c) { return c ? c->launches : 0; }
unsigned long long gzb_incremental_compare_count(const gzb_ctx* c) { return c ? c->incremental_compares : 0; }
unsigned long long gzb_fine_bdm_compare_count(const gzb_ctx* c) { return c ? c->fine_bdm_compares : 0; }

// ---- stage entry points ----------------------------------------------------------------------
int gzb_blur(int device, float* plane, size_t xsize, size_t ysize, double sigma, double border_ratio) {
  if (!plane || xsize == 0 || ysize == 0) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_blur: bad argument");
  try {
    CK(cudaSetDevice(device));
    init_device_tables(device);
    HostKernel hk;
    if (!make_host_kernel(sigma, &hk)) return fail(nullptr, GZB_ERR_UNSUPPORTED, "gzb_blur: sigma too large (radius > 32)");
    if (hk.step > 4) return fail(nullptr, GZB_ERR_UNSUPPORTED, "gzb_blur: decimation > 4");
    const int W = static_cast<int>(xsize), H = static_cast<int>(ysize), P = round_up(W, 32);
    float taps[kMaxTaps] = {0};
    memcpy(taps, hk.taps, sizeof(float) * (2 * hk.r + 1));
    CK(cudaMemcpyToSymbol(c_taps, taps, sizeof(taps), sizeof(float) * kMaxTaps * kBUser));
    BlurPlan pl;
    pl.build_decimated(hk, kBUser, W, H, P, border_ratio, 1);
    pl.upload_own();
    float *d_in = nullptr, *d_tmp = nullptr, *d_out = nullptr;
    dmalloc(&d_in, static_cast<size_t>(P) * H);
    dmalloc(&d_tmp, pl.tmp_floats());
    dmalloc(&d_out, pl.out_floats());
    CK(cudaMemcpy2D(d_in, P * sizeof(float), plane, W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice));
    dim3 blk(32, 8);
    dim3 gh((pl.g.nx + pl.g.oxn - 1) / pl.g.oxn, (H + kBhRows - 1) / kBhRows, 1);
    k_blur_h<1><<<gh, blk>>>(d_in, 0, pl.g, pl.d_sx, d_tmp, 0, kAllDirty);
    dim3 gv((pl.g.nx + 31) / 32, (pl.g.ny + pl.g.oyn - 1) / pl.g.oyn, 1);
    k_blur_v<<<gv, blk>>>(d_tmp, 0, pl.g, pl.d_sy, d_out, 0, pl.g.tmp_pitch, kAllDirty);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    std::vector<float> small(pl.out_floats());
    CK(cudaMemcpy(small.data(), d_out, small.size() * sizeof(float), cudaMemcpyDeviceToHost));
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        plane[static_cast<size_t>(y) * W + x] = small[static_cast<size_t>(y / hk.step) * pl.g.tmp_pitch + x / hk.step];
    cudaFree(d_in); cudaFree(d_tmp); cudaFree(d_out);
    pl.release();
  } catch (const std::string& e) { return fail(nullptr, GZB_ERR_CUDA, e); }
  return GZB_OK;
}

int gzb_opsin_dynamics_image(int device, float* r, float* g, float* b, size_t xsize, size_t ysize) {
  if (!r || !g || !b || xsize == 0 || ysize == 0) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_opsin_dynamics_image: bad argument");
  try {
    CK(cudaSetDevice(device));
    init_device_tables(device);
    const int W = static_cast<int>(xsize), H = static_cast<int>(ysize), P = round_up(W, 32);
    const size_t ps = static_cast<size_t>(P) * H;
    BlurPlan pl;
    pl.build_decimated(g_hk[kB11], kB11, W, H, P, 0.0, 1);
    pl.upload_own();
    float *d_in = nullptr, *d_out = nullptr;
    dmalloc(&d_in, 3 * ps);
    dmalloc(&d_out, 3 * ps);
    float* src[3] = {r, g, b};
    for (int k = 0; k < 3; ++k)
      CK(cudaMemcpy2D(d_in + k * ps, P * sizeof(float), src[k], W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice));
    dim3 blk(32, 8), grd((W + 31) / 32, (H + 31) / 32);
    k_opsin_dynamics_f32<<<grd, blk>>>(d_in, ps, W, H, P, pl.d_sx, pl.d_sy, d_out, ps);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    for (int k = 0; k < 3; ++k)
      CK(cudaMemcpy2D(src[k], W * sizeof(float), d_out + k * ps, P * sizeof(float), W * sizeof(float), H, cudaMemcpyDeviceToHost));
    cudaFree(d_in); cudaFree(d_out);
    pl.release();
  } catch (const std::string& e) { return fail(nullptr, GZB_ERR_CUDA, e); }
  return GZB_OK;
}

int gzb_diffmap_opsin_dynamics_image(int device, float* result, const float* r, const float* g, const float* b,
                                     const float* r2, const float* g2, const float* b2, size_t xsize, size_t ysize,
                                     size_t step) {
  if (step != 3) return fail(nullptr, GZB_ERR_UNSUPPORTED, "gzb_diffmap_opsin_dynamics_image: step must be 3");
  if (!result || !r || !g || !b || !r2 || !g2 || !b2) return fail(nullptr, GZB_ERR_BAD_ARG, "null plane");
  if (xsize < 32 || ysize < 32) return fail(nullptr, GZB_ERR_TOO_SMALL, "image smaller than 32x32");
  gzb_ctx* c = alloc_ctx(device, static_cast<int>(xsize), static_cast<int>(ysize), 1.0f);
  if (!c) return GZB_ERR_CUDA;
  int rc = GZB_OK;
  try {
    const float* s0[3] = {r, g, b};
    const float* s1[3] = {r2, g2, b2};
    for (int k = 0; k < 3; ++k) {
      CK(cudaMemcpy2DAsync(c->d_xyb0 + k * c->ps, c->P * sizeof(float), s0[k], c->W * sizeof(float), c->W * sizeof(float), c->H, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H));
      CK(cudaMemcpy2DAsync(c->d_xyb1 + k * c->ps, c->P * sizeof(float), s1[k], c->W * sizeof(float), c->W * sizeof(float), c->H, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H));
    }
    run_diffmap(c, c->d_xyb0, c->d_xyb1);
    CK(cudaMemcpy2DAsync(result, c->W * sizeof(float), c->d_diffmap, c->P * sizeof(float), c->W * sizeof(float), c->H, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H));
    sync_check(c);
  } catch (const std::string& e) { rc = fail(nullptr, GZB_ERR_CUDA, e); }
  free_ctx(c);
  return rc;
}

// butteraugli::Mask (butteraugli.cc:1505-1566) as a standalone stage: the cuMask hook of the reference
// (clguetzli/cuguetzli.h:42-47, dispatched from clbutter_comparator.cpp:1651-1666). Host planes in and out.
int gzb_mask(int device, float* mask_r, float* mask_g, float* mask_b, float* maskdc_r, float* maskdc_g, float* maskdc_b,
             size_t xsize, size_t ysize, const float* r, const float* g, const float* b, const float* r2, const float* g2,
             const float* b2) {
  if (!mask_r || !mask_g || !mask_b || !maskdc_r || !maskdc_g || !maskdc_b || !r || !g || !b || !r2 || !g2 || !b2)
    return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_mask: null plane");
  if (xsize < 32 || ysize < 32) return fail(nullptr, GZB_ERR_TOO_SMALL, "gzb_mask: image smaller than 32x32");
  if (xsize >= (1u << 16) || ysize >= (1u << 16)) return fail(nullptr, GZB_ERR_BAD_ARG, "gzb_mask: image too large");
  gzb_ctx* c = alloc_ctx(device, static_cast<int>(xsize), static_cast<int>(ysize), 1.0f);
  if (!c) return GZB_ERR_CUDA;
  int rc = GZB_OK;
  BlurPlan pl[3];
  float* lat[3] = {nullptr, nullptr, nullptr};
  try {
    const int W = c->W, H = c->H, P = c->P;
    const float* s0[3] = {r, g, b};
    const float* s1[3] = {r2, g2, b2};
    for (int k = 0; k < 3; ++k) {
      CK(cudaMemcpy2DAsync(c->d_xyb0 + k * c->ps, P * sizeof(float), s0[k], W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice, c->stream));
      CK(cudaMemcpy2DAsync(c->d_xyb1 + k * c->ps, P * sizeof(float), s1[k], W * sizeof(float), W * sizeof(float), H, cudaMemcpyHostToDevice, c->stream));
      c->h2d_bytes += 2ull * W * sizeof(float) * H;
    }
    // DiffPrecompute + Average5x5 + MinSquareVal -> d_bl[0..2]; then the three blurs on their full decimated lattices
    dim3 blk(32, 8), grd((W + 31) / 32, (H + 31) / 32, 3);
    KLAUNCH(c, KC_MASK_FRONT, k_mask_front<<<grd, blk, 0, c->stream>>>(c->d_xyb0, c->d_xyb1, c->ps, W, H, P, c->d_bl, kAllDirty));
    const int kinds[3] = {kB9657, kB14264, kB4533};
    MaskSample ms;
    for (int k = 0; k < 3; ++k) {
      pl[k].build_decimated(g_hk[kinds[k]], kinds[k], W, H, P, 0.0, 1);
      pl[k].upload_own();
      dmalloc(&lat[k], pl[k].out_floats());
      run_blur(c, pl[k], c->d_bl + k * c->ps, 0, 1, lat[k], 0, pl[k].g.tmp_pitch);
      ms.m[k] = lat[k];
      ms.pitch[k] = pl[k].g.tmp_pitch;
      ms.x0[k] = pl[k].g.x0; ms.sx[k] = pl[k].g.sx; ms.y0[k] = pl[k].g.y0; ms.sy[k] = pl[k].g.sy;
    }
    // the six output planes reuse the MaskHighIntensityChange planes
    dim3 gpx((W + 31) / 32, (H + 7) / 8);
    KLAUNCH(c, KC_COMBINE, k_mask_full<<<gpx, blk, 0, c->stream>>>(ms, W, H, P, c->d_mh, c->d_mh + 3 * c->ps, c->ps));
    float* outs[6] = {mask_r, mask_g, mask_b, maskdc_r, maskdc_g, maskdc_b};
    for (int k = 0; k < 6; ++k) {
      CK(cudaMemcpy2DAsync(outs[k], W * sizeof(float), c->d_mh + k * c->ps, P * sizeof(float), W * sizeof(float), H, cudaMemcpyDeviceToHost, c->stream));
      c->d2h_bytes += static_cast<unsigned long long>(W) * sizeof(float) * H;
    }
    sync_check(c);
  } catch (const std::string& e) { rc = fail(nullptr, GZB_ERR_CUDA, e); }
  for (int k = 0; k < 3; ++k) { if (lat[k]) cudaFree(lat[k]); pl[k].release(); }
  free_ctx(c);
  return rc;
}

int gzb_butteraugli_srgb(int device, const uint8_t* rgb0, const uint8_t* rgb1, int width, int height, float* distance,
                         float* diffmap_out) {
  if (!rgb0 || !rgb1) return fail(nullptr, GZB_ERR_BAD_ARG, "null image");
  gzb_ctx* c = nullptr;
  int rc = gzb_create(device, width, height, rgb0, 1.0f, &c);
  if (rc != GZB_OK) return rc;
  try {
    const size_t n = static_cast<size_t>(3) * c->W * c->H;
    CK(cudaMemcpyAsync(c->d_stage_u8, rgb1, n, cudaMemcpyHostToDevice, c->stream)); c->h2d_bytes += (n);
    dim3 grd((c->W + 255) / 256, c->H);
    KLAUNCH(c, KC_MISC, k_deinterleave_rgb<<<grd, 256, 0, c->stream>>>(c->d_stage_u8, c->W, c->H, c->P, c->d_rgb1, c->ps));
    opsin_from_u8(c, c->d_rgb1, c->d_xyb1);
    run_diffmap(c, c->d_xyb0, c->d_xyb1);
    CK(cudaMemcpyAsync(c->h_pinned, c->d_scalars, sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (sizeof(unsigned int));
    if (diffmap_out)
      CK(cudaMemcpy2DAsync(diffmap_out, c->W * sizeof(float), c->d_diffmap, c->P * sizeof(float), c->W * sizeof(float), c->H, cudaMemcpyDeviceToHost, c->stream)); c->d2h_bytes += (static_cast<unsigned long long>(c->W * sizeof(float)) * (c->H));
    sync_check(c);
    if (distance) memcpy(distance, c->h_pinned, sizeof(float));
  } catch (const std::string& e) { rc = fail(nullptr, GZB_ERR_CUDA, e); }
  free_ctx(c);
  return rc;
}

}  // extern "C"
