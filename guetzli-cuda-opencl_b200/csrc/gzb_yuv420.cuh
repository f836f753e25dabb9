// gzb_yuv420.cuh -- the YUV 4:2:0 branch of the search on sm_100a:
//   * OutputImage::Downsample with the default DownsampleConfig (guetzli/output_image.cc:496-571):
//     ToFloatPixels (double IDCT, 100-122), PreProcessChannel (guetzli/preprocess_downsample.cc:
//     157-281) for channel 2 then 1, SetDownsampledCoefficients (2x2 average, double DCT, 496-531);
//   * the candidate image with factor-2 chroma: IDCT of the half-resolution chroma blocks, libjpeg's
//     "fancy" 9-3-3-1 upsampling (UpdatePixelsForBlock, output_image.cc:147-210) and ToPixels'
//     (p + 8 - (x & 1)) >> 4 (68-98);
//   * ComputeBlockZeroingOrder over 16x16 macro-blocks (guetzli/processor.cc:376-487 with
//     factor_x = factor_y = 2): a candidate changes the Cb/Cr block, the error is the maximum of the
//     CompareBlock errors of the (up to) four 8x8 luma-sized sub-blocks.
//
// Why the upsampled chroma can be computed per pixel instead of through the reference's
// block-by-block pixels_ updates: a stored sub-sampled value is idct << 4, so the upsampler's
// (9a + 3b + 3c + d) >> 4 is exact, and the "inverse upsampler" that reconstructs the neighbours'
// border samples, (9p00 - 3p01 - 3p10 + p11) >> 2, returns exactly the neighbour's sample. pixels_ is
// therefore always the fancy upsampling of ONE global sub-sampled plane S (with samples beyond the
// image replaced by the nearest sample inside: the x0 >= width_ / y0 >= height_ / x0 < 0 / y0 < 0
// rules of UpdatePixelsForBlock), whatever order blocks were updated in.
#pragma once
#include "gzb_device_math.cuh"
#include "gzb_zeroing.cuh"

namespace gzb {

// ---------------------------------------------------------------------------------------------
// One component: (optional quantise / scale) + integer IDCT -> u8 samples at (8bx, 8by) of `plane`.
// Same arithmetic as k_coeffs_to_rgb8; Y blocks are laid out cbw per row (MCU-padded), chroma blocks
// ceil(W/16) per row.
// ---------------------------------------------------------------------------------------------
template <int OP>
__global__ void __launch_bounds__(256)
k420_idct_comp(const int16_t* __restrict__ src, int16_t* __restrict__ dst, const int* __restrict__ q64,
               int cbw, int nblocks, int P, uint8_t* __restrict__ plane) {
  __shared__ int s_in[32][8][9];
  const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + lb;
  const bool live = b < nblocks;
  if (live) {
    const int4 v = reinterpret_cast<const int4*>(src)[static_cast<size_t>(b) * 8 + t];
    int e[8] = {static_cast<int16_t>(v.x & 0xffff), v.x >> 16, static_cast<int16_t>(v.y & 0xffff), v.y >> 16,
                static_cast<int16_t>(v.z & 0xffff), v.z >> 16, static_cast<int16_t>(v.w & 0xffff), v.w >> 16};
    if (OP != kCoeffKeep) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const int q = q64[8 * t + k];
        e[k] = (OP == kCoeffQuantize || OP == kCoeffQuantizeSrc) ? quantize_coeff(e[k], q) : static_cast<int>(static_cast<int16_t>(e[k] * q));
      }
      int4 o;
      o.x = (e[0] & 0xffff) | (e[1] << 16);
      o.y = (e[2] & 0xffff) | (e[3] << 16);
      o.z = (e[4] & 0xffff) | (e[5] << 16);
      o.w = (e[6] & 0xffff) | (e[7] << 16);
      reinterpret_cast<int4*>(dst)[static_cast<size_t>(b) * 8 + t] = o;
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) s_in[lb][t][k] = e[k];
  }
  __syncthreads();
  int col[8], out[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) col[u] = s_in[lb][u][t];
  __syncthreads();
  idct_1d(col, out);
#pragma unroll
  for (int y = 0; y < 8; ++y) s_in[lb][y][t] = idct_col_round(out[y]);
  __syncthreads();
  int row[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) row[u] = s_in[lb][t][u];
  idct_1d(row, out);
  if (!live) return;
  uint32_t w0 = 0, w1 = 0;
#pragma unroll
  for (int x = 0; x < 4; ++x) {
    w0 |= static_cast<uint32_t>(idct_row_round(out[x])) << (8 * x);
    w1 |= static_cast<uint32_t>(idct_row_round(out[4 + x])) << (8 * x);
  }
  const int bx = b % cbw, by = b / cbw;
  *reinterpret_cast<uint2*>(plane + static_cast<size_t>(8 * by + t) * P + 8 * bx) = make_uint2(w0, w1);
}

// Fancy-upsampled chroma sample of pixel (x, y): S is the half-resolution sample plane (pitch P),
// (xmax, ymax) the last sample position inside the image.
__device__ __forceinline__ int fancy_upsample_at(const uint8_t* __restrict__ S, int P, int x, int y, int xmax, int ymax) {
  const int X = min(x >> 1, xmax), Y = min(y >> 1, ymax);
  const int Xn = min(max(X + ((x & 1) ? 1 : -1), 0), xmax);
  const int Yn = min(max(Y + ((y & 1) ? 1 : -1), 0), ymax);
  const int v = 9 * S[static_cast<size_t>(Y) * P + X] + 3 * S[static_cast<size_t>(Yn) * P + X] +
                3 * S[static_cast<size_t>(Y) * P + Xn] + S[static_cast<size_t>(Yn) * P + Xn];
  return (v + 8 - (x & 1)) >> 4;
}

// OutputImage::ToSRGB of a 4:2:0 image: Y samples + upsampled chroma -> sRGB8 planes; the upsampled
// Cb / Cr samples are kept (cup) for the luma pass of the zeroing search.
__global__ void __launch_bounds__(256)
k420_render(const uint8_t* __restrict__ ycc, size_t plane_stride, int W, int H, int P,
            uint8_t* __restrict__ rgb, uint8_t* __restrict__ cup) {
  const int x0 = 4 * (blockIdx.x * blockDim.x + threadIdx.x);
  const int y = blockIdx.y;
  if (x0 >= W || y >= H) return;
  const int xmax = (W - 1) >> 1, ymax = (H - 1) >> 1;
  const size_t o = static_cast<size_t>(y) * P + x0;
  const uint32_t yw = *reinterpret_cast<const uint32_t*>(ycc + o);
  uint32_t wr = 0, wg = 0, wb = 0, wcb = 0, wcr = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = min(x0 + k, W - 1);
    const int cb = fancy_upsample_at(ycc + plane_stride, P, x, y, xmax, ymax);
    const int cr = fancy_upsample_at(ycc + 2 * plane_stride, P, x, y, xmax, ymax);
    int r, g, b;
    ycbcr_to_rgb(static_cast<int>((yw >> (8 * k)) & 0xff), cb, cr, r, g, b);
    wr |= static_cast<uint32_t>(r) << (8 * k);
    wg |= static_cast<uint32_t>(g) << (8 * k);
    wb |= static_cast<uint32_t>(b) << (8 * k);
    wcb |= static_cast<uint32_t>(cb) << (8 * k);
    wcr |= static_cast<uint32_t>(cr) << (8 * k);
  }
  *reinterpret_cast<uint32_t*>(rgb + o) = wr;
  *reinterpret_cast<uint32_t*>(rgb + plane_stride + o) = wg;
  *reinterpret_cast<uint32_t*>(rgb + 2 * plane_stride + o) = wb;
  *reinterpret_cast<uint32_t*>(cup + o) = wcb;
  *reinterpret_cast<uint32_t*>(cup + plane_stride + o) = wcr;
}

// ---------------------------------------------------------------------------------------------
// OutputImage::Downsample, default config.
// ---------------------------------------------------------------------------------------------
// ToFloatPixels (output_image.cc:100-122): double IDCT of a 4:4:4 component + 128 -> float plane.
__global__ void __launch_bounds__(256)
k420_to_float(const int16_t* __restrict__ coef, int bw, int nblocks, int W, int H, int P, float* __restrict__ out) {
  __shared__ double s_in[4][64];
  __shared__ double s_tmp[4][64];
  const int lb = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int b = blockIdx.x * 4 + lb;
  const bool live = b < nblocks;
  s_in[lb][t] = live ? static_cast<double>(coef[static_cast<size_t>(b) * 64 + t]) : 0.0;
  __syncthreads();
  {
    const int xo = t >> 3, col = t & 7;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += c_dct_matrix[8 * u + xo] * s_in[lb][8 * u + col];
    s_tmp[lb][8 * xo + col] = acc;
  }
  __syncthreads();
  const int row = t >> 3, xo = t & 7;
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += c_dct_matrix[8 * u + xo] * s_tmp[lb][8 * row + u];
  if (!live) return;
  const int x = 8 * (b % bw) + xo, y = 8 * (b / bw) + row;
  if (x < W && y < H) out[static_cast<size_t>(y) * P + x] = static_cast<float>(acc + 128.0);
}

struct PreProcTaps { float sharpen[5], blur[5]; float sharpen_mul, blur_mul; };

// Normalisation to [0,1] / [-0.5,0.5] and the dark / red maps (preprocess_downsample.cc:163-223).
__global__ void __launch_bounds__(256)
k_pp_norm(const float* __restrict__ yuv, size_t ps, int W, int H, int P, int channel, float* __restrict__ nrm,
          uint8_t* __restrict__ dark, uint8_t* __restrict__ red) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t i = static_cast<size_t>(y) * P + x;
  const float yy = static_cast<float>(static_cast<double>(yuv[i]) / 255.0);
  const float u = yuv[ps + i] / 255.0f - 0.5f;
  const float v = yuv[2 * ps + i] / 255.0f - 0.5f;
  nrm[i] = yy; nrm[ps + i] = u; nrm[2 * ps + i] = v;
  const float r = yy + 1.402f * v;
  const float g = yy - 0.34414f * u - 0.71414f * v;
  const float b = yy + 1.772f * u;
  bool d, rd;
  if (channel == 2) {
    d = static_cast<double>(g) < 0.85 && static_cast<double>(b) < 0.85 && static_cast<double>(r) < 0.9;
    rd = 2.116 * static_cast<double>(v) > -0.34414 * static_cast<double>(u) + 0.2 &&
         1.402 * static_cast<double>(v) > 1.772 * static_cast<double>(u) + 0.2;
  } else {
    d = static_cast<double>(r) < 0.85 && static_cast<double>(g) < 0.85 && static_cast<double>(b) < 0.9;
    rd = static_cast<double>(v) < 1.263 * static_cast<double>(u) - 0.1 &&
         static_cast<double>(u) > -0.33741 * static_cast<double>(v);
  }
  dark[i] = d ? 1 : 0;
  red[i] = rd ? 1 : 0;
}

// Erode / Dilate with the plus-shaped element, interior pixels only (preprocess_downsample.cc:109-137).
__global__ void __launch_bounds__(256)
k_pp_morph(const uint8_t* __restrict__ in, uint8_t* __restrict__ out, int W, int H, int P, int dilate) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t i = static_cast<size_t>(y) * P + x;
  uint8_t v = in[i];
  if (x >= 1 && x + 1 < W && y >= 1 && y + 1 < H) {
    if (dilate) v = (v || in[i - 1] || in[i + 1] || in[i - P] || in[i + P]) ? 1 : 0;
    else v = (v && in[i - 1] && in[i + 1] && in[i - P] && in[i + P]) ? 1 : 0;
  }
  out[i] = v;
}

// sharpenmap = red && dark; blurmap before its two erosions (preprocess_downsample.cc:225-258).
__global__ void __launch_bounds__(256)
k_pp_blurmap(const float* __restrict__ nrm, size_t ps, int W, int H, int P, int channel, const uint8_t* __restrict__ dark,
             const uint8_t* __restrict__ red, uint8_t* __restrict__ sharp, uint8_t* __restrict__ blurm) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t i = static_cast<size_t>(y) * P + x;
  const float* img = nrm + channel * ps;
  float edge = img[i];
  if (x >= 1 && x + 1 < W && y >= 1 && y + 1 < H) {   // Convolve2D with {0,-1,0,-1,4,-1,0,-1,0}
    float v = 0.0f;
    v += -1.0f * img[i - P];
    v += -1.0f * img[i - 1];
    v += 4.0f * img[i];
    v += -1.0f * img[i + 1];
    v += -1.0f * img[i + P];
    edge = v;
  }
  const bool sh = red[i] && dark[i];
  const double threshold = (channel == 2 ? 0.02 : 1.0) * 127.5;
  const float u = nrm[ps + i], v = nrm[2 * ps + i];
  bool bl = false;
  if (!sh && dark[i] && static_cast<double>(fabsf(edge)) < threshold &&
      static_cast<double>(v) < -0.162 * static_cast<double>(u)) bl = true;
  sharp[i] = sh ? 1 : 0;
  blurm[i] = bl ? 1 : 0;
}

// Horizontal pass of Convolve2X for the sharpen and the blur kernel at once (50-65).
__global__ void __launch_bounds__(256)
k_pp_conv_h(const float* __restrict__ img, int W, int H, int P, PreProcTaps tp, float* __restrict__ ts,
            float* __restrict__ tb) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t i = static_cast<size_t>(y) * P + x;
  float s = img[i], b = img[i];
  if (x >= 2 && x + 2 < W) {
    float vs = 0.0f, vb = 0.0f;
#pragma unroll
    for (int j = 0; j < 5; ++j) {
      const float p = img[i + j - 2];
      vs += tp.sharpen[j] * p;
      vb += tp.blur[j] * p;
    }
    s = vs * tp.sharpen_mul;
    b = vb * tp.blur_mul;
  }
  ts[i] = s;
  tb[i] = b;
}

// Vertical pass, Sharpen's unsharp mask, the per-pixel choice and the way back to 0..255 (66-80,
// 89-106, 260-280).
__global__ void __launch_bounds__(256)
k_pp_final(const float* __restrict__ nrm, size_t ps, int W, int H, int P, int channel, PreProcTaps tp, float amount,
           const float* __restrict__ ts, const float* __restrict__ tb, const uint8_t* __restrict__ sharp,
           const uint8_t* __restrict__ blurm, float* __restrict__ out) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y;
  if (x >= W || y >= H) return;
  const size_t i = static_cast<size_t>(y) * P + x;
  float val[3] = {nrm[i], nrm[ps + i], nrm[2 * ps + i]};
  if (sharp[i] || blurm[i]) {
    float rs = ts[i], rb = tb[i];
    if (y >= 2 && y + 2 < H) {
      float vs = 0.0f, vb = 0.0f;
#pragma unroll
      for (int j = 0; j < 5; ++j) {
        const size_t k = i + static_cast<size_t>(j) * P - 2 * static_cast<size_t>(P);
        vs += tp.sharpen[j] * ts[k];
        vb += tp.blur[j] * tb[k];
      }
      rs = vs * tp.sharpen_mul;
      rb = vb * tp.blur_mul;
    }
    const float im = val[channel];
    if (sharp[i]) val[channel] = im + (im - rs) * amount;
    else val[channel] = rb;
  }
  out[i] = static_cast<float>(static_cast<double>(val[0]) * 255.0);
  out[ps + i] = (val[1] + 0.5f) * 255.0f;
  out[2 * ps + i] = (val[2] + 0.5f) * 255.0f;
}

// SetDownsampledCoefficients with factor 2 (output_image.cc:496-531): 2x2 float average, double
// DCT, DC - 1024, round half away from zero.
__global__ void __launch_bounds__(256)
k420_downsample_chroma(const float* __restrict__ plane, int W, int H, int P, int mcw, int nblocks,
                       int16_t* __restrict__ coef) {
  __shared__ double s_in[4][64];
  __shared__ double s_tmp[4][64];
  const int lb = threadIdx.x >> 6, t = threadIdx.x & 63;
  const int b = blockIdx.x * 4 + lb;
  const bool live = b < nblocks;
  if (live) {
    const int ix = t & 7, iy = t >> 3;
    const int x0 = 16 * (b % mcw), y0 = 16 * (b / mcw);
    float avg = 0.0f;
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int x = min(x0 + 2 * ix + i, W - 1), y = min(y0 + 2 * iy + j, H - 1);
        avg += plane[static_cast<size_t>(y) * P + x];
      }
    avg /= 4;
    s_in[lb][t] = avg;
  } else {
    s_in[lb][t] = 0.0;
  }
  __syncthreads();
  {
    const int xo = t >> 3, col = t & 7;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += c_dct_matrix[8 * xo + u] * s_in[lb][8 * u + col];
    s_tmp[lb][8 * xo + col] = acc;
  }
  __syncthreads();
  const int row = t >> 3, xo = t & 7;
  double acc = 0.0;
#pragma unroll
  for (int u = 0; u < 8; ++u) acc += c_dct_matrix[8 * xo + u] * s_tmp[lb][8 * row + u];
  if (t == 0) acc -= 1024.0;
  if (live) coef[static_cast<size_t>(b) * 64 + t] = static_cast<int16_t>(round(acc));
}

// SaveToJpegData's MCU padding of the luma component (output_image.cc:608-632): block (bx, by) of
// the cbw x cbh layout is the image's block when it has one, else {DC of the raster predecessor, 0...}
// = the DC of the last image block of row min(by, bh - 1).
__global__ void __launch_bounds__(256)
k420_pad_luma(const int16_t* __restrict__ src, int bw, int bh, int cbw, int cbh, int16_t* __restrict__ dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;   // one thread per 8 coefficients
  if (i >= cbw * cbh * 8) return;
  const int b = i >> 3, part = i & 7;
  const int bx = b % cbw, by = b / cbw;
  int4 v;
  if (bx < bw && by < bh) {
    v = reinterpret_cast<const int4*>(src)[(static_cast<size_t>(by) * bw + bx) * 8 + part];
  } else {
    v = make_int4(0, 0, 0, 0);
    if (part == 0) v.x = static_cast<uint16_t>(src[(static_cast<size_t>(min(by, bh - 1)) * bw + bw - 1) * 64]);
  }
  reinterpret_cast<int4*>(dst)[static_cast<size_t>(b) * 8 + part] = v;
}

// ---------------------------------------------------------------------------------------------
// ComputeBlockZeroingOrder with factor 2 (comp_mask 6): one CTA of four warps per 16x16 macro-block,
// warp w owns the 8x8 sub-block (w & 1, w >> 1). Every warp keeps its own copy of the chroma state
// (coefficients, column pass, sub-sampled samples) so that only the trial's error crosses warps.
// ---------------------------------------------------------------------------------------------
struct ZeroMbSmem {
  ZeroWarpSmem w[4];
  unsigned char ypx[256];       // luma samples of the macro-block (fixed during this pass)
  unsigned char sb[2][100];     // 10x10 neighbourhood of chroma samples; own 8x8 cells unused
  unsigned char win[4][128];    // per warp: upsampled Cb / Cr of its 8x8 window
  float err[4];
  int blk;
};

__device__ __forceinline__ void mb_window(const ZeroMbSmem& m, const unsigned char* own, int cc, int mx, int my,
                                          int ix, int iy, int xmax, int ymax, unsigned char* win, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h;
    const int gx = 16 * mx + 8 * ix + (p & 7), gy = 16 * my + 8 * iy + (p >> 3);
    const int X = min(gx >> 1, xmax), Y = min(gy >> 1, ymax);
    const int Xn = min(max(X + ((gx & 1) ? 1 : -1), 0), xmax);
    const int Yn = min(max(Y + ((gy & 1) ? 1 : -1), 0), ymax);
    auto at = [&](int xa, int ya) -> int {
      const int li = xa - 8 * mx, lj = ya - 8 * my;
      if (li >= 0 && li < 8 && lj >= 0 && lj < 8) return own[8 * lj + li];
      return m.sb[cc][10 * (lj + 1) + li + 1];
    };
    const int v = 9 * at(X, Y) + 3 * at(X, Yn) + 3 * at(Xn, Y) + at(Xn, Yn);
    win[p] = static_cast<unsigned char>((v + 8 - (gx & 1)) >> 4);
  }
}

__global__ void __launch_bounds__(128, 5)
k_zeroing_order_mb(const int16_t* __restrict__ orig, const int16_t* __restrict__ cur, size_t comp_stride,
                   const uint8_t* __restrict__ rgb_planes, const uint8_t* __restrict__ ycc, size_t plane_stride,
                   int P, int W, int H, int bw, int mcw, int nmb, const float* __restrict__ mask_scale, float limit,
                   int lookahead, int mb_begin, CoeffDataDev* __restrict__ out, unsigned int* __restrict__ counter,
                   const int* __restrict__ lpt_order, unsigned int* __restrict__ tie_counter) {
  __shared__ ZeroMbSmem m;
  __shared__ int s_basis[64];
  const float* s_lut = g_tab.srgb_lin;
  if (threadIdx.x < 64) s_basis[threadIdx.x] = kIdctBasis[threadIdx.x];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ZeroWarpSmem& s = m.w[warp];
  const double csf_a = kCsf8x8[4 + lane], csf_b = kCsf8x8[36];
  const int xmax = (W - 1) >> 1, ymax = (H - 1) >> 1;
  const int ix = warp & 1, iy = warp >> 1;
  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) m.blk = static_cast<int>(atomicAdd(counter, 1u)) + mb_begin;
    __syncthreads();
    int mb = m.blk;
    if (mb >= nmb) break;
    if (lpt_order) mb = lpt_order[mb - mb_begin];
    const int mx = mb % mcw, my = mb / mcw;
    const int bxx = 2 * mx + ix, byy = 2 * my + iy;
    const bool live = 8 * bxx < W && 8 * byy < H;
    const int vx = min(8, W - 8 * bxx), vy = min(8, H - 8 * byy);
    // ---- shared, fixed state: luma samples and the chroma neighbourhood ----
    for (int i = threadIdx.x; i < 256; i += 128)
      m.ypx[i] = ycc[static_cast<size_t>(16 * my + (i >> 4)) * P + 16 * mx + (i & 15)];
    for (int i = threadIdx.x; i < 200; i += 128) {
      const int cc = i / 100, e = i - 100 * cc, lj = e / 10 - 1, li = e % 10 - 1;
      const int xa = min(max(8 * mx + li, 0), xmax), ya = min(max(8 * my + lj, 0), ymax);
      m.sb[cc][e] = ycc[(1 + cc) * plane_stride + static_cast<size_t>(ya) * P + xa];
    }
    // ---- per-warp state ----
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = lane + 32 * k, c = i >> 6;
      s.cf[i] = c >= 1 ? cur[c * comp_stride + static_cast<size_t>(mb) * 64 + (i & 63)] : 0;
    }
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        const int x = min(8 * bxx + (p & 7), W - 1), y = min(8 * byy + (p >> 3), H - 1);
        const size_t g = static_cast<size_t>(y) * P + x;
#pragma unroll
        for (int c = 0; c < 3; ++c) s.bufA[64 * c + p] = s_lut[rgb_planes[c * plane_stride + g]];
      }
    }
    __syncwarp();
    if (live) warp_block_opsin(s.bufA, s.bufB, s.pg0, lane);
    warp_full_idct(s, s_basis, 1, lane);
    warp_full_idct(s, s_basis, 2, lane);
    float scale[3] = {0.f, 0.f, 0.f};
    if (live) {
      const int blk = byy * bw + bxx;
      scale[0] = mask_scale[3 * blk]; scale[1] = mask_scale[3 * blk + 1]; scale[2] = mask_scale[3 * blk + 2];
    }
    __syncthreads();   // ypx / sb visible
    if (live) {
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        s.pix[p] = m.ypx[16 * (8 * iy + (p >> 3)) + 8 * ix + (p & 7)];
      }
    }
    // ---- input order over the Cb / Cr coefficients ----
    int n = 0;
#pragma unroll 1
    for (int k = 2; k < 6; ++k) {
      const int i = lane + 32 * k, c = i >> 6;
      const bool take = (i & 63) != 0 && s.cf[i] != 0;
      const unsigned int mk = __ballot_sync(0xffffffffu, take);
      if (take) {
        const int pos = n + __popc(mk & ((1u << lane) - 1));
        const int o = orig[c * comp_stride + static_cast<size_t>(mb) * 64 + (i & 63)];
        s.ent[pos] = static_cast<unsigned char>(i);
        s.key[pos] = abs(o) * c_zero_csf[i] + c_zero_bias[i];
      }
      n += __popc(mk);
    }
    __syncwarp();
    // (every warp of the macro-block orders its private copy; only the first one reports a tie)
    warp_input_order(s.key, s.ent, n, s.order, reinterpret_cast<ZeroSortPair*>(s.bufA), (threadIdx.x >> 5) == 0 ? tie_counter : nullptr, lane);
    CoeffDataDev* o = out + static_cast<size_t>(mb) * 192;
    int win[3];
    int nwin = min(lookahead, n), next = nwin, nout = 0;
    for (int i = 0; i < 3; ++i) win[i] = i < nwin ? s.order[i] : 0;
    while (nwin > 0) {
      float best_err = 1e17f;
      int best_i = 0;
      for (int i = 0; i < nwin; ++i) {
        const int zidx = win[i], zc = zidx >> 6;
        float err = 0.0f;
        if (live) {
          warp_candidate_idct(s, s_basis, zidx, lane);
          mb_window(m, zc == 1 ? s.cpx : s.pix + 64, 0, mx, my, ix, iy, xmax, ymax, m.win[warp], lane);
          mb_window(m, zc == 2 ? s.cpx : s.pix + 128, 1, mx, my, ix, iy, xmax, ymax, m.win[warp] + 64, lane);
          __syncwarp();
          err = warp_compare_pixels(s, s_lut, s.pix, m.win[warp], m.win[warp] + 64, vx, vy, scale, csf_a, csf_b, lane);
        }
        if (lane == 0) m.err[warp] = err;
        __syncthreads();
        const float max_err = fmaxf(fmaxf(0.0f, m.err[0]), fmaxf(m.err[1], fmaxf(m.err[2], m.err[3])));
        __syncthreads();
        if (max_err < best_err) { best_err = max_err; best_i = i; }
      }
      const int idx = win[best_i];
      warp_commit_zero(s, s_basis, idx, lane);
      if (threadIdx.x == 0) { o[nout].idx = idx; o[nout].block_err = best_err; }
      ++nout;
      for (int i = best_i; i + 1 < nwin; ++i) win[i] = win[i + 1];
      if (next < n) win[nwin - 1] = s.order[next++];
      else --nwin;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float min_err = 1e10f;
      for (int i = nout - 1; i >= 0; --i) {
        min_err = fminf(min_err, o[i].block_err);
        o[i].block_err = min_err;
      }
      int keep = 0;
      while (keep < nout && o[keep].block_err <= limit) ++keep;
      for (int i = keep; i < nout; ++i) { o[i].idx = 0; o[i].block_err = 0.0f; }
    }
  }
}

}  // namespace gzb
