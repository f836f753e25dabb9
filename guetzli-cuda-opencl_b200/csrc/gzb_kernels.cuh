// gzb_kernels.cuh -- sm_100a kernels of the full-image butteraugli Compare pipeline.
//
// Data layout in HBM (all resident per context, see gzb_context.cuh):
//   * full-resolution float planes: row pitch P = round_up(W, 32) floats (128-byte rows)
//   * candidate / original sRGB8 planes: planar u8, same pitch P, height padded to 8*ceil(H/8)
//   * coefficients: int16 [3][B8][64], block-major (OutputImageComponent::coeffs layout)
//   * res-grid maps (step 3): [ceil(H/3)][ceil(W/3)] with 3 interleaved floats where noted
//   * decimated blur outputs: [ny][pitch round_up(nx,32)]
// Grids are sized from the image; the heavy warp-per-cell kernels run persistent CTAs in
// multiples of the SM count.
#pragma once
#include <cuda.h>   // CUtensorMap (the encoder entry point is resolved at run time: no libcuda link dependency)
#include "gzb_device_math.cuh"

namespace gzb {

constexpr int kMaxTaps = 65;           // radius <= 32
__constant__ float c_taps[12][kMaxTaps];  // one row per blur kernel (see BlurKind)

__host__ __device__ inline int round_up(int v, int m) { return (v + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------
// Incremental Compare: between two Compares of the back end only a few blocks change, and every
// stage has bounded support, so each stage owns a byte mask over 32x32-pixel tiles of the image
// ("an output anchored in this tile may change") and skips clean outputs; their values are still in
// the context's buffers from the previous Compare. m == nullptr: everything is dirty.
// ---------------------------------------------------------------------------------------------
struct DirtyMask { const uint8_t* m; int tw, th; };
__device__ __forceinline__ bool dirty_at(const DirtyMask& d, int x, int y) {
  if (!d.m) return true;
  const int tx = min(max(x, 0) >> 5, d.tw - 1), ty = min(max(y, 0) >> 5, d.th - 1);
  return d.m[ty * d.tw + tx] != 0;
}
// any tile intersecting the pixel rectangle [x0, x1] x [y0, y1] dirty? (uniform per CTA when the arguments are)
__device__ __forceinline__ bool dirty_any(const DirtyMask& d, int x0, int y0, int x1, int y1) {
  if (!d.m) return true;
  const int tx0 = min(max(x0, 0) >> 5, d.tw - 1), ty0 = min(max(y0, 0) >> 5, d.th - 1);
  const int tx1 = min(max(x1, 0) >> 5, d.tw - 1), ty1 = min(max(y1, 0) >> 5, d.th - 1);
  for (int ty = ty0; ty <= ty1; ++ty)
    for (int tx = tx0; tx <= tx1; ++tx)
      if (d.m[ty * d.tw + tx]) return true;
  return false;
}

// ---------------------------------------------------------------------------------------------
// K0: interleaved sRGB8 -> planar u8 with pitch (upload helper)
// ---------------------------------------------------------------------------------------------
__global__ void k_deinterleave_rgb(const uint8_t* __restrict__ rgb, int W, int H, int P,
                                   uint8_t* __restrict__ planes, size_t plane_stride) {
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= W || y >= H) return;
  const uint8_t* s = rgb + 3 * (static_cast<size_t>(y) * W + x);
  const size_t o = static_cast<size_t>(y) * P + x;
  planes[o] = s[0];
  planes[plane_stride + o] = s[1];
  planes[2 * plane_stride + o] = s[2];
}

// Sparse coefficient update (SetCoeffBlock on touched blocks; guetzli/processor.cc:867-874).
// A coefficient is touched at most once per batch in the back-end loop except when an "up" and a
// later record hit the same slot; the highest record index must win, so only the last record of a
// slot writes.
__global__ void k_scatter_coeffs(const int* __restrict__ block_ix, const int16_t* __restrict__ val,
                                 const uint8_t* __restrict__ idx, size_t n, size_t comp_stride,
                                 int16_t* __restrict__ coef, uint8_t* __restrict__ blk_changed) {
  const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
  if (i >= n) return;
  const int id = idx[i];
  coef[(id >> 6) * comp_stride + static_cast<size_t>(block_ix[i]) * 64 + (id & 63)] = val[i];
  if (blk_changed) blk_changed[block_ix[i]] = 1;   // 4:4:4: the coefficient block is the image block (see BlockChanges)
}

// ---------------------------------------------------------------------------------------------
// K0b: the RGB front end on the device: EncodeRGBToJpeg with the all-ones quantiser
//   (guetzli/jpeg_data_encoder.cc:28-117): RGB -> YCbCr in 16.16 fixed point, the integer forward DCT
//   (guetzli/fdct.cc:28-240: column pass then row pass, truncating shifts, int16 stores) and the q = 1
//   "quantisation" ((v * 65537 + (0x80 << 12)) >> 20). Window columns / rows past the image replicate
//   the last one. CTA = 256 threads = 32 blocks x 8 threads; thread t owns row t, column t, row t.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int fd_mul16(int a, int b) { return (a * b) >> 16; }
__device__ __forceinline__ int fd_s16(int v) { return static_cast<int>(static_cast<int16_t>(v)); }

__device__ __forceinline__ void fdct_column8(int v[8]) {
  const int i0 = v[0], i1 = v[1], i2 = v[2], i3 = v[3], i4 = v[4], i5 = v[5], i6 = v[6], i7 = v[7];
  int d07 = i0 - i7, s07 = i0 + i7, d25 = i2 - i5, s25 = i2 + i5;
  int d34 = i3 - i4, s34 = i3 + i4, d16 = i1 - i6, s16 = i1 + i6;
  int e0 = s07 - s34, e1 = s07 + s34, e2 = s16 - s25, e3 = s16 + s25;
  e1 <<= 3; e3 <<= 3;
  v[0] = fd_s16(e1 + e3);
  v[4] = fd_s16(e1 - e3);
  e0 <<= 3; e2 <<= 3; d34 <<= 3; d07 <<= 3;
  v[2] = fd_s16(fd_mul16(27146, e2) + e0);
  v[6] = fd_s16(fd_mul16(27146, e0) - e2);
  d25 <<= 4; d16 <<= 4;
  const int r = fd_mul16(d16 + d25, 23170), s = fd_mul16(d16 - d25, 23170);
  const int p3 = d34 - s, p1 = d34 + s, p0 = d07 - r, p2 = d07 + r;
  const int q3 = fd_mul16(p3, -21746) + p3 + 1;
  const int q1 = fd_mul16(p1, 13036) + p2 + 1;
  const int q4 = fd_mul16(-21746, p0) + p0;
  const int q5 = fd_mul16(13036, p2);
  v[1] = fd_s16(q1);
  v[3] = fd_s16(p0 - q3);
  v[5] = fd_s16(p3 + q4);
  v[7] = fd_s16(q5 - p1);
}

__constant__ const int kFdctRowTab[4][7] = {{22725, 21407, 19266, 16384, 12873, 8867, 4520},
                                            {31521, 29692, 26722, 22725, 17855, 12299, 6270},
                                            {29692, 27969, 25172, 21407, 16819, 11585, 5906},
                                            {26722, 25172, 22654, 19266, 15137, 10426, 5315}};

__device__ __forceinline__ void fdct_row8(int in[8], int row) {
  const int sel = row == 0 || row == 4 ? 0 : (row == 1 || row == 7) ? 1 : (row == 2 || row == 6) ? 2 : 3;
  const int* t = kFdctRowTab[sel];
  const int a0 = in[0] + in[7], b0 = in[0] - in[7], a1 = in[1] + in[6], b1 = in[1] - in[6];
  const int a2 = in[2] + in[5], b2 = in[2] - in[5], a3 = in[3] + in[4], b3 = in[3] - in[4];
  const int C1 = t[0], C2 = t[1], C3 = t[2], C4 = t[3], C5 = t[4], C6 = t[5], C7 = t[6];
  const int c0 = a0 + a3, c1 = a0 - a3, c2 = a1 + a2, c3 = a1 - a2;
  in[0] = fd_s16((C4 * (c0 + c2)) >> 16);
  in[4] = fd_s16((C4 * (c0 - c2)) >> 16);
  in[2] = fd_s16((C2 * c1 + C6 * c3) >> 16);
  in[6] = fd_s16((C6 * c1 - C2 * c3) >> 16);
  in[1] = fd_s16((C1 * b0 + C3 * b1 + C5 * b2 + C7 * b3) >> 16);
  in[3] = fd_s16((C3 * b0 - C7 * b1 - C1 * b2 - C5 * b3) >> 16);
  in[5] = fd_s16((C5 * b0 - C1 * b1 + C7 * b2 + C3 * b3) >> 16);
  in[7] = fd_s16((C7 * b0 - C5 * b1 + C3 * b2 - C1 * b3) >> 16);
}

__global__ void __launch_bounds__(256)
k_rgb_to_coeffs(const uint8_t* __restrict__ planes, size_t plane_stride, int W, int H, int P, int bw, int nblocks,
                int16_t* __restrict__ out, size_t comp_stride) {
  __shared__ int s[3][32][8][9];
  const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + lb;
  const bool live = b < nblocks;
  const int bx = live ? b % bw : 0, by = live ? b / bw : 0;
  {  // row t of the window: RGB -> YCbCr (int16 like the reference's block arrays)
    const int y = min(H - 1, 8 * by + t);
#pragma unroll
    for (int ix = 0; ix < 8; ++ix) {
      const int x = min(W - 1, 8 * bx + ix);
      const size_t g = static_cast<size_t>(y) * P + x;
      const int r = planes[g], gg = planes[plane_stride + g], bl = planes[2 * plane_stride + g];
      s[0][lb][t][ix] = fd_s16((19595 * r + 38469 * gg + 7471 * bl - (128 << 16) + 32768) >> 16);
      s[1][lb][t][ix] = fd_s16((-11059 * r - 21709 * gg + 32768 * bl + 32768 - 1) >> 16);
      s[2][lb][t][ix] = fd_s16((32768 * r - 27439 * gg - 5329 * bl + 32768 - 1) >> 16);
    }
  }
  __syncthreads();
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    int v[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = s[c][lb][k][t];   // column t
    fdct_column8(v);
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) s[c][lb][k][t] = v[k];
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 8; ++k) v[k] = s[c][lb][t][k];   // row t
    fdct_row8(v, t);
    if (live) {
      int e[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) e[k] = fd_s16((v[k] * 65537 + (0x80 << 12)) >> 20);
      int4 o;
      o.x = (e[0] & 0xffff) | (e[1] << 16);
      o.y = (e[2] & 0xffff) | (e[3] << 16);
      o.z = (e[4] & 0xffff) | (e[5] << 16);
      o.w = (e[6] & 0xffff) | (e[7] << 16);
      reinterpret_cast<int4*>(out + c * comp_stride)[static_cast<size_t>(b) * 8 + t] = o;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// K1: coefficients -> candidate sRGB8 planes (integer IDCT + YCbCr->RGB).
//   guetzli/output_image.cc:124-146 (SetCoeffBlock), 68-98 (ToPixels), 642-652 (ToSRGB);
//   in 4:4:4 the stored pixel is idct<<4 and (p+8-(x&1))>>4 returns the idct value.
// CTA = 256 threads = 32 blocks x 8 threads; thread t owns column t, then row t.
// Optionally applies a global quantiser first (ApplyGlobalQuantization, output_image.cc:349-360)
// or the coeff*quant copy (CopyFromJpegComponent, 212-228) and writes the coefficients back.
// ---------------------------------------------------------------------------------------------
// kCoeffQuantizeSrc: quantise `src` (the q = 1 input) into `dst`: CopyFromJpegData with the all-ones matrix
// followed by ApplyGlobalQuantization, in one pass.
enum CoeffOp { kCoeffKeep = 0, kCoeffQuantize = 1, kCoeffScale = 2, kCoeffQuantizeSrc = 3 };

template <int OP>
__global__ void __launch_bounds__(256)
k_coeffs_to_rgb8(const int16_t* __restrict__ src, int16_t* __restrict__ dst, size_t comp_stride,
                 const int* __restrict__ q192, int bw, int nblocks, int P,
                 uint8_t* __restrict__ planes, size_t plane_stride) {
  __shared__ int s_in[32][8][9];
  __shared__ uint8_t s_px[3][32][8][8];
  const int lb = threadIdx.x >> 3, t = threadIdx.x & 7;
  const int b = blockIdx.x * 32 + lb;
  const bool live = b < nblocks;
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    if (live) {
      const int4 v = reinterpret_cast<const int4*>(src + c * comp_stride)[static_cast<size_t>(b) * 8 + t];
      int e[8] = {static_cast<int16_t>(v.x & 0xffff), v.x >> 16, static_cast<int16_t>(v.y & 0xffff), v.y >> 16,
                  static_cast<int16_t>(v.z & 0xffff), v.z >> 16, static_cast<int16_t>(v.w & 0xffff), v.w >> 16};
      if (OP != kCoeffKeep) {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const int q = q192[64 * c + 8 * t + k];
          e[k] = (OP == kCoeffQuantize || OP == kCoeffQuantizeSrc) ? quantize_coeff(e[k], q)
                                                                   : static_cast<int>(static_cast<int16_t>(e[k] * q));
        }
        int4 o;
        o.x = (e[0] & 0xffff) | (e[1] << 16);
        o.y = (e[2] & 0xffff) | (e[3] << 16);
        o.z = (e[4] & 0xffff) | (e[5] << 16);
        o.w = (e[6] & 0xffff) | (e[7] << 16);
        reinterpret_cast<int4*>(dst + c * comp_stride)[static_cast<size_t>(b) * 8 + t] = o;
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) s_in[lb][t][k] = e[k];  // row t of the coefficient block
    }
    __syncthreads();
    int col[8], out[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) col[u] = s_in[lb][u][t];  // column t
    __syncthreads();
    idct_1d(col, out);
#pragma unroll
    for (int y = 0; y < 8; ++y) s_in[lb][y][t] = idct_col_round(out[y]);
    __syncthreads();
    int row[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) row[u] = s_in[lb][t][u];
    idct_1d(row, out);
#pragma unroll
    for (int x = 0; x < 8; ++x) s_px[c][lb][t][x] = static_cast<uint8_t>(idct_row_round(out[x]));
    __syncthreads();
  }
  if (!live) return;
  const int bx = b % bw, by = b / bw;
  uint32_t pr[2] = {0, 0}, pg[2] = {0, 0}, pb[2] = {0, 0};
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    int r, g, bl;
    ycbcr_to_rgb(s_px[0][lb][t][x], s_px[1][lb][t][x], s_px[2][lb][t][x], r, g, bl);
    pr[x >> 2] |= static_cast<uint32_t>(r) << (8 * (x & 3));
    pg[x >> 2] |= static_cast<uint32_t>(g) << (8 * (x & 3));
    pb[x >> 2] |= static_cast<uint32_t>(bl) << (8 * (x & 3));
  }
  const size_t o = static_cast<size_t>(8 * by + t) * P + 8 * bx;
  *reinterpret_cast<uint2*>(planes + o) = make_uint2(pr[0], pr[1]);
  *reinterpret_cast<uint2*>(planes + plane_stride + o) = make_uint2(pg[0], pg[1]);
  *reinterpret_cast<uint2*>(planes + 2 * plane_stride + o) = make_uint2(pb[0], pb[1]);
}

// ---------------------------------------------------------------------------------------------
// K2: sRGB8 planes -> XYB planes: LUT -> blur sigma 1.1 (5 taps, separable, border-renormalised)
//     -> opsin dynamics. One 32x32 output tile per CTA, halo 2, all three channels in shared
//     memory. butteraugli.cc:943-974 with Blur (100-148) inlined; scale tables from the host.
// ---------------------------------------------------------------------------------------------
constexpr int kOpsTile = 32;
__global__ void __launch_bounds__(256)
k_opsin_dynamics(const uint8_t* __restrict__ planes, size_t plane_stride, int W, int H, int P,
                 const double* __restrict__ scale_x, const double* __restrict__ scale_y,
                 float* __restrict__ xyb, size_t xyb_stride, DirtyMask dm) {
  __shared__ float s_lin[3][kOpsTile + 4][kOpsTile + 4 + 1];
  __shared__ float s_h[3][kOpsTile + 4][kOpsTile + 1];
  __shared__ float s_lut[256];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * kOpsTile, y0 = blockIdx.y * kOpsTile;
  if (!dirty_at(dm, x0, y0)) return;   // the CTA's tile is the mask's tile
  s_lut[tid] = g_tab.srgb_lin[tid];
  __syncthreads();
  for (int i = tid; i < 3 * 36 * 36; i += 256) {
    const int c = i / (36 * 36), rem = i - c * 36 * 36, ly = rem / 36, lx = rem - ly * 36;
    const int gx = x0 + lx - 2, gy = y0 + ly - 2;
    float v = 0.0f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H)
      v = s_lut[planes[c * plane_stride + static_cast<size_t>(gy) * P + gx]];
    s_lin[c][ly][lx] = v;
  }
  __syncthreads();
  const float t0 = c_taps[0][0], t1 = c_taps[0][1], t2 = c_taps[0][2], t3 = c_taps[0][3], t4 = c_taps[0][4];
  // horizontal pass for 36 rows x 32 columns x 3 channels
  for (int i = tid; i < 3 * 36 * 32; i += 256) {
    const int c = i / (36 * 32), rem = i - c * 36 * 32, ly = rem >> 5, lx = rem & 31;
    const int gx = x0 + lx;
    float r = 0.0f;
    if (gx < W) {
      const float* p = &s_lin[c][ly][lx];
      double acc = 0.0;
      acc += static_cast<double>(p[0] * t0);
      acc += static_cast<double>(p[1] * t1);
      acc += static_cast<double>(p[2] * t2);
      acc += static_cast<double>(p[3] * t3);
      acc += static_cast<double>(p[4] * t4);
      r = static_cast<float>(acc * scale_x[gx]);
    }
    s_h[c][ly][lx] = r;
  }
  __syncthreads();
  const int lx = threadIdx.x, gx = x0 + lx;
  if (gx >= W) return;
#pragma unroll 1
  for (int ly = threadIdx.y; ly < kOpsTile; ly += 8) {
    const int gy = y0 + ly;
    if (gy >= H) break;
    const double sy = scale_y[gy];
    float bl[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
      // rows outside the image hold zeros; rows of the H pass outside the image are skipped by
      // the reference -- they contribute +0.0 here.
      const bool in0 = gy - 2 >= 0, in1 = gy - 1 >= 0, in3 = gy + 1 < H, in4 = gy + 2 < H;
      acc += static_cast<double>((in0 ? s_h[c][ly][lx] : 0.0f) * t0);
      acc += static_cast<double>((in1 ? s_h[c][ly + 1][lx] : 0.0f) * t1);
      acc += static_cast<double>(s_h[c][ly + 2][lx] * t2);
      acc += static_cast<double>((in3 ? s_h[c][ly + 3][lx] : 0.0f) * t3);
      acc += static_cast<double>((in4 ? s_h[c][ly + 4][lx] : 0.0f) * t4);
      bl[c] = static_cast<float>(acc * sy);
    }
    float X, Y, B;
    opsin_pixel(bl[0], bl[1], bl[2], s_lin[0][ly + 2][lx + 2], s_lin[1][ly + 2][lx + 2],
                s_lin[2][ly + 2][lx + 2], X, Y, B);
    const size_t o = static_cast<size_t>(gy) * P + gx;
    xyb[o] = X;
    xyb[xyb_stride + o] = Y;
    xyb[2 * xyb_stride + o] = B;
  }
}

// Float-plane variant used by the reference-shaped stage entry point (gzb_opsin_dynamics_image):
// input linear RGB float planes instead of sRGB8.
__global__ void __launch_bounds__(256)
k_opsin_dynamics_f32(const float* __restrict__ lin, size_t lin_stride, int W, int H, int P,
                     const double* __restrict__ scale_x, const double* __restrict__ scale_y,
                     float* __restrict__ xyb, size_t xyb_stride) {
  __shared__ float s_lin[3][kOpsTile + 4][kOpsTile + 4 + 1];
  __shared__ float s_h[3][kOpsTile + 4][kOpsTile + 1];
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * kOpsTile, y0 = blockIdx.y * kOpsTile;
  for (int i = tid; i < 3 * 36 * 36; i += 256) {
    const int c = i / (36 * 36), rem = i - c * 36 * 36, ly = rem / 36, lx = rem - ly * 36;
    const int gx = x0 + lx - 2, gy = y0 + ly - 2;
    float v = 0.0f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = lin[c * lin_stride + static_cast<size_t>(gy) * P + gx];
    s_lin[c][ly][lx] = v;
  }
  __syncthreads();
  const float t0 = c_taps[0][0], t1 = c_taps[0][1], t2 = c_taps[0][2], t3 = c_taps[0][3], t4 = c_taps[0][4];
  for (int i = tid; i < 3 * 36 * 32; i += 256) {
    const int c = i / (36 * 32), rem = i - c * 36 * 32, ly = rem >> 5, lx = rem & 31;
    const int gx = x0 + lx;
    float r = 0.0f;
    if (gx < W) {
      const float* p = &s_lin[c][ly][lx];
      double acc = 0.0;
      acc += static_cast<double>(p[0] * t0);
      acc += static_cast<double>(p[1] * t1);
      acc += static_cast<double>(p[2] * t2);
      acc += static_cast<double>(p[3] * t3);
      acc += static_cast<double>(p[4] * t4);
      r = static_cast<float>(acc * scale_x[gx]);
    }
    s_h[c][ly][lx] = r;
  }
  __syncthreads();
  const int lx = threadIdx.x, gx = x0 + lx;
  if (gx >= W) return;
#pragma unroll 1
  for (int ly = threadIdx.y; ly < kOpsTile; ly += 8) {
    const int gy = y0 + ly;
    if (gy >= H) break;
    const double sy = scale_y[gy];
    float bl[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
      const bool in0 = gy - 2 >= 0, in1 = gy - 1 >= 0, in3 = gy + 1 < H, in4 = gy + 2 < H;
      acc += static_cast<double>((in0 ? s_h[c][ly][lx] : 0.0f) * t0);
      acc += static_cast<double>((in1 ? s_h[c][ly + 1][lx] : 0.0f) * t1);
      acc += static_cast<double>(s_h[c][ly + 2][lx] * t2);
      acc += static_cast<double>((in3 ? s_h[c][ly + 3][lx] : 0.0f) * t3);
      acc += static_cast<double>((in4 ? s_h[c][ly + 4][lx] : 0.0f) * t4);
      bl[c] = static_cast<float>(acc * sy);
    }
    float X, Y, B;
    opsin_pixel(bl[0], bl[1], bl[2], s_lin[0][ly + 2][lx + 2], s_lin[1][ly + 2][lx + 2],
                s_lin[2][ly + 2][lx + 2], X, Y, B);
    const size_t o = static_cast<size_t>(gy) * P + gx;
    xyb[o] = X;
    xyb[xyb_stride + o] = Y;
    xyb[2 * xyb_stride + o] = B;
  }
}

// ---------------------------------------------------------------------------------------------
// K3: MaskHighIntensityChange (butteraugli.cc:791-843). One pixel per thread; neighbours via L1.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_mask_high_intensity_change(const float* __restrict__ a, const float* __restrict__ b, size_t stride,
                             int W, int H, int P, float* __restrict__ oa, float* __restrict__ ob, DirtyMask dm) {
  const int x = blockIdx.x * 32 + threadIdx.x;
  const int y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  if (!dirty_at(dm, x, y)) return;
  const size_t ix = static_cast<size_t>(y) * P + x;
  const float c0[3] = {a[ix], a[stride + ix], a[2 * stride + ix]};
  const float c1[3] = {b[ix], b[stride + ix], b[2 * stride + ix]};
  const float* ya = a + stride;
  const float* yb = b + stride;
  double worst = -1;
  if (x > 0) { const double d = mhic_sqdiff(ya[ix - 1], yb[ix - 1], c0[1], c1[1]); if (worst < d) worst = d; }
  if (x + 1 < W) { const double d = mhic_sqdiff(ya[ix + 1], yb[ix + 1], c0[1], c1[1]); if (worst < d) worst = d; }
  if (y > 0) { const double d = mhic_sqdiff(ya[ix - P], yb[ix - P], c0[1], c1[1]); if (worst < d) worst = d; }
  if (y + 1 < H) { const double d = mhic_sqdiff(ya[ix + P], yb[ix + P], c0[1], c1[1]); if (worst < d) worst = d; }
  float o0[3], o1[3];
  mhic_pixel(c0, c1, worst, o0, o1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    oa[c * stride + ix] = o0[c];
    ob[c * stride + ix] = o1[c];
  }
}

// ---------------------------------------------------------------------------------------------
// K4: separable Gaussian on an output lattice (butteraugli.cc:68-148).
//   H pass: tmp[y][ox]  = float( sum_j double(in[y][j] * tap[j-x+r]) * scale_x[ox] ), x = x0+ox*sx
//   V pass: out[oy][ox] = float( sum_j double(tmp[j][ox] * tap[j-y+r]) * scale_y[oy] ), y = y0+oy*sy
// Taps outside the image read 0.0f (adds +0.0: identical to the reference skipping them); the
// border renormalisation lives in the host-built scale tables. UPS>1 reads a nearest-lower
// replicated input (in[(y/UPS)][x/UPS]) -- used for the 3x3-upsampled diffmap.
// Shared-memory tiles with halos; rows are read with coalesced 128-byte lines.
// ---------------------------------------------------------------------------------------------
struct BlurGeom {
  int in_w, in_h;    // logical input size (after UPS replication)
  int in_pitch;      // physical input pitch (floats)
  int r, kind;       // radius, row of c_taps
  int x0, sx, nx;    // output columns
  int y0, sy, ny;    // output rows
  int tmp_pitch;     // pitch of tmp and out
  int ups;
  int oxn, oyn;      // outputs per CTA along x (H pass) / y (V pass): tile must fit shared memory
};

constexpr int kBhOx = 32, kBhRows = 16, kBhMaxSpan = kBhOx * 4 + 2 * 32;  // 192
template <int UPS>
__global__ void __launch_bounds__(256)
k_blur_h(const float* __restrict__ in, size_t in_stride, BlurGeom g,
         const double* __restrict__ scale_x, float* __restrict__ tmp, size_t tmp_stride, DirtyMask dm) {
  __shared__ float s[kBhRows][kBhMaxSpan + 1];
  in += blockIdx.z * in_stride;
  tmp += blockIdx.z * tmp_stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int ox0 = blockIdx.x * g.oxn, yb = blockIdx.y * kBhRows;
  // outputs are anchored at pixel (x0 + ox * sx, row)
  if (!dirty_any(dm, g.x0 + ox0 * g.sx, yb, g.x0 + (ox0 + g.oxn - 1) * g.sx, yb + kBhRows - 1)) return;
  const int xs = g.x0 + ox0 * g.sx - g.r;  // first input column of the tile
  const int span = (g.oxn - 1) * g.sx + 2 * g.r + 1;
  for (int i = tid; i < kBhRows * span; i += 256) {
    const int ly = i / span, k = i - ly * span;
    const int gx = xs + k, gy = yb + ly;
    float v = 0.0f;
    if (gx >= 0 && gx < g.in_w && gy < g.in_h)
      v = UPS == 1 ? in[static_cast<size_t>(gy) * g.in_pitch + gx]
                   : in[static_cast<size_t>(gy / UPS) * g.in_pitch + gx / UPS];
    s[ly][k] = v;
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x;
  if (ox >= g.nx || threadIdx.x >= g.oxn) return;
  const double sc = scale_x[ox];
  const float* taps = c_taps[g.kind];
  const int nt = 2 * g.r + 1;
#pragma unroll
  for (int rr = 0; rr < kBhRows; rr += 8) {
    const int ly = threadIdx.y + rr, gy = yb + ly;
    if (gy >= g.in_h) break;
    if (!dirty_at(dm, g.x0 + ox * g.sx, gy)) continue;
    const float* p = &s[ly][threadIdx.x * g.sx];
    double acc = 0.0;
    for (int k = 0; k < nt; ++k) acc += static_cast<double>(p[k] * taps[k]);
    tmp[static_cast<size_t>(gy) * g.tmp_pitch + ox] = static_cast<float>(acc * sc);
  }
}

// The same H pass with its halo tile brought in by the TMA unit: one elected thread arms an mbarrier with
// the tile's byte count and issues cp.async.bulk.tensor (3-D map: x, y, plane) for the box
// [xa, xa + box_w) x [yb, yb + 16) of plane blockIdx.z; coordinates left of / beyond the plane are zero
// filled by the hardware -- exactly the "taps outside the image read 0.0f" rule of the scalar loader.
// The innermost coordinate of a non-swizzled tiled load must be a multiple of 16 bytes (measured: an odd
// start faults with "illegal instruction", tests/probes/tma_probe.cu), so the box starts at xs rounded
// down to 4 floats and is up to 3 floats wider; the taps read from the offset xs - xa.
constexpr int kBhTmaMaxBox = 196;
__global__ void __launch_bounds__(256)
k_blur_h_tma(const __grid_constant__ CUtensorMap tmap, BlurGeom g, int box_w,
             const double* __restrict__ scale_x, float* __restrict__ tmp, size_t tmp_stride, DirtyMask dm) {
  __shared__ __align__(128) float s[kBhRows * kBhTmaMaxBox];
  __shared__ __align__(8) unsigned long long mbar;
  tmp += blockIdx.z * tmp_stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int ox0 = blockIdx.x * g.oxn, yb = blockIdx.y * kBhRows;
  if (!dirty_any(dm, g.x0 + ox0 * g.sx, yb, g.x0 + (ox0 + g.oxn - 1) * g.sx, yb + kBhRows - 1)) return;
  const int xs = g.x0 + ox0 * g.sx - g.r;  // first input column of the tile (may be negative)
  const int xa = xs & ~3;                  // ... rounded down to a 16-byte boundary (two's complement: also for xs < 0)
  const unsigned mbar_addr = static_cast<unsigned>(__cvta_generic_to_shared(&mbar));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the initialised barrier must be visible to the async proxy
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned bytes = static_cast<unsigned>(box_w) * kBhRows * sizeof(float);
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(mbar_addr), "r"(xa), "r"(yb),
          "r"(static_cast<int>(blockIdx.z))
        : "memory");
  }
  {  // every thread waits for the bytes to land (phase 0 of the barrier)
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(mbar_addr) : "memory");
    }
  }
  const int ox = ox0 + threadIdx.x;
  if (ox >= g.nx || threadIdx.x >= g.oxn) return;
  const double sc = scale_x[ox];
  const float* taps = c_taps[g.kind];
  const int nt = 2 * g.r + 1;
#pragma unroll
  for (int rr = 0; rr < kBhRows; rr += 8) {
    const int ly = threadIdx.y + rr, gy = yb + ly;
    if (gy >= g.in_h) break;
    if (!dirty_at(dm, g.x0 + ox * g.sx, gy)) continue;
    double acc = 0.0;
    if (g.sx == 4) {
      // Outputs four pixels apart: scalar reads would put the 32 lanes on 8 banks (4-way conflict, the
      // shared pipe was the limiter of the sigma-14 passes). Each lane instead streams 16-byte chunks from
      // its own 16-byte aligned position (box_w and 4 * lane are multiples of four floats): a warp reads
      // 512 contiguous bytes per instruction. Stream element j is tap j - off; same order of additions.
      const float4* q = reinterpret_cast<const float4*>(&s[ly * box_w + 4 * threadIdx.x]);
      const int off = xs - xa;   // 0..3, the same for the whole CTA
      int k = 0, m = 0;
      if (off) {
        const float4 v = q[m++];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 1; t < 4; ++t)
          if (t >= off) acc += static_cast<double>(e[t] * taps[k++]);
      }
      for (; k + 4 <= nt; k += 4) {
        const float4 v = q[m++];
        acc += static_cast<double>(v.x * taps[k]);
        acc += static_cast<double>(v.y * taps[k + 1]);
        acc += static_cast<double>(v.z * taps[k + 2]);
        acc += static_cast<double>(v.w * taps[k + 3]);
      }
      if (k < nt) {
        const float4 v = q[m];
        const float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int t = 0; t < 3; ++t)
          if (k + t < nt) acc += static_cast<double>(e[t] * taps[k + t]);
      }
    } else {
      const float* p = &s[ly * box_w + (xs - xa) + threadIdx.x * g.sx];
      for (int k = 0; k < nt; ++k) acc += static_cast<double>(p[k] * taps[k]);
    }
    tmp[static_cast<size_t>(gy) * g.tmp_pitch + ox] = static_cast<float>(acc * sc);
  }
}

constexpr int kBvOy = 16, kBvMaxRows = (kBvOy - 1) * 4 + 2 * 32 + 1;  // 125
__global__ void __launch_bounds__(256)
k_blur_v(const float* __restrict__ tmp, size_t tmp_stride, BlurGeom g,
         const double* __restrict__ scale_y, float* __restrict__ out, size_t out_stride,
         int out_pitch, DirtyMask dm) {
  __shared__ float s[kBvMaxRows][33];
  tmp += blockIdx.z * tmp_stride;
  out += blockIdx.z * out_stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * g.oyn;
  // outputs are anchored at pixel (x0 + ox * sx, y0 + oy * sy)
  if (!dirty_any(dm, g.x0 + ox0 * g.sx, g.y0 + oy0 * g.sy, g.x0 + (ox0 + 31) * g.sx, g.y0 + (oy0 + g.oyn - 1) * g.sy)) return;
  const int ys = g.y0 + oy0 * g.sy - g.r;
  const int rows = (g.oyn - 1) * g.sy + 2 * g.r + 1;
  for (int i = tid; i < rows * 32; i += 256) {
    const int ly = i >> 5, lx = i & 31;
    const int gy = ys + ly, gx = ox0 + lx;
    float v = 0.0f;
    if (gy >= 0 && gy < g.in_h && gx < g.nx) v = tmp[static_cast<size_t>(gy) * g.tmp_pitch + gx];
    s[ly][lx] = v;
  }
  __syncthreads();
  const int ox = ox0 + threadIdx.x;
  if (ox >= g.nx) return;
  const float* taps = c_taps[g.kind];
  const int nt = 2 * g.r + 1;
#pragma unroll
  for (int rr = 0; rr < kBvOy; rr += 8) {
    const int loy = threadIdx.y + rr, oy = oy0 + loy;
    if (oy >= g.ny || loy >= g.oyn) break;
    if (!dirty_at(dm, g.x0 + ox * g.sx, g.y0 + oy * g.sy)) continue;
    double acc = 0.0;
    const int base = loy * g.sy;
    for (int k = 0; k < nt; ++k) acc += static_cast<double>(s[base + k][threadIdx.x] * taps[k]);
    out[static_cast<size_t>(oy) * out_pitch + ox] = static_cast<float>(acc * scale_y[oy]);
  }
}

// The V pass with its (rows x 32 columns) tile of the H-pass scratch brought in by TMA: box = 32 x rows x 1 at
// (ox0, ys, plane); ox0 is a multiple of 32 (16-byte rule), rows above / below the plane and columns >= nx
// are zero filled like the scalar loader's bounds checks. Dense 32-float rows in shared memory (a warp reads
// 32 consecutive columns of one row: conflict-free without padding).
__global__ void __launch_bounds__(256)
k_blur_v_tma(const __grid_constant__ CUtensorMap tmap, BlurGeom g, int box_rows,
             const double* __restrict__ scale_y, float* __restrict__ out, size_t out_stride, int out_pitch, DirtyMask dm) {
  __shared__ __align__(128) float s[kBvMaxRows * 32];
  __shared__ __align__(8) unsigned long long mbar;
  out += blockIdx.z * out_stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int ox0 = blockIdx.x * 32, oy0 = blockIdx.y * g.oyn;
  if (!dirty_any(dm, g.x0 + ox0 * g.sx, g.y0 + oy0 * g.sy, g.x0 + (ox0 + 31) * g.sx, g.y0 + (oy0 + g.oyn - 1) * g.sy)) return;
  const int ys = g.y0 + oy0 * g.sy - g.r;
  const unsigned mbar_addr = static_cast<unsigned>(__cvta_generic_to_shared(&mbar));
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_addr));
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const unsigned bytes = static_cast<unsigned>(box_rows) * 32 * sizeof(float);
    const unsigned dst = static_cast<unsigned>(__cvta_generic_to_shared(s));
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_addr), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
        ::"r"(dst), "l"(reinterpret_cast<unsigned long long>(&tmap)), "r"(mbar_addr), "r"(ox0), "r"(ys),
          "r"(static_cast<int>(blockIdx.z))
        : "memory");
  }
  {
    unsigned done = 0;
    while (!done) {
      asm volatile(
          "{\n\t.reg .pred p;\n\t"
          "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
          "selp.u32 %0, 1, 0, p;\n\t}"
          : "=r"(done) : "r"(mbar_addr) : "memory");
    }
  }
  const int ox = ox0 + threadIdx.x;
  if (ox >= g.nx) return;
  const float* taps = c_taps[g.kind];
  const int nt = 2 * g.r + 1;
#pragma unroll
  for (int rr = 0; rr < kBvOy; rr += 8) {
    const int loy = threadIdx.y + rr, oy = oy0 + loy;
    if (oy >= g.ny || loy >= g.oyn) break;
    if (!dirty_at(dm, g.x0 + ox * g.sx, g.y0 + oy * g.sy)) continue;
    double acc = 0.0;
    const int base = loy * g.sy;
    for (int k = 0; k < nt; ++k) acc += static_cast<double>(s[(base + k) * 32 + threadIdx.x] * taps[k]);
    out[static_cast<size_t>(oy) * out_pitch + ox] = static_cast<float>(acc * scale_y[oy]);
  }
}

// ---------------------------------------------------------------------------------------------
// K4b: both passes of a small-radius (r <= 3, step 1) blur fused, for the six planes EdgeDetectorMap
// blurs (three channels with their own sigma, two images; butteraugli.cc:1124-1133): one launch
// instead of six. 32x32 output tile per CTA, halo r; the H-pass result of the (32 + 2r) tile rows
// stays in shared memory as float, exactly what the two-kernel path stores between its passes.
// blockIdx.z = image * 3 + channel.
// ---------------------------------------------------------------------------------------------
struct SmallBlur3 {
  int r[3], kind[3];
  const double* sx[3];
  const double* sy[3];
};
constexpr int kSbT = 32, kSbR = 3;
// The only reader of these six planes is k_edge_detector_map below. For the res cell at (3rx, 3ry) it reads the
// columns 3rx - 3, 3rx, 3rx + 3 and 3rx + 4, 3rx + 7, 3rx + 10 (x = px or px + 7, and x -+ 3), i.e. only columns
// with x mod 3 in {0, 1} -- except for the cells clamped to the right border (px = W - 8), which read columns
// >= W - 11. The same holds for rows. So only the outputs with (x mod 3 != 2 or x >= W - 12) and (y mod 3 != 2
// or y >= H - 12) are computed: 4/9 of the V pass and 2/3 of the H pass. The needed columns / rows of the tile are
// listed once per CTA and the work items are spread over the 256 threads in compact form.
__device__ __forceinline__ bool sb_needed(int g, int size) { return g % 3 != 2 || g >= size - 12; }
// (the two passes with the tap count known at compile time: the loops unroll)
template <int R>
__device__ __forceinline__ void sb_passes(float (*s_in)[kSbT + 2 * kSbR + 1], float (*s_h)[kSbT + 1], const unsigned char* s_cols,
                                          const unsigned char* s_rows, int nc, int nr, const float* __restrict__ taps,
                                          const double* __restrict__ scale_x, const double* __restrict__ scale_y, float* __restrict__ out,
                                          int x0, int y0, int H, int P, int tid) {
  constexpr int nt = 2 * R + 1, span = kSbT + 2 * R;
  float tp[nt];
#pragma unroll
  for (int k = 0; k < nt; ++k) tp[k] = taps[k];
  const int dq = 256 / nc, dr = 256 - dq * nc;
  // H pass: every tile row (the V pass needs R rows above and below each needed row), needed columns only
  {
    int ly = tid / nc, ci = tid - ly * nc;
    for (; ly < span; ly += dq, ci += dr) {
      if (ci >= nc) { ci -= nc; ++ly; if (ly >= span) break; }
      const int lx = s_cols[ci];
      const int gx = x0 + lx, gy = y0 + ly - R;
      float v = 0.0f;
      if (gy >= 0 && gy < H) {
        const float* p = &s_in[ly][lx];
        double acc = 0.0;
#pragma unroll
        for (int k = 0; k < nt; ++k) acc += static_cast<double>(p[k] * tp[k]);
        v = static_cast<float>(acc * scale_x[gx]);
      }
      s_h[ly][lx] = v;
    }
  }
  __syncthreads();
  // V pass: needed rows x needed columns
  {
    int ri = tid / nc, ci = tid - ri * nc;
    for (; ri < nr; ri += dq, ci += dr) {
      if (ci >= nc) { ci -= nc; ++ri; if (ri >= nr) break; }
      const int ly = s_rows[ri], lx = s_cols[ci];
      const int gx = x0 + lx, gy = y0 + ly;
      double acc = 0.0;
#pragma unroll
      for (int k = 0; k < nt; ++k) acc += static_cast<double>(s_h[ly + k][lx] * tp[k]);
      out[static_cast<size_t>(gy) * P + gx] = static_cast<float>(acc * scale_y[gy]);
    }
  }
}
__global__ void __launch_bounds__(256)
k_blur_small_hv(const float* __restrict__ in, float* __restrict__ out, size_t plane_stride, int W, int H, int P,
                SmallBlur3 sb, DirtyMask dm) {
  __shared__ float s_in[kSbT + 2 * kSbR][kSbT + 2 * kSbR + 1];
  __shared__ float s_h[kSbT + 2 * kSbR][kSbT + 1];
  __shared__ unsigned char s_cols[kSbT], s_rows[kSbT];
  __shared__ int s_nc, s_nr;
  if (!dirty_at(dm, blockIdx.x * kSbT, blockIdx.y * kSbT)) return;   // the CTA's tile is the mask's tile
  const int ch = blockIdx.z % 3;
  const int r = sb.r[ch];
  const float* taps = c_taps[sb.kind[ch]];
  const double* __restrict__ scale_x = sb.sx[ch];
  const double* __restrict__ scale_y = sb.sy[ch];
  in += blockIdx.z * plane_stride;
  out += blockIdx.z * plane_stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * kSbT, y0 = blockIdx.y * kSbT;
  const int span = kSbT + 2 * r;
  if (tid < 32) {   // the needed columns of the tile, compacted (warp 0)
    const bool need = x0 + tid < W && sb_needed(x0 + tid, W);
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (need) s_cols[__popc(m & ((1u << tid) - 1))] = static_cast<unsigned char>(tid);
    if (tid == 0) s_nc = __popc(m);
  } else if (tid < 64) {   // ... and rows (warp 1)
    const int l = tid - 32;
    const bool need = y0 + l < H && sb_needed(y0 + l, H);
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (need) s_rows[__popc(m & ((1u << l) - 1))] = static_cast<unsigned char>(l);
    if (l == 0) s_nr = __popc(m);
  }
  // (work items are walked as (row, column) pairs with a carried remainder: one division per thread and loop)
  {
    const int dq = 256 / span, dr = 256 - dq * span;
    int ly = tid / span, lx = tid - ly * span;
    for (; ly < span; ly += dq, lx += dr) {
      if (lx >= span) { lx -= span; ++ly; if (ly >= span) break; }
      const int gx = x0 + lx - r, gy = y0 + ly - r;
      float v = 0.0f;
      if (gx >= 0 && gx < W && gy >= 0 && gy < H) v = in[static_cast<size_t>(gy) * P + gx];
      s_in[ly][lx] = v;
    }
  }
  __syncthreads();
  const int nc = s_nc, nr = s_nr;
  if (nc == 0) return;
  if (r == 3) sb_passes<3>(s_in, s_h, s_cols, s_rows, nc, nr, taps, scale_x, scale_y, out, x0, y0, H, P, tid);
  else if (r == 2) sb_passes<2>(s_in, s_h, s_cols, s_rows, nc, nr, taps, scale_x, scale_y, out, x0, y0, H, P, tid);
  else sb_passes<1>(s_in, s_h, s_cols, s_rows, nc, nr, taps, scale_x, scale_y, out, x0, y0, H, P, tid);
}

// ---------------------------------------------------------------------------------------------
// K5: EdgeDetectorMap consumer (butteraugli.cc:689-738, 1135-1148). One res cell per thread,
// reads the six small-sigma blurred planes.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_edge_detector_map(const float* __restrict__ bl0, const float* __restrict__ bl1, size_t stride,
                    int W, int H, int P, int rxs, float* __restrict__ out, DirtyMask dm) {
  const int rx = blockIdx.x * 32 + threadIdx.x, ry = blockIdx.y * 8 + threadIdx.y;
  const int res_x = 3 * rx, res_y = 3 * ry;
  if (!(res_x + 5 < W && res_y + 5 < H)) return;
  if (!dirty_at(dm, res_x, res_y)) return;
  const int px = min(res_x, W - 8), py = min(res_y, H - 8);
  const double w = 0.711100840192;
  double acc[3] = {0.0, 0.0, 0.0};
  int count = 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int x = px + (k >= 2 ? 7 : 0), y = py + ((k & 1) ? 7 : 0);
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      size_t i1, i2;
      if (dir == 0) {
        if (!(x >= 3 && x + 3 < W)) continue;
        i1 = static_cast<size_t>(y) * P + (x - 3);
        i2 = i1 + 6;
      } else {
        if (!(y >= 3 && y + 3 < H)) continue;
        i1 = static_cast<size_t>(y - 3) * P + x;
        i2 = i1 + 6 * static_cast<size_t>(P);
      }
      double d0[3], d1[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        d0[c] = w * (bl0[c * stride + i1] - bl0[c * stride + i2]);  // float difference, widened
        d1[c] = w * (bl1[c * stride + i1] - bl1[c * stride + i2]);
      }
      lowfreq_sq_acc(d0, d1, 1.0, acc);
      ++count;
    }
  }
  const double mul = 0.01617112696 * 8.0 / count;
  float* o = out + 3 * (static_cast<size_t>(ry) * rxs + rx);
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = static_cast<float>(0.0 + mul * acc[c]);
}

// ---------------------------------------------------------------------------------------------
// K6: BlockDiffMap (butteraugli.cc:1081-1117): one warp per res cell, persistent CTAs.
// ---------------------------------------------------------------------------------------------
constexpr int kBdmWarps = 4;
// Which cells must be recomputed after coefficient flips: a cell reads the MaskHighIntensityChange planes
// in its 8x8 window; a sample there depends on the candidate's pixels within 3 (opsin blur 2 + the
// neighbour test 1), and a flipped coefficient changes the pixels of its own 8x8 block only (for a
// sub-sampled chroma block the caller marks every luma block its upsampled samples reach). `chg` has one
// byte per 8x8 block, set by whoever flipped a coefficient since the last Compare (k_be_apply_prefix,
// k_be_apply_walk, k_scatter_coeffs); *enable == 0: no such record, every cell is recomputed.
struct BlockChanges { const uint8_t* chg; const unsigned int* enable; int bw, bh; };
__device__ __forceinline__ bool cell_window_changed(const BlockChanges& bc, int ox, int oy) {
  const int bx0 = max(ox - 3, 0) >> 3, bx1 = min((ox + 10) >> 3, bc.bw - 1);
  const int by0 = max(oy - 3, 0) >> 3, by1 = min((oy + 10) >> 3, bc.bh - 1);
  for (int by = by0; by <= by1; ++by)
    for (int bx = bx0; bx <= bx1; ++bx)
      if (bc.chg[by * bc.bw + bx]) return true;
  return false;
}
__global__ void __launch_bounds__(32 * kBdmWarps)
k_block_diff_map(const float* __restrict__ a, const float* __restrict__ b, size_t stride, int W,
                 int H, int P, int rxs, int ncx, int ncy, float* __restrict__ dc_out,
                 float* __restrict__ ac_out, DirtyMask dm, BlockChanges bc) {
  __shared__ float s_a[kBdmWarps][192];
  __shared__ float s_b[kBdmWarps][192];
  __shared__ double s_ws[kBdmWarps][kBlockDiffScratchDoubles];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int total = ncx * ncy;
  const double csf_a = kCsf8x8[4 + lane], csf_b = kCsf8x8[36];
  const bool fine = bc.chg != nullptr && *bc.enable != 0;
  for (int cell = blockIdx.x * kBdmWarps + warp; cell < total; cell += gridDim.x * kBdmWarps) {
    const int ry = cell / ncx, rx = cell - ry * ncx;
    if (!dirty_at(dm, 3 * rx, 3 * ry)) continue;
    const int ox = min(3 * rx, W - 8), oy = min(3 * ry, H - 8);
    if (fine && !cell_window_changed(bc, ox, oy)) continue;   // (warp-uniform)
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = lane + 32 * k, c = i >> 6, y = (i >> 3) & 7, x = i & 7;
      const size_t g = c * stride + static_cast<size_t>(oy + y) * P + ox + x;
      s_a[warp][i] = a[g];
      s_b[warp][i] = b[g];
    }
    __syncwarp();
    double dc[3], ac[3], edge[3];
    warp_block_diff<false>(s_a[warp], s_b[warp], s_ws[warp], s_ws[warp] + 4 * kBdPlane, csf_a, csf_b, dc, ac, edge);
    if (lane < 3) {
      const size_t o = 3 * (static_cast<size_t>(ry) * rxs + rx) + lane;
      ac_out[o] = static_cast<float>(lane == 0 ? ac[0] : lane == 1 ? ac[1] : ac[2]);
    }
  }
}

// K6': the same AC term, a CTA per vertical strip of kBsCells cells. Cells are 3 px apart and their
// windows 8 px high, so a row of a window is shared by 8/3 cells: the strip computes the average /
// half-difference planes and the row transforms ONCE per image row (29 rows instead of 8 x 8), and every stage
// is a flat loop of independent tasks over the CTA's threads -- the warp-per-cell kernel above leaves 12 of 32
// lanes idle in the column transforms, runs the 33 per-frequency terms as two passes of 32 lanes and the three
// ordered sums on 3 lanes. Per-element arithmetic (and so every bit of the result) is that of warp_block_diff.
constexpr int kBsCells = 8, kBsThreads = 128, kBsMinCtas = 7;
constexpr int kBsRows = 3 * (kBsCells - 1) + 8;      // 29 image rows under a strip
constexpr int kBsPlane = kBsRows * 9;                // a plane: rows of 8 doubles padded to 9
// Strides (in doubles) chosen so that the 16 lanes of a half-warp hit 16 different bank pairs where it matters:
constexpr int kBsSpecRow = 40;                       // a (plane, bin) column of the row spectra; = 8 mod 16: the column
                                                     // transforms run cell-fastest, 8 cells (3 rows apart) x 2 columns
constexpr int kBsPow = 41;                           // power spectrum of one (cell, plane): 40 used bins; odd
constexpr int kBsPowCell = 178;                      // ... of one cell (4 planes); = 2 mod 16
struct BsSmem {
  // Two regions, each reused once its first tenant is dead: the samples and the planes (read by stages 1 and
  // 3) make room for the power spectra (written by stage 4), the row spectra (read by stage 4) for the
  // per-frequency terms (written by stage 5).
  union {
    struct {
      float in[2][3][kBsRows][9];                    // the two images' samples under the strip (rows padded: stage 1 runs row-fastest)
      double pl[4][kBsPlane];                        // y_avg, x/y/z half-differences
    } a;
    double pw[kBsCells][kBsPowCell];                 // power spectra: bin 8u + v of plane p at kBsPow * p + 8u + (v ^ u)
  } u1;
  union {
    struct { double sre[4 * 5][kBsSpecRow], sim[4 * 5][kBsSpecRow]; } b;   // row spectra: [plane * 5 + bin][row]
    double term[kBsCells][3][33];
  } u2;
  double csf[37];                                    // kCsf8x8 (indexed per lane: not from the constant bank)
  int oy[kBsCells];                                  // window origin of each cell (clamped at the bottom border)
  int list[kBsCells], nlist;
  unsigned char row_need[kBsRows + 3];
};
__global__ void __launch_bounds__(kBsThreads, kBsMinCtas)
k_block_diff_strip(const float* __restrict__ a, const float* __restrict__ b, size_t stride, int W, int H, int P, int rxs,
                   int ncx, int ncy, float* __restrict__ ac_out, DirtyMask dm, BlockChanges bc) {
  extern __shared__ unsigned long long bs_dyn[];
  BsSmem& s = *reinterpret_cast<BsSmem*>(bs_dyn);
  const int tid = threadIdx.x;
  const int nsy = (ncy + kBsCells - 1) / kBsCells, total = ncx * nsy;
  const bool fine = bc.chg != nullptr && *bc.enable != 0;
  const bool all = dm.m == nullptr && !fine;   // a full Compare: every cell
  if (tid < 37) s.csf[tid] = kCsf8x8[tid];
  for (int strip = blockIdx.x; strip < total; strip += gridDim.x) {
    const int sy = strip / ncx, rx = strip - sy * ncx;
    const int ry0 = sy * kBsCells, ncell = min(kBsCells, ncy - ry0);
    const int ox = min(3 * rx, W - 8), y_lo = min(3 * ry0, H - 8);
    __syncthreads();   // the previous strip's terms have been summed
    if (tid < 32) {    // warp 0: which cells, in a compact list
      bool need = false;
      int oy = y_lo;
      if (tid < ncell) {
        const int ry = ry0 + tid;
        oy = min(3 * ry, H - 8);
        need = all || (dirty_at(dm, 3 * rx, 3 * ry) && (!fine || cell_window_changed(bc, ox, oy)));
      }
      const unsigned m = __ballot_sync(0xffffffffu, need);
      if (tid < kBsCells) s.oy[tid] = oy;
      if (need) s.list[__popc(m & ((1u << tid) - 1u))] = tid;
      if (tid == 0) s.nlist = __popc(m);
      // rows some needed cell covers
      bool rn = false;
#pragma unroll
      for (int k = 0; k < kBsCells; ++k) {
        const int r0 = __shfl_sync(0xffffffffu, oy, k) - y_lo;
        rn = rn || ((m >> k & 1u) && tid >= r0 && tid < r0 + 8);
      }
      if (tid < kBsRows) s.row_need[tid] = rn ? 1 : 0;
    }
    __syncthreads();
    const int nlist = s.nlist;
    if (nlist == 0) continue;   // (uniform)
    const int nrows = s.oy[ncell - 1] + 8 - y_lo;
    // samples: 2 images x 3 channels x nrows x 8 (all loads of a thread in flight before the first store)
    {
      constexpr int kPer = (6 * kBsRows * 8 + kBsThreads - 1) / kBsThreads;
      float v[kPer];
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kBsThreads;
        const int x = i & 7, r = (i >> 3) % kBsRows, pc = i / (8 * kBsRows);   // pc = image * 3 + channel
        v[j] = 0.0f;
        if (i < 6 * kBsRows * 8 && r < nrows && s.row_need[r]) {
          const int ch = pc >= 3 ? pc - 3 : pc;
          v[j] = __ldg((pc >= 3 ? b : a) + ch * stride + static_cast<size_t>(y_lo + r) * P + ox + x);
        }
      }
#pragma unroll
      for (int j = 0; j < kPer; ++j) {
        const int i = tid + j * kBsThreads;
        if (i < 6 * kBsRows * 8) (&s.u1.a.in[0][0][0][0])[(i >> 3) * 9 + (i & 7)] = v[j];
      }
    }
    __syncthreads();
    // (1) planes (row-fastest: the strides of 9 floats / 9 doubles keep the lanes on different banks)
    for (int i = tid; i < kBsRows * 8; i += kBsThreads) {
      const int x = i / kBsRows, r = i - x * kBsRows;
      if (r < nrows && s.row_need[r]) {
        const double a0 = s.u1.a.in[0][0][r][x], a1 = s.u1.a.in[0][1][r][x], a2 = s.u1.a.in[0][2][r][x];
        const double b0 = s.u1.a.in[1][0][r][x], b1 = s.u1.a.in[1][1][r][x], b2 = s.u1.a.in[1][2][r][x];
        const int o = r * 9 + x;
        s.u1.a.pl[0][o] = (a1 + b1) / 2;
        s.u1.a.pl[1][o] = (a0 - b0) / 2;
        s.u1.a.pl[2][o] = (a1 - b1) / 2;
        s.u1.a.pl[3][o] = (a2 - b2) / 2;
      }
    }
    __syncthreads();
    // (3) row transforms, once per (plane, image row)
    for (int t = tid; t < 4 * kBsRows; t += kBsThreads) {
      const int plane = t / kBsRows, r = t - plane * kBsRows;
      if (r < nrows && s.row_need[r]) {
        double x[8], re[5], im[5];
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = s.u1.a.pl[plane][9 * r + k];
        rfft8_half(x, re, im);
#pragma unroll
        for (int u = 0; u < 5; ++u) {
          s.u2.b.sre[plane * 5 + u][r] = re[u];
          s.u2.b.sim[plane * 5 + u][r] = im[u];
        }
      }
    }
    __syncthreads();
    // (4) column transforms: task = (needed cell, plane, bin u). The power spectra overwrite the samples and
    // planes, which nobody reads any more.
    for (int t = tid; t < 20 * nlist; t += kBsThreads) {
      const int pu = t / nlist, li = t - pu * nlist, cell = s.list[li], plane = pu / 5, u = pu - 5 * plane;   // cell-fastest
      const int r0 = s.oy[cell] - y_lo;
      double re[8], im[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) { re[k] = s.u2.b.sre[pu][r0 + k]; im[k] = s.u2.b.sim[pu][r0 + k]; }
      cfft8(re, im);
      double* dst = &s.u1.pw[cell][kBsPow * plane + 8 * u];
#pragma unroll
      for (int v = 0; v < 8; ++v) dst[v ^ u] = (re[v] * re[v] + im[v] * im[v]) * 0.000064;
    }
    __syncthreads();
    // (5) per-frequency terms i = 4..36 of every needed cell (over the row spectra)
    for (int t = tid; t < 33 * nlist; t += kBsThreads) {
      const int li = t / 33, k = t - 33 * li, i = 4 + k, cell = s.list[li];
      const double d = s.csf[i];
      const int pi = i ^ (i >> 3);
      const double* pw = s.u1.pw[cell];
      const double tx = d * 64.8 * pw[kBsPow + pi];
      const double tz = d * 2.4 * pw[3 * kBsPow + pi];
      const double ya = sqrt(pw[pi]), yh = sqrt(pw[2 * kBsPow + pi]);
      const double y0 = remove_range_around_zero(ya - yh, 0.04);
      const double y1 = remove_range_around_zero(ya + yh, 0.04);
      double ty = 0.0;
      if (y0 != y1) {
        const double v0 = interp_signed21(g_tab.lut21[1], y0 * 1.51983458269);
        const double v1 = interp_signed21(g_tab.lut21[1], y1 * 1.51983458269);
        const double vy = 1.753123908348329 * (v0 - v1);
        ty = d * vy * vy;
      }
      s.u2.term[cell][0][k] = tx;
      s.u2.term[cell][1][k] = ty;
      s.u2.term[cell][2][k] = tz;
    }
    __syncthreads();
    // the ordered sums: one lane per (cell, channel)
    if (tid < 3 * nlist) {
      const int li = tid / 3, c = tid - 3 * li, cell = s.list[li];
      const double* t = s.u2.term[cell][c];
      double sum = 0.0;
#pragma unroll
      for (int i = 0; i < 33; ++i) sum += t[i];
      ac_out[3 * (static_cast<size_t>(ry0 + cell) * rxs + rx) + c] = static_cast<float>(sum);
    }
  }
}

// K6b: the DC term of BlockDiffMap (butteraugli.cc:602-617, 1098-1106), one res cell per thread: the
// mean of the 64 per-pixel half-differences in the reference's order, then the low-frequency
// colour metric. (The warp-per-cell kernel would spend 64 dependent additions on three lanes.)
__global__ void __launch_bounds__(128)
k_block_dc(const float* __restrict__ a, const float* __restrict__ b, size_t stride, int W, int H, int P,
           int rxs, int ncx, int ncy, float* __restrict__ dc_out, DirtyMask dm, BlockChanges bc) {
  const int cell = blockIdx.x * blockDim.x + threadIdx.x;
  if (cell >= ncx * ncy) return;
  const int ry = cell / ncx, rx = cell - ry * ncx;
  if (!dirty_at(dm, 3 * rx, 3 * ry)) return;
  const int ox = min(3 * rx, W - 8), oy = min(3 * ry, H - 8);
  if (bc.chg != nullptr && *bc.enable != 0 && !cell_window_changed(bc, ox, oy)) return;
  double m[3];
#pragma unroll 1
  for (int c = 0; c < 3; ++c) {
    const float* pa = a + c * stride + static_cast<size_t>(oy) * P + ox;
    const float* pb = b + c * stride + static_cast<size_t>(oy) * P + ox;
    double acc = 0.0;
#pragma unroll 1
    for (int y = 0; y < 8; ++y) {
#pragma unroll
      for (int x = 0; x < 8; ++x) acc += (static_cast<double>(pa[x]) - static_cast<double>(pb[x])) / 2;
      pa += P;
      pb += P;
    }
    m[c] = acc / 32;
  }
  double sq[3] = {0.0, 0.0, 0.0};
  lowfreq_sq_acc0(m, kCsf8x8[0], sq);
  float* o = dc_out + 3 * (static_cast<size_t>(ry) * rxs + rx);
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = static_cast<float>(sq[c]);
}

// K6b': the same DC term, a CTA per 64 horizontally adjacent cells. Cells are 3 px apart and their windows 8 px
// wide, so a sample is read by 8/3 cells of a row: the CTA loads the 8 image rows under its cells once (coalesced),
// keeps the half-differences (a - b) / 2 of the three channels as doubles in shared memory -- two conversions
// per SAMPLE instead of per term -- and one thread per (cell, channel) adds its 64 in the reference's order.
constexpr int kBdcCells = 64, kBdcThreads = 224;   // >= the 197 samples of a row and >= 3 x 64 sums
constexpr int kBdcWidth = 3 * (kBdcCells - 1) + 8;   // 197 samples under 64 cells
constexpr int kBdcPitch = kBdcWidth + 1;             // (cells are 3 doubles apart: any pitch keeps 16 lanes on 16 bank pairs)
__global__ void __launch_bounds__(kBdcThreads)
k_block_dc_rows(const float* __restrict__ a, const float* __restrict__ b, size_t stride, int W, int H, int P,
                int rxs, int ncx, int ncy, float* __restrict__ dc_out, DirtyMask dm, BlockChanges bc) {
  __shared__ double s_hd[3][8][kBdcPitch];
  __shared__ double s_m[3][kBdcCells];
  __shared__ unsigned char s_need[kBdcCells];
  const int tid = threadIdx.x;
  const int ry = blockIdx.y, rx0 = blockIdx.x * kBdcCells;
  const int ncell = min(kBdcCells, ncx - rx0);
  const int oy = min(3 * ry, H - 8), x_lo = min(3 * rx0, W - 8);
  const bool fine = bc.chg != nullptr && *bc.enable != 0;
  bool need = false;
  if (tid < kBdcCells) {
    if (tid < ncell) {
      const int rx = rx0 + tid, ox = min(3 * rx, W - 8);
      need = dirty_at(dm, 3 * rx, 3 * ry) && (!fine || cell_window_changed(bc, ox, oy));
    }
    s_need[tid] = need ? 1 : 0;
  }
  if (!__syncthreads_or(need ? 1 : 0)) return;
  const int wid = min(3 * (rx0 + ncell - 1), W - 8) + 8 - x_lo;   // samples under the CTA's cells
  // half-differences of the three channels: 24 rows of wid samples, a thread per sample of a row
  if (tid < wid) {
    const float* pa = a + static_cast<size_t>(oy) * P + x_lo + tid;
    const float* pb = b + static_cast<size_t>(oy) * P + x_lo + tid;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
#pragma unroll
      for (int y = 0; y < 8; ++y) {
        const size_t g = c * stride + static_cast<size_t>(y) * P;
        s_hd[c][y][tid] = (static_cast<double>(__ldg(pa + g)) - static_cast<double>(__ldg(pb + g))) / 2;
      }
    }
  }
  __syncthreads();
  {
    const int cell = tid % kBdcCells, c = tid / kBdcCells;
    if (c < 3 && cell < ncell && s_need[cell]) {
      const int x0 = min(3 * (rx0 + cell), W - 8) - x_lo;
      double acc = 0.0;
#pragma unroll 1
      for (int y = 0; y < 8; ++y) {
        const double* r = &s_hd[c][y][x0];
#pragma unroll
        for (int x = 0; x < 8; ++x) acc += r[x];
      }
      s_m[c][cell] = acc / 32;
    }
  }
  __syncthreads();
  if (tid < ncell && s_need[tid]) {
    const double m[3] = {s_m[0][tid], s_m[1][tid], s_m[2][tid]};
    double sq[3] = {0.0, 0.0, 0.0};
    lowfreq_sq_acc0(m, kCsf8x8[0], sq);
    float* o = dc_out + 3 * (static_cast<size_t>(ry) * rxs + rx0 + tid);
#pragma unroll
    for (int c = 0; c < 3; ++c) o[c] = static_cast<float>(sq[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// K7: EdgeDetectorLowFreq consumer (butteraugli.cc:1164-1204). Reads the decimated sigma-14 maps:
// blurred[y][x] == small[y/4][x/4]. One lattice point per thread. The reference adds the term into
// block_diff_ac; here it is stored in its own res map (same cell, two to the right of the lattice
// point) and k_combine performs the float addition, so that the term can be refreshed on its own.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_edge_lowfreq(const float* __restrict__ s0, const float* __restrict__ s1, size_t stride,
               int spitch, int step, int W, int H, int rxs, float* __restrict__ term_out, DirtyMask dm) {
  const int cx = blockIdx.x * 32 + threadIdx.x, cy = blockIdx.y * 8 + threadIdx.y;
  const int x = 3 * cx, y = 3 * cy;
  if (!(x + 8 < W && y + 8 < H)) return;
  if (!dirty_at(dm, x, y)) return;
  const int ox[4] = {x + 8, x, x + 6, x - 6};
  const int oy[4] = {y, y + 8, y + 6, y + 6};
  const size_t i0 = static_cast<size_t>(y / step) * spitch + x / step;
  double best[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    double d[3] = {0.0, 0.0, 0.0};
    if (!(k == 3 && x < 8)) {
      const size_t i2 = static_cast<size_t>(oy[k] / step) * spitch + ox[k] / step;
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        const float f = (s1[c * stride + i0] - s0[c * stride + i0]) +
                        (s0[c * stride + i2] - s1[c * stride + i2]);
        d[c] = f;
      }
    }
    double sq[3] = {0.0, 0.0, 0.0};
    lowfreq_sq_acc0(d, 1.0, sq);
#pragma unroll
    for (int c = 0; c < 3; ++c) best[c] = best[c] < sq[c] ? sq[c] : best[c];
  }
  float* o = term_out + 3 * (static_cast<size_t>(cy) * rxs + cx + 2);
#pragma unroll
  for (int c = 0; c < 3; ++c) o[c] = static_cast<float>(10 * best[c]);
}

// ---------------------------------------------------------------------------------------------
// K8: front of Mask (butteraugli.cc:1440-1493, 1379-1438, 1332-1376) fused per channel:
//   DiffPrecompute -> Average5x5 (reference float summation order) -> MinSquareVal(4,0).
// 32x32 output tile per CTA; blockIdx.z = channel (the three channels are independent).
// ---------------------------------------------------------------------------------------------
constexpr int kMpT = 32;
__global__ void __launch_bounds__(256)
k_mask_front(const float* __restrict__ a, const float* __restrict__ b, size_t stride, int W, int H,
             int P, float* __restrict__ out, DirtyMask dm) {
  if (!dirty_at(dm, blockIdx.x * kMpT, blockIdx.y * kMpT)) return;   // the CTA's tile is the mask's tile
  __shared__ float s_a[kMpT + 7][kMpT + 7 + 1];   // xyb tile, origin (-2,-2), 39 x 39
  __shared__ float s_b[kMpT + 7][kMpT + 7 + 1];
  __shared__ float s_p[kMpT + 5][kMpT + 5 + 1];   // precompute, origin (-1,-1)
  __shared__ float s_v[kMpT + 3][kMpT + 3 + 1];   // averaged, origin (0,0)
  const int c = blockIdx.z;
  a += c * stride;
  b += c * stride;
  out += c * stride;
  const int tid = threadIdx.y * 32 + threadIdx.x;
  const int x0 = blockIdx.x * kMpT, y0 = blockIdx.y * kMpT;
  // xyb tile, origin (-2,-2): the last column/row use their left/upper neighbour instead.
  for (int i = tid; i < 39 * 39; i += 256) {
    const int ly = i / 39, lx = i - ly * 39;
    const int gx = x0 + lx - 2, gy = y0 + ly - 2;
    float va = 0.0f, vb = 0.0f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
      const size_t g = static_cast<size_t>(gy) * P + gx;
      va = a[g];
      vb = b[g];
    }
    s_a[ly][lx] = va;
    s_b[ly][lx] = vb;
  }
  __syncthreads();
  for (int i = tid; i < 37 * 37; i += 256) {
    const int ly = i / 37, lx = i - ly * 37;
    const int gx = x0 + lx - 1, gy = y0 + ly - 1;
    float r = 0.0f;
    if (gx >= 0 && gx < W && gy >= 0 && gy < H) {
      const int tx = lx + 1, ty = ly + 1;  // position in the xyb tile
      const int hx = gx + 1 < W ? tx + 1 : tx - 1;
      const int vy = gy + 1 < H ? ty + 1 : ty - 1;
      const double h0 = highfreq_val(c, s_a[ty][tx] - s_a[ty][hx]);
      const double h1 = highfreq_val(c, s_b[ty][tx] - s_b[ty][hx]);
      const double v0 = highfreq_val(c, s_a[ty][tx] - s_a[vy][tx]);
      const double v1 = highfreq_val(c, s_b[ty][tx] - s_b[vy][tx]);
      const double sup0 = fabs(h0) + fabs(v0), sup1 = fabs(h1) + fabs(v1);
      r = static_cast<float>(sup1 < sup0 ? sup1 : sup0);
    }
    s_p[ly][lx] = r;
  }
  __syncthreads();
  const float w = 0.679144890667f;
  const float scale = 1.0f / (5.0f + 4 * w);
  for (int i = tid; i < 35 * 35; i += 256) {
    const int ly = i / 35, lx = i - ly * 35;
    const int gx = x0 + lx, gy = y0 + ly;
    float r = __int_as_float(0x7f800000);  // +inf outside the image: ignored by the min
    if (gx < W && gy < H) {
      const int px = lx + 1, py = ly + 1;
      // Outside-image neighbours hold 0.0f in s_p; the reference drops them (x + 0 == x).
      float acc = s_p[py][px];
      acc += s_p[py - 1][px - 1] * w;
      acc += s_p[py - 1][px];
      acc += s_p[py - 1][px + 1] * w;
      acc += s_p[py][px - 1];
      acc += s_p[py][px + 1];
      acc += s_p[py + 1][px - 1] * w;
      acc += s_p[py + 1][px];
      acc += s_p[py + 1][px + 1] * w;
      r = acc * scale;
    }
    s_v[ly][lx] = r;
  }
  __syncthreads();
  const int gx = x0 + threadIdx.x;
  if (gx >= W) return;
#pragma unroll
  for (int ly = threadIdx.y; ly < kMpT; ly += 8) {
    const int gy = y0 + ly;
    if (gy >= H) break;
    float m = s_v[ly][threadIdx.x];
#pragma unroll
    for (int dy = 0; dy < 4; ++dy)
#pragma unroll
      for (int dx = 0; dx < 4; ++dx) m = fminf(m, s_v[ly + dy][threadIdx.x + dx]);
    out[static_cast<size_t>(gy) * P + gx] = m;
  }
}

// ---------------------------------------------------------------------------------------------
// K9: mask LUTs at the sampled pixel + CombineChannels + sqrt of CalculateDiffmap
//     (butteraugli.cc:1538-1566, 1207-1231, 1002-1008). One res cell per thread.
//     Blurred mask planes are read from their lattices: channel c value at pixel (px,py) is
//     m[c][(py - my0[c]) / msy[c]][(px - mx0[c]) / msx[c]].
// ---------------------------------------------------------------------------------------------
struct MaskSample {
  const float* m[3];
  int pitch[3], x0[3], sx[3], y0[3], sy[3];
};
__device__ __forceinline__ void mask_at(const MaskSample& ms, int px, int py, double mask[3],
                                        double mask_dc[3], bool want_dc) {
  const double wmul[3] = {232.206464018, 22.9455222245, 503.962310606};
  const float gs = static_cast<float>((1.0 / 14.921561160295326) * (1.0 / 14.921561160295326));
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float s = ms.m[c][static_cast<size_t>((py - ms.y0[c]) / ms.sy[c]) * ms.pitch[c] +
                            (px - ms.x0[c]) / ms.sx[c]];
    const double p = wmul[c] * static_cast<double>(s);
    float m = static_cast<float>(interp_clamp512(g_tab.mask_lut[c], p));
    m *= gs;
    mask[c] = m;
    if (want_dc) {
      float d = static_cast<float>(interp_clamp512(g_tab.mask_lut[3 + c], p));
      d *= gs;
      mask_dc[c] = d;
    }
  }
}

__global__ void __launch_bounds__(256)
k_combine(MaskSample ms, const float* __restrict__ dc, const float* __restrict__ ac, const float* __restrict__ lft,
          const float* __restrict__ edm, int W, int H, int rxs, int rys, int sq_pitch,
          float* __restrict__ sq, DirtyMask dm) {
  const int rx = blockIdx.x * 32 + threadIdx.x, ry = blockIdx.y * 8 + threadIdx.y;
  if (rx >= rxs || ry >= rys) return;
  if (!dirty_at(dm, 3 * rx, 3 * ry)) return;
  float r = 0.0f;
  if (3 * rx + 5 < W && 3 * ry + 5 < H) {
    double mask[3], mdc[3];
    mask_at(ms, 3 * rx + 3, 3 * ry + 3, mask, mdc, true);
    const size_t o = 3 * (static_cast<size_t>(ry) * rxs + rx);
    const double t0 = dc[o] * mdc[0] + dc[o + 1] * mdc[1] + dc[o + 2] * mdc[2];
    // block_diff_ac after EdgeDetectorLowFreq: the float sum the reference keeps (butteraugli.cc:1201)
    const float a0 = ac[o] + lft[o], a1 = ac[o + 1] + lft[o + 1], a2 = ac[o + 2] + lft[o + 2];
    const double t1 = a0 * mask[0] + a1 * mask[1] + a2 * mask[2];
    const double t2 = edm[o] * mask[0] + edm[o + 1] * mask[1] + edm[o + 2] * mask[2];
    const float v = static_cast<float>(t0 + t1 + t2);
    r = v < (1.0 / (100.0f * 100.0f)) ? 100.0f * v : sqrtf(v);
  }
  sq[static_cast<size_t>(ry) * sq_pitch + rx] = r;
}

// Sampled mask only (StartBlockComparisons: mask_xyz_ at each 8x8 block's top-left pixel;
// guetzli/butteraugli_comparator.cc:72-79, 148-151). out: [B8][3] floats.
__global__ void k_block_mask_scale(MaskSample ms, int bw, int bh, float* __restrict__ out) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bw * bh) return;
  double mask[3], unused[3];
  mask_at(ms, 8 * (b % bw), 8 * (b / bw), mask, unused, false);
#pragma unroll
  for (int c = 0; c < 3; ++c) out[3 * b + c] = static_cast<float>(mask[c]);
}

// The full-resolution mask / mask_dc planes of butteraugli::Mask (butteraugli.cc:1505-1566), for the
// standalone gzb_mask entry point only: the Compare pipeline never materialises them (k_combine looks
// the tables up at the one pixel per res cell it needs).
__global__ void __launch_bounds__(256)
k_mask_full(MaskSample ms, int W, int H, int P, float* __restrict__ mask, float* __restrict__ mask_dc, size_t stride) {
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (x >= W || y >= H) return;
  double m[3], d[3];
  mask_at(ms, x, y, m, d, true);
  const size_t o = static_cast<size_t>(y) * P + x;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    mask[c * stride + o] = static_cast<float>(m[c]);
    mask_dc[c * stride + o] = static_cast<float>(d[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// K10: tail of CalculateDiffmap (butteraugli.cc:1009-1043) + score (1233-1240):
//   diffmap = (up + 24.82f * blur(up_crop)) * 1/25.82, max-reduced into *dist_bits.
// `sq` is the res-grid sqrt map, `small` the decimated (step 2) sigma-8.85 blur of the crop.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_diffmap_final(const float* __restrict__ sq, int sq_pitch, const float* __restrict__ small,
                int small_pitch, int bstep, int W, int H, int out_pitch, float* __restrict__ out,
                unsigned int* __restrict__ cta_max, DirtyMask dm) {
  __shared__ unsigned int s_max;
  const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
  if (!dirty_at(dm, blockIdx.x * 32, blockIdx.y * 8)) return;   // a 32x8 CTA lies inside one mask tile
  if (threadIdx.x == 0 && threadIdx.y == 0) s_max = 0u;
  __syncthreads();
  float v = 0.0f;
  if (x < W && y < H) {
    float up = 0.0f;
    if (x >= 2 && y >= 2) {
      const int cx = (x - 2) / 3, cy = (y - 2) / 3;
      if (3 * cx + 5 < W && 3 * cy + 5 < H) up = sq[static_cast<size_t>(cy) * sq_pitch + cx];
    }
    v = up;
    if (x >= 2 && y >= 2 && x - 2 < W - 5 && y - 2 < H - 5) {
      const float mul1 = static_cast<float>(24.8235314874);
      v += mul1 * small[static_cast<size_t>((y - 2) / bstep) * small_pitch + (x - 2) / bstep];
    }
    v *= static_cast<float>(1.0 / (1.0 + 24.8235314874));
    out[static_cast<size_t>(y) * out_pitch + x] = v;
  }
  // all values are >= 0: unsigned ordering of the bit patterns equals float ordering
  unsigned int bits = __float_as_uint(v > 0.0f ? v : 0.0f);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) bits = max(bits, __shfl_xor_sync(0xffffffffu, bits, o));
  if (threadIdx.x == 0 && bits != 0) atomicMax(&s_max, bits);
  __syncthreads();
  // the maximum of every CTA's pixels stays in cta_max, so that a later Compare that skips this CTA
  // still finds it (k_max_u32 reduces the array)
  if (threadIdx.x == 0 && threadIdx.y == 0) cta_max[blockIdx.y * gridDim.x + blockIdx.x] = s_max;
}

// Maximum of n unsigned values into *out (which the caller zeroes): the score of the diffmap.
__global__ void __launch_bounds__(256)
k_max_u32(const unsigned int* __restrict__ v, int n, unsigned int* __restrict__ out) {
  unsigned int m = 0u;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) m = max(m, v[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m != 0u) atomicMax(out, m);
}

// ---------------------------------------------------------------------------------------------
// K11: ComputeBlockErrorAdjustmentWeights (guetzli/butteraugli_comparator.cc:169-233); bs = 8 * factor.
// ---------------------------------------------------------------------------------------------
__global__ void k_block_max(const float* __restrict__ dm, int pitch, int W, int H, int bw, int bh, int bs,
                            float* __restrict__ bmax) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bw * bh) return;
  const int bx = b % bw, by = b / bw;
  float m = 0.0f;
  for (int y = bs * by; y < min(H, bs * by + bs); ++y)
    for (int x = bs * bx; x < min(W, bs * bx + bs); ++x) m = fmaxf(m, dm[static_cast<size_t>(y) * pitch + x]);
  bmax[b] = m;
}
// flags[b] = 1 if block b passes the per-direction test of the reference.
__global__ void k_block_flags(const float* __restrict__ bmax, int bw, int bh, int direction, int rad,
                              double target, unsigned char* __restrict__ flags) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bw * bh) return;
  const int bx = b % bw, by = b / bw;
  float local = static_cast<float>(target);
  for (int y = max(0, by - rad); y < min(bh, by + 1 + rad); ++y)
    for (int x = max(0, bx - rad); x < min(bw, bx + 1 + rad); ++x) local = fmaxf(local, bmax[y * bw + x]);
  const double mine = bmax[b];
  bool f;
  if (direction > 0) f = mine <= target && static_cast<double>(local) <= 1.1 * target;
  else f = !(mine <= (1 - 0.5) * target + 0.5 * static_cast<double>(local));
  flags[b] = f ? 1 : 0;
}
__global__ void k_block_weights(const unsigned char* __restrict__ flags, int bw, int bh,
                                int direction, int rad, float* __restrict__ weight) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= bw * bh) return;
  const int bx = b % bw, by = b / bw;
  float w = 0.0f;
  if (direction > 0) {
    w = flags[b] ? 1.0f : 0.0f;
  } else {
    for (int y = max(0, by - rad); y < min(bh, by + 1 + rad); ++y)
      for (int x = max(0, bx - rad); x < min(bw, bx + 1 + rad); ++x)
        if (flags[y * bw + x]) {
          const int d = max(abs(y - by), abs(x - bx));
          w = fmaxf(w, 1.0f / (d + 1.0f));
        }
  }
  weight[b] = w;
}

// ---------------------------------------------------------------------------------------------
// K12: double-precision 8x8 DCT / IDCT (guetzli/dct_double.cc:28-85), batched: 64 threads per block,
// columns then rows, each output an in-order 8-term double sum against the 10-decimal matrix.
// Only the 4:2:0 path of the reference reaches it (output_image.cc:100-122, 496-531).
// ---------------------------------------------------------------------------------------------
__constant__ double c_dct_matrix[64];   // kDCTMatrix[8*u + x], filled by init_device_tables()

__global__ void __launch_bounds__(256)
k_dct_double(double* __restrict__ blocks, size_t nblocks, int inverse) {
  __shared__ double s_in[4][64];
  __shared__ double s_tmp[4][64];
  const int lb = threadIdx.x >> 6, t = threadIdx.x & 63;
  const size_t b = blockIdx.x * static_cast<size_t>(4) + lb;
  const bool live = b < nblocks;
  if (live) s_in[lb][t] = blocks[b * 64 + t];
  __syncthreads();
  {  // column pass: tmp[8*xo + col] = sum_u M(xo,u) * in[8*u + col]
    const int xo = t >> 3, col = t & 7;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (inverse ? c_dct_matrix[8 * u + xo] : c_dct_matrix[8 * xo + u]) * s_in[lb][8 * u + col];
    s_tmp[lb][8 * xo + col] = acc;
  }
  __syncthreads();
  {  // row pass: out[8*row + xo] = sum_u M(xo,u) * tmp[8*row + u]
    const int row = t >> 3, xo = t & 7;
    double acc = 0.0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += (inverse ? c_dct_matrix[8 * u + xo] : c_dct_matrix[8 * xo + u]) * s_tmp[lb][8 * row + u];
    if (live) blocks[b * 64 + 8 * row + xo] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// K13: candidate packing of SelectFrequencyMasking (guetzli/processor.cc:694-712) on the device:
// per block the records with 0 < err <= limit, in order; offsets by an exclusive scan.
// ---------------------------------------------------------------------------------------------
struct CoeffRec { int idx; float err; };
__global__ void k_count_candidates(const CoeffRec* __restrict__ order, int nblocks, float limit,
                                   int* __restrict__ counts) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nblocks) return;
  const CoeffRec* p = order + static_cast<size_t>(warp) * 192;
  int n = 0;
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const float e = p[lane + 32 * k].err;
    n += __popc(__ballot_sync(0xffffffffu, e > 0 && e <= limit));
  }
  if (lane == 0) counts[warp] = n;
}
// Single-CTA exclusive scan (nblocks <= a few hundred thousand): offsets[0..nblocks].
__global__ void __launch_bounds__(1024)
k_scan_counts(const int* __restrict__ counts, int nblocks, int* __restrict__ offsets) {
  __shared__ int s_part[1024];
  const int t = threadIdx.x;
  const int per = (nblocks + 1023) / 1024;
  const int b0 = min(nblocks, t * per), b1 = min(nblocks, b0 + per);
  int sum = 0;
  for (int b = b0; b < b1; ++b) sum += counts[b];
  s_part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int v = t >= off ? s_part[t - off] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  int run = s_part[t] - sum;
  for (int b = b0; b < b1; ++b) { offsets[b] = run; run += counts[b]; }
  if (t == 1023) offsets[nblocks] = s_part[1023];
}
// The same scan over many CTAs (the back end scans one count per block every iteration; the single-CTA
// version above walks strided memory and takes 0.2 ms at 12 MPix): chunk sums, a scan of the <= 1024 chunk
// sums, then every chunk scans itself on top of its offset. Chunks of kScanChunk counts.
constexpr int kScanChunk = 2048;   // 256 threads x 8
__global__ void __launch_bounds__(256)
k_scan_chunk_sums(const int* __restrict__ counts, int n, int* __restrict__ sums) {
  __shared__ int s_w[8];
  const int base = blockIdx.x * kScanChunk;
  int v = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int i = base + j * 256 + threadIdx.x;
    if (i < n) v += counts[i];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    int t = 0;
    for (int w = 0; w < 8; ++w) t += s_w[w];
    sums[blockIdx.x] = t;
  }
}
// exclusive scan of up to 1024 chunk sums in place; offsets[n] = total
__global__ void __launch_bounds__(1024)
k_scan_chunk_offsets(int* __restrict__ sums, int nchunks, int* __restrict__ offsets, int n) {
  __shared__ int s[1024];
  const int t = threadIdx.x;
  const int v = t < nchunks ? sums[t] : 0;
  s[t] = v;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const int a = t >= off ? s[t - off] : 0;
    __syncthreads();
    s[t] += a;
    __syncthreads();
  }
  if (t < nchunks) sums[t] = s[t] - v;
  if (t == 1023) offsets[n] = s[1023];
}
__global__ void __launch_bounds__(256)
k_scan_chunk_apply(const int* __restrict__ counts, int n, const int* __restrict__ sums, int* __restrict__ offsets) {
  __shared__ int s_w[8];
  const int base = blockIdx.x * kScanChunk + threadIdx.x * 8;   // 8 consecutive counts per thread
  int c[8], tot = 0;
#pragma unroll
  for (int j = 0; j < 8; ++j) { c[j] = base + j < n ? counts[base + j] : 0; tot += c[j]; }
  int inc = tot;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int a = __shfl_up_sync(0xffffffffu, inc, off);
    if ((threadIdx.x & 31) >= off) inc += a;
  }
  if ((threadIdx.x & 31) == 31) s_w[threadIdx.x >> 5] = inc;
  __syncthreads();
  int run = sums[blockIdx.x] + inc - tot;
  for (int w = 0; w < (threadIdx.x >> 5); ++w) run += s_w[w];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    if (base + j < n) offsets[base + j] = run;
    run += c[j];
  }
}

__global__ void k_pack_candidates(const CoeffRec* __restrict__ order, int nblocks, float limit,
                                  const int* __restrict__ offsets, uint8_t* __restrict__ out_idx,
                                  float* __restrict__ out_err) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= nblocks) return;
  const CoeffRec* p = order + static_cast<size_t>(warp) * 192;
  int base = offsets[warp];
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const CoeffRec r = p[lane + 32 * k];
    const bool take = r.err > 0 && r.err <= limit;
    const unsigned m = __ballot_sync(0xffffffffu, take);
    if (take) {
      const int o = base + __popc(m & ((1u << lane) - 1));
      out_idx[o] = static_cast<uint8_t>(r.idx);
      out_err[o] = r.err;
    }
    base += __popc(m);
  }
}

// ---------------------------------------------------------------------------------------------
// Measurement aid: the non-FMA double-precision rate of the device. The search kernels are compiled
// with -fmad=false (the reference's CPU arithmetic has no contraction), so their ceiling is one DADD or
// DMUL per FP64 lane and clock, not the FMA rate of the data sheet. Eight independent mul/add chains per
// thread keep the pipe full; `iters` x 16 flops per thread.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_fp64_peak(double* __restrict__ out, int iters, double m, double c) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
#pragma unroll 4
  for (int i = 0; i < iters; ++i) {
    a0 = a0 * m; a1 = a1 * m; a2 = a2 * m; a3 = a3 * m; a4 = a4 * m; a5 = a5 * m; a6 = a6 * m; a7 = a7 * m;
    a0 = a0 + c; a1 = a1 + c; a2 = a2 + c; a3 = a3 + c; a4 = a4 + c; a5 = a5 + c; a6 = a6 + c; a7 = a7 + c;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

}  // namespace gzb
