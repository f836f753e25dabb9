/* oracle/gzoracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C, single-threaded restatement of the Guetzli hot path (butteraugli-guided quantisation
 * search) with the reference's MODE_CPU arithmetic: float storage, double intermediates, the same
 * operation order. It exists only to CHECK the CUDA path: tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline leg may load it; nothing under guetzli-cuda-opencl_b200/ does.
 *
 * Parity status: PINNED. Every function here is checked bit-for-bit against the unmodified
 * reference compiled into oracle/_ref/libgzref.so (tests/test_oracle_vs_ref.py) and, through it,
 * against the reference's one offline-reproducible golden vector (tests/golden_checksums.txt:3,
 * bees.png at q95) -- see tests/test_golden_bees.py.
 *
 * Citations are to files under /root/reference/.
 * Build: gcc -std=gnu11 -O2 -ffp-contract=off (no FMA contraction: the reference build has none).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gzoracle_tables.h"

#define GZO_API __attribute__((visibility("default")))

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------------------------------
 * sRGB8 -> linear LUT.  guetzli/gamma_correct.cc:23-33
 * ---------------------------------------------------------------------------------------- */
GZO_API void gzo_srgb8_to_linear_table(double* t) {
  for (int i = 0; i < 256; ++i)
    t[i] = i < 11 ? i / 12.92 : 255.0 * pow(((i / 255.0) + 0.055) / 1.055, 2.4);
}

/* ------------------------------------------------------------------------------------------
 * Separable Gaussian blur.  third_party/butteraugli/butteraugli/butteraugli.cc:68-148
 * ---------------------------------------------------------------------------------------- */
/* One pass: convolve rows at every `step`-th column and write the result transposed. */
static void conv_pass(int xs, int ys, int step, int radius, const float* taps, const float* in,
                      double border_ratio, float* out /* [ceil(xs/step)][ys] */) {
  double full = 0.0;
  for (int j = 0; j <= 2 * radius; ++j) full += taps[j];
  for (int x = 0, ox = 0; x < xs; x += step, ++ox) {
    const int lo = imax(0, x - radius);
    const int hi = imin(xs, x + radius + 1) - 1;
    double wsum = 0.0;
    for (int j = lo; j <= hi; ++j) wsum += taps[j - x + radius];
    wsum = (1.0 - border_ratio) * wsum + border_ratio * full;
    const double scale = 1.0 / wsum;
    for (int y = 0; y < ys; ++y) {
      double acc = 0.0;
      for (int j = lo; j <= hi; ++j) {
        const float prod = in[(size_t)y * xs + j] * taps[j - x + radius]; /* float product */
        acc += prod;
      }
      out[(size_t)ox * ys + y] = (float)(acc * scale);
    }
  }
}

typedef struct { int radius, step; float taps[2 * 40 + 1]; } gzo_kernel;

static void make_kernel(double sigma, gzo_kernel* k) {
  const double scaler = -1.0 / (2 * sigma * sigma);
  k->radius = imax(1, (int)(2.25 * fabs(sigma)));
  for (int i = -k->radius; i <= k->radius; ++i) k->taps[i + k->radius] = (float)exp(scaler * i * i);
  k->step = imax(1, (int)(sigma / 3));
}

GZO_API void gzo_blur_params(double sigma, int* radius, int* step, float* taps) {
  gzo_kernel k;
  make_kernel(sigma, &k);
  *radius = k.radius;
  *step = k.step;
  memcpy(taps, k.taps, (2 * k.radius + 1) * sizeof(float));
}

GZO_API void gzo_blur(float* plane, int xs, int ys, double sigma, double border_ratio) {
  gzo_kernel k;
  make_kernel(sigma, &k);
  const int dxs = (xs + k.step - 1) / k.step, dys = (ys + k.step - 1) / k.step;
  float* tmp = (float*)malloc(sizeof(float) * (size_t)dxs * ys);
  conv_pass(xs, ys, k.step, k.radius, k.taps, plane, border_ratio, tmp);
  if (k.step == 1) {
    conv_pass(ys, dxs, 1, k.radius, k.taps, tmp, border_ratio, plane);
  } else {
    float* small = (float*)malloc(sizeof(float) * (size_t)dxs * dys);
    conv_pass(ys, dxs, k.step, k.radius, k.taps, tmp, border_ratio, small);
    for (int y = 0; y < ys; ++y)
      for (int x = 0; x < xs; ++x)
        plane[(size_t)y * xs + x] = small[(size_t)(y / k.step) * dxs + x / k.step];
    free(small);
  }
  free(tmp);
}

/* ------------------------------------------------------------------------------------------
 * Opsin dynamics.  butteraugli.cc:741-764 (absorbance), 868-941 (gamma), 283-292 (XYB), 943-974
 * ---------------------------------------------------------------------------------------- */
static const double kAbsorb[12] = {
    0.348036746003, 0.577814843137, 0.0544556093735, 0.774145581713,
    0.26922717275,  0.767247733938, 0.0366922708552, 0.920130265014,
    0.0882062883536, 0.158581714673, 0.712857943858, 10.6524069248};

static void absorbance(const double in[3], double out[3]) {
  for (int r = 0; r < 3; ++r) {
    const double* m = kAbsorb + 4 * r;
    out[r] = m[0] * in[0] + m[1] * in[1] + m[2] * in[2] + m[3];
  }
}

/* Chebyshev series via Clenshaw; the last step has no doubling. */
static double clenshaw6(double x, const double c[6]) {
  double b1 = 0.0, b2 = 0.0;
  for (int i = 5; i >= 1; --i) {
    const double xb = x * b1;
    const double t = (xb + xb) - b2 + c[i];
    b2 = b1;
    b1 = t;
  }
  return x * b1 - b2 + c[0];
}

static double gamma_poly(double v) {
  static const double P[6] = {881.979476556478289, 1496.058452015812463, 908.662212739659481,
                              373.566100223287378, 85.840860336314364,  6.683258861509244};
  static const double Q[6] = {12.262350348616792, 20.557285797683576, 12.161463238367844,
                              4.711532733641639,  0.899112889751053,  0.035662329617191};
  const float xf = (float)v; /* argument passes through float (butteraugli.cc:938-941) */
  const double lo = 0.770000000000000, hi = 274.579999999999984;
  const double x01 = (xf - lo) / (hi - lo);
  const double xc = 2.0 * x01 - 1.0;
  const double yp = clenshaw6(xc, P), yq = clenshaw6(xc, Q);
  if (yq == 0.0) return 0.0;
  return (float)(yp / yq); /* result passes through float */
}

GZO_API void gzo_opsin_dynamics_image(float* planes, int xs, int ys) {
  const size_t n = (size_t)xs * ys;
  float* blurred = (float*)malloc(3 * n * sizeof(float));
  memcpy(blurred, planes, 3 * n * sizeof(float));
  for (int c = 0; c < 3; ++c) gzo_blur(blurred + c * n, xs, ys, 1.1, 0.0);
  for (size_t i = 0; i < n; ++i) {
    double pre[3] = {blurred[i], blurred[n + i], blurred[2 * n + i]};
    double pm[3], sens[3];
    absorbance(pre, pm);
    for (int c = 0; c < 3; ++c) sens[c] = gamma_poly(pm[c]) / pm[c];
    double cur[3] = {planes[i], planes[n + i], planes[2 * n + i]};
    double cm[3];
    absorbance(cur, cm);
    for (int c = 0; c < 3; ++c) cm[c] *= sens[c];
    planes[i] = (float)(1.01611726948 * cm[0] - 0.982482243696 * cm[1]);
    planes[n + i] = (float)(1.43571362627 * cm[0] + 0.896039849412 * cm[1]);
    planes[2 * n + i] = (float)cm[2];
  }
  free(blurred);
}

/* interleaved sRGB8 -> XYB planes.  guetzli/butteraugli_comparator.cc:31-46 */
GZO_API void gzo_srgb_to_xyb(const uint8_t* rgb, int xs, int ys, float* xyb) {
  double lut[256];
  gzo_srgb8_to_linear_table(lut);
  const size_t n = (size_t)xs * ys;
  for (int c = 0; c < 3; ++c)
    for (size_t i = 0; i < n; ++i) xyb[c * n + i] = (float)lut[rgb[3 * i + c]];
  gzo_opsin_dynamics_image(xyb, xs, ys);
}

/* ------------------------------------------------------------------------------------------
 * MaskHighIntensityChange.  butteraugli.cc:791-843
 * ---------------------------------------------------------------------------------------- */
GZO_API void gzo_mask_high_intensity_change(const float* c0, const float* c1, int xs, int ys,
                                            float* o0, float* o1) {
  const size_t n = (size_t)xs * ys;
  static const double red[3] = {275.19165240059317, 18599.41286306991, 410.8995306951065};
  for (int y = 0; y < ys; ++y)
    for (int x = 0; x < xs; ++x) {
      const size_t ix = (size_t)y * xs + x;
      double ave[3];
      for (int c = 0; c < 3; ++c) ave[c] = (c0[c * n + ix] + c1[c * n + ix]) * 0.5;
      double worst = -1;
      const long nb[4] = {-1, 1, -(long)xs, (long)xs};
      const int off[4] = {x == 0, x + 1 == xs, y == 0, y + 1 == ys};
      for (int d = 0; d < 4; ++d) {
        if (off[d]) continue;
        const size_t j = ix + nb[d];
        double diff = 0.5 * (c0[n + j] + c1[n + j]) - ave[1];
        diff *= diff;
        if (worst < diff) worst = diff;
      }
      const double chroma = 106.95800948271017 / (ave[1] + 106.95800948271017);
      double mix[3];
      mix[0] = chroma * red[0] / (worst + red[0]);
      mix[1] = red[1] / (worst + red[1]);
      mix[2] = chroma * red[2] / (worst + red[2]);
      for (int c = 0; c < 3; ++c) {
        o0[c * n + ix] = (float)(mix[c] * c0[c * n + ix] + (1 - mix[c]) * ave[c]);
        o1[c * n + ix] = (float)(mix[c] * c1[c * n + ix] + (1 - mix[c]) * ave[c]);
      }
    }
}

/* ------------------------------------------------------------------------------------------
 * Piecewise-linear LUTs.  butteraugli.cc:200-281
 * ---------------------------------------------------------------------------------------- */
static double ramp21(double first, double inc, int i) { /* lut[i] built by repeated addition */
  if (i == 0) return 0.0;
  double v = first;
  for (int k = 2; k <= i; ++k) v += inc;
  return v;
}
static double lut21[3][21];
static int lut21_ready = 0;
static void init_lut21(void) {
  if (lut21_ready) return;
  for (int i = 0; i < 21; ++i) {
    lut21[0][i] = ramp21(11.38708334481672, 14.550189611520716, i); /* HighFreqColorDiffDx */
    lut21[1][i] = ramp21(1.4103373714040413, 0.7084088867024, i);   /* HighFreqColorDiffDy */
    lut21[2][i] = ramp21(5.2511644570349185, 5.2511644570349185, i); /* LowFreqColorDiffDy */
  }
  lut21_ready = 1;
}
static double interp_signed(const double* a, int size, double sx) {
  const double ax = fabs(sx);
  const int b = (int)ax;
  double r;
  if (b >= size - 1) r = a[size - 1];
  else r = a[b] + (ax - b) * (a[b + 1] - a[b]);
  return sx < 0 ? -r : r;
}
static double interp_clamp(const double* a, int size, double sx) {
  if (sx < 0) sx = 0;
  const int b = (int)sx;
  if (b >= size - 1) return a[size - 1];
  return a[b] + (sx - b) * (a[b + 1] - a[b]);
}

/* XybLowFreqToVals + XybDiffLowFreqSquaredAccumulate.  butteraugli.cc:305-350 */
static void lowfreq_vals(double x, double y, double z, double v[3]) {
  z += 0.0812519812628 * y;
  v[2] = z * 7.34905756986;
  v[0] = x * 6.64482198135;
  v[1] = interp_signed(lut21[2], 21, y * 0.837846224276);
}
static void lowfreq_sq_acc(const double a[3], const double b[3], double factor, double res[3]) {
  double va[3], vb[3];
  lowfreq_vals(a[0], a[1], a[2], va);
  if (b[0] == 0.0 && b[1] == 0.0 && b[2] == 0.0) {
    for (int c = 0; c < 3; ++c) res[c] += factor * va[c] * va[c];
    return;
  }
  lowfreq_vals(b[0], b[1], b[2], vb);
  for (int c = 0; c < 3; ++c) {
    const double d = va[c] - vb[c];
    res[c] += factor * d * d;
  }
}

/* ------------------------------------------------------------------------------------------
 * 8x8 FFT power spectrum.  butteraugli.cc:371-597 (DJB split-radix butterflies, restated in
 * closed form; the floating-point DAG -- operand pairing of every add/mul -- is unchanged).
 * ---------------------------------------------------------------------------------------- */
static const double kR = 0.70710678118654752440084436210484903;

typedef struct { double re, im; } cpx;

static void rfft8(const double* x, int stride, cpx F[8]) {
  const double s04 = x[0] + x[4 * stride], d04 = x[0] - x[4 * stride];
  const double s26 = x[2 * stride] + x[6 * stride], d26 = x[2 * stride] - x[6 * stride];
  const double s15 = x[stride] + x[5 * stride], d15 = x[stride] - x[5 * stride];
  const double s37 = x[3 * stride] + x[7 * stride], d37 = x[3 * stride] - x[7 * stride];
  const double a = (d15 - d37) * kR, b = (d15 + d37) * kR;
  const double ev = s04 + s26, od = s15 + s37;
  F[0].re = ev + od;       F[0].im = 0;
  F[4].re = ev - od;       F[4].im = 0;
  F[2].re = s04 - s26;     F[2].im = -(s15 - s37);
  F[6].re = F[2].re;       F[6].im = s15 - s37;
  F[1].re = a + d04;       F[1].im = -(b + d26);
  F[7].re = F[1].re;       F[7].im = b + d26;
  F[3].re = d04 - a;       F[3].im = d26 - b;
  F[5].re = F[3].re;       F[5].im = b - d26;
}

static void cfft8(cpx* a) {
  const double sr04 = a[0].re + a[4].re, dr04 = a[0].re - a[4].re;
  const double si04 = a[0].im + a[4].im, di04 = a[0].im - a[4].im;
  const double sr26 = a[2].re + a[6].re, dr26 = a[2].re - a[6].re;
  const double si26 = a[2].im + a[6].im, di26 = a[2].im - a[6].im;
  const double sr15 = a[1].re + a[5].re, dr15 = a[1].re - a[5].re;
  const double si15 = a[1].im + a[5].im, di15 = a[1].im - a[5].im;
  const double sr37 = a[3].re + a[7].re, dr37 = a[3].re - a[7].re;
  const double si37 = a[3].im + a[7].im, di37 = a[3].im - a[7].im;
  /* odd half */
  const double e4r = dr04 - di26, e4i = di04 + dr26;
  const double e6r = dr04 + di26, e6i = di04 - dr26;
  const double m1 = dr15 - di37, m3 = dr15 + di37;
  const double m2 = di15 - dr37, m4 = di15 + dr37;
  const double P = (m1 - m4) * kR, Q = (m1 + m4) * kR;
  const double U = (m2 - m3) * kR, V = (m2 + m3) * kR;
  cpx o4, o5, o6, o7;
  o5.re = e4r - P; o5.im = e4i - Q;
  o4.re = P + e4r; o4.im = Q + e4i;
  o7.im = e6i - U; o7.re = e6r - V;
  o6.re = V + e6r; o6.im = U + e6i;
  /* even half: 4-point transform of the pair sums */
  const double A = sr04 + sr26, B = sr15 + sr37, C = si04 + si26, D = si15 + si37;
  const double t1 = sr04 - sr26, t3 = sr15 - sr37, t2 = si04 - si26, t4 = si15 - si37;
  cpx o0, o1, o2, o3;
  o0.re = A + B; o0.im = C + D;
  o1.re = A - B; o1.im = C - D;
  o2.re = t1 - t4; o2.im = t2 + t3;
  o3.re = t1 + t4; o3.im = t2 - t3;
  a[0] = o0; a[1] = o6; a[2] = o3; a[3] = o5; a[4] = o1; a[5] = o7; a[6] = o2; a[7] = o4;
}

/* Overwrites block[4..36] with the scaled power of frequency bins (u=i/8 horizontal, v=i%8). */
static void fft_power(double block[64]) {
  cpx rows[8][8];
  for (int y = 0; y < 8; ++y) rfft8(block + 8 * y, 1, rows[y]);
  cpx col[5][8];
  double r0[8], r4[8];
  for (int y = 0; y < 8; ++y) { r0[y] = rows[y][0].re; r4[y] = rows[y][4].re; }
  rfft8(r0, 1, col[0]);
  rfft8(r4, 1, col[4]);
  for (int u = 1; u < 4; ++u) {
    for (int y = 0; y < 8; ++y) col[u][y] = rows[y][u];
    cfft8(col[u]);
  }
  for (int i = 4; i < 37; ++i) {
    const cpx c = col[i / 8][i % 8];
    double p = c.re * c.re + c.im * c.im;
    p *= 0.000064;
    block[i] = p;
  }
}

static const double kCsf8x8[37] = {
    5.28270670524, 0.0, 0.0, 0.0, 0.3831134973, 0.676303603859, 3.58927792424, 18.6104367002,
    18.6104367002, 3.09093131948, 1.0, 0.498250875965, 0.36198671102, 0.308982169883,
    0.1312701920435, 2.37370549629, 3.58927792424, 1.0, 2.37370549629, 0.991205724152,
    1.05178802919, 0.627264168628, 0.4, 0.1312701920435, 0.676303603859, 0.498250875965,
    0.991205724152, 0.5, 0.3831134973, 0.349686450518, 0.627264168628, 0.308982169883,
    0.3831134973, 0.36198671102, 1.05178802919, 0.3831134973, 0.12};

static double drop_deadzone(double v, double range) {
  if (v >= -range && v < range) return 0;
  return v < 0 ? v + range : v - range;
}

/* ButteraugliBlockDiff.  butteraugli.cc:602-684.  a,b: [3][64] doubles, destroyed. */
static void block_diff(double* a, double* b, double dc[3], double ac[3], double edge[3]) {
  init_lut21();
  double mean[3] = {0, 0, 0};
  double em[3][4] = {{0}};
  for (int i = 0; i < 192; ++i) {
    const double d = a[i] - b[i];
    const int c = i / 64, k = i % 64, kx = k % 8, ky = k / 8;
    mean[c] += d / 64;
    if (ky == 0) em[c][1] += d / 8;
    else if (ky == 7) em[c][3] += d / 8;
    if (kx == 0) em[c][0] += d / 8;
    else if (kx == 7) em[c][2] += d / 8;
  }
  const double zero[3] = {0, 0, 0};
  lowfreq_sq_acc(mean, zero, kCsf8x8[0], dc);
  for (int e = 0; e < 4; ++e) {
    const double v[3] = {em[0][e], em[1][e], em[2][e]};
    lowfreq_sq_acc(v, zero, kCsf8x8[0], edge);
  }
  for (int i = 0; i < 192; ++i) {
    const double avg = (a[i] + b[i]) / 2, hd = (a[i] - b[i]) / 2;
    a[i] = avg;
    b[i] = hd;
  }
  double* y_avg = a + 64;
  double *x_hd = b, *y_hd = b + 64, *z_hd = b + 128;
  fft_power(y_avg);
  fft_power(x_hd);
  fft_power(y_hd);
  fft_power(z_hd);
  for (int i = 4; i < 37; ++i) {
    const double d = kCsf8x8[i];
    ac[0] += d * 64.8 * x_hd[i];
    ac[2] += d * 2.4 * z_hd[i];
    const double ya = sqrt(y_avg[i]), yh = sqrt(y_hd[i]);
    const double y0 = drop_deadzone(ya - yh, 0.04), y1 = drop_deadzone(ya + yh, 0.04);
    if (y0 != y1) {
      const double v0 = interp_signed(lut21[1], 21, y0 * 1.51983458269);
      const double v1 = interp_signed(lut21[1], 21, y1 * 1.51983458269);
      const double vy = 1.753123908348329 * (v0 - v1);
      ac[1] += d * vy * vy;
    }
  }
}

GZO_API void gzo_block_diff(const double* b0, const double* b1, double* dc, double* ac,
                            double* edge) {
  double a[192], b[192];
  memcpy(a, b0, sizeof(a));
  memcpy(b, b1, sizeof(b));
  for (int c = 0; c < 3; ++c) dc[c] = ac[c] = edge[c] = 0.0;
  block_diff(a, b, dc, ac, edge);
}

/* ------------------------------------------------------------------------------------------
 * Res-grid stages (step 3).  butteraugli.cc:689-738, 1081-1231
 * ---------------------------------------------------------------------------------------- */
enum { STEP = 3 };

static void edge_detector_map(const float* xyb0, const float* xyb1, int xs, int ys, float* out) {
  init_lut21();
  const size_t n = (size_t)xs * ys;
  const int rxs = (xs + STEP - 1) / STEP;
  static const double sig[3] = {1.5, 0.586, 0.4};
  float* bl0 = (float*)malloc(3 * n * sizeof(float));
  float* bl1 = (float*)malloc(3 * n * sizeof(float));
  memcpy(bl0, xyb0, 3 * n * sizeof(float));
  memcpy(bl1, xyb1, 3 * n * sizeof(float));
  for (int c = 0; c < 3; ++c) {
    gzo_blur(bl0 + c * n, xs, ys, sig[c], 0.0);
    gzo_blur(bl1 + c * n, xs, ys, sig[c], 0.0);
  }
  const double w = 0.711100840192;
  for (int ry = 0; ry + (8 - STEP) < ys; ry += STEP)
    for (int rx = 0; rx + (8 - STEP) < xs; rx += STEP) {
      const size_t rix = ((size_t)ry * rxs + rx) / STEP;
      const int px = imin(rx, xs - 8), py = imin(ry, ys - 8);
      double acc[3] = {0, 0, 0};
      int count = 0;
      for (int k = 0; k < 4; ++k) {
        const int x = px + (k >= 2 ? 7 : 0), y = py + ((k & 1) ? 7 : 0);
        for (int dir = 0; dir < 2; ++dir) {
          size_t i1, i2;
          if (dir == 0) {
            if (!(x >= 3 && x + 3 < xs)) continue;
            i1 = (size_t)y * xs + (x - 3);
            i2 = i1 + 6;
          } else {
            if (!(y >= 3 && y + 3 < ys)) continue;
            i1 = (size_t)(y - 3) * xs + x;
            i2 = i1 + 6 * (size_t)xs;
          }
          double d0[3], d1[3];
          for (int c = 0; c < 3; ++c) {
            d0[c] = w * (bl0[c * n + i1] - bl0[c * n + i2]);
            d1[c] = w * (bl1[c * n + i1] - bl1[c * n + i2]);
          }
          lowfreq_sq_acc(d0, d1, 1.0, acc);
          ++count;
        }
      }
      const double mul = 0.01617112696 * 8.0 / count;
      for (int c = 0; c < 3; ++c) {
        double v = 0.0;
        v += mul * acc[c];
        out[3 * rix + c] = (float)v;
      }
    }
  free(bl0);
  free(bl1);
}

static void block_diff_map(const float* xyb0, const float* xyb1, int xs, int ys, float* dc_out,
                           float* ac_out) {
  const size_t n = (size_t)xs * ys;
  const int rxs = (xs + STEP - 1) / STEP;
  for (int ry = 0; ry + (8 - STEP - 1) < ys; ry += STEP)
    for (int rx = 0; rx + (8 - STEP - 1) < xs; rx += STEP) {
      const size_t rix = ((size_t)ry * rxs + rx) / STEP;
      const size_t off = (size_t)imin(ry, ys - 8) * xs + imin(rx, xs - 8);
      double a[192], b[192];
      for (int c = 0; c < 3; ++c)
        for (int y = 0; y < 8; ++y)
          for (int x = 0; x < 8; ++x) {
            a[64 * c + 8 * y + x] = xyb0[c * n + off + (size_t)y * xs + x];
            b[64 * c + 8 * y + x] = xyb1[c * n + off + (size_t)y * xs + x];
          }
      double dc[3] = {0}, ac[3] = {0}, edge[3] = {0};
      block_diff(a, b, dc, ac, edge);
      for (int c = 0; c < 3; ++c) {
        dc_out[3 * rix + c] = (float)dc[c];
        ac_out[3 * rix + c] = (float)ac[c];
      }
    }
}

static void edge_detector_lowfreq(const float* xyb0, const float* xyb1, int xs, int ys,
                                  float* ac_io) {
  init_lut21();
  const size_t n = (size_t)xs * ys;
  const int rxs = (xs + STEP - 1) / STEP;
  float* bl0 = (float*)malloc(3 * n * sizeof(float));
  float* bl1 = (float*)malloc(3 * n * sizeof(float));
  memcpy(bl0, xyb0, 3 * n * sizeof(float));
  memcpy(bl1, xyb1, 3 * n * sizeof(float));
  for (int c = 0; c < 3; ++c) {
    gzo_blur(bl0 + c * n, xs, ys, 14.0, 0.0);
    gzo_blur(bl1 + c * n, xs, ys, 14.0, 0.0);
  }
  const double zero[3] = {0, 0, 0};
  for (int y = 0; y + 8 < ys; y += STEP) {
    const int resy = y / STEP;
    int resx = 8 / STEP;
    for (int x = 0; x + 8 < xs; x += STEP, ++resx) {
      const long ix = (long)y * xs + x;
      const long rix = (long)resy * rxs + resx;
      const long far[4] = {ix + 8, ix + 8L * xs, ix + 6L * xs + 6, ix + 6L * xs - 6};
      double best[3] = {0, 0, 0};
      for (int k = 0; k < 4; ++k) {
        double d[3];
        for (int c = 0; c < 3; ++c) {
          if (k == 3 && x < 8) { d[c] = 0; continue; }
          /* float differences, float sum, widened on assignment */
          d[c] = (bl1[c * n + ix] - bl0[c * n + ix]) + (bl0[c * n + far[k]] - bl1[c * n + far[k]]);
        }
        double sq[3] = {0, 0, 0};
        lowfreq_sq_acc(d, zero, 1.0, sq);
        for (int c = 0; c < 3; ++c) best[c] = best[c] > sq[c] ? best[c] : sq[c];
      }
      for (int c = 0; c < 3; ++c) ac_io[3 * rix + c] += (float)(10 * best[c]);
    }
  }
  free(bl0);
  free(bl1);
}

/* ------------------------------------------------------------------------------------------
 * Mask.  butteraugli.cc:1242-1567
 * ---------------------------------------------------------------------------------------- */
static double mask_lut[6][512];
static int mask_ready = 0;
static void init_mask_lut(void) {
  if (mask_ready) return;
  /* {extmul, extoff, offset, scaler, mul} for MaskX, MaskY, MaskB, MaskDcX, MaskDcY, MaskDcB */
  static const double P[6][5] = {
      {0.975741017749, -4.25328244168, 0.454909521427, 0.0738288224836, 20.8029176447},
      {0.373995618954, 1.5307267433, 0.911952641929, 1.1731667845, 16.2447033988},
      {0.61582234137, -4.25376118646, 1.05105070921, 0.47434643535, 31.1444967089},
      {1.79116943438, -3.86797479189, 0.670960225853, 0.486575865525, 20.4563479139},
      {0.212223514236, -3.65647120524, 1.73396799447, 0.170392660501, 21.6566724788},
      {0.349376011816, -0.894711072781, 0.901647926679, 0.380086095024, 18.0373825149}};
  for (int t = 0; t < 6; ++t)
    for (int i = 0; i < 512; ++i) {
      const double c = P[t][4] / ((0.01 * P[t][3] * i) + P[t][2]);
      double v = 1.0 + P[t][0] * (c + P[t][1]);
      v *= v;
      mask_lut[t][i] = v;
    }
  mask_ready = 1;
}
GZO_API void gzo_mask_luts(double* out /* [6][512] */) {
  init_mask_lut();
  memcpy(out, mask_lut, sizeof(mask_lut));
}

static void xyb_to_vals(double x, double y, double z, double v[3]) { /* butteraugli.cc:294-302 */
  v[0] = interp_signed(lut21[0], 21, x * 0.758304045695);
  v[1] = interp_signed(lut21[1], 21, y * 2.28148649801);
  v[2] = 1.87816926918 * z;
}

GZO_API void gzo_diff_precompute(const float* xyb0, const float* xyb1, int xs, int ys,
                                 float* out) {
  init_lut21();
  const size_t n = (size_t)xs * ys;
  for (int y = 0; y < ys; ++y)
    for (int x = 0; x < xs; ++x) {
      const size_t ix = (size_t)y * xs + x;
      const size_t ih = x + 1 < xs ? ix + 1 : ix - 1;
      const size_t iv = y + 1 < ys ? ix + xs : ix - xs;
      double h0[3], h1[3], v0[3], v1[3];
      xyb_to_vals(xyb0[ix] - xyb0[ih], xyb0[n + ix] - xyb0[n + ih],
                  xyb0[2 * n + ix] - xyb0[2 * n + ih], h0);
      xyb_to_vals(xyb1[ix] - xyb1[ih], xyb1[n + ix] - xyb1[n + ih],
                  xyb1[2 * n + ix] - xyb1[2 * n + ih], h1);
      xyb_to_vals(xyb0[ix] - xyb0[iv], xyb0[n + ix] - xyb0[n + iv],
                  xyb0[2 * n + ix] - xyb0[2 * n + iv], v0);
      xyb_to_vals(xyb1[ix] - xyb1[iv], xyb1[n + ix] - xyb1[n + iv],
                  xyb1[2 * n + ix] - xyb1[2 * n + iv], v1);
      for (int c = 0; c < 3; ++c) {
        const double s0 = fabs(h0[c]) + fabs(v0[c]), s1 = fabs(h1[c]) + fabs(v1[c]);
        out[c * n + ix] = (float)(s0 < s1 ? s0 : s1); /* std::min(sup0, sup1) */
      }
    }
}

/* _Average5x5 in gather form with the scatter code's float summation order
 * (butteraugli.cc:1379-1438; order derived in SURVEY.md Appendix A). */
GZO_API void gzo_average5x5(float* plane, int xs, int ys) {
  if (xs < 4 || ys < 4) return;
  const float w = 0.679144890667f;
  const float scale = 1.0f / (5.0f + 4 * w);
  const size_t n = (size_t)xs * ys;
  float* src = (float*)malloc(n * sizeof(float));
  memcpy(src, plane, n * sizeof(float));
#define SRC(xx, yy) src[(size_t)(yy) * xs + (xx)]
  for (int y = 0; y < ys; ++y)
    for (int x = 0; x < xs; ++x) {
      float acc = SRC(x, y);
      if (y > 0) {
        if (x > 0) acc += SRC(x - 1, y - 1) * w;
        acc += SRC(x, y - 1);
        if (x + 1 < xs) acc += SRC(x + 1, y - 1) * w;
      }
      if (x > 0) acc += SRC(x - 1, y);
      if (x + 1 < xs) acc += SRC(x + 1, y);
      if (y + 1 < ys) {
        if (x > 0) acc += SRC(x - 1, y + 1) * w;
        acc += SRC(x, y + 1);
        if (x + 1 < xs) acc += SRC(x + 1, y + 1) * w;
      }
      plane[(size_t)y * xs + x] = acc * scale;
    }
#undef SRC
  free(src);
}

/* _MinSquareVal(4, 0): forward 4x4 minimum, columns then rows.  butteraugli.cc:1332-1376 */
GZO_API void gzo_min_square_val(float* plane, int xs, int ys, int square, int offset) {
  const size_t n = (size_t)xs * ys;
  float* tmp = (float*)malloc(n * sizeof(float));
  for (int y = 0; y < ys; ++y) {
    const int lo = imax(0, y - offset), hi = imin(ys, y + square - offset);
    for (int x = 0; x < xs; ++x) {
      double m = plane[(size_t)lo * xs + x];
      for (int j = lo + 1; j < hi; ++j) m = fmin(m, plane[(size_t)j * xs + x]);
      tmp[(size_t)y * xs + x] = (float)m;
    }
  }
  for (int x = 0; x < xs; ++x) {
    const int lo = imax(0, x - offset), hi = imin(xs, x + square - offset);
    for (int y = 0; y < ys; ++y) {
      double m = tmp[(size_t)y * xs + lo];
      for (int j = lo + 1; j < hi; ++j) m = fmin(m, tmp[(size_t)y * xs + j]);
      plane[(size_t)y * xs + x] = (float)m;
    }
  }
  free(tmp);
}

GZO_API void gzo_mask(const float* xyb0, const float* xyb1, int xs, int ys, float* mask,
                      float* mask_dc) {
  init_mask_lut();
  const size_t n = (size_t)xs * ys;
  static const double sig[3] = {9.65781083553, 14.2644604355, 4.53358927369};
  static const double wmul[3] = {232.206464018, 22.9455222245, 503.962310606};
  static const double kGlobalScale = 1.0 / 14.921561160295326;
  const float gs = (float)(kGlobalScale * kGlobalScale);
  gzo_diff_precompute(xyb0, xyb1, xs, ys, mask);
  for (int c = 0; c < 3; ++c) {
    gzo_average5x5(mask + c * n, xs, ys);
    gzo_min_square_val(mask + c * n, xs, ys, 4, 0);
    gzo_blur(mask + c * n, xs, ys, sig[c], 0.0);
  }
  for (size_t i = 0; i < n; ++i)
    for (int c = 0; c < 3; ++c) {
      const double p = wmul[c] * (double)mask[c * n + i];
      float m = (float)interp_clamp(mask_lut[c], 512, p);
      float d = (float)interp_clamp(mask_lut[3 + c], 512, p);
      m *= gs;
      d *= gs;
      mask[c * n + i] = m;
      if (mask_dc) mask_dc[c * n + i] = d;
    }
}

/* ------------------------------------------------------------------------------------------
 * CombineChannels + CalculateDiffmap + score.  butteraugli.cc:985-1044, 1207-1240
 * ---------------------------------------------------------------------------------------- */
static void combine_channels(const float* mask, const float* mask_dc, const float* dc,
                             const float* ac, const float* edm, int xs, int ys, float* res) {
  const size_t n = (size_t)xs * ys;
  const int rxs = (xs + STEP - 1) / STEP;
  for (int ry = 0; ry + (8 - STEP) < ys; ry += STEP)
    for (int rx = 0; rx + (8 - STEP) < xs; rx += STEP) {
      const size_t rix = ((size_t)ry * rxs + rx) / STEP;
      const size_t pix = (size_t)(ry + 3) * xs + (rx + 3);
      double m[3], mdc[3];
      for (int c = 0; c < 3; ++c) { m[c] = mask[c * n + pix]; mdc[c] = mask_dc[c * n + pix]; }
      const float* pdc = dc + 3 * rix; const float* pac = ac + 3 * rix; const float* ped = edm + 3 * rix;
      const double a = pdc[0] * mdc[0] + pdc[1] * mdc[1] + pdc[2] * mdc[2];
      const double b = pac[0] * m[0] + pac[1] * m[1] + pac[2] * m[2];
      const double e = ped[0] * m[0] + ped[1] * m[1] + ped[2] * m[2];
      res[rix] = (float)(a + b + e);
    }
}

GZO_API void gzo_calculate_diffmap(const float* res, int xs, int ys, float* diffmap) {
  const size_t n = (size_t)xs * ys;
  const int rxs = (xs + STEP - 1) / STEP;
  const int s2 = (8 - STEP) / 2, s = 8 - STEP;
  memset(diffmap, 0, n * sizeof(float));
  for (int ry = 0; ry + s < ys; ry += STEP)
    for (int rx = 0; rx + s < xs; rx += STEP) {
      const float o = res[((size_t)ry * rxs + rx) / STEP];
      const double v = o < (1.0 / (100.0f * 100.0f)) ? 100.0f * o : sqrt(o);
      for (int dy = 0; dy < STEP; ++dy)
        for (int dx = 0; dx < STEP; ++dx)
          diffmap[(size_t)(ry + dy + s2) * xs + rx + dx + s2] = (float)v;
    }
  const int cw = xs - s, ch = ys - s;
  float* crop = (float*)malloc((size_t)cw * ch * sizeof(float));
  for (int y = 0; y < ch; ++y)
    for (int x = 0; x < cw; ++x) crop[(size_t)y * cw + x] = diffmap[(size_t)(y + s2) * xs + x + s2];
  gzo_blur(crop, cw, ch, 8.8510880283, 0.03027655136);
  const float mul1 = (float)24.8235314874;
  for (int y = 0; y < ch; ++y)
    for (int x = 0; x < cw; ++x)
      diffmap[(size_t)(y + s2) * xs + x + s2] += mul1 * crop[(size_t)y * cw + x];
  const float sc = (float)(1.0 / (1.0 + 24.8235314874));
  for (size_t i = 0; i < n; ++i) diffmap[i] *= sc;
  free(crop);
}

/* DiffmapOpsinDynamicsImage with optional stage outputs (NULL to skip).  butteraugli.cc:1046-1079 */
GZO_API void gzo_diffmap_stages(const float* xyb0_in, const float* xyb1_in, int xs, int ys,
                                float* mhic0, float* mhic1, float* edge_map, float* block_dc,
                                float* block_ac_pre, float* block_ac, float* mask_out,
                                float* mask_dc_out, float* combined, float* diffmap) {
  const size_t n = (size_t)xs * ys;
  const int rxs = (xs + STEP - 1) / STEP, rys = (ys + STEP - 1) / STEP;
  const size_t rn = (size_t)rxs * rys;
  float* x0 = (float*)malloc(3 * n * sizeof(float));
  float* x1 = (float*)malloc(3 * n * sizeof(float));
  gzo_mask_high_intensity_change(xyb0_in, xyb1_in, xs, ys, x0, x1);
  if (mhic0) memcpy(mhic0, x0, 3 * n * sizeof(float));
  if (mhic1) memcpy(mhic1, x1, 3 * n * sizeof(float));
  float* edm = (float*)calloc(3 * rn, sizeof(float));
  float* dc = (float*)calloc(3 * rn, sizeof(float));
  float* ac = (float*)calloc(3 * rn, sizeof(float));
  edge_detector_map(x0, x1, xs, ys, edm);
  if (edge_map) memcpy(edge_map, edm, 3 * rn * sizeof(float));
  block_diff_map(x0, x1, xs, ys, dc, ac);
  if (block_dc) memcpy(block_dc, dc, 3 * rn * sizeof(float));
  if (block_ac_pre) memcpy(block_ac_pre, ac, 3 * rn * sizeof(float));
  edge_detector_lowfreq(x0, x1, xs, ys, ac);
  if (block_ac) memcpy(block_ac, ac, 3 * rn * sizeof(float));
  float* m = (float*)malloc(3 * n * sizeof(float));
  float* mdc = (float*)malloc(3 * n * sizeof(float));
  gzo_mask(x0, x1, xs, ys, m, mdc);
  if (mask_out) memcpy(mask_out, m, 3 * n * sizeof(float));
  if (mask_dc_out) memcpy(mask_dc_out, mdc, 3 * n * sizeof(float));
  float* res = (float*)calloc(rn, sizeof(float));
  combine_channels(m, mdc, dc, ac, edm, xs, ys, res);
  if (combined) memcpy(combined, res, rn * sizeof(float));
  if (diffmap) gzo_calculate_diffmap(res, xs, ys, diffmap);
  free(x0); free(x1); free(edm); free(dc); free(ac); free(m); free(mdc); free(res);
}

GZO_API void gzo_diffmap(const float* xyb0, const float* xyb1, int xs, int ys, float* diffmap) {
  gzo_diffmap_stages(xyb0, xyb1, xs, ys, 0, 0, 0, 0, 0, 0, 0, 0, 0, diffmap);
}

GZO_API float gzo_score_from_diffmap(const float* d, size_t n) {
  float r = 0.0f;
  for (size_t i = 0; i < n; ++i) r = r > d[i] ? r : d[i];
  return r;
}

/* ------------------------------------------------------------------------------------------
 * Integer IDCT, quantiser, YCbCr->RGB.  guetzli/idct.cc:29-161, quantize.h:24-29,
 * color_transform.h:22-219 (libjpeg 16.16 fixed-point tables, regenerated from their formula)
 * ---------------------------------------------------------------------------------------- */
static const int kIdctBasis[64] = {
    8192, 11363, 10703, 9633,   8192,  6437,   4433,   2260,   8192, 9633,   4433,   -2259, -8192,
    -11362, -10704, -6436,      8192,  6437,   -4433,  -11362, -8192, 2261,  10704,  9633,  8192,
    2260,  -10703, -6436, 8192, 9633,  -4433,  -11363, 8192,   -2260, -10703, 6436,  8192,  -9633,
    -4433, 11363,  8192,  -6437, -4433, 11362, -8192,  -2261,  10704, -9633, 8192,   -9633, 4433,
    2259,  -8192,  11362, -10704, 6436, 8192,  -11363, 10703,  -9633, 8192,  -6437,  4433,  -2260};

static void idct_1d(const int16_t* in, int stride, int out[8]) {
  for (int x = 0; x < 8; ++x) {
    int acc = 0;
    for (int u = 0; u < 8; ++u) acc += kIdctBasis[8 * x + u] * in[u * stride];
    out[x] = acc;
  }
}

GZO_API void gzo_idct(const int16_t* block, uint8_t* out) {
  int16_t cols[64];
  for (int x = 0; x < 8; ++x) {
    int v[8];
    idct_1d(block + x, 8, v);
    for (int y = 0; y < 8; ++y) cols[8 * y + x] = (int16_t)((v[y] + (1 << 10)) >> 11);
  }
  for (int y = 0; y < 8; ++y) {
    int v[8];
    idct_1d(cols + 8 * y, 1, v);
    for (int x = 0; x < 8; ++x) {
      const int p = (v[x] + (257 << 17)) >> 18;
      out[8 * y + x] = (uint8_t)imax(0, imin(255, p));
    }
  }
}

GZO_API int gzo_quantize(int coeff, int q) {
  const int r = coeff % q;
  const int delta = 2 * r > q ? q - r : (-2) * r > q ? -q - r : -r;
  return (int16_t)(coeff + delta);
}

static int clamp255(int v) { return v < 0 ? 0 : v > 255 ? 255 : v; }
GZO_API void gzo_ycbcr_to_rgb(uint8_t* px, int npix) {
  for (int i = 0; i < npix; ++i, px += 3) {
    const int y = px[0], cb = px[1] - 128, cr = px[2] - 128;
    const int r = y + ((91881 * cr + 32768) >> 16);
    const int g = y + ((-46802 * cr + (-22554 * cb + 32768)) >> 16);
    const int b = y + ((116130 * cb + 32768) >> 16);
    px[0] = (uint8_t)clamp255(r);
    px[1] = (uint8_t)clamp255(g);
    px[2] = (uint8_t)clamp255(b);
  }
}

/* 4:4:4 candidate image: block-major coefficient planes -> interleaved sRGB8.
 * guetzli/output_image.cc:124-146 (pixel = idct<<4), 68-98 ((p+8-(x&1))>>4 == idct), 642-652 */
GZO_API void gzo_coeffs_to_srgb(const int16_t* c0, const int16_t* c1, const int16_t* c2, int xs,
                                int ys, uint8_t* rgb) {
  const int bw = (xs + 7) / 8, bh = (ys + 7) / 8;
  const int16_t* cc[3] = {c0, c1, c2};
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      uint8_t px[3][64];
      for (int c = 0; c < 3; ++c) gzo_idct(cc[c] + ((size_t)by * bw + bx) * 64, px[c]);
      for (int iy = 0; iy < 8; ++iy)
        for (int ix = 0; ix < 8; ++ix) {
          const int x = 8 * bx + ix, y = 8 * by + iy;
          if (x >= xs || y >= ys) continue;
          uint8_t* o = rgb + 3 * ((size_t)y * xs + x);
          o[0] = px[0][8 * iy + ix]; o[1] = px[1][8 * iy + ix]; o[2] = px[2][8 * iy + ix];
          gzo_ycbcr_to_rgb(o, 1);
        }
    }
}

/* ApplyGlobalQuantization on block-major planes.  output_image.cc:349-360 */
GZO_API void gzo_apply_global_quant(int16_t* coeffs, size_t nblocks, const int* q64) {
  for (size_t b = 0; b < nblocks; ++b)
    for (int k = 0; k < 64; ++k) coeffs[b * 64 + k] = (int16_t)gzo_quantize(coeffs[b * 64 + k], q64[k]);
}

/* ButteraugliComparator::Compare for a 4:4:4 candidate.  guetzli/butteraugli_comparator.cc:60-70 */
GZO_API float gzo_compare(const uint8_t* rgb_orig, const int16_t* c0, const int16_t* c1,
                          const int16_t* c2, int xs, int ys, float* distmap_out) {
  const size_t n = (size_t)xs * ys;
  float* xyb0 = (float*)malloc(3 * n * sizeof(float));
  float* xyb1 = (float*)malloc(3 * n * sizeof(float));
  uint8_t* cand = (uint8_t*)malloc(3 * n);
  float* dm = (float*)malloc(n * sizeof(float));
  gzo_srgb_to_xyb(rgb_orig, xs, ys, xyb0);
  gzo_coeffs_to_srgb(c0, c1, c2, xs, ys, cand);
  gzo_srgb_to_xyb(cand, xs, ys, xyb1);
  gzo_diffmap(xyb0, xyb1, xs, ys, dm);
  const float dist = gzo_score_from_diffmap(dm, n);
  if (distmap_out) memcpy(distmap_out, dm, n * sizeof(float));
  free(xyb0); free(xyb1); free(cand); free(dm);
  return dist;
}

/* ------------------------------------------------------------------------------------------
 * Block comparisons and the greedy zeroing order (4:4:4, factor 1).
 * guetzli/butteraugli_comparator.cc:72-163, guetzli/processor.cc:376-487
 * ---------------------------------------------------------------------------------------- */
/* StartBlockComparisons: mask_xyz_ = Mask(opsin(orig), opsin(orig)) (3*xs*ys floats). */
GZO_API void gzo_block_mask(const uint8_t* rgb_orig, int xs, int ys, float* mask_xyz) {
  const size_t n = (size_t)xs * ys;
  float* xyb = (float*)malloc(3 * n * sizeof(float));
  gzo_srgb_to_xyb(rgb_orig, xs, ys, xyb);
  gzo_mask(xyb, xyb, xs, ys, mask_xyz, 0);
  free(xyb);
}

/* SwitchBlock: opsin dynamics of the edge-clamped original 8x8 window (block-local blur). */
GZO_API void gzo_block_pregamma(const uint8_t* rgb_orig, int xs, int ys, int bx, int by,
                                float* out192) {
  double lut[256];
  gzo_srgb8_to_linear_table(lut);
  for (int iy = 0; iy < 8; ++iy)
    for (int ix = 0; ix < 8; ++ix) {
      const int x = imin(8 * bx + ix, xs - 1), y = imin(8 * by + iy, ys - 1);
      for (int c = 0; c < 3; ++c)
        out192[64 * c + 8 * iy + ix] = (float)lut[rgb_orig[3 * ((size_t)y * xs + x) + c]];
    }
  gzo_opsin_dynamics_image(out192, 8, 8);
}

/* CompareBlock for candidate coefficients [3][64] of block (bx,by); `pregamma` from
 * gzo_block_pregamma, `scale[3]` = mask_xyz[c][8*by*xs + 8*bx]. */
GZO_API double gzo_compare_block(const int16_t* cand192, int xs, int ys, int bx, int by,
                                 const float* pregamma192, const float* scale3) {
  double lut[256];
  gzo_srgb8_to_linear_table(lut);
  uint8_t px[3][64];
  for (int c = 0; c < 3; ++c) gzo_idct(cand192 + 64 * c, px[c]);
  /* ToPixels window: columns/rows past the image edge replicate the last valid one. */
  const int vx = imin(8, xs - 8 * bx), vy = imin(8, ys - 8 * by);
  float lin[192];
  for (int iy = 0; iy < 8; ++iy)
    for (int ix = 0; ix < 8; ++ix) {
      const int sx = imin(ix, vx - 1), sy = imin(iy, vy - 1);
      uint8_t p[3] = {px[0][8 * sy + sx], px[1][8 * sy + sx], px[2][8 * sy + sx]};
      gzo_ycbcr_to_rgb(p, 1);
      for (int c = 0; c < 3; ++c) lin[64 * c + 8 * iy + ix] = (float)lut[p[c]];
    }
  gzo_opsin_dynamics_image(lin, 8, 8);
  float m0[192], m1[192];
  gzo_mask_high_intensity_change(pregamma192, lin, 8, 8, m0, m1);
  double a[192], b[192];
  for (int i = 0; i < 192; ++i) { a[i] = m0[i]; b[i] = m1[i]; }
  double dc[3] = {0}, ac[3] = {0}, edge[3] = {0};
  block_diff(a, b, dc, ac, edge);
  double diff = 0.0, diff_edge = 0.0;
  for (int c = 0; c < 3; ++c) {
    const double s = scale3[c];
    diff += dc[c] * s;
    diff += ac[c] * s;
    diff_edge += edge[c] * s;
  }
  return sqrt((1 - 0.05) * diff + 0.05 * diff_edge);
}

typedef struct { int idx; float block_err; } gzo_coeff_data;

/* ComputeBlockZeroingOrder for one block.  Returns the number of entries kept; sets *ties to the
 * number of adjacent equal sort keys (the reference's std::sort is unstable on ties). */
static int zeroing_order_block(const int16_t* cur192, const int16_t* orig192, int comp_mask,
                               int xs, int ys, int bx, int by, const float* pregamma,
                               const float* scale3, float limit, gzo_coeff_data* out, int* ties) {
  int idx[192];
  float key[192];
  int n = 0;
  for (int c = 0; c < 3; ++c) {
    if (!(comp_mask & (1 << c))) continue;
    for (int k = 1; k < 64; ++k) {
      const int i = 64 * c + k;
      if (cur192[i] != 0) {
        idx[n] = i;
        key[n] = abs(orig192[i]) * gzo_order_csf[i] + gzo_order_bias[i];
        ++n;
      }
    }
  }
  for (int i = 1; i < n; ++i) { /* stable insertion sort, ascending key */
    const int ti = idx[i];
    const float tk = key[i];
    int j = i - 1;
    while (j >= 0 && key[j] > tk) { idx[j + 1] = idx[j]; key[j + 1] = key[j]; --j; }
    idx[j + 1] = ti;
    key[j + 1] = tk;
  }
  for (int i = 1; i < n; ++i)
    if (key[i] == key[i - 1]) ++*ties;
  int16_t work[192];
  memcpy(work, cur192, sizeof(work));
  int nout = 0;
  while (n > 0) {
    float best_err = 1e17f;
    int best_i = 0;
    for (int i = 0; i < imin(3, n); ++i) {
      int16_t cand[192];
      memcpy(cand, work, sizeof(cand));
      cand[idx[i]] = 0;
      const float err = (float)gzo_compare_block(cand, xs, ys, bx, by, pregamma, scale3);
      float max_err = 0;
      max_err = max_err > err ? max_err : err;
      if (max_err < best_err) { best_err = max_err; best_i = i; }
    }
    work[idx[best_i]] = 0;
    out[nout].idx = idx[best_i];
    out[nout].block_err = best_err;
    ++nout;
    for (int i = best_i; i + 1 < n; ++i) { idx[i] = idx[i + 1]; key[i] = key[i + 1]; }
    --n;
  }
  float min_err = 1e10f;
  for (int i = nout - 1; i >= 0; --i) {
    min_err = min_err < out[i].block_err ? min_err : out[i].block_err;
    out[i].block_err = min_err;
  }
  int keep = 0;
  while (keep < nout && out[keep].block_err <= limit) ++keep;
  for (int i = keep; i < nout; ++i) { out[i].idx = 0; out[i].block_err = 0; }
  return keep;
}

/* The MODE_CPU loop of SelectFrequencyMasking (processor.cc:638-672) over blocks
 * [block_begin, block_end). out: (block_end-block_begin)*192 records, zero-filled tails.
 * mask_xyz from gzo_block_mask. Returns the number of sort-key ties seen. */
GZO_API int gzo_zeroing_order(const uint8_t* rgb_orig, int xs, int ys, const int16_t* orig0,
                              const int16_t* orig1, const int16_t* orig2, const int16_t* cur0,
                              const int16_t* cur1, const int16_t* cur2, const float* mask_xyz,
                              int comp_mask, float limit, int block_begin, int block_end,
                              gzo_coeff_data* out) {
  const int bw = (xs + 7) / 8;
  const size_t n = (size_t)xs * ys;
  const int16_t* orig[3] = {orig0, orig1, orig2};
  const int16_t* cur[3] = {cur0, cur1, cur2};
  int ties = 0;
  for (int b = block_begin; b < block_end; ++b) {
    const int bx = b % bw, by = b / bw;
    int16_t c192[192] = {0}, o192[192] = {0};
    for (int c = 0; c < 3; ++c)
      if (comp_mask & (1 << c)) {
        memcpy(c192 + 64 * c, cur[c] + (size_t)b * 64, 128);
        memcpy(o192 + 64 * c, orig[c] + (size_t)b * 64, 128);
      }
    float pregamma[192], scale[3];
    gzo_block_pregamma(rgb_orig, xs, ys, bx, by, pregamma);
    for (int c = 0; c < 3; ++c) scale[c] = mask_xyz[c * n + (size_t)(8 * by) * xs + 8 * bx];
    gzo_coeff_data* o = out + (size_t)(b - block_begin) * 192;
    memset(o, 0, 192 * sizeof(*o));
    zeroing_order_block(c192, o192, comp_mask, xs, ys, bx, by, pregamma, scale, limit, o, &ties);
  }
  return ties;
}

/* ComputeBlockErrorAdjustmentWeights, factor 1.  guetzli/butteraugli_comparator.cc:169-233 */
GZO_API void gzo_block_weights(const float* distmap, int xs, int ys, float target_distance_f,
                               int direction, int max_block_dist, double target_mul,
                               float* weight_io) {
  const double target = target_distance_f * target_mul;
  const int bw = (xs + 7) / 8, bh = (ys + 7) / 8;
  float* bmax = (float*)malloc(sizeof(float) * (size_t)bw * bh);
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      float m = 0.0f;
      for (int y = 8 * by; y < imin(ys, 8 * by + 8); ++y)
        for (int x = 8 * bx; x < imin(xs, 8 * bx + 8); ++x) {
          const float v = distmap[(size_t)y * xs + x];
          m = m > v ? m : v;
        }
      bmax[by * bw + bx] = m;
    }
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      const int ix = by * bw + bx;
      float local = (float)target;
      const int x0 = imax(0, bx - max_block_dist), y0 = imax(0, by - max_block_dist);
      const int x1 = imin(bw, bx + 1 + max_block_dist), y1 = imin(bh, by + 1 + max_block_dist);
      for (int y = y0; y < y1; ++y)
        for (int x = x0; x < x1; ++x) local = local > bmax[y * bw + x] ? local : bmax[y * bw + x];
      if (direction > 0) {
        if (bmax[ix] <= target && local <= 1.1 * target) weight_io[ix] = 1.0;
      } else {
        if (bmax[ix] <= (1 - 0.5) * target + 0.5 * local) continue;
        for (int y = y0; y < y1; ++y)
          for (int x = x0; x < x1; ++x) {
            const int d = imax(abs(y - by), abs(x - bx));
            const float wv = 1.0f / (d + 1.0f);
            if (weight_io[y * bw + x] < wv) weight_io[y * bw + x] = wv;
          }
      }
    }
  free(bmax);
}

/* ScoreJPEG.  guetzli/score.cc:23-41 */
GZO_API double gzo_score_jpeg(double dist, int size, double target) {
  const double diff = dist - target;
  if (diff <= 0.0) return size;
  const double e = 50 * diff;
  if (e > 10) return 1e30 * exp(10.0) * diff + size;
  return exp(e) * size;
}

/* Double-precision 8x8 DCT/IDCT (reached only on the 4:2:0 path).  guetzli/dct_double.cc:28-85
 * The reference's matrix is 0.5*alpha(u)*cos((2x+1)u*pi/16) written with 10 decimals; it is
 * regenerated here by rounding the formula to 10 decimals (decimal -> nearest double). */
static double dctm[64];
static int dctm_ready = 0;
static void init_dctm(void) {
  if (dctm_ready) return;
  for (int u = 0; u < 8; ++u)
    for (int x = 0; x < 8; ++x) {
      const double v = 0.5 * (u == 0 ? sqrt(0.5) : 1.0) * cos((2 * x + 1) * u * M_PI / 16);
      char buf[64];
      snprintf(buf, sizeof(buf), "%.10f", v);
      dctm[8 * u + x] = strtod(buf, 0);
    }
  dctm_ready = 1;
}
static void dct1d_fwd(const double* in, int stride, double* out) {
  for (int x = 0; x < 8; ++x) {
    double acc = 0.0;
    for (int u = 0; u < 8; ++u) acc += dctm[8 * x + u] * in[u * stride];
    out[x * stride] = acc;
  }
}
static void dct1d_inv(const double* in, int stride, double* out) {
  for (int x = 0; x < 8; ++x) {
    double acc = 0.0;
    for (int u = 0; u < 8; ++u) acc += dctm[8 * u + x] * in[u * stride];
    out[x * stride] = acc;
  }
}
GZO_API void gzo_dct_double(double* block, int inverse) {
  init_dctm();
  double tmp[64];
  for (int x = 0; x < 8; ++x) (inverse ? dct1d_inv : dct1d_fwd)(block + x, 8, tmp + x);
  for (int y = 0; y < 8; ++y) (inverse ? dct1d_inv : dct1d_fwd)(tmp + 8 * y, 1, block + 8 * y);
}

/* ------------------------------------------------------------------------------------------
 * Front end: RGB -> YCbCr (16.16 fixed point) -> integer forward DCT (output
 * scaled by 16) -> q=1 quantisation.  guetzli/jpeg_data_encoder.cc:28-117, guetzli/fdct.cc:28-240
 * The transform's shifts truncate, so the operation order below is the reference's.
 * ---------------------------------------------------------------------------------------- */
static int mul16(int a, int b) { return (a * b) >> 16; }

static void fdct_column(int16_t* v /* stride 8 */) {
  const int i0 = v[0], i1 = v[8], i2 = v[16], i3 = v[24], i4 = v[32], i5 = v[40], i6 = v[48], i7 = v[56];
  int d07 = i0 - i7, s07 = i0 + i7, d25 = i2 - i5, s25 = i2 + i5;
  int d34 = i3 - i4, s34 = i3 + i4, d16 = i1 - i6, s16 = i1 + i6;
  int e0 = s07 - s34, e1 = s07 + s34;   /* even part */
  int e2 = s16 - s25, e3 = s16 + s25;
  e1 <<= 3; e3 <<= 3;
  v[0] = (int16_t)(e1 + e3);
  v[32] = (int16_t)(e1 - e3);
  e0 <<= 3; e2 <<= 3; d34 <<= 3; d07 <<= 3;
  v[16] = (int16_t)(mul16(27146, e2) + e0);
  v[48] = (int16_t)(mul16(27146, e0) - e2);
  d25 <<= 4; d16 <<= 4;
  int r = mul16(d16 + d25, 23170), s = mul16(d16 - d25, 23170);
  int p3 = d34 - s, p1 = d34 + s;      /* BUTTERFLY(m3, m1) */
  int p0 = d07 - r, p2 = d07 + r;      /* BUTTERFLY(m0, m2) */
  int q3 = mul16(p3, -21746) + p3 + 1;
  int q1 = mul16(p1, 13036) + p2 + 1;
  int q4 = mul16(-21746, p0) + p0;
  int q5 = mul16(13036, p2);
  v[8] = (int16_t)q1;
  v[24] = (int16_t)(p0 - q3);
  v[40] = (int16_t)(p3 + q4);
  v[56] = (int16_t)(q5 - p1);
}

static void fdct_row(int16_t* in, const int16_t* t) {
  const int a0 = in[0] + in[7], b0 = in[0] - in[7], a1 = in[1] + in[6], b1 = in[1] - in[6];
  const int a2 = in[2] + in[5], b2 = in[2] - in[5], a3 = in[3] + in[4], b3 = in[3] - in[4];
  const int C1 = t[0], C2 = t[1], C3 = t[2], C4 = t[3], C5 = t[4], C6 = t[5], C7 = t[6];
  const int c0 = a0 + a3, c1 = a0 - a3, c2 = a1 + a2, c3 = a1 - a2;
  in[0] = (int16_t)((C4 * (c0 + c2)) >> 16);
  in[4] = (int16_t)((C4 * (c0 - c2)) >> 16);
  in[2] = (int16_t)((C2 * c1 + C6 * c3) >> 16);
  in[6] = (int16_t)((C6 * c1 - C2 * c3) >> 16);
  in[1] = (int16_t)((C1 * b0 + C3 * b1 + C5 * b2 + C7 * b3) >> 16);
  in[3] = (int16_t)((C3 * b0 - C7 * b1 - C1 * b2 - C5 * b3) >> 16);
  in[5] = (int16_t)((C5 * b0 - C1 * b1 + C7 * b2 + C3 * b3) >> 16);
  in[7] = (int16_t)((C7 * b0 - C5 * b1 + C3 * b2 - C1 * b3) >> 16);
}

GZO_API void gzo_fdct(int16_t* block) {
  static const int16_t T04[7] = {22725, 21407, 19266, 16384, 12873, 8867, 4520};
  static const int16_t T17[7] = {31521, 29692, 26722, 22725, 17855, 12299, 6270};
  static const int16_t T26[7] = {29692, 27969, 25172, 21407, 16819, 11585, 5906};
  static const int16_t T35[7] = {26722, 25172, 22654, 19266, 15137, 10426, 5315};
  const int16_t* rows[8] = {T04, T17, T26, T35, T04, T35, T26, T17};
  for (int x = 0; x < 8; ++x) fdct_column(block + x);
  for (int y = 0; y < 8; ++y) fdct_row(block + 8 * y, rows[y]);
}

/* EncodeRGBToJpeg with the all-ones quantiser: block-major int16 coefficient planes. */
GZO_API void gzo_rgb_to_jpeg_coeffs(const uint8_t* rgb, int xs, int ys, int16_t* c0, int16_t* c1,
                                    int16_t* c2) {
  const int bw = (xs + 7) / 8, bh = (ys + 7) / 8;
  int16_t* out[3] = {c0, c1, c2};
  for (int by = 0; by < bh; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      int16_t blk[192];
      for (int iy = 0; iy < 8; ++iy)
        for (int ix = 0; ix < 8; ++ix) {
          const int y = imin(ys - 1, 8 * by + iy), x = imin(xs - 1, 8 * bx + ix);
          const uint8_t* p = rgb + 3 * ((size_t)y * xs + x);
          const int r = p[0], g = p[1], b = p[2], k = 8 * iy + ix;
          blk[k] = (int16_t)((19595 * r + 38469 * g + 7471 * b - (128 << 16) + 32768) >> 16);
          blk[64 + k] = (int16_t)((-11059 * r - 21709 * g + 32768 * b + 32768 - 1) >> 16);
          blk[128 + k] = (int16_t)((32768 * r - 27439 * g - 5329 * b + 32768 - 1) >> 16);
        }
      for (int c = 0; c < 3; ++c) {
        gzo_fdct(blk + 64 * c);
        int16_t* o = out[c] + ((size_t)by * bw + bx) * 64;
        for (int k = 0; k < 64; ++k) o[k] = (int16_t)((blk[64 * c + k] * 65537 + (0x80 << 12)) >> 20);
      }
    }
}

/* ------------------------------------------------------------------------------------------
 * YUV 4:2:0 branch: Downsample, the factor-2 candidate, zeroing over macro-blocks.
 * ---------------------------------------------------------------------------------------- */
#include "gzoracle_yuv420.inc"
