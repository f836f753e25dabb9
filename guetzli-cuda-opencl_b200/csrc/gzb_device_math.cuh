// gzb_device_math.cuh -- per-pixel / per-block device arithmetic of the butteraugli metric.
//
// Every function follows the MODE_CPU arithmetic of the reference (float storage, double
// intermediates, identical operand pairing) so that results are bit-identical to the CPU path:
// the translation unit is compiled with -fmad=false (no FMA contraction), IEEE div/sqrt.
// Citations: third_party/butteraugli/butteraugli/butteraugli.cc (abbrev. "ba.cc") and guetzli/*.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gzb {

// ---------------------------------------------------------------------------------------------
// Device-resident constant tables (filled once per process by gzb_tables_init()).
// ---------------------------------------------------------------------------------------------
struct DeviceTables {
  double lut21[3][21];    // HighFreqColorDiffDx, HighFreqColorDiffDy, LowFreqColorDiffDy (ba.cc:200-247)
  double mask_lut[6][512];  // MaskX,Y,B,DcX,DcY,DcB (ba.cc:1242-1326)
  float srgb_lin[256];    // float(Srgb8ToLinearTable[i]) (guetzli/gamma_correct.cc:23-33)
};
__device__ DeviceTables g_tab;  // global memory (divergent indexing, L1-cached); single-TU build

__constant__ const double kCsf8x8[37] = {  // ba.cc:157-198
    5.28270670524, 0.0, 0.0, 0.0, 0.3831134973, 0.676303603859, 3.58927792424, 18.6104367002,
    18.6104367002, 3.09093131948, 1.0, 0.498250875965, 0.36198671102, 0.308982169883,
    0.1312701920435, 2.37370549629, 3.58927792424, 1.0, 2.37370549629, 0.991205724152,
    1.05178802919, 0.627264168628, 0.4, 0.1312701920435, 0.676303603859, 0.498250875965,
    0.991205724152, 0.5, 0.3831134973, 0.349686450518, 0.627264168628, 0.308982169883,
    0.3831134973, 0.36198671102, 1.05178802919, 0.3831134973, 0.12};

// Scalar constants of the per-pixel arithmetic, in constant memory and deliberately NOT const-qualified: a
// double-precision literal costs two UMOV instructions every time it is used (the profile of
// k_zeroing_order showed 10 % of its issue slots going there), a constant-bank operand costs none.
struct MathConsts {
  double absorb[3][4];          // OpsinAbsorbance rows: r, g, b weights and bias (ba.cc:741-764)
  double gamma_lo, gamma_range, gamma_range_inv;
  double gamma_p[6], gamma_q[6];   // Chebyshev coefficients in Clenshaw order, the constant term last (ba.cc:868-941)
  double xyb[4];                // RgbToXyb (ba.cc:283-292)
  double mhic[4];               // MaskHighIntensityChange (ba.cc:798-841)
};
__constant__ MathConsts g_mc = {
    {{0.348036746003, 0.577814843137, 0.0544556093735, 0.774145581713},
     {0.26922717275, 0.767247733938, 0.0366922708552, 0.920130265014},
     {0.0882062883536, 0.158581714673, 0.712857943858, 10.6524069248}},
    0.770000000000000, 274.579999999999984 - 0.770000000000000, 1.0 / (274.579999999999984 - 0.770000000000000),
    {6.683258861509244, 85.840860336314364, 373.566100223287378, 908.662212739659481, 1496.058452015812463, 881.979476556478289},
    {0.035662329617191, 0.899112889751053, 4.711532733641639, 12.161463238367844, 20.557285797683576, 12.262350348616792},
    {1.01611726948, 0.982482243696, 1.43571362627, 0.896039849412},
    {106.95800948271017, 275.19165240059317, 18599.41286306991, 410.8995306951065}};

// ---------------------------------------------------------------------------------------------
// Piecewise-linear table lookups (ba.cc:249-281)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double interp_signed21(const double* __restrict__ lut, double sx) {
  const double ax = fabs(sx);
  const int b = static_cast<int>(ax);
  double r;
  if (b >= 20) {
    r = lut[20];
  } else {
    const double lo = lut[b], hi = lut[b + 1];
    r = lo + (ax - b) * (hi - lo);
  }
  return sx < 0 ? -r : r;
}

__device__ __forceinline__ double interp_clamp512(const double* __restrict__ lut, double sx) {
  if (sx < 0) sx = 0;
  const int b = static_cast<int>(sx);
  if (b >= 511) return lut[511];
  const double lo = lut[b], hi = lut[b + 1];
  return lo + (sx - b) * (hi - lo);
}

// ---------------------------------------------------------------------------------------------
// Opsin dynamics (ba.cc:741-764, 868-941, 283-292, 951-973)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void opsin_absorbance(double r, double g, double b, double out[3]) {
#pragma unroll
  for (int c = 0; c < 3; ++c)
    out[c] = g_mc.absorb[c][0] * r + g_mc.absorb[c][1] * g + g_mc.absorb[c][2] * b + g_mc.absorb[c][3];
}

constexpr double kGammaRange = 274.579999999999984 - 0.770000000000000;
constexpr double kGammaRangeInv = 1.0 / kGammaRange;
__host__ __device__ __forceinline__ double div_by_gamma_range(double x) {
  const double q0 = x * kGammaRangeInv;
  const double r = fma(-kGammaRange, q0, x);
  return fma(r, kGammaRangeInv, q0);
}

// Degree-5/5 Chebyshev rational via Clenshaw; argument and result pass through float.
__device__ __forceinline__ double gamma_rational(double v) {
  const float xf = static_cast<float>(v);
  // (xf - 0.77) / (274.58 - 0.77), correctly rounded, without the division routine: with y = RN(1 / c),
  // q0 = RN(x * y) is within an ulp of x / c, r = x - c * q0 is exact in one fused operation, and
  // RN(q0 + r * y) is the correctly rounded quotient (Markstein's theorem for division through a correctly
  // rounded reciprocal). gzb_test_gamma_division checks all floats in [0, 1024] against IEEE division on the host.
  const double xd = static_cast<double>(xf) - g_mc.gamma_lo;
  const double q0 = xd * g_mc.gamma_range_inv;
  const double x01 = fma(fma(-g_mc.gamma_range, q0, xd), g_mc.gamma_range_inv, q0);
  const double x = 2.0 * x01 - 1.0;
  double p1 = 0.0, p2 = 0.0, q1 = 0.0, q2 = 0.0, t, xb;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    xb = x * p1; t = (xb + xb) - p2 + g_mc.gamma_p[k]; p2 = p1; p1 = t;
    xb = x * q1; t = (xb + xb) - q2 + g_mc.gamma_q[k]; q2 = q1; q1 = t;
  }
  const double yp = x * p1 - p2 + g_mc.gamma_p[5];
  const double yq = x * q1 - q2 + g_mc.gamma_q[5];
  if (yq == 0.0) return 0.0;
  return static_cast<double>(static_cast<float>(yp / yq));
}

// One pixel: (blurred linear rgb, sharp linear rgb) -> XYB floats.
__device__ __forceinline__ void opsin_pixel(float br, float bg, float bb, float r, float g,
                                            float b, float& X, float& Y, float& B) {
  double pm[3], cm[3];
  opsin_absorbance(br, bg, bb, pm);
  opsin_absorbance(r, g, b, cm);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double sens = gamma_rational(pm[c]) / pm[c];
    cm[c] *= sens;
  }
  X = static_cast<float>(g_mc.xyb[0] * cm[0] - g_mc.xyb[1] * cm[1]);
  Y = static_cast<float>(g_mc.xyb[2] * cm[0] + g_mc.xyb[3] * cm[1]);
  B = static_cast<float>(cm[2]);
}

// ---------------------------------------------------------------------------------------------
// MaskHighIntensityChange, one pixel (ba.cc:798-841). `worst` = max squared Y-average difference
// to the in-image 4-neighbours (-1 if none).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void mhic_pixel(const float c0[3], const float c1[3], double worst,
                                           float o0[3], float o1[3]) {
  const double ave1 = (c0[1] + c1[1]) * 0.5;  // float sum, widened by the double multiply
  const double chroma = g_mc.mhic[0] / (ave1 + g_mc.mhic[0]);
  double mix[3];
  mix[0] = chroma * g_mc.mhic[1] / (worst + g_mc.mhic[1]);
  mix[1] = g_mc.mhic[2] / (worst + g_mc.mhic[2]);
  mix[2] = chroma * g_mc.mhic[3] / (worst + g_mc.mhic[3]);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double ave = (c0[c] + c1[c]) * 0.5;
    o0[c] = static_cast<float>(mix[c] * c0[c] + (1 - mix[c]) * ave);
    o1[c] = static_cast<float>(mix[c] * c1[c] + (1 - mix[c]) * ave);
  }
}
// Squared difference of the neighbour's Y average to the centre's (ba.cc:817-818).
__device__ __forceinline__ double mhic_sqdiff(float n0, float n1, float c0, float c1) {
  const double ave1 = (c0 + c1) * 0.5;
  double d = 0.5 * (n0 + n1) - ave1;
  return d * d;
}

// ---------------------------------------------------------------------------------------------
// Low-frequency colour metric (ba.cc:305-350)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void lowfreq_vals(double x, double y, double z, double v[3]) {
  z += 0.0812519812628 * y;
  v[2] = z * 7.34905756986;
  v[0] = x * 6.64482198135;
  v[1] = interp_signed21(g_tab.lut21[2], y * 0.837846224276);
}
// res += factor * (vals(a) - vals(b))^2 ; b == 0 short-circuits like the reference.
__device__ __forceinline__ void lowfreq_sq_acc(const double a[3], const double b[3],
                                               double factor, double res[3]) {
  double va[3];
  lowfreq_vals(a[0], a[1], a[2], va);
  if (b[0] == 0.0 && b[1] == 0.0 && b[2] == 0.0) {
#pragma unroll
    for (int c = 0; c < 3; ++c) res[c] += factor * va[c] * va[c];
    return;
  }
  double vb[3];
  lowfreq_vals(b[0], b[1], b[2], vb);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double d = va[c] - vb[c];
    res[c] += factor * d * d;
  }
}
__device__ __forceinline__ void lowfreq_sq_acc0(const double a[3], double factor, double res[3]) {
  double va[3];
  lowfreq_vals(a[0], a[1], a[2], va);
#pragma unroll
  for (int c = 0; c < 3; ++c) res[c] += factor * va[c] * va[c];
}

// XybToVals, one channel at a time (ba.cc:294-302): the three outputs are independent.
__device__ __forceinline__ double highfreq_val(int c, double d) {
  if (c == 0) return interp_signed21(g_tab.lut21[0], d * 0.758304045695);
  if (c == 1) return interp_signed21(g_tab.lut21[1], d * 2.28148649801);
  return 1.87816926918 * d;
}

// ---------------------------------------------------------------------------------------------
// 8-point transforms of the 8x8 power spectrum (ba.cc:371-570), closed form of the split-radix
// butterflies with the reference's operand pairing.
// ---------------------------------------------------------------------------------------------
#define GZB_SQRT_HALF 0.70710678118654752440084436210484903

// Real input x[0..7] -> bins 0..4 (bins 5..7 are conjugates). re[0..4], im[0..4].
__device__ __forceinline__ void rfft8_half(const double x[8], double re[5], double im[5]) {
  const double s04 = x[0] + x[4], d04 = x[0] - x[4];
  const double s26 = x[2] + x[6], d26 = x[2] - x[6];
  const double s15 = x[1] + x[5], d15 = x[1] - x[5];
  const double s37 = x[3] + x[7], d37 = x[3] - x[7];
  const double a = (d15 - d37) * GZB_SQRT_HALF, b = (d15 + d37) * GZB_SQRT_HALF;
  const double ev = s04 + s26, od = s15 + s37;
  re[0] = ev + od;   im[0] = 0.0;
  re[4] = ev - od;   im[4] = 0.0;
  re[2] = s04 - s26; im[2] = -(s15 - s37);
  re[1] = a + d04;   im[1] = -(b + d26);
  re[3] = d04 - a;   im[3] = d26 - b;
}

// Complex 8-point transform in place (natural output order).
__device__ __forceinline__ void cfft8(double re[8], double im[8]) {
  const double sr04 = re[0] + re[4], dr04 = re[0] - re[4];
  const double si04 = im[0] + im[4], di04 = im[0] - im[4];
  const double sr26 = re[2] + re[6], dr26 = re[2] - re[6];
  const double si26 = im[2] + im[6], di26 = im[2] - im[6];
  const double sr15 = re[1] + re[5], dr15 = re[1] - re[5];
  const double si15 = im[1] + im[5], di15 = im[1] - im[5];
  const double sr37 = re[3] + re[7], dr37 = re[3] - re[7];
  const double si37 = im[3] + im[7], di37 = im[3] - im[7];
  const double e4r = dr04 - di26, e4i = di04 + dr26;
  const double e6r = dr04 + di26, e6i = di04 - dr26;
  const double m1 = dr15 - di37, m3 = dr15 + di37;
  const double m2 = di15 - dr37, m4 = di15 + dr37;
  const double P = (m1 - m4) * GZB_SQRT_HALF, Q = (m1 + m4) * GZB_SQRT_HALF;
  const double U = (m2 - m3) * GZB_SQRT_HALF, V = (m2 + m3) * GZB_SQRT_HALF;
  const double A = sr04 + sr26, Bv = sr15 + sr37, C = si04 + si26, D = si15 + si37;
  const double t1 = sr04 - sr26, t3 = sr15 - sr37, t2 = si04 - si26, t4 = si15 - si37;
  re[0] = A + Bv;   im[0] = C + D;
  re[4] = A - Bv;   im[4] = C - D;
  re[6] = t1 - t4;  im[6] = t2 + t3;
  re[2] = t1 + t4;  im[2] = t2 - t3;
  re[3] = e4r - P;  im[3] = e4i - Q;
  re[7] = P + e4r;  im[7] = Q + e4i;
  re[5] = e6r - V;  im[5] = e6i - U;
  re[1] = V + e6r;  im[1] = U + e6i;
}

__device__ __forceinline__ double remove_range_around_zero(double v, double range) {
  if (v >= -range && v < range) return 0;
  return v < 0 ? v + range : v - range;
}

// ---------------------------------------------------------------------------------------------
// ButteraugliBlockDiff, warp-cooperative (ba.cc:602-684).
//   fa, fb : the two 8x8x3 blocks as floats in shared memory, [c*64 + 8*y + x]
//   pl     : per-warp scratch of 4 * kBdPlane doubles in shared memory (planes, then power spectra)
//   spec   : per-warp scratch of kBdSpecDoubles doubles (row spectra, then the per-frequency terms). It
//            MAY ALIAS fa / fb: they are last read in step (1), spec is first written in step (3).
//   csf_a / csf_b : kCsf8x8[4 + lane] and kCsf8x8[36] (hoisted by the caller: loop invariant)
// Returns dc[3], ac[3], edge[3] in ALL lanes. Sequential sums keep the reference's order.
// kWithDcEdge = false skips the mean / edge-mean part (dc and edge are then not written): the
// full-image BlockDiffMap computes its DC term in a lane-per-cell kernel (k_block_dc) where the 64
// strictly ordered additions cost one instruction stream per 32 cells instead of per cell.
// All 32 lanes must call. Ends with __syncwarp(); fa/fb are not modified.
//
// Shared-memory layout (doubles), padded so that the 64-bit accesses of a half-warp fall into
// distinct banks: planes [4][8 rows][9] (stride 72 per plane), row spectra re/im [4][56] with bin u
// at +9u and row r at +r.
// ---------------------------------------------------------------------------------------------
constexpr int kBdPlane = 72, kBdSpec = 56;
constexpr int kBdSpecHalf = 3 * kBdSpec + 44;        // re (or im) of four planes: the last one needs 9*4 + 8 = 44
constexpr int kBdSpecDoubles = 2 * kBdSpecHalf;      // 424
constexpr int kBlockDiffScratchDoubles = 4 * kBdPlane + kBdSpecDoubles;  // 712

template <bool kWithDcEdge = true>
__device__ __forceinline__ void warp_block_diff(const float* fa, const float* fb, double* pl, double* spec,
                                                double csf_a, double csf_b, double dc[3],
                                                double ac[3], double edge[3]) {
  const int lane = threadIdx.x & 31;
  // pl: [4][72]: y_avg, x_halfdiff, y_halfdiff, z_halfdiff
  double* sre = spec;                    // [4][56] (44 used of the last)
  double* sim = spec + kBdSpecHalf;

  // (1) average / half-difference planes (8 values per lane).
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int i = lane + 32 * k, o = (i >> 3) * 9 + (i & 7);
    const double a0 = fa[i], b0 = fb[i], a1 = fa[64 + i], b1 = fb[64 + i], a2 = fa[128 + i],
                 b2 = fb[128 + i];
    pl[o] = (a1 + b1) / 2;
    pl[kBdPlane + o] = (a0 - b0) / 2;
    pl[2 * kBdPlane + o] = (a1 - b1) / 2;
    pl[3 * kBdPlane + o] = (a2 - b2) / 2;
  }
  __syncwarp();
  if (kWithDcEdge) {
  // (2) means and edge means: 15 independent in-order accumulations. The reference adds
  // diff/64 (diff/8 for edges) term by term; half-differences are diff/2 exactly and scaling by a
  // power of two commutes with rounding, so sum(diff/64) == sum(halfdiff)/32 bit for bit.
  double acc = 0.0;
  if (lane < 3) {
    const double* h = pl + (1 + lane) * kBdPlane;
#pragma unroll
    for (int y = 0; y < 8; ++y)
#pragma unroll
      for (int x = 0; x < 8; ++x) acc += h[9 * y + x];
    acc = acc / 32;
  } else if (lane < 15) {
    const int e = lane - 3, c = e >> 2, side = e & 3;
    const double* h = pl + (1 + c) * kBdPlane;
    // side 0: kx==0, 1: ky==0, 2: kx==7, 3: ky==7 (ba.cc:619-626)
    const int base = side == 0 ? 0 : side == 1 ? 0 : side == 2 ? 7 : 63;
    const int stride = (side & 1) ? 1 : 9;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc += h[base + t * stride];
    acc = acc / 4;
  }
  // five low-frequency colour evaluations on lanes 0..4 (0: mean, 1..4: the four edges)
  double sq[3] = {0.0, 0.0, 0.0};
  {
    const int e = lane - 1;
    const int s0 = lane == 0 ? 0 : 3 + e, s1 = lane == 0 ? 1 : 7 + e, s2 = lane == 0 ? 2 : 11 + e;
    const double v0 = __shfl_sync(0xffffffffu, acc, s0 & 31);
    const double v1 = __shfl_sync(0xffffffffu, acc, s1 & 31);
    const double v2 = __shfl_sync(0xffffffffu, acc, s2 & 31);
    if (lane < 5) {
      const double v[3] = {v0, v1, v2};
      lowfreq_sq_acc0(v, kCsf8x8[0], sq);
    }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    dc[c] = __shfl_sync(0xffffffffu, sq[c], 0);
    double ed = 0.0;
#pragma unroll
    for (int e = 1; e <= 4; ++e) ed += __shfl_sync(0xffffffffu, sq[c], e);
    edge[c] = ed;
  }
  }
  // (3) row transforms: lane = plane*8 + row.
  {
    const int plane = lane >> 3, row = lane & 7;
    double x[8], re[5], im[5];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = pl[kBdPlane * plane + 9 * row + k];
    rfft8_half(x, re, im);
#pragma unroll
    for (int u = 0; u < 5; ++u) {
      sre[plane * kBdSpec + 9 * u + row] = re[u];
      sim[plane * kBdSpec + 9 * u + row] = im[u];
    }
  }
  __syncwarp();
  // (4) column transforms: 20 tasks = plane*5 + u; power into pl[plane*72 + 8u + (v ^ u)]: the row of
  // eight is XOR-swizzled by u, otherwise the 20 lanes (plane strides of 72, rows of 8 doubles) store
  // bin v into only two bank pairs.
  // Columns u = 0 and u = 4 of a real input are themselves real (their stored imaginary parts are
  // exactly 0.0); running them through the complex transform gives the same bits as the reference's
  // real transform up to the sign of zeros (x - 0 = x, 0 - x = -x, (-x - y) = -(x + y) are exact), and
  // only re^2 + im^2 is used. One code path for all 20 lanes instead of two divergent ones. Bins the
  // reference does not compute (u = 0: v < 4; u = 4: v > 4) land in slots nobody reads.
  // Tasks of planes 0, 1 run on lanes 0..9 and those of planes 2, 3 on lanes 16..25: a 64-bit access is
  // served per half-warp, and planes p and p + 2 sit an even number of 128-byte lines apart (same banks).
  if ((lane & 15) < 10) {
    const int hl = lane & 15, plane = 2 * (lane >> 4) + (hl >= 5 ? 1 : 0), u = hl >= 5 ? hl - 5 : hl;
    const double* cre = sre + plane * kBdSpec + 9 * u;
    const double* cim = sim + plane * kBdSpec + 9 * u;
    double* dst = pl + kBdPlane * plane + 8 * u;
    double re[8], im[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) { re[k] = cre[k]; im[k] = cim[k]; }
    cfft8(re, im);
#pragma unroll
    for (int v = 0; v < 8; ++v) dst[v ^ u] = (re[v] * re[v] + im[v] * im[v]) * 0.000064;
  }
  __syncwarp();
  // (5) per-frequency terms i = 4..36 (lane L -> i = 4+L; lane 0 also i = 36), then in-order sums.
  double* term = sre;  // [3][33]
  for (int i = 4 + lane; i < 37; i += 32) {
    const double d = i == 36 ? csf_b : csf_a;
    const int pi = i ^ (i >> 3);   // bin i = 8u + v sits at 8u + (v ^ u), u <= 4
    term[i - 4] = d * 64.8 * pl[kBdPlane + pi];
    term[66 + i - 4] = d * 2.4 * pl[3 * kBdPlane + pi];
    const double ya = sqrt(pl[pi]), yh = sqrt(pl[2 * kBdPlane + pi]);
    const double y0 = remove_range_around_zero(ya - yh, 0.04);
    const double y1 = remove_range_around_zero(ya + yh, 0.04);
    double ty = 0.0;
    if (y0 != y1) {
      const double v0 = interp_signed21(g_tab.lut21[1], y0 * 1.51983458269);
      const double v1 = interp_signed21(g_tab.lut21[1], y1 * 1.51983458269);
      const double vy = 1.753123908348329 * (v0 - v1);
      ty = d * vy * vy;
    }
    term[33 + i - 4] = ty;
  }
  __syncwarp();
  double s = 0.0;
  if (lane < 3) {
    const double* t = term + 33 * lane;
#pragma unroll
    for (int i = 0; i < 33; ++i) s += t[i];
  }
#pragma unroll
  for (int c = 0; c < 3; ++c) ac[c] = __shfl_sync(0xffffffffu, s, c);
  __syncwarp();
}

// ---------------------------------------------------------------------------------------------
// Integer 8x8 IDCT (guetzli/idct.cc:29-161) in direct matrix form: exact integer arithmetic, so
// the result is identical to the reference's factored evaluation.
// ---------------------------------------------------------------------------------------------
__constant__ const int kIdctBasis[64] = {
    8192, 11363, 10703, 9633,   8192,  6437,   4433,   2260,   8192, 9633,   4433,   -2259, -8192,
    -11362, -10704, -6436,      8192,  6437,   -4433,  -11362, -8192, 2261,  10704,  9633,  8192,
    2260,  -10703, -6436, 8192, 9633,  -4433,  -11363, 8192,   -2260, -10703, 6436,  8192,  -9633,
    -4433, 11363,  8192,  -6437, -4433, 11362, -8192,  -2261,  10704, -9633, 8192,   -9633, 4433,
    2259,  -8192,  11362, -10704, 6436, 8192,  -11363, 10703,  -9633, 8192,  -6437,  4433,  -2260};

// out[x] = sum_u basis[8x+u] * in[u]
__device__ __forceinline__ void idct_1d(const int in[8], int out[8]) {
#pragma unroll
  for (int x = 0; x < 8; ++x) {
    int acc = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += kIdctBasis[8 * x + u] * in[u];
    out[x] = acc;
  }
}
__device__ __forceinline__ int idct_col_round(int v) {  // >>11 with rounding, kept as coeff_t
  return static_cast<int>(static_cast<int16_t>((v + (1 << 10)) >> 11));
}
__device__ __forceinline__ int idct_row_round(int v) {  // >>18 with rounding, +128, clamp
  const int p = (v + (257 << 17)) >> 18;
  return min(255, max(0, p));
}

// YCbCr -> RGB, libjpeg 16.16 tables regenerated arithmetically (guetzli/color_transform.h:22-219)
__device__ __forceinline__ void ycbcr_to_rgb(int y, int cb, int cr, int& r, int& g, int& b) {
  cb -= 128;
  cr -= 128;
  r = min(255, max(0, y + ((91881 * cr + 32768) >> 16)));
  g = min(255, max(0, y + ((-46802 * cr + (-22554 * cb + 32768)) >> 16)));
  b = min(255, max(0, y + ((116130 * cb + 32768) >> 16)));
}

// Quantize (guetzli/quantize.h:24-29)
__device__ __forceinline__ int quantize_coeff(int raw, int q) {
  const int r = raw % q;
  const int delta = 2 * r > q ? q - r : (-2) * r > q ? -q - r : -r;
  return static_cast<int>(static_cast<int16_t>(raw + delta));
}

}  // namespace gzb
