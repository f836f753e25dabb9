// gzb_zeroing.cuh -- the per-8x8-block greedy coefficient-zeroing search on sm_100a.
//
// One warp owns one 8x8 image block for the whole greedy loop of
// Processor::ComputeBlockZeroingOrder (guetzli/processor.cc:376-487, MODE_CPU semantics: NO early
// break) and evaluates ButteraugliComparator::CompareBlock (guetzli/butteraugli_comparator.cc:
// 113-163) for every trial with all block state in shared memory:
//   candidate coefficients -> integer IDCT (incremental: only the touched column is re-transformed)
//   -> YCbCr->sRGB->linear -> 8x8 block-local blur + opsin dynamics -> MaskHighIntensityChange
//   against the original block -> ButteraugliBlockDiff (warp-cooperative FFTs) -> masked error.
// Blocks are handed out through an atomic counter (work per block varies with its number of
// non-zero coefficients). 4:4:4 only (factor 1, comp_mask selects components).
#pragma once
#include "gzb_device_math.cuh"
#include "gzb_zeroing_model.h"
#include "gzb_exact_sort.h"
#include <cstddef>

namespace gzb {

struct CoeffDataDev { int idx; float block_err; };  // guetzli/processor.h:29-32

__constant__ float c_zero_csf[192];
__constant__ float c_zero_bias[192];
__constant__ double c_scale8[8];   // border scales of the sigma-1.1 blur on an 8-sample line
__constant__ float c_taps11[5];    // sigma-1.1 taps

// Per-warp state, 7888 bytes: with four warps per CTA and 72 registers per thread seven CTAs (28
// warps) are resident per SM. The block-diff row spectra (424 doubles) live on top of bufA / bufB / cx /
// key (+ spec_tail): those are dead by the time ButteraugliBlockDiff transforms its rows (fa / fb are read
// in its first step only; cx is rewritten by every trial; key is only used to order the candidates
// before the greedy loop).
struct ZeroWarpSmem {
  double pl[4 * kBdPlane];   // block-diff planes / power spectra
  float bufA[192];   // linear rgb -> fa (original after MaskHighIntensityChange)   } also the block-diff
  float bufB[192];   // H-pass -> fb (candidate after MaskHighIntensityChange)      } spectra, together
  float cx[192];     // candidate block, opsin dynamics                             } with spec_tail
  float key[192];    // candidate ordering keys                                     }
  double spec_tail[kBdSpecDoubles - 4 * 192 / 2];
  float pg0[192];    // original block, opsin dynamics (SwitchBlock)
  short cf[192];     // processed coefficients
  short colv[192];   // column-pass values of the processed state
  short ccol[8];     // candidate's replacement column
  unsigned char pix[192];   // Y, Cb, Cr pixels of the processed state
  unsigned char cpx[64];    // candidate pixels of the touched component
  unsigned char order[192]; // sorted input order
  unsigned char ent[192];   // unsorted entries
};
static_assert(offsetof(ZeroWarpSmem, spec_tail) == offsetof(ZeroWarpSmem, bufA) + 4 * 192 * sizeof(float), "spectra must be contiguous");
static_assert(offsetof(ZeroWarpSmem, bufA) % 8 == 0, "spectra must be 8-byte aligned");

// 8x8 opsin dynamics of linear rgb in `lin` (smem, [c*64+8y+x]) -> `dst`; `hb` is scratch.
__device__ __forceinline__ void warp_block_opsin(const float* lin, float* hb, float* dst, int lane) {
  // horizontal pass: 192 outputs, 6 per lane
#pragma unroll
  for (int k = 0; k < 6; ++k) {
    const int i = lane + 32 * k, x = i & 7;
    const float* row = lin + (i & ~7);
    double acc = 0.0;
#pragma unroll
    for (int d = -2; d <= 2; ++d) {
      const int j = x + d;
      if (j >= 0 && j < 8) acc += static_cast<double>(row[j] * c_taps11[d + 2]);
    }
    hb[i] = static_cast<float>(acc * c_scale8[x]);
  }
  __syncwarp();
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int p = lane + 32 * k, x = p & 7, y = p >> 3;
    float bl[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      double acc = 0.0;
#pragma unroll
      for (int d = -2; d <= 2; ++d) {
        const int j = y + d;
        if (j >= 0 && j < 8) acc += static_cast<double>(hb[64 * c + 8 * j + x] * c_taps11[d + 2]);
      }
      bl[c] = static_cast<float>(acc * c_scale8[y]);
    }
    float X, Y, B;
    opsin_pixel(bl[0], bl[1], bl[2], lin[p], lin[64 + p], lin[128 + p], X, Y, B);
    dst[p] = X;
    dst[64 + p] = Y;
    dst[128 + p] = B;
  }
  __syncwarp();
}

// Candidate = processed state with coefficient `zidx` zeroed: the touched component's samples -> s.cpx
// (incremental IDCT: one replacement column, then the rows).
__device__ __forceinline__ void warp_candidate_idct(ZeroWarpSmem& s, const int* s_basis, int zidx, int lane) {
  const int zc = zidx >> 6, k = zidx & 63, kx = k & 7, ky = k >> 3;
  if (lane < 8) {
    int acc = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) {
      const int cv = v == ky ? 0 : s.cf[64 * zc + 8 * v + kx];
      acc += s_basis[8 * lane + v] * cv;
    }
    s.ccol[lane] = static_cast<short>(idct_col_round(acc));
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h, x = p & 7, y = p >> 3;
    int acc = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int cv = u == kx ? s.ccol[y] : s.colv[64 * zc + 8 * y + u];
      acc += s_basis[8 * x + u] * cv;
    }
    s.cpx[p] = static_cast<unsigned char>(idct_row_round(acc));
  }
  __syncwarp();
}

// CompareBlock (guetzli/butteraugli_comparator.cc:113-163) from the point where the candidate window is
// linear rgb in s.bufA ([c*64 + 8y + x]), against the original block in s.pg0. Returns the error in
// all lanes.
__device__ __forceinline__ float warp_compare_linear(ZeroWarpSmem& s, const float scale[3], double csf_a, double csf_b,
                                                     int lane);

// The same from the window's Y / Cb / Cr samples pY / pCb / pCr (64 bytes each).
__device__ __forceinline__ float warp_compare_pixels(ZeroWarpSmem& s, const float* lut, const unsigned char* pY,
                                                     const unsigned char* pCb, const unsigned char* pCr,
                                                     int vx, int vy, const float scale[3],
                                                     double csf_a, double csf_b, int lane) {
  // pixels -> linear rgb (window replication past the image edge, output_image.cc:85-97)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h, x = min(p & 7, vx - 1), y = min(p >> 3, vy - 1), sp = 8 * y + x;
    int r, g, b;
    ycbcr_to_rgb(pY[sp], pCb[sp], pCr[sp], r, g, b);
    s.bufA[p] = lut[r];
    s.bufA[64 + p] = lut[g];
    s.bufA[128 + p] = lut[b];
  }
  __syncwarp();
  return warp_compare_linear(s, scale, csf_a, csf_b, lane);
}

__device__ __forceinline__ float warp_compare_linear(ZeroWarpSmem& s, const float scale[3], double csf_a, double csf_b,
                                                     int lane) {
  warp_block_opsin(s.bufA, s.bufB, s.cx, lane);
  // MaskHighIntensityChange(8, 8, orig, cand) -> fa (bufA), fb (bufB)
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h, x = p & 7, y = p >> 3;
    const float c0[3] = {s.pg0[p], s.pg0[64 + p], s.pg0[128 + p]};
    const float c1[3] = {s.cx[p], s.cx[64 + p], s.cx[128 + p]};
    double worst = -1;
    if (x > 0) { const double d = mhic_sqdiff(s.pg0[64 + p - 1], s.cx[64 + p - 1], c0[1], c1[1]); if (worst < d) worst = d; }
    if (x < 7) { const double d = mhic_sqdiff(s.pg0[64 + p + 1], s.cx[64 + p + 1], c0[1], c1[1]); if (worst < d) worst = d; }
    if (y > 0) { const double d = mhic_sqdiff(s.pg0[64 + p - 8], s.cx[64 + p - 8], c0[1], c1[1]); if (worst < d) worst = d; }
    if (y < 7) { const double d = mhic_sqdiff(s.pg0[64 + p + 8], s.cx[64 + p + 8], c0[1], c1[1]); if (worst < d) worst = d; }
    float o0[3], o1[3];
    mhic_pixel(c0, c1, worst, o0, o1);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      s.bufA[64 * c + p] = o0[c];
      s.bufB[64 * c + p] = o1[c];
    }
  }
  __syncwarp();
  double dc[3], ac[3], edge[3];
  warp_block_diff(s.bufA, s.bufB, s.pl, reinterpret_cast<double*>(s.bufA), csf_a, csf_b, dc, ac, edge);
  double diff = 0.0, diff_edge = 0.0;
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const double sc = scale[c];
    diff += dc[c] * sc;
    diff += ac[c] * sc;
    diff_edge += edge[c] * sc;
  }
  return static_cast<float>(sqrt((1 - 0.05) * diff + 0.05 * diff_edge));
}

// CompareBlock on the state in `s`: candidate = processed with coefficient `zidx` zeroed
// (zidx < 0: the processed block itself).
__device__ __forceinline__ float warp_compare_block(ZeroWarpSmem& s, const int* s_basis, const float* lut, int zidx,
                                                    int vx, int vy, const float scale[3],
                                                    double csf_a, double csf_b, int lane) {
  const int zc = zidx >= 0 ? zidx >> 6 : -1;
  if (zidx >= 0) warp_candidate_idct(s, s_basis, zidx, lane);
  return warp_compare_pixels(s, lut, zc == 0 ? s.cpx : s.pix, zc == 1 ? s.cpx : s.pix + 64,
                             zc == 2 ? s.cpx : s.pix + 128, vx, vy, scale, csf_a, csf_b, lane);
}

// Commits the zeroing of coefficient idx: refreshes the touched column and the component's samples.
__device__ __forceinline__ void warp_commit_zero(ZeroWarpSmem& s, const int* s_basis, int idx, int lane) {
  const int c = idx >> 6, k = idx & 63, kx = k & 7;
  if (lane == 0) s.cf[idx] = 0;
  __syncwarp();
  if (lane < 8) {
    int acc = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) acc += s_basis[8 * lane + v] * s.cf[64 * c + 8 * v + kx];
    s.colv[64 * c + 8 * lane + kx] = static_cast<short>(idct_col_round(acc));
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h, x = p & 7, y = p >> 3;
    int acc = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += s_basis[8 * x + u] * s.colv[64 * c + 8 * y + u];
    s.pix[64 * c + p] = static_cast<unsigned char>(idct_row_round(acc));
  }
  __syncwarp();
}

// Rebuilds colv / pix of component c from s.cf (full IDCT, warp-cooperative).
__device__ __forceinline__ void warp_full_idct(ZeroWarpSmem& s, const int* s_basis, int c, int lane) {
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h, x = p & 7, y = p >> 3;
    int acc = 0;
#pragma unroll
    for (int v = 0; v < 8; ++v) acc += s_basis[8 * y + v] * s.cf[64 * c + 8 * v + x];
    s.colv[64 * c + p] = static_cast<short>(idct_col_round(acc));
  }
  __syncwarp();
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h, x = p & 7, y = p >> 3;
    int acc = 0;
#pragma unroll
    for (int u = 0; u < 8; ++u) acc += s_basis[8 * x + u] * s.colv[64 * c + 8 * y + u];
    s.pix[64 * c + p] = static_cast<unsigned char>(idct_row_round(acc));
  }
  __syncwarp();
}

constexpr int kZeroWarps = 4;

// Input order of a block's candidates: ascending key, as std::sort leaves them (processor.cc:410-412).
// Without equal keys any sort gives that order, and up to 16 entries std::sort is a plain (stable)
// insertion sort, so the ranks computed by the whole warp are exact. A block with more than 16 candidates
// AND two equal keys (possible: |57| * csf[84] == |116| * csf[70] in float, see
// tests/test_host_cpu.py::test_zeroing_key_ties_are_possible) gets the library's own arrangement: one
// lane runs the restated introsort (gzb_exact_sort.h) over the (key, index) pairs in the order the
// reference builds them. `pairs` is scratch for 192 pairs (bufA + bufB, dead at this point).
struct ZeroSortPair { float key; int ent; };
struct ZeroSortLess {
  __host__ __device__ bool operator()(const ZeroSortPair& a, const ZeroSortPair& b) const { return a.key < b.key; }
};
// (kept out of line: the rare path must not cost the search loop registers)
__device__ __noinline__ void zero_exact_sort(ZeroSortPair* pairs, int n, unsigned char* order) {
  xsort::sort<16>(pairs, pairs + n, ZeroSortLess());
  for (int i = 0; i < n; ++i) order[i] = static_cast<unsigned char>(pairs[i].ent);
}
__device__ __forceinline__ void warp_input_order(const float* key, const unsigned char* ent, int n, unsigned char* order,
                                                 ZeroSortPair* pairs, unsigned int* tie_counter, int lane) {
  int ties = 0;
  for (int e = lane; e < n; e += 32) {
    const float ke = key[e];
    int rank = 0;
    for (int j = 0; j < n; ++j) {
      const float kj = key[j];
      rank += (kj < ke || (kj == ke && j < e)) ? 1 : 0;
      ties += (kj == ke && j != e) ? 1 : 0;
    }
    order[rank] = ent[e];
  }
  const bool tied = __any_sync(0xffffffffu, ties != 0);
  __syncwarp();
  if (tied && n > 16) {
    for (int e = lane; e < n; e += 32) pairs[e] = ZeroSortPair{key[e], ent[e]};
    __syncwarp();
    if (lane == 0) {
      zero_exact_sort(pairs, n, order);
      if (tie_counter) atomicAdd(tie_counter, 1u);
    }
    __syncwarp();
  }
}

// Longest-processing-time-first order of the blocks [b0, b1): a block's search costs about three
// CompareBlock trials per non-zero AC coefficient, and at ~1 MPix a warp only gets five or six blocks,
// so handing out the expensive blocks first shortens the tail of the launch. Counting sort by the
// number of non-zero AC coefficients, descending (order inside a bucket is arbitrary: blocks are
// independent, the results do not depend on it): cost + histogram, 193-bin scan, scatter.
__global__ void __launch_bounds__(256)
k_zero_block_cost(const int16_t* __restrict__ cur, size_t comp_stride, int comp_mask, int b0, int b1, int bw, int coef_bw,
                  unsigned char* __restrict__ cost, unsigned int* __restrict__ bins) {
  const int b = b0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= b1) return;
  const size_t cb = coef_bw == bw ? static_cast<size_t>(b) : static_cast<size_t>(b / bw) * coef_bw + b % bw;
  int nz = 0;
  for (int c = 0; c < 3; ++c) {
    if (!((comp_mask >> c) & 1)) continue;
    const uint4* p = reinterpret_cast<const uint4*>(cur + c * comp_stride + cb * 64);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint4 v = p[k];
      const unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) nz += ((w[j] & 0xffffu) != 0) + ((w[j] >> 16) != 0);
    }
    nz -= cur[c * comp_stride + cb * 64] != 0;   // DC is not a candidate
  }
  cost[b - b0] = static_cast<unsigned char>(nz);
  atomicAdd(&bins[192 - nz], 1u);
}
__global__ void k_zero_cost_scan(unsigned int* __restrict__ bins) {
  if (threadIdx.x != 0) return;
  unsigned int run = 0;
  for (int i = 0; i < 193; ++i) { const unsigned int v = bins[i]; bins[i] = run; run += v; }
}
__global__ void __launch_bounds__(256)
k_zero_lpt_scatter(const unsigned char* __restrict__ cost, unsigned int* __restrict__ bins, int b0, int b1,
                   int* __restrict__ order) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (b0 + i >= b1) return;
  order[atomicAdd(&bins[192 - cost[i]], 1u)] = b0 + i;
}

// mode 0: full zeroing order -> out[block*192 + r]
// mode 1: single CompareBlock of `cur` with no zeroing -> err_out[block] (stage test entry)
// mode 2: CompareBlock of ONE block `single_block` whose candidate coefficients are the 192 values
//         at `cur` (comp_stride 64, nblocks 1) -> err_out[0] (Comparator::CompareBlock adaptor)
// Blocks [block_begin, nblocks) are processed (a group of GPUs splits the image by block range).
#ifndef GZB_ZERO_MIN_CTAS
#define GZB_ZERO_MIN_CTAS 7
#endif
// TEAM = 1: one warp per block, the <= 3 look-ahead trials of a round one after the other (large images: the
// GPU is full of blocks anyway). TEAM = 3: three warps per block, each with its own private copy of the block
// state, one trial of the round each; they exchange the three errors through shared memory and all commit the
// same winner (the redundant set-up and commits cost ~6 % more instructions, which is why this variant is only
// launched when there are too few blocks to fill the GPU with one warp each: tests/bees.png has 1848).
constexpr int kZeroTeamWarps = 6;   // two teams of three per CTA: 6 x 7888 bytes of state stay under 48 KB of static smem
__device__ __forceinline__ void zero_team_bar(int team) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + team), "r"(96) : "memory");
}
template <int TEAM>
__global__ void __launch_bounds__(TEAM == 1 ? 32 * kZeroWarps : 32 * kZeroTeamWarps, TEAM == 1 ? GZB_ZERO_MIN_CTAS : 4)
k_zeroing_order(const int16_t* __restrict__ orig, const int16_t* __restrict__ cur,
                size_t comp_stride, const uint8_t* __restrict__ rgb_planes, size_t plane_stride,
                int P, int W, int H, int bw, int nblocks, const float* __restrict__ mask_scale,
                int comp_mask, float limit, int lookahead, int mode, int single_block, int block_begin,
                CoeffDataDev* __restrict__ out, float* __restrict__ err_out,
                float* __restrict__ pregamma_out, unsigned int* __restrict__ counter,
                const int* __restrict__ lpt_order, int coef_bw, const uint8_t* __restrict__ fixed_chroma,
                unsigned int* __restrict__ tie_counter) {
  constexpr int kWarps = TEAM == 1 ? kZeroWarps : kZeroTeamWarps;
  __shared__ ZeroWarpSmem sm[kWarps];
  __shared__ int s_basis[64];
  __shared__ unsigned int s_team_blk[2];
  __shared__ float s_team_err[2][2][3];   // [team][round parity][member]
  const float* s_lut = g_tab.srgb_lin;   // 1 KB, L1-resident; shared memory is the occupancy limiter
  if (threadIdx.x < 64) s_basis[threadIdx.x] = kIdctBasis[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int team = warp / TEAM, member = warp - team * TEAM;
  ZeroWarpSmem& s = sm[warp];
  const double csf_a = kCsf8x8[4 + lane], csf_b = kCsf8x8[36];
  for (;;) {
    unsigned int b = 0;
    if (TEAM == 1) {
      if (lane == 0) b = atomicAdd(counter, 1u);
      b = __shfl_sync(0xffffffffu, b, 0) + static_cast<unsigned int>(block_begin);
    } else {
      if (member == 0 && lane == 0) s_team_blk[team] = atomicAdd(counter, 1u);
      zero_team_bar(team);
      b = s_team_blk[team] + static_cast<unsigned int>(block_begin);
    }
    if (b >= static_cast<unsigned int>(nblocks)) break;
    if (lpt_order) b = static_cast<unsigned int>(lpt_order[b - block_begin]);
    const int blk = mode == 2 ? single_block : static_cast<int>(b);  // image block (b indexes coefficients)
    const int bx = blk % bw, by = blk / bw;
    const int vx = min(8, W - 8 * bx), vy = min(8, H - 8 * by);
    // coefficient block: the luma plane of a 4:2:0 image is MCU-padded to coef_bw blocks per row
    const size_t cb = (mode == 2 || coef_bw == bw) ? static_cast<size_t>(b) : static_cast<size_t>(by) * coef_bw + bx;
    // ---- load block state ----
#pragma unroll
    for (int k = 0; k < 6; ++k) {
      const int i = lane + 32 * k, c = i >> 6;
      s.cf[i] = (comp_mask >> c) & 1 ? cur[c * comp_stride + cb * 64 + (i & 63)] : 0;
    }
    // original pixels, coordinates clamped to the image (SwitchBlock, butteraugli_comparator.cc:96-105)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int p = lane + 32 * h;
      const int x = min(8 * bx + (p & 7), W - 1), y = min(8 * by + (p >> 3), H - 1);
      const size_t g = static_cast<size_t>(y) * P + x;
#pragma unroll
      for (int c = 0; c < 3; ++c) s.bufA[64 * c + p] = s_lut[rgb_planes[c * plane_stride + g]];
    }
    __syncwarp();
    warp_block_opsin(s.bufA, s.bufB, s.pg0, lane);
    if (pregamma_out && member == 0) {
#pragma unroll
      for (int k = 0; k < 6; ++k) pregamma_out[static_cast<size_t>(b) * 192 + lane + 32 * k] = s.pg0[lane + 32 * k];
    }
#pragma unroll 1
    for (int c = 0; c < 3; ++c) warp_full_idct(s, s_basis, c, lane);
    if (fixed_chroma) {   // 4:2:0 luma pass: Cb / Cr of the window are the image's upsampled samples
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const int p = lane + 32 * h;
        const size_t g = static_cast<size_t>(8 * by + (p >> 3)) * P + 8 * bx + (p & 7);
        s.pix[64 + p] = fixed_chroma[g];
        s.pix[128 + p] = fixed_chroma[plane_stride + g];
      }
      __syncwarp();
    }
    const float scale[3] = {mask_scale[3 * blk], mask_scale[3 * blk + 1], mask_scale[3 * blk + 2]};
    if (mode >= 1) {
      const float e = warp_compare_block(s, s_basis, s_lut, -1, vx, vy, scale, csf_a, csf_b, lane);
      if (lane == 0) err_out[b] = e;
      continue;
    }
    // ---- input order: non-zero AC coefficients sorted by |orig|*csf + bias (stable) ----
    int n = 0;
#pragma unroll 1
    for (int k = 0; k < 6; ++k) {
      const int i = lane + 32 * k, c = i >> 6;
      const bool take = ((comp_mask >> c) & 1) && (i & 63) != 0 && s.cf[i] != 0;
      const unsigned int m = __ballot_sync(0xffffffffu, take);
      if (take) {
        const int pos = n + __popc(m & ((1u << lane) - 1));
        const int o = orig[c * comp_stride + cb * 64 + (i & 63)];
        s.ent[pos] = static_cast<unsigned char>(i);
        s.key[pos] = abs(o) * c_zero_csf[i] + c_zero_bias[i];
      }
      n += __popc(m);
    }
    __syncwarp();
    warp_input_order(s.key, s.ent, n, s.order, reinterpret_cast<ZeroSortPair*>(s.bufA), member == 0 ? tie_counter : nullptr, lane);
    // ---- greedy loop ----
    CoeffDataDev* o = out + static_cast<size_t>(b) * 192;
    int win[3];
    int nwin = min(lookahead, n), next = nwin, nout = 0;
    for (int i = 0; i < 3; ++i) win[i] = i < nwin ? s.order[i] : 0;
    int round = 0;
    while (nwin > 0) {
      float best_err = 1e17f;
      int best_i = 0;
      if (TEAM == 1) {
        for (int i = 0; i < nwin; ++i) {
          const float err = warp_compare_block(s, s_basis, s_lut, win[i], vx, vy, scale, csf_a, csf_b, lane);
          const float max_err = fmaxf(0.0f, err);
          if (max_err < best_err) { best_err = max_err; best_i = i; }
        }
      } else {
        // one trial of the round per member; the errors meet in shared memory (slots alternate by round, so one
        // barrier per round orders the writes of round r + 2 after the reads of round r)
        float mine = 0.0f;
        if (member < nwin) mine = fmaxf(0.0f, warp_compare_block(s, s_basis, s_lut, win[member], vx, vy, scale, csf_a, csf_b, lane));
        if (lane == 0) s_team_err[team][round & 1][member] = mine;
        zero_team_bar(team);
        for (int i = 0; i < nwin; ++i) {
          const float max_err = s_team_err[team][round & 1][i];
          if (max_err < best_err) { best_err = max_err; best_i = i; }
        }
        ++round;
      }
      const int idx = win[best_i];
      warp_commit_zero(s, s_basis, idx, lane);
      if (lane == 0 && member == 0) { o[nout].idx = idx; o[nout].block_err = best_err; }
      ++nout;
      for (int i = best_i; i + 1 < nwin; ++i) win[i] = win[i + 1];
      if (next < n) win[nwin - 1] = s.order[next++];
      else --nwin;
    }
    __syncwarp();
    // ---- monotone from the tail, cut at the limit (processor.cc:467-479) ----
    if (lane == 0 && member == 0) {
      float min_err = 1e10f;
      for (int i = nout - 1; i >= 0; --i) {
        min_err = fminf(min_err, o[i].block_err);
        o[i].block_err = min_err;
      }
      int keep = 0;
      while (keep < nout && o[keep].block_err <= limit) ++keep;
      for (int i = keep; i < nout; ++i) { o[i].idx = 0; o[i].block_err = 0.0f; }
    }
    __syncwarp();
  }
}

// Comparator::CompareBlock for a caller-rendered window: rgb192 = the 8x8 window of the candidate as
// interleaved sRGB8 (OutputImage::ToSRGB(xmin, ymin, 8, 8), any chroma sampling), compared against the
// original's block (bx, by). One warp. -> err_out[0]
__global__ void __launch_bounds__(32)
k_compare_block_rgb(const uint8_t* __restrict__ rgb192, const uint8_t* __restrict__ rgb_planes, size_t plane_stride,
                    int P, int W, int H, int bw, int bx, int by, const float* __restrict__ mask_scale,
                    float* __restrict__ err_out) {
  __shared__ ZeroWarpSmem s;
  const float* lut = g_tab.srgb_lin;
  const int lane = threadIdx.x;
  const double csf_a = kCsf8x8[4 + lane], csf_b = kCsf8x8[36];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h;
    const int x = min(8 * bx + (p & 7), W - 1), y = min(8 * by + (p >> 3), H - 1);
    const size_t g = static_cast<size_t>(y) * P + x;
#pragma unroll
    for (int c = 0; c < 3; ++c) s.bufA[64 * c + p] = lut[rgb_planes[c * plane_stride + g]];
  }
  __syncwarp();
  warp_block_opsin(s.bufA, s.bufB, s.pg0, lane);
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int p = lane + 32 * h;
#pragma unroll
    for (int c = 0; c < 3; ++c) s.bufA[64 * c + p] = lut[rgb192[3 * p + c]];
  }
  __syncwarp();
  const int blk = by * bw + bx;
  const float scale[3] = {mask_scale[3 * blk], mask_scale[3 * blk + 1], mask_scale[3 * blk + 2]};
  const float e = warp_compare_linear(s, scale, csf_a, csf_b, lane);
  if (lane == 0) err_out[0] = e;
}

}  // namespace gzb
