// gzb_huffman.cuh -- the entropy-coded segment of WriteJpeg on the device.
//
// The reference serialises every candidate to a JPEG just to learn its size
// (TryQuantMatrix / SelectFrequencyBackEnd -> OutputImage::SaveToJpegData + WriteJpeg,
// guetzli/processor.cc:279-308, 897-915; guetzli/jpeg_data_writer.cc:361-553): a sequential Huffman
// coder over all coefficients. Here the candidate's coefficients already live in HBM, so the scan
// is coded there:
//   k_huff_histogram : DC/AC symbol histograms per component (BuildDCHistograms /
//                      BuildACHistograms, jpeg_data_writer.cc:189-247) -> host builds the codes
//   k_huff_code<false> : bits of every (block, component) unit under those codes, summed per CTA
//   k_scan_u64         : exclusive scan of the CTA totals (a few hundred to a few thousand values)
//   k_huff_code<true>  : CTA-local scan of the unit bits + the CTA's offset = each unit's bit position;
//                        every unit writes its code words there (EncodeScan/EncodeDCTBlockSequential,
//                        jpeg_data_writer.cc:249-359), MSB first, big-endian bytes
//   k_huff_finish    : pads the last byte with ones (JumpToByteBoundary) and counts 0xff bytes
// The file size is header + bytes + (#0xff, one stuffing byte each) + 2 (EOI); the bytes themselves
// are copied to the host only for a candidate that becomes the best so far.
// One thread per unit; a CTA stages its units' 128-byte coefficient rows through padded shared
// memory with coalesced 16-byte loads. HBM-bound: 2 B per coefficient in, ~1 bit per coefficient out.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gzb {

struct HuffDeviceTables {      // uploaded per file
  uint16_t dc_code[3][16];
  uint16_t ac_code[3][256];
  uint8_t dc_len[3][16];
  uint8_t ac_len[3][256];
};

__constant__ uint8_t c_natural_order[64] = {  // zig-zag position -> natural index
    0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,  12, 19, 26, 33, 40, 48,
    41, 34, 27, 20, 13, 6,  7,  14, 21, 28, 35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23,
    30, 37, 44, 51, 58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};

// Scan order of the coding units (EncodeScan, jpeg_data_writer.cc:506-536):
//   mode 0: 1x1 MCUs, unit u = block * ncomp + comp (4:4:4, or one component);
//   mode 1: 4:2:0 MCUs, six units per MCU: the four luma blocks (2mx+ix, 2my+iy) of the MCU-padded luma
//           plane (cbw0 blocks per row), then Cb and Cr block (mx, my);
//   mode 2: luma only, image blocks (bw per row) read from the MCU-padded plane (SaveToJpegData drops
//           all-zero chroma planes and then writes a grey 1x1 file, output_image.cc:588-603).
struct HuffLayout { int mode, ncomp, mcw, bw, cbw0; };

// unit -> component, coefficient block, coefficient block of the unit's DC predecessor (-1: none)
__device__ __forceinline__ void huff_map(const HuffLayout& L, long long u, int* c, long long* cb, long long* pcb) {
  if (L.mode == 0) {
    const long long b = u / L.ncomp;
    *c = static_cast<int>(u - b * L.ncomp);
    *cb = b;
    *pcb = b - 1;
  } else if (L.mode == 1) {
    const long long mcu = u / 6;
    const int r = static_cast<int>(u - mcu * 6);
    const long long mx = mcu % L.mcw, my = mcu / L.mcw;
    if (r < 4) {
      *c = 0;
      *cb = (2 * my + (r >> 1)) * L.cbw0 + 2 * mx + (r & 1);
      if (r > 0) {
        *pcb = (2 * my + ((r - 1) >> 1)) * L.cbw0 + 2 * mx + ((r - 1) & 1);
      } else if (mcu > 0) {
        const long long px = (mcu - 1) % L.mcw, py = (mcu - 1) / L.mcw;
        *pcb = (2 * py + 1) * L.cbw0 + 2 * px + 1;
      } else {
        *pcb = -1;
      }
    } else {
      *c = r - 3;
      *cb = mcu;
      *pcb = mcu - 1;
    }
  } else {
    *c = 0;
    *cb = (u / L.bw) * L.cbw0 + u % L.bw;
    *pcb = u > 0 ? ((u - 1) / L.bw) * L.cbw0 + (u - 1) % L.bw : -1;
  }
}

constexpr int kHuffThreads = 128;
constexpr int kHuffRow = 66;   // shorts per staged row: 132 bytes = 33 words -> conflict-free columns

__device__ __forceinline__ int bit_length_u(unsigned v) { return 32 - __clz(v); }

// Stages the quantised indices of units [u0, u0 + kHuffThreads) into shared memory.
// coef holds dequantised values (multiples of q), so index = coef / q.
__device__ __forceinline__ void huff_stage_rows(const int16_t* __restrict__ coef, size_t comp_stride,
                                                const int* __restrict__ s_q, const HuffLayout& L, long long u0,
                                                long long nunits, int16_t* s_rows) {
  // 128 rows x 8 chunks of 16 bytes
  for (int i = threadIdx.x; i < kHuffThreads * 8; i += kHuffThreads) {
    const int r = i >> 3, ch = i & 7;
    const long long u = u0 + r;
    if (u >= nunits) continue;
    int c;
    long long b, pb;
    huff_map(L, u, &c, &b, &pb);
    const uint4 v = *reinterpret_cast<const uint4*>(coef + c * comp_stride + static_cast<size_t>(b) * 64 + ch * 8);
    const int16_t* pv = reinterpret_cast<const int16_t*>(&v);
    int16_t* dst = s_rows + r * kHuffRow + ch * 8;
#pragma unroll
    for (int k = 0; k < 8; ++k) dst[k] = static_cast<int16_t>(pv[k] / s_q[64 * c + ch * 8 + k]);
  }
}

// DC index of the unit that precedes this one in its component's scan order (pb < 0: none).
__device__ __forceinline__ int huff_prev_dc(const int16_t* __restrict__ coef, size_t comp_stride, const int* s_q,
                                            long long pb, int c) {
  if (pb < 0) return 0;
  return coef[c * comp_stride + static_cast<size_t>(pb) * 64] / s_q[64 * c];
}

__global__ void __launch_bounds__(kHuffThreads)
k_huff_histogram(const int16_t* __restrict__ coef, size_t comp_stride, const int* __restrict__ q192, HuffLayout L,
                 long long nunits, unsigned int* __restrict__ dc_hist /*[3][16]*/,
                 unsigned int* __restrict__ ac_hist /*[3][256]*/) {
  __shared__ int16_t s_rows[kHuffThreads * kHuffRow];
  __shared__ int s_q[192];
  __shared__ unsigned int s_dc[3 * 16], s_ac[3 * 256];
  for (int i = threadIdx.x; i < 192; i += kHuffThreads) s_q[i] = q192[i];
  for (int i = threadIdx.x; i < 48; i += kHuffThreads) s_dc[i] = 0;
  for (int i = threadIdx.x; i < 768; i += kHuffThreads) s_ac[i] = 0;
  __syncthreads();
  const long long u0 = static_cast<long long>(blockIdx.x) * kHuffThreads;
  huff_stage_rows(coef, comp_stride, s_q, L, u0, nunits, s_rows);
  __syncthreads();
  const long long u = u0 + threadIdx.x;
  if (u < nunits) {
    int c;
    long long b, pb;
    huff_map(L, u, &c, &b, &pb);
    const int16_t* row = s_rows + threadIdx.x * kHuffRow;
    const int diff = row[0] - huff_prev_dc(coef, comp_stride, s_q, pb, c);
    atomicAdd(&s_dc[16 * c + bit_length_u(static_cast<unsigned>(abs(diff)))], 1u);
    int run = 0;
#pragma unroll 1
    for (int k = 1; k < 64; ++k) {
      const int v = row[c_natural_order[k]];
      if (v == 0) { ++run; continue; }
      while (run > 15) { atomicAdd(&s_ac[256 * c + 0xf0], 1u); run -= 16; }
      atomicAdd(&s_ac[256 * c + (run << 4) + bit_length_u(static_cast<unsigned>(abs(v)))], 1u);
      run = 0;
    }
    if (run > 0) atomicAdd(&s_ac[256 * c], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 48; i += kHuffThreads) if (s_dc[i]) atomicAdd(&dc_hist[i], s_dc[i]);
  for (int i = threadIdx.x; i < 768; i += kHuffThreads) if (s_ac[i]) atomicAdd(&ac_hist[i], s_ac[i]);
}

// Bit sink of one unit. EMIT=false only counts.
template <bool EMIT>
struct HuffSink {
  unsigned int* words;        // scan buffer as 32-bit words (big-endian byte order inside each word)
  unsigned long long pos;     // next bit position (absolute)
  unsigned long long acc;     // pending bits, right-aligned
  int nacc;                   // number of pending bits (< 32 after put), includes the leading gap
  unsigned int count;
  __device__ __forceinline__ void init(unsigned int* w, unsigned long long bitpos) {
    words = w; pos = bitpos & ~31ull; acc = 0; nacc = static_cast<int>(bitpos & 31); count = 0;
  }
  __device__ __forceinline__ void put(int nbits, unsigned int bits) {  // nbits <= 27
    if (!EMIT) { count += nbits; return; }
    acc = (acc << nbits) | bits;
    nacc += nbits;
    if (nacc >= 32) {
      nacc -= 32;
      const unsigned int w = static_cast<unsigned int>(acc >> nacc);
      atomicOr(words + (pos >> 5), __byte_perm(w, 0, 0x0123));
      pos += 32;
      acc &= (1ull << nacc) - 1;
    }
  }
  __device__ __forceinline__ void flush() {
    if (!EMIT || nacc == 0) return;
    const unsigned int w = static_cast<unsigned int>(acc << (32 - nacc));
    atomicOr(words + (pos >> 5), __byte_perm(w, 0, 0x0123));
  }
};

template <bool EMIT>
__global__ void __launch_bounds__(kHuffThreads)
k_huff_code(const int16_t* __restrict__ coef, size_t comp_stride, const int* __restrict__ q192, HuffLayout L,
            long long nunits, const HuffDeviceTables* __restrict__ tab, unsigned int* __restrict__ unit_bits,
            unsigned int* __restrict__ cta_bits, const unsigned long long* __restrict__ cta_off,
            unsigned int* __restrict__ words) {
  __shared__ unsigned int s_warp[kHuffThreads / 32];
  __shared__ int16_t s_rows[kHuffThreads * kHuffRow];
  __shared__ int s_q[192];
  __shared__ HuffDeviceTables s_tab;
  for (int i = threadIdx.x; i < 192; i += kHuffThreads) s_q[i] = q192[i];
  {
    const unsigned int* src = reinterpret_cast<const unsigned int*>(tab);
    unsigned int* dst = reinterpret_cast<unsigned int*>(&s_tab);
    for (int i = threadIdx.x; i < static_cast<int>(sizeof(HuffDeviceTables) / 4); i += kHuffThreads) dst[i] = src[i];
  }
  __syncthreads();
  const long long u0 = static_cast<long long>(blockIdx.x) * kHuffThreads;
  huff_stage_rows(coef, comp_stride, s_q, L, u0, nunits, s_rows);
  __syncthreads();
  const long long u = u0 + threadIdx.x;
  const bool live = u < nunits;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned long long start = 0;
  if (EMIT) {  // exclusive scan of this CTA's unit bits
    const unsigned int mine = live ? unit_bits[u] : 0u;
    unsigned int incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned int v = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += v; }
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    unsigned int base = 0;
    for (int w = 0; w < warp; ++w) base += s_warp[w];
    start = cta_off[blockIdx.x] + base + (incl - mine);
  }
  int c = 0;
  long long b = 0, pb = -1;
  if (live) huff_map(L, u, &c, &b, &pb);
  const int16_t* row = s_rows + threadIdx.x * kHuffRow;
  HuffSink<EMIT> sink;
  sink.init(words, start);
  if (live) {
  {  // DC: coeff_t arithmetic as in the writer (jpeg_data_writer.cc:262-276)
    const int diff = static_cast<int16_t>(row[0] - huff_prev_dc(coef, comp_stride, s_q, pb, c));
    int mag = diff, low = diff;
    if (diff < 0) { mag = -diff; low = diff - 1; }
    mag = static_cast<int16_t>(mag);
    const int nbits = bit_length_u(static_cast<unsigned>(mag));
    sink.put(s_tab.dc_len[c][nbits] + nbits,
             (static_cast<unsigned int>(s_tab.dc_code[c][nbits]) << nbits) | (static_cast<unsigned int>(low) & ((1u << nbits) - 1)));
  }
  int run = 0;
#pragma unroll 1
  for (int k = 1; k < 64; ++k) {
    const int v = row[c_natural_order[k]];
    if (v == 0) { ++run; continue; }
    while (run > 15) { sink.put(s_tab.ac_len[c][0xf0], s_tab.ac_code[c][0xf0]); run -= 16; }
    const int a = abs(v);
    const int nbits = bit_length_u(static_cast<unsigned>(a));
    const int sym = (run << 4) + nbits;
    const int lowbits = v + (v >> 31);  // v < 0 ? ~a : a
    sink.put(s_tab.ac_len[c][sym] + nbits,
             (static_cast<unsigned int>(s_tab.ac_code[c][sym]) << nbits) | (static_cast<unsigned int>(lowbits) & ((1u << nbits) - 1)));
    run = 0;
  }
  if (run > 0) sink.put(s_tab.ac_len[c][0], s_tab.ac_code[c][0]);
  }
  if (EMIT) {
    if (live) sink.flush();
  } else {
    if (live) unit_bits[u] = sink.count;
    unsigned int sum = live ? sink.count : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
    if (lane == 0) s_warp[warp] = sum;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned int t = 0;
      for (int w = 0; w < kHuffThreads / 32; ++w) t += s_warp[w];
      cta_bits[blockIdx.x] = t;
    }
  }
}

// Single-CTA exclusive scan of n 32-bit counts into 64-bit offsets[0..n].
__global__ void __launch_bounds__(1024)
k_scan_u64(const unsigned int* __restrict__ counts, long long n, unsigned long long* __restrict__ offsets) {
  __shared__ unsigned long long s_part[1024];
  const int t = threadIdx.x;
  const long long per = (n + 1023) / 1024;
  const long long b0 = min(n, t * per), b1 = min(n, b0 + per);
  unsigned long long sum = 0;
  for (long long b = b0; b < b1; ++b) sum += counts[b];
  s_part[t] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {
    const unsigned long long v = t >= off ? s_part[t - off] : 0;
    __syncthreads();
    s_part[t] += v;
    __syncthreads();
  }
  unsigned long long run = s_part[t] - sum;
  for (long long b = b0; b < b1; ++b) { offsets[b] = run; run += counts[b]; }
  if (t == 1023) offsets[n] = s_part[1023];
}

// Pads the last byte with ones and counts the 0xff bytes of the scan (each gets a stuffing byte).
// out2[0] = scan bytes (before stuffing), out2[1] = number of 0xff bytes.
__global__ void k_huff_finish(unsigned char* __restrict__ bytes, const unsigned long long* __restrict__ total_bits_p,
                              unsigned long long* __restrict__ out2) {
  const unsigned long long total_bits = *total_bits_p;
  const unsigned long long nbytes = (total_bits + 7) >> 3;
  const unsigned long long tid = static_cast<unsigned long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * blockDim.x;
  const int tail = static_cast<int>(total_bits & 7);
  unsigned int ff = 0;
  for (unsigned long long i = tid; i < nbytes; i += stride) {
    unsigned int v = bytes[i];
    if (i == nbytes - 1 && tail) {
      v |= (1u << (8 - tail)) - 1;
      bytes[i] = static_cast<unsigned char>(v);
    }
    ff += v == 0xff;
  }
  for (int o = 16; o > 0; o >>= 1) ff += __shfl_down_sync(0xffffffffu, ff, o);
  if ((threadIdx.x & 31) == 0 && ff) atomicAdd(out2 + 1, static_cast<unsigned long long>(ff));
  if (tid == 0) out2[0] = nbytes;
}

}  // namespace gzb
