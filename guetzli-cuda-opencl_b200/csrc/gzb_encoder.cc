// gzb_encoder.cc -- host search driver: guetzli::Process(rgb -> jpeg) with every butteraugli
// evaluation, IDCT/quantisation pass and the block-zeroing search running on the B200 through the
// C ABI of this library. The control flow and every scalar decision follow the reference's
// Processor (guetzli/processor.cc:151-372, 559-1020, 1157-1185) for clear_metadata=true,
// zeroing_greedy_lookahead=3, new_zeroing_model=true, use_silver_screen=false and any try_420 /
// force_420 (the 4:4:4 pass and the YUV420 pass with its two frequency-masking passes), so that the
// emitted JPEG is byte-identical to the CPU reference's.
//
// Host-side accelerations that do not change any result:
//   * ComputeEntropyCodes is evaluated only at the steps whose value can be observed
//     (processor.cc:879-888 recomputes it every 10th step but reads it only in the break test);
//   * the entropy-coded size is tracked incrementally between code rebuilds;
//   * the global candidate order is produced by a lazy evaluation of libstdc++'s introsort that
//     yields exactly std::sort's permutation for the prefix actually consumed;
//   * WriteJpeg runs block-row bands on host threads and overlaps the GPU Compare.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <condition_variable>
#include <memory>
#include <mutex>
#include <set>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/gzb200.h"
#include "gzb_exact_sort.h"
#include "gzb_jpeg.h"
#include "gzb_quant_search.h"

namespace {

using gzb::jpeg::Frame;
using gzb::jpeg::Histogram;

double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---- front end: RGB -> YCbCr -> integer forward DCT -> q=1 indices ---------------------------
// (guetzli/jpeg_data_encoder.cc:28-117, guetzli/fdct.cc:28-240; truncating shifts keep the order)
inline int mul16(int a, int b) { return (a * b) >> 16; }

void fdct_column(int16_t* v) {
  const int i0 = v[0], i1 = v[8], i2 = v[16], i3 = v[24], i4 = v[32], i5 = v[40], i6 = v[48], i7 = v[56];
  int d07 = i0 - i7, s07 = i0 + i7, d25 = i2 - i5, s25 = i2 + i5;
  int d34 = i3 - i4, s34 = i3 + i4, d16 = i1 - i6, s16 = i1 + i6;
  int e0 = s07 - s34, e1 = s07 + s34, e2 = s16 - s25, e3 = s16 + s25;
  e1 <<= 3; e3 <<= 3;
  v[0] = static_cast<int16_t>(e1 + e3);
  v[32] = static_cast<int16_t>(e1 - e3);
  e0 <<= 3; e2 <<= 3; d34 <<= 3; d07 <<= 3;
  v[16] = static_cast<int16_t>(mul16(27146, e2) + e0);
  v[48] = static_cast<int16_t>(mul16(27146, e0) - e2);
  d25 <<= 4; d16 <<= 4;
  const int r = mul16(d16 + d25, 23170), s = mul16(d16 - d25, 23170);
  const int p3 = d34 - s, p1 = d34 + s, p0 = d07 - r, p2 = d07 + r;
  const int q3 = mul16(p3, -21746) + p3 + 1;
  const int q1 = mul16(p1, 13036) + p2 + 1;
  const int q4 = mul16(-21746, p0) + p0;
  const int q5 = mul16(13036, p2);
  v[8] = static_cast<int16_t>(q1);
  v[24] = static_cast<int16_t>(p0 - q3);
  v[40] = static_cast<int16_t>(p3 + q4);
  v[56] = static_cast<int16_t>(q5 - p1);
}

void fdct_row(int16_t* in, const int16_t* t) {
  const int a0 = in[0] + in[7], b0 = in[0] - in[7], a1 = in[1] + in[6], b1 = in[1] - in[6];
  const int a2 = in[2] + in[5], b2 = in[2] - in[5], a3 = in[3] + in[4], b3 = in[3] - in[4];
  const int C1 = t[0], C2 = t[1], C3 = t[2], C4 = t[3], C5 = t[4], C6 = t[5], C7 = t[6];
  const int c0 = a0 + a3, c1 = a0 - a3, c2 = a1 + a2, c3 = a1 - a2;
  in[0] = static_cast<int16_t>((C4 * (c0 + c2)) >> 16);
  in[4] = static_cast<int16_t>((C4 * (c0 - c2)) >> 16);
  in[2] = static_cast<int16_t>((C2 * c1 + C6 * c3) >> 16);
  in[6] = static_cast<int16_t>((C6 * c1 - C2 * c3) >> 16);
  in[1] = static_cast<int16_t>((C1 * b0 + C3 * b1 + C5 * b2 + C7 * b3) >> 16);
  in[3] = static_cast<int16_t>((C3 * b0 - C7 * b1 - C1 * b2 - C5 * b3) >> 16);
  in[5] = static_cast<int16_t>((C5 * b0 - C1 * b1 + C7 * b2 + C3 * b3) >> 16);
  in[7] = static_cast<int16_t>((C7 * b0 - C5 * b1 + C3 * b2 - C1 * b3) >> 16);
}

void fdct_block(int16_t* block) {
  static const int16_t T04[7] = {22725, 21407, 19266, 16384, 12873, 8867, 4520};
  static const int16_t T17[7] = {31521, 29692, 26722, 22725, 17855, 12299, 6270};
  static const int16_t T26[7] = {29692, 27969, 25172, 21407, 16819, 11585, 5906};
  static const int16_t T35[7] = {26722, 25172, 22654, 19266, 15137, 10426, 5315};
  static const int16_t* const rows[8] = {T04, T17, T26, T35, T04, T35, T26, T17};
  for (int x = 0; x < 8; ++x) fdct_column(block + x);
  for (int y = 0; y < 8; ++y) fdct_row(block + 8 * y, rows[y]);
}

void rgb_to_coeffs_rows(const uint8_t* rgb, int w, int h, int bw, int by0, int by1, int16_t* const out[3]) {
  for (int by = by0; by < by1; ++by)
    for (int bx = 0; bx < bw; ++bx) {
      int16_t blk[192];
      for (int iy = 0; iy < 8; ++iy)
        for (int ix = 0; ix < 8; ++ix) {
          const int y = std::min(h - 1, 8 * by + iy), x = std::min(w - 1, 8 * bx + ix);
          const uint8_t* p = rgb + 3 * (static_cast<size_t>(y) * w + x);
          const int r = p[0], g = p[1], b = p[2], k = 8 * iy + ix;
          blk[k] = static_cast<int16_t>((19595 * r + 38469 * g + 7471 * b - (128 << 16) + 32768) >> 16);
          blk[64 + k] = static_cast<int16_t>((-11059 * r - 21709 * g + 32768 * b + 32768 - 1) >> 16);
          blk[128 + k] = static_cast<int16_t>((32768 * r - 27439 * g - 5329 * b + 32768 - 1) >> 16);
        }
      for (int c = 0; c < 3; ++c) {
        fdct_block(blk + 64 * c);
        int16_t* o = out[c] + (static_cast<size_t>(by) * bw + bx) * 64;
        for (int k = 0; k < 64; ++k) o[k] = static_cast<int16_t>((blk[64 * c + k] * 65537 + (0x80 << 12)) >> 20);
      }
    }
}

template <typename F>
void parallel_rows(int n, gzb::WorkerPool* pool, F fn) {
  const int T = std::max(1, std::min(pool ? pool->size() : 1, n));
  if (T == 1) { fn(0, n); return; }
  pool->run(T, [&](int t) {
    fn(static_cast<int>(static_cast<int64_t>(n) * t / T), static_cast<int>(static_cast<int64_t>(n) * (t + 1) / T));
  });
}

// ---- quantiser (guetzli/quantize.h:24-29) -----------------------------------------------------
inline int16_t quantize_coeff(int16_t raw, int q) {
  const int r = raw % q;
  const int16_t delta = static_cast<int16_t>(2 * r > q ? q - r : (-2) * r > q ? -q - r : -r);
  return static_cast<int16_t>(raw + delta);
}

using gzb::QuantData;
using gzb::QuantGenerator;
using gzb::quant_heuristic_score;
using gzb::quant_data_better;

// ---- std::sort's permutation, restated ---------------------------------------------------------
// The reference orders the back end's candidates with std::sort on float keys (processor.cc:825-828);
// ties between different blocks are the norm, so WHICH of the tied entries falls inside the consumed
// part of the order decides which coefficients are flipped. The reference's golden files come from a
// libstdc++ build, whose std::sort is introsort: median-of-three + unguarded Hoare partitioning down to
// ranges of 16 with a depth budget of 2 * floor(log2(n)), heap sort when the budget runs out, one final
// insertion sort. The long ranges are partitioned on the device with the same element movements
// (gzb_backend.cuh: k_be_tiles_* / k_be_swap / k_be_local); the short ranges it hands back are finished here. This is a
// restatement of the algorithm, not a call into the library's private functions; sort_emulation_ok()
// checks it once per process against the std::sort this library was built with and the driver falls back
// to sorting the whole order with std::sort itself when they disagree (a different standard library).
typedef std::pair<int, float> OrderEntry;
static_assert(sizeof(OrderEntry) == sizeof(gzb_order_entry), "std::pair<int, float> must be 8 bytes");
struct OrderLess {
  bool operator()(const OrderEntry& a, const OrderEntry& b) const { return a.second < b.second; }
};

namespace exact_sort {

// the algorithm itself: gzb_exact_sort.h (shared with the device code of the block-zeroing search)
inline OrderEntry* partition_pivot(OrderEntry* first, OrderEntry* last) { return gzb::xsort::partition_pivot(first, last, OrderLess()); }
inline void heap_sort(OrderEntry* first, OrderEntry* last) { gzb::xsort::heap_sort(first, last, OrderLess()); }
inline void finish_range(OrderEntry* first, OrderEntry* last, int depth) { gzb::xsort::finish_range<72>(first, last, depth, OrderLess()); }
inline int depth_budget(size_t n) { return gzb::xsort::depth_budget(n); }

// One-time check of the restatement against the std::sort of this build on tie-heavy inputs.
bool emulation_ok() {
  static const bool ok = [] {
    uint32_t seed = 12345;
    auto rnd = [&] { seed = seed * 1664525u + 1013904223u; return seed >> 8; };
    for (int trial = 0; trial < 6; ++trial) {
      const size_t n = trial < 2 ? 40 + 977 * trial : 3000 + 1777 * trial;
      std::vector<OrderEntry> a(n), b;
      const unsigned levels = trial % 3 == 0 ? 3 : trial % 3 == 1 ? 40 : 100000;
      for (size_t i = 0; i < n; ++i) a[i] = std::make_pair(static_cast<int>(i), static_cast<float>(rnd() % levels) * 0.25f);
      if (trial == 5) for (size_t i = 0; i < n; ++i) a[i].second = static_cast<float>(i / 7);   // pre-sorted runs: deep recursion
      b = a;
      std::sort(b.begin(), b.end(), OrderLess());
      finish_range(a.data(), a.data() + n, depth_budget(n));
      if (a != b) return false;
      // the heap-sort fallback (depth budget exhausted) against std::partial_sort, which is the same library path
      std::vector<OrderEntry> c2 = b, d2;
      for (size_t i = 0; i + 1 < n; i += 2) std::swap(c2[i], c2[n - 1 - i / 2]);
      d2 = c2;
      std::partial_sort(d2.begin(), d2.end(), d2.end(), OrderLess());
      heap_sort(c2.data(), c2.data() + n);
      if (c2 != d2) return false;
    }
    return true;
  }();
  static const bool forced_off = getenv("GZB_NO_SORT_EMULATION") != nullptr;
  return ok && !forced_off;
}


// Lazy std::sort of one short range of the order on the host: the same partition tree as
// xsort::introsort_loop, but a right-hand part is only finished when the walk reaches it.
struct HostLazy {
  std::vector<OrderEntry> buf;
  size_t done = 0;       // buf[0, done) is final
  size_t appended = 0;   // ... and this much of it has been handed to the walk
  struct R { size_t first, last; int depth; };
  std::vector<R> pending;   // back() = leftmost
  void reset(size_t n, int depth) {
    done = appended = 0;
    pending.clear();
    if (n > 0) pending.push_back({0, n, depth});
  }
  // several consecutive ranges of buf (leftmost first), each isolated by the sort with its own depth budget
  void reset_ranges(const R* r, int count) {
    done = appended = 0;
    pending.clear();
    for (int i = count - 1; i >= 0; --i) if (r[i].last > r[i].first) pending.push_back(r[i]);
  }
  void clear() { buf.clear(); reset(0, 0); }
  // buf[0, p) becomes the SET std::sort would leave there (in no particular order); whatever the split
  // finalises beyond p is reflected in `done`.
  void split_set(size_t p) {
    while (!pending.empty() && pending.back().first < p) {
      R r = pending.back();
      pending.pop_back();
      if (r.last <= p) continue;   // wholly inside the set
      if (r.last - r.first <= 16 || r.depth == 0) {
        finish_range(buf.data() + r.first, buf.data() + r.last, r.depth);
        done = r.last;
        break;   // ranges are disjoint and ordered: nothing else starts before p
      }
      --r.depth;
      const size_t cut = static_cast<size_t>(partition_pivot(buf.data() + r.first, buf.data() + r.last) - buf.data());
      pending.push_back({cut, r.last, r.depth});
      pending.push_back({r.first, cut, r.depth});
    }
    done = std::max(done, p);
  }
  // finalises the leftmost pending range (at most 16 entries, or a heap-sorted one); false: nothing pending
  bool advance() {
    if (pending.empty()) return false;
    R r = pending.back();
    pending.pop_back();
    while (r.last - r.first > 16 && r.depth > 0) {
      --r.depth;
      const size_t cut = static_cast<size_t>(partition_pivot(buf.data() + r.first, buf.data() + r.last) - buf.data());
      pending.push_back({cut, r.last, r.depth});
      r.last = cut;
    }
    finish_range(buf.data() + r.first, buf.data() + r.last, r.depth);
    done = r.last;
    return true;
  }
};

}  // namespace exact_sort

// ---- the candidate as a JPEG, coded on the device ------------------------------------------------
// Header (DQT/SOF/DHT/SOS with the clustered Huffman codes) on the host, scan on the GPU
// (gzb_candidate_entropy_code). Only the size is needed to score a candidate; the bytes are fetched
// for a candidate that becomes the best so far.
struct DeviceJpeg {
  std::string header;
  uint64_t scan_bytes = 0, ff_bytes = 0;
  size_t size() const { return header.size() + static_cast<size_t>(scan_bytes + ff_bytes) + 2; }
};

void histogram_from_counts(const uint32_t* counts, int n, Histogram* h) {
  *h = Histogram();
  for (int i = 0; i < n; ++i) if (counts[i]) h->add(i, static_cast<int>(counts[i]));
}

// SaveToJpegData drops chroma planes that are entirely zero (output_image.cc:588): every block then
// codes a zero DC difference and a lone end-of-block.
int ncomp_from_histograms(const Histogram* dc, const Histogram* ac, int nblocks) {
  (void)nblocks;
  // a chroma histogram may also be empty: the back end counts only the components SaveToJpegData
  // keeps at its start (processor.cc:744-750), like the reference
  for (int c = 1; c < 3; ++c)
    for (int i = 1; i + 1 < Histogram::kSize; ++i)
      if (dc[c].counts[i] != 0 || ac[c].counts[i] != 0) return 3;
  return 1;
}

// Codes the resident candidate. dc_hist/ac_hist: its per-component histograms if the caller has them
// (the back end maintains them incrementally), else they are counted on the device.
bool device_code_candidate(gzb_ctx* ctx, int w, int h, int nblocks, const int q[3][64], bool input_tables,
                           const Histogram* dc_hist, const Histogram* ac_hist, DeviceJpeg* out, bool yuv420 = false) {
  Histogram hd[3], ha[3];
  auto count_on_device = [&](int ncomp) {
    uint32_t dc[48], ac[768];
    if (gzb_candidate_symbol_histograms_n(ctx, nullptr, ncomp, dc, ac) != GZB_OK) return false;
    for (int c = 0; c < 3; ++c) {
      histogram_from_counts(dc + 16 * c, 16, &hd[c]);
      histogram_from_counts(ac + 256 * c, 256, &ha[c]);
    }
    dc_hist = hd;
    ac_hist = ha;
    return true;
  };
  if ((!dc_hist || !ac_hist) && !count_on_device(3)) return false;
  Frame f;
  f.width = w; f.height = h; f.bw = (w + 7) / 8; f.bh = (h + 7) / 8;
  f.yuv420 = yuv420;
  if (input_tables) {
    f.ncomp = 3;
    gzb::jpeg::frame_set_quant_input(&f, q);
  } else {
    f.ncomp = ncomp_from_histograms(dc_hist, ac_hist, nblocks);
    // a 4:2:0 candidate without chroma is written as a plain grey file: its luma DC differences then
    // follow image block order instead of MCU order
    if (yuv420 && f.ncomp == 1 && !count_on_device(1)) return false;
    gzb::jpeg::frame_set_quant(&f, q);
  }
  gzb::jpeg::CodeTable dct[3], act[3];
  gzb::jpeg::write_jpeg_header(f, dc_hist, ac_hist, &out->header, dct, act);
  uint16_t dc_code[3][16], ac_code[3][256];
  uint8_t dc_len[3][16], ac_len[3][256];
  memset(dc_code, 0, sizeof(dc_code)); memset(ac_code, 0, sizeof(ac_code));
  memset(dc_len, 0, sizeof(dc_len)); memset(ac_len, 0, sizeof(ac_len));
  for (int c = 0; c < f.ncomp; ++c) {
    for (int i = 0; i < 16; ++i) { dc_code[c][i] = dct[c].code[i]; dc_len[c][i] = dct[c].depth[i] == 255 ? 0 : dct[c].depth[i]; }
    for (int i = 0; i < 256; ++i) { ac_code[c][i] = act[c].code[i]; ac_len[c][i] = act[c].depth[i] == 255 ? 0 : act[c].depth[i]; }
  }
  // the scan's length follows from the histograms and the code lengths (every symbol is counted twice in a
  // Histogram): DC symbols carry `symbol` extra bits, AC symbols `symbol & 15`
  uint64_t total_bits = 0;
  for (int c = 0; c < f.ncomp; ++c) {
    for (int i = 0; i < 16; ++i) total_bits += static_cast<uint64_t>(dc_hist[c].counts[i] / 2) * (dc_len[c][i] + i);
    for (int i = 0; i < 256; ++i) total_bits += static_cast<uint64_t>(ac_hist[c].counts[i] / 2) * (ac_len[c][i] + (i & 15));
  }
  return gzb_candidate_entropy_code_sized(ctx, f.ncomp, &dc_code[0][0], &dc_len[0][0], &ac_code[0][0], &ac_len[0][0],
                                          total_bits, &out->scan_bytes, &out->ff_bytes) == GZB_OK;
}

// header + byte-stuffed scan + EOI
void assemble_jpeg(const std::string& header, const uint8_t* scan, size_t scan_bytes, size_t ff_bytes, std::string* out) {
  out->resize(header.size() + scan_bytes + ff_bytes + 2);
  memcpy(&(*out)[0], header.data(), header.size());
  uint8_t* o = reinterpret_cast<uint8_t*>(&(*out)[0]) + header.size();
  const uint8_t* p = scan;
  const uint8_t* end = scan + scan_bytes;
  while (p < end) {
    const uint8_t* ff = static_cast<const uint8_t*>(memchr(p, 0xff, static_cast<size_t>(end - p)));
    const size_t n = static_cast<size_t>((ff ? ff + 1 : end) - p);
    memcpy(o, p, n);
    o += n;
    p += n;
    if (ff) *o++ = 0;
  }
  *o++ = 0xff;
  *o++ = 0xd9;
}

bool device_fetch_jpeg(gzb_ctx* ctx, const DeviceJpeg& dj, std::string* out) {
  std::vector<uint8_t> scan(static_cast<size_t>(dj.scan_bytes) + 8);
  if (gzb_candidate_fetch_scan(ctx, scan.data(), dj.scan_bytes) != GZB_OK) return false;
  assemble_jpeg(dj.header, scan.data(), static_cast<size_t>(dj.scan_bytes), static_cast<size_t>(dj.ff_bytes), out);
  return true;
}

// ---- encoder state ----------------------------------------------------------------------------
struct Encoder {
  int w = 0, h = 0, bw = 0, bh = 0, nb = 0;
  // Coefficient arrays per component (orig / idx): 4:4:4 = bw x bh blocks each; after the 4:2:0
  // downsampling the luma plane is MCU-padded (2*ceil(w/16) x 2*ceil(h/16)) and the chroma planes have
  // ceil(w/16) x ceil(h/16) blocks (the layout of JPEGData after SaveToJpegData).
  bool yuv420 = false;
  int cbw[3] = {0, 0, 0}, cbh[3] = {0, 0, 0};
  size_t cnb[3] = {0, 0, 0};
  bool try_420 = false, force_420 = false;   // Params (guetzli/processor.h:34-42)
  void set_geometry(bool m420) {
    yuv420 = m420;
    const int mcw = (w + 15) / 16, mch = (h + 15) / 16;
    for (int c = 0; c < 3; ++c) {
      cbw[c] = !m420 ? bw : (c == 0 ? 2 * mcw : mcw);
      cbh[c] = !m420 ? bh : (c == 0 ? 2 * mch : mch);
      cnb[c] = static_cast<size_t>(cbw[c]) * cbh[c];
    }
  }
  const char* frame_type() const { return yuv420 ? "f112222" : "f111111"; }   // OutputImage::FrameTypeStr
  int nthreads = 1;
  std::unique_ptr<gzb::WorkerPool> pool;
  gzb::jpeg::WriteTimers wt;
  float target = 0.f;
  gzb_ctx* ctx = nullptr;
  int quant[3][64];
  std::string best_jpeg;
  double best_score = -1;
  gzb_encode_stats st{};
  std::string trace;
  bool want_trace = false, ran = false;
  float distance = 0.f;
  gzb::Group group;                 // multi-GPU candidate sharding (gzb_encoder_set_group)
  size_t best_size = 0;             // size of the best file so far (MaybeOutput, processor.cc:151-160)
  bool defer_trial_bytes = false;   // a SelectQuantMatrix trial that becomes the best is rebuilt at the end instead of fetched
  bool best_remote = false;         // the best file so far is not in best_jpeg: rebuild it from best_trial at the end
  gzb::Trial best_trial{};          // ... and this is the trial that produced it
  int search_rounds = 0, search_trials = 0;
  unsigned long long code_gen = 0;  // counts device codings: tells whether a trial's scan is still resident

  void log(const char* fmt, ...) {
    if (!want_trace) return;
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    trace += buf;
  }

  // MaybeOutput (processor.cc:151-160) for a trial that may have been evaluated on another rank.
  void maybe_output_trial(const gzb::Trial& t, const gzb::TrialOutcome& o);

  bool compare_begin() { t_cmp = now_ms(); return gzb_compare_begin(ctx) == GZB_OK; }
  bool compare_end() {
    if (gzb_compare_end(ctx, &distance) != GZB_OK) return false;
    st.device_compare_ms += gzb_last_device_ms(ctx);
    st.compare_wall_ms += now_ms() - t_cmp;
    st.num_compares++;
    return true;
  }
  double t_cmp = 0;
  bool compare(bool quiet = false) {
    const double t0 = now_ms();
    if (gzb_compare(ctx, &distance) != GZB_OK) return false;
    st.device_compare_ms += gzb_last_device_ms(ctx);
    st.compare_wall_ms += now_ms() - t0;
    st.num_compares++;
    if (!quiet) log(" BA[100.00%%] D[%6.4f]", distance);
    return true;
  }

  // img.CopyFromJpegData(jpg) ; img.ApplyGlobalQuantization(q) on the device-resident candidate
  bool set_global_quant(const int q[3][64]) {
    memcpy(quant, q, sizeof(quant));
    return gzb_quantize_from_jpeg(ctx, &q[0][0]) == GZB_OK;
  }
};


// ScoreJPEG (guetzli/score.cc:23-41) for a given distance (the writer stage scores a file after the
// comparator has moved on to the next iteration).
double score_jpeg(double distance, int size, double target) {
  const double diff = distance - target;
  if (diff <= 0.0) return size;
  const double ex = 50 * diff;
  if (ex > 10) return 1e30 * std::exp(10.0) * diff + size;
  return std::exp(ex) * size;
}

void Encoder::maybe_output_trial(const gzb::Trial& t, const gzb::TrialOutcome& o) {
  const double score = score_jpeg(o.distance, static_cast<int>(o.jpg_size), target);
  log(" Score[%.4f]", score);
  if (score < best_score || best_score < 0) {
    best_score = score;
    best_size = static_cast<size_t>(o.jpg_size);
    if (defer_trial_bytes) {
      // Only the best file of the WHOLE search is ever read. A trial of SelectQuantMatrix is rarely that file
      // (the back end almost always improves on it), and its bytes -- tens of megabytes for the first ones --
      // can be produced again from its matrix at the end, so they are not fetched now.
      best_jpeg.clear();
      best_remote = true;
      best_trial = t;
    } else if (o.owner == group.rank && (o.resident_gen == 0 || o.resident_gen == code_gen)) {
      if (o.resident_gen != 0) {   // still on the device: fetch it now
        DeviceJpeg dj;
        dj.header = o.jpeg;
        dj.scan_bytes = o.scan_bytes;
        dj.ff_bytes = o.ff_bytes;
        if (!device_fetch_jpeg(ctx, dj, &best_jpeg)) { best_jpeg.clear(); best_remote = true; best_trial = t; }
        else best_remote = false;
      } else {
        assemble_jpeg(o.jpeg, reinterpret_cast<const uint8_t*>(o.scan.data()), o.scan.size(),
                      static_cast<size_t>(o.jpg_size) - o.jpeg.size() - o.scan.size() - 2, &best_jpeg);
        best_remote = false;
      }
    } else {  // the bytes live on another rank: remember how to rebuild them if they stay the best
      best_jpeg.clear();
      best_remote = true;
      best_trial = t;
    }
    log(" (*)");
  }
  log("\n");
}

}  // namespace

// =============================================================================================
extern "C" {

double gzb_butteraugli_score_for_quality(double quality) {
  // guetzli/quality.cc:31-85 (median butteraugli scores of libjpeg-turbo output per quality level)
  static const double kScore[] = {
      2.810761, 2.729300, 2.689687, 2.636811, 2.547863, 2.525400, 2.473416, 2.366133, 2.338078, 2.318654,
      2.201674, 2.145517, 2.087322, 2.009328, 1.945456, 1.900112, 1.805701, 1.750194, 1.644175, 1.562165,
      1.473608, 1.382021, 1.294298, 1.185402, 1.066781, 0.971769, 0.852901, 0.724544, 0.611302, 0.443185,
      0.211578, 0.209462, 0.207346, 0.205230, 0.203114, 0.200999, 0.198883, 0.196767, 0.194651, 0.192535,
      0.190420, 0.190420};
  if (quality < 70) quality = 70;
  if (quality > 110) quality = 110;
  const int index = static_cast<int>(quality);
  const double mix = quality - index;
  return kScore[index - 70] * (1 - mix) + kScore[index - 70 + 1] * mix;
}

void gzb_free(void* p) { free(p); }

static thread_local std::string g_encode_err;
const char* gzb_encode_last_error(void) { return g_encode_err.c_str(); }

int gzb_rgb_to_jpeg_coeffs(const uint8_t* rgb, int width, int height, int16_t* c0, int16_t* c1, int16_t* c2) {
  if (!rgb || !c0 || !c1 || !c2 || width <= 0 || height <= 0 || width >= (1 << 16) || height >= (1 << 16))
    return GZB_ERR_BAD_ARG;
  int16_t* out[3] = {c0, c1, c2};
  const int bw = (width + 7) / 8, bh = (height + 7) / 8;
  gzb::WorkerPool pool(static_cast<int>(std::max(1u, std::min(16u, std::thread::hardware_concurrency()))));
  parallel_rows(bh, &pool, [&](int y0, int y1) { rgb_to_coeffs_rows(rgb, width, height, bw, y0, y1, out); });
  return GZB_OK;
}

long gzb_write_jpeg(const int16_t* c0, const int16_t* c1, const int16_t* c2, int width, int height,
                    const int* q192, int input_tables, int host_threads, uint8_t* out, long cap) {
  if (!c0 || !c1 || !c2 || !q192 || width <= 0 || height <= 0) return GZB_ERR_BAD_ARG;
  const int bw = (width + 7) / 8, bh = (height + 7) / 8;
  const size_t n = static_cast<size_t>(bw) * bh * 64;
  const int16_t* src[3] = {c0, c1, c2};
  std::vector<int16_t> idx[3];
  int q[3][64];
  memcpy(q, q192, sizeof(q));
  bool chroma = false;
  for (int c = 0; c < 3; ++c) {
    idx[c].resize(n);
    for (size_t i = 0; i < n; ++i) {
      if (q[c][i & 63] <= 0) return GZB_ERR_BAD_ARG;
      idx[c][i] = static_cast<int16_t>(src[c][i] / q[c][i & 63]);
      if (c > 0 && src[c][i] != 0) chroma = true;
    }
  }
  Frame f;
  f.width = width; f.height = height; f.bw = bw; f.bh = bh;
  f.ncomp = (chroma || input_tables) ? 3 : 1;
  for (int c = 0; c < 3; ++c) f.coeffs[c] = idx[c].data();
  if (input_tables) gzb::jpeg::frame_set_quant_input(&f, q); else gzb::jpeg::frame_set_quant(&f, q);
  std::string s;
  gzb::WorkerPool pool(host_threads > 0 ? host_threads : 1);
  gzb::jpeg::write_jpeg(f, &s, &pool);
  if (out && static_cast<long>(s.size()) <= cap) memcpy(out, s.data(), s.size());
  return static_cast<long>(s.size());
}

// Micro-benchmark hook: repeated WriteJpeg of one frame of quantised indices; returns ms per write
// and fills parts[4] = {hist, code, encode, stitch} ms per write.
double gzb_bench_write_jpeg(const int16_t* c0, const int16_t* c1, const int16_t* c2, int width, int height,
                            int host_threads, int reps, int with_hist, double* parts) {
  const int bw = (width + 7) / 8, bh = (height + 7) / 8;
  Frame f;
  f.width = width; f.height = height; f.bw = bw; f.bh = bh; f.ncomp = 3;
  f.coeffs[0] = c0; f.coeffs[1] = c1; f.coeffs[2] = c2;
  int q[3][64];
  for (int c = 0; c < 3; ++c) for (int k = 0; k < 64; ++k) q[c][k] = 1 + c;
  gzb::jpeg::frame_set_quant(&f, q);
  gzb::WorkerPool pool(host_threads > 0 ? host_threads : 1);
  Histogram dc[3], ac[3];
  gzb::jpeg::build_histograms(f, dc, ac, &pool);
  gzb::jpeg::WriteTimers tm;
  std::string s;
  const double t0 = now_ms();
  for (int r = 0; r < reps; ++r) gzb::jpeg::write_jpeg(f, &s, &pool, with_hist ? dc : nullptr, with_hist ? ac : nullptr, &tm);
  const double dt = (now_ms() - t0) / reps;
  if (parts) { parts[0] = tm.hist_ms / reps; parts[1] = tm.code_ms / reps; parts[2] = tm.encode_ms / reps; parts[3] = tm.stitch_ms / reps; }
  return dt;
}

// Micro-benchmark hook: ComputeEntropyCodes-style clustering of a frame's three AC histograms with a
// small perturbation per repetition; returns microseconds per call.
double gzb_bench_cluster(const int16_t* c0, const int16_t* c1, const int16_t* c2, int width, int height, int reps,
                         int use_cache) {
  gzb::jpeg::HuffCache caches[5];
  const int bw = (width + 7) / 8, bh = (height + 7) / 8;
  Frame f;
  f.width = width; f.height = height; f.bw = bw; f.bh = bh; f.ncomp = 3;
  f.coeffs[0] = c0; f.coeffs[1] = c1; f.coeffs[2] = c2;
  Histogram dc[3], ac[3];
  gzb::jpeg::build_histograms(f, dc, ac, nullptr);
  size_t sink = 0;
  const double t0 = now_ms();
  for (int r = 0; r < reps; ++r) {
    ac[r % 3].counts[(r * 7) % 200 + 1] += 2;
    Histogram clustered[3] = {ac[0], ac[1], ac[2]};
    size_t num = 3;
    int indexes[4];
    uint8_t cd[3 * Histogram::kSize];
    sink += gzb::jpeg::cluster_histograms(clustered, &num, indexes, cd, use_cache ? caches : nullptr);
  }
  const double us = (now_ms() - t0) * 1e3 / reps;
  return sink == 12345 ? -us : us;
}

// Test hook: huffman_depths on a 257-entry histogram (optionally through a warm HuffCache).
void gzb_test_huffman_depths(const uint32_t* counts257, uint8_t* depth257, const uint32_t* warm_counts257) {
  memset(depth257, 0, 257);
  if (warm_counts257) {
    gzb::jpeg::HuffCache cache;
    uint8_t tmp[257] = {0};
    gzb::jpeg::huffman_depths(warm_counts257, 257, 16, tmp, &cache);
    gzb::jpeg::huffman_depths(counts257, 257, 16, depth257, &cache);
  } else {
    gzb::jpeg::huffman_depths(counts257, 257, 16, depth257, nullptr);
  }
}

int gzb_write_candidate_jpeg(gzb_ctx* ctx, const int* q192, int input_tables, uint8_t* out, size_t cap, size_t* size_out) {
  if (!ctx || !q192 || !size_out) return GZB_ERR_BAD_ARG;
  int w = 0, h = 0;
  if (gzb_image_size(ctx, &w, &h) != GZB_OK) return GZB_ERR_BAD_ARG;
  int q[3][64];
  memcpy(q, q192, sizeof(q));
  uint32_t dc[48], ac[768];
  int rc = gzb_candidate_symbol_histograms(ctx, q192, dc, ac);  // also makes q192 the device's matrix
  if (rc != GZB_OK) return rc;
  Histogram hd[3], ha[3];
  for (int c = 0; c < 3; ++c) {
    histogram_from_counts(dc + 16 * c, 16, &hd[c]);
    histogram_from_counts(ac + 256 * c, 256, &ha[c]);
  }
  DeviceJpeg dj;
  int factor = 1;
  gzb_component_dims(ctx, 1, nullptr, nullptr, &factor);
  if (!device_code_candidate(ctx, w, h, ((w + 7) / 8) * ((h + 7) / 8), q, input_tables != 0, hd, ha, &dj, factor == 2)) return GZB_ERR_CUDA;
  *size_out = dj.size();
  if (out && cap >= dj.size()) {
    std::string bytes;
    if (!device_fetch_jpeg(ctx, dj, &bytes)) return GZB_ERR_CUDA;
    memcpy(out, bytes.data(), bytes.size());
  }
  return GZB_OK;
}

// Test hook (CPU only): the speculative SelectQuantMatrix search with a caller-provided evaluator.
// visited: rows of {original, heuristic score, distance, jpg_size}; info: {best.dist_ok, rounds,
// evaluated on this rank, evaluated by the group}.
int gzb_test_quant_search_mode(int rank, int world, gzb_allgather_fn allgather, void* user,
                               int (*eval_fn)(void*, int, const int*, float*, uint64_t*), void* eval_user, float target,
                               double* visited, int cap, int* nvisited, int* best_q192, int* info, int batch, int mode);
int gzb_test_quant_search(int rank, int world, gzb_allgather_fn allgather, void* user,
                          int (*eval_fn)(void*, int, const int*, float*, uint64_t*), void* eval_user, float target,
                          double* visited, int cap, int* nvisited, int* best_q192, int* info, int batch) {
  return gzb_test_quant_search_mode(rank, world, allgather, user, eval_fn, eval_user, target, visited, cap, nvisited,
                                    best_q192, info, batch, 0);
}
// mode 2: the search of a downsampled (YUV420) image -- no original, generator starts at score 0.
int gzb_test_quant_search_mode(int rank, int world, gzb_allgather_fn allgather, void* user,
                               int (*eval_fn)(void*, int, const int*, float*, uint64_t*), void* eval_user, float target,
                               double* visited, int cap, int* nvisited, int* best_q192, int* info, int batch, int mode) {
  gzb::Group g;
  g.rank = rank; g.world = world; g.allgather = allgather; g.user = user;
  gzb::QuantSearch search(g, target, batch, mode);
  int n = 0;
  const bool ok = search.run(
      [&](const std::vector<gzb::Trial>& ts, std::vector<gzb::TrialOutcome>* os) {
        for (size_t j = 0; j < ts.size(); ++j) {
          float d = 0.f;
          uint64_t sz = 0;
          if (eval_fn(eval_user, ts[j].original, &ts[j].q[0][0], &d, &sz) != 0) return false;
          (*os)[j].distance = d;
          (*os)[j].jpg_size = sz;
        }
        return true;
      },
      [&](const gzb::Trial& t, const gzb::TrialOutcome& o) {
        if (n < cap) {
          visited[4 * n] = t.original;
          visited[4 * n + 1] = quant_heuristic_score(t.q);
          visited[4 * n + 2] = o.distance;
          visited[4 * n + 3] = static_cast<double>(o.jpg_size);
        }
        ++n;
      });
  if (!ok) return GZB_ERR_STATE;
  *nvisited = std::min(n, cap);
  memcpy(best_q192, search.best().q, sizeof(int) * 192);
  info[0] = search.best().dist_ok ? 1 : 0;
  info[1] = search.rounds();
  info[2] = search.evaluated_here();
  info[3] = search.evaluated_total();
  return GZB_OK;
}

// WorkerPool stress (CPU test hook): `jobs` back-to-back run() calls of varying width on a pool of
// `threads`; every task adds its index to a per-job sum. Returns the number of jobs whose sum is wrong.
long gzb_test_pool_stress(int threads, int jobs) {
  gzb::WorkerPool pool(threads);
  long bad = 0;
  for (int j = 0; j < jobs; ++j) {
    const int n = 1 + (j * 7919) % 67;
    std::atomic<long> sum(0);
    std::vector<int> hits(n, 0);
    const std::function<void(int)> fn = [&](int i) { sum.fetch_add(i + 1); ++hits[i]; if ((j & 255) == 0 && i == 0) std::this_thread::yield(); };
    if (j % 3 == 2) {
      // the two-part form: the caller does something else (here: a nested run(), which must then stay on this
      // thread) between handing the items out and joining
      pool.begin(n, fn);
      std::atomic<int> inner(0);
      pool.run(3, [&](int) { inner.fetch_add(1); });
      pool.finish();
      if (inner.load() != 3) ++bad;
    } else {
      pool.run(n, fn);
    }
    long want = static_cast<long>(n) * (n + 1) / 2;
    bool once = true;
    for (int h : hits) once = once && h == 1;
    bad += (sum.load() != want || !once) ? 1 : 0;
  }
  return bad;
}

// Test hooks (CPU): the restated introsort must give std::sort's permutation. gzb_test_exact_sort runs it
// whole; gzb_test_exact_sort_split partitions the top `levels` levels one range at a time, the way the
// device kernel does (leftmost first, right-hand ranges pending), and finishes the pieces separately.
void gzb_test_exact_sort(int* first, float* second, size_t n) {
  std::vector<OrderEntry> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(first[i], second[i]);
  exact_sort::finish_range(v.data(), v.data() + n, exact_sort::depth_budget(n));
  for (size_t i = 0; i < n; ++i) { first[i] = v[i].first; second[i] = v[i].second; }
}
void gzb_test_exact_sort_split(int* first, float* second, size_t n, size_t small_max) {
  std::vector<OrderEntry> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(first[i], second[i]);
  struct R { size_t first, last; int depth; };
  std::vector<R> pending;
  if (n > 0) pending.push_back({0, n, exact_sort::depth_budget(n)});
  while (!pending.empty()) {
    R r = pending.back();
    pending.pop_back();
    if (r.last - r.first <= std::max<size_t>(small_max, 16) || r.depth == 0) {
      if (r.last - r.first > 16 && r.depth == 0) exact_sort::heap_sort(v.data() + r.first, v.data() + r.last);
      else exact_sort::finish_range(v.data() + r.first, v.data() + r.last, r.depth);
      continue;
    }
    --r.depth;
    const size_t cut = static_cast<size_t>(exact_sort::partition_pivot(v.data() + r.first, v.data() + r.last) - v.data());
    pending.push_back({cut, r.last, r.depth});
    pending.push_back({r.first, cut, r.depth});
  }
  for (size_t i = 0; i < n; ++i) { first[i] = v[i].first; second[i] = v[i].second; }
}
// heap path alone, and std::partial_sort(first, last, last) for the checker
void gzb_test_exact_heap_sort(int* first, float* second, size_t n, int use_std) {
  std::vector<OrderEntry> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(first[i], second[i]);
  if (use_std) std::partial_sort(v.begin(), v.end(), v.end(), OrderLess());
  else exact_sort::heap_sort(v.data(), v.data() + n);
  for (size_t i = 0; i < n; ++i) { first[i] = v[i].first; second[i] = v[i].second; }
}
// HostLazy: split_set(p), then advance() to the end; [p, n) must be std::sort's arrangement, [0, p) its set.
void gzb_test_host_lazy(int* first, float* second, size_t n, size_t p) {
  exact_sort::HostLazy lz;
  lz.buf.resize(n);
  for (size_t i = 0; i < n; ++i) lz.buf[i] = std::make_pair(first[i], second[i]);
  lz.reset(n, exact_sort::depth_budget(n));
  if (p > 0) lz.split_set(p);
  while (lz.advance()) {}
  for (size_t i = 0; i < n; ++i) { first[i] = lz.buf[i].first; second[i] = lz.buf[i].second; }
}
// The multi-range hand-over on the host side: the restated partition plays the device (leftmost first, right-hand
// ranges pending, down to small_max), and the short ranges are given to HostLazy `group` consecutive ones at a time
// (HostLazy::reset_ranges), the first group with the prefix p split off as a set.
void gzb_test_host_lazy_ranges(int* first, float* second, size_t n, size_t p, size_t small_max, int group) {
  std::vector<OrderEntry> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(first[i], second[i]);
  typedef exact_sort::HostLazy::R R;
  std::vector<R> pending, small;
  if (n > 0) pending.push_back({0, n, exact_sort::depth_budget(n)});
  while (!pending.empty()) {
    R r = pending.back();
    pending.pop_back();
    if (r.last <= p) continue;   // consumed as a set
    if (r.last - r.first <= std::max<size_t>(small_max, 16) || r.depth == 0) { small.push_back(r); continue; }
    --r.depth;
    const size_t cut = static_cast<size_t>(exact_sort::partition_pivot(v.data() + r.first, v.data() + r.last) - v.data());
    pending.push_back({cut, r.last, r.depth});
    pending.push_back({r.first, cut, r.depth});
  }
  size_t have_end = p;
  for (size_t g = 0; g < small.size();) {
    size_t ge = g + 1;
    while (ge < small.size() && static_cast<int>(ge - g) < group && small[ge].first == small[ge - 1].last) ++ge;
    const size_t rf = small[g].first, rl = small[ge - 1].last;
    exact_sort::HostLazy lz;
    lz.buf.assign(v.begin() + rf, v.begin() + rl);
    std::vector<R> rel;
    for (size_t k = g; k < ge; ++k) rel.push_back({small[k].first - rf, small[k].last - rf, small[k].depth});
    lz.reset_ranges(rel.data(), static_cast<int>(rel.size()));
    if (have_end > rf) lz.split_set(have_end - rf);
    while (lz.advance()) {}
    std::copy(lz.buf.begin(), lz.buf.end(), v.begin() + rf);
    have_end = rl;
    g = ge;
  }
  for (size_t i = 0; i < n; ++i) { first[i] = v[i].first; second[i] = v[i].second; }
}
int gzb_test_sort_emulation_ok(void) { return exact_sort::emulation_ok() ? 1 : 0; }
// Test hook (GPU): sorts `entries` through the back end's device path -- long ranges partitioned by
// the k_be_* kernels, short ones finished by exact_sort -- optionally consuming [0, prefix) as a set first (those
// entries come back in unspecified order). Everything from `prefix` on must equal std::sort's arrangement.
int gzb_test_device_sort_depth(gzb_ctx* ctx, gzb_order_entry* entries, size_t n, size_t prefix, int small_max, int depth);
int gzb_test_device_sort(gzb_ctx* ctx, gzb_order_entry* entries, size_t n, size_t prefix, int small_max) {
  return gzb_test_device_sort_depth(ctx, entries, n, prefix, small_max, -1);
}
void gzb_test_exact_sort_depth(int* first, float* second, size_t n, int depth) {
  std::vector<OrderEntry> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(first[i], second[i]);
  exact_sort::finish_range(v.data(), v.data() + n, depth);
  for (size_t i = 0; i < n; ++i) { first[i] = v[i].first; second[i] = v[i].second; }
}
int gzb_test_device_sort_depth(gzb_ctx* ctx, gzb_order_entry* entries, size_t n, size_t prefix, int small_max, int depth) {
  if (!ctx || !entries || n == 0 || prefix > n) return GZB_ERR_BAD_ARG;
  int rc = gzb_be_test_load_order_depth(ctx, entries, n, depth);
  if (rc != GZB_OK) return rc;
  std::vector<OrderEntry> buf(GZB_BE_MAX_ENTRIES), big;
  size_t have_end = prefix;
  // small_max < 0: several ranges per round trip (gzb_be_select_ranges), as many as fit
  const size_t want = small_max < 0 ? GZB_BE_MAX_ENTRIES : 0;
  if (small_max < 0) small_max = -small_max;
  for (;;) {
    int status = 0, nranges = 0;
    gzb_be_range rr[GZB_BE_MAX_RANGES];
    rc = gzb_be_select_ranges(ctx, have_end, small_max, have_end + want, &status, &nranges, rr, reinterpret_cast<gzb_order_entry*>(buf.data()));
    if (rc != GZB_OK) return rc;
    if (status == 3) break;
    if (nranges < 1 || (want == 0 && nranges != 1)) return GZB_ERR_STATE;
    const size_t rf = static_cast<size_t>(rr[0].first), rl = static_cast<size_t>(rr[nranges - 1].last);
    OrderEntry* r0 = buf.data();
    if (status == 2) {
      big.resize(rl - rf);
      rc = gzb_be_fetch_order(ctx, rf, reinterpret_cast<gzb_order_entry*>(big.data()), rl - rf);
      if (rc != GZB_OK) return rc;
      exact_sort::heap_sort(big.data(), big.data() + big.size());
      r0 = big.data();
    } else {
      for (int i = 0; i < nranges; ++i) {
        if (i > 0 && rr[i].first != rr[i - 1].last) return GZB_ERR_STATE;   // consecutive
        exact_sort::finish_range(r0 + (rr[i].first - rf), r0 + (rr[i].last - rf), rr[i].depth);
      }
    }
    rc = gzb_be_store_order(ctx, rf, reinterpret_cast<const gzb_order_entry*>(r0), rl - rf);
    if (rc != GZB_OK) return rc;
    have_end = rl;
  }
  return gzb_be_fetch_order(ctx, 0, entries, n);
}

void gzb_test_std_sort(int* first, float* second, size_t n) {
  std::vector<OrderEntry> v(n);
  for (size_t i = 0; i < n; ++i) v[i] = std::make_pair(first[i], second[i]);
  std::sort(v.begin(), v.end(), [](const OrderEntry& a, const OrderEntry& b) { return a.second < b.second; });
  for (size_t i = 0; i < n; ++i) { first[i] = v[i].first; second[i] = v[i].second; }
}

struct gzb_encoder { Encoder e; double t_start = 0; };

int gzb_encoder_create(int device, const uint8_t* rgb, int width, int height, float butteraugli_target,
                       int host_threads, gzb_encoder** out) {
  g_encode_err.clear();
  if (!rgb || !out) { g_encode_err = "gzb_encoder_create: null argument"; return GZB_ERR_BAD_ARG; }
  *out = nullptr;
  if (butteraugli_target > 2.0f) {  // processor.cc:939-945
    g_encode_err = "gzb_encoder_create: quality below 84 is refused (butteraugli target > 2.0)";
    return GZB_ERR_BAD_ARG;
  }
  if (width < 32 || height < 32) {
    g_encode_err = "gzb_encoder_create: images smaller than 32x32 skip butteraugli in the reference; not accelerated";
    return GZB_ERR_TOO_SMALL;
  }
  if (width >= (1 << 16) || height >= (1 << 16)) { g_encode_err = "gzb_encoder_create: image too large"; return GZB_ERR_BAD_ARG; }
  gzb_encoder* enc = new gzb_encoder;
  enc->t_start = now_ms();
  Encoder& e = enc->e;
  e.w = width; e.h = height; e.bw = (width + 7) / 8; e.bh = (height + 7) / 8; e.nb = e.bw * e.bh;
  e.set_geometry(false);
  e.target = butteraugli_target;
  const unsigned hc = std::thread::hardware_concurrency();
  e.nthreads = host_threads > 0 ? host_threads : static_cast<int>(std::max(1u, std::min(16u, hc)));
  e.pool.reset(new gzb::WorkerPool(e.nthreads));
  const double t_create = now_ms();
  int rc = gzb_create(device, width, height, rgb, butteraugli_target, &e.ctx);
  if (rc != GZB_OK) { g_encode_err = gzb_last_error(nullptr); delete enc; return rc; }
  e.st.create_ms = now_ms() - t_create;
  // EncodeRGBToJpeg (q = 1): on the device, from the image gzb_create has just uploaded. The coefficients
  // never come to the host: the whole search reads them in HBM.
  bool ok = true;
  {
    const double t0 = now_ms();
    ok = gzb_rgb_to_jpeg_coeffs_device(e.ctx) == GZB_OK;
    e.st.host_frontend_ms = now_ms() - t0;
  }
  if (!ok) {
    g_encode_err = gzb_last_error(e.ctx);
    gzb_destroy(e.ctx);
    delete enc;
    return GZB_ERR_CUDA;
  }
  e.st.prepare_ms = now_ms() - enc->t_start;
  *out = enc;
  return GZB_OK;
}

void gzb_encoder_destroy(gzb_encoder* enc) {
  if (!enc) return;
  if (enc->e.ctx) gzb_destroy(enc->e.ctx);
  delete enc;
}

gzb_ctx* gzb_encoder_context(gzb_encoder* enc) { return enc ? enc->e.ctx : nullptr; }

int gzb_encoder_set_params(gzb_encoder* enc, int try_420, int force_420) {
  if (!enc) return GZB_ERR_BAD_ARG;
  if (enc->e.ran) { g_encode_err = "gzb_encoder_set_params: the encoder has already run"; return GZB_ERR_STATE; }
  enc->e.try_420 = try_420 != 0;
  enc->e.force_420 = force_420 != 0;
  return GZB_OK;
}

int gzb_encoder_set_group(gzb_encoder* enc, int rank, int world, gzb_allgather_fn allgather, void* user) {
  if (!enc || world < 1 || rank < 0 || rank >= world || (world > 1 && !allgather)) {
    g_encode_err = "gzb_encoder_set_group: bad argument";
    return GZB_ERR_BAD_ARG;
  }
  if (enc->e.ran) { g_encode_err = "gzb_encoder_set_group: the encoder has already run"; return GZB_ERR_STATE; }
  enc->e.group.rank = rank;
  enc->e.group.world = world;
  enc->e.group.allgather = allgather;
  enc->e.group.user = user;
  return GZB_OK;
}

int gzb_encoder_set_group_device(gzb_encoder* enc, gzb_allgather_device_fn allgather_device) {
  if (!enc) { g_encode_err = "gzb_encoder_set_group_device: null encoder"; return GZB_ERR_BAD_ARG; }
  if (enc->e.ran) { g_encode_err = "gzb_encoder_set_group_device: the encoder has already run"; return GZB_ERR_STATE; }
  enc->e.group.allgather_device = allgather_device;
  return GZB_OK;
}

int gzb_encoder_run(gzb_encoder* enc, uint8_t** jpeg_out, size_t* jpeg_size, gzb_encode_stats* stats,
                    char** trace_out) {
  if (!enc || !jpeg_out || !jpeg_size) { g_encode_err = "gzb_encoder_run: null argument"; return GZB_ERR_BAD_ARG; }
  *jpeg_out = nullptr; *jpeg_size = 0;
  if (trace_out) *trace_out = nullptr;
  Encoder& e = enc->e;
  if (e.ran) { g_encode_err = "gzb_encoder_run: an encoder runs once"; return GZB_ERR_STATE; }
  e.ran = true;
  const double t_start = now_ms();
  const int width = e.w, height = e.h;
  e.want_trace = trace_out != nullptr;
  auto fail = [&](int code) {
    g_encode_err = gzb_last_error(e.ctx);
    return code;
  };
  int ones[3][64];
  for (int c = 0; c < 3; ++c) for (int k = 0; k < 64; ++k) ones[c][k] = 1;
  // ---- the original + SelectQuantMatrix (processor.cc:310-372, 986-1003) ----
  // Trials are evaluated one per rank of the group (gzb_quant_search.h) and visited in the
  // reference's order; with a group of one this is the reference's sequential loop.
  // mode: 0 = the original, then the search (the 4:4:4 pass); 1 = the original only (before a forced
  // downsampling); 2 = the search of a downsampled image (no original, generator starts at score 0).
  auto select_quant = [&](const int mode, int best_q[3][64]) -> int {
    const double t_search = now_ms();
    // TryQuantMatrix (processor.cc:279-308) entirely on the device: quantise + IDCT, Huffman-code the
    // candidate (its size is what the search needs), Compare. The scan is kept (unstuffed) so that
    // the file can be assembled if this trial turns out to be the best so far.
    gzb::QuantSearch search(e.group, e.target, 1, mode);
    auto evaluate = [&](const std::vector<gzb::Trial>& ts, std::vector<gzb::TrialOutcome>* os) -> bool {
      for (size_t j = 0; j < ts.size(); ++j) {
        const gzb::Trial& t = ts[j];
        gzb::TrialOutcome& o = (*os)[j];
        const double t0 = now_ms();
        if ((t.original ? gzb_copy_from_jpeg(e.ctx, &ones[0][0]) : gzb_quantize_from_jpeg(e.ctx, &t.q[0][0])) != GZB_OK) return false;
        if (!e.compare_begin()) return false;   // the Compare runs while the file is coded on the second stream
        DeviceJpeg dj;
        if (!device_code_candidate(e.ctx, width, height, e.nb, t.q, t.original != 0, nullptr, nullptr, &dj, e.yuv420)) return false;
        ++e.code_gen;
        o.scan_bytes = dj.scan_bytes;
        o.ff_bytes = dj.ff_bytes;
        if (e.group.world == 1) {
          o.resident_gen = e.code_gen;   // visited right after this round: fetched only if it becomes the best
        } else {
          o.scan.resize(static_cast<size_t>(dj.scan_bytes));
          if (dj.scan_bytes && gzb_candidate_fetch_scan(e.ctx, reinterpret_cast<uint8_t*>(&o.scan[0]), dj.scan_bytes) != GZB_OK) return false;
        }
        o.jpg_size = dj.size();
        o.jpeg.swap(dj.header);
        e.st.num_jpeg_writes++;
        e.st.device_write_ms += now_ms() - t0;
        if (!e.compare_end()) return false;
        o.distance = e.distance;
        e.st.trial_device_ms += now_ms() - t0;
      }
      return true;
    };
    auto visit = [&](const gzb::Trial& t, const gzb::TrialOutcome& o) {
      if (t.original) {
        e.log("Original Out[%7zd]", static_cast<size_t>(o.jpg_size));
      } else {
        e.log("Iter %2d: %s GQ[%5.2f] Out[%7zd]", e.st.num_iterations + 1, e.frame_type(), quant_heuristic_score(t.q),
              static_cast<size_t>(o.jpg_size));
        ++e.st.num_iterations;
      }
      e.log(" BA[100.00%%] D[%6.4f]", o.distance);
      e.distance = o.distance;   // the comparator's state after this trial, wherever it ran
      e.maybe_output_trial(t, o);
    };
    if (!search.run(evaluate, visit)) {
      if (g_encode_err.empty()) g_encode_err = gzb_last_error(e.ctx);
      if (g_encode_err.empty()) g_encode_err = "gzb_encoder_run: the group exchange failed";
      return GZB_ERR_CUDA;
    }
    e.st.search_wall_ms += now_ms() - t_search;
    e.search_rounds += search.rounds();
    e.search_trials += search.evaluated_total();
    if (mode == 1) return GZB_OK;
    const QuantData& best = search.best();
    memcpy(best_q, best.q, 192 * sizeof(int));
    if (!best.dist_ok)
      for (int c = 0; c < 3; ++c) for (int k = 0; k < 64; ++k) best_q[c][k] = 1;
    return GZB_OK;
  };

  // ---- SelectFrequencyMasking(jpg, img, comp_mask, target_mul, stop_early) (processor.cc:559-919) ----
  // Returns 1 when this rank of a group has nothing left to do (the back end runs on rank 0).
  // jpg_ncomp: jpg.components.size() of the pass (1 only for a grey image after a forced "downsampling").
  // shard: the zeroing search is split over the ranks of the group (all ranks hold the same candidate). The
  // chroma pass of a YUV420 image starts from the candidate rank 0's luma back end has left, which the other
  // ranks do not have: rank 0 then searches alone (shard = false) and the others return at once.
  auto select_frequency_masking = [&](const int comp_mask, const double target_mul, const bool stop_early,
                                      const int jpg_ncomp, const bool shard = true) -> int {
    if (!shard && e.group.rank != 0) return 1;
    // units of the pass: 8x8 blocks, or the 16x16 macro-blocks of the sub-sampled chroma planes
    const int factor = (e.yuv420 && (comp_mask & 6)) ? 2 : 1;
    const int pass_bw = (width + 8 * factor - 1) / (8 * factor), pass_bh = (height + 8 * factor - 1) / (8 * factor);
    const int num_blocks = pass_bw * pass_bh;
    std::vector<int> cand_offsets(num_blocks + 1);
    std::vector<uint8_t> cand_coeffs;
    std::vector<float> cand_errors;   // only a group needs them on the host (for the exchange)
    const int world = shard ? e.group.world : 1, rank = shard ? e.group.rank : 0;
    bool be_begun = false;   // the device-side exchange leaves the lists in the back end's state
    {
      // In a group a failing rank still takes part in the exchanges and reports its status there, so
      // that all ranks leave together instead of waiting for it inside a collective.
      int zrc = GZB_OK;
      if (gzb_start_block_comparisons(e.ctx) != GZB_OK) zrc = GZB_ERR_CUDA;
      if (zrc != GZB_OK && world == 1) return fail(zrc);
      const double t0 = now_ms();
      // the blocks are independent: rank r of the group searches blocks [nb*r/world, nb*(r+1)/world)
      const int b0 = static_cast<int>(static_cast<int64_t>(num_blocks) * rank / world);
      const int b1 = static_cast<int>(static_cast<int64_t>(num_blocks) * (rank + 1) / world);
      const int nloc = b1 - b0;
      std::vector<int> loc_off(nloc + 1, 0);
      size_t ncand = 0;
      cand_coeffs.resize(static_cast<size_t>(nloc) * 48 + 16);
      // alone, or with a device-side exchange, the errors stay on the device (gzb_be_begin / gzb_be_begin_gathered
      // pick the packed lists up where they are)
      const bool host_exchange = world > 1 && e.group.allgather_device == nullptr;
      if (host_exchange) cand_errors.resize(cand_coeffs.size());
      float* errs = host_exchange ? cand_errors.data() : nullptr;
      if (zrc == GZB_OK && gzb_compute_block_zeroing_candidates_range(e.ctx, comp_mask, b0, b1, loc_off.data(), cand_coeffs.data(), errs,
                                                                      cand_coeffs.size(), &ncand) != GZB_OK) zrc = GZB_ERR_CUDA;
      if (zrc == GZB_OK && ncand > cand_coeffs.size()) {  // more than 48 candidates per block on average: fetch again (no recompute)
        cand_coeffs.resize(ncand);
        if (host_exchange) cand_errors.resize(ncand);
        errs = host_exchange ? cand_errors.data() : nullptr;
        if (gzb_compute_block_zeroing_candidates_range(e.ctx, comp_mask, b0, b1, loc_off.data(), cand_coeffs.data(), errs, ncand,
                                                       &ncand) != GZB_OK) zrc = GZB_ERR_CUDA;
      }
      if (zrc != GZB_OK) { ncand = 0; std::fill(loc_off.begin(), loc_off.end(), 0); g_encode_err = gzb_last_error(e.ctx); }
      if (zrc != GZB_OK && world == 1) return zrc;
      cand_coeffs.resize(ncand);
      if (host_exchange) cand_errors.resize(ncand);
      e.st.device_zeroing_ms += gzb_last_device_ms(e.ctx);
      if (world == 1) {
        cand_offsets = loc_off;
        // the back end's state is set up while the candidate lists still sit in the blur scratch
        if (gzb_be_begin(e.ctx, comp_mask, nullptr, nullptr, nullptr, ncand) != GZB_OK) return fail(GZB_ERR_CUDA);
      } else {
        // all-gather 1: every rank's per-block offsets (padded to the longest range), its candidate count
        // and its status -> global offsets
        const int maxloc = (num_blocks + world - 1) / world + 1;
        const int rec1 = maxloc + 2;
        std::vector<int> send_off(rec1, 0), all_off(static_cast<size_t>(world) * rec1);
        memcpy(send_off.data(), loc_off.data(), sizeof(int) * (nloc + 1));
        send_off[maxloc] = static_cast<int>(ncand);
        send_off[maxloc + 1] = zrc;
        if (e.group.allgather(e.group.user, send_off.data(), sizeof(int) * rec1, all_off.data()) != 0) {
          g_encode_err = "gzb_encoder_run: the group exchange failed";
          return GZB_ERR_CUDA;
        }
        for (int r = 0; r < world; ++r)
          if (all_off[static_cast<size_t>(r) * rec1 + maxloc + 1] != GZB_OK) {
            if (zrc == GZB_OK) g_encode_err = "gzb_encoder_run: the zeroing search failed on rank " + std::to_string(r);
            return GZB_ERR_CUDA;   // every rank sees the same statuses and leaves here
          }
        size_t maxn = 0, total = 0;
        std::vector<size_t> base(world + 1, 0);
        for (int r = 0; r < world; ++r) {
          const size_t n = static_cast<size_t>(all_off[static_cast<size_t>(r) * rec1 + maxloc]);
          maxn = std::max(maxn, n);
          base[r + 1] = base[r] + n;
          total += n;
        }
        for (int r = 0; r < world; ++r) {
          const int rb0 = static_cast<int>(static_cast<int64_t>(num_blocks) * r / world);
          const int rb1 = static_cast<int>(static_cast<int64_t>(num_blocks) * (r + 1) / world);
          const int* ro = &all_off[static_cast<size_t>(r) * rec1];
          for (int b = rb0; b < rb1; ++b) cand_offsets[b] = static_cast<int>(base[r]) + ro[b - rb0];
        }
        cand_offsets[num_blocks] = static_cast<int>(total);
        if (!host_exchange) {
          // all-gather 2 on the devices: the lists go from GPU to GPU and become the back end's on rank 0, whose
          // host only needs the coefficient indices
          std::vector<uint64_t> counts(world);
          for (int r = 0; r < world; ++r) counts[r] = base[r + 1] - base[r];
          if (rank == 0) cand_coeffs.resize(total);
          int grc = gzb_be_begin_gathered(e.ctx, comp_mask, world, rank, cand_offsets.data(), counts.data(), e.group.allgather_device,
                                          e.group.user, rank == 0 ? 1 : 0, rank == 0 ? cand_coeffs.data() : nullptr);
          // (a rank that fails here has already taken part in the collective or made it fail for everybody)
          if (grc != GZB_OK) return fail(GZB_ERR_CUDA);
          be_begun = true;
        } else {
        // all-gather 2: the packed candidates, [errors | coefficient indices], padded to the longest
        const size_t rec = maxn * 5;
        std::vector<uint8_t> send(rec + 1, 0), all(static_cast<size_t>(world) * (rec + 1));
        memcpy(send.data(), cand_errors.data(), ncand * sizeof(float));
        memcpy(send.data() + maxn * 4, cand_coeffs.data(), ncand);
        if (e.group.allgather(e.group.user, send.data(), rec + 1, all.data()) != 0) {
          g_encode_err = "gzb_encoder_run: the group exchange failed";
          return GZB_ERR_CUDA;
        }
        cand_coeffs.resize(total);
        cand_errors.resize(total);
        for (int r = 0; r < world; ++r) {
          const size_t n = base[r + 1] - base[r];
          const uint8_t* src = all.data() + static_cast<size_t>(r) * (rec + 1);
          memcpy(cand_errors.data() + base[r], src, n * sizeof(float));
          memcpy(cand_coeffs.data() + base[r], src + maxn * 4, n);
        }
        }
      }
      e.st.zeroing_wall_ms += now_ms() - t0;
      gzb_finish_block_comparisons(e.ctx);
    }

    // The back end is one sequential walk: rank 0 of a group finishes the image alone.
    if (rank != 0) return 1;
    if (world > 1 && !be_begun &&
        gzb_be_begin(e.ctx, comp_mask, cand_offsets.data(), cand_coeffs.data(), cand_errors.data(), cand_coeffs.size()) != GZB_OK)
      return fail(GZB_ERR_CUDA);
    std::vector<float>().swap(cand_errors);

    // ---- SelectFrequencyBackEnd (processor.cc:723-919) ----
    // The order, its sort, the prefix of every iteration and the per-block state live on the device
    // (gzb_backend.cuh); this thread keeps the histograms and walks the last, observable steps.
    {
      const double t_be = now_ms();
      const int ncomp = jpg_ncomp;
      // What the back end needs from the quantised image: header size, DC/AC histograms, DC size estimate
      // (processor.cc:742-755). back_ncomp: the components SaveToJpegData keeps (output_image.cc:588).
      Histogram ac_hist[3], dc_hist[3];
      int back_ncomp = 3, header_size = 0, dc_size = 0;
      {
        uint32_t dc[48], ac[768];
        if (gzb_candidate_symbol_histograms_n(e.ctx, nullptr, 3, dc, ac) != GZB_OK) return fail(GZB_ERR_CUDA);
        for (int c = 0; c < 3; ++c) {
          histogram_from_counts(dc + 16 * c, 16, &dc_hist[c]);
          histogram_from_counts(ac + 256 * c, 256, &ac_hist[c]);
        }
        back_ncomp = ncomp_from_histograms(dc_hist, ac_hist, e.nb);
        if (back_ncomp == 1) {
          // a grey file: its luma DC differences follow image block order (4:2:0: not MCU order), and the
          // dropped planes have no histograms
          if (e.yuv420) {
            if (gzb_candidate_symbol_histograms_n(e.ctx, nullptr, 1, dc, ac) != GZB_OK) return fail(GZB_ERR_CUDA);
            histogram_from_counts(dc, 16, &dc_hist[0]);
            histogram_from_counts(ac, 256, &ac_hist[0]);
          }
          for (int c = 1; c < 3; ++c) { dc_hist[c] = Histogram(); ac_hist[c] = Histogram(); }
        }
        Frame f;
        f.width = width; f.height = height; f.bw = e.bw; f.bh = e.bh; f.ncomp = back_ncomp;
        f.yuv420 = e.yuv420;
        gzb::jpeg::frame_set_quant(&f, e.quant);
        header_size = static_cast<int>(gzb::jpeg::header_size(f));
        // EstimateDCSize (processor.cc:548-555)
        Histogram tmp[3] = {dc_hist[0], dc_hist[1], dc_hist[2]};
        size_t num = f.ncomp;
        int ix[4];
        uint8_t dd[3 * Histogram::kSize];
        dc_size = static_cast<int>(gzb::jpeg::cluster_histograms(tmp, &num, ix, dd));
      }
      std::vector<uint8_t> ac_depths(3 * Histogram::kSize);
      // ComputeEntropyCodes (processor.cc:517-536)
      gzb::jpeg::HuffCache huff_caches[5];
      auto compute_entropy_codes = [&]() -> size_t {
        Histogram clustered[3] = {ac_hist[0], ac_hist[1], ac_hist[2]};
        size_t num = ncomp;
        int indexes[4];
        uint8_t cd[3 * Histogram::kSize];
        gzb::jpeg::cluster_histograms(clustered, &num, indexes, cd, huff_caches);
        for (int i = 0; i < ncomp; ++i)
          memcpy(&ac_depths[i * Histogram::kSize], &cd[indexes[i] * Histogram::kSize], Histogram::kSize);
        size_t hs = 0;
        for (size_t i = 0; i < num; ++i) hs += gzb::jpeg::header_cost_bits(clustered[i]) / 8;
        ++e.st.num_entropy_code_builds;
        return hs;
      };
      // EntropyCodedDataSize (processor.cc:538-546), with the raw bit sums cached per component.
      uint64_t raw_bits[3] = {0, 0, 0};
      auto recount_bits = [&]() {
        for (int c = 0; c < ncomp; ++c) {
          uint64_t bits = 0;
          const uint8_t* d = &ac_depths[c * Histogram::kSize];
          for (int i = 0; i + 1 < Histogram::kSize; ++i) bits += static_cast<uint64_t>(ac_hist[c].counts[i] / 2) * (d[i] + (i & 0xf));
          raw_bits[c] = bits;
        }
      };
      auto coded_size = [&]() -> size_t {
        size_t numbits = 0;
        for (int c = 0; c < ncomp; ++c) numbits += raw_bits[c] + ((raw_bits[c] * 3 + 512) >> 10);
        return (numbits + 7) / 8;
      };
      int ac_histogram_size = static_cast<int>(compute_entropy_codes());
      recount_bits();
      const int base_size = header_size + dc_size + ac_histogram_size + static_cast<int>(coded_size());
      int prev_size = base_size;

      // ---- the part of the order this thread sees: exactly sorted entries from position `wbase` on ----
      // and the state of the blocks they name (gzb_be_gather), which the walk flips locally; the flips
      // that stay applied go back to the device at the end of the iteration.
      struct WalkBlock {
        int block, last_index;
        bool in_prefix, touched;
        uint64_t zmask[3];
      };
      std::vector<WalkBlock> wblocks;
      // wstates[slot]: the block's coefficients as the device reported them, edited by the walk. A plain buffer:
      // the records are filled by gzb_be_gather, and value-initialising 800 bytes per block first costs as much
      // as copying them.
      struct StateBuf {
        gzb_be_block_state* p = nullptr;
        size_t n = 0, cap = 0;
        ~StateBuf() { free(p); }
        void clear() { n = 0; }
        bool resize(size_t want) {
          if (want > cap) {
            const size_t nc = std::max(want, std::max<size_t>(8192, cap * 2));
            void* q = realloc(p, nc * sizeof(gzb_be_block_state));
            if (!q) return false;
            p = static_cast<gzb_be_block_state*>(q);
            cap = nc;
          }
          n = want;
          return true;
        }
        gzb_be_block_state* data() { return p; }
        gzb_be_block_state& operator[](size_t i) { return p[i]; }
      } wstates;
      std::vector<int> wslot(num_blocks, -1);   // unit -> index into wblocks / wstates, -1: not fetched in this iteration
      std::vector<OrderEntry> went;
      size_t wbase = 0;
      bool order_done = false;         // every entry of the order has been fetched
      std::vector<OrderEntry> range_buf(GZB_BE_MAX_ENTRIES);
      exact_sort::HostLazy lazy;
      std::vector<int> req_blocks;
      // The device partitions down to ranges of at most small_max entries and hands over, in one round trip, as
      // many consecutive ones as cover want_more entries past the prefix: about what the walk of the previous
      // iteration asked for (the state of every block named in them is fetched too, so more is not free).
      static const int small_max_env = getenv("GZB_BE_SMALL_MAX") ? std::max(16, std::min(4096, atoi(getenv("GZB_BE_SMALL_MAX")))) : 0;
      static const int want_env = getenv("GZB_BE_WANT") ? std::max(0, atoi(getenv("GZB_BE_WANT"))) : -1;
      const int small_max = small_max_env ? small_max_env : 1024;
      // (one estimate per direction: "up" and "down" walks differ by an order of magnitude on some images; within
      // an iteration every further round trip asks for twice as much)
      const size_t want_cap = GZB_BE_MAX_ENTRIES - 1024;
      size_t want_dir[2] = {1024, 1024};
      size_t want_more = want_env >= 0 ? want_env : 1024;
      // the coefficient flips of the sequential walk
      struct Flips {
        std::vector<int32_t> block; std::vector<uint8_t> cidx; std::vector<int16_t> val;
        void clear() { block.clear(); cidx.clear(); val.clear(); }
      } flips;
      // windowed evaluation of the entropy-code rebuilds (see the walk below)
      struct SymDelta { int16_t sym; int8_t c; int8_t w; };
      struct UndoRec { int slot; uint8_t c, k; int16_t old_idx; uint64_t old_mask; bool newly_touched; uint32_t delta_begin; };
      struct CodeWindow {
        size_t first = 0;
        int nsteps = 0, changed_first = 0, break_step = -1, ac_histogram_size = 0;
        uint32_t delta_begin[11];
        Histogram hist[3];
        uint8_t depths[3 * Histogram::kSize];
        uint64_t raw[10][3];
        int est[10];
        gzb::jpeg::HuffCache caches[5];
      };
      // A batch of windows: walked by this thread (symbol deltas and undo records logged), then evaluated by the
      // pool. Two of them: the next batch is walked while the pool evaluates the current one.
      struct Batch {
        std::vector<SymDelta> dlog;
        std::vector<UndoRec> ulog;
        std::vector<CodeWindow> windows;
        Histogram base[3];   // the AC histograms before the batch's first step
        size_t i0 = 0;
        int nw = 0;
      };
      Batch batches[2];
      bool first_up_iter = true;
      const int directions[2] = {1, -1};
      const int n_ccoef = static_cast<int>(cand_coeffs.size());
      for (int direction : directions) {
        for (;;) {
          // down-adjusting only makes the output larger (processor.cc:766-774)
          if (stop_early && direction == -1 && prev_size > 1.01 * e.best_size) break;
          double tt = now_ms();
          // distmap is all zeros until the first iteration has compared (processor.cc:777-780)
          if (first_up_iter && gzb_clear_distmap(e.ctx) != GZB_OK) return fail(GZB_ERR_CUDA);
          // block weights for rblock = 1.. and global_order in block order (processor.cc:775-819), on the device
          uint64_t order_n = 0, below = 0;
          int blocks_to_change = 0;
          const float below_limit = 0.75f * gzb_block_error_limit(e.ctx);
          if (gzb_be_build_order(e.ctx, direction, target_mul, below_limit, &order_n, &blocks_to_change, &below, nullptr) != GZB_OK)
            return fail(GZB_ERR_CUDA);
          { const double t1 = now_ms(); e.st.be_order_ms += t1 - tt; tt = t1; }
          if (order_n == 0) break;
          const size_t order_size = static_cast<size_t>(order_n);

          double rel_size_delta = direction > 0 ? 0.01 : 0.0005;
          // DistanceOK(1.0) of the LAST Compare in the reference's order (in a group that may be a trial
          // another rank evaluated, so the context's own last distance must not be used)
          if (direction > 0 && static_cast<double>(e.distance) <= 1.0 * static_cast<double>(e.target)) rel_size_delta = 0.05;
          const double min_size_delta = base_size * rel_size_delta;
          const float coeffs_to_change_per_block = direction > 0 ? 2.0f : factor * factor * 0.2f;
          int min_coeffs_to_change = static_cast<int>(coeffs_to_change_per_block * blocks_to_change);
          if (first_up_iter) {
            // partition_point on the sorted order == number of entries below the limit (processor.cc:840-848)
            min_coeffs_to_change = std::max<int>(min_coeffs_to_change, static_cast<int>(below));
            first_up_iter = false;
          }
          float val_threshold = 0.0;
          int changed_coeffs = 0;
          int est_jpg_size = prev_size;
          Flips* job = &flips;
          job->clear();
          // ---- the silent prefix ----
          // While i + 9 < min_coeffs_to_change and i + 9 < order_size - 1 the reference's loop can
          // neither stop nor have its entropy-code rebuild observed (processor.cc:879-903), so the
          // first `prefix` entries of the sorted order are consumed as a SET on the device: the order is
          // partitioned std::sort's way only as far as needed to know which entries these are, every
          // block takes as many of its next candidates as it has entries in the set, and the symbol
          // histograms are recounted (sums are order-independent).
          size_t prefix = 0;
          if (min_coeffs_to_change > 9 && order_size > 10)
            prefix = std::min(static_cast<size_t>(min_coeffs_to_change - 9), order_size - 10);
          for (const WalkBlock& wb : wblocks) wslot[wb.block] = -1;
          wblocks.clear();
          wstates.clear();
          went.clear();
          lazy.clear();
          wbase = prefix;
          if (want_env < 0) want_more = want_dir[direction > 0 ? 1 : 0];
          order_done = false;
          int walk_changed_blocks = 0, prefix_changed_blocks = 0;
          bool prefix_applied = false;
          // Blocks named by entries [from, to) of `src` that the walk has not seen yet -> their state from the
          // device. The first call of an iteration also consumes the prefix (the head of the straddling range
          // must be in place on the device by then).
          auto gather_blocks = [&](const OrderEntry* src, size_t count) -> bool {
            const double tg = now_ms();
            size_t at = 0;
            do {
              req_blocks.clear();
              while (at < count && req_blocks.size() < GZB_BE_MAX_ENTRIES) {
                const int b = src[at++].first;
                if (wslot[b] < 0) { wslot[b] = static_cast<int>(wblocks.size() + req_blocks.size()); req_blocks.push_back(b); }
              }
              const size_t slot0 = wblocks.size();
              if (!wstates.resize(slot0 + req_blocks.size())) return false;
              gzb_be_block_state* req_states = wstates.data() + slot0;   // filled in place
              const int nreq = static_cast<int>(req_blocks.size());
              if (!prefix_applied) {
                uint32_t ac[768];
                if (gzb_be_apply_prefix(e.ctx, prefix, direction, back_ncomp, prefix > 0 ? ac : nullptr, &prefix_changed_blocks,
                                        req_blocks.data(), nreq, req_states) != GZB_OK) return false;
                if (prefix > 0) {
                  for (int c = 0; c < back_ncomp; ++c) histogram_from_counts(ac + 256 * c, 256, &ac_hist[c]);
                  recount_bits();  // raw bit sums for the current codes and the new histograms
                  changed_coeffs = static_cast<int>(prefix);
                  e.st.be_steps += prefix;
                  e.st.be_prefix_steps += prefix;
                }
                prefix_applied = true;
              } else if (nreq > 0) {
                if (gzb_be_gather(e.ctx, req_blocks.data(), nreq, direction, req_states) != GZB_OK) return false;
              }
              wblocks.resize(slot0 + req_blocks.size());
              for (size_t i = 0; i < req_blocks.size(); ++i) {
                WalkBlock& wb = wblocks[slot0 + i];
                wb.block = req_blocks[i];
                wb.last_index = req_states[i].last_index;
                wb.in_prefix = req_states[i].prefix_count > 0;
                wb.touched = false;
                for (int c = 0; c < 3; ++c) wb.zmask[c] = req_states[i].zmask[c];
              }
            } while (at < count);
            e.st.be_gather_ms += now_ms() - tg;
            return true;
          };
          // Makes at least one more entry of the order final (appended to `went`), or sets order_done. A range
          // comes from the device partitioned down to at most `small_max` entries (gzb_be_select); this thread
          // finishes it lazily, std::sort's way, only as far as the walk gets. The first range of an iteration
          // straddles the end of the prefix: its head is split off as a set and goes back to the device.
          auto fetch_more = [&]() -> bool {
            const double ts = now_ms();
            struct Timer { double t0; double& acc; ~Timer() { acc += now_ms() - t0; } } timer{ts, e.st.be_sort_ms};
            if (lazy.advance()) {
              went.insert(went.end(), lazy.buf.begin() + lazy.appended, lazy.buf.begin() + lazy.done);
              lazy.appended = lazy.done;
              e.st.be_lazy_ms += now_ms() - ts;
              return true;
            }
            const size_t have_end = wbase + went.size();
            if (!exact_sort::emulation_ok()) {
              // a standard library whose std::sort this file does not restate: sort the whole order with it
              std::vector<OrderEntry> all(order_size);
              if (gzb_be_fetch_order(e.ctx, 0, reinterpret_cast<gzb_order_entry*>(all.data()), order_size) != GZB_OK) return false;
              std::sort(all.begin(), all.end(), OrderLess());
              if (prefix > 0 && gzb_be_store_order(e.ctx, 0, reinterpret_cast<const gzb_order_entry*>(all.data()), prefix) != GZB_OK) return false;
              went.assign(all.begin() + prefix, all.end());
              order_done = true;
              return gather_blocks(went.data(), went.size());
            }
            int status = 0, nranges = 0;
            gzb_be_range rr[GZB_BE_MAX_RANGES];
            const double tsel = now_ms();
            // (the walk of the previous iteration is the best guess of how far this one will get)
            if (gzb_be_select_ranges(e.ctx, have_end, small_max, have_end + want_more, &status, &nranges, rr,
                                     reinterpret_cast<gzb_order_entry*>(range_buf.data())) != GZB_OK)
              return false;
            if (want_env < 0) want_more = std::min(want_cap, 2 * want_more);
            e.st.be_select_ms += now_ms() - tsel;
            if (status == 3 || nranges <= 0) { order_done = true; return true; }
            const size_t rf = static_cast<size_t>(rr[0].first), rl = static_cast<size_t>(rr[nranges - 1].last);
            if (status == 2) {   // depth budget exhausted on a long range: std::sort heap-sorts it
              lazy.buf.resize(rl - rf);
              if (gzb_be_fetch_order(e.ctx, rf, reinterpret_cast<gzb_order_entry*>(lazy.buf.data()), rl - rf) != GZB_OK) return false;
              lazy.reset(rl - rf, 0);
            } else {
              lazy.buf.assign(range_buf.begin(), range_buf.begin() + (rl - rf));
              exact_sort::HostLazy::R lr[GZB_BE_MAX_RANGES];
              for (int i = 0; i < nranges; ++i)
                lr[i] = {static_cast<size_t>(rr[i].first) - rf, static_cast<size_t>(rr[i].last) - rf, rr[i].depth};
              lazy.reset_ranges(lr, nranges);
              e.st.be_host_ranges += nranges - 1;
            }
            ++e.st.be_host_ranges;
            const size_t head = have_end > rf ? have_end - rf : 0;   // entries of the range that belong to the prefix
            if (head > 0) {
              const double tl = now_ms();
              lazy.split_set(head);
              e.st.be_lazy_ms += now_ms() - tl;
              if (gzb_be_store_order(e.ctx, rf, reinterpret_cast<const gzb_order_entry*>(lazy.buf.data()), head) != GZB_OK) return false;
            }
            lazy.appended = head;
            if (!gather_blocks(lazy.buf.data() + head, lazy.buf.size() - head)) return false;
            if (lazy.done <= head) lazy.advance();
            went.insert(went.end(), lazy.buf.begin() + lazy.appended, lazy.buf.begin() + lazy.done);
            lazy.appended = lazy.done;
            return true;
          };
          bool fetch_failed = false;
          auto have_entry = [&](size_t i) -> bool {
            while (i - wbase >= went.size()) {
              if (order_done) return false;
              if (!fetch_more()) { fetch_failed = true; return false; }
            }
            return true;
          };
          if (!have_entry(prefix)) return fail(GZB_ERR_CUDA);   // consumes the prefix; the walk has at least one step
          // One step of the walk (processor.cc:843-876): flips the next candidate of the order's i-th
          // block and updates the AC histograms by the symbols that change. With `dlog` the symbol
          // deltas and an undo record are logged instead of being priced with the current codes.
          auto flip = [&](size_t i, std::vector<SymDelta>* dlog, std::vector<UndoRec>* ulog) {
            const int b = went[i - wbase].first;
            const int slot = wslot[b];
            WalkBlock& wb = wblocks[slot];
            const int last_idx = wb.last_index;
            const int offset = std::max(0, std::min(cand_offsets[b], n_ccoef - 1));
            const uint8_t* candidates = cand_coeffs.data() + offset;
            const int cidx = candidates[last_idx + std::min(direction, 0)];
            const int c = cidx / 64, k = cidx % 64, z = gzb::jpeg::kZigZag[k];
            const int* qc = e.quant[c];
            int16_t* blk_idx = wstates[slot].idx[c];
            const int16_t newval = direction > 0 ? 0 : wstates[slot].requant[c][k];   // Quantize(jpg coefficient, q)
            const int16_t new_idx = static_cast<int16_t>(newval / qc[k]);
            const int16_t old_idx = blk_idx[k];
            // UpdateACHistogram(-1, old block); UpdateACHistogram(+1, new block) (processor.cc:491-515,
            // 871-873) restricted to the symbols that differ: those between the previous (p) and the
            // next (n) non-zero coefficient around zig-zag position z.
            uint64_t& m = wb.zmask[c];
            if (ulog) ulog->push_back(UndoRec{slot, static_cast<uint8_t>(c), static_cast<uint8_t>(k), old_idx, m, !wb.touched,
                                              static_cast<uint32_t>(dlog->size())});
            const uint64_t lower = m & ((1ULL << z) - 1);
            const int p = lower ? 63 - __builtin_clzll(lower) : 0;
            const uint64_t upper = z < 63 ? (m >> (z + 1)) : 0;
            const int n = upper ? z + 1 + __builtin_ctzll(upper) : 64;
            const int v_n = n < 64 ? blk_idx[gzb::jpeg::kNaturalOrder[n]] : 0;
            Histogram& hh = ac_hist[c];
            const uint8_t* d = &ac_depths[c * Histogram::kSize];
            auto add_sym = [&](int sym, int weight) {
              hh.add(sym, weight);
              if (dlog) dlog->push_back(SymDelta{static_cast<int16_t>(sym), static_cast<int8_t>(c), static_cast<int8_t>(weight)});
              else raw_bits[c] += static_cast<int64_t>(weight) * (d[sym] + (sym & 0xf));
            };
            auto add_run = [&](int run, int v, int weight) {
              while (run > 15) { add_sym(0xf0, weight); run -= 16; }
              add_sym((run << 4) + (32 - __builtin_clz(static_cast<unsigned>(std::abs(v)))), weight);
            };
            auto emit = [&](int v_z, int weight) {
              if (v_z != 0) {
                add_run(z - p - 1, v_z, weight);
                if (n < 64) add_run(n - z - 1, v_n, weight);
                else if (z != 63) add_sym(0, weight);
              } else {
                if (n < 64) add_run(n - p - 1, v_n, weight);
                else add_sym(0, weight);
              }
            };
            emit(old_idx, -1);
            emit(new_idx, 1);
            blk_idx[k] = new_idx;
            if (new_idx != 0) m |= 1ULL << z; else m &= ~(1ULL << z);
            job->block.push_back(b); job->cidx.push_back(static_cast<uint8_t>(cidx)); job->val.push_back(newval);
            wb.last_index += direction;
            if (!wb.touched) { wb.touched = true; if (!wb.in_prefix) ++walk_changed_blocks; }
            ++changed_coeffs;
            ++e.st.be_steps;
          };
          auto rebuild_needed = [&](size_t i) {
            // Evaluate only where the result can be observed: by the break test within the next 10
            // steps, or as prev_size when the order is about to run out.
            return static_cast<long long>(i) + 9 >= min_coeffs_to_change || i + 9 >= order_size - 1;
          };
          const bool windowed = e.pool->size() >= 4;
          size_t i = prefix;
          bool stopped = false;
          size_t last_step = prefix;   // index of the last step that stays applied
          // ---- the steps up to the first observable entropy-code rebuild: the reference's loop as is ----
          for (; i < order_size; ++i) {
            if (windowed && i % 10 == 0 && rebuild_needed(i)) break;
            if (!have_entry(i)) return fail(GZB_ERR_CUDA);
            flip(i, nullptr, nullptr);
            last_step = i;
            if (i % 10 == 0 && rebuild_needed(i)) {
              const double tb = now_ms();
              ac_histogram_size = static_cast<int>(compute_entropy_codes());
              recount_bits();
              e.st.be_codes_ms += now_ms() - tb;
            }
            if (changed_coeffs > min_coeffs_to_change || i + 1 == order_size) {
              est_jpg_size = header_size + dc_size + ac_histogram_size + static_cast<int>(coded_size());
              if (changed_coeffs > min_coeffs_to_change && std::abs(est_jpg_size - prev_size) > min_size_delta) { stopped = true; break; }
            }
          }
          // ---- from there on: windows of ten steps, each opening with a ComputeEntropyCodes rebuild ----
          // The flips do not depend on the codes (only the stopping test does), so this thread walks a
          // batch of windows ahead, logging symbol deltas and undo records, the pool rebuilds the codes
          // of every window and replays its size estimates in parallel, and the flips past the first
          // window that stops are undone. Decisions are those of the one-step-at-a-time loop.
          if (windowed && !stopped && i < order_size) {
            const int NB = 2 * e.pool->size();
            // walks up to NB windows from step i on; false: the order could not be fetched
            auto walk_batch = [&](Batch& B) -> bool {
              B.dlog.clear();
              B.ulog.clear();
              B.i0 = i;
              B.nw = 0;
              for (int c = 0; c < 3; ++c) B.base[c] = ac_hist[c];
              if (static_cast<int>(B.windows.size()) < NB) B.windows.resize(NB);
              while (B.nw < NB && i < order_size) {
                CodeWindow& W = B.windows[B.nw++];
                W.first = i;
                W.nsteps = 0;
                W.break_step = -1;
                for (int st = 0; st < 10 && i < order_size; ++st, ++i) {
                  W.delta_begin[st] = static_cast<uint32_t>(B.dlog.size());
                  if (!have_entry(i)) return false;
                  flip(i, &B.dlog, &B.ulog);
                  if (st == 0) W.changed_first = changed_coeffs;   // (its histograms: base + the deltas so far, rebuilt by the pool)
                  ++W.nsteps;
                }
                W.delta_begin[W.nsteps] = static_cast<uint32_t>(B.dlog.size());
              }
              return true;
            };
            // one window on a pool thread: its entropy codes, then the size estimate step by step
            auto eval_window = [&](Batch& B, int w) {
              CodeWindow& W = B.windows[w];
              const std::vector<SymDelta>& dlog = B.dlog;
              // the histograms after the window's first step: the batch's base plus every delta up to there
              for (int c = 0; c < 3; ++c) W.hist[c] = B.base[c];
              for (uint32_t j = 0; j < W.delta_begin[1]; ++j) W.hist[dlog[j].c].add(dlog[j].sym, dlog[j].w);
              Histogram clustered[3] = {W.hist[0], W.hist[1], W.hist[2]};
              size_t num = ncomp;
              int indexes[4];
              uint8_t cd[3 * Histogram::kSize];
              gzb::jpeg::cluster_histograms(clustered, &num, indexes, cd, W.caches);
              for (int c = 0; c < ncomp; ++c)
                memcpy(&W.depths[c * Histogram::kSize], &cd[indexes[c] * Histogram::kSize], Histogram::kSize);
              size_t hs = 0;
              for (size_t k = 0; k < num; ++k) hs += gzb::jpeg::header_cost_bits(clustered[k]) / 8;
              W.ac_histogram_size = static_cast<int>(hs);
              uint64_t raw[3] = {0, 0, 0};
              for (int c = 0; c < ncomp; ++c) {
                const uint8_t* d = &W.depths[c * Histogram::kSize];
                uint64_t bits = 0;
                for (int k = 0; k + 1 < Histogram::kSize; ++k) bits += static_cast<uint64_t>(W.hist[c].counts[k] / 2) * (d[k] + (k & 0xf));
                raw[c] = bits;
              }
              for (int st = 0; st < W.nsteps; ++st) {
                if (st > 0)
                  for (uint32_t j = W.delta_begin[st]; j < W.delta_begin[st + 1]; ++j) {
                    const SymDelta& dl = dlog[j];
                    raw[dl.c] += static_cast<int64_t>(dl.w) * (W.depths[dl.c * Histogram::kSize + dl.sym] + (dl.sym & 0xf));
                  }
                for (int c = 0; c < 3; ++c) W.raw[st][c] = raw[c];
                const int changed = W.changed_first + st;
                W.est[st] = -1;
                if (changed > min_coeffs_to_change || W.first + st + 1 == order_size) {
                  size_t numbits = 0;
                  for (int c = 0; c < ncomp; ++c) numbits += raw[c] + ((raw[c] * 3 + 512) >> 10);
                  const int est = header_size + dc_size + W.ac_histogram_size + static_cast<int>((numbits + 7) / 8);
                  W.est[st] = est;
                  if (changed > min_coeffs_to_change && std::abs(est - prev_size) > min_size_delta) { W.break_step = st; break; }
                }
              }
            };
            // takes back the steps of a batch from its keep-th on (the latest batch first)
            auto undo_from = [&](Batch& B, size_t keep) {
              if (keep >= B.ulog.size()) return;
              for (size_t j = B.dlog.size(); j-- > B.ulog[keep].delta_begin;) ac_hist[B.dlog[j].c].add(B.dlog[j].sym, -B.dlog[j].w);
              for (size_t j = B.ulog.size(); j-- > keep;) {
                const UndoRec& u = B.ulog[j];
                WalkBlock& wb = wblocks[u.slot];
                wstates[u.slot].idx[u.c][u.k] = u.old_idx;
                wb.zmask[u.c] = u.old_mask;
                wb.last_index -= direction;
                if (u.newly_touched) { wb.touched = false; if (!wb.in_prefix) --walk_changed_blocks; }
                job->block.pop_back(); job->cidx.pop_back(); job->val.pop_back();
                --changed_coeffs;
                --e.st.be_steps;
              }
            };
            const std::function<void(int)> evals[2] = {[&](int w) { eval_window(batches[0], w); }, [&](int w) { eval_window(batches[1], w); }};
            const double tb = now_ms();
            int cur = 0;
            if (!walk_batch(batches[0])) return fail(GZB_ERR_CUDA);
            for (;;) {
              Batch& B = batches[cur];
              Batch& N = batches[1 - cur];
              // The pool evaluates this batch while this thread walks the next one ahead of the decision: if the
              // loop stops inside this batch, the next one is taken back whole.
              e.pool->begin(B.nw, evals[cur]);
              bool have_next = false, ok = true;
              if (i < order_size) { ok = walk_batch(N); have_next = ok; }
              const double tpool = now_ms();
              e.pool->finish();
              e.st.be_pool_ms += now_ms() - tpool;
              if (!ok) return fail(GZB_ERR_CUDA);
              int final_w = B.nw - 1, final_st = B.windows[B.nw - 1].nsteps - 1;
              for (int w = 0; w < B.nw; ++w)
                if (B.windows[w].break_step >= 0) { final_w = w; final_st = B.windows[w].break_step; stopped = true; break; }
              const CodeWindow& F = B.windows[final_w];
              e.st.num_entropy_code_builds += final_w + 1;
              last_step = F.first + final_st;
              if (stopped) {  // undo the steps walked past the stop
                if (have_next) undo_from(N, 0);
                undo_from(B, last_step + 1 - B.i0);
              }
              memcpy(ac_depths.data(), F.depths, 3 * Histogram::kSize);
              ac_histogram_size = F.ac_histogram_size;
              for (int c = 0; c < 3; ++c) raw_bits[c] = F.raw[final_st][c];
              if (F.est[final_st] >= 0) est_jpg_size = F.est[final_st];
              if (stopped || !have_next) break;
              cur = 1 - cur;
            }
            e.st.be_codes_ms += now_ms() - tb;
          }
          if (fetch_failed) return fail(GZB_ERR_CUDA);
          if (changed_coeffs > 0) val_threshold = went[last_step - wbase].second;
          if (want_env < 0) want_dir[direction > 0 ? 1 : 0] = std::min(want_cap, std::max<size_t>(256, went.size() + went.size() / 8 + 64));
          const size_t changed_blocks = static_cast<size_t>(prefix_changed_blocks + walk_changed_blocks);
          { const double t1 = now_ms(); e.st.be_walk_ms += t1 - tt; tt = t1; }
          ++e.st.num_iterations;
          if (direction > 0) ++e.st.num_iterations_up; else ++e.st.num_iterations_down;
          // the walked flips, max_block_error += block_weight * val_threshold * direction (processor.cc:893-895),
          // and the candidate's samples for the Compare
          if (gzb_be_finish_iteration(e.ctx, job->block.data(), job->cidx.data(), job->val.data(), job->block.size(), direction,
                                      val_threshold) != GZB_OK) return fail(GZB_ERR_CUDA);
          { const double t1 = now_ms(); e.st.be_update_ms += t1 - tt; tt = t1; }
          if (!e.compare_begin()) return fail(GZB_ERR_CUDA);
          // the iteration's file: coded on the device (second stream, concurrently with the Compare)
          // with the histograms the walk maintains; only its size is needed unless it becomes the best
          // (processor.cc:897-915, MaybeOutput 151-160)
          const double tw = now_ms();
          DeviceJpeg dj;
          if (!device_code_candidate(e.ctx, width, height, e.nb, e.quant, false, dc_hist, ac_hist, &dj, e.yuv420)) return fail(GZB_ERR_CUDA);
          ++e.code_gen;
          e.st.num_jpeg_writes++;
          e.st.device_write_ms += now_ms() - tw;
          if (!e.compare_end()) return fail(GZB_ERR_CUDA);
          const size_t jpg_size = dj.size();
          e.log("Iter %2d: %s(%d) %s Coeffs[%d/%zd] Blocks[%zd/%d/%d] ValThres[%.4f] Out[%7zd] EstErr[%.2f%%]",
                e.st.num_iterations, e.frame_type(), comp_mask, direction > 0 ? "up" : "down", changed_coeffs, order_size,
                changed_blocks, blocks_to_change, num_blocks, val_threshold, jpg_size,
                100.0 - (100.0 * est_jpg_size) / jpg_size);
          e.log(" BA[100.00%%] D[%6.4f]", e.distance);
          const double score = score_jpeg(e.distance, static_cast<int>(jpg_size), e.target);
          e.log(" Score[%.4f]", score);
          if (score < e.best_score || e.best_score < 0) {
            if (!device_fetch_jpeg(e.ctx, dj, &e.best_jpeg)) return fail(GZB_ERR_CUDA);
            e.best_size = jpg_size;
            e.best_score = score;
            e.best_remote = false;
            e.log(" (*)");
          }
          e.log("\n");
          prev_size = est_jpg_size;
        }
      }
      e.st.backend_wall_ms += now_ms() - t_be;
    }
  return GZB_OK;
  };

  // ---- the passes of ProcessJpegData (processor.cc:986-1016) ----
  int gray_i = 0;   // IsGrayscale(jpg_in) (processor.cc:920-928)
  if (gzb_input_is_gray(e.ctx, &gray_i) != GZB_OK) return fail(GZB_ERR_CUDA);
  const bool gray = gray_i != 0;
  const int try_420 = (e.force_420 || (e.try_420 && !gray)) ? 1 : 0;
  // (a YUV420 pass replaces the input coefficients on the device: a 4:4:4 trial could not be rendered again after it)
  static const bool no_defer = getenv("GZB_NO_DEFER_BYTES") != nullptr;
  e.defer_trial_bytes = !try_420 && !no_defer;
  const int force_420 = e.force_420 ? 1 : 0;
  // force_420 on a grey image: OutputImage::Downsample returns early (output_image.cc:536-539), the
  // image stays 4:4:4 and SaveToJpegData leaves ONE component in jpg, so the pass is the luma-only one
  // with target_mul 1 (processor.cc:1011-1014; the chroma pass returns at processor.cc:735).
  const bool gray_force = force_420 && gray;
  // The best file so far is a SelectQuantMatrix trial whose bytes are not here (evaluated on another rank, or
  // not fetched): render and code it again. ("Original": the q=1 input with three index-0 tables,
  // processor.cc:967-985.)
  auto rebuild_best_trial = [&]() -> bool {
    const gzb::Trial& t = e.best_trial;
    if ((t.original ? gzb_copy_from_jpeg(e.ctx, &ones[0][0]) : gzb_quantize_from_jpeg(e.ctx, &t.q[0][0])) != GZB_OK) return false;
    DeviceJpeg dj;
    if (!device_code_candidate(e.ctx, width, height, e.nb, t.q, t.original != 0, nullptr, nullptr, &dj, e.yuv420) ||
        !device_fetch_jpeg(e.ctx, dj, &e.best_jpeg)) return false;
    e.st.num_jpeg_writes++;
    e.best_remote = false;
    return true;
  };
  bool idle_rank = false;   // a rank other than 0 of a group: it takes part in the searches, rank 0 returns the file
  if (force_420) {   // the original is compared and output before any pass (processor.cc:967-985)
    const int rc = select_quant(1, nullptr);
    if (rc != GZB_OK) return rc;
  }
  for (int downsample = force_420; downsample <= try_420; ++downsample) {
    int best_q[3][64];
    if (downsample && !gray_force) {
      // The downsampling replaces the input coefficients: a 4:4:4 trial that is the best so far and whose bytes
      // are elsewhere has to be rebuilt now.
      if (e.best_remote && !idle_rank && !rebuild_best_trial()) return fail(GZB_ERR_CUDA);
      // DownsampleImage + SaveToJpegData on the q=1 input, on the device (every rank of a group on its own copy)
      const double t0 = now_ms();
      if (gzb_downsample_420(e.ctx) != GZB_OK) return fail(GZB_ERR_CUDA);
      e.set_geometry(true);
      e.st.downsample_ms += now_ms() - t0;
    }
    int rc = select_quant(downsample ? 2 : 0, best_q);
    if (rc != GZB_OK) return rc;
    if (!e.set_global_quant(best_q)) return fail(GZB_ERR_CUDA);
    if (!downsample) {
      rc = select_frequency_masking(7, 1.0, false, 3);
    } else if (gray_force) {
      rc = select_frequency_masking(1, 1.0, false, 1);
    } else {
      rc = select_frequency_masking(1, static_cast<double>(0.97f), false, 3);   // ymul (processor.cc:1011)
      if (rc >= 0) rc = select_frequency_masking(6, 1.0, true, 3, /*shard=*/false);
    }
    if (rc < 0) return rc;
    idle_rank = rc == 1;
  }
  // The back end is one sequential walk: rank 0 of a group finishes the image alone.
  if (idle_rank) {
    e.st.launches = gzb_launch_count(e.ctx);
    gzb_get_transfer_bytes(e.ctx, &e.st.h2d_bytes, &e.st.d2h_bytes);
    e.st.run_ms = now_ms() - t_start;
    e.st.total_wall_ms = e.st.prepare_ms + e.st.run_ms;
    e.st.search_rounds = e.search_rounds;
    e.st.search_trials = e.search_trials;
    *jpeg_out = static_cast<uint8_t*>(malloc(1));
    if (stats) *stats = e.st;
    if (trace_out) {
      *trace_out = static_cast<char*>(malloc(e.trace.size() + 1));
      memcpy(*trace_out, e.trace.c_str(), e.trace.size() + 1);
    }
    return GZB_OK;
  }

  if (e.best_remote && !rebuild_best_trial()) return fail(GZB_ERR_CUDA);
  gzb_be_stats(e.ctx, &e.st.be_selects, &e.st.be_levels);
  e.st.num_fine_bdm_compares = gzb_fine_bdm_compare_count(e.ctx);
  e.st.search_rounds = e.search_rounds;
  e.st.search_trials = e.search_trials;
  e.st.launches = gzb_launch_count(e.ctx);
  e.st.write_hist_ms = e.wt.hist_ms; e.st.write_code_ms = e.wt.code_ms;
  e.st.write_encode_ms = e.wt.encode_ms; e.st.write_stitch_ms = e.wt.stitch_ms;
  gzb_get_transfer_bytes(e.ctx, &e.st.h2d_bytes, &e.st.d2h_bytes);
  *jpeg_size = e.best_jpeg.size();
  *jpeg_out = static_cast<uint8_t*>(malloc(std::max<size_t>(1, e.best_jpeg.size())));
  memcpy(*jpeg_out, e.best_jpeg.data(), e.best_jpeg.size());
  e.st.run_ms = now_ms() - t_start;
  e.st.total_wall_ms = e.st.prepare_ms + e.st.run_ms;
  e.st.final_distance = e.distance;
  e.st.final_score = e.best_score;
  if (stats) *stats = e.st;
  if (trace_out) {
    *trace_out = static_cast<char*>(malloc(e.trace.size() + 1));
    memcpy(*trace_out, e.trace.c_str(), e.trace.size() + 1);
  }
  return GZB_OK;
}

int gzb_encode_rgb(int device, const uint8_t* rgb, int width, int height, float butteraugli_target,
                   int host_threads, uint8_t** jpeg_out, size_t* jpeg_size, gzb_encode_stats* stats,
                   char** trace_out) {
  return gzb_encode_rgb_params(device, rgb, width, height, butteraugli_target, 0, 0, host_threads, jpeg_out, jpeg_size,
                               stats, trace_out);
}

int gzb_encode_rgb_params(int device, const uint8_t* rgb, int width, int height, float butteraugli_target,
                          int try_420, int force_420, int host_threads, uint8_t** jpeg_out, size_t* jpeg_size,
                          gzb_encode_stats* stats, char** trace_out) {
  if (jpeg_out) *jpeg_out = nullptr;
  if (jpeg_size) *jpeg_size = 0;
  if (trace_out) *trace_out = nullptr;
  gzb_encoder* enc = nullptr;
  int rc = gzb_encoder_create(device, rgb, width, height, butteraugli_target, host_threads, &enc);
  if (rc != GZB_OK) return rc;
  gzb_encoder_set_params(enc, try_420, force_420);
  rc = gzb_encoder_run(enc, jpeg_out, jpeg_size, stats, trace_out);
  gzb_encoder_destroy(enc);
  return rc;
}

// A batch of images on ONE GPU with `inflight` encodes running concurrently (each on its own host
// threads, device context and streams): one encode's sequential host phases (sort, entropy-code
// rebuilds) overlap the other encodes' kernels. Results are those of n independent gzb_encode_rgb_params
// calls. status[i] receives each image's return code; the function returns the first failure (or GZB_OK).
int gzb_encode_rgb_batch(int device, int n, const uint8_t* const* rgb, const int* width, const int* height,
                         float butteraugli_target, int try_420, int force_420, int inflight, int host_threads_per_encode,
                         uint8_t** jpeg_out, size_t* jpeg_size, gzb_encode_stats* stats, int* status) {
  if (n < 0 || (n > 0 && (!rgb || !width || !height || !jpeg_out || !jpeg_size))) return GZB_ERR_BAD_ARG;
  if (inflight < 1) inflight = 1;
  if (host_threads_per_encode <= 0) {
    const unsigned hc = std::max(1u, std::thread::hardware_concurrency());
    host_threads_per_encode = static_cast<int>(std::max(1u, std::min(16u, hc / static_cast<unsigned>(std::min(inflight, std::max(n, 1))))));
  }
  std::atomic<int> next(0), first_error(GZB_OK);
  std::mutex err_mu;
  std::string err_msg;
  auto work = [&] {
    for (;;) {
      const int i = next.fetch_add(1);
      if (i >= n) break;
      const int rc = gzb_encode_rgb_params(device, rgb[i], width[i], height[i], butteraugli_target, try_420, force_420,
                                           host_threads_per_encode, &jpeg_out[i], &jpeg_size[i], stats ? &stats[i] : nullptr, nullptr);
      if (status) status[i] = rc;
      if (rc != GZB_OK) {
        std::lock_guard<std::mutex> l(err_mu);
        if (first_error.load() == GZB_OK) { first_error = rc; err_msg = gzb_encode_last_error(); }
      }
    }
  };
  std::vector<std::thread> th;
  for (int t = 1; t < std::min(inflight, n); ++t) th.emplace_back(work);
  work();
  for (std::thread& t : th) t.join();
  if (first_error.load() != GZB_OK) g_encode_err = err_msg;
  return first_error.load();
}

}  // extern "C"
