// gzb_jpeg.h -- host-side JPEG back end of the search: Huffman code construction, histogram
// clustering, size estimation and the baseline bit-stream writer.
//
// Written from scratch; the OUTPUT BYTES must equal the reference's, because the search scores
// every candidate by its true encoded size and the golden checksum covers the final file:
//   guetzli/jpeg_data_writer.cc (WriteJpeg 540-553, BuildAndEncodeHuffmanCodes 361-456,
//   ClusterHistograms 295-342, EncodeScan 506-536), guetzli/entropy_encode.cc (CreateHuffmanTree
//   68-143), guetzli/jpeg_bit_writer.h.
#pragma once
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <atomic>
#include <vector>

namespace gzb {

// Minimal persistent worker pool: run(n, fn) calls fn(i) for i in [0, n) on the workers plus the
// calling thread and returns when all are done. One pool per encoder (not shared across threads).
// The search driver issues thousands of sub-millisecond parallel steps per encode, so workers spin
// for a short while on the job epoch before they go to sleep on the condition variable: a dispatch
// then costs about a microsecond instead of a futex wake-up per worker.
class WorkerPool {
 public:
  explicit WorkerPool(int nthreads);
  ~WorkerPool();
  int size() const { return nthreads_; }
  void run(int n, const std::function<void(int)>& fn);
  // The same in two halves: begin() hands the items to the workers and returns; finish() makes the caller work
  // on what is left and returns when all items are done. fn must stay alive in between, and run() must not be
  // called in between (it would then do its items on the calling thread).
  void begin(int n, const std::function<void(int)>& fn);
  void finish();

 private:
  struct Job {
    const std::function<void(int)>* fn;
    int n;
    std::atomic<int> next{0}, pending{0};
  };
  void worker();
  int nthreads_;
  std::vector<std::thread> threads_;
  std::mutex mu_;
  std::condition_variable cv_start_;
  std::atomic<Job*> job_{nullptr};
  std::atomic<unsigned long long> epoch_{0};
  std::atomic<int> active_{0}, sleepers_{0};
  std::atomic<bool> stop_{false};
  Job async_job_;
  bool async_open_ = false;
};

namespace jpeg {

extern const int kNaturalOrder[64];  // zig-zag position -> natural index
extern const int kZigZag[64];        // natural index -> zig-zag position

// Symbol histogram with every occurrence counted twice plus one sentinel of weight 1 that takes
// the all-ones code (JpegHistogram, guetzli/jpeg_data_writer.h:54-90).
struct Histogram {
  static const int kSize = 257;
  uint32_t counts[kSize];
  Histogram() { clear(); }
  void clear() { memset(counts, 0, sizeof(counts)); counts[kSize - 1] = 1; }
  void add(int symbol, int weight = 1) { counts[symbol] += 2 * weight; }
  void merge(const Histogram& o) {
    for (int i = 0; i + 1 < kSize; ++i) counts[i] += o.counts[i];
    counts[kSize - 1] = 1;
  }
  int num_symbols() const {
    int n = 0;
    for (int i = 0; i + 1 < kSize; ++i) n += counts[i] > 0;
    return n;
  }
};

// Leaf order of a previous huffman_depths call on a nearby histogram (speeds up the next one;
// never changes the result).
struct HuffCache { int n = 0; int16_t order[257]; };
// Length-limited Huffman depths (CreateHuffmanTree).
void huffman_depths(const uint32_t* counts, int length, int limit, uint8_t* depth, HuffCache* cache = nullptr);
size_t header_cost_bits(const Histogram& h);                        // HistogramHeaderCost
size_t entropy_cost_bits(const Histogram& h, const uint8_t* depth);  // HistogramEntropyCost
// ClusterHistograms: merges trailing histograms while that is cheaper; returns bytes.
// caches: optional HuffCache[5] (three inputs + the two possible merges), kept between calls.
size_t cluster_histograms(Histogram* histo, size_t* num, int* indexes, uint8_t* depth, HuffCache* caches = nullptr);

uint64_t zigzag_nonzero_mask(const int16_t* q);
// AC symbols of one block of QUANTISED indices in natural order (UpdateACHistogramForDCTBlock).
void ac_histogram_add_block(const int16_t* q, int weight, Histogram* h);

// A frame ready to be serialised: quantised indices, block-major, per component.
struct Frame {
  int width = 0, height = 0, ncomp = 3;
  int bw = 0, bh = 0;
  // YUV 4:2:0 (2x2 luma blocks per MCU): only the header is written on the host for such frames --
  // their scan is coded on the device (gzb_huffman.cuh). With ncomp == 1 the file is a plain grey one.
  bool yuv420 = false;
  const int16_t* coeffs[3] = {nullptr, nullptr, nullptr};  // quantised indices [bw*bh*64]
  // quantisation tables as the file will carry them
  int num_tables = 0;
  int table_values[3][64];
  int table_index[3];     // DQT index byte of table t
  int comp_table[3];      // table used by component c
};

// Deduplicates q[3][64] into tables like SaveQuantTables (guetzli/jpeg_data.cc:67-102).
void frame_set_quant(Frame* f, const int q[3][64]);
// The reference's RGB front end writes three tables that all carry index 0
// (guetzli/jpeg_data_encoder.cc:66-83 with JPEGQuantTable::index defaulting to 0).
void frame_set_quant_input(Frame* f, const int q[3][64]);

size_t header_size(const Frame& f);             // JpegHeaderSize with stripped metadata
size_t estimate_dc_size(const Frame& f);        // EstimateDCSize (processor.cc:548-555)
void build_ac_histograms(const Frame& f, Histogram* h /*[ncomp]*/);

// Per-component DC / AC histograms of a frame (unclustered), band-parallel.
void build_histograms(const Frame& f, Histogram* dc /*[ncomp]*/, Histogram* ac /*[ncomp]*/, WorkerPool* pool);

// Canonical Huffman code of one table: depth 255 marks an unused symbol.
struct CodeTable { uint8_t depth[256]; uint16_t code[256]; };
// Everything of WriteJpeg up to and including the SOS header (SOI, APP0, DQT, SOF, DHT, SOS):
// clusters the per-component DC and AC histograms, builds the codes and returns the per-component
// code tables the scan is coded with.
void write_jpeg_header(const Frame& f, const Histogram* dc_hist /*[ncomp]*/, const Histogram* ac_hist /*[ncomp]*/,
                       std::string* out, CodeTable* dc_tab /*[3]*/, CodeTable* ac_tab /*[3]*/);

struct WriteTimers { double hist_ms = 0, code_ms = 0, encode_ms = 0, stitch_ms = 0; };

// WriteJpeg(jpg, strip_metadata=true, out). With a pool, block-row bands are Huffman-coded in
// parallel and stitched at bit granularity (also in parallel); the bytes are identical to the
// sequential writer. dc_hist / ac_hist, when given, are the frame's per-component histograms
// (the back end maintains the AC ones incrementally) and save the two coefficient scans.
void write_jpeg(const Frame& f, std::string* out, WorkerPool* pool = nullptr,
                const Histogram* dc_hist = nullptr, const Histogram* ac_hist = nullptr,
                WriteTimers* tm = nullptr);

}  // namespace jpeg
}  // namespace gzb
