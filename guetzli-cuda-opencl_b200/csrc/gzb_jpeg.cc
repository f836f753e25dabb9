// gzb_jpeg.cc -- see gzb_jpeg.h. Output bytes are identical to the reference writer's.
#include "gzb_jpeg.h"

#include <algorithm>
#include <cassert>
#include <chrono>
#include <cstdlib>
#include <thread>

namespace gzb {
namespace jpeg {

const int kNaturalOrder[64] = {0,  1,  8,  16, 9,  2,  3,  10, 17, 24, 32, 25, 18, 11, 4,  5,
                               12, 19, 26, 33, 40, 48, 41, 34, 27, 20, 13, 6,  7,  14, 21, 28,
                               35, 42, 49, 56, 57, 50, 43, 36, 29, 22, 15, 23, 30, 37, 44, 51,
                               58, 59, 52, 45, 38, 31, 39, 46, 53, 60, 61, 54, 47, 55, 62, 63};
const int kZigZag[64] = {0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42,
                         3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
                         10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60,
                         21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};

static inline int log2_floor_nz(uint32_t n) { return 31 ^ __builtin_clz(n); }
static inline int bit_length(uint32_t n) { return n == 0 ? 0 : log2_floor_nz(n) + 1; }

// ---------------------------------------------------------------------------------------------
// Huffman depths: package-free two-queue construction with a rising count floor until the tree
// fits `limit` bits (entropy_encode.cc:68-143). Ties: equal counts order by descending symbol.
// ---------------------------------------------------------------------------------------------
void huffman_depths(const uint32_t* counts, int length, int limit, uint8_t* depth, HuffCache* cache) {
  // Node k: leaves [0, n), a sentinel at n, parents [n + 1, 2n), as parallel arrays. An attempt
  // tracks node heights while it merges, so one that cannot fit `limit` bits stops at the first
  // node that is too high and never walks its tree.
  constexpr int kMaxNodes = 2 * Histogram::kSize + 2;
  uint32_t cnt[kMaxNodes];
  int16_t lhs[kMaxNodes], rhs[kMaxNodes], sym[Histogram::kSize];
  uint16_t level[kMaxNodes], height[kMaxNodes];
  // Leaves ordered by (count ascending, symbol descending): a strict total order, so any sorting
  // algorithm gives the reference's arrangement. One 64-bit key per leaf: count << 16 | (0xffff - symbol).
  uint64_t keys[Histogram::kSize];
  int n = 0;
  if (cache && cache->n > 0) {
    // start from the previous call's order (the histogram moved by a few counts): insertion sort
    // is then close to linear
    int nonzero = 0;
    for (int i = 0; i < length; ++i) nonzero += counts[i] != 0;
    for (int k = 0; k < cache->n; ++k) {
      const int i = cache->order[k];
      if (i < length && counts[i]) keys[n++] = (static_cast<uint64_t>(counts[i]) << 16) | static_cast<uint64_t>(0xffff - i);
    }
    if (n != nonzero) {  // symbols the previous histogram did not have
      bool seen[Histogram::kSize] = {false};
      for (int k = 0; k < cache->n; ++k) seen[cache->order[k]] = true;
      for (int i = length; i-- > 0;)
        if (counts[i] && !seen[i]) keys[n++] = (static_cast<uint64_t>(counts[i]) << 16) | static_cast<uint64_t>(0xffff - i);
    }
    for (int k = 1; k < n; ++k) {
      const uint64_t v = keys[k];
      int j = k - 1;
      while (j >= 0 && keys[j] > v) { keys[j + 1] = keys[j]; --j; }
      keys[j + 1] = v;
    }
  } else {
    for (int i = length; i-- > 0;)
      if (counts[i]) keys[n++] = (static_cast<uint64_t>(counts[i]) << 16) | static_cast<uint64_t>(0xffff - i);
    if (n > 1) std::sort(keys, keys + n);
  }
  if (n == 0) return;
  for (int k = 0; k < n; ++k) {
    cnt[k] = static_cast<uint32_t>(keys[k] >> 16);
    sym[k] = static_cast<int16_t>(0xffff - static_cast<int>(keys[k] & 0xffff));
  }
  if (cache) {
    cache->n = n;
    std::copy(sym, sym + n, cache->order);
  }
  if (n == 1) {
    depth[sym[0]] = 1;
    return;
  }
  // Attempts with a rising count floor until the tree fits `limit` bits. Raising the floor only
  // changes the leading leaves (count <= floor): they all tie at the floor and must then order by
  // descending symbol; the tail keeps its order. So each attempt re-sorts just that prefix.
  for (int k = 0; k < n; ++k) height[k] = 0;
  int P = 0;  // leaves [0, P) have count <= floor_count
  for (uint32_t floor_count = 1;; floor_count *= 2) {
    if (floor_count > 1) {
      const int P0 = P;
      while (P < n && cnt[P] <= floor_count) ++P;
      if (P > P0) std::sort(sym, sym + P, [](int16_t x, int16_t y) { return x > y; });
      for (int k = 0; k < P; ++k) cnt[k] = floor_count;
    }
    cnt[n] = ~0u;
    cnt[n + 1] = ~0u;
    int i = 0, j = n + 1;
    bool fits = true;
    for (int parent = n + 1; parent < 2 * n; ++parent) {
      int left, right;
      if (cnt[i] <= cnt[j]) left = i++; else left = j++;
      if (cnt[i] <= cnt[j]) right = i++; else right = j++;
      cnt[parent] = cnt[left] + cnt[right];
      lhs[parent] = static_cast<int16_t>(left);
      rhs[parent] = static_cast<int16_t>(right);
      // height above the leaves: the root's is the longest code of this attempt
      const uint16_t hl = height[left], hr = height[right];
      const int hp = (hl > hr ? hl : hr) + 1;
      if (hp > limit) { fits = false; break; }  // the root can only be higher
      height[parent] = static_cast<uint16_t>(hp);
      cnt[parent + 1] = ~0u;
    }
    if (!fits) continue;
    // Children always precede their parent, so one downward sweep gives every node's level.
    level[2 * n - 1] = 0;
    for (int p = 2 * n - 1; p > n; --p) {
      const uint16_t d = static_cast<uint16_t>(level[p] + 1);
      level[lhs[p]] = d;
      level[rhs[p]] = d;
    }
    for (int k = 0; k < n; ++k) depth[sym[k]] = static_cast<uint8_t>(level[k]);
    return;
  }
}

size_t header_cost_bits(const Histogram& h) {
  size_t bits = 17 * 8;
  for (int i = 0; i + 1 < Histogram::kSize; ++i) bits += h.counts[i] > 0 ? 8 : 0;
  return bits;
}

size_t entropy_cost_bits(const Histogram& h, const uint8_t* depth) {
  size_t bits = 0;
  for (int i = 0; i + 1 < Histogram::kSize; ++i) bits += (h.counts[i] / 2) * (depth[i] + (i & 0xf));
  bits += (bits * 3 + 512) >> 10;  // escape-byte estimate
  return bits;
}

size_t cluster_histograms(Histogram* histo, size_t* num, int* indexes, uint8_t* depth, HuffCache* caches) {
  memset(depth, 0, *num * Histogram::kSize);
  size_t costs[4];
  for (size_t i = 0; i < *num; ++i) {
    indexes[i] = static_cast<int>(i);
    huffman_depths(histo[i].counts, Histogram::kSize, 16, depth + i * Histogram::kSize, caches ? &caches[i] : nullptr);
    costs[i] = header_cost_bits(histo[i]) + entropy_cost_bits(histo[i], depth + i * Histogram::kSize);
  }
  const size_t orig_num = *num;
  while (*num > 1) {
    const size_t last = *num - 1, prev = *num - 2;
    Histogram both(histo[last]);
    both.merge(histo[prev]);
    uint8_t depth_both[Histogram::kSize] = {0};
    huffman_depths(both.counts, Histogram::kSize, 16, depth_both, caches ? &caches[3 + (orig_num - *num)] : nullptr);
    const size_t cost_both = header_cost_bits(both) + entropy_cost_bits(both, depth_both);
    if (!(cost_both < costs[last] + costs[prev])) break;
    histo[prev] = both;
    histo[last] = Histogram();
    costs[prev] = cost_both;
    memcpy(depth + prev * Histogram::kSize, depth_both, sizeof(depth_both));
    for (size_t i = 0; i < orig_num; ++i)
      if (indexes[i] == static_cast<int>(last)) indexes[i] = static_cast<int>(prev);
    --*num;
  }
  size_t total = 0;
  for (size_t i = 0; i < *num; ++i) total += costs[i];
  return (total + 7) / 8;
}

// Bit k (1..63) set iff the coefficient at zig-zag position k is non-zero. Branch-free: the
// zero/non-zero pattern of real coefficient data is what makes a per-coefficient branch slow.
uint64_t zigzag_nonzero_mask(const int16_t* q) {
  uint64_t m = 0;
#pragma GCC unroll 63
  for (int k = 1; k < 64; ++k) m |= static_cast<uint64_t>(q[kNaturalOrder[k]] != 0) << k;
  return m;
}

void ac_histogram_add_block(const int16_t* q, int weight, Histogram* h) {
  uint64_t m = zigzag_nonzero_mask(q);
  int prev = 0;
  while (m) {
    const int k = __builtin_ctzll(m);
    m &= m - 1;
    int run = k - prev - 1;
    prev = k;
    while (run > 15) { h->add(0xf0, weight); run -= 16; }
    h->add((run << 4) + bit_length(static_cast<uint32_t>(std::abs(static_cast<int>(q[kNaturalOrder[k]])))), weight);
  }
  if (prev != 63) h->add(0, weight);
}

// ---------------------------------------------------------------------------------------------
// Frame helpers
// ---------------------------------------------------------------------------------------------
void frame_set_quant(Frame* f, const int q[3][64]) {
  f->num_tables = 0;
  for (int c = 0; c < f->ncomp; ++c) {
    int found = -1;
    for (int t = 0; t < f->num_tables; ++t)
      if (memcmp(q[c], f->table_values[t], 64 * sizeof(int)) == 0) { found = t; break; }
    if (found < 0) {
      found = f->num_tables++;
      memcpy(f->table_values[found], q[c], 64 * sizeof(int));
      f->table_index[found] = found;
    }
    f->comp_table[c] = found;
  }
}

void frame_set_quant_input(Frame* f, const int q[3][64]) {
  f->num_tables = 3;
  for (int c = 0; c < 3; ++c) {
    memcpy(f->table_values[c], q[c], 64 * sizeof(int));
    f->table_index[c] = 0;
    f->comp_table[c] = c;
  }
}

static int table_precision(const int* v) {
  for (int k = 0; k < 64; ++k) if (v[k] > 0xff) return 1;
  return 0;
}

size_t header_size(const Frame& f) {
  size_t n = 2 + 18 + 4;  // SOI, APP0, DQT marker
  for (int t = 0; t < f.num_tables; ++t) n += 1 + (table_precision(f.table_values[t]) ? 2 : 1) * 64;
  n += 10 + 3 * f.ncomp;  // SOF
  n += 4;                 // DHT without code data
  n += 8 + 2 * f.ncomp;   // SOS
  n += 2;                 // EOI
  return n;
}

static void build_dc_histograms(const Frame& f, Histogram* h) {
  const size_t nb = static_cast<size_t>(f.bw) * f.bh;
  for (int c = 0; c < f.ncomp; ++c) {
    int last = 0;
    const int16_t* p = f.coeffs[c];
    for (size_t b = 0; b < nb; ++b) {
      const int dc = p[b * 64];
      h[c].add(bit_length(static_cast<uint32_t>(std::abs(dc - last))));
      last = dc;
    }
  }
}

size_t estimate_dc_size(const Frame& f) {
  Histogram h[3];
  build_dc_histograms(f, h);
  size_t num = f.ncomp;
  int idx[4];
  uint8_t depth[3 * Histogram::kSize];
  return cluster_histograms(h, &num, idx, depth);
}

void build_ac_histograms(const Frame& f, Histogram* h) {
  const size_t nb = static_cast<size_t>(f.bw) * f.bh;
  for (int c = 0; c < f.ncomp; ++c) {
    const int16_t* p = f.coeffs[c];
    for (size_t b = 0; b < nb; ++b) ac_histogram_add_block(p + b * 64, 1, &h[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// Bit-stream writer
// ---------------------------------------------------------------------------------------------
namespace {

// Canonical code from depths; the sentinel (symbol 256, deepest, last) is dropped
// (BuildHuffmanCode + BuildHuffmanCodeTable, jpeg_data_writer.cc:136-187).
void make_code(const uint8_t* depth257, int counts[17], int values[257], CodeTable* table) {
  memset(counts, 0, 17 * sizeof(int));
  for (int i = 0; i < 257; ++i) if (depth257[i] > 0) ++counts[depth257[i]];
  int offset[17] = {0};
  for (int l = 1; l <= 16; ++l) offset[l] = offset[l - 1] + counts[l - 1];
  for (int i = 0; i < 257; ++i) if (depth257[i] > 0) values[offset[depth257[i]]++] = i;
  memset(table->depth, 255, sizeof(table->depth));
  memset(table->code, 0, sizeof(table->code));
  int total = 0;
  for (int l = 1; l <= 16; ++l) total += counts[l];
  if (total == 0) return;
  int code = 0, p = 0;
  for (int l = 1; l <= 16; ++l) {
    for (int i = 0; i < counts[l]; ++i, ++p) {
      if (p < total - 1) {  // all but the sentinel
        table->depth[values[p]] = static_cast<uint8_t>(l);
        table->code[values[p]] = static_cast<uint16_t>(code);
      }
      ++code;
    }
    code <<= 1;
  }
}

// MSB-first bit accumulator without byte stuffing (stuffing is applied when bands are stitched).
// Bits collect in a 64-bit register and leave 32 at a time into a buffer that the caller keeps
// large enough (reserve() before every block).
struct BitBuf {
  std::vector<uint8_t> bytes;   // valid prefix: [0, size_) after finish()
  uint8_t* ptr = nullptr;
  uint64_t acc = 0;
  int nacc = 0;                 // bits held in acc
  size_t total_bits = 0;
  void reserve(size_t extra) {
    const size_t used = ptr ? static_cast<size_t>(ptr - bytes.data()) : 0;
    if (bytes.size() < used + extra) {
      bytes.resize(std::max(bytes.size() * 2, used + extra + 4096));
      ptr = bytes.data() + used;
    }
    if (!ptr) ptr = bytes.data();
  }
  inline void put(int nbits, uint64_t bits) {  // nbits <= 32
    acc = (acc << nbits) | bits;
    nacc += nbits;
    if (nacc >= 32) {
      nacc -= 32;
      const uint32_t w = __builtin_bswap32(static_cast<uint32_t>(acc >> nacc));
      memcpy(ptr, &w, 4);
      ptr += 4;
    }
  }
  // Leaves whole bytes in `bytes` and the remaining (< 8) bits in acc/nacc.
  void finish() {
    size_t used = ptr ? static_cast<size_t>(ptr - bytes.data()) : 0;
    total_bits = used * 8 + nacc;
    reserve(8);
    while (nacc >= 8) {
      nacc -= 8;
      *ptr++ = static_cast<uint8_t>(acc >> nacc);
    }
    used = static_cast<size_t>(ptr - bytes.data());
    bytes.resize(used);
    ptr = nullptr;
  }
};

void encode_band(const Frame& f, const CodeTable* dc, const CodeTable* ac, int by0, int by1, BitBuf* out) {
  int last_dc[3] = {0, 0, 0};
  if (by0 > 0)
    for (int c = 0; c < f.ncomp; ++c) last_dc[c] = f.coeffs[c][(static_cast<size_t>(by0) * f.bw - 1) * 64];
  // code and magnitude bits of a symbol merged into one put: (code << nbits) | lowbits
  for (int by = by0; by < by1; ++by)
    for (int bx = 0; bx < f.bw; ++bx) {
      const size_t b = static_cast<size_t>(by) * f.bw + bx;
      out->reserve(static_cast<size_t>(f.ncomp) * 64 * 4 + 16);
      for (int c = 0; c < f.ncomp; ++c) {
        const int16_t* q = f.coeffs[c] + b * 64;
        const CodeTable& dct = dc[c];
        const CodeTable& act = ac[c];
        int diff = static_cast<int16_t>(q[0] - last_dc[c]);  // coeff_t arithmetic
        last_dc[c] = q[0];
        int mag = diff, low = diff;
        if (diff < 0) { mag = -diff; low = diff - 1; }
        mag = static_cast<int16_t>(mag);
        int nbits = bit_length(static_cast<uint32_t>(mag));
        out->put(dct.depth[nbits] + nbits,
                 (static_cast<uint64_t>(dct.code[nbits]) << nbits) | (static_cast<uint32_t>(low) & ((1u << nbits) - 1)));
        uint64_t m = zigzag_nonzero_mask(q);
        int prev = 0;
        while (m) {
          const int k = __builtin_ctzll(m);
          m &= m - 1;
          int run = k - prev - 1;
          prev = k;
          const int v = q[kNaturalOrder[k]];
          const int a = v < 0 ? -v : v;
          const int lowbits = v + (v >> 31);  // v < 0 ? ~a : a
          while (run > 15) { out->put(act.depth[0xf0], act.code[0xf0]); run -= 16; }
          nbits = bit_length(static_cast<uint32_t>(a));
          const int sym = (run << 4) + nbits;
          out->put(act.depth[sym] + nbits,
                   (static_cast<uint64_t>(act.code[sym]) << nbits) | (static_cast<uint32_t>(lowbits) & ((1u << nbits) - 1)));
        }
        if (prev != 63) out->put(act.depth[0], act.code[0]);
      }
    }
  out->finish();
}

inline void emit_stuffed(std::string* out, uint8_t b) {
  out->push_back(static_cast<char>(b));
  if (b == 0xff) out->push_back('\0');
}

}  // namespace

static double now_ms_() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

void build_histograms(const Frame& f, Histogram* dc, Histogram* ac, WorkerPool* pool) {
  const int T = pool ? std::max(1, std::min(pool->size(), f.bh)) : 1;
  if (T == 1) {
    for (int c = 0; c < f.ncomp; ++c) { dc[c] = Histogram(); ac[c] = Histogram(); }
    build_dc_histograms(f, dc);
    build_ac_histograms(f, ac);
    return;
  }
  std::vector<Histogram> part(static_cast<size_t>(T) * 6);
  pool->run(T, [&](int t) {
    const int y0 = static_cast<int>(static_cast<int64_t>(f.bh) * t / T);
    const int y1 = static_cast<int>(static_cast<int64_t>(f.bh) * (t + 1) / T);
    for (int c = 0; c < f.ncomp; ++c) {
      Histogram& hd = part[t * 6 + c];
      Histogram& ha = part[t * 6 + 3 + c];
      const int16_t* p = f.coeffs[c];
      const size_t b0 = static_cast<size_t>(y0) * f.bw, b1 = static_cast<size_t>(y1) * f.bw;
      int last = b0 > 0 ? p[(b0 - 1) * 64] : 0;
      for (size_t b = b0; b < b1; ++b) {
        const int v = p[b * 64];
        hd.add(bit_length(static_cast<uint32_t>(std::abs(v - last))));
        last = v;
        ac_histogram_add_block(p + b * 64, 1, &ha);
      }
    }
  });
  for (int c = 0; c < f.ncomp; ++c) {
    dc[c] = Histogram(); ac[c] = Histogram();
    for (int t = 0; t < T; ++t) { dc[c].merge(part[t * 6 + c]); ac[c].merge(part[t * 6 + 3 + c]); }
  }
}

void write_jpeg_header(const Frame& f, const Histogram* dc_hist, const Histogram* ac_hist, std::string* out,
                       CodeTable* dc_tab, CodeTable* ac_tab) {
  out->clear();
  const int ncomp = f.ncomp;
  // SOI + APP0 (stripped metadata always writes the fixed JFIF header)
  static const uint8_t kHead[] = {0xff, 0xd8, 0xff, 0xe0, 0x00, 0x10, 0x4a, 0x46, 0x49, 0x46, 0x00,
                                  0x01, 0x01, 0x00, 0x00, 0x01, 0x00, 0x01, 0x00, 0x00};
  out->append(reinterpret_cast<const char*>(kHead), sizeof(kHead));
  {  // DQT
    int len = 2;
    for (int t = 0; t < f.num_tables; ++t) len += 1 + (table_precision(f.table_values[t]) ? 2 : 1) * 64;
    out->push_back(static_cast<char>(0xff)); out->push_back(static_cast<char>(0xdb));
    out->push_back(static_cast<char>(len >> 8)); out->push_back(static_cast<char>(len & 0xff));
    for (int t = 0; t < f.num_tables; ++t) {
      const int prec = table_precision(f.table_values[t]);
      out->push_back(static_cast<char>((prec << 4) + f.table_index[t]));
      for (int k = 0; k < 64; ++k) {
        const int v = f.table_values[t][kNaturalOrder[k]];
        if (prec) out->push_back(static_cast<char>(v >> 8));
        out->push_back(static_cast<char>(v & 0xff));
      }
    }
  }
  {  // SOF (marker 0xc1, as the reference writes)
    const int len = 8 + 3 * ncomp;
    const uint8_t h[] = {0xff, 0xc1, static_cast<uint8_t>(len >> 8), static_cast<uint8_t>(len & 0xff), 8,
                         static_cast<uint8_t>(f.height >> 8), static_cast<uint8_t>(f.height & 0xff),
                         static_cast<uint8_t>(f.width >> 8), static_cast<uint8_t>(f.width & 0xff),
                         static_cast<uint8_t>(ncomp)};
    out->append(reinterpret_cast<const char*>(h), sizeof(h));
    for (int c = 0; c < ncomp; ++c) {
      out->push_back(static_cast<char>(c));      // component id
      out->push_back(static_cast<char>(f.yuv420 && ncomp == 3 && c == 0 ? 0x22 : 0x11));   // h/v sampling factors
      out->push_back(static_cast<char>(f.table_index[f.comp_table[c]]));
    }
  }
  // Huffman codes: DC histograms clustered, then AC histograms clustered.
  Histogram histo[6];
  uint8_t depths[6 * Histogram::kSize];
  for (int c = 0; c < ncomp; ++c) histo[c] = dc_hist[c];
  size_t num_dc = ncomp;
  int dc_idx[4], ac_idx[4];
  cluster_histograms(histo, &num_dc, dc_idx, depths);
  for (int c = 0; c < ncomp; ++c) histo[num_dc + c] = ac_hist[c];
  size_t num_ac = ncomp;
  cluster_histograms(histo + num_dc, &num_ac, ac_idx, depths + num_dc * Histogram::kSize);
  const int num_histo = static_cast<int>(num_dc + num_ac);
  int total_symbols = 0;
  for (int i = 0; i < num_histo; ++i) total_symbols += histo[i].num_symbols();
  const int dht_len = 2 + num_histo * 17 + total_symbols;
  out->push_back(static_cast<char>(0xff)); out->push_back(static_cast<char>(0xc4));
  out->push_back(static_cast<char>(dht_len >> 8)); out->push_back(static_cast<char>(dht_len & 0xff));
  for (int i = 0; i < num_histo; ++i) {
    const bool is_dc = i < static_cast<int>(num_dc);
    const int idx = is_dc ? i : i - static_cast<int>(num_dc);
    int counts[17], values[257] = {0};
    CodeTable table;
    make_code(depths + i * Histogram::kSize, counts, values, &table);
    for (int c = 0; c < ncomp; ++c) {
      if (is_dc && dc_idx[c] == idx) dc_tab[c] = table;
      if (!is_dc && ac_idx[c] == idx) ac_tab[c] = table;
    }
    int max_len = 16;
    while (max_len > 0 && counts[max_len] == 0) --max_len;
    --counts[max_len];
    int total = 0;
    for (int l = 1; l <= max_len; ++l) total += counts[l];
    out->push_back(static_cast<char>(is_dc ? i : idx + 0x10));
    for (int l = 1; l <= 16; ++l) out->push_back(static_cast<char>(counts[l]));
    for (int j = 0; j < total; ++j) out->push_back(static_cast<char>(values[j]));
  }
  {  // SOS
    const int len = 6 + 2 * ncomp;
    out->push_back(static_cast<char>(0xff)); out->push_back(static_cast<char>(0xda));
    out->push_back(static_cast<char>(len >> 8)); out->push_back(static_cast<char>(len & 0xff));
    out->push_back(static_cast<char>(ncomp));
    for (int c = 0; c < ncomp; ++c) {
      out->push_back(static_cast<char>(c));
      out->push_back(static_cast<char>((dc_idx[c] << 4) | ac_idx[c]));
    }
    out->push_back(0); out->push_back(63); out->push_back(0);
  }
}

void write_jpeg(const Frame& f, std::string* out, WorkerPool* pool, const Histogram* dc_hist,
                const Histogram* ac_hist, WriteTimers* tm) {
  double t0 = now_ms_();
  const int ncomp = f.ncomp;
  Histogram hdc[3], hac[3];
  if (!dc_hist || !ac_hist) {
    build_histograms(f, hdc, hac, pool);
    dc_hist = hdc;
    ac_hist = hac;
  }
  if (tm) { const double t1 = now_ms_(); tm->hist_ms += t1 - t0; t0 = t1; }
  CodeTable dc_tab[3], ac_tab[3];
  write_jpeg_header(f, dc_hist, ac_hist, out, dc_tab, ac_tab);
  (void)ncomp;
  if (tm) { const double t1 = now_ms_(); tm->code_ms += t1 - t0; t0 = t1; }
  // Entropy-coded segment: bands of block rows coded in parallel.
  const int T = pool ? std::max(1, std::min(pool->size(), f.bh)) : 1;
  std::vector<BitBuf> bands(T);
  auto band_rows = [&](int t, int* y0, int* y1) {
    *y0 = static_cast<int>(static_cast<int64_t>(f.bh) * t / T);
    *y1 = static_cast<int>(static_cast<int64_t>(f.bh) * (t + 1) / T);
  };
  if (T == 1) {
    encode_band(f, dc_tab, ac_tab, 0, f.bh, &bands[0]);
  } else {
    pool->run(T, [&](int t) {
      int y0, y1;
      band_rows(t, &y0, &y1);
      encode_band(f, dc_tab, ac_tab, y0, y1, &bands[t]);
    });
  }
  if (tm) { const double t1 = now_ms_(); tm->encode_ms += t1 - t0; t0 = t1; }
  bool parallel_stitch = T > 1;
  for (auto& b : bands) if (b.total_bits < 64) parallel_stitch = false;
  if (!parallel_stitch) {
    size_t est = 0;
    for (auto& b : bands) est += b.bytes.size() + 8;
    out->reserve(out->size() + est + est / 64 + 16);
    uint32_t acc = 0;  // pending bits (< 8) carried between bands
    int nacc = 0;
    for (auto& b : bands) {
      if (nacc == 0) {
        for (uint8_t v : b.bytes) emit_stuffed(out, v);
      } else {
        for (uint8_t v : b.bytes) {
          acc = (acc << 8) | v;
          emit_stuffed(out, static_cast<uint8_t>(acc >> nacc));
          acc &= (1u << nacc) - 1;
        }
      }
      if (b.nacc > 0) {  // band tail: b.nacc (< 8) leftover bits
        acc = (acc << b.nacc) | static_cast<uint32_t>(b.acc & ((1u << b.nacc) - 1));
        nacc += b.nacc;
        if (nacc >= 8) {
          nacc -= 8;
          emit_stuffed(out, static_cast<uint8_t>(acc >> nacc));
          acc &= (1u << nacc) - 1;
        }
      }
    }
    if (nacc > 0) {  // pad the last byte with ones (JumpToByteBoundary)
      const uint32_t pad = (1u << (8 - nacc)) - 1;
      emit_stuffed(out, static_cast<uint8_t>((acc << (8 - nacc)) | pad));
    }
  } else {
    // Parallel stitch. Band t contributes merged bits [off[t], off[t+1]); it owns the merged bytes
    // whose first bit falls in that range and borrows up to 7 head bits of the next band (ones
    // after the last band) for its final byte.
    std::vector<uint64_t> off(T + 1, 0);
    for (int t = 0; t < T; ++t) off[t + 1] = off[t] + bands[t].total_bits;
    std::vector<uint16_t> head16(T + 1, 0xffff);
    for (int t = 0; t < T; ++t) head16[t] = static_cast<uint16_t>((bands[t].bytes[0] << 8) | bands[t].bytes[1]);
    // pass 1: per band, append the two virtual tail bytes and count owned bytes and 0xff bytes
    std::vector<size_t> nbytes(T + 1, 0);
    pool->run(T, [&](int t) {
      BitBuf& b = bands[t];
      const uint32_t tailbits = b.nacc ? static_cast<uint32_t>(b.acc & ((1u << b.nacc) - 1)) : 0;
      const uint32_t V = ((tailbits << (16 - b.nacc)) | (static_cast<uint32_t>(head16[t + 1]) >> b.nacc)) & 0xffff;
      b.bytes.push_back(static_cast<uint8_t>(V >> 8));
      b.bytes.push_back(static_cast<uint8_t>(V & 0xff));
      const uint64_t j0 = (off[t] + 7) / 8, j1 = (off[t + 1] + 7) / 8;
      const int s0 = static_cast<int>(8 * j0 - off[t]);
      const size_t count = static_cast<size_t>(j1 - j0);
      const uint8_t* L = b.bytes.data();
      size_t ff = 0;
      if (s0 == 0) {
        for (size_t m = 0; m < count; ++m) ff += L[m] == 0xff;
      } else {
        for (size_t m = 0; m < count; ++m) ff += static_cast<uint8_t>(((L[m] << 8) | L[m + 1]) >> (8 - s0)) == 0xff;
      }
      nbytes[t + 1] = count + ff;
    });
    const size_t base = out->size();
    for (int t = 0; t < T; ++t) nbytes[t + 1] += nbytes[t];
    out->resize(base + nbytes[T]);
    char* dst0 = &(*out)[0] + base;
    // pass 2: shifted + stuffed bytes straight into the output
    pool->run(T, [&](int t) {
      const BitBuf& b = bands[t];
      const uint64_t j0 = (off[t] + 7) / 8, j1 = (off[t + 1] + 7) / 8;
      const int s0 = static_cast<int>(8 * j0 - off[t]);
      const size_t count = static_cast<size_t>(j1 - j0);
      const uint8_t* L = b.bytes.data();
      uint8_t* o = reinterpret_cast<uint8_t*>(dst0) + nbytes[t];
      if (s0 == 0) {
        for (size_t m = 0; m < count; ++m) { const uint8_t v = L[m]; *o++ = v; if (v == 0xff) *o++ = 0; }
      } else {
        for (size_t m = 0; m < count; ++m) {
          const uint8_t v = static_cast<uint8_t>(((L[m] << 8) | L[m + 1]) >> (8 - s0));
          *o++ = v;
          if (v == 0xff) *o++ = 0;
        }
      }
    });
  }
  out->push_back(static_cast<char>(0xff));
  out->push_back(static_cast<char>(0xd9));
  if (tm) { const double t1 = now_ms_(); tm->stitch_ms += t1 - t0; }
}

}  // namespace jpeg

// ---------------------------------------------------------------------------------------------
// WorkerPool
// ---------------------------------------------------------------------------------------------
WorkerPool::WorkerPool(int nthreads) : nthreads_(std::max(1, nthreads)) {
  for (int i = 1; i < nthreads_; ++i) threads_.emplace_back(&WorkerPool::worker, this);
}
WorkerPool::~WorkerPool() {
  {
    std::lock_guard<std::mutex> l(mu_);
    stop_.store(true);
  }
  cv_start_.notify_all();
  for (auto& t : threads_) t.join();
}

static inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

// Protocol: run() publishes a Job (on its stack), bumps the epoch, works, waits for pending == 0,
// clears job_ and then waits until no worker is still inside a job section (active_ == 0). A worker
// raises active_ BEFORE it reads job_, so it either sees nullptr or keeps run() from returning while
// it holds the pointer.
void WorkerPool::worker() {
  unsigned long long seen = 0;
  const int kSpin = 4000;   // ~20-50 us of polling before sleeping
  for (;;) {
    int spins = 0;
    while (epoch_.load(std::memory_order_acquire) == seen && !stop_.load(std::memory_order_relaxed)) {
      if (++spins < kSpin) { cpu_relax(); continue; }
      std::unique_lock<std::mutex> l(mu_);
      sleepers_.fetch_add(1);
      cv_start_.wait(l, [&] { return stop_.load() || epoch_.load() != seen; });
      sleepers_.fetch_sub(1);
    }
    if (stop_.load()) return;
    seen = epoch_.load(std::memory_order_acquire);
    active_.fetch_add(1);
    Job* j = job_.load();
    if (j) {
      for (;;) {
        const int i = j->next.fetch_add(1);
        if (i >= j->n) break;
        (*j->fn)(i);
        j->pending.fetch_sub(1);
      }
    }
    active_.fetch_sub(1);
  }
}

void WorkerPool::run(int n, const std::function<void(int)>& fn) {
  if (n <= 0) return;
  if (nthreads_ == 1 || n == 1 || async_open_) { for (int i = 0; i < n; ++i) fn(i); return; }
  begin(n, fn);
  finish();
}

void WorkerPool::begin(int n, const std::function<void(int)>& fn) {
  Job& job = async_job_;
  job.fn = &fn;
  job.n = std::max(0, n);
  job.next.store(0);
  job.pending.store(job.n);
  async_open_ = true;
  if (job.n == 0 || nthreads_ == 1) return;   // (finish() does the items)
  job_.store(&job);
  epoch_.fetch_add(1, std::memory_order_release);
  if (sleepers_.load() > 0) {
    std::lock_guard<std::mutex> l(mu_);   // a sleeper checks the epoch under this mutex
    cv_start_.notify_all();
  }
}

void WorkerPool::finish() {
  if (!async_open_) return;
  Job& job = async_job_;
  for (;;) {  // the caller works too
    const int i = job.next.fetch_add(1);
    if (i >= job.n) break;
    (*job.fn)(i);
    job.pending.fetch_sub(1);
  }
  int spins = 0;
  while (job.pending.load(std::memory_order_acquire) != 0) { if (++spins < 64) cpu_relax(); else std::this_thread::yield(); }
  job_.store(nullptr);
  while (active_.load() != 0) cpu_relax();
  async_open_ = false;
}

}  // namespace gzb
