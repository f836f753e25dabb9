// gzb_backend.cuh -- SelectFrequencyBackEnd (guetzli/processor.cc:723-919) on the device.
//
// Per adjustment iteration the reference builds `global_order` -- one (block, value) pair for every
// remaining zeroing candidate of every block with a non-zero weight, up to 10^7 pairs at 12 MPix --
// std::sort()s it and walks it from the front until the estimated file size has moved far enough.
// The walk cannot stop (nor be observed) before `min_coeffs_to_change` entries, so all but its last
// few hundred to few thousand steps are a SET of entries: the first p of the sorted order, in any
// order. Here everything up to that point stays in HBM:
//
//   k_be_count / k_scan_counts / k_be_fill : the order, in the reference's block-major arrangement
//       (processor.cc:786-813), values (err - max_err) / weight in IEEE float like the host code;
//   k_be_tiles_* / k_be_swap / k_be_local : libstdc++'s introsort
//       (std::sort = __introsort_loop + __final_insertion_sort) evaluated lazily: they run the library's
//       median-of-three + unguarded Hoare partition steps -- the same element swaps, so that ties
//       between blocks end up exactly where std::sort puts them -- only on the ranges that straddle
//       p, postponing every right-hand range on a device-side stack, until the straddling range is
//       short enough for the host to finish it with the library's own code. A partition of a long
//       range is data parallel although the library's loop is sequential: the loop swaps the k-th
//       element from the left that is not less than the pivot with the k-th from the right that is
//       not greater, while the former lies left of the latter; both lists can be read off the
//       array before any swap, so ranks come from prefix sums and the swaps commute;
//   k_be_prefix_count / k_be_apply_prefix : each block takes as many of its next candidates as it
//       has entries among the first p (processor.cc:854-876 for a whole prefix at once);
//   k_be_gather : the state of the few blocks the sequential walk will touch (quantised indices,
//       last_index), for the host;
//   k_be_apply_walk / k_be_update_max_err : the walked flips and processor.cc:893-895.
//
// The array never leaves the device; the host sees the short range around p (a few KB).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "gzb_device_math.cuh"

namespace gzb {

// std::pair<int, float> of the reference's global_order, as raw 64 bits: low word = block index,
// high word = the float's bits (little endian: .first at the lower address).
typedef unsigned long long BeEntry;
__host__ __device__ __forceinline__ BeEntry be_make(int block, float val) {
#ifdef __CUDA_ARCH__
  return static_cast<unsigned long long>(static_cast<unsigned>(block)) | (static_cast<unsigned long long>(__float_as_uint(val)) << 32);
#else
  unsigned u;
  memcpy(&u, &val, 4);
  return static_cast<unsigned long long>(static_cast<unsigned>(block)) | (static_cast<unsigned long long>(u) << 32);
#endif
}
__device__ __forceinline__ float be_val(BeEntry e) { return __uint_as_float(static_cast<unsigned>(e >> 32)); }
__device__ __forceinline__ int be_block(BeEntry e) { return static_cast<int>(static_cast<unsigned>(e)); }

struct BeRange { unsigned first, last; int depth, pad; };
constexpr int kBeStack = 128;      // >= 2 * log2(n) + 2 pending ranges
constexpr int kBeTile = 2048;      // entries per partition tile (256 threads x 8)
constexpr int kBeThreads = 256;
constexpr int kBeSmallMax = 8192;  // entries one select can hand to the host (capacity of BeState::small), blocks one gather returns
constexpr unsigned kBeLocalMax = 16384;   // ranges up to this length are partitioned by one CTA (no grid barriers)
constexpr int kBeMaxRet = 16;      // short ranges one gzb_be_select_ranges can hand over

enum { BE_RUNNING = 0, BE_SMALL = 1, BE_HEAP = 2, BE_EMPTY = 3 };

// Device-resident state of one back-end pass; the head (up to `stack`) is what the host reads back.
struct BeState {
  // ---- results of the order build ----
  unsigned n;                 // entries in the order
  unsigned blocks_to_change;  // blocks that contributed at least one entry
  unsigned below;             // entries with value < limit (partition_point of the first "up" iteration)
  unsigned changed_blocks;    // blocks flipped by the prefix of this iteration
  // ---- lazy sort ----
  int status;                 // BE_*
  int top;                    // pending ranges on the stack; stack[top - 1] is the leftmost
  unsigned p_set;             // ranges that end at or before p_set are consumed as a set (dropped unsorted)
  unsigned small_max;         // a range of at most this many entries goes to the host
  unsigned levels, levels_total;
  // the partition in flight
  unsigned first, last;       // range being partitioned; the pivot sits at `first`
  int depth;
  float pv;
  unsigned ntiles, NL, NR, K;
  unsigned done_count, done_swap;   // CTAs that have finished the count / swap phase of the level in flight
  // short ranges handed to the host by this select (consecutive, leftmost first; entries back to back in `small`)
  unsigned want_end;          // keep handing over ranges while they end before this position (and fit)
  unsigned nret, ret_total, pad_;
  BeRange ret[kBeMaxRet];
  BeRange stack[kBeStack];
};
static_assert(sizeof(BeState) % 8 == 0, "the 8-byte entries of the handed-over ranges follow the state");

__device__ __forceinline__ BeEntry be_ld(const BeEntry* p) { return __ldcg(p); }
__device__ __forceinline__ void be_st(BeEntry* p, BeEntry v) { __stcg(p, v); }

// std::__move_median_to_first(first, first + 1, mid, last - 1) with comp(a, b) = a.second < b.second.
__device__ inline void be_median_to_first(BeEntry* a, unsigned first, unsigned last) {
  const unsigned ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
  const BeEntry ea = be_ld(a + ia), eb = be_ld(a + ib), ec = be_ld(a + ic);
  const float va = be_val(ea), vb = be_val(eb), vc = be_val(ec);
  unsigned pick;
  if (va < vb) {
    if (vb < vc) pick = ib;
    else if (va < vc) pick = ic;
    else pick = ia;
  } else if (va < vc) pick = ia;
  else if (vb < vc) pick = ic;
  else pick = ib;
  const BeEntry ef = be_ld(a + first);
  const BeEntry ep = pick == ia ? ea : pick == ib ? eb : ec;
  be_st(a + first, ep);
  be_st(a + pick, ef);
}

// The control step between two partitions (one thread): drops consumed ranges, stops at a short range
// or an exhausted depth budget, else pops the leftmost range and moves its pivot into place.
__device__ inline void be_next_range(BeState* st, BeEntry* a) {
  for (;;) {
    if (st->top == 0) { st->status = BE_EMPTY; return; }
    const BeRange r = st->stack[st->top - 1];
    if (r.last <= st->p_set) { --st->top; continue; }
    if (r.last - r.first <= st->small_max) { st->status = BE_SMALL; return; }
    if (r.depth == 0) { st->status = BE_HEAP; return; }
    --st->top;
    st->first = r.first;
    st->last = r.last;
    st->depth = r.depth - 1;
    be_median_to_first(a, r.first, r.last);
    st->pv = be_val(be_ld(a + r.first));
    st->ntiles = (r.last - r.first - 1 + kBeTile - 1) / kBeTile;
    st->K = 0;
    st->done_count = 0;
    st->done_swap = 0;
    st->status = BE_RUNNING;
    ++st->levels;
    return;
  }
}

// ---- one std::__unguarded_partition_pivot of the range in *st (pivot already at `first`) -----------
// Five phases with a grid-wide dependency between them. A long range runs them as three launches on a
// grid of CTAs (k_be_tiles_count, k_be_tiles_lists, k_be_swap); a range of at most kBeLocalMax entries is
// partitioned by ONE CTA in shared memory (k_be_local), level after level while the ranges stay short.
// No kernel ever waits for another CTA, so any number of contexts can sort concurrently.
struct BeLevel {
  unsigned first, last, m, ntiles;
  float pv;
  BeEntry* A;   // a + first + 1: the entries the two scans walk
};
__device__ __forceinline__ BeLevel be_level(BeEntry* a, const BeState* st) {
  const volatile BeState* v = st;
  BeLevel L;
  L.first = v->first; L.last = v->last; L.pv = v->pv; L.ntiles = v->ntiles;
  L.m = L.last - L.first - 1;
  L.A = a + L.first + 1;
  return L;
}
// a grid-level partition is in flight (the range is long)?
__device__ __forceinline__ bool be_grid_level(const BeState* st) {
  const volatile BeState* v = st;
  return v->status == BE_RUNNING && v->last - v->first > kBeLocalMax;
}

struct BeSmem {
  unsigned cell_l[8][8], cell_r[8][8];   // [iteration][warp] stopper counts of a tile
  unsigned scan[kBeThreads];
  unsigned tile_r;
};

// phase 1: per-tile stopper counts
__device__ __forceinline__ void be_phase_count(const BeLevel& L, unsigned tile0, unsigned tile_step, unsigned* tcl, unsigned* tcr, BeSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (unsigned t = tile0; t < L.ntiles; t += tile_step) {
    unsigned nl = 0, nr = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned i = t * kBeTile + j * kBeThreads + tid;
      bool fl = false, fr = false;
      if (i < L.m) {
        const float v = be_val(be_ld(L.A + i));
        fl = !(v < L.pv);
        fr = !(L.pv < v);
      }
      nl += __popc(__ballot_sync(0xffffffffu, fl));
      nr += __popc(__ballot_sync(0xffffffffu, fr));
    }
    if (lane == 0) { sm.cell_l[0][warp] = nl; sm.cell_r[0][warp] = nr; }
    __syncthreads();
    if (tid == 0) {
      unsigned sl = 0, sr = 0;
      for (int w = 0; w < 8; ++w) { sl += sm.cell_l[0][w]; sr += sm.cell_r[0][w]; }
      __stcg(tcl + t, sl);
      __stcg(tcr + t, sr);
    }
    __syncthreads();
  }
}

// phase 2 (one CTA): tile offsets. tcl[t] <- stoppers left of tile t; tcr[t] <- right stoppers right of tile t
__device__ __forceinline__ void be_phase_scan(const BeLevel& L, unsigned* tcl, unsigned* tcr, BeState* st, BeSmem& sm) {
  const int tid = threadIdx.x;
  const unsigned per = (L.ntiles + kBeThreads - 1) / kBeThreads;
  const unsigned t0 = min(L.ntiles, tid * per), t1 = min(L.ntiles, t0 + per);
  unsigned sl = 0, sr = 0;
  for (unsigned t = t0; t < t1; ++t) { sl += __ldcg(tcl + t); sr += __ldcg(tcr + t); }
  // inclusive scans of the per-thread sums (left counts forward, right counts backward)
  sm.scan[tid] = sl;
  __syncthreads();
  for (int off = 1; off < kBeThreads; off <<= 1) {
    const unsigned v = tid >= off ? sm.scan[tid - off] : 0;
    __syncthreads();
    sm.scan[tid] += v;
    __syncthreads();
  }
  unsigned runl = sm.scan[tid] - sl;
  const unsigned NL = sm.scan[kBeThreads - 1];
  __syncthreads();
  sm.scan[tid] = sr;
  __syncthreads();
  for (int off = 1; off < kBeThreads; off <<= 1) {
    const unsigned v = tid + off < kBeThreads ? sm.scan[tid + off] : 0;
    __syncthreads();
    sm.scan[tid] += v;
    __syncthreads();
  }
  const unsigned NR = sm.scan[0];
  unsigned runr = sm.scan[tid];   // right stoppers in this thread's tiles and everything to their right
  for (unsigned t = t0; t < t1; ++t) {
    const unsigned cl = __ldcg(tcl + t), cr = __ldcg(tcr + t);
    __stcg(tcl + t, runl);
    runl += cl;
    runr -= cr;
    __stcg(tcr + t, runr);
  }
  if (tid == 0) { st->NL = NL; st->NR = NR; }
  __syncthreads();
}

// phase 3: the two stopper lists
__device__ __forceinline__ void be_phase_lists(const BeLevel& L, unsigned tile0, unsigned tile_step, const unsigned* tcl, const unsigned* tcr,
                                               unsigned* lpos, unsigned* rpos, BeSmem& sm) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (unsigned t = tile0; t < L.ntiles; t += tile_step) {
    unsigned bl[8], br[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned i = t * kBeTile + j * kBeThreads + tid;
      bool fl = false, fr = false;
      if (i < L.m) {
        const float v = be_val(be_ld(L.A + i));
        fl = !(v < L.pv);
        fr = !(L.pv < v);
      }
      bl[j] = __ballot_sync(0xffffffffu, fl);
      br[j] = __ballot_sync(0xffffffffu, fr);
      if (lane == 0) { sm.cell_l[j][warp] = __popc(bl[j]); sm.cell_r[j][warp] = __popc(br[j]); }
    }
    __syncthreads();
    // exclusive scan of the 64 cells in index order (iteration-major, then warp): two cells per lane of warp 0
    if (warp == 0) {
      unsigned* cl = &sm.cell_l[0][0];
      unsigned* cr = &sm.cell_r[0][0];
      const unsigned l0 = cl[2 * lane], l1 = cl[2 * lane + 1], r0 = cr[2 * lane], r1 = cr[2 * lane + 1];
      unsigned sl = l0 + l1, sr = r0 + r1;
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        const unsigned vl = __shfl_up_sync(0xffffffffu, sl, off), vr = __shfl_up_sync(0xffffffffu, sr, off);
        if (lane >= off) { sl += vl; sr += vr; }
      }
      cl[2 * lane] = sl - l0 - l1; cl[2 * lane + 1] = sl - l1;
      cr[2 * lane] = sr - r0 - r1; cr[2 * lane + 1] = sr - r1;
      if (lane == 31) sm.tile_r = sr;   // right stoppers in the tile
    }
    __syncthreads();
    const unsigned base_l = __ldcg(tcl + t);
    // right stoppers of this tile take ranks [tcr[t], tcr[t] + tile_r) counted from the right
    const unsigned base_r = __ldcg(tcr + t), tile_r = sm.tile_r;
    const unsigned lt = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const unsigned i = t * kBeTile + j * kBeThreads + tid;
      if (bl[j] >> lane & 1u) __stcg(lpos + base_l + sm.cell_l[j][warp] + __popc(bl[j] & lt), i);
      if (br[j] >> lane & 1u) {
        const unsigned asc = sm.cell_r[j][warp] + __popc(br[j] & lt);   // rank from the left inside the tile
        __stcg(rpos + base_r + (tile_r - 1u - asc), i);
      }
    }
    __syncthreads();
  }
}

// phase 4: the swaps. Pair k is swapped iff lpos[k] < rpos[k] (monotone in k); K = their number
__device__ __forceinline__ void be_phase_swap(const BeLevel& L, unsigned k0, unsigned kstep, const unsigned* lpos, const unsigned* rpos, BeState* st) {
  const volatile BeState* v = st;
  const unsigned lim = min(v->NL, v->NR);
  unsigned cnt = 0;
  for (unsigned k = k0; k < lim; k += kstep) {
    const unsigned l = __ldcg(lpos + k), r = __ldcg(rpos + k);
    if (!(l < r)) break;
    const BeEntry el = be_ld(L.A + l), er = be_ld(L.A + r);
    be_st(L.A + l, er);
    be_st(L.A + r, el);
    ++cnt;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if ((threadIdx.x & 31) == 0 && cnt) atomicAdd(&st->K, cnt);
}

// phase 5 (one thread): the cut, the two new ranges, the next pivot
__device__ __forceinline__ void be_phase_finalize(const BeLevel& L, BeEntry* a, const unsigned* lpos, const unsigned* rpos, BeState* st) {
  const volatile BeState* v = st;
  const unsigned K = v->K, NL = v->NL;
  unsigned cut = L.m;   // the scans are guarded by the median-of-three: a stopper exists
  if (K < NL) cut = min(cut, __ldcg(lpos + K));
  if (K > 0) cut = min(cut, __ldcg(rpos + K - 1));
  const unsigned gcut = L.first + 1 + cut;
  const int depth = v->depth;
  st->stack[st->top++] = BeRange{gcut, L.last, depth, 0};
  st->stack[st->top++] = BeRange{L.first, gcut, depth, 0};
  be_next_range(st, a);
}

// Three launches per long level: the last CTA to finish the count phase scans the tile counts, and the
// last one to finish the swaps takes the control step (ticket counters; nobody waits for anybody).
__device__ __forceinline__ bool be_last_cta(unsigned* counter) {
  __shared__ bool s_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  __syncthreads();
  if (s_last) __threadfence();
  return s_last;
}
__global__ void __launch_bounds__(kBeThreads)
k_be_tiles_count(BeEntry* a, unsigned* tcl, unsigned* tcr, BeState* st) {
  __shared__ BeSmem sm;
  if (!be_grid_level(st)) return;
  const BeLevel L = be_level(a, st);
  be_phase_count(L, blockIdx.x, gridDim.x, tcl, tcr, sm);
  if (be_last_cta(&st->done_count)) be_phase_scan(L, tcl, tcr, st, sm);
}
__global__ void __launch_bounds__(kBeThreads)
k_be_tiles_lists(BeEntry* a, unsigned* lpos, unsigned* rpos, const unsigned* tcl, const unsigned* tcr, BeState* st) {
  __shared__ BeSmem sm;
  if (!be_grid_level(st)) return;
  be_phase_lists(be_level(a, st), blockIdx.x, gridDim.x, tcl, tcr, lpos, rpos, sm);
}
__global__ void __launch_bounds__(kBeThreads)
k_be_swap(BeEntry* a, const unsigned* lpos, const unsigned* rpos, BeState* st) {
  if (!be_grid_level(st)) return;
  const BeLevel L = be_level(a, st);
  be_phase_swap(L, blockIdx.x * kBeThreads + threadIdx.x, gridDim.x * kBeThreads, lpos, rpos, st);
  if (be_last_cta(&st->done_swap) && threadIdx.x == 0) be_phase_finalize(L, a, lpos, rpos, st);
}

// ---- short ranges: one CTA, the range in shared memory ------------------------------------------------
// A range of at most kBeLocalMax entries is loaded into shared memory once and partitioned there level after
// level (the library's partition restated with ballots: the k-th stopper from either side is found by a
// binary search over per-chunk counts instead of from stored lists), then written back. When the sort has
// stopped at a range for the host, the kernel copies it out; if the caller wants more than that (want_end)
// and the next pending range is again one this CTA can partition, it carries on with it in the same launch:
// every range handed over saves the host a round trip.
//   small : receives the entries of the ranges handed over (BeState::ret), back to back
constexpr int kBeLocalThreads = 1024;
constexpr int kBeLocalChunks = kBeLocalMax / 32;   // ballot words of a level
struct BeLocalSmem {
  BeState st;                                          // working copy of the state
  unsigned bl[kBeLocalChunks], br[kBeLocalChunks];     // stopper ballots: bit = entry is not less / not greater than the pivot
  unsigned exl[kBeLocalChunks];                        // left stoppers before chunk c
  unsigned inr[kBeLocalChunks + 1];                    // right stoppers in chunks >= c
  unsigned wsum[2][32];
  unsigned NL, NR, K;
  unsigned wlo, whi;                                   // entries [wlo, whi) of the order are held in shared memory (none: whi == wlo)
  int action;
};
constexpr size_t kBeLocalSmemBytes = sizeof(BeEntry) * kBeLocalMax + sizeof(BeLocalSmem);
enum { BE_ACT_LEVEL = 0, BE_ACT_LOAD = 1, BE_ACT_LOAD_PIVOT = 2, BE_ACT_BIG = 3, BE_ACT_STOP = 4, BE_ACT_EXIT = 5 };

// position of the n-th (0-based, from bit 0) set bit of m
__device__ __forceinline__ unsigned be_nth_set_bit(unsigned m, unsigned n) {
  unsigned pos = 0;
#pragma unroll
  for (int w = 16; w >= 1; w >>= 1) {
    const unsigned c = __popc((m >> pos) & ((1u << w) - 1u));
    if (n >= c) { n -= c; pos += w; }
  }
  return pos;
}
// position (relative to A) of the right stopper of rank k from the right / the left stopper of rank k from the left
__device__ __forceinline__ unsigned be_local_select_right(const BeLocalSmem& s, unsigned nchunks, unsigned k) {
  unsigned lo = 0, hi = nchunks;   // inr[lo] > k >= inr[hi]
  while (hi - lo > 1) {
    const unsigned mid = (lo + hi) >> 1;
    if (s.inr[mid] > k) lo = mid; else hi = mid;
  }
  const unsigned m = s.br[lo], j = k - s.inr[lo + 1];   // rank j counted from the top bit
  return 32u * lo + be_nth_set_bit(m, __popc(m) - 1u - j);
}
__device__ __forceinline__ unsigned be_local_select_left(const BeLocalSmem& s, unsigned nchunks, unsigned k) {
  unsigned lo = 0, hi = nchunks;   // exl[lo] <= k < exl[hi]
  while (hi - lo > 1) {
    const unsigned mid = (lo + hi) >> 1;
    if (s.exl[mid] <= k) lo = mid; else hi = mid;
  }
  return 32u * lo + be_nth_set_bit(s.bl[lo], k - s.exl[lo]);
}

// std::__move_median_to_first on the window (indices relative to it)
__device__ inline void be_local_median(BeEntry* e, unsigned first, unsigned last) {
  const unsigned ia = first + 1, ib = first + (last - first) / 2, ic = last - 1;
  const BeEntry ea = e[ia], eb = e[ib], ec = e[ic];
  const float va = be_val(ea), vb = be_val(eb), vc = be_val(ec);
  unsigned pick;
  if (va < vb) {
    if (vb < vc) pick = ib;
    else if (va < vc) pick = ic;
    else pick = ia;
  } else if (va < vc) pick = ia;
  else if (vb < vc) pick = ic;
  else pick = ib;
  const BeEntry ef = e[first];
  e[first] = pick == ia ? ea : pick == ib ? eb : ec;
  e[pick] = ef;
}

// one std::__unguarded_partition_pivot of [st.first, st.last) inside the window (pivot at its first entry),
// then (thread 0) the two new ranges
__device__ __forceinline__ void be_local_level(BeEntry* e, BeLocalSmem& s) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const unsigned f = s.st.first - s.wlo, m = s.st.last - s.st.first - 1;
  BeEntry* A = e + f + 1;
  const float pv = s.st.pv;
  const unsigned nchunks = (m + 31) >> 5;
  for (unsigned c = warp; c < nchunks; c += kBeLocalThreads / 32) {
    const unsigned i = 32 * c + lane;
    bool fl = false, fr = false;
    if (i < m) {
      const float v = be_val(A[i]);
      fl = !(v < pv);
      fr = !(pv < v);
    }
    const unsigned b1 = __ballot_sync(0xffffffffu, fl), b2 = __ballot_sync(0xffffffffu, fr);
    if (lane == 0) { s.bl[c] = b1; s.br[c] = b2; }
  }
  __syncthreads();
  // chunk t belongs to thread t: left counts scanned forward, right counts backward
  unsigned cl = 0, cr = 0;
  if (tid < kBeLocalChunks && static_cast<unsigned>(tid) < nchunks) { cl = __popc(s.bl[tid]); cr = __popc(s.br[tid]); }
  unsigned il = cl, ir = cr;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const unsigned vl = __shfl_up_sync(0xffffffffu, il, off), vr = __shfl_down_sync(0xffffffffu, ir, off);
    if (lane >= off) il += vl;
    if (lane + off < 32) ir += vr;
  }
  if (warp < kBeLocalChunks / 32) {
    if (lane == 31) s.wsum[0][warp] = il;
    if (lane == 0) s.wsum[1][warp] = ir;
  }
  __syncthreads();
  if (warp == 0) {
    const bool in = lane < kBeLocalChunks / 32;
    const unsigned wl = in ? s.wsum[0][lane] : 0u, wr = in ? s.wsum[1][lane] : 0u;
    unsigned sl = wl, sr = wr;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned vl = __shfl_up_sync(0xffffffffu, sl, off), vr = __shfl_down_sync(0xffffffffu, sr, off);
      if (lane >= off) sl += vl;
      if (lane + off < 32) sr += vr;
    }
    s.wsum[0][lane] = sl - wl;   // left stoppers in the warps before this one
    s.wsum[1][lane] = sr - wr;   // right stoppers in the warps after this one
    if (lane == 31) s.NL = sl;
    if (lane == 0) { s.NR = sr; s.K = 0; s.inr[kBeLocalChunks] = 0; }
  }
  __syncthreads();
  if (tid < kBeLocalChunks) {
    s.exl[tid] = s.wsum[0][warp] + il - cl;
    s.inr[tid] = s.wsum[1][warp] + ir;
  }
  __syncthreads();
  // the swaps: pair k is swapped iff its left stopper lies left of its right stopper (monotone in k)
  const unsigned NL = s.NL, NR = s.NR, lim = min(NL, NR);
  unsigned cnt = 0;
  for (unsigned c = warp; c < nchunks; c += kBeLocalThreads / 32) {
    const unsigned b1 = s.bl[c];
    if (b1 >> lane & 1u) {
      const unsigned k = s.exl[c] + __popc(b1 & ((1u << lane) - 1u));
      if (k < lim) {
        const unsigned i = 32 * c + lane, r = be_local_select_right(s, nchunks, k);
        if (i < r) {
          const BeEntry x = A[i], y = A[r];
          A[i] = y;
          A[r] = x;
          ++cnt;
        }
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, off);
  if (lane == 0 && cnt) atomicAdd(&s.K, cnt);
  __syncthreads();
  if (tid == 0) {
    const unsigned K = s.K;
    unsigned cut = m;   // the scans are guarded by the median-of-three: a stopper exists
    if (K < NL) cut = min(cut, be_local_select_left(s, nchunks, K));
    if (K > 0) cut = min(cut, be_local_select_right(s, nchunks, K - 1));
    const unsigned gcut = s.st.first + 1 + cut;
    const int depth = s.st.depth;
    s.st.stack[s.st.top++] = BeRange{gcut, s.st.last, depth, 0};
    s.st.stack[s.st.top++] = BeRange{s.st.first, gcut, depth, 0};
  }
}

// be_next_range on the working copy (thread 0): the action the CTA takes next
__device__ inline int be_local_next(BeLocalSmem& s, BeEntry* e) {
  BeState& st = s.st;
  for (;;) {
    if (st.top == 0) { st.status = BE_EMPTY; return BE_ACT_STOP; }
    const BeRange r = st.stack[st.top - 1];
    if (r.last <= st.p_set) { --st.top; continue; }
    const unsigned len = r.last - r.first;
    if (len <= st.small_max) { st.status = BE_SMALL; return BE_ACT_STOP; }
    if (r.depth == 0) { st.status = BE_HEAP; return BE_ACT_STOP; }
    if (len > kBeLocalMax) return BE_ACT_BIG;   // a long range: the grid kernels' turn
    --st.top;
    st.first = r.first;
    st.last = r.last;
    st.depth = r.depth - 1;
    st.status = BE_RUNNING;
    ++st.levels;
    if (s.whi > s.wlo && r.first >= s.wlo && r.last <= s.whi) {
      be_local_median(e, r.first - s.wlo, r.last - s.wlo);
      st.pv = be_val(e[r.first - s.wlo]);
      return BE_ACT_LEVEL;
    }
    return BE_ACT_LOAD_PIVOT;
  }
}

__global__ void __launch_bounds__(kBeLocalThreads)
k_be_local(BeEntry* a, BeState* st, BeEntry* small) {
  extern __shared__ unsigned long long be_local_dyn[];
  BeEntry* e = be_local_dyn;
  BeLocalSmem& s = *reinterpret_cast<BeLocalSmem*>(be_local_dyn + kBeLocalMax);
  const int tid = threadIdx.x;
  {
    const unsigned* src = reinterpret_cast<const unsigned*>(st);
    unsigned* dst = reinterpret_cast<unsigned*>(&s.st);
    for (unsigned i = tid; i < sizeof(BeState) / 4; i += kBeLocalThreads) dst[i] = __ldcg(src + i);
  }
  if (tid == 0) { s.wlo = 0; s.whi = 0; }
  __syncthreads();
  if (tid == 0) {
    // (a range in flight has its pivot in place already: the control step that started it ran on the array in HBM)
    if (s.st.status == BE_RUNNING) s.action = s.st.last - s.st.first <= kBeLocalMax ? BE_ACT_LOAD : BE_ACT_EXIT;
    else s.action = BE_ACT_STOP;
  }
  __syncthreads();
  auto flush = [&]() {
    const unsigned wlo = s.wlo, n = s.whi - s.wlo;
    for (unsigned i = tid; i < n; i += kBeLocalThreads) be_st(a + wlo + i, e[i]);
  };
  for (;;) {
    int act = s.action;
    if (act == BE_ACT_LOAD || act == BE_ACT_LOAD_PIVOT) {
      flush();
      __syncthreads();
      const unsigned first = s.st.first, n = s.st.last - s.st.first;
      for (unsigned i = tid; i < n; i += kBeLocalThreads) e[i] = be_ld(a + first + i);
      __syncthreads();
      if (tid == 0) {
        s.wlo = first;
        s.whi = first + n;
        if (act == BE_ACT_LOAD_PIVOT) {
          be_local_median(e, 0, n);
          s.st.pv = be_val(e[0]);
        }
      }
      __syncthreads();
      act = BE_ACT_LEVEL;
    }
    if (act == BE_ACT_LEVEL) {
      be_local_level(e, s);
      if (tid == 0) s.action = be_local_next(s, e);
      __syncthreads();
      continue;
    }
    if (act == BE_ACT_STOP && s.st.status == BE_SMALL) {
      // hand the range over; maybe go on with the next one
      const BeRange r = s.st.stack[s.st.top - 1];
      const unsigned base = s.st.ret_total, n = r.last - r.first;
      if (s.whi > s.wlo && r.first >= s.wlo && r.last <= s.whi) {
        for (unsigned i = tid; i < n; i += kBeLocalThreads) small[base + i] = e[r.first - s.wlo + i];
      } else {
        for (unsigned i = tid; i < n; i += kBeLocalThreads) small[base + i] = be_ld(a + r.first + i);
      }
      __syncthreads();
      if (tid == 0) {
        BeState& w = s.st;
        w.ret[w.nret++] = r;
        w.ret_total = base + n;
        int next = BE_ACT_EXIT;
        if (w.nret < static_cast<unsigned>(kBeMaxRet) && r.last < w.want_end && w.top >= 2) {
          const BeRange nx = w.stack[w.top - 2];
          const unsigned nlen = nx.last - nx.first;
          if (nlen <= kBeLocalMax && (nlen <= w.small_max || nx.depth > 0) &&
              w.ret_total + min(nlen, w.small_max) <= static_cast<unsigned>(kBeSmallMax)) {
            w.p_set = r.last;   // the range just handed over is the host's now
            next = be_local_next(s, e);
          }
        }
        s.action = next;
      }
      __syncthreads();
      if (s.action == BE_ACT_EXIT) break;
      continue;
    }
    break;   // BE_ACT_BIG, BE_ACT_EXIT, or stopped at an exhausted depth budget / an empty stack
  }
  const int last_act = s.action;
  __syncthreads();
  flush();
  {
    unsigned* dst = reinterpret_cast<unsigned*>(st);
    const unsigned* src = reinterpret_cast<const unsigned*>(&s.st);
    for (unsigned i = tid; i < sizeof(BeState) / 4; i += kBeLocalThreads) __stcg(dst + i, src[i]);
  }
  if (last_act == BE_ACT_BIG) {
    // the next range is a long one: its control step (median of three, level set-up) runs on the array in HBM
    __threadfence();
    __syncthreads();
    if (tid == 0) be_next_range(st, a);
  }
}

// ---- the order ---------------------------------------------------------------------------------
// Candidate list addressing with the reference's clamp (processor.cc:789-791, 859): a block whose
// offset equals the total reads the last entry of the arrays.
struct BeCands {
  const int* off;          // [num_blocks + 1]
  const uint8_t* idx;      // [total]
  const float* err;        // [total]
  int total;
};
__device__ __forceinline__ int be_cand_offset(const BeCands& c, int b) { return max(0, min(c.off[b], c.total - 1)); }

// entries block b contributes (processor.cc:794-811)
__global__ void k_be_count(BeCands c, const float* __restrict__ weight, const int* __restrict__ last_index, int num_blocks,
                           int direction, int* __restrict__ counts, BeState* st) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  int cnt = 0;
  if (b < num_blocks && weight[b] != 0.f) {
    const int offset = be_cand_offset(c, b);
    const int num = c.off[b + 1] - offset;
    const int li = last_index[b];
    cnt = direction > 0 ? max(0, num - li) : max(0, li);
  }
  if (b < num_blocks) counts[b] = cnt;
  const unsigned any = __popc(__ballot_sync(0xffffffffu, cnt > 0));
  if ((threadIdx.x & 31) == 0 && any) atomicAdd(&st->blocks_to_change, any);
}

// warp per block: its entries in the reference's order, at the block's offset of the scan
__global__ void k_be_fill(BeCands c, const float* __restrict__ weight, const int* __restrict__ last_index,
                          const float* __restrict__ max_err, const int* __restrict__ counts, const int* __restrict__ offsets,
                          int num_blocks, int direction, float limit, BeEntry* __restrict__ order, BeState* st) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= num_blocks) return;
  const int cnt = counts[b];
  if (cnt == 0) return;
  const int offset = be_cand_offset(c, b);
  const int li = last_index[b];
  const float me = max_err[b], w = weight[b];
  const float* errs = c.err + offset;
  BeEntry* o = order + offsets[b];
  unsigned below = 0;
  for (int j = lane; j < cnt; j += 32) {
    const float e = errs[direction > 0 ? li + j : li - 1 - j];
    const float val = direction > 0 ? (e - me) / w : (me - e) / w;
    o[j] = be_make(b, val);
    below += val < limit ? 1u : 0u;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) below += __shfl_xor_sync(0xffffffffu, below, off);
  if (lane == 0 && below) atomicAdd(&st->below, below);
}

__global__ void k_be_set_n(const int* __restrict__ offsets, int num_blocks, BeState* st) {
  st->n = static_cast<unsigned>(offsets[num_blocks]);
}

// (re)starts the lazy sort over the n entries: one pending range with introsort's depth budget
__global__ void k_be_sort_begin(BeState* st, int depth_override = -1) {
  const unsigned n = st->n;
  st->top = 0;
  st->levels_total = 0;
  if (n >= 1) {   // (a single entry is a short range: the host takes it as it is)
    // 2 * std::__lg(n); a test may start with a smaller budget to reach the heap-sort fallback
    st->stack[0] = BeRange{0u, n, depth_override >= 0 ? depth_override : 2 * (31 - __clz(n)), 0};
    st->top = 1;
  }
}

// starts a gzb_be_select: the parameters, then the first control step
__global__ void k_be_select_begin(BeEntry* a, BeState* st, unsigned p_set, unsigned small_max, unsigned want_end) {
  st->p_set = p_set;
  st->small_max = small_max;
  st->want_end = want_end;
  st->nret = 0;
  st->ret_total = 0;
  st->levels = 0;
  be_next_range(st, a);
}

// ---- the prefix, consumed as a set ---------------------------------------------------------------
__global__ void k_be_prefix_count(const BeEntry* __restrict__ order, unsigned p, unsigned* __restrict__ pcount) {
  const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p) atomicAdd(&pcount[be_block(order[i])], 1u);
}

// Geometry of the pass: unit b of the search owns coefficient block (b / pass_bw) * coef_bw + b % pass_bw
// of the searched planes (the 4:2:0 luma plane is MCU-padded).
struct BeGeom {
  int num_blocks, pass_bw, coef_bw; size_t cs;
  int factor, bw, bh;        // sampling factor of the searched planes; 8x8 blocks of the image
  uint8_t* blk_changed;      // one byte per 8x8 image block: its samples changed since the last Compare
};
// Marks the image blocks whose samples a flip in unit b changes: the unit's own block, or -- for a
// sub-sampled chroma block -- the 2x2 luma blocks of its macro-block plus one block all around (the fancy
// upsampling reaches one sample past the macro-block).
__device__ __forceinline__ void be_mark_changed(const BeGeom& g, int b) {
  if (!g.blk_changed) return;
  if (g.factor == 1) { g.blk_changed[b] = 1; return; }
  const int mx = b % g.pass_bw, my = b / g.pass_bw;
  for (int by = max(2 * my - 1, 0); by <= min(2 * my + 2, g.bh - 1); ++by)
    for (int bx = max(2 * mx - 1, 0); bx <= min(2 * mx + 2, g.bw - 1); ++bx) g.blk_changed[by * g.bw + bx] = 1;
}
__device__ __forceinline__ size_t be_cblock(const BeGeom& g, int b) {
  return static_cast<size_t>(b / g.pass_bw) * g.coef_bw + b % g.pass_bw;
}

// thread per block: the next pcount[b] candidates of the block are flipped (processor.cc:858-875)
__global__ void k_be_apply_prefix(BeCands c, BeGeom g, const unsigned* __restrict__ pcount, int direction,
                                  const int16_t* __restrict__ orig, const int* __restrict__ q192,
                                  int16_t* __restrict__ coef, int* __restrict__ last_index, BeState* st) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  unsigned times = 0;
  if (b < g.num_blocks) times = pcount[b];
  const unsigned any = __popc(__ballot_sync(0xffffffffu, times > 0));
  if ((threadIdx.x & 31) == 0 && any) atomicAdd(&st->changed_blocks, any);
  if (!times) return;
  be_mark_changed(g, b);
  const uint8_t* cands = c.idx + be_cand_offset(c, b);
  const size_t cb = be_cblock(g, b);
  int li = last_index[b];
  for (unsigned rep = 0; rep < times; ++rep) {
    const int cidx = cands[li + min(direction, 0)];
    const int comp = cidx >> 6, k = cidx & 63;
    const size_t at = comp * g.cs + cb * 64 + k;
    coef[at] = direction > 0 ? static_cast<int16_t>(0) : static_cast<int16_t>(quantize_coeff(orig[at], q192[cidx]));
    li += direction;
  }
  last_index[b] = li;
}

// ---- the walk's blocks ---------------------------------------------------------------------------
// One record per requested block: what the sequential walk on the host needs to flip the block's next
// candidates and to price the symbols that change.
struct BeBlockState {
  int last_index;
  unsigned prefix_count;      // entries of the block in this iteration's prefix (> 0: already counted as changed)
  unsigned long long zmask[3];   // bit z set: the coefficient at zig-zag position z of the component is non-zero
  int16_t idx[3][64];         // quantised indices (coefficient / q) of the three components
  int16_t requant[3][64];     // Quantize(original, q): the value a "down" step restores
};
__constant__ uint8_t c_be_zigzag[64] = {   // natural index -> zig-zag position
    0,  1,  5,  6,  14, 15, 27, 28, 2,  4,  7,  13, 16, 26, 29, 42, 3,  8,  12, 17, 25, 30, 41, 43, 9,  11, 18, 24, 31, 40, 44, 53,
    10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61, 35, 36, 48, 49, 57, 58, 62, 63};
__global__ void k_be_gather(BeGeom g, const int* __restrict__ blocks, int nreq, const int16_t* __restrict__ coef,
                            const int16_t* __restrict__ orig, const int* __restrict__ q192, const int* __restrict__ last_index,
                            const unsigned* __restrict__ pcount, int comp_mask, int want_requant, char* __restrict__ out, size_t out_stride) {
  // (records out_stride bytes apart: without the requantised values only the head of each record exists, so that
  // what goes to the host is one contiguous copy)
  const int r = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (r >= nreq) return;
  const int b = blocks[r];
  const size_t cb = be_cblock(g, b);
  BeBlockState* o = reinterpret_cast<BeBlockState*>(out + static_cast<size_t>(r) * out_stride);
  if (lane == 0) { o->last_index = last_index[b]; o->prefix_count = pcount[b]; }
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int i = lane + 32 * j, comp = i >> 6, k = i & 63;   // a warp covers half a component per step
    int16_t vi = 0, vr = 0;
    if (comp_mask >> comp & 1) {
      const size_t at = comp * g.cs + cb * 64 + k;
      const int q = q192[i];
      vi = static_cast<int16_t>(coef[at] / q);
      if (want_requant) vr = static_cast<int16_t>(quantize_coeff(orig[at], q));
    }
    o->idx[comp][k] = vi;
    if (want_requant) o->requant[comp][k] = vr;
    // non-zero mask in zig-zag positions: OR over the warp (two steps per component)
    unsigned long long m = vi != 0 ? 1ull << c_be_zigzag[k] : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) m |= __shfl_xor_sync(0xffffffffu, m, off);
    if (lane == 0) {
      if (j & 1) o->zmask[comp] |= m; else o->zmask[comp] = m;
    }
  }
}

// the flips of the sequential walk that stay applied
__global__ void k_be_apply_walk(BeGeom g, const int* __restrict__ blocks, const uint8_t* __restrict__ cidx,
                                const int16_t* __restrict__ val, int n, int direction, int16_t* __restrict__ coef,
                                int* __restrict__ last_index) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int b = blocks[i], ci = cidx[i];
  coef[(ci >> 6) * g.cs + be_cblock(g, b) * 64 + (ci & 63)] = val[i];
  atomicAdd(&last_index[b], direction);
  be_mark_changed(g, b);
}

// processor.cc:893-895
__global__ void k_be_update_max_err(const float* __restrict__ weight, float val_threshold, int direction, int num_blocks,
                                    float* __restrict__ max_err) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= num_blocks) return;
  const float d = (weight[b] * val_threshold) * static_cast<float>(direction);
  max_err[b] = max_err[b] + d;
}

// IsGrayscale (processor.cc:921-929): any non-zero coefficient in the chroma planes of the q=1 input?
__global__ void k_any_nonzero(const int16_t* __restrict__ p, size_t n, unsigned* flag) {
  size_t i = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) * 8;
  bool nz = false;
  if (i + 8 <= n) {
    const uint4 v = *reinterpret_cast<const uint4*>(p + i);
    nz = (v.x | v.y | v.z | v.w) != 0;
  } else {
    for (; i < n; ++i) nz = nz || p[i] != 0;
  }
  if (__any_sync(0xffffffffu, nz) && (threadIdx.x & 31) == 0) atomicOr(flag, 1u);
}

}  // namespace gzb
