// gzb_quant_search.h -- SelectQuantMatrix (guetzli/processor.cc:310-372) as a resumable search that
// can evaluate several TryQuantMatrix candidates at once, one per GPU of a group.
//
// The reference visits trials strictly one after the other: the "original" (q = 1 input), the
// all-ones matrix, then the matrices QuantMatrixGenerator::GetNext (processor.cc:206-271) derives
// from the dist_ok bits of the earlier trials (expansion by total_csf steps, then bisection on the
// heuristic score). A trial's *result* depends only on its matrix, so the group speculates: from the
// current generator state it enumerates the decision tree of not-yet-known dist_ok outcomes
// (best-first by path probability), hands the first `world` distinct matrices to ranks 0..world-1,
// all-gathers {distance, jpg_size} (16 bytes per rank) and then REPLAYS the reference's sequential
// logic over the cached results in the reference's visiting order. Decisions, the verbose trace and
// the chosen matrix are therefore identical to the single-GPU run for any world size; speculative
// results that the replay never reaches are simply unused. A rank may also take `batch` trials per
// round (their host legs -- quantise + Huffman-code the file -- run concurrently on host threads
// while the GPU compares them one after the other), so even a single GPU needs fewer serial rounds.
#pragma once
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <functional>
#include <queue>
#include <string>
#include <vector>

#include "gzb_jpeg.h"

namespace gzb {

// ---- QuantMatrixGenerator (guetzli/processor.cc:162-308) --------------------------------------
struct QuantData { int q[3][64]; size_t jpg_size; bool dist_ok; };

inline double contrast_sensitivity(int k) { return 1.0 / (1.0 + jpeg::kZigZag[k] / 2.0); }

inline double quant_heuristic_score(const int q[3][64]) {
  double score = 0.0;
  for (int c = 0; c < 3; ++c)
    for (int k = 0; k < 64; ++k) score += 0.5 * (q[c][k] - 1.0) * contrast_sensitivity(k);
  return score;
}

class QuantGenerator {
 public:
  // downsample: the search of a YUV420 image starts at heuristic score 0 (processor.cc:224)
  explicit QuantGenerator(bool downsample = false) : a_(-1.0), b_(-1.0), total_csf_(0.0), downsample_(downsample) {
    for (int k = 0; k < 64; ++k) total_csf_ += 3.0 * contrast_sensitivity(k);
  }
  bool next(int q[3][64]) {
    for (int iter = 0; iter < 1000; ++iter) {
      double hscore;
      if (b_ == -1.0) {
        if (a_ == -1.0) hscore = downsample_ ? 0.0 : total_csf_;
        else if (a_ < 5.0 * total_csf_) hscore = a_ + total_csf_;
        else hscore = 2 * (a_ + total_csf_);
        if (hscore > 100 * total_csf_) return false;
      } else if (b_ == 0.0) {
        return false;
      } else if (a_ == -1.0) {
        hscore = 0.0;
      } else {
        int lo[3][64], hi[3][64];
        const double eps = 0.05;
        matrix_for((1 - eps) * a_ + eps * 0.5 * (a_ + b_), lo);
        matrix_for((1 - eps) * b_ + eps * 0.5 * (a_ + b_), hi);
        if (memcmp(lo, hi, sizeof(lo)) == 0) return false;
        hscore = (a_ + b_) * 0.5;
      }
      matrix_for(hscore, q);
      bool retry = false;
      for (const QuantData& d : seen_)
        if (memcmp(q, d.q, sizeof(d.q)) == 0) {
          if (d.dist_ok) a_ = hscore; else b_ = hscore;
          retry = true;
          break;
        }
      if (!retry) return true;
    }
    return false;
  }
  void add(const QuantData& d) {
    seen_.push_back(d);
    const double hs = quant_heuristic_score(d.q);
    if (d.dist_ok) a_ = std::max(a_, hs);
    else b_ = b_ == -1.0 ? hs : std::min(b_, hs);
  }
  bool expanding() const { return b_ == -1.0; }

 private:
  void matrix_for(double score, int q[3][64]) const {
    const int level = static_cast<int>(score / total_csf_);
    score -= level * total_csf_;
    for (int k = 63; k >= 0; --k) {
      const int nat = jpeg::kNaturalOrder[k];
      for (int c = 0; c < 3; ++c) q[c][nat] = 2 * level + (score > 0.0 ? 3 : 1);
      score -= 3.0 * contrast_sensitivity(nat);
    }
  }
  double a_, b_, total_csf_;
  bool downsample_;
  std::vector<QuantData> seen_;
};

inline bool quant_data_better(const QuantData& a, const QuantData& b) {
  if (a.dist_ok && !b.dist_ok) return true;
  if (!a.dist_ok && b.dist_ok) return false;
  return a.jpg_size < b.jpg_size;
}

// ---- the group -----------------------------------------------------------------------------------
// all-gather of `nbytes` from every rank into recv[world * nbytes], rank-major; returns 0 on success.
typedef int (*AllGatherFn)(void* user, const void* send, size_t nbytes, void* recv);
struct Group {
  int rank = 0, world = 1;
  AllGatherFn allgather = nullptr;
  void* user = nullptr;
  // optional: all-gather of device memory, for the bulky exchange of the zeroing candidates (gzb_allgather_device_fn)
  int (*allgather_device)(void* user, const void* d_send, size_t nbytes, void* d_recv) = nullptr;
};

// One TryQuantMatrix-like trial: the q=1 "original" written with the input's tables
// (processor.cc:967-985), or CopyFromJpegData + ApplyGlobalQuantization(q) (processor.cc:279-308).
struct Trial {
  int original;
  int q[3][64];
  bool operator==(const Trial& o) const { return original == o.original && memcmp(q, o.q, sizeof(q)) == 0; }
};
struct TrialOutcome {
  float distance = 0.f;
  uint64_t jpg_size = 0;
  int owner = 0;      // rank that evaluated it (its JPEG bytes exist there only)
  std::string jpeg;   // on the owner: the file's header (everything up to the scan) ...
  std::string scan;   // ... and its entropy-coded segment before byte stuffing
  // single-rank groups leave the scan on the device until the trial turns out to be the best so far:
  uint64_t scan_bytes = 0, ff_bytes = 0;
  unsigned long long resident_gen = 0;   // != 0: the scan is the device's coding number `resident_gen`
};

class QuantSearch {
 public:
  // evaluate: run the trial on this rank's GPU. visit: called once per trial in the reference's order.
  typedef std::function<bool(const std::vector<Trial>&, std::vector<TrialOutcome>*)> EvalFn;
  typedef std::function<void(const Trial&, const TrialOutcome&)> VisitFn;

  // mode 0: original, all-ones matrix, generator loop (the 4:4:4 pass of ProcessJpegData);
  // mode 1: the original only; mode 2: all-ones matrix + generator loop of a downsampled image.
  QuantSearch(const Group& g, float target, int batch = 1, int mode = 0)
      : g_(g), target_(target), batch_(std::max(1, batch)), mode_(mode) {}

  bool run(const EvalFn& evaluate, const VisitFn& visit) {
    State st;
    if (mode_ == 2) { st.phase = 1; st.gen = QuantGenerator(true); }
    st.only_original = mode_ == 1;
    for (;;) {
      Trial t;
      if (!st.next(&t)) break;
      const TrialOutcome* o = find(t);
      if (!o) {
        if (!round(st, evaluate)) return false;
        o = find(t);
        if (!o) return false;  // cannot happen: the needed trial is always first in the round
      }
      visit(t, *o);
      ++visited_;
      const bool ok = dist_ok(o->distance, 0.97f);   // target_mul_high (processor.cc:347)
      const int phase = st.phase;
      st.advance(ok, o->jpg_size);
      if (phase == 1) {
        best_ = st.last;
      } else if (phase == 2 && quant_data_better(st.last, best_)) {
        best_ = st.last;
        if (best_.dist_ok && !dist_ok(o->distance, 0.95f)) break;  // target_mul_low (processor.cc:348, 362)
      }
    }
    return true;
  }
  const QuantData& best() const { return best_; }
  int rounds() const { return rounds_; }
  int evaluated_here() const { return evaluated_here_; }
  int evaluated_total() const { return static_cast<int>(cache_.size()); }
  int visited() const { return visited_; }

  bool dist_ok(float distance, float mul) const {  // DistanceOK (butteraugli_comparator.h:52-54)
    return static_cast<double>(distance) <= static_cast<double>(mul) * static_cast<double>(target_);
  }

 private:
  struct State {
    int phase = 0;  // 0: original, 1: all-ones matrix, 2: generator loop, 3: exhausted
    QuantGenerator gen;
    QuantData last;
    bool next(Trial* t) {
      if (phase <= 1) {
        t->original = phase == 0;
        for (int c = 0; c < 3; ++c) for (int k = 0; k < 64; ++k) t->q[c][k] = 1;
        return true;
      }
      if (phase == 2) {
        t->original = 0;
        if (gen.next(t->q)) { memcpy(pending, t->q, sizeof(pending)); return true; }
        phase = 3;
      }
      return false;
    }
    void advance(bool ok, size_t jpg_size) {  // the generator-visible effect of one trial
      if (phase == 0) { phase = only_original ? 3 : 1; return; }
      if (phase == 1) {
        for (int c = 0; c < 3; ++c) for (int k = 0; k < 64; ++k) last.q[c][k] = 1;
      } else {
        memcpy(last.q, pending, sizeof(pending));
      }
      last.dist_ok = ok;
      last.jpg_size = jpg_size;
      if (phase == 2) gen.add(last);
      phase = 2;
    }
    int pending[3][64];
    bool only_original = false;
  };
  struct Node {
    double p;
    int seq;
    State st;
    bool operator<(const Node& o) const { return p != o.p ? p < o.p : seq > o.seq; }
  };

  const TrialOutcome* find(const Trial& t) const {
    for (size_t i = 0; i < keys_.size(); ++i) if (keys_[i] == t) return &cache_[i];
    return nullptr;
  }

  // Enumerates up to `world` trials the reference may visit next and evaluates one per rank.
  bool round(const State& from, const EvalFn& evaluate) {
    std::vector<Trial> list;
    std::priority_queue<Node> pq;
    int seq = 0;
    pq.push(Node{1.0, seq++, from});
    int expanded = 0;
    const int want = g_.world * batch_;
    while (static_cast<int>(list.size()) < want && !pq.empty() && expanded < 64 * want) {
      Node n = pq.top();
      pq.pop();
      ++expanded;
      Trial t;
      if (!n.st.next(&t)) continue;
      const TrialOutcome* known = find(t);
      bool listed = false;
      for (const Trial& l : list) listed = listed || l == t;
      if (!known && !listed) list.push_back(t);
      if (known) {
        n.st.advance(dist_ok(known->distance, 0.97f), known->jpg_size);
        pq.push(Node{n.p, seq++, n.st});
      } else if (n.st.phase < 2) {
        n.st.advance(true, 0);   // the generator ignores the first two trials' outcome
        pq.push(Node{n.p, seq++, n.st});
      } else {
        // while no failing matrix is known the next one usually still passes; in the bisection
        // both outcomes are equally likely
        const double p_ok = n.st.gen.expanding() ? 0.65 : 0.5;
        Node yes{n.p * p_ok, seq++, n.st}, no{n.p * (1.0 - p_ok), seq++, n.st};
        yes.st.advance(true, 0);
        no.st.advance(false, 0);
        pq.push(yes);
        pq.push(no);
      }
    }
    ++rounds_;
    // list[r * batch + j] goes to rank r
    struct Rec { int32_t valid; float distance; uint64_t jpg_size; };
    std::vector<Rec> mine(batch_, Rec{0, 0.f, 0});
    std::vector<Trial> my_trials;
    for (int j = 0; j < batch_; ++j) {
      const size_t k = static_cast<size_t>(g_.rank) * batch_ + j;
      if (k < list.size()) my_trials.push_back(list[k]);
    }
    std::vector<TrialOutcome> local(my_trials.size());
    if (!my_trials.empty()) {
      // A rank whose evaluation fails still takes part in the exchange below with valid = 0 for its trials:
      // every rank then sees the missing outcome and leaves together, instead of the others waiting for
      // this rank inside the collective.
      if (evaluate(my_trials, &local)) {
        evaluated_here_ += static_cast<int>(my_trials.size());
        for (size_t j = 0; j < my_trials.size(); ++j) mine[j] = Rec{1, local[j].distance, local[j].jpg_size};
      } else if (g_.world == 1) {
        return false;
      }
    }
    std::vector<Rec> all(static_cast<size_t>(g_.world) * batch_);
    if (g_.world > 1) {
      if (!g_.allgather || g_.allgather(g_.user, mine.data(), sizeof(Rec) * batch_, all.data()) != 0) return false;
    } else {
      all = mine;
    }
    for (size_t k = 0; k < list.size(); ++k) {
      if (!all[k].valid) return false;
      const int r = static_cast<int>(k / batch_);
      TrialOutcome o;
      o.distance = all[k].distance;
      o.jpg_size = all[k].jpg_size;
      o.owner = r;
      if (r == g_.rank) {
        TrialOutcome& mine_o = local[k - static_cast<size_t>(r) * batch_];
        o.jpeg.swap(mine_o.jpeg);
        o.scan.swap(mine_o.scan);
        o.scan_bytes = mine_o.scan_bytes;
        o.ff_bytes = mine_o.ff_bytes;
        o.resident_gen = mine_o.resident_gen;
      }
      keys_.push_back(list[k]);
      cache_.push_back(std::move(o));
    }
    return true;
  }

  Group g_;
  float target_;
  int batch_, mode_;
  std::vector<Trial> keys_;
  std::vector<TrialOutcome> cache_;
  QuantData best_{};
  int rounds_ = 0, evaluated_here_ = 0, visited_ = 0;
};

}  // namespace gzb
