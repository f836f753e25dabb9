"""CPU tests of the multi-GPU SelectQuantMatrix search (gzb_quant_search.h): for any group size the
speculative, one-trial-per-rank search must visit exactly the trials -- in exactly the order -- of
the reference's sequential loop (guetzli/processor.cc:310-372), restated here in Python with
QuantMatrixGenerator (processor.cc:162-308). The GPU trial is replaced by a deterministic function
of the matrix; the all-gather is a thread barrier (and gloo in test_multirank_cpu.py)."""
import math
import threading

import numpy as np
import pytest

import __graft_entry__ as ge

ZIGZAG = [0, 1, 5, 6, 14, 15, 27, 28, 2, 4, 7, 13, 16, 26, 29, 42, 3, 8, 12, 17, 25, 30, 41, 43, 9, 11, 18, 24, 31,
          40, 44, 53, 10, 19, 23, 32, 39, 45, 52, 54, 20, 22, 33, 38, 46, 51, 55, 60, 21, 34, 37, 47, 50, 56, 59, 61,
          35, 36, 48, 49, 57, 58, 62, 63]
NATURAL = [ZIGZAG.index(i) for i in range(64)]


def csf(k):
    return 1.0 / (1.0 + ZIGZAG[k] / 2.0)


def hscore(q):
    s = 0.0
    for c in range(3):
        for k in range(64):
            s += 0.5 * (q[64 * c + k] - 1.0) * csf(k)
    return s


class RefGenerator:
    """QuantMatrixGenerator, processor.cc:162-271."""

    def __init__(self, downsample=False):
        self.downsample = downsample
        self.a = self.b = -1.0
        self.total = 0.0
        for k in range(64):
            self.total += 3.0 * csf(k)
        self.seen = []

    def matrix(self, score):
        level = int(score / self.total)
        score -= level * self.total
        q = [0] * 192
        for k in range(63, -1, -1):
            nat = NATURAL[k]
            for c in range(3):
                q[64 * c + nat] = 2 * level + (3 if score > 0.0 else 1)
            score -= 3.0 * csf(nat)
        return q

    def next(self):
        for _ in range(1000):
            if self.b == -1.0:
                if self.a == -1.0:
                    hs = 0.0 if self.downsample else self.total
                elif self.a < 5.0 * self.total:
                    hs = self.a + self.total
                else:
                    hs = 2 * (self.a + self.total)
                if hs > 100 * self.total:
                    return None
            elif self.b == 0.0:
                return None
            elif self.a == -1.0:
                hs = 0.0
            else:
                eps = 0.05
                if self.matrix((1 - eps) * self.a + eps * 0.5 * (self.a + self.b)) == \
                        self.matrix((1 - eps) * self.b + eps * 0.5 * (self.a + self.b)):
                    return None
                hs = (self.a + self.b) * 0.5
            q = self.matrix(hs)
            hit = [d for d in self.seen if d[0] == q]
            if hit:
                if hit[0][1]:
                    self.a = hs
                else:
                    self.b = hs
                continue
            return q
        return None

    def add(self, q, ok):
        self.seen.append((q, ok))
        hs = hscore(q)
        if ok:
            self.a = max(self.a, hs)
        else:
            self.b = hs if self.b == -1.0 else min(self.b, hs)


def reference_search(eval_fn, target, downsample=False):
    """processor.cc:986-1003 (original) + SelectQuantMatrix 336-372; downsample: the YUV420 pass (no
    original, QuantMatrixGenerator(downsample=true))."""
    f32 = np.float32

    def ok_at(d, mul):
        return float(f32(d)) <= float(f32(mul)) * float(f32(target))
    visited = []
    ones = [1] * 192
    if not downsample:
        d, s = eval_fn(True, ones)
        visited.append((1.0, 0.0, float(f32(d)), float(s)))
    d, s = eval_fn(False, ones)
    visited.append((0.0, 0.0, float(f32(d)), float(s)))
    best = (ones, ok_at(d, 0.97), s)
    gen = RefGenerator(downsample)
    while True:
        q = gen.next()
        if q is None:
            break
        d, s = eval_fn(False, q)
        visited.append((0.0, hscore(q), float(f32(d)), float(s)))
        data = (q, ok_at(d, 0.97), s)
        gen.add(q, data[1])
        better = (data[1] and not best[1]) or (data[1] == best[1] and data[2] < best[2])
        if better:
            best = data
            if data[1] and not ok_at(d, 0.95):
                break
    return visited, best


def make_eval(kind):
    def ev(original, q):
        h = hscore(q)
        wob = 0.013 * math.sin(h * 0.37)
        if kind == "typical":
            d = 0.43 + 0.0175 * h + wob
        elif kind == "never_ok":
            d = 3.0 + 0.01 * h
        elif kind == "always_ok":
            d = 0.2 + 1e-4 * h
        elif kind == "early_exit":
            d = 0.9 + 0.004 * h            # quickly lands between 0.95 and 0.97 of the target
        else:
            d = 0.6 + 0.03 * h + 5 * wob   # non-monotone around the threshold
        size = int(150000 / (1.0 + 0.05 * h)) + (77 if original else 0)
        return d, size
    return ev


class BarrierAllGather:
    def __init__(self, world):
        self.world, self.slots, self.bar = world, [None] * world, threading.Barrier(world)

    def for_rank(self, r):
        def ag(buf):
            self.slots[r] = bytes(buf)
            self.bar.wait()
            out = b"".join(self.slots)
            self.bar.wait()
            return out
        return ag


def run_group(gz, world, target, ev, batch=1, mode=0):
    if world == 1:
        return [gz.QuantSearchSimulate(0, 1, None, target, ev, batch, mode)]
    ag = BarrierAllGather(world)
    res = [None] * world

    def work(r):
        res[r] = gz.QuantSearchSimulate(r, world, ag.for_rank(r), target, ev, batch, mode)
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(120)
    assert all(r is not None for r in res)
    return res


@pytest.fixture(scope="module")
def gz():
    return ge.build()


@pytest.mark.parametrize("kind", ["typical", "never_ok", "always_ok", "early_exit", "wobbly"])
@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_group_search_replays_the_reference_sequence(gz, kind, world):
    target = 0.971769 if kind != "wobbly" else 1.473608
    ev = make_eval(kind)
    want_visited, want_best = reference_search(ev, target)
    res = run_group(gz, world, target, ev)
    for r in res:
        got = r["visited"]
        assert len(got) == len(want_visited)
        for g, w in zip(got, want_visited):
            assert g[0] == w[0] and abs(g[1] - w[1]) < 1e-9 and g[2] == w[2] and g[3] == w[3]
        assert r["best_q"] == want_best[0] and r["best_ok"] == want_best[1]
    # every rank made the same decisions in the same number of rounds
    assert len({r["rounds"] for r in res}) == 1
    assert sum(r["evaluated_here"] for r in res) == res[0]["evaluated_total"]
    # the reference may ask for the all-ones matrix twice (the generator does not know the first
    # TryQuantMatrix): the deterministic result is reused, the visit is not skipped
    distinct = len({(v[0], v[1]) for v in want_visited})
    if world == 1:
        assert res[0]["rounds"] == distinct == res[0]["evaluated_total"]
    else:
        assert res[0]["rounds"] < distinct                   # fewer serial rounds than trials
        assert res[0]["evaluated_total"] >= distinct


def test_eight_ranks_need_about_a_third_of_the_rounds(gz):
    ev = make_eval("typical")
    want_visited, _ = reference_search(ev, 0.971769)
    r8 = run_group(gz, 8, 0.971769, ev)[0]
    r2 = run_group(gz, 2, 0.971769, ev)[0]
    assert r8["rounds"] <= math.ceil(len(want_visited) / 2.5)
    assert r8["rounds"] <= r2["rounds"] < len(want_visited)


@pytest.mark.parametrize("kind", ["typical", "never_ok", "always_ok", "early_exit", "wobbly"])
@pytest.mark.parametrize("world,batch", [(1, 2), (1, 4), (2, 3), (4, 4)])
def test_batched_trials_per_rank_replay_the_reference_sequence(gz, kind, world, batch):
    """Several trials per rank and round (concurrent host legs on one GPU) change nothing either."""
    target = 0.971769
    ev = make_eval(kind)
    want_visited, want_best = reference_search(ev, target)
    res = run_group(gz, world, target, ev, batch)
    for r in res:
        assert len(r["visited"]) == len(want_visited)
        for g, w in zip(r["visited"], want_visited):
            assert g[0] == w[0] and abs(g[1] - w[1]) < 1e-9 and g[2] == w[2] and g[3] == w[3]
        assert r["best_q"] == want_best[0] and r["best_ok"] == want_best[1]
    distinct = len({(v[0], v[1]) for v in want_visited})
    assert res[0]["rounds"] < distinct
    assert sum(r["evaluated_here"] for r in res) == res[0]["evaluated_total"]


@pytest.mark.parametrize("kind", ["typical", "never_ok", "always_ok", "early_exit", "wobbly"])
@pytest.mark.parametrize("world", [1, 2, 4])
def test_downsampled_search_replays_the_reference_sequence(gz, kind, world):
    """The YUV420 pass: SelectQuantMatrix(jpg, downsample=true) -- the generator starts at heuristic
    score 0, so the all-ones matrix is asked for twice and then the search expands from there."""
    target = 0.971769
    ev = make_eval(kind)
    want_visited, want_best = reference_search(ev, target, downsample=True)
    assert want_visited[0][:2] == (0.0, 0.0) and want_visited[1][:2] == (0.0, 0.0)
    for r in run_group(gz, world, target, ev, mode=2):
        got = r["visited"]
        assert len(got) == len(want_visited)
        for g, w in zip(got, want_visited):
            assert g[0] == w[0] and abs(g[1] - w[1]) < 1e-9 and g[2] == w[2] and g[3] == w[3]
        assert r["best_q"] == want_best[0] and r["best_ok"] == want_best[1]


@pytest.mark.parametrize("world", [2, 4])
def test_a_failing_rank_fails_the_whole_group_without_a_hang(gz, world):
    """A rank whose trial fails (CUDA error, out of memory) must still enter the exchange: every rank then
    returns an error in the same round instead of the others blocking inside the all-gather."""
    target = 0.971769
    ok_ev = make_eval("typical")
    calls = [0] * world
    ag = BarrierAllGather(world)
    res = [None] * world

    def work(r):
        def ev(original, q):
            calls[r] += 1
            if r == world - 1 and calls[r] >= 2:
                return None          # second trial of the last rank fails
            return ok_ev(original, q)
        try:
            gz.QuantSearchSimulate(r, world, ag.for_rank(r), target, ev)
            res[r] = "ok"
        except Exception as ex:   # noqa
            res[r] = "error"
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(60)
    assert not any(t.is_alive() for t in th), "a rank is stuck in the exchange"
    assert res == ["error"] * world
