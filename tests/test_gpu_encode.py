"""End-to-end parity of the B200 encoder (gzb_encode_rgb == guetzli::Process) against the golden
fixtures generated from the unmodified reference (tests/golden/make_golden.py):
  * tests/bees.png at q95 must reproduce the reference's own golden checksum
    (tests/golden_checksums.txt:3) byte for byte;
  * seeded synthetic images must reproduce the reference's bytes and its iteration trace."""
import hashlib
import json
import os
import re

import ctypes as C
import numpy as np
import pytest

from _libs import synth_image, bees, ROOT
import __graft_entry__ as ge

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")


@pytest.fixture(scope="module")
def gz():
    mod = ge.load_package()
    assert mod.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mod


def trace_records(lines):
    """(out_bytes, distance) per iteration line of a GUETZLI_LOG trace."""
    out = []
    for l in lines:
        m = re.search(r"Out\[\s*(\d+)\].*D\[\s*([0-9.]+)\]", l)
        if m:
            out.append((int(m.group(1)), m.group(2)))
    return out


def check(gz, img, gold):
    jpg, st, trace = gz.Process(img, np.float32(gold["target"]), want_trace=True)
    got = trace_records([l for l in trace.splitlines() if "Out[" in l])
    want = trace_records(gold["trace"])
    first_bad = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), None)
    assert first_bad is None and len(got) == len(want), (
        "trace diverges at record %s: got %s want %s" % (first_bad, got[first_bad:first_bad + 2] if first_bad is not None else len(got),
                                                          want[first_bad:first_bad + 2] if first_bad is not None else len(want)))
    assert st["num_iterations"] == gold["iterations"]
    assert len(jpg) == gold["size"]
    assert hashlib.sha256(jpg).hexdigest() == gold["sha256"]
    return st


def test_bees_golden_checksum(gz):
    gold = json.load(open(os.path.join(GOLD, "bees_q95.json")))
    assert gold["sha256"] == "39cfc110d3d389ddf3b9c47adece290ef077b78467f1ed39679b691c6fb7c96d"
    st = check(gz, bees(), gold)
    assert st["launches"] > 100 and st["device_compare_ms"] > 0 and st["device_zeroing_ms"] > 0


@pytest.mark.parametrize("key", ["128x96_q90_s1234", "160x120_q95_s1244", "97x61_q84_s1254", "256x256_q90_s1234"])
def test_synthetic_encodes_equal_reference(gz, key):
    gold = json.load(open(os.path.join(GOLD, "synth_encodes.json")))[key]
    m = re.match(r"(\d+)x(\d+)_q(\d+)_s(\d+)", key)
    w, h, q, seed = map(int, m.groups())
    assert abs(gz.ButteraugliScoreForQuality(q) - gold["target"]) < 1e-12
    check(gz, synth_image(w, h, seed), gold)


def test_encode_is_deterministic_and_thread_count_independent(gz):
    img = synth_image(128, 96, 1234)
    t = np.float32(gz.ButteraugliScoreForQuality(90))
    a, _, _ = gz.Process(img, t, host_threads=1)
    b, _, _ = gz.Process(img, t, host_threads=7)
    assert a == b


def test_encode_rejects_bad_input(gz):
    img = synth_image(64, 48)
    with pytest.raises(gz.GzbError):
        gz.Process(img, 2.5)            # quality < 84 is refused (processor.cc:939-945)
    with pytest.raises(gz.GzbError):
        gz.Process(img[:16, :16], 1.0)  # < 32 px


def test_unmodified_reference_processor_drives_b200_comparator(gz):
    """Drop-in: the reference's own Processor::ProcessJpegData (compiled from /root/reference into
    oracle/_ref) running against integration/gzb_comparator.cc must emit the reference's bytes."""
    import ctypes as C
    import _libs
    if not _libs.have_ref():
        pytest.skip("oracle/_ref not present")
    L = _libs.ref_dropin()
    L.ref_process_rgb_b200.restype = C.c_long
    gold_all = json.load(open(os.path.join(GOLD, "synth_encodes.json")))
    for key in ["97x61_q84_s1254", "128x96_q90_s1234"]:
        gold = gold_all[key]
        m = re.match(r"(\d+)x(\d+)_q(\d+)_s(\d+)", key)
        w, h, q, seed = map(int, m.groups())
        img = synth_image(w, h, seed)
        out = np.zeros(w * h * 3 + (1 << 16), np.uint8)
        iters = C.c_int()
        n = L.ref_process_rgb_b200(_libs.p(img), w, h, C.c_float(gold["target"]), 0, _libs.p(out), C.c_long(out.size),
                                   C.byref(iters))
        assert n == gold["size"] and iters.value == gold["iterations"]
        assert hashlib.sha256(out[:n].tobytes()).hexdigest() == gold["sha256"]


def test_compare_block_single_equals_batched(gz):
    import ctypes as C
    img = synth_image(64, 48)
    c = gz.ButteraugliComparator(64, 48, img, 0.97)
    co = gz.RgbToJpegCoeffs(img)
    c.SetJpegCoeffs(co); c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
    cur = c.GetCoeffs()
    c.StartBlockComparisons()
    want = c.CompareBlocks()
    L = gz.lib()
    L.gzb_compare_block.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.POINTER(C.c_double)]
    for b in range(c.num_blocks):
        cand = np.ascontiguousarray(cur[:, b, :].reshape(192))
        e = C.c_double()
        assert L.gzb_compare_block(c._ctx, b % c.block_width, b // c.block_width, cand.ctypes.data, C.byref(e)) == 0
        assert np.float32(e.value) == want[b]
    c.close()


class BarrierDeviceAllGather:
    """gzb_allgather_device_fn for ranks that are threads on ONE GPU: every rank publishes its send pointer, then
    copies all of them into its own receive buffer (device to device)."""
    def __init__(self, world, lib):
        import threading
        self.world, self.slots, self.bar, self.lib, self.calls = world, [None] * world, threading.Barrier(world), lib, 0
        lib.gzb_test_memcpy_d2d.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]

    def for_rank(self, r):
        def fn(d_send, nbytes, d_recv):
            self.slots[r] = d_send
            self.bar.wait()
            for k in range(self.world):
                assert self.lib.gzb_test_memcpy_d2d(d_recv + k * nbytes, self.slots[k], nbytes) == 0
            if r == 0:
                self.calls += 1
            self.bar.wait()
        return fn


def run_thread_group(gz, img, target, world, device_exchange=None, **params):
    """`world` encoders of the same image as the ranks of a group, one host thread each, on the one
    GPU of the test box (the exchange is a thread barrier; across GPUs it is NCCL, bench.py)."""
    import threading
    from test_quant_search_cpu import BarrierAllGather
    ag = BarrierAllGather(world)
    res, err = [None] * world, []

    def work(r):
        try:
            enc = gz.Encoder(img, target, host_threads=2, **params)
            enc.set_group(r, world, ag.for_rank(r), device_exchange.for_rank(r) if device_exchange else None)
            res[r] = enc.run(want_trace=True)
            enc.close()
        except Exception as e:  # a dead rank would leave the others at the barrier
            err.append(e)
            ag.bar.abort()
    th = [threading.Thread(target=work, args=(r,)) for r in range(world)]
    for t in th:
        t.start()
    for t in th:
        t.join(600)
    assert not err, err
    return res


@pytest.mark.parametrize("world", [2, 3, 8])
def test_group_encode_equals_single_gpu_encode(gz, world):
    """SelectQuantMatrix candidates and zeroing blocks sharded over a group produce the
    same trace and the same bytes as the single-GPU encode (== the reference's)."""
    gold = json.load(open(os.path.join(GOLD, "synth_encodes.json")))["160x120_q95_s1244"]
    img = synth_image(160, 120, 1244)
    res = run_thread_group(gz, img, np.float32(gold["target"]), world)
    jpg, st, trace = res[0]
    assert hashlib.sha256(jpg).hexdigest() == gold["sha256"]
    got = trace_records([l for l in trace.splitlines() if "Out[" in l])
    assert got == trace_records(gold["trace"])
    assert st["num_iterations"] == gold["iterations"]
    for r in range(1, world):
        assert res[r][0] == b""                      # only rank 0 returns the file
        assert res[r][1]["search_rounds"] == st["search_rounds"]
    quant_trials = sum(1 for l in gold["trace"] if "GQ[" in l) + 1
    assert st["search_rounds"] < quant_trials        # fewer serial rounds than trials
    assert sum(r[1]["num_compares"] for r in res[1:]) > 0   # the other ranks did evaluate trials


@pytest.mark.parametrize("world", [2, 3])
def test_group_encode_with_device_side_candidate_exchange(gz, world):
    """gzb_encoder_set_group_device: the zeroing candidates are all-gathered on the device(s) and become the back
    end's lists without visiting the host. Same bytes; also for the YUV420 branch (luma pass sharded)."""
    gold = json.load(open(os.path.join(GOLD, "synth_encodes.json")))["160x120_q95_s1244"]
    img = synth_image(160, 120, 1244)
    dx = BarrierDeviceAllGather(world, gz.lib())
    res = run_thread_group(gz, img, np.float32(gold["target"]), world, device_exchange=dx)
    assert hashlib.sha256(res[0][0]).hexdigest() == gold["sha256"]
    assert dx.calls == 1
    for r in range(1, world):
        assert res[r][0] == b""
    single = gz.Process(img, np.float32(gold["target"]), try_420=True)[0]
    dx = BarrierDeviceAllGather(world, gz.lib())
    res = run_thread_group(gz, img, np.float32(gold["target"]), world, device_exchange=dx, try_420=True)
    assert res[0][0] == single
    assert dx.calls == 2   # the 4:4:4 pass and the luma pass of the 4:2:0 branch


def test_group_encode_bees_golden(gz):
    gold = json.load(open(os.path.join(GOLD, "bees_q95.json")))
    res = run_thread_group(gz, bees(), np.float32(gold["target"]), 4)
    assert hashlib.sha256(res[0][0]).hexdigest() == gold["sha256"]


def test_zeroing_candidates_by_block_range_equal_whole(gz):
    img = synth_image(96, 72)
    c = gz.ButteraugliComparator(96, 72, img, 0.97)
    c.SetJpegCoeffs(gz.RgbToJpegCoeffs(img)); c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 2, np.int32))
    c.StartBlockComparisons()
    off, idx, err = c.ComputeBlockZeroingCandidates(7)
    nb = c.num_blocks
    for (b0, b1) in [(0, nb // 3), (nb // 3, nb - 1), (nb - 1, nb), (5, 5)]:
        o, i, e = c.ComputeBlockZeroingCandidates(7, b0, b1)
        assert np.array_equal(o, off[b0:b1 + 1] - off[b0])
        assert np.array_equal(i, idx[off[b0]:off[b1]]) and np.array_equal(e, err[off[b0]:off[b1]])
    c.close()


@pytest.mark.parametrize("w,h,q,seed", [(64, 48, 1, 1234), (97, 61, 5, 1244), (200, 133, 2, 1254), (444, 258, 3, 0),
                                        (160, 120, 40, 1264), (256, 192, 200, 1274)])
def test_device_huffman_writer_bytes_equal_host_writer(gz, w, h, q, seed):
    """gzb_write_candidate_jpeg (scan coded on the GPU) == the host writer (which test_host_cpu pins
    byte for byte to the reference's WriteJpeg), including grey-only output when chroma quantises to
    zero (q=200) and the input-table variant of the q=1 original."""
    img = bees() if seed == 0 else synth_image(w, h, seed)
    c = gz.ButteraugliComparator(w, h, img, 0.97)
    co = gz.RgbToJpegCoeffs(img)
    c.SetJpegCoeffs(co)
    c.CopyFromJpegData()
    qm = np.full(192, q, np.int32)
    qm[64:] += (seed % 3)              # different chroma tables
    c.ApplyGlobalQuantization(qm)
    cur = c.GetCoeffs()
    want = gz.WriteJpeg(cur, w, h, qm, input_tables=False, host_threads=3)   # takes dequantised values
    size, got = c.WriteJpeg(qm)
    assert size == len(want) and got == want
    assert c.WriteJpeg(qm, want_bytes=False)[0] == len(want)
    if q == 1:
        want_in = gz.WriteJpeg(cur, w, h, qm, input_tables=True, host_threads=1)
        assert c.WriteJpeg(qm, input_tables=True)[1] == want_in
    c.close()


from _libs import SPECIAL_IMAGES, special_image


@pytest.mark.parametrize("name", SPECIAL_IMAGES)
def test_edge_case_encodes_equal_reference(gz, name):
    """Grey-only input (chroma dropped from the file), 32x32, sizes off the 8x8 grid, a flat image,
    white noise, hard edges: bytes and iteration trace of the reference (tests/golden/edge_encodes.json)."""
    gold = json.load(open(os.path.join(GOLD, "edge_encodes.json")))[name]
    img, q = special_image(name)
    assert abs(gz.ButteraugliScoreForQuality(q) - gold["target"]) < 1e-12
    check(gz, img, gold)


@pytest.mark.parametrize("name", SPECIAL_IMAGES)
@pytest.mark.parametrize("world", [2, 5])
def test_group_encode_edge_cases(gz, name, world):
    """The group path on the edge-case inputs. (The grey image once exposed that DistanceOK(1.0) of
    the first back-end iteration must see the last trial in the REFERENCE's order, which another
    rank may have evaluated.)"""
    gold = json.load(open(os.path.join(GOLD, "edge_encodes.json")))[name]
    img, _ = special_image(name)
    res = run_thread_group(gz, img, np.float32(gold["target"]), world)
    assert hashlib.sha256(res[0][0]).hexdigest() == gold["sha256"]
    got = trace_records([l for l in res[0][2].splitlines() if "Out[" in l])
    assert got == trace_records(gold["trace"])


@pytest.mark.parametrize("key", ["1024x1024_q90_s1234", "4000x3000_q95_s1234", "1920x1080_q95_s1234", "1920x1080_q95_s1244"])
def test_full_bench_workloads_equal_reference(gz, key):
    """The FULL workloads of BASELINE.json (the 1 MPix q90 image of configs[1], bench.py's 12 MPix q95 image,
    two images of the 64-image batch of configs[3]):
    bytes and iteration trace of the single-threaded CPU reference (tests/golden/full_encodes.json,
    two minutes and a quarter of an hour of CPU Guetzli, made by tests/golden/make_full_golden.py)."""
    path = os.path.join(GOLD, "full_encodes.json")
    gold_all = json.load(open(path)) if os.path.exists(path) else {}
    if key not in gold_all:
        pytest.skip("golden for %s not generated" % key)
    gold = gold_all[key]
    m = re.match(r"(\d+)x(\d+)_q(\d+)_s(\d+)", key)
    w, h, q, seed = map(int, m.groups())
    check(gz, synth_image(w, h, seed), gold)


def test_batch_with_encodes_in_flight_equals_single_encodes(gz):
    """gzb_encode_rgb_batch: several encodes running concurrently on one GPU (separate contexts, host
    threads split between them) give exactly the bytes of one-at-a-time encodes."""
    imgs = [synth_image(160, 120, 1244), synth_image(128, 96, 1234), synth_image(97, 61, 1254),
            synth_image(200, 136, 31), synth_image(64, 64, 5), synth_image(256, 256, 1234)]
    t = np.float32(gz.ButteraugliScoreForQuality(92))
    single = [gz.Process(im, t)[0] for im in imgs]
    for inflight in (1, 3, 6):
        got = gz.ProcessBatch(imgs, t, inflight=inflight)
        assert [g[0] for g in got] == single
    got = gz.ProcessBatch(imgs[:3], t, inflight=2, try_420=True)
    assert [g[0] for g in got] == [gz.Process(im, t, try_420=True)[0] for im in imgs[:3]]


def test_std_sort_fallback_gives_the_same_bytes():
    """If the restated introsort ever disagreed with the std::sort of the build (another standard library), the
    back end fetches the order and sorts it with std::sort itself. GZB_NO_SORT_EMULATION=1 forces that path (the
    switch is read once per process, hence the subprocess): same golden bytes on bees.png."""
    import subprocess, sys
    gold = json.load(open(os.path.join(GOLD, "bees_q95.json")))
    code = ("import sys, hashlib, numpy as np; sys.path.insert(0, %r); sys.path.insert(0, %r); "
            "import __graft_entry__ as ge; from _libs import bees; gz = ge.load_package(); "
            "jpg, st, _ = gz.Process(bees(), np.float32(%r)); print(hashlib.sha256(jpg).hexdigest(), st['be_selects'])"
            % (ROOT, os.path.join(ROOT, "tests"), gold["target"]))
    env = dict(os.environ, GZB_NO_SORT_EMULATION="1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    sha, selects = out.stdout.split()[-2:]
    assert sha == gold["sha256"]
    assert int(selects) == 0      # the device's lazy sort was not used
