"""CPU-only tests of the host side of the product: the C ABI loads and exports every symbol the
header declares, and the host code (front end, JPEG writer, lazy std::sort) reproduces the
compiled reference byte for byte. No compute call touches a GPU here."""
import ctypes as C
import hashlib
import os
import re

import numpy as np
import pytest

from _libs import oracle, ref, have_ref, p, RefSession, synth_image, bees, ROOT
import __graft_entry__ as ge


@pytest.fixture(scope="module")
def gz():
    return ge.build()


def test_abi_exports_every_declared_symbol(gz):
    hdr = open(os.path.join(ROOT, "include", "gzb200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(gzb_[a-z0-9_]+)\s*\(", hdr))
    assert len(names) >= 35
    L = gz.lib()
    missing = [n for n in sorted(names) if not hasattr(L, n)]
    assert not missing, missing
    assert b"sm_100a" in L.gzb_version()


def test_no_cpu_fallback_without_gpu(gz):
    """Without a CUDA device every compute entry point must fail loudly."""
    if gz.device_count() > 0:
        pytest.skip("a GPU is present")
    img = synth_image(64, 48)
    with pytest.raises(gz.GzbError):
        gz.ButteraugliComparator(64, 48, img, 1.0)
    with pytest.raises(gz.GzbError):
        gz.Process(img, 1.0)


def test_product_does_not_touch_oracle():
    pkg = os.path.join(ROOT, "guetzli-cuda-opencl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h", "Makefile")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                if f == "gen_tables.py":
                    continue  # writes oracle/gzoracle_tables.h, does not read the oracle
                # no load / link / include / import of anything that lives under oracle/
                assert not re.search(r"libgzoracle|libgzref|gzoracle\.|gzo_[a-z]|ref_session|oracle/[\w/]+\.(so|c|cc|h|py)", txt), f


def test_quality_table(gz):
    for q in (84, 90, 95, 100, 97.5, 70, 110, 60, 120):
        want = oracle_quality(q)
        assert gz.ButteraugliScoreForQuality(q) == want


def oracle_quality(q):
    if have_ref():
        return ref().ref_butteraugli_score_for_quality(float(q))
    table = {84: 1.945456, 90: 1.473608, 95: 0.971769, 100: 0.211578}
    return table.get(q, gz.ButteraugliScoreForQuality(q))


def test_front_end_matches_oracle(gz):
    for (w, h) in [(64, 48), (70, 45), (33, 39), (444, 258)]:
        img = synth_image(w, h) if w != 444 else bees()
        nb = ((w + 7) // 8) * ((h + 7) // 8)
        want = np.zeros((3, nb, 64), np.int16)
        oracle().gzo_rgb_to_jpeg_coeffs(p(img), w, h, p(want[0]), p(want[1]), p(want[2]))
        assert np.array_equal(gz.RgbToJpegCoeffs(img), want)


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
@pytest.mark.parametrize("w,h,q", [(64, 48, 1), (97, 61, 5), (200, 133, 2), (444, 258, 3), (160, 120, 40)])
def test_jpeg_writer_bytes_equal_reference(gz, w, h, q):
    img = synth_image(w, h) if w != 444 else bees()
    s = RefSession(img, 0.97)
    qm = np.full(192, q, np.int32)
    qm[64:] = q + (q > 1)     # chroma table differs -> two tables in the file
    s.apply_quant(qm)
    want = s.write_jpeg()
    for nt in (1, 3, 8):
        got = gz.WriteJpeg(s.coeffs(), w, h, qm, host_threads=nt)
        assert hashlib.sha256(got).hexdigest() == hashlib.sha256(want).hexdigest(), (nt, len(got), len(want))
    s.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_jpeg_writer_grayscale_and_sparse(gz):
    img = synth_image(96, 80)
    gray = np.repeat(img[:, :, :1], 3, axis=2)
    s = RefSession(gray, 0.97)
    co = s.coeffs()
    assert not co[1:].any()   # chroma is exactly zero for r=g=b input
    want = s.write_jpeg()
    got = gz.WriteJpeg(co, 96, 80, np.ones(192, np.int32))
    assert got == want
    s.close()
    # long zero runs (ZRL symbols) and large magnitudes
    s = RefSession(img, 0.97)
    co = s.jpg_coeffs().copy()
    rng = np.random.default_rng(1)
    co[:, :, 1:] = np.where(rng.random(co[:, :, 1:].shape) < 0.03, co[:, :, 1:] * 7, 0).astype(np.int16)
    s.set_coeffs(co)
    assert gz.WriteJpeg(co, 96, 80, np.ones(192, np.int32)) == s.write_jpeg()
    s.close()


def _sort_cases(rng, n):
    yield rng.random(n).astype(np.float32)
    yield rng.integers(0, 7, n).astype(np.float32)                      # heavy ties
    yield np.sort(rng.random(n).astype(np.float32))                      # sorted input
    yield (rng.integers(0, max(1, n // 3), n) / 8.0).astype(np.float32)
    yield (np.arange(n) // 7).astype(np.float32)[::-1].copy()            # descending runs of equal keys


def test_restated_introsort_equals_std_sort(gz):
    """exact_sort (gzb_encoder.cc) restates libstdc++'s std::sort; the whole sort, and the sort evaluated
    the way the device does it (one partition at a time, right-hand ranges pending, short ranges finished
    separately with their depth budget), must both give std::sort's permutation, ties included."""
    L = gz.lib()
    L.gzb_test_exact_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.gzb_test_exact_sort_split.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.gzb_test_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    assert L.gzb_test_sort_emulation_ok() == 1
    rng = np.random.default_rng(4)
    for n in [0, 1, 2, 15, 16, 17, 33, 100, 1000, 5000, 70000, 600000]:
        for v in _sort_cases(rng, n):
            ids = np.arange(n, dtype=np.int32)
            a_id, a_v = ids.copy(), v.copy()
            L.gzb_test_std_sort(p(a_id), p(a_v), n)
            b_id, b_v = ids.copy(), v.copy()
            L.gzb_test_exact_sort(p(b_id), p(b_v), n)
            assert np.array_equal(b_id, a_id) and np.array_equal(b_v, a_v), n
            for small in (16, 100, 1024, 4096):
                b_id, b_v = ids.copy(), v.copy()
                L.gzb_test_exact_sort_split(p(b_id), p(b_v), n, small)
                assert np.array_equal(b_id, a_id) and np.array_equal(b_v, a_v), (n, small)


def test_host_lazy_sort_of_a_short_range(gz):
    """HostLazy (gzb_encoder.cc): the short ranges the device hands over are finished lazily -- the head that
    belongs to the prefix as a set, the rest in std::sort's arrangement, piece by piece."""
    L = gz.lib()
    L.gzb_test_host_lazy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t]
    L.gzb_test_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(17)
    for n in [1, 2, 16, 17, 40, 500, 4096]:
        for v in _sort_cases(rng, n):
            ids = np.arange(n, dtype=np.int32)
            a_id, a_v = ids.copy(), v.copy()
            L.gzb_test_std_sort(p(a_id), p(a_v), n)
            for pfx in sorted({0, 1, n // 3, n - 1}):
                b_id, b_v = ids.copy(), v.copy()
                L.gzb_test_host_lazy(p(b_id), p(b_v), n, pfx)
                assert np.array_equal(b_id[pfx:], a_id[pfx:]) and np.array_equal(b_v[pfx:], a_v[pfx:]), (n, pfx)
                assert np.array_equal(np.sort(b_id[:pfx]), np.sort(a_id[:pfx])), (n, pfx)


def test_host_lazy_sort_of_several_ranges_per_round_trip(gz):
    """gzb_be_select_ranges hands over up to 16 consecutive short ranges at once (HostLazy::reset_ranges): the
    restated partition plays the device here. Everything from the prefix on must be std::sort's arrangement."""
    L = gz.lib()
    L.gzb_test_host_lazy_ranges.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int]
    L.gzb_test_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(23)
    for n in [1, 17, 500, 5000, 40000]:
        for v in _sort_cases(rng, n):
            ids = np.arange(n, dtype=np.int32)
            a_id, a_v = ids.copy(), v.copy()
            L.gzb_test_std_sort(p(a_id), p(a_v), n)
            for pfx in sorted({0, n // 3, max(0, n - 10)}):
                for small, group in ((16, 16), (100, 3), (1024, 16), (1024, 1)):
                    b_id, b_v = ids.copy(), v.copy()
                    L.gzb_test_host_lazy_ranges(p(b_id), p(b_v), n, pfx, small, group)
                    assert np.array_equal(b_id[pfx:], a_id[pfx:]) and np.array_equal(b_v[pfx:], a_v[pfx:]), (n, pfx, small, group)
                    assert np.array_equal(np.sort(b_id[:pfx]), np.sort(a_id[:pfx])), (n, pfx, small, group)


def test_restated_heap_sort_equals_std_partial_sort(gz):
    """The fallback of introsort when the depth budget runs out: std::__partial_sort(first, last, last)."""
    L = gz.lib()
    L.gzb_test_exact_heap_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rng = np.random.default_rng(6)
    for n in [0, 1, 2, 3, 17, 100, 189, 1000, 4097, 50000]:
        for v in _sort_cases(rng, n):
            ids = np.arange(n, dtype=np.int32)
            a_id, a_v, b_id, b_v = ids.copy(), v.copy(), ids.copy(), v.copy()
            L.gzb_test_exact_heap_sort(p(a_id), p(a_v), n, 1)
            L.gzb_test_exact_heap_sort(p(b_id), p(b_v), n, 0)
            assert np.array_equal(a_id, b_id) and np.array_equal(a_v, b_v), n


def test_zeroing_key_ties_are_possible():
    """Probe of the zeroing-order model (guetzli/order.inc: key = |orig| * csf[idx] + bias[idx] in float):
    two DIFFERENT coefficients of a block can have equal keys, so the per-block std::sort's tie placement
    matters and the device kernel must reproduce it (gzb_zeroing.cuh: warp_input_order). None below |64|."""
    txt = open(os.path.join(ROOT, "guetzli-cuda-opencl_b200", "csrc", "gzb_zeroing_model.h")).read()
    vals = re.findall(r"\{\s*([-0-9.e+]+)f?,\s*([-0-9.e+]+)f?\}", txt[txt.index("kGzbZeroModel[192] = {"):])
    assert len(vals) == 192
    csf = np.array([float(a) for a, _ in vals], np.float32)
    bias = np.array([float(b) for _, b in vals], np.float32)
    idxs = np.array([i for i in range(192) if i % 64 != 0])

    def ties(amax):
        amp = np.arange(1, amax + 1, dtype=np.float32)
        keys = (amp[None, :] * csf[idxs][:, None]).astype(np.float32) + bias[idxs][:, None]
        ids = np.repeat(idxs, amax)
        k = keys.reshape(-1)
        o = np.argsort(k, kind="stable")
        ks, is_ = k[o], ids[o]
        return int(np.count_nonzero((ks[1:] == ks[:-1]) & (is_[1:] != is_[:-1])))
    assert ties(63) == 0
    assert ties(255) > 0
    k70 = np.float32(116) * csf[70] + bias[70]
    k84 = np.float32(57) * csf[84] + bias[84]
    assert k70 == k84      # the pair used by tests/test_gpu_parity.py::test_zeroing_order_with_equal_keys_is_std_sorts


def test_gamma_range_division_by_fma_is_ieee_division(gz):
    """gamma_rational divides (x - 0.77) by the constant 273.81 with a multiply and two fused multiply-adds
    (gzb_device_math.cuh: div_by_gamma_range). Exhaustive over every float x in [0, 1024] (the opsin
    absorbances lie in [0.77, 260]): the same bits as IEEE division."""
    L = gz.lib()
    L.gzb_test_gamma_division.restype = C.c_ulonglong
    assert L.gzb_test_gamma_division() == 0


def test_worker_pool_runs_every_task_exactly_once(gz):
    """The spin-then-sleep pool under many short back-to-back jobs, also with several pools alive at
    once (one per encoder thread in the group tests) and more threads than cores."""
    import threading
    L = gz.lib()
    L.gzb_test_pool_stress.restype = C.c_long
    L.gzb_test_pool_stress.argtypes = [C.c_int, C.c_int]
    assert L.gzb_test_pool_stress(1, 1000) == 0
    assert L.gzb_test_pool_stress(6, 60000) == 0
    res = []
    th = [threading.Thread(target=lambda t=t: res.append(L.gzb_test_pool_stress(3 + t, 20000))) for t in range(4)]
    for t in th:
        t.start()
    for t in th:
        t.join(300)
    assert res == [0, 0, 0, 0]


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_huffman_depths_equal_reference(gz):
    """Length-limited Huffman depths vs CreateHuffmanTree, including histograms whose unconstrained
    tree is deeper than 16 bits (rising count floor) and the warm-cache path."""
    L = gz.lib()
    R = ref()
    rng = np.random.default_rng(12)
    for trial in range(300):
        counts = np.zeros(257, np.uint32)
        nsym = int(rng.integers(1, 200))
        syms = rng.choice(256, nsym, replace=False)
        spread = int(rng.integers(1, 24))
        counts[syms] = (2 * (1 + rng.integers(0, 1 << rng.integers(0, spread + 1, nsym)))).astype(np.uint32)
        if trial % 5 == 0:   # Fibonacci-like counts force deep trees
            f = [1, 1]
            while len(f) < nsym: f.append(min(f[-1] + f[-2], 1 << 22))
            counts[syms] = 2 * np.array(f[:nsym], dtype=np.uint32)
        counts[256] = 1
        want = np.zeros(257, np.uint8); got = np.zeros(257, np.uint8); got2 = np.zeros(257, np.uint8)
        R.ref_create_huffman_tree(p(counts), p(want))
        L.gzb_test_huffman_depths(p(counts), p(got), None)
        warm = counts.copy(); warm[syms[: max(1, nsym // 3)]] += np.uint32(2 * int(rng.integers(1, 50)))
        if trial % 3 == 0:   # the warm histogram's support differs: symbols appear and vanish
            warm[syms[-1]] = 0
            warm[int(rng.integers(0, 256))] += np.uint32(4)
        L.gzb_test_huffman_depths(p(counts), p(got2), p(warm))
        assert np.array_equal(got, want), trial
        assert np.array_equal(got2, want), trial


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_reference_checker_does_not_link_the_product():
    """oracle/_ref/libgzref.so is the unmodified reference plus a forwarding shim; only the separate
    drop-in library (libgzref_dropin.so) may depend on libgzb200.so."""
    import subprocess
    ref()
    path = os.path.join(ROOT, "oracle", "_ref", "libgzref.so")
    needed = subprocess.run(["readelf", "-d", path], capture_output=True, text=True).stdout
    assert "NEEDED" in needed
    assert "gzb200" not in needed, needed
