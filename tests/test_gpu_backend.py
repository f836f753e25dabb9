"""The device side of SelectFrequencyBackEnd (csrc/gzb_backend.cuh) on its own.

The partition kernels (k_be_tiles_*, k_be_swap, k_be_local) must move the entries of the order exactly like libstdc++'s std::sort does (median-of-three,
unguarded Hoare partition): the reference sorts `global_order` with it (guetzli/processor.cc:825-828) and
ties between different blocks are common, so which of the tied entries ends up inside the consumed part
decides the output bytes. Checker: std::sort itself (gzb_test_std_sort), bit-exact."""
import ctypes as C

import numpy as np
import pytest

from _libs import synth_image, p
import __graft_entry__ as ge

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gz():
    mod = ge.load_package()
    assert mod.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mod


@pytest.fixture(scope="module")
def ctx(gz):
    c = gz.ButteraugliComparator(64, 64, synth_image(64, 64), 1.0)
    yield c
    c.close()


def _cases(rng, n):
    yield "uniform", rng.random(n).astype(np.float32)
    yield "ties7", rng.integers(0, 7, n).astype(np.float32)
    yield "sorted", np.sort(rng.random(n).astype(np.float32))
    yield "ties_n3", (rng.integers(0, max(1, n // 3), n) / 8.0).astype(np.float32)
    yield "runs_desc", (np.arange(n) // 7).astype(np.float32)[::-1].copy()
    yield "const", np.full(n, 0.5, np.float32)


@pytest.mark.parametrize("n", [1, 2, 17, 1000, 5000, 70000, 600000, 3000000])
def test_device_lazy_sort_equals_std_sort(gz, ctx, n):
    L = gz.lib()
    L.gzb_test_device_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int]
    L.gzb_test_std_sort.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    rng = np.random.default_rng(100 + n % 97)
    entry = np.dtype([("block", np.int32), ("value", np.float32)])
    for name, v in _cases(rng, n):
        ids = np.arange(n, dtype=np.int32)
        a_id, a_v = ids.copy(), v.copy()
        L.gzb_test_std_sort(p(a_id), p(a_v), n)
        prefixes = sorted({0, n // 50, n // 3, max(0, n - 10)})
        # (a negative length asks for several short ranges per round trip: gzb_be_select_ranges)
        smalls = (16, 1024, 4096, -64, -1024) if n <= 70000 else (1024, -512)
        for small in smalls:
            for pfx in prefixes:
                e = np.zeros(n, entry)
                e["block"], e["value"] = ids, v
                rc = L.gzb_test_device_sort(ctx._ctx, p(e), n, pfx, small)
                assert rc == 0, (name, n, pfx, small, rc)
                assert np.array_equal(e["block"][pfx:], a_id[pfx:]), (name, n, pfx, small)
                assert np.array_equal(e["value"][pfx:], a_v[pfx:]), (name, n, pfx, small)
                assert np.array_equal(np.sort(e["block"][:pfx]), np.sort(a_id[:pfx])), (name, n, pfx, small)


@pytest.mark.parametrize("n,depth", [(5000, 0), (5000, 2), (70000, 3), (70000, 6), (600000, 4)])
def test_device_lazy_sort_with_exhausted_depth_budget(gz, ctx, n, depth):
    """introsort heap-sorts a range whose depth budget is used up. The real budget (2 log2 n) is almost never
    exhausted, so the sort is started with a tiny one: the device must stop partitioning such a range (status 2),
    the host heap-sorts it, and the result must equal the restated introsort run whole with the same budget
    (whose heap path is checked against std::partial_sort in tests/test_host_cpu.py)."""
    L = gz.lib()
    L.gzb_test_device_sort_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_int]
    L.gzb_test_exact_sort_depth.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    rng = np.random.default_rng(7 + n + depth)
    entry = np.dtype([("block", np.int32), ("value", np.float32)])
    for name, v in _cases(rng, n):
        ids = np.arange(n, dtype=np.int32)
        a_id, a_v = ids.copy(), v.copy()
        L.gzb_test_exact_sort_depth(p(a_id), p(a_v), n, depth)
        for pfx, small in ((0, 1024), (n // 3, 1024), (n // 3, -256)):
            e = np.zeros(n, entry)
            e["block"], e["value"] = ids, v
            assert L.gzb_test_device_sort_depth(ctx._ctx, p(e), n, pfx, small, depth) == 0
            assert np.array_equal(e["block"][pfx:], a_id[pfx:]), (name, n, depth, pfx)
            assert np.array_equal(e["value"][pfx:], a_v[pfx:]), (name, n, depth, pfx)
            assert np.array_equal(np.sort(e["block"][:pfx]), np.sort(a_id[:pfx])), (name, n, depth, pfx)


def test_input_is_gray(gz):
    L = gz.lib()
    img = synth_image(96, 64)
    g = img.reshape(-1, 3).copy()
    g[:, 1] = g[:, 0]
    g[:, 2] = g[:, 0]
    for im, want in ((img, 0), (g.reshape(img.shape), 1)):
        c = gz.ButteraugliComparator(96, 64, im, 1.0)
        assert L.gzb_rgb_to_jpeg_coeffs_device(c._ctx) == 0
        flag = C.c_int(-1)
        assert L.gzb_input_is_gray(c._ctx, C.byref(flag)) == 0
        assert flag.value == want
        c.close()
