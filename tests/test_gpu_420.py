"""GPU parity tests of the YUV 4:2:0 branch against the compiled reference
(oracle/_ref/libgzref.so): Downsample + SaveToJpegData, the factor-2 candidate image (fancy
upsampling), Compare, the zeroing search over luma blocks (comp_mask 1) and 16x16 macro-blocks
(comp_mask 6), the block weights with factor 2, the 4:2:0 JPEG writer and the whole encoder with
Params::force_420 / try_420. Bit-exact (tolerance 0) everywhere."""
import hashlib
import json
import os

import numpy as np
import pytest

from _libs import (oracle, have_ref, p, RefSession420, synth_image, bees, ref_process_params, image_420)
import __graft_entry__ as ge

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_ref(), reason="compiled reference missing")]

SIZES = [("red", 96, 80), ("red", 100, 75), ("synth", 33, 47), ("red", 129, 66), ("bees", 0, 0)]


@pytest.fixture(scope="module")
def gz():
    mod = ge.load_package()
    assert mod.device_count() > 0, "no CUDA device: the product has no CPU fallback"
    return mod


def jpeg_coeffs(img):
    h, w = img.shape[:2]
    nb = ((w + 7) // 8) * ((h + 7) // 8)
    c = np.zeros((3, nb, 64), np.int16)
    oracle().gzo_rgb_to_jpeg_coeffs(p(img), w, h, p(c[0]), p(c[1]), p(c[2]))
    return c


def report(name, a, b):
    a = np.asarray(a).reshape(-1); b = np.asarray(b).reshape(-1)
    assert a.shape == b.shape, (name, a.shape, b.shape)
    bad = np.flatnonzero(a != b)
    if len(bad):
        i = bad[0]
        pytest.fail("%s: %d/%d differ; first at %d: %r vs %r" % (name, len(bad), a.size, i, a[i], b[i]))


def make(gz, kind, w, h, target=0.97):
    img = image_420(kind, w, h)
    h, w = img.shape[:2]
    cmp_ = gz.ButteraugliComparator(w, h, img, target)
    cmp_.SetJpegCoeffs(jpeg_coeffs(img))
    cmp_.Downsample420()
    rs = RefSession420(img, target)
    return img, cmp_, rs


def quant_matrix(seed):
    rng = np.random.default_rng(seed)
    q = rng.integers(1, 9, (3, 64)).astype(np.int32)
    q[:, 0] = rng.integers(1, 4, 3)
    return q


@pytest.mark.parametrize("kind,w,h", SIZES)
def test_downsample_coeffs(gz, kind, w, h):
    img, cmp_, rs = make(gz, kind, w, h)
    for c in range(3):
        assert cmp_.ComponentDims(c)[:2] == rs.jpg_dims[c]
        assert cmp_.ComponentDims(c)[3] == rs.factor[c]
    got, want = cmp_.GetJpegCoeffs(), rs.jpg_coeffs()
    for c in range(3):
        report("downsampled jpg coeffs c%d" % c, got[c], want[c])
    cmp_.close(); rs.close()


@pytest.mark.parametrize("kind,w,h", SIZES)
def test_candidate_and_compare(gz, kind, w, h):
    img, cmp_, rs = make(gz, kind, w, h)
    q = quant_matrix(7)
    cmp_.CopyFromJpegData()
    rs.reset()
    report("srgb q=1", cmp_.ToSRGB(), rs.to_srgb())
    cmp_.ApplyGlobalQuantization(q)
    rs.apply_quant(q)
    got, want = cmp_.GetCoeffs(), rs.to_padded(rs.coeffs())
    for c in range(3):
        report("quantised coeffs c%d" % c, got[c], want[c])
    report("srgb", cmp_.ToSRGB(), rs.to_srgb())
    d = cmp_.Compare()
    d_ref, dm_ref = rs.compare()
    report("diffmap", cmp_.distmap(), dm_ref)
    assert float(d) == float(d_ref)
    # sparse update == SetCoeffBlock on the touched blocks
    rng = np.random.default_rng(11)
    coeffs = rs.coeffs()
    blocks, idxs, vals = [], [], []
    for c in range(3):
        nb = coeffs[c].shape[0]
        (bw, bh), (jw, jh) = rs.img_dims[c], rs.jpg_dims[c]
        for b in rng.choice(nb, size=min(nb, 12), replace=False):
            k = int(rng.integers(1, 64))
            coeffs[c][b, k] = 0
            blocks.append((b // bw) * jw + b % bw); idxs.append(64 * c + k); vals.append(0)
    rs.set_coeffs(coeffs)
    cmp_.UpdateCoeffs(blocks, idxs, vals)
    report("srgb after update", cmp_.ToSRGB(), rs.to_srgb())
    cmp_.close(); rs.close()


@pytest.mark.parametrize("kind,w,h", SIZES[:4])
@pytest.mark.parametrize("comp_mask", [1, 6])
def test_zeroing_order(gz, kind, w, h, comp_mask):
    # a generous block error limit so that whole zeroing orders (not only their cut-off) are compared
    img, cmp_, rs = make(gz, kind, w, h, target=3.0 if comp_mask == 6 else 1.5)
    q = quant_matrix(3)
    q[:] = np.minimum(q, 3)
    cmp_.CopyFromJpegData(); cmp_.ApplyGlobalQuantization(q)
    rs.reset(); rs.apply_quant(q)
    cmp_.StartBlockComparisons(); rs.start_block_comparisons()
    got = cmp_.ComputeBlockZeroingOrder(comp_mask)
    want = rs.zeroing_order_f(comp_mask)
    report("zeroing idx", got["idx"], want["idx"])
    report("zeroing err", got["err"], want["err"])
    assert (want["err"] > 0).sum() > 20
    off, ci, ce = cmp_.ComputeBlockZeroingCandidates(comp_mask)
    keep = want["err"] > 0
    report("candidate idx", ci, want["idx"][keep].astype(np.uint8))
    report("candidate err", ce, want["err"][keep])
    cmp_.close(); rs.close()


@pytest.mark.parametrize("direction,rblock", [(1, 1), (1, 3), (-1, 1), (-1, 4)])
def test_block_weights_factor2(gz, direction, rblock):
    img, cmp_, rs = make(gz, "red", 129, 66, target=0.6)
    q = np.full((3, 64), 5, np.int32)
    cmp_.CopyFromJpegData(); cmp_.ApplyGlobalQuantization(q)
    rs.reset(); rs.apply_quant(q)
    cmp_.Compare()
    d, dm = rs.compare()
    for factor in (1, 2):
        got = cmp_.ComputeBlockErrorAdjustmentWeights(direction, rblock, 1.0, factor=factor)
        want = rs.block_weights_f(direction, rblock, 1.0, factor, dm)
        report("weights f%d" % factor, got, want)
    cmp_.close(); rs.close()


@pytest.mark.parametrize("kind,w,h", SIZES)
def test_candidate_jpeg_bytes(gz, kind, w, h):
    """SaveToJpegData + WriteJpeg of a 4:2:0 candidate (2x2 MCUs, MCU-padded luma, SOF sampling 0x22):
    header on the host, scan Huffman-coded on the device -- the reference's bytes."""
    img, cmp_, rs = make(gz, kind, w, h)
    q = quant_matrix(5)
    cmp_.CopyFromJpegData(); cmp_.ApplyGlobalQuantization(q)
    rs.reset(); rs.apply_quant(q)
    want = rs.write_jpeg()
    n, got = cmp_.WriteJpeg(q)
    assert n == len(want)
    assert got == want
    cmp_.close(); rs.close()


def trace_records(lines):
    out = []
    for l in lines:
        out.append((l.split("Out[")[1].split("]")[0].strip(), l.split("D[")[1].split("]")[0].strip()))
    return out


GOLD_420 = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "yuv420_encodes.json")


def _gold_420():
    return json.load(open(GOLD_420)) if os.path.exists(GOLD_420) else {}


@pytest.mark.parametrize("name", sorted(_gold_420().keys()))
def test_encode_420_golden(gz, name):
    """guetzli::Process with try_420 / force_420: bytes and iteration trace of the reference
    (tests/golden/yuv420_encodes.json, generated by tests/golden/make_golden.py --420)."""
    gold = _gold_420()[name]
    kind, size, q, mode = name.split("_")
    w, h = (int(v) for v in size.split("x"))
    img = image_420(kind, w, h)
    jpg, st, trace = gz.Process(img, np.float32(gold["target"]), want_trace=True,
                                try_420=mode == "try", force_420=mode == "force")
    got = trace_records([l for l in trace.splitlines() if "Out[" in l])
    want = trace_records(gold["trace"])
    first_bad = next((i for i, (a, b) in enumerate(zip(got, want)) if a != b), None)
    assert first_bad is None and len(got) == len(want), \
        "trace diverges at record %s: got %s want %s" % (first_bad, got[first_bad:first_bad + 2] if first_bad is not None else len(got),
                                                         want[first_bad:first_bad + 2] if first_bad is not None else len(want))
    assert len(jpg) == gold["size"]
    assert hashlib.sha256(jpg).hexdigest() == gold["sha256"]
    assert st["num_iterations"] == gold["iterations"]


@pytest.mark.parametrize("name", ["red_96x80_q95_force", "synth_33x47_q84_force", "synth_96x80_q95_try"])
def test_unmodified_reference_processor_420_through_b200_comparator(gz, name):
    """Drop-in: the reference's own ProcessJpegData with Params::force_420 / try_420, running against the
    Comparator adaptor of integration/gzb_comparator.cc (Compare of a factor-2 OutputImage,
    SwitchBlock / CompareBlock with sampling factors through gzb_compare_block_srgb, block weights with
    factor 2), must emit the reference's bytes."""
    import ctypes as C
    import _libs
    L = _libs.ref_dropin()
    L.ref_process_rgb_b200_params.restype = C.c_long
    gold = _gold_420()[name]
    kind, size, q, mode = name.split("_")
    w, h = (int(v) for v in size.split("x"))
    img = image_420(kind, w, h)
    out = np.zeros(w * h * 3 + (1 << 16), np.uint8)
    iters = C.c_int()
    n = L.ref_process_rgb_b200_params(p(img), w, h, C.c_float(gold["target"]), int(mode == "try"), int(mode == "force"),
                                      0, p(out), C.c_long(out.size), C.byref(iters))
    assert n == gold["size"] and iters.value == gold["iterations"]
    assert hashlib.sha256(out[:n].tobytes()).hexdigest() == gold["sha256"]


@pytest.mark.parametrize("world", [2, 3])
@pytest.mark.parametrize("name", [n for n in ("bees_444x258_q90_force", "bees_444x258_q95_try", "synth_33x47_q84_force", "synth_96x80_q95_try",
                                               "gray_80x64_q90_force", "gray_80x64_q90_try", "red_129x66_q92_try") if n in _gold_420()])
def test_group_encode_420_equals_reference(gz, name, world):
    """One image on several GPUs with Params::try_420 / force_420: the quant searches of both passes and the
    zeroing searches of the 4:4:4 and the luma pass are sharded over the group (the chroma pass, which starts
    from rank 0's luma result, runs on rank 0); rank 0 must return the reference's bytes and trace."""
    from test_gpu_encode import run_thread_group
    gold = _gold_420()[name]
    kind, size, q, mode = name.split("_")
    w, h = (int(v) for v in size.split("x"))
    img = image_420(kind, w, h)
    res = run_thread_group(gz, img, np.float32(gold["target"]), world, try_420=mode == "try", force_420=mode == "force")
    jpg, st, trace = res[0]
    got = trace_records([l for l in trace.splitlines() if "Out[" in l])
    assert got == trace_records(gold["trace"])
    assert hashlib.sha256(jpg).hexdigest() == gold["sha256"]
    assert st["num_iterations"] == gold["iterations"]
    for r in range(1, world):
        assert res[r][0] == b""
