"""Pins the plain-C oracle (oracle/gzoracle.c) to the unmodified reference (oracle/_ref/libgzref.so),
stage by stage and BIT-EXACTLY, on seeded inputs -- the method of the reference's own --checkcl mode
(clguetzli/clguetzli_test.cpp:18-36) with the tolerance tightened to zero. CPU only."""
import ctypes as C
import numpy as np
import pytest

from _libs import oracle, ref, have_ref, p, RefSession, synth_image, bees, COEFF_DATA

pytestmark = pytest.mark.skipif(not have_ref(), reason="oracle/_ref not built and no /root/reference")

SIZES = [(64, 48), (97, 61), (200, 133)]


def rand_planes(w, h, seed, lo=0.0, hi=255.0):
    rng = np.random.default_rng(seed)
    return rng.uniform(lo, hi, (3, h, w)).astype(np.float32)


def xyb_pair(w, h, seed=7):
    """Two related XYB images: a synthetic photo and a perturbed copy, through the reference."""
    img = synth_image(w, h, seed)
    rng = np.random.default_rng(seed)
    img2 = np.clip(img.astype(np.int32) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    a = np.zeros((3, h, w), np.float32)
    b = np.zeros((3, h, w), np.float32)
    ref().ref_compute_opsin_dynamics_image(p(img), w, h, p(a))
    ref().ref_compute_opsin_dynamics_image(p(img2), w, h, p(b))
    return img, img2, a, b


def test_srgb_lut_and_scalars():
    a = np.zeros(256); b = np.zeros(256)
    oracle().gzo_srgb8_to_linear_table(p(a)); ref().ref_srgb8_to_linear_table(p(b))
    assert np.array_equal(a, b)
    for d, s, t in [(0.5, 1000, 0.97), (0.98, 1234, 0.971769), (1.5, 999, 0.97), (1.0, 5, 1.0)]:
        assert oracle().gzo_score_jpeg(d, s, t) == ref().ref_score_jpeg(d, s, t)


@pytest.mark.parametrize("sigma,br", [(1.1, 0.0), (1.5, 0.0), (0.586, 0.0), (0.4, 0.0), (14.0, 0.0),
                                      (9.65781083553, 0.0), (14.2644604355, 0.0),
                                      (4.53358927369, 0.0), (8.8510880283, 0.03027655136)])
@pytest.mark.parametrize("w,h", [(8, 8), (67, 45), (130, 40)])
def test_blur(sigma, br, w, h):
    x = rand_planes(w, h, 3)[0].copy()
    a = x.copy(); b = x.copy()
    oracle().gzo_blur(a, w, h, sigma, br); ref().ref_blur(b, w, h, sigma, br)
    assert np.array_equal(a, b)


@pytest.mark.parametrize("w,h", [(8, 8)] + SIZES)
def test_opsin_dynamics(w, h):
    x = rand_planes(w, h, 5)
    a = x.copy(); b = x.copy()
    oracle().gzo_opsin_dynamics_image(p(a), w, h); ref().ref_opsin_dynamics_image(p(b), w, h)
    assert np.array_equal(a, b)
    img = synth_image(w, h)
    a = np.zeros((3, h, w), np.float32); b = a.copy()
    oracle().gzo_srgb_to_xyb(p(img), w, h, p(a)); ref().ref_compute_opsin_dynamics_image(p(img), w, h, p(b))
    assert np.array_equal(a, b)


@pytest.mark.parametrize("w,h", [(8, 8)] + SIZES)
def test_mask_high_intensity_change(w, h):
    _, _, x0, x1 = xyb_pair(w, h)
    oa = np.zeros_like(x0); ob = oa.copy(); ra = oa.copy(); rb = oa.copy()
    oracle().gzo_mask_high_intensity_change(p(x0), p(x1), w, h, p(oa), p(ob))
    ref().ref_mask_high_intensity_change(p(x0), p(x1), w, h, p(ra), p(rb))
    assert np.array_equal(oa, ra) and np.array_equal(ob, rb)


def test_block_diff():
    rng = np.random.default_rng(11)
    for trial in range(200):
        b0 = rng.normal(0, 30, 192)
        b1 = b0 + rng.normal(0, [0.01, 0.5, 5.0][trial % 3], 192)
        if trial == 7:
            b1 = b0.copy()
        od = [np.zeros(3) for _ in range(3)]; rd = [np.zeros(3) for _ in range(3)]
        oracle().gzo_block_diff(p(b0), p(b1), p(od[0]), p(od[1]), p(od[2]))
        ref().ref_block_diff(p(b0), p(b1), p(rd[0]), p(rd[1]), p(rd[2]))
        for a, b in zip(od, rd):
            assert np.array_equal(a, b), trial


@pytest.mark.parametrize("w,h", SIZES)
def test_mask_chain(w, h):
    _, _, x0, x1 = xyb_pair(w, h)
    a = np.zeros_like(x0); b = a.copy()
    oracle().gzo_diff_precompute(p(x0), p(x1), w, h, p(a)); ref().ref_diff_precompute(p(x0), p(x1), w, h, p(b))
    assert np.array_equal(a, b)
    pa = a[0].copy(); pb = a[0].copy()
    oracle().gzo_average5x5(p(pa), w, h); ref().ref_average5x5(p(pb), w, h)
    assert np.array_equal(pa, pb)
    oracle().gzo_min_square_val(p(pa), w, h, 4, 0); ref().ref_min_square_val(p(pb), w, h, 4, 0)
    assert np.array_equal(pa, pb)
    m = np.zeros_like(x0); mdc = m.copy(); rm = m.copy(); rmdc = m.copy()
    oracle().gzo_mask(p(x0), p(x1), w, h, p(m), p(mdc)); ref().ref_mask(p(x0), p(x1), w, h, p(rm), p(rmdc))
    assert np.array_equal(m, rm) and np.array_equal(mdc, rmdc)


@pytest.mark.parametrize("w,h", [(32, 32), (33, 35), (34, 40)] + SIZES)
def test_diffmap_stages(w, h):
    _, _, x0, x1 = xyb_pair(w, h)
    n = w * h; rn = ((w + 2) // 3) * ((h + 2) // 3)
    shapes = [3 * n, 3 * n, 3 * rn, 3 * rn, 3 * rn, 3 * rn, 3 * n, 3 * n, rn, n]
    names = ["mhic0", "mhic1", "edge_map", "block_dc", "block_ac_pre", "block_ac", "mask",
             "mask_dc", "combined", "diffmap"]
    oo = [np.zeros(s, np.float32) for s in shapes]
    rr = [np.zeros(s, np.float32) for s in shapes]
    oracle().gzo_diffmap_stages(p(x0), p(x1), w, h, *[p(a) for a in oo])
    ref().ref_diffmap_stages(p(x0), p(x1), w, h, *[p(a) for a in rr])
    for nm, a, b in zip(names, oo, rr):
        assert np.array_equal(a, b), nm
    d = np.zeros(n, np.float32)
    ref().ref_diffmap(p(x0), p(x1), w, h, p(d))
    assert np.array_equal(d, oo[-1])


def test_integer_stages():
    rng = np.random.default_rng(5)
    for t in range(300):
        blk = (rng.normal(0, [4, 60, 400][t % 3], 64)).astype(np.int16)
        blk[0] = rng.integers(-1024, 1024)
        a = np.zeros(64, np.uint8); b = a.copy()
        oracle().gzo_idct(p(blk), p(a)); ref().ref_idct(p(blk), p(b))
        assert np.array_equal(a, b)
    for q in [1, 2, 3, 5, 8, 17, 64, 255]:
        for c in list(range(-300, 301)) + [-32768, -32767, 32767, 2047, -2048]:
            assert oracle().gzo_quantize(c, q) == ref().ref_quantize(c, q), (c, q)
    px = rng.integers(0, 256, (4096, 3)).astype(np.uint8)
    grid = np.stack(np.meshgrid(np.arange(0, 256, 5), np.arange(256), np.arange(256), indexing="ij"), -1)
    px = np.concatenate([px, grid.reshape(-1, 3).astype(np.uint8)])
    a = px.copy(); b = px.copy()
    oracle().gzo_ycbcr_to_rgb(p(a), len(a)); ref().ref_ycbcr_to_rgb(p(b), len(b))
    assert np.array_equal(a, b)
    for t in range(20):
        blk = rng.normal(0, 50, 64)
        for inv in (0, 1):
            a = blk.copy(); b = blk.copy()
            oracle().gzo_dct_double(p(a), inv)
            (ref().ref_idct_double if inv else ref().ref_dct_double)(p(b))
            assert np.array_equal(a, b)


@pytest.mark.parametrize("w,h,q", [(64, 48, 3), (97, 61, 5), (200, 133, 2)])
def test_compare_and_candidate_image(w, h, q):
    img = synth_image(w, h)
    s = RefSession(img, 0.97)
    s.apply_quant(np.full(192, q, np.int32))
    co = s.coeffs()
    jc = s.jpg_coeffs()
    mine = jc.copy()
    for c in range(3):
        oracle().gzo_apply_global_quant(p(mine[c]), s.nblocks, p(np.full(64, q, np.int32)))
    assert np.array_equal(mine, co)
    srgb = np.zeros((h, w, 3), np.uint8)
    oracle().gzo_coeffs_to_srgb(p(co[0]), p(co[1]), p(co[2]), w, h, p(srgb))
    assert np.array_equal(srgb, s.to_srgb())
    d_ref, dm_ref = s.compare()
    dm = np.zeros((h, w), np.float32)
    d = oracle().gzo_compare(p(img), p(co[0]), p(co[1]), p(co[2]), w, h, p(dm))
    assert d == d_ref and np.array_equal(dm, dm_ref)
    # ComputeBlockErrorAdjustmentWeights, both directions, all radii
    for direction in (1, -1):
        for rblock in (1, 2, 4):
            for tm in (1.0, 0.6, 2.5):
                wr = s.block_weights(direction, rblock, tm, dm)
                wo = np.zeros(s.nblocks, np.float32)
                oracle().gzo_block_weights(dm.reshape(-1), w, h, np.float32(0.97), direction, rblock, tm, wo)
                assert np.array_equal(wr, wo)
    s.close()


@pytest.mark.parametrize("w,h", [(64, 48), (70, 45)])
def test_block_comparisons_and_zeroing_order(w, h):
    img = synth_image(w, h)
    target = 0.971769
    s = RefSession(img, target)
    s.apply_quant(np.full(192, 3, np.int32))
    cur = s.coeffs(); orig = s.jpg_coeffs()
    mask_ref = s.start_block_comparisons()
    mask = np.zeros((3, h, w), np.float32)
    oracle().gzo_block_mask(p(img), w, h, p(mask))
    assert np.array_equal(mask, mask_ref)
    rng = np.random.default_rng(2)
    for b in [0, s.bw - 1, s.nblocks - 1, s.nblocks // 2]:
        bx, by = b % s.bw, b // s.bw
        pg_ref = s.switch_block(bx, by)
        pg = np.zeros(192, np.float32)
        oracle().gzo_block_pregamma(p(img), w, h, bx, by, p(pg))
        assert np.array_equal(pg, pg_ref)
        scale = np.ascontiguousarray(mask[:, 8 * by, 8 * bx])
        for t in range(6):
            cand = cur[:, b, :].reshape(192).copy()
            nz = np.flatnonzero(cand)
            if len(nz):
                cand[rng.choice(nz, min(len(nz), 1 + 3 * t), replace=False)] = 0
            e_ref = s.compare_block(bx, by, cand)
            e = oracle().gzo_compare_block(p(cand), w, h, bx, by, p(pg), p(scale))
            assert e == e_ref, (b, t)
    zo_ref = s.zeroing_order(7)
    zo = np.zeros((s.nblocks, 192), COEFF_DATA)
    ties = oracle().gzo_zeroing_order(p(img), w, h, p(orig[0]), p(orig[1]), p(orig[2]), p(cur[0]),
                                      p(cur[1]), p(cur[2]), p(mask), 7, C.c_float(target), 0,
                                      s.nblocks, p(zo))
    assert ties == 0
    assert np.array_equal(zo["idx"], zo_ref["idx"])
    assert np.array_equal(zo["err"], zo_ref["err"])
    assert (zo_ref["err"] > 0).sum() > s.nblocks  # non-trivial
    s.finish_block_comparisons()
    s.close()


def test_front_end_fdct_and_rgb_to_coeffs():
    rng = np.random.default_rng(9)
    for t in range(300):
        blk = rng.integers(-128, 128, 64).astype(np.int16)
        if t % 3 == 0:
            blk = (rng.normal(0, 20, 64) + rng.integers(-100, 100)).clip(-128, 127).astype(np.int16)
        a = blk.copy(); b = blk.copy()
        oracle().gzo_fdct(p(a)); ref().ref_fdct(p(b))
        assert np.array_equal(a, b)
    for (w, h) in [(64, 48), (70, 45), (33, 39)]:
        img = synth_image(w, h)
        s = RefSession(img, 1.0)
        nb = s.nblocks
        c = np.zeros((3, nb, 64), np.int16)
        oracle().gzo_rgb_to_jpeg_coeffs(p(img), w, h, p(c[0]), p(c[1]), p(c[2]))
        assert np.array_equal(c, s.jpg_coeffs())
        s.close()
    img = bees()
    h, w = img.shape[:2]
    s = RefSession(img, 1.0)
    c = np.zeros((3, s.nblocks, 64), np.int16)
    oracle().gzo_rgb_to_jpeg_coeffs(p(img), w, h, p(c[0]), p(c[1]), p(c[2]))
    assert np.array_equal(c, s.jpg_coeffs())
    s.close()


# ---- YUV 4:2:0 branch (oracle/gzoracle_yuv420.inc) -------------------------------------------------
from _libs import RefSession420, image_420   # noqa: E402

SIZES_420 = [("red", 48, 40), ("red", 100, 75), ("synth", 33, 47), ("red", 129, 66)]


def _jpeg_coeffs(img):
    h, w = img.shape[:2]
    nb = ((w + 7) // 8) * ((h + 7) // 8)
    c = np.zeros((3, nb, 64), np.int16)
    oracle().gzo_rgb_to_jpeg_coeffs(p(img), w, h, p(c[0]), p(c[1]), p(c[2]))
    return c


def _oracle_downsample(img):
    h, w = img.shape[:2]
    c = _jpeg_coeffs(img)
    mcw, mch = (w + 15) // 16, (h + 15) // 16
    out = [np.zeros((4 * mcw * mch, 64), np.int16), np.zeros((mcw * mch, 64), np.int16), np.zeros((mcw * mch, 64), np.int16)]
    assert oracle().gzo420_downsample(p(c[0]), p(c[1]), p(c[2]), w, h, p(out[0]), p(out[1]), p(out[2])) == 1
    return out


@pytest.mark.parametrize("kind,w,h", SIZES_420 + [("bees", 0, 0)])
def test_420_downsample(kind, w, h):
    """OutputImage::Downsample (PreProcessChannel for V then U, 2x2 average, double DCT) + SaveToJpegData."""
    img = image_420(kind, w, h)
    rs = RefSession420(img, 0.97)
    got, want = _oracle_downsample(img), rs.jpg_coeffs()
    for c in range(3):
        assert got[c].shape == want[c].shape and np.array_equal(got[c], want[c]), "component %d" % c
    rs.close()


def test_420_downsample_skips_grey():
    g = synth_image(48, 40, 3)[:, :, 0]
    img = np.ascontiguousarray(np.stack([g, g, g], axis=2))
    c = _jpeg_coeffs(img)
    dummy = np.zeros((64, 64), np.int16)
    assert oracle().gzo420_downsample(p(c[0]), p(c[1]), p(c[2]), 48, 40, p(dummy), p(dummy), p(dummy)) == 0


@pytest.mark.parametrize("kind,w,h", SIZES_420)
def test_420_candidate_closed_form_upsampling_and_compare(kind, w, h):
    """The factor-2 candidate as a per-pixel function of the coefficients equals the reference's
    block-by-block UpdatePixelsForBlock state -- after a full copy, after quantisation and after
    sparse SetCoeffBlock updates in an arbitrary order."""
    img = image_420(kind, w, h)
    h, w = img.shape[:2]
    rs = RefSession420(img, 0.97)
    rng = np.random.default_rng(5)
    q = rng.integers(1, 9, (3, 64)).astype(np.int32)
    rs.reset(); rs.apply_quant(q)
    coeffs = rs.coeffs()
    # scattered single-block updates (each re-renders an 18x18 pixel neighbourhood in the reference)
    for c in (1, 2, 1, 2, 0):
        nb = coeffs[c].shape[0]
        for b in rng.choice(nb, size=min(nb, 5), replace=False):
            coeffs[c][b, int(rng.integers(0, 64))] = int(rng.integers(-3, 4)) * int(q[c].max())
            rs.set_coeffs(coeffs)
    pad = rs.to_padded(rs.coeffs())
    got = np.zeros((h, w, 3), np.uint8)
    oracle().gzo420_to_srgb(p(pad[0]), p(pad[1]), p(pad[2]), w, h, p(got))
    assert np.array_equal(got, rs.to_srgb())
    dm = np.zeros((h, w), np.float32)
    d = oracle().gzo420_compare(p(img), p(pad[0]), p(pad[1]), p(pad[2]), w, h, p(dm))
    d_ref, dm_ref = rs.compare()
    assert np.array_equal(dm, dm_ref) and d == d_ref
    rs.close()


@pytest.mark.parametrize("kind,w,h", SIZES_420[:3])
@pytest.mark.parametrize("comp_mask", [1, 6])
def test_420_zeroing_order(kind, w, h, comp_mask):
    img = image_420(kind, w, h)
    h, w = img.shape[:2]
    target = 3.0 if comp_mask == 6 else 1.5
    rs = RefSession420(img, target)
    q = np.minimum(np.random.default_rng(3).integers(1, 9, (3, 64)), 3).astype(np.int32)
    rs.reset(); rs.apply_quant(q)
    mask = rs.start_block_comparisons()
    want = rs.zeroing_order_f(comp_mask)
    n = want.shape[0] if kind != "red" or w < 64 else min(want.shape[0], 24)   # bounded CPU time
    orig, cur = rs.jpg_coeffs(), rs.to_padded(rs.coeffs())
    got = np.zeros((n, 192), COEFF_DATA)
    oracle().gzo420_zeroing_order(p(img), w, h, p(orig[0]), p(orig[1]), p(orig[2]), p(cur[0]), p(cur[1]), p(cur[2]),
                                  p(mask), comp_mask, C.c_float(target), 0, n, p(got))
    assert np.array_equal(got["idx"], want["idx"][:n]) and np.array_equal(got["err"], want["err"][:n])
    assert (want["err"][:n] > 0).sum() > 10
    rs.close()


@pytest.mark.parametrize("direction,rblock", [(1, 1), (1, 3), (-1, 2), (-1, 4)])
@pytest.mark.parametrize("factor", [1, 2])
def test_block_weights_with_factor(direction, rblock, factor):
    img = image_420("red", 129, 66)
    rs = RefSession420(img, 0.6)
    rs.reset(); rs.apply_quant(np.full((3, 64), 5, np.int32))
    d, dm = rs.compare()
    mul = 1.0 if direction < 0 else 0.8 * float(dm.max()) / 0.6   # all but the worst region pass
    want = rs.block_weights_f(direction, rblock, mul, factor, dm)
    got = np.zeros_like(want)
    oracle().gzo_block_weights_f.argtypes = [np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS"), C.c_int, C.c_int,
                                             C.c_float, C.c_int, C.c_int, C.c_double, C.c_int,
                                             np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")]
    oracle().gzo_block_weights_f(np.ascontiguousarray(dm), 129, 66, np.float32(0.6), direction, rblock, mul, factor, got)
    assert np.array_equal(got, want) and got.max() > 0
    rs.close()
