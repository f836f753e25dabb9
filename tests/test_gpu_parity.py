"""GPU parity tests: the sm_100a CUDA path (through the C ABI of libgzb200.so) against the plain-C
oracle and, when present, the compiled reference (oracle/_ref). Bit-exact unless stated."""
import ctypes as C
import numpy as np
import pytest

from _libs import oracle, ref, have_ref, p, RefSession, synth_image, bees, COEFF_DATA
import __graft_entry__ as ge

pytestmark = pytest.mark.gpu


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
        rel = np.abs(a[bad].astype(np.float64) - b[bad]) / np.maximum(np.abs(b[bad]), 1e-30)
        pytest.fail("%s: %d/%d differ; first at %d: %r vs %r; max rel %.3g" %
                    (name, len(bad), a.size, i, a[i], b[i], rel.max()))


@pytest.mark.parametrize("sigma,br", [(1.1, 0.0), (1.5, 0.0), (0.586, 0.0), (0.4, 0.0), (14.0, 0.0),
                                      (9.65781083553, 0.0), (14.2644604355, 0.0),
                                      (4.53358927369, 0.0), (8.8510880283, 0.03027655136)])
@pytest.mark.parametrize("w,h", [(32, 32), (67, 45), (200, 133)])
def test_blur(gz, sigma, br, w, h):
    rng = np.random.default_rng(3)
    x = rng.uniform(0, 255, (h, w)).astype(np.float32)
    want = x.copy()
    oracle().gzo_blur(want, w, h, sigma, br)
    report("blur", gz.Blur(x, sigma, br), want)


@pytest.mark.parametrize("w,h", [(32, 32), (97, 61), (200, 133)])
def test_opsin_dynamics(gz, w, h):
    rng = np.random.default_rng(5)
    x = rng.uniform(0, 255, (3, h, w)).astype(np.float32)
    want = x.copy()
    oracle().gzo_opsin_dynamics_image(p(want), w, h)
    report("opsin", gz.OpsinDynamicsImage(x), want)


@pytest.mark.parametrize("w,h", [(32, 32), (33, 35), (64, 48), (97, 61), (200, 133), (444, 258)])
def test_diffmap_opsin_dynamics_image(gz, w, h):
    img = synth_image(w, h, 7)
    rng = np.random.default_rng(7)
    img2 = np.clip(img.astype(np.int32) + rng.integers(-6, 7, img.shape), 0, 255).astype(np.uint8)
    a = np.zeros((3, h, w), np.float32); b = a.copy()
    oracle().gzo_srgb_to_xyb(p(img), w, h, p(a)); oracle().gzo_srgb_to_xyb(p(img2), w, h, p(b))
    want = np.zeros((h, w), np.float32)
    oracle().gzo_diffmap(p(a), p(b), w, h, p(want))
    report("diffmap", gz.DiffmapOpsinDynamicsImage(a, b), want)
    d, dm = gz.ButteraugliSrgb(img, img2)
    report("butteraugli_srgb diffmap", dm, want)
    assert d == want.max()


@pytest.mark.parametrize("w,h", [(32, 32), (97, 61), (200, 133), (444, 258)])
def test_mask(gz, w, h):
    """gzb_mask == butteraugli::Mask (the cuMask hook): full-resolution mask and mask_dc planes, bit-exact."""
    img = synth_image(w, h, 11)
    rng = np.random.default_rng(11)
    img2 = np.clip(img.astype(np.int32) + rng.integers(-9, 10, img.shape), 0, 255).astype(np.uint8)
    a = np.zeros((3, h, w), np.float32); b = a.copy()
    oracle().gzo_srgb_to_xyb(p(img), w, h, p(a)); oracle().gzo_srgb_to_xyb(p(img2), w, h, p(b))
    want_m = np.zeros((3, h, w), np.float32); want_d = want_m.copy()
    oracle().gzo_mask(p(a), p(b), w, h, p(want_m), p(want_d))
    m, d = gz.Mask(a, b)
    report("mask", m, want_m)
    report("mask_dc", d, want_d)


@pytest.mark.parametrize("w,h,q", [(64, 48, 3), (97, 61, 5), (200, 133, 2), (70, 45, 4)])
def test_compare_stages(gz, w, h, q):
    img = synth_image(w, h)
    orig = jpeg_coeffs(img)
    cmp_ = gz.ButteraugliComparator(w, h, img, 0.97)
    cmp_.SetJpegCoeffs(orig)
    cmp_.CopyFromJpegData()
    report("copy_from_jpeg", cmp_.GetCoeffs(), orig)
    cmp_.ApplyGlobalQuantization(np.full(192, q, np.int32))
    cur = orig.copy()
    for c in range(3):
        oracle().gzo_apply_global_quant(p(cur[c]), cur.shape[1], p(np.full(64, q, np.int32)))
    report("apply_global_quant", cmp_.GetCoeffs(), cur)
    srgb = np.zeros((h, w, 3), np.uint8)
    oracle().gzo_coeffs_to_srgb(p(cur[0]), p(cur[1]), p(cur[2]), w, h, p(srgb))
    report("to_srgb", cmp_.ToSRGB(), srgb)
    d = cmp_.Compare()
    # stage-level comparison against the oracle
    x0 = np.zeros((3, h, w), np.float32); x1 = x0.copy()
    oracle().gzo_srgb_to_xyb(p(img), w, h, p(x0)); oracle().gzo_srgb_to_xyb(p(srgb), w, h, p(x1))
    report("xyb0", cmp_.debug_fetch("xyb0"), x0)
    report("xyb1", cmp_.debug_fetch("xyb1"), x1)
    n = w * h; rxs, rys = (w + 2) // 3, (h + 2) // 3; rn = rxs * rys
    shapes = [3 * n, 3 * n, 3 * rn, 3 * rn, 3 * rn, 3 * rn, 3 * n, 3 * n, rn, n]
    st = [np.zeros(s, np.float32) for s in shapes]
    oracle().gzo_diffmap_stages(p(x0), p(x1), w, h, *[p(a) for a in st])
    mh0, mh1, edm, dc, _, ac, _, _, comb, dm = st
    report("mhic0", cmp_.debug_fetch("mhic0"), mh0)
    report("mhic1", cmp_.debug_fetch("mhic1"), mh1)
    report("edge_map", cmp_.debug_fetch("edge_map"), edm)
    report("block_dc", cmp_.debug_fetch("block_dc"), dc)
    # block_ac: cells the reference never reads may differ only if never written; compare all
    report("block_ac", cmp_.debug_fetch("block_ac"), ac)
    comb_sq = np.where(comb < 1e-4, np.float32(100.0) * comb, np.sqrt(comb)).astype(np.float32)
    valid = np.zeros((rys, rxs), bool)
    valid[:len(range(0, h - 5, 3)), :len(range(0, w - 5, 3))] = True
    got = cmp_.debug_fetch("combined_sqrt").reshape(rys, rxs)
    report("combined_sqrt", got[valid], comb_sq.reshape(rys, rxs)[valid])
    report("diffmap", cmp_.distmap(), dm)
    assert d == dm.max()
    assert cmp_.DistanceOK(1.0) == bool(np.float32(d) <= 1.0 * float(np.float32(0.97)))
    assert cmp_.ScoreOutputSize(12345) == oracle().gzo_score_jpeg(float(d), 12345, float(np.float32(0.97)))
    # ComputeBlockErrorAdjustmentWeights from the resident map and from a host map
    for direction in (1, -1):
        for rblock in (1, 2, 4):
            for tm in (1.0, 0.6, 2.5):
                wo = np.zeros(cmp_.num_blocks, np.float32)
                oracle().gzo_block_weights(dm.reshape(-1), w, h, np.float32(0.97), direction, rblock, tm, wo)
                report("weights", cmp_.ComputeBlockErrorAdjustmentWeights(direction, rblock, tm), wo)
    report("weights(host map)", cmp_.ComputeBlockErrorAdjustmentWeights(-1, 2, 0.8, dm), (lambda wo: (oracle().gzo_block_weights(dm.reshape(-1), w, h, np.float32(0.97), -1, 2, 0.8, wo), wo)[1])(np.zeros(cmp_.num_blocks, np.float32)))
    cmp_.close()


@pytest.mark.parametrize("w,h", [(64, 48), (70, 45), (96, 64)])
def test_block_comparisons_and_zeroing_order(gz, w, h):
    img = synth_image(w, h)
    target = 0.971769
    orig = jpeg_coeffs(img)
    cmp_ = gz.ButteraugliComparator(w, h, img, target)
    cmp_.SetJpegCoeffs(orig)
    cmp_.CopyFromJpegData()
    cmp_.ApplyGlobalQuantization(np.full(192, 3, np.int32))
    cur = cmp_.GetCoeffs()
    cmp_.StartBlockComparisons()
    mask = np.zeros((3, h, w), np.float32)
    oracle().gzo_block_mask(p(img), w, h, p(mask))
    ms, ob = cmp_.BlockLists()
    bw = cmp_.block_width
    want_ms = np.stack([mask[:, 8 * (b // bw), 8 * (b % bw)] for b in range(cmp_.num_blocks)])
    report("mask_scale", ms, want_ms)
    want_ob = np.zeros((cmp_.num_blocks, 192), np.float32)
    for b in range(cmp_.num_blocks):
        oracle().gzo_block_pregamma(p(img), w, h, b % bw, b // bw, p(want_ob[b]))
    report("opsin_blocks", ob, want_ob)
    errs = cmp_.CompareBlocks()
    want = np.zeros(cmp_.num_blocks, np.float32)
    for b in range(cmp_.num_blocks):
        cand = np.ascontiguousarray(cur[:, b, :].reshape(192))
        sc = np.ascontiguousarray(want_ms[b])
        want[b] = np.float32(oracle().gzo_compare_block(p(cand), w, h, b % bw, b // bw, p(want_ob[b]), p(sc)))
    report("compare_blocks", errs, want)
    zo = cmp_.ComputeBlockZeroingOrder(7)
    zo_o = np.zeros((cmp_.num_blocks, 192), COEFF_DATA)
    ties = oracle().gzo_zeroing_order(p(img), w, h, p(orig[0]), p(orig[1]), p(orig[2]), p(cur[0]), p(cur[1]),
                                      p(cur[2]), p(mask), 7, C.c_float(target), 0, cmp_.num_blocks, p(zo_o))
    assert ties == 0
    report("zeroing idx", zo["idx"], zo_o["idx"])
    report("zeroing err", zo["err"], zo_o["err"])
    assert (zo["err"] > 0).sum() > cmp_.num_blocks
    # device-side candidate packing (processor.cc:694-712) against the same filter on the host
    off, cidx, cerr = cmp_.ComputeBlockZeroingCandidates(7)
    keep = (zo_o["err"] > 0) & (zo_o["err"] <= np.float32(target))
    want_off = np.concatenate([[0], np.cumsum(keep.sum(axis=1))]).astype(np.int32)
    report("candidate offsets", off, want_off)
    report("candidate idx", cidx, zo_o["idx"][keep].astype(np.uint8))
    report("candidate err", cerr, zo_o["err"][keep])
    cmp_.FinishBlockComparisons()
    cmp_.close()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_compare_against_compiled_reference_bees(gz):
    """bees.png (the reference's own test image): Compare + zeroing order vs the compiled reference."""
    img = bees()
    h, w = img.shape[:2]
    target = 0.971769
    s = RefSession(img, target)
    s.apply_quant(np.full(192, 3, np.int32))
    d_ref, dm_ref = s.compare()
    cmp_ = gz.ButteraugliComparator(w, h, img, target)
    cmp_.SetJpegCoeffs(s.jpg_coeffs())
    cmp_.CopyFromJpegData()
    cmp_.ApplyGlobalQuantization(np.full(192, 3, np.int32))
    report("coeffs", cmp_.GetCoeffs(), s.coeffs())
    d = cmp_.Compare()
    report("bees diffmap", cmp_.distmap(), dm_ref)
    assert d == d_ref
    s.start_block_comparisons()
    cmp_.StartBlockComparisons()
    nb = 4 * s.bw  # four block rows on the CPU reference (~1 s)
    zo_ref = s.zeroing_order(7, 0, nb)
    zo = cmp_.ComputeBlockZeroingOrder(7)
    report("bees zeroing idx", zo["idx"][:nb], zo_ref["idx"])
    report("bees zeroing err", zo["err"][:nb], zo_ref["err"])
    cmp_.close(); s.close()


@pytest.mark.parametrize("w,h", [(2048, 2048), (4000, 3000)])
def test_compare_at_bench_sizes_equals_oracle(gz, w, h):
    """Compare at the sizes of the bench workloads (4 and 12 MPix; TMA tile loads, decimated lattices and the
    persistent grids all at full scale) against the plain-C oracle: diffmap and distance bit-exact. The
    candidate is the image after ApplyGlobalQuantization with the all-3 matrix (SURVEY 8d's M2 input)."""
    img = synth_image(w, h, 1234)
    orig = jpeg_coeffs(img)
    cmp_ = gz.ButteraugliComparator(w, h, img, 0.971769)
    cmp_.SetJpegCoeffs(orig)
    cmp_.CopyFromJpegData()
    cmp_.ApplyGlobalQuantization(np.full(192, 3, np.int32))
    cur = cmp_.GetCoeffs()
    d = cmp_.Compare()
    dm = cmp_.distmap()
    cmp_.close()
    want = np.zeros((h, w), np.float32)
    d_o = oracle().gzo_compare(p(img), p(cur[0]), p(cur[1]), p(cur[2]), w, h, p(want))
    report("diffmap %dx%d" % (w, h), dm, want)
    assert d == d_o == want.max()


@pytest.mark.skipif(not have_ref(), reason="oracle/_ref not present")
def test_zeroing_order_with_equal_keys_is_std_sorts(gz):
    """Two candidates of one block can have the same ordering key (|orig| * csf, guetzli/processor.cc:401) --
    e.g. |116| * csf[70] == |57| * csf[84] in float. std::sort places such ties its own (unstable) way once a
    block has more than 16 candidates; the device detects the tie and runs the restated introsort. Checker:
    the compiled reference, whose std::sort is the real one. Bit-exact idx and err."""
    w, h = 64, 48
    img = synth_image(w, h, 21)
    s = RefSession(img, 0.971769)
    co = s.jpg_coeffs().copy()
    # (component, k, amplitude) pairs with equal float keys, found by tests/test_host_cpu.py's probe
    pairs = [((1, 6, 116), (1, 20, 57)), ((1, 40, 13), (2, 59, 174)), ((1, 6, 232), (1, 20, 114)),
             ((1, 27, 208), (2, 36, 109)), ((0, 46, 164), (0, 54, 193))]
    nb = co.shape[1]
    for b in range(nb):
        (c0, k0, a0), (c1, k1, a1) = pairs[b % len(pairs)]
        sg = -1 if (b // len(pairs)) % 2 else 1
        co[c0, b, k0] = sg * a0
        co[c1, b, k1] = -sg * a1
        if b % 3 == 0:          # a second tie in the same block
            (d0, l0, e0), (d1, l1, e1) = pairs[(b + 2) % len(pairs)]
            co[d0, b, l0] = e0
            co[d1, b, l1] = e1
    s.set_jpg_coeffs(co)
    s.reset()
    s.start_block_comparisons()
    zo_ref = s.zeroing_order(7)
    cmp_ = gz.ButteraugliComparator(w, h, img, 0.971769)
    cmp_.SetJpegCoeffs(co)
    cmp_.CopyFromJpegData()
    report("coeffs", cmp_.GetCoeffs(), s.coeffs())
    cmp_.StartBlockComparisons()
    zo = cmp_.ComputeBlockZeroingOrder(7)
    L = gz.lib()
    L.gzb_last_zeroing_tie_blocks.restype = C.c_uint
    assert L.gzb_last_zeroing_tie_blocks(cmp_._ctx) == nb      # every block was crafted to tie
    report("tied zeroing idx", zo["idx"], zo_ref["idx"])
    report("tied zeroing err", zo["err"], zo_ref["err"])
    cmp_.close(); s.close()


def test_dct_double(gz):
    """ComputeBlockDCTDouble / IDCTDouble (guetzli/dct_double.cc:47-85), bit-exact."""
    rng = np.random.default_rng(8)
    blocks = rng.normal(0, 60, (1000, 64))
    for inv in (False, True):
        want = blocks.copy()
        for b in want:
            oracle().gzo_dct_double(p(b), 1 if inv else 0)
        report("dct_double inverse=%s" % inv, gz.DctDouble(blocks, inv), want)
    # round trip is the identity up to the matrix's 10-decimal rounding
    rt = gz.DctDouble(gz.DctDouble(blocks, False), True)
    assert np.abs(rt - blocks).max() < 1e-6


def test_error_paths(gz):
    img = synth_image(64, 48)
    with pytest.raises(gz.GzbError):
        gz.ButteraugliComparator(16, 16, img[:16, :16], 1.0)   # < 32 px: explicit error, no fallback
    cmp_ = gz.ButteraugliComparator(64, 48, img, 1.0)
    with pytest.raises(gz.GzbError):
        cmp_.Compare()                                           # no candidate yet
    with pytest.raises(gz.GzbError):
        cmp_.ComputeBlockZeroingOrder(7)                         # StartBlockComparisons missing
    cmp_.close()


@pytest.mark.parametrize("w,h,yuv420,maxchg", [(200, 133, False, 8), (1000, 700, False, 2), (2048, 1536, False, 6),
                                               (257, 131, True, 4), (1200, 800, True, 2)])
def test_incremental_compare_equals_full(gz, w, h, yuv420, maxchg):
    """Compares that follow sparse coefficient updates recompute only the tiles the changed blocks can
    reach; diffmap and distance must equal those of a full Compare of the same coefficients (a second
    context that never saw the earlier states), bit for bit, over a chain of updates."""
    img = synth_image(w, h, 21)
    target = 0.97
    a = gz.ButteraugliComparator(w, h, img, target)
    b = gz.ButteraugliComparator(w, h, img, target)
    co = gz.RgbToJpegCoeffs(img)
    for c in (a, b):
        c.SetJpegCoeffs(co)
        if yuv420:
            c.Downsample420()
        c.CopyFromJpegData()
        c.ApplyGlobalQuantization(np.full(192, 4, np.int32))
    a.Compare()
    rng = np.random.default_rng(9)
    base = a.incremental_compare_count()
    assert base == 0
    for step in range(12):
        cur = a.GetCoeffs()
        blocks, idxs, vals = [], [], []
        nchg = int(rng.integers(1, maxchg + 1))
        for _ in range(nchg):
            comp = int(rng.integers(0, 3))
            nb = cur[comp].shape[0]
            # favour the image border and corners, where the clamped windows and blur borders live
            blk = int(rng.integers(0, nb)) if rng.random() < 0.6 else int(rng.choice([0, nb - 1, nb // 2]))
            k = int(rng.integers(0, 64))
            v = int(rng.integers(-8, 9)) * 4
            blocks.append(blk); idxs.append(64 * comp + k); vals.append(v)
        a.UpdateCoeffs(blocks, idxs, vals)
        d = a.Compare()
        b.SetCoeffs(a.GetCoeffs())
        d_full = b.Compare()
        report("incremental diffmap step %d" % step, a.distmap(), b.distmap())
        assert float(d) == float(d_full)
    if w * h >= 700000:
        assert a.incremental_compare_count() >= 8, a.incremental_compare_count()
    else:   # the reachable area covers most of a small image: the Compares stay full
        assert a.incremental_compare_count() == 0
    assert b.incremental_compare_count() == 0
    a.close(); b.close()


@pytest.mark.parametrize("w,h,yuv420,frac", [(200, 133, False, 0.05), (1000, 700, False, 0.03), (2048, 1536, False, 0.01),
                                             (1000, 700, False, 0.4), (257, 131, True, 0.05), (1200, 800, True, 0.02)])
def test_compare_after_many_scattered_updates_equals_full(gz, w, h, yuv420, frac):
    """The back end's iterations flip coefficients in hundreds to thousands of scattered blocks: too many for
    the tile masks, but BlockDiffMap -- a third of a Compare -- only recomputes the cells whose window a flipped
    block can reach (k_block_diff_map / k_block_dc with BlockChanges). Must equal a full Compare bit for bit.
    (4:2:0 updates through gzb_update_coeffs are not recorded per block: those Compares are simply full.)"""
    img = synth_image(w, h, 23)
    a = gz.ButteraugliComparator(w, h, img, 0.97)
    b = gz.ButteraugliComparator(w, h, img, 0.97)
    co = gz.RgbToJpegCoeffs(img)
    for c in (a, b):
        c.SetJpegCoeffs(co)
        if yuv420:
            c.Downsample420()
        c.CopyFromJpegData()
        c.ApplyGlobalQuantization(np.full(192, 4, np.int32))
    a.Compare()
    rng = np.random.default_rng(29)
    for step in range(6):
        cur = a.GetCoeffs()
        blocks, idxs, vals = [], [], []
        for comp in range(3):
            nb = cur[comp].shape[0]
            n = max(1, int(frac * nb))
            pick = rng.choice(nb, n, replace=False)
            if step % 2 == 0:      # also the corners and the last block row / column
                pick = np.unique(np.concatenate([pick, [0, nb - 1]]))
            for blk in pick:
                blocks.append(int(blk)); idxs.append(64 * comp + int(rng.integers(0, 64))); vals.append(int(rng.integers(-8, 9)) * 4)
        a.UpdateCoeffs(blocks, idxs, vals)
        d = a.Compare()
        b.SetCoeffs(a.GetCoeffs())
        d_full = b.Compare()
        report("diffmap after scattered updates, step %d" % step, a.distmap(), b.distmap())
        assert float(d) == float(d_full)
    assert a.fine_bdm_compare_count() == (0 if yuv420 else 6)
    assert b.fine_bdm_compare_count() == 0
    a.close(); b.close()


@pytest.mark.parametrize("w,h", [(32, 32), (33, 47), (97, 61), (200, 133), (444, 258), (1024, 768)])
def test_rgb_front_end_on_device(gz, w, h):
    """EncodeRGBToJpeg with q = 1 (RGB -> YCbCr fixed point, integer forward DCT) on the device equals the
    oracle's restatement (itself pinned to the reference) and the library's host implementation."""
    rng = np.random.default_rng(4)
    img = synth_image(w, h, 77) if (w + h) % 2 else rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    c = gz.ButteraugliComparator(w, h, img, 0.97)
    got = c.RgbToJpegCoeffsDevice()
    report("device front end vs oracle", got, jpeg_coeffs(img))
    report("device front end vs host", got, gz.RgbToJpegCoeffs(img))
    c.close()
