"""ctypes loaders for the two checkers (test infrastructure):
   * oracle/libgzoracle.so      -- plain-C restatement (oracle/gzoracle.c)
   * oracle/_ref/libgzref.so    -- the unmodified reference compiled from /root/reference
plus small numpy-friendly wrappers and the seeded synthetic image generator."""
import ctypes as C
import os
import subprocess
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")

c_f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
c_f64p = np.ctypeslib.ndpointer(np.float64, flags="C_CONTIGUOUS")
c_u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
c_i16p = np.ctypeslib.ndpointer(np.int16, flags="C_CONTIGUOUS")
c_i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")

COEFF_DATA = np.dtype([("idx", np.int32), ("err", np.float32)])


def _build(target):
    subprocess.run(["make", "-C", ORACLE_DIR, target], check=True, stdout=subprocess.DEVNULL)


_oracle = None
_ref = None


def oracle():
    global _oracle
    if _oracle is None:
        path = os.path.join(ORACLE_DIR, "libgzoracle.so")
        src = os.path.join(ORACLE_DIR, "gzoracle.c")
        inc = os.path.join(ORACLE_DIR, "gzoracle_yuv420.inc")
        if not os.path.exists(path) or os.path.getmtime(path) < max(os.path.getmtime(src), os.path.getmtime(inc)):
            _build("port")
        _oracle = C.CDLL(path)
        L = _oracle
        L.gzo_compare.restype = C.c_float
        L.gzo420_compare.restype = C.c_float
        L.gzo_compare_block.restype = C.c_double
        L.gzo_score_from_diffmap.restype = C.c_float
        L.gzo_score_jpeg.restype = C.c_double
        L.gzo_score_jpeg.argtypes = [C.c_double, C.c_int, C.c_double]
        L.gzo_blur.argtypes = [c_f32p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.gzo_block_weights.argtypes = [c_f32p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int,
                                        C.c_double, c_f32p]
    return _oracle


def have_ref():
    return os.path.exists(os.path.join(ORACLE_DIR, "_ref", "libgzref.so")) or \
        os.path.exists("/root/reference/guetzli/processor.cc")


_dropin = None


def ref_dropin():
    """oracle/_ref/libgzref_dropin.so: the unmodified reference Processor over the product's comparator
    (the only checker library that links libgzb200.so; GPU tests only)."""
    global _dropin
    if _dropin is None:
        ref()
        path = os.path.join(ORACLE_DIR, "_ref", "libgzref_dropin.so")
        if not os.path.exists(path):
            _build("dropin")
        _dropin = C.CDLL(path)
    return _dropin


def ref():
    global _ref
    if _ref is None:
        path = os.path.join(ORACLE_DIR, "_ref", "libgzref.so")
        if not os.path.exists(path):
            _build("ref")
        _ref = C.CDLL(path)
        L = _ref
        L.ref_butteraugli_score_for_quality.restype = C.c_double
        L.ref_butteraugli_score_for_quality.argtypes = [C.c_double]
        L.ref_score_jpeg.restype = C.c_double
        L.ref_score_jpeg.argtypes = [C.c_double, C.c_int, C.c_double]
        L.ref_blur.argtypes = [c_f32p, C.c_int, C.c_int, C.c_double, C.c_double]
        L.ref_session_new.restype = C.c_void_p
        L.ref_session_new.argtypes = [c_u8p, C.c_int, C.c_int, C.c_float]
        L.ref_session_compare.restype = C.c_float
        L.ref_session_compare_block.restype = C.c_double
        L.ref_process_rgb.restype = C.c_long
        L.ref_process_rgb_params.restype = C.c_long
        L.ref_session_block_weights_f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int,
                                                  c_f32p, c_f32p, C.c_int]
        L.ref_session_block_weights.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double,
                                                c_f32p, c_f32p]
        for name in ("ref_session_free", "ref_session_dims", "ref_session_jpg_coeffs",
                     "ref_session_reset", "ref_session_apply_quant", "ref_session_get_coeffs",
                     "ref_session_set_coeffs", "ref_session_to_srgb", "ref_session_to_linear",
                     "ref_session_compare", "ref_session_jpeg_size", "ref_session_write_jpeg",
                     "ref_session_start_block_comparisons",
                     "ref_session_finish_block_comparisons", "ref_session_switch_block",
                     "ref_session_compare_block", "ref_session_zeroing_order", "ref_session_downsample",
                     "ref_session_comp_dims", "ref_session_zeroing_order_f"):
            fn = getattr(L, name)
            if fn.argtypes is None:
                fn.argtypes = None  # first arg passed explicitly as c_void_p by RefSession
    return _ref


def p(a):
    """numpy array -> void* (keeps `a` alive only for the call: pass named arrays)."""
    return a.ctypes.data_as(C.c_void_p)


class RefSession:
    """Encoder state of the compiled reference (oracle/ref_shim.cc `Session`)."""

    def __init__(self, rgb, target):
        self.L = ref()
        self.h, self.w = rgb.shape[:2]
        self.rgb = np.ascontiguousarray(rgb, np.uint8)
        self.s = C.c_void_p(self.L.ref_session_new(self.rgb, self.w, self.h, C.c_float(target)))
        assert self.s.value
        bw, bh = C.c_int(), C.c_int()
        self.L.ref_session_dims(self.s, C.byref(bw), C.byref(bh))
        self.bw, self.bh = bw.value, bh.value
        self.nblocks = self.bw * self.bh

    def close(self):
        if self.s:
            self.L.ref_session_free(self.s)
            self.s = None

    def __del__(self):
        self.close()

    def jpg_coeffs(self):
        out = np.zeros((3, self.nblocks, 64), np.int16)
        for c in range(3):
            self.L.ref_session_jpg_coeffs(self.s, c, p(out[c]))
        return out

    def reset(self):
        self.L.ref_session_reset(self.s)

    def set_jpg_coeffs(self, coeffs):
        """Overwrites the q=1 input coefficients of the session (jpg.components[c].coeffs)."""
        coeffs = np.ascontiguousarray(coeffs, np.int16)
        self.L.ref_session_set_jpg_coeffs(self.s, p(coeffs[0]), p(coeffs[1]), p(coeffs[2]))

    def apply_quant(self, q):
        q = np.ascontiguousarray(q, np.int32).reshape(192)
        self.L.ref_session_apply_quant(self.s, p(q))

    def coeffs(self):
        out = np.zeros((3, self.nblocks, 64), np.int16)
        for c in range(3):
            self.L.ref_session_get_coeffs(self.s, c, p(out[c]))
        return out

    def set_coeffs(self, coeffs):
        coeffs = np.ascontiguousarray(coeffs, np.int16)
        for c in range(3):
            self.L.ref_session_set_coeffs(self.s, c, p(coeffs[c]))

    def to_srgb(self):
        out = np.zeros((self.h, self.w, 3), np.uint8)
        self.L.ref_session_to_srgb(self.s, p(out))
        return out

    def compare(self):
        dm = np.zeros((self.h, self.w), np.float32)
        d = self.L.ref_session_compare(self.s, p(dm))
        return d, dm

    def write_jpeg(self):
        buf = np.zeros(self.w * self.h * 3 + (1 << 16), np.uint8)
        n = self.L.ref_session_write_jpeg(self.s, p(buf), C.c_int(buf.size))
        return buf[:n].tobytes()

    def start_block_comparisons(self):
        m = np.zeros((3, self.h, self.w), np.float32)
        self.L.ref_session_start_block_comparisons(self.s, p(m))
        return m

    def finish_block_comparisons(self):
        self.L.ref_session_finish_block_comparisons(self.s)

    def switch_block(self, bx, by):
        out = np.zeros(192, np.float32)
        self.L.ref_session_switch_block(self.s, bx, by, p(out))
        return out

    def compare_block(self, bx, by, cand192):
        cand = np.ascontiguousarray(cand192, np.int16).reshape(192)
        return self.L.ref_session_compare_block(self.s, bx, by, p(cand))

    def zeroing_order(self, comp_mask=7, begin=0, end=None):
        end = self.nblocks if end is None else end
        out = np.zeros((end - begin, 192), COEFF_DATA)
        self.L.ref_session_zeroing_order(self.s, comp_mask, begin, end, p(out))
        return out

    def block_weights(self, direction, rblock, target_mul, distmap, weights=None):
        w = np.zeros(self.nblocks, np.float32) if weights is None else weights.copy()
        dm = np.ascontiguousarray(distmap, np.float32).reshape(-1)
        self.L.ref_session_block_weights(self.s, direction, rblock, float(target_mul), dm, w)
        return w


class RefSession420(RefSession):
    """The session after Processor::DownsampleImage + SaveToJpegData (processor.cc:994-997)."""

    def __init__(self, rgb, target):
        RefSession.__init__(self, rgb, target)
        self.L.ref_session_downsample(self.s)
        self.img_dims, self.jpg_dims, self.factor = [], [], []
        for c in range(3):
            v = [C.c_int() for _ in range(5)]
            self.L.ref_session_comp_dims(self.s, c, *[C.byref(x) for x in v])
            self.img_dims.append((v[0].value, v[1].value))
            self.jpg_dims.append((v[2].value, v[3].value))
            self.factor.append(v[4].value)

    def jpg_coeffs(self):
        """jpg.components[c].coeffs (MCU-padded layout): list of [blocks, 64]."""
        out = [np.zeros((bw * bh, 64), np.int16) for (bw, bh) in self.jpg_dims]
        for c in range(3):
            self.L.ref_session_jpg_coeffs(self.s, c, p(out[c]))
        return out

    def coeffs(self):
        """OutputImage coefficients (image layout): list of [blocks, 64]."""
        out = [np.zeros((bw * bh, 64), np.int16) for (bw, bh) in self.img_dims]
        for c in range(3):
            self.L.ref_session_get_coeffs(self.s, c, p(out[c]))
        return out

    def set_coeffs(self, coeffs):
        for c in range(3):
            a = np.ascontiguousarray(coeffs[c], np.int16)
            self.L.ref_session_set_coeffs(self.s, c, p(a))

    def to_padded(self, coeffs):
        """Image-layout coefficient arrays -> the MCU-padded layout of SaveToJpegData (padding
        blocks = {DC of the raster predecessor, 0...}, output_image.cc:608-632)."""
        out = []
        for c in range(3):
            (bw, bh), (jw, jh) = self.img_dims[c], self.jpg_dims[c]
            src = np.asarray(coeffs[c]).reshape(bh, bw, 64)
            dst = np.zeros((jh, jw, 64), np.int16)
            last = 0
            for by in range(jh):
                for bx in range(jw):
                    if by < bh and bx < bw:
                        dst[by, bx] = src[by, bx]
                    else:
                        dst[by, bx, 0] = last
                    last = dst[by, bx, 0]
            out.append(dst.reshape(-1, 64))
        return out

    def from_padded(self, coeffs):
        out = []
        for c in range(3):
            (bw, bh), (jw, jh) = self.img_dims[c], self.jpg_dims[c]
            out.append(np.ascontiguousarray(np.asarray(coeffs[c]).reshape(jh, jw, 64)[:bh, :bw].reshape(-1, 64)))
        return out

    def zeroing_order_f(self, comp_mask, begin=0, end=None):
        f = self.factor[2 if comp_mask & 4 else (1 if comp_mask & 2 else 0)]
        n = ((self.w + 8 * f - 1) // (8 * f)) * ((self.h + 8 * f - 1) // (8 * f))
        end = n if end is None else end
        out = np.zeros((end - begin, 192), COEFF_DATA)
        self.L.ref_session_zeroing_order_f(self.s, comp_mask, begin, end, p(out))
        return out

    def block_weights_f(self, direction, rblock, target_mul, factor, distmap):
        bs = 8 * factor
        n = ((self.w + bs - 1) // bs) * ((self.h + bs - 1) // bs)
        w = np.zeros(n, np.float32)
        dm = np.ascontiguousarray(distmap, np.float32).reshape(-1)
        self.L.ref_session_block_weights_f(self.s, direction, rblock, float(target_mul), factor, dm, w, n)
        return w


def ref_process_params(rgb, target, try_420=False, force_420=False, want_trace=False):
    L = ref()
    h, w = rgb.shape[:2]
    rgb = np.ascontiguousarray(rgb, np.uint8)
    out = np.zeros(w * h * 3 + (1 << 16), np.uint8)
    trace = C.create_string_buffer(1 << 22) if want_trace else None
    iters = C.c_int()
    n = L.ref_process_rgb_params(p(rgb), w, h, C.c_float(target), int(try_420), int(force_420), p(out),
                                 C.c_long(out.size), trace, C.c_long(1 << 22), C.byref(iters))
    assert n > 0
    return out[:n].tobytes(), iters.value, (trace.value.decode() if want_trace else None)


def ref_process(rgb, target, want_trace=False):
    L = ref()
    h, w = rgb.shape[:2]
    rgb = np.ascontiguousarray(rgb, np.uint8)
    out = np.zeros(w * h * 3 + (1 << 16), np.uint8)
    trace = C.create_string_buffer(1 << 22) if want_trace else None
    iters = C.c_int()
    n = L.ref_process_rgb(p(rgb), w, h, C.c_float(target), p(out), C.c_long(out.size), trace,
                          C.c_long(1 << 22), C.byref(iters))
    assert n > 0
    return out[:n].tobytes(), iters.value, (trace.value.decode() if want_trace else None)


def synth_image(w, h, seed=1234):
    """Deterministic 'photo-like' test image (SURVEY.md section 8d recipe)."""
    import cv2
    img = np.zeros((h, w, 3), np.uint8)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    for c in range(3):
        rng = np.random.Generator(np.random.PCG64(seed + c))
        phi = rng.uniform(0, 2 * np.pi)
        base = 128 + 80 * np.sin(2 * np.pi * 3 * xx / w + phi) * np.cos(2 * np.pi * 2 * yy / h + 0.5 * c)
        n1 = cv2.GaussianBlur(rng.standard_normal((h, w)), (0, 0), 2.0)
        n1 /= n1.std()
        n2 = cv2.GaussianBlur(rng.standard_normal((h, w)), (0, 0), 0.7)
        n2 /= n2.std()
        img[:, :, c] = np.clip(np.rint(base + 24 * n1 + 6 * n2), 0, 255).astype(np.uint8)
    return img


def image_420(kind, w, h):
    """Inputs of the 4:2:0 tests. "red": saturated red texture (PreProcessChannel sharpens V there),
    dark smooth blue-green (it blurs there), a bright patch and photo-like noise side by side."""
    if kind == "bees":
        return bees()
    if kind == "synth":
        return synth_image(w, h, 777)
    if kind == "gray":
        g = synth_image(w, h, 4321)[:, :, 0]
        return np.ascontiguousarray(np.stack([g, g, g], axis=2))
    rng = np.random.Generator(np.random.PCG64(4242))
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float64)
    img = synth_image(w, h, 31).astype(np.float64)
    a, b = w // 3, 2 * w // 3
    tex = 25 * np.sin(xx / 2.3) * np.cos(yy / 3.1) + 6 * rng.standard_normal((h, w))
    img[:, :a, 0] = 170 + tex[:, :a]
    img[:, :a, 1] = 30 + 0.2 * tex[:, :a]
    img[:, :a, 2] = 35 + 0.1 * tex[:, :a]
    img[:, a:b, 0] = 20 + 10 * yy[:, a:b] / h
    img[:, a:b, 1] = 70 + 20 * xx[:, a:b] / w
    img[:, a:b, 2] = 110 + 25 * yy[:, a:b] / h
    img[h // 4:h // 2, a + 4:b - 4, :] = 245
    return np.ascontiguousarray(np.clip(np.rint(img), 0, 255).astype(np.uint8))


def bees():
    from PIL import Image
    path = os.path.join(ROOT, "tests", "golden", "bees.png")
    return np.ascontiguousarray(np.asarray(Image.open(path).convert("RGB")))


SPECIAL_IMAGES = ["gray_80x64_q90", "min_32x32_q95", "odd_40x33_q100", "flat_64x64_q90", "noise_64x48_q84",
                  "edges_120x72_q92", "graynoise_48x40_q88"]


def special_image(name):
    """Edge-case inputs of the encoder tests: grey-only (chroma planes quantise to nothing and are
    dropped from the file), the smallest size the search accepts, sizes that are no multiple of 8, a
    flat image (no AC energy), white noise (no quant matrix passes), hard edges."""
    kind, size, q = name.split("_")
    w, h = (int(v) for v in size.split("x"))
    quality = int(q[1:])
    rng = np.random.Generator(np.random.PCG64(77))
    if kind in ("gray", "graynoise"):
        g = synth_image(w, h, 4321)[:, :, 0]
        if kind == "graynoise":
            g = rng.integers(0, 256, (h, w)).astype(np.uint8)
        img = np.stack([g, g, g], axis=2)
    elif kind in ("min", "odd"):
        img = synth_image(w, h, 999)
    elif kind == "flat":
        img = np.full((h, w, 3), 128, np.uint8)
    elif kind == "noise":
        img = rng.integers(0, 256, (h, w, 3)).astype(np.uint8)
    elif kind == "edges":
        img = np.zeros((h, w, 3), np.uint8)
        img[:, w // 3:, 0] = 255
        img[h // 2:, :, 1] = 200
        img[::7, :, 2] = 90
        img[:, ::11, :] //= 2
    else:
        raise ValueError(name)
    return np.ascontiguousarray(img), quality
