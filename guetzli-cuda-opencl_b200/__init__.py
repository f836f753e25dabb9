"""gzb200 -- B200-native (sm_100a CUDA) butteraugli-guided quantisation search for Guetzli.

Thin ctypes binding over the C ABI of ``libgzb200.so`` (include/gzb200.h). The class and method
names mirror the reference's operator interface for this path -- ``guetzli::Comparator`` /
``ButteraugliComparator`` (guetzli/comparator.h:29-96, guetzli/butteraugli_comparator.h:33-83) and
the ``cu*`` free functions (clguetzli/cuguetzli.h) -- so the parity tests read like the reference.

There is NO CPU fallback: importing works without a GPU (so the symbol table can be checked), but
every compute call fails loudly if the CUDA library or a device is missing. Nothing here imports
or links anything under ``oracle/``.

The directory name is not a Python identifier; load it with ``__graft_entry__.load_package()``
(importlib, module name ``gzb200``).
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GZB200_LIB") or os.path.join(_HERE, "libgzb200.so")   # GZB200_LIB: A/B timing of two builds

COEFF_DATA = np.dtype([("idx", np.int32), ("err", np.float32)])


class GzbError(RuntimeError):
    pass


_lib = None


def lib():
    """Loads libgzb200.so (raises if it has not been built: no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GzbError("libgzb200.so is not built (run __graft_entry__.build()); "
                           "there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.gzb_version.restype = C.c_char_p
        L.gzb_last_error.restype = C.c_char_p
        L.gzb_last_error.argtypes = [C.c_void_p]
        L.gzb_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_float,
                                 C.POINTER(C.c_void_p)]
        L.gzb_destroy.argtypes = [C.c_void_p]
        L.gzb_destroy.restype = None
        L.gzb_score_output_size.restype = C.c_double
        L.gzb_score_output_size.argtypes = [C.c_void_p, C.c_int]
        L.gzb_block_error_limit.restype = C.c_float
        L.gzb_block_error_limit.argtypes = [C.c_void_p]
        L.gzb_distance_ok.argtypes = [C.c_void_p, C.c_double]
        L.gzb_last_device_ms.restype = C.c_float
        L.gzb_last_device_ms.argtypes = [C.c_void_p]
        L.gzb_launch_count.restype = C.c_ulonglong
        L.gzb_launch_count.argtypes = [C.c_void_p]
        L.gzb_compute_block_error_adjustment_weights.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_void_p]
        L.gzb_blur.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_double, C.c_double]
        L.gzb_opsin_dynamics_image.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                               C.c_size_t, C.c_size_t]
        L.gzb_diffmap_opsin_dynamics_image.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_size_t] * 3
        L.gzb_butteraugli_srgb.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                           C.c_void_p, C.c_void_p]
        L.gzb_debug_fetch.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p, C.c_size_t,
                                      C.POINTER(C.c_size_t)]
        L.gzb_update_coeffs.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t]
        for name in ("gzb_set_jpeg_coeffs", "gzb_set_coeffs", "gzb_get_coeffs"):
            getattr(L, name).argtypes = [C.c_void_p] * 4
        for name in ("gzb_copy_from_jpeg", "gzb_apply_global_quantization", "gzb_to_srgb",
                     "gzb_compare", "gzb_get_distmap", "gzb_compare_blocks"):
            getattr(L, name).argtypes = [C.c_void_p, C.c_void_p]
        for name in ("gzb_start_block_comparisons", "gzb_finish_block_comparisons"):
            getattr(L, name).argtypes = [C.c_void_p]
        L.gzb_get_block_lists.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.gzb_compute_block_zeroing_order.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc, ctx=None):
    if rc != 0:
        msg = lib().gzb_last_error(ctx).decode(errors="replace")
        raise GzbError("gzb200 error %d: %s" % (rc, msg))


def device_count():
    return lib().gzb_device_count()


class ButteraugliComparator:
    """guetzli::ButteraugliComparator on a B200, with the candidate OutputImage held on the device.

    Mirrors guetzli/butteraugli_comparator.h:33-83; the OutputImage methods the search driver
    needs (CopyFromJpegData, ApplyGlobalQuantization, ToSRGB, SetCoeffBlock) are methods here
    because the candidate lives in HBM next to the comparator state.
    """

    def __init__(self, width, height, rgb, target_distance, device=0):
        rgb = np.ascontiguousarray(rgb, np.uint8)
        if rgb.size != width * height * 3:
            raise ValueError("rgb must hold width*height*3 bytes")
        self.width, self.height = int(width), int(height)
        self.block_width, self.block_height = (width + 7) // 8, (height + 7) // 8
        self.num_blocks = self.block_width * self.block_height
        self.target_distance = np.float32(target_distance)
        self._ctx = C.c_void_p()
        _check(lib().gzb_create(device, width, height, _p(rgb), C.c_float(target_distance),
                                C.byref(self._ctx)))
        self.distance = np.float32(0.0)

    def close(self):
        if self._ctx:
            lib().gzb_destroy(self._ctx)
            self._ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- candidate image ------------------------------------------------------------------
    def SetJpegCoeffs(self, coeffs):
        c = np.ascontiguousarray(coeffs, np.int16).reshape(3, self.num_blocks, 64)
        _check(lib().gzb_set_jpeg_coeffs(self._ctx, _p(c[0]), _p(c[1]), _p(c[2])), self._ctx)
        self.is_420 = False

    def RgbToJpegCoeffsDevice(self):
        """EncodeRGBToJpeg (q = 1) of the context's original image on the device; returns [3, blocks, 64]."""
        L = lib()
        L.gzb_rgb_to_jpeg_coeffs_device.argtypes = [C.c_void_p]
        _check(L.gzb_rgb_to_jpeg_coeffs_device(self._ctx), self._ctx)
        self.is_420 = False
        return np.stack(self.GetJpegCoeffs())

    def CopyFromJpegData(self, quant=None):
        q = np.ones(192, np.int32) if quant is None else np.ascontiguousarray(quant, np.int32).reshape(192)
        _check(lib().gzb_copy_from_jpeg(self._ctx, _p(q)), self._ctx)

    def ApplyGlobalQuantization(self, q):
        q = np.ascontiguousarray(q, np.int32).reshape(192)
        _check(lib().gzb_apply_global_quantization(self._ctx, _p(q)), self._ctx)

    def SetCoeffs(self, coeffs):
        if self.is_420:
            c = [np.ascontiguousarray(coeffs[k], np.int16).reshape(self.ComponentDims(k)[2], 64) for k in range(3)]
        else:
            c = np.ascontiguousarray(coeffs, np.int16).reshape(3, self.num_blocks, 64)
        _check(lib().gzb_set_coeffs(self._ctx, _p(c[0]), _p(c[1]), _p(c[2])), self._ctx)

    def GetCoeffs(self):
        """[3, num_blocks, 64] int16; for a 4:2:0 image a list of three [blocks_c, 64] arrays."""
        if self.is_420:
            c = [np.zeros((self.ComponentDims(k)[2], 64), np.int16) for k in range(3)]
        else:
            c = np.zeros((3, self.num_blocks, 64), np.int16)
        _check(lib().gzb_get_coeffs(self._ctx, _p(c[0]), _p(c[1]), _p(c[2])), self._ctx)
        return c

    # ---- YUV 4:2:0 (OutputImage::Downsample + SaveToJpegData) -------------------------------
    is_420 = False

    def Downsample420(self):
        L = lib()
        L.gzb_downsample_420.argtypes = [C.c_void_p]
        _check(L.gzb_downsample_420(self._ctx), self._ctx)
        self.is_420 = True

    def ComponentDims(self, comp):
        """(blocks per row, blocks per column, blocks, sampling factor) of a component's coefficient array."""
        L = lib()
        L.gzb_component_dims.argtypes = [C.c_void_p, C.c_int] + [C.POINTER(C.c_int)] * 3
        bw, bh, f = C.c_int(), C.c_int(), C.c_int()
        _check(L.gzb_component_dims(self._ctx, comp, C.byref(bw), C.byref(bh), C.byref(f)), self._ctx)
        return bw.value, bh.value, bw.value * bh.value, f.value

    def GetJpegCoeffs(self):
        """The input coefficients (jpg.components[c].coeffs): list of three [blocks_c, 64] arrays."""
        L = lib()
        L.gzb_get_jpeg_coeffs.argtypes = [C.c_void_p] * 4
        c = [np.zeros((self.ComponentDims(k)[2], 64), np.int16) for k in range(3)]
        _check(L.gzb_get_jpeg_coeffs(self._ctx, _p(c[0]), _p(c[1]), _p(c[2])), self._ctx)
        return c

    def ZeroingUnits(self, comp_mask):
        if self.is_420 and (comp_mask & 6):
            return ((self.width + 15) // 16) * ((self.height + 15) // 16)
        return self.num_blocks

    def UpdateCoeffs(self, block_ix, idx, val):
        b = np.ascontiguousarray(block_ix, np.int32)
        i = np.ascontiguousarray(idx, np.uint8)
        v = np.ascontiguousarray(val, np.int16)
        _check(lib().gzb_update_coeffs(self._ctx, _p(b), _p(i), _p(v), len(b)), self._ctx)

    def ToSRGB(self):
        out = np.zeros((self.height, self.width, 3), np.uint8)
        _check(lib().gzb_to_srgb(self._ctx, _p(out)), self._ctx)
        return out

    def WriteJpeg(self, q, input_tables=False, want_bytes=True):
        """SaveToJpegData + WriteJpeg of the resident candidate, the scan Huffman-coded on the device.
        q[192]: the matrix the candidate's values are multiples of. Returns (size, bytes or None)."""
        L = lib()
        L.gzb_write_candidate_jpeg.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        qq = np.ascontiguousarray(q, np.int32).reshape(192)
        n = C.c_size_t()
        _check(L.gzb_write_candidate_jpeg(self._ctx, _p(qq), int(input_tables), None, 0, C.byref(n)), self._ctx)
        if not want_bytes:
            return n.value, None
        buf = np.zeros(n.value, np.uint8)
        _check(L.gzb_write_candidate_jpeg(self._ctx, _p(qq), int(input_tables), _p(buf), buf.size, C.byref(n)), self._ctx)
        return n.value, buf.tobytes()

    # ---- Comparator interface -------------------------------------------------------------
    def Compare(self):
        d = C.c_float()
        _check(lib().gzb_compare(self._ctx, C.byref(d)), self._ctx)
        self.distance = np.float32(d.value)
        return self.distance

    def distmap(self):
        out = np.zeros((self.height, self.width), np.float32)
        _check(lib().gzb_get_distmap(self._ctx, _p(out)), self._ctx)
        return out

    def distmap_aggregate(self):
        return self.distance

    def DistanceOK(self, target_mul):
        return bool(lib().gzb_distance_ok(self._ctx, float(target_mul)))

    def ScoreOutputSize(self, size):
        return lib().gzb_score_output_size(self._ctx, int(size))

    def BlockErrorLimit(self):
        return np.float32(lib().gzb_block_error_limit(self._ctx))

    def StartBlockComparisons(self):
        _check(lib().gzb_start_block_comparisons(self._ctx), self._ctx)

    def FinishBlockComparisons(self):
        _check(lib().gzb_finish_block_comparisons(self._ctx), self._ctx)

    def BlockLists(self, want_opsin=True):
        """(imgMaskXyzScaleBlockList [B,3], imgOpsinDynamicsBlockList [B,192])."""
        ms = np.zeros((self.num_blocks, 3), np.float32)
        ob = np.zeros((self.num_blocks, 192), np.float32) if want_opsin else None
        _check(lib().gzb_get_block_lists(self._ctx, _p(ms), _p(ob) if want_opsin else None), self._ctx)
        return ms, ob

    def CompareBlocks(self):
        """CompareBlock of every 8x8 block of the resident candidate (factor 1, comp_mask 7)."""
        out = np.zeros(self.num_blocks, np.float32)
        _check(lib().gzb_compare_blocks(self._ctx, _p(out)), self._ctx)
        return out

    def ComputeBlockZeroingOrder(self, comp_mask=7):
        out = np.zeros((self.ZeroingUnits(comp_mask), 192), COEFF_DATA)
        _check(lib().gzb_compute_block_zeroing_order(self._ctx, comp_mask, _p(out)), self._ctx)
        return out

    def ComputeBlockZeroingCandidates(self, comp_mask=7, block_begin=0, block_end=None):
        """(candidate_coeff_offsets[n+1], candidate_coeffs u8, candidate_coeff_errors f32) of the
        blocks [block_begin, block_end) (default: all)."""
        L = lib()
        L.gzb_compute_block_zeroing_candidates_range.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_size_t)]
        b1 = self.ZeroingUnits(comp_mask) if block_end is None else block_end
        nloc = b1 - block_begin
        off = np.zeros(nloc + 1, np.int32)
        cap = max(1, nloc * 192)
        idx = np.zeros(cap, np.uint8)
        err = np.zeros(cap, np.float32)
        n = C.c_size_t()
        _check(L.gzb_compute_block_zeroing_candidates_range(self._ctx, comp_mask, block_begin, b1, _p(off), _p(idx), _p(err),
                                                            cap, C.byref(n)), self._ctx)
        return off, idx[:n.value].copy(), err[:n.value].copy()

    def ComputeBlockErrorAdjustmentWeights(self, direction, max_block_dist, target_mul, distmap=None, factor=1):
        bs = 8 * factor
        w = np.zeros(((self.width + bs - 1) // bs) * ((self.height + bs - 1) // bs), np.float32)
        dm = None if distmap is None else np.ascontiguousarray(distmap, np.float32)
        L = lib()
        L.gzb_compute_block_error_adjustment_weights_f.argtypes = [
            C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_int, C.c_void_p, C.c_void_p]
        _check(L.gzb_compute_block_error_adjustment_weights_f(
            self._ctx, direction, max_block_dist, float(target_mul), int(factor),
            None if dm is None else _p(dm), _p(w)), self._ctx)
        return w

    # ---- debugging / measurement ----------------------------------------------------------
    def debug_fetch(self, name):
        n = C.c_size_t()
        _check(lib().gzb_debug_fetch(self._ctx, name.encode(), None, 0, C.byref(n)), self._ctx)
        out = np.zeros(n.value, np.float32)
        _check(lib().gzb_debug_fetch(self._ctx, name.encode(), _p(out), out.size, C.byref(n)), self._ctx)
        return out

    def last_device_ms(self):
        return float(lib().gzb_last_device_ms(self._ctx))

    def profile(self, on=True):
        profile_enable(self._ctx, on)

    def kernel_times(self):
        return profile_get(self._ctx)

    def launch_count(self):
        return int(lib().gzb_launch_count(self._ctx))

    def fine_bdm_compare_count(self):
        L = lib()
        L.gzb_fine_bdm_compare_count.restype = C.c_ulonglong
        return int(L.gzb_fine_bdm_compare_count(self._ctx))

    def incremental_compare_count(self):
        L = lib()
        L.gzb_incremental_compare_count.restype = C.c_ulonglong
        L.gzb_incremental_compare_count.argtypes = [C.c_void_p]
        return int(L.gzb_incremental_compare_count(self._ctx))


# ---- stage entry points (roles of the reference's cu* free functions) ----------------------
def OpsinDynamicsImage(planes, device=0):
    """cuOpsinDynamicsImage: planes [3,H,W] float32 linear rgb -> XYB (returns a new array)."""
    a = np.ascontiguousarray(planes, np.float32).copy()
    _, h, w = a.shape
    _check(lib().gzb_opsin_dynamics_image(device, _p(a[0]), _p(a[1]), _p(a[2]), w, h))
    return a


def DiffmapOpsinDynamicsImage(xyb0, xyb1, device=0):
    """cuDiffmapOpsinDynamicsImage: two [3,H,W] XYB images -> diffmap [H,W]."""
    a = np.ascontiguousarray(xyb0, np.float32)
    b = np.ascontiguousarray(xyb1, np.float32)
    _, h, w = a.shape
    out = np.zeros((h, w), np.float32)
    _check(lib().gzb_diffmap_opsin_dynamics_image(device, _p(out), _p(a[0]), _p(a[1]), _p(a[2]),
                                                  _p(b[0]), _p(b[1]), _p(b[2]), w, h, 3))
    return out


def Mask(xyb0, xyb1, device=0):
    """cuMask (butteraugli::Mask): two [3,H,W] XYB images -> (mask [3,H,W], mask_dc [3,H,W])."""
    a = np.ascontiguousarray(xyb0, np.float32)
    b = np.ascontiguousarray(xyb1, np.float32)
    _, h, w = a.shape
    m = np.zeros((3, h, w), np.float32)
    d = np.zeros((3, h, w), np.float32)
    L = lib()
    L.gzb_mask.argtypes = [C.c_int] + [C.c_void_p] * 6 + [C.c_size_t, C.c_size_t] + [C.c_void_p] * 6
    _check(L.gzb_mask(device, _p(m[0]), _p(m[1]), _p(m[2]), _p(d[0]), _p(d[1]), _p(d[2]), w, h,
                      _p(a[0]), _p(a[1]), _p(a[2]), _p(b[0]), _p(b[1]), _p(b[2])))
    return m, d


def Blur(plane, sigma, border_ratio=0.0, device=0):
    a = np.ascontiguousarray(plane, np.float32).copy()
    h, w = a.shape
    _check(lib().gzb_blur(device, _p(a), w, h, float(sigma), float(border_ratio)))
    return a


def ButteraugliSrgb(rgb0, rgb1, want_diffmap=True, device=0):
    a = np.ascontiguousarray(rgb0, np.uint8)
    b = np.ascontiguousarray(rgb1, np.uint8)
    h, w = a.shape[:2]
    d = C.c_float()
    dm = np.zeros((h, w), np.float32) if want_diffmap else None
    _check(lib().gzb_butteraugli_srgb(device, _p(a), _p(b), w, h, C.byref(d),
                                      _p(dm) if want_diffmap else None))
    return np.float32(d.value), dm


# ---- whole encoder and host-side pieces --------------------------------------------------------
class EncodeStats(C.Structure):
    _fields_ = [("num_iterations", C.c_int), ("num_iterations_up", C.c_int), ("num_iterations_down", C.c_int),
                ("num_compares", C.c_int), ("num_jpeg_writes", C.c_int), ("num_entropy_code_builds", C.c_int),
                ("total_wall_ms", C.c_double), ("host_frontend_ms", C.c_double), ("host_quant_ms", C.c_double),
                ("host_write_ms", C.c_double), ("compare_wall_ms", C.c_double), ("device_compare_ms", C.c_double),
                ("zeroing_wall_ms", C.c_double), ("device_zeroing_ms", C.c_double), ("backend_wall_ms", C.c_double),
                ("write_hist_ms", C.c_double), ("write_code_ms", C.c_double), ("write_encode_ms", C.c_double),
                ("write_stitch_ms", C.c_double), ("be_weights_ms", C.c_double), ("be_order_ms", C.c_double),
                ("be_walk_ms", C.c_double), ("be_update_ms", C.c_double), ("create_ms", C.c_double),
                ("be_codes_ms", C.c_double), ("be_sort_ms", C.c_double), ("be_steps", C.c_ulonglong),
                ("prepare_ms", C.c_double), ("run_ms", C.c_double), ("h2d_bytes", C.c_ulonglong), ("d2h_bytes", C.c_ulonglong),
                ("final_score", C.c_double), ("final_distance", C.c_float), ("launches", C.c_ulonglong),
                ("be_prefix_steps", C.c_ulonglong), ("device_write_ms", C.c_double), ("search_wall_ms", C.c_double), ("trial_host_ms", C.c_double),
                ("trial_device_ms", C.c_double), ("search_rounds", C.c_int), ("search_trials", C.c_int),
                ("downsample_ms", C.c_double), ("be_selects", C.c_ulonglong), ("be_levels", C.c_ulonglong),
                ("be_host_ranges", C.c_ulonglong), ("be_lazy_ms", C.c_double), ("num_fine_bdm_compares", C.c_ulonglong), ("be_select_ms", C.c_double), ("be_gather_ms", C.c_double),
                ("be_pool_ms", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def ButteraugliScoreForQuality(quality):
    L = lib()
    L.gzb_butteraugli_score_for_quality.restype = C.c_double
    L.gzb_butteraugli_score_for_quality.argtypes = [C.c_double]
    return L.gzb_butteraugli_score_for_quality(float(quality))


def MeasureFp64Peak(device=0):
    """Non-FMA FP64 peak of the device in Gflop/s (gzb_measure_fp64_peak)."""
    g = C.c_double()
    L = lib()
    L.gzb_measure_fp64_peak.argtypes = [C.c_int, C.POINTER(C.c_double)]
    _check(L.gzb_measure_fp64_peak(device, C.byref(g)))
    return float(g.value)


def Process(rgb, butteraugli_target, device=0, host_threads=0, want_trace=False, try_420=False, force_420=False):
    """guetzli::Process(params, stats, rgb, w, h, &out) on the B200 (try_420 / force_420: Params of
    guetzli/processor.h:34-42). Returns (jpeg_bytes, stats_dict, trace_or_None)."""
    L = lib()
    a = np.ascontiguousarray(rgb, np.uint8)
    h, w = a.shape[:2]
    out = C.c_void_p()
    n = C.c_size_t()
    st = EncodeStats()
    tr = C.c_void_p()
    L.gzb_encode_rgb_params.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int, C.c_int, C.c_int,
                                        C.POINTER(C.c_void_p), C.POINTER(C.c_size_t), C.POINTER(EncodeStats),
                                        C.POINTER(C.c_void_p)]
    L.gzb_encode_last_error.restype = C.c_char_p
    L.gzb_free.argtypes = [C.c_void_p]
    L.gzb_free.restype = None
    rc = L.gzb_encode_rgb_params(device, _p(a), w, h, C.c_float(butteraugli_target), int(try_420), int(force_420),
                                 host_threads, C.byref(out), C.byref(n), C.byref(st), C.byref(tr) if want_trace else None)
    if rc != 0:
        raise GzbError("gzb_encode_rgb failed (%d): %s" % (rc, L.gzb_encode_last_error().decode(errors="replace")))
    data = C.string_at(out, n.value)
    L.gzb_free(out)
    trace = None
    if want_trace and tr:
        trace = C.string_at(tr).decode(errors="replace")
        L.gzb_free(tr)
    return data, st.as_dict(), trace


def ProcessBatch(images, butteraugli_target, device=0, inflight=3, host_threads_per_encode=0, try_420=False, force_420=False):
    """gzb_encode_rgb_batch: the images (HxWx3 uint8 arrays) encoded on one GPU with `inflight` encodes
    running concurrently. Returns [(jpeg_bytes, stats_dict)] in input order."""
    L = lib()
    n = len(images)
    arrs = [np.ascontiguousarray(im, np.uint8) for im in images]
    ptrs = (C.c_void_p * n)(*[a.ctypes.data for a in arrs])
    ws = (C.c_int * n)(*[a.shape[1] for a in arrs])
    hs = (C.c_int * n)(*[a.shape[0] for a in arrs])
    outs = (C.c_void_p * n)()
    sizes = (C.c_size_t * n)()
    stats = (EncodeStats * n)()
    status = (C.c_int * n)()
    L.gzb_encode_rgb_batch.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_float, C.c_int, C.c_int,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gzb_encode_last_error.restype = C.c_char_p
    L.gzb_free.argtypes = [C.c_void_p]
    L.gzb_free.restype = None
    rc = L.gzb_encode_rgb_batch(device, n, ptrs, ws, hs, C.c_float(butteraugli_target), int(try_420), int(force_420),
                                int(inflight), int(host_threads_per_encode), outs, sizes, stats, status)
    res = []
    for i in range(n):
        if outs[i]:
            res.append((C.string_at(outs[i], sizes[i]), stats[i].as_dict()))
            L.gzb_free(outs[i])
        else:
            res.append((None, None))
    if rc != 0:
        raise GzbError("gzb_encode_rgb_batch failed (%d): %s" % (rc, L.gzb_encode_last_error().decode(errors="replace")))
    return res


def RgbToJpegCoeffs(rgb):
    """guetzli::EncodeRGBToJpeg with q=1 (host code): int16 [3, nblocks, 64]."""
    a = np.ascontiguousarray(rgb, np.uint8)
    h, w = a.shape[:2]
    nb = ((w + 7) // 8) * ((h + 7) // 8)
    c = np.zeros((3, nb, 64), np.int16)
    lib().gzb_rgb_to_jpeg_coeffs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    _check(lib().gzb_rgb_to_jpeg_coeffs(_p(a), w, h, _p(c[0]), _p(c[1]), _p(c[2])))
    return c


def WriteJpeg(coeffs, width, height, q, input_tables=False, host_threads=1):
    """OutputImage::SaveToJpegData + WriteJpeg (host code): JPEG bytes."""
    c = np.ascontiguousarray(coeffs, np.int16)
    qq = np.ascontiguousarray(q, np.int32).reshape(192)
    L = lib()
    L.gzb_write_jpeg.restype = C.c_long
    L.gzb_write_jpeg.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                 C.c_int, C.c_void_p, C.c_long]
    buf = np.zeros(width * height * 3 + (1 << 16), np.uint8)
    n = L.gzb_write_jpeg(_p(c[0]), _p(c[1]), _p(c[2]), width, height, _p(qq), int(input_tables), host_threads,
                         _p(buf), buf.size)
    if n < 0:
        raise GzbError("gzb_write_jpeg failed (%d)" % n)
    return buf[:n].tobytes()


class Encoder:
    """Two-step encoder (gzb_encoder_create / gzb_encoder_run): create() leaves the image, its
    opsin-dynamics image and the q=1 coefficients resident in HBM; run() performs the search."""

    def __init__(self, rgb, butteraugli_target, device=0, host_threads=0, profile=False, try_420=False, force_420=False):
        L = lib()
        a = np.ascontiguousarray(rgb, np.uint8)
        self.h, self.w = a.shape[:2]
        self._params = (int(try_420), int(force_420))
        L.gzb_encoder_create.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_int,
                                         C.POINTER(C.c_void_p)]
        L.gzb_encoder_run.argtypes = [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_size_t),
                                      C.POINTER(EncodeStats), C.POINTER(C.c_void_p)]
        L.gzb_encoder_context.restype = C.c_void_p
        L.gzb_encoder_context.argtypes = [C.c_void_p]
        L.gzb_encoder_destroy.argtypes = [C.c_void_p]
        L.gzb_encoder_destroy.restype = None
        L.gzb_encode_last_error.restype = C.c_char_p
        L.gzb_free.argtypes = [C.c_void_p]
        L.gzb_free.restype = None
        self._enc = C.c_void_p()
        rc = L.gzb_encoder_create(device, _p(a), self.w, self.h, C.c_float(butteraugli_target), host_threads,
                                  C.byref(self._enc))
        if rc != 0:
            raise GzbError("gzb_encoder_create failed (%d): %s" % (rc, L.gzb_encode_last_error().decode(errors="replace")))
        self._ctx = C.c_void_p(L.gzb_encoder_context(self._enc))
        L.gzb_encoder_set_params.argtypes = [C.c_void_p, C.c_int, C.c_int]
        L.gzb_encoder_set_params(self._enc, *self._params)
        if profile:
            profile_enable(self._ctx, True)

    def run(self, want_trace=False):
        L = lib()
        out = C.c_void_p(); n = C.c_size_t(); st = EncodeStats(); tr = C.c_void_p()
        rc = L.gzb_encoder_run(self._enc, C.byref(out), C.byref(n), C.byref(st), C.byref(tr) if want_trace else None)
        if rc != 0:
            raise GzbError("gzb_encoder_run failed (%d): %s" % (rc, L.gzb_encode_last_error().decode(errors="replace")))
        data = C.string_at(out, n.value)
        L.gzb_free(out)
        trace = None
        if want_trace and tr:
            trace = C.string_at(tr).decode(errors="replace")
            L.gzb_free(tr)
        return data, st.as_dict(), trace

    def set_group(self, rank, world, allgather, allgather_device=None):
        """gzb_encoder_set_group: this encoder becomes rank `rank` of `world` encoders of the SAME
        image (one per GPU). `allgather(send: bytes-like) -> bytes` must return the concatenation of
        every rank's buffer in rank order (see torch_allgather). `allgather_device(d_send, nbytes, d_recv)`
        (optional, gzb_encoder_set_group_device) all-gathers DEVICE memory given as integer pointers (see
        torch_allgather_device): the zeroing candidates then go from GPU to GPU."""
        self._cb = make_allgather_callback(allgather, world)   # keep the thunk alive
        L = lib()
        L.gzb_encoder_set_group.argtypes = [C.c_void_p, C.c_int, C.c_int, ALLGATHER_FN, C.c_void_p]
        rc = L.gzb_encoder_set_group(self._enc, rank, world, self._cb, None)
        if rc != 0:
            raise GzbError("gzb_encoder_set_group failed (%d): %s" % (rc, L.gzb_encode_last_error().decode(errors="replace")))
        if allgather_device is not None:
            def thunk(user, d_send, nbytes, d_recv):
                try:
                    allgather_device(int(d_send or 0), int(nbytes), int(d_recv or 0))
                    return 0
                except Exception:  # never let an exception cross the C boundary
                    import traceback
                    traceback.print_exc()
                    return -2
            self._cb_dev = ALLGATHER_FN(thunk)
            L.gzb_encoder_set_group_device.argtypes = [C.c_void_p, ALLGATHER_FN]
            rc = L.gzb_encoder_set_group_device(self._enc, self._cb_dev)
            if rc != 0:
                raise GzbError("gzb_encoder_set_group_device failed (%d): %s" % (rc, L.gzb_encode_last_error().decode(errors="replace")))

    def kernel_times(self):
        return profile_get(self._ctx)

    def close(self):
        if self._enc:
            lib().gzb_encoder_destroy(self._enc)
            self._enc = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


# ---- multi-GPU group plumbing (gzb_allgather_fn) -------------------------------------------------
ALLGATHER_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p)


def make_allgather_callback(allgather, world):
    def thunk(user, send, nbytes, recv):
        try:
            into = getattr(allgather, "into", None)
            if into is not None:   # no intermediate copies: numpy views of the two C buffers
                s = np.ctypeslib.as_array(C.cast(send, C.POINTER(C.c_uint8)), shape=(nbytes,))
                r = np.ctypeslib.as_array(C.cast(recv, C.POINTER(C.c_uint8)), shape=(world * nbytes,))
                into(s, r)
                return 0
            out = allgather(C.string_at(send, nbytes))
            if len(out) != world * nbytes:
                return -1
            C.memmove(recv, bytes(out), world * nbytes)
            return 0
        except Exception:  # never let an exception cross the C boundary
            import traceback
            traceback.print_exc()
            return -2
    return ALLGATHER_FN(thunk)


def torch_allgather(dist, device):
    """An allgather(bytes)->bytes over torch.distributed: NCCL over NVLink/NVSwitch when `device` is
    a CUDA device (the exchange buffers are staged through HBM), gloo for the CPU tests. Its `into(send, recv)`
    form works on numpy views of the caller's buffers (one copy in, one copy out)."""
    import torch
    world = dist.get_world_size()

    def allgather(buf):
        t = torch.frombuffer(bytearray(buf), dtype=torch.uint8).to(device)
        outs = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(outs, t)
        return b"".join(o.cpu().numpy().tobytes() for o in outs)

    def into(send, recv):
        t = torch.from_numpy(send).to(device)
        out = torch.empty((world, t.numel()), dtype=torch.uint8, device=device)
        dist.all_gather(list(out.unbind(0)), t)
        torch.from_numpy(recv).copy_(out.view(-1))
    allgather.into = into
    return allgather


def torch_allgather_device(dist, device):
    """gzb_allgather_device_fn over torch.distributed (NCCL): the two device pointers are wrapped as uint8
    tensors without a copy and all-gathered over NVLink / NVSwitch."""
    import torch
    world = dist.get_world_size()
    dev = torch.device("cuda", device) if isinstance(device, int) else device

    class _Ptr(object):   # a raw device pointer as a 1-D uint8 CUDA array
        def __init__(self, ptr, n):
            self.__cuda_array_interface__ = {"shape": (n,), "typestr": "|u1", "data": (ptr, False), "version": 2}

    def fn(d_send, nbytes, d_recv):
        with torch.cuda.device(dev):
            s = torch.as_tensor(_Ptr(d_send, nbytes), device=dev)
            r = torch.as_tensor(_Ptr(d_recv, nbytes * world), device=dev)
            dist.all_gather_into_tensor(r, s)
            torch.cuda.current_stream(dev).synchronize()
    return fn


def ProcessGroup(rgb, butteraugli_target, dist, device=0, host_threads=0, want_trace=False):
    """guetzli::Process of ONE image by all ranks of a torch.distributed group (one rank per GPU):
    SelectQuantMatrix candidates and zeroing-search blocks are sharded, rank 0 returns the JPEG
    (other ranks return b""). Returns (jpeg_bytes, stats_dict, trace_or_None)."""
    import torch
    dev = torch.device("cuda", device) if dist.get_backend() == "nccl" else torch.device("cpu")
    enc = Encoder(rgb, butteraugli_target, device=device, host_threads=host_threads)
    try:
        enc.set_group(dist.get_rank(), dist.get_world_size(), torch_allgather(dist, dev),
                      torch_allgather_device(dist, dev) if dev.type == "cuda" else None)
        return enc.run(want_trace)
    finally:
        enc.close()


def QuantSearchSimulate(rank, world, allgather, target, eval_fn, batch=1, mode=0):
    """Test hook (no GPU): runs the speculative SelectQuantMatrix search of gzb_quant_search.h with
    `eval_fn(original, q[192]) -> (distance, jpg_size)` standing in for the GPU trial. Returns
    {"visited": [(original, hscore, distance, size)], "best_q": [192], "best_ok": bool, "rounds": n,
     "evaluated_here": n, "evaluated_total": n}."""
    L = lib()
    EVAL = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint64))

    def ev(user, original, q, dist_out, size_out):
        r = eval_fn(bool(original), [q[i] for i in range(192)])
        if r is None:      # the stand-in for a failed GPU trial
            return 1
        dist_out[0], size_out[0] = r
        return 0
    ev_c = EVAL(ev)
    cb = make_allgather_callback(allgather, world) if world > 1 else ALLGATHER_FN(0)
    cap = 256
    vis = np.zeros((cap, 4), np.float64)
    nvis = C.c_int()
    best_q = np.zeros(192, np.int32)
    info = np.zeros(4, np.int32)
    L.gzb_test_quant_search_mode.argtypes = [C.c_int, C.c_int, ALLGATHER_FN, C.c_void_p, EVAL, C.c_void_p, C.c_float,
                                             C.c_void_p, C.c_int, C.POINTER(C.c_int), C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    rc = L.gzb_test_quant_search_mode(rank, world, cb, None, ev_c, None, C.c_float(target), _p(vis), cap, C.byref(nvis),
                                      _p(best_q), _p(info), batch, mode)
    if rc != 0:
        raise GzbError("gzb_test_quant_search failed (%d)" % rc)
    return {"visited": [tuple(v) for v in vis[:nvis.value].tolist()], "best_q": best_q.tolist(), "best_ok": bool(info[0]),
            "rounds": int(info[1]), "evaluated_here": int(info[2]), "evaluated_total": int(info[3])}


def profile_enable(ctx, on=True):
    L = lib()
    L.gzb_profile_enable.argtypes = [C.c_void_p, C.c_int]
    _check(L.gzb_profile_enable(ctx, 1 if on else 0), ctx)


def profile_get(ctx):
    """{kernel name: (accumulated device ms, launches)} from CUDA events on the context's stream."""
    L = lib()
    L.gzb_profile_name.restype = C.c_char_p
    L.gzb_profile_get.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_double), C.POINTER(C.c_ulonglong)]
    out = {}
    for i in range(L.gzb_profile_count()):
        ms = C.c_double(); n = C.c_ulonglong()
        _check(L.gzb_profile_get(ctx, i, C.byref(ms), C.byref(n)), ctx)
        if n.value:
            out[L.gzb_profile_name(i).decode()] = (ms.value, n.value)
    return out


def DctDouble(blocks, inverse=False, device=0):
    """ComputeBlockDCTDouble / ComputeBlockIDCTDouble on [n, 64] float64 blocks (returns a copy)."""
    a = np.ascontiguousarray(blocks, np.float64).copy().reshape(-1, 64)
    lib().gzb_dct_double.argtypes = [C.c_int, C.c_void_p, C.c_size_t, C.c_int]
    _check(lib().gzb_dct_double(device, _p(a), a.shape[0], 1 if inverse else 0))
    return a
