#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Guetzli hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size WxH] [--quality Q]

A "step" is one full guetzli::Process of one synthetic image per GPU (RGB -> q=1 coefficients ->
SelectQuantMatrix -> block-zeroing search -> back-end iterations -> best JPEG). Default workload: the
configuration the metric and the >=100x target are quoted on, 4000x3000 at quality 95 (BASELINE.json
configs[2]'s image on ONE B200). With N>1 (torchrun, one rank per GPU) every rank encodes its own image
of that size per step: weak scaling, no data-path collective (images are independent units).

  value  : MPix/s of gzb_encoder_run with the image, its XYB and its q=1 coefficients already
           resident in HBM (gzb_encoder_create is outside the timed region)
  e2e    : MPix/s of gzb_encode_rgb from a pinned host RGB buffer to host JPEG bytes (context creation,
           H2D of the image, every per-iteration copy and the D2H results inside the timed region)
  roofline : dominant kernel (k_zeroing_order), algorithmic bytes U2 = 3084 B per 8x8 block
            / its CUDA-event duration on the launching stream, vs MEASURED_PEAKS.json;
           the FP64 rate of the same kernel against a non-FMA FP64 peak measured in this run
  cpu_baseline : the unmodified reference (oracle/_ref) on one host core on a bounded crop of the same
           image, plus the stored single-core time of the WHOLE workload image (tests/golden/full_encodes.json)
  batch64 : BASELINE configs[3], 64 images of 1920x1080 at q95 split over the ranks, several encodes in
           flight per GPU (gzb_encode_rgb_batch) -- extra key, separately timed
  group  : BASELINE configs[2] at N>1, the 12 MPix image encoded ONCE by all ranks together
           (SelectQuantMatrix candidates + zeroing blocks sharded, NCCL all-gather) -- extra key

--impl reference runs the reference's own CPU encoder on all host cores at the same quality on crops
of the same workload image (one crop per core per step) and prints the same JSON shape.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

ALGO_BYTES_PER_BLOCK = 3084      # U2: 2x384 coeffs + 768 original XYB + 12 mask scale + 1536 out
ALGO_BYTES_COMPARE_PER_PX = 22   # U1: 6 coeffs + 12 cached XYB + 4 diffmap
FALLBACK_HBM_GBS = 6650.0


def ncu_evidence(kernel, w, h):
    """Per-launch ncu counters of `kernel` from the committed `ncu --set full` capture of THIS workload
    (profiles/r2_ncu_traffic.json holds one entry per image size): dram bytes, FP64-pipe utilisation and
    the double-precision instruction counts the FP64 rate is computed from."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r2_ncu_traffic.json")))
        k = d["sizes"]["%dx%d" % (w, h)]["kernels"][kernel]
        return k, d["sizes"]["%dx%d" % (w, h)].get("source")
    except Exception:
        return None, None


def peaks():
    try:
        d = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return FALLBACK_HBM_GBS, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc, self.mark_at = index, [], None, 0

    def mark(self):
        """Rows from here on belong to the timed region (the sampler is started earlier, during warm-up,
        so that nvidia-smi's start-up time does not eat a short timed region)."""
        self.mark_at = len(self.rows)

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        rows = self.rows[self.mark_at:] if len(self.rows) > self.mark_at else self.rows
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7),
                              ("sw_power_cap", 8)):
                if len(r) > col and r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


_REAL_STDOUT = None


def guard_stdout():
    """The contract is ONE JSON line on stdout. Libraries (NCCL's version banner, torchrun notices) also
    write to file descriptor 1, so it is pointed at stderr for the whole run and the line goes to a saved
    duplicate of the original stdout."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(text):
    data = (text + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(text + "\n")
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


# ------------------------------------------------------------------------------------------------
# the workload image and bounded samples of it
# ------------------------------------------------------------------------------------------------
CROP = 320   # side of the crops the CPU reference is timed on (a whole 12 MPix encode takes half an hour per core)


def workload_image(w, h, seed):
    import _libs
    return _libs.synth_image(w, h, seed)


def crop_origins(w, h, count, first=0):
    """`count` crop origins spread over the image on a grid (deterministic), starting at index `first`."""
    nx, ny = max(1, w // CROP), max(1, h // CROP)
    out = []
    for i in range(first, first + count):
        j = (i * 7919) % (nx * ny)
        out.append(((j % nx) * CROP, (j // nx) * CROP))
    return out


def stored_full_reference(w, h, quality, seed):
    """Single-core time of the unmodified reference on the WHOLE workload image, measured once in the build
    container and committed with the golden bytes (tests/golden/full_encodes.json, make_full_golden.py)."""
    try:
        d = json.load(open(os.path.join(ROOT, "tests", "golden", "full_encodes.json")))
        g = d["%dx%d_q%d_s%d" % (w, h, int(quality), seed)]
        return g
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------
# reference arm: the unmodified reference CPU encoder on all host cores
# ------------------------------------------------------------------------------------------------
_REF_IMG = None


def _ref_worker(args):
    x0, y0, target = args
    import _libs
    crop = np.ascontiguousarray(_REF_IMG[y0:y0 + CROP, x0:x0 + CROP])
    t0 = time.perf_counter()
    jpg, iters, _ = _libs.ref_process(crop, target)
    return time.perf_counter() - t0, len(jpg), iters


def run_reference(a, rank, world):
    global _REF_IMG
    if rank != 0:
        return
    import multiprocessing as mp
    import _libs
    if not _libs.have_ref():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgzref.so not built"}))
        return
    cores = os.cpu_count() or 1
    w, h = a.size
    # bounded sample of the workload: crops of the SAME image at the SAME quality, one per host core per
    # step (the reference has no threading; its own test runs one process per core, tests/golden_test.sh:25)
    _REF_IMG = workload_image(w, h, 1234)
    target = float(np.float32(_libs.ref().ref_butteraugli_score_for_quality(float(a.quality))))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for step in range(a.warmup + a.steps):
            jobs = [(x0, y0, target) for (x0, y0) in crop_origins(w, h, cores, step * cores)]
            t0 = time.perf_counter()
            pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= a.warmup:
                times.append(dt)
    mpix_step = cores * CROP * CROP / 1e6
    total = sum(times)
    value = mpix_step * len(times) / total
    full = stored_full_reference(w, h, a.quality, 1234)
    sample = ("%d crops of %dx%d of the synthetic %dx%d workload image per step, one per host core, quality %g"
              % (cores, CROP, CROP, w, h, a.quality))
    line = {
        "impl": "reference", "metric": "end-to-end encode MPix/s", "value": value, "unit": "MPix/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %dx%d sRGB (SURVEY 8d generator, seed 1234), quality %g, guetzli::Process of the "
                               "reference on the host CPU; bounded sample: %s" % (w, h, a.quality, sample)},
        "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": cores, "kind": "reference", "sample": sample,
                         "full_workload_one_core": None if not full else {
                             "seconds": full["reference_seconds_one_core"], "value": w * h / 1e6 / full["reference_seconds_one_core"],
                             "unit": "MPix/s", "note": "the whole workload image on one core, measured in the build container "
                                                       "(tests/golden/full_encodes.json): per-pixel cost grows with image size"}},
        "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# speed-only comparator: the reference's OWN CUDA path (guetzli --cuda) on the same GPU
# ------------------------------------------------------------------------------------------------
def run_reference_cuda(a, rank, world):
    """`--impl reference-cuda`: the reference built with -D__USE_CUDA__ (oracle/Makefile: refcuda), its kernels JIT-ed
    from PTX for compute_100, g_mathMode = MODE_CUDA -- Compare and the zeroing search on the GPU, everything
    else on one host thread, float-only arithmetic (its bytes differ from the CPU encoder's: never a parity
    reference). One 1024x1024 crop of the workload image per step: a whole 12 MPix encode takes it minutes."""
    if rank != 0:
        return
    import ctypes as C
    import _libs
    base = os.path.join(ROOT, "oracle", "_ref", "cuda")
    so = os.path.join(base, "libgzref_cuda.so")
    if not os.path.exists(so) or not os.path.exists(os.path.join(base, "clguetzli", "clguetzli.cu.ptx64")):
        emit(json.dumps({"impl": "reference-cuda", "unavailable": "oracle/_ref/cuda not built (make -C oracle refcuda)"}))
        return
    w, h = a.size
    side = min(1024, w, h)
    img = workload_image(w, h, 1234)
    crop = np.ascontiguousarray(img[:side, :side])
    os.chdir(base)   # ocu.cpp:44 reads clguetzli/clguetzli.cu.ptx64 relative to the working directory
    L = C.CDLL(so)
    L.ref_butteraugli_score_for_quality.restype = C.c_double
    L.ref_butteraugli_score_for_quality.argtypes = [C.c_double]
    L.ref_process_rgb.restype = C.c_long
    target = float(np.float32(L.ref_butteraugli_score_for_quality(float(a.quality))))
    L.ref_set_math_mode(3)   # MODE_CUDA (clguetzli/clguetzli.h:17-25)
    out = np.zeros(side * side * 3 + (1 << 16), np.uint8)
    iters = C.c_int()
    times, sizes = [], []
    for step in range(a.warmup + a.steps):
        t0 = time.perf_counter()
        n = L.ref_process_rgb(_libs.p(crop), side, side, C.c_float(target), _libs.p(out), C.c_long(out.size), None,
                              C.c_long(0), C.byref(iters))
        dt = time.perf_counter() - t0
        if n <= 0:
            emit(json.dumps({"impl": "reference-cuda", "unavailable": "guetzli::Process failed in MODE_CUDA (%d)" % n}))
            return
        if step >= a.warmup:
            times.append(dt); sizes.append(int(n))
    value = side * side / 1e6 * len(times) / sum(times)
    emit(json.dumps({
        "impl": "reference-cuda", "metric": "end-to-end encode MPix/s", "value": value, "unit": "MPix/s", "n_gpus": 1,
        "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sum(times) / len(times), "higher_is_better": True,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": "the reference's own --cuda path (kernels of clguetzli/clguetzli.cu JIT-ed for compute_100, host search on "
                               "one thread) on a %dx%d crop of the synthetic %dx%d workload image, quality %g; float-only arithmetic, "
                               "output differs from the CPU encoder's (speed comparison only)" % (side, side, w, h, a.quality),
                   "iterations": iters.value, "bytes": sizes[-1]}}))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def cpu_baseline_single_core(img, w, h, quality):
    """The unmodified reference on ONE core on a bounded sample: one CROP x CROP crop of the workload image."""
    import _libs
    if not _libs.have_ref():
        return None
    x0, y0 = crop_origins(w, h, 1)[0]
    crop = np.ascontiguousarray(img[y0:y0 + CROP, x0:x0 + CROP])
    target = float(np.float32(_libs.ref().ref_butteraugli_score_for_quality(float(quality))))
    t0 = time.perf_counter()
    jpg, iters, _ = _libs.ref_process(crop, target)
    dt = time.perf_counter() - t0
    full = stored_full_reference(w, h, quality, 1234)
    return {"value": crop.shape[0] * crop.shape[1] / 1e6 / dt, "unit": "MPix/s", "cores": 1, "kind": "reference",
            "sample": "one %dx%d crop of the workload image, quality %g, guetzli::Process single-threaded: %.1f s, %d iterations"
                      % (crop.shape[1], crop.shape[0], quality, dt, iters),
            "full_workload_one_core": None if not full else {
                "seconds": full["reference_seconds_one_core"], "value": w * h / 1e6 / full["reference_seconds_one_core"], "unit": "MPix/s",
                "note": "the whole workload image (same generator and seed) on one core of the build container, stored with "
                        "the golden bytes in tests/golden/full_encodes.json"}}


def pinned_copy(torch, img):
    """The image in page-locked host memory (the e2e arm copies its input from there)."""
    t = torch.from_numpy(np.ascontiguousarray(img)).pin_memory()
    return t, t.numpy()


def run_ours(a, rank, world, local):
    import hashlib
    import torch
    import __graft_entry__ as ge
    import _libs
    gz = ge.load_package()
    if not torch.cuda.is_available() or gz.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        # rank 0 prints exactly one line on stdout: keep NCCL's version banner off it
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        import torch.distributed as dist_mod
        dist = dist_mod
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w, h = a.size
    target = np.float32(gz.ButteraugliScoreForQuality(a.quality))
    cores = os.cpu_count() or 1
    host_threads = min(16, max(1, cores // max(1, world)))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def reduce_max(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(x):
        if dist is None:
            return x
        t = torch.tensor([float(x)], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    # One image per rank, the same at every step: nothing is carried from one encode to the next except the
    # allocator's cached slab, and L2 is flushed in between. Weak scaling needs the SAME work on every GPU, so all
    # ranks encode the seed-1234 image of the SURVEY 8d generator (whose bytes are pinned to the reference's).
    # --distinct-images gives rank r the image of seed 1234 + 10 r instead: the time of an encode then depends on
    # the image (the observable tail of the back-end walk is 135 000 to 800 000 steps on these eight), and the
    # line's max-over-ranks is the slowest image's; its `per_rank` shows them. `batch64` always uses distinct images.
    seed = 1234 + (10 * rank if a.distinct_images else 0)
    pin_t, img = pinned_copy(torch, workload_image(w, h, seed))
    sampler = ClockSampler(local)
    run_ms, e2e_ms, launches, h2d, d2h = [], [], 0, 0, 0
    z_ms_sum = cmp_ms_sum = 0.0
    n_cmp = 0
    last_st = None
    jpg = b""
    for step in range(a.warmup + a.steps):
        timed = step >= a.warmup
        if step == 0:
            sampler.start()
        if timed and step == a.warmup:
            sampler.mark()
        # ---- device-resident arm: create outside, run inside the timed region
        enc = gz.Encoder(img, target, device=local, host_threads=host_threads, profile=False)
        flush.fill_(step & 0xff)
        barrier()
        t0 = time.perf_counter()
        jpg, st, _ = enc.run()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if timed:
            run_ms.append(dt)
            launches += st["launches"]
            n_cmp += st["num_compares"]
            # the dominant kernel and the Compare pipeline are bracketed by CUDA events on their own
            # stream inside the library on every step (no per-launch profiling in the timed region)
            z_ms_sum += st["device_zeroing_ms"]
            cmp_ms_sum += st["device_compare_ms"]
            last_st = st
        enc.close()
        # ---- end-to-end arm: pinned host RGB buffer -> host JPEG bytes
        flush.fill_((step + 1) & 0xff)
        barrier()
        t0 = time.perf_counter()
        jpg2, st2, _ = gz.Process(img, target, device=local, host_threads=host_threads)
        dt2 = (time.perf_counter() - t0) * 1e3
        assert jpg2 == jpg
        if timed:
            e2e_ms.append(dt2)
            h2d += st2["h2d_bytes"]; d2h += st2["d2h_bytes"]
    clocks = sampler.stop()
    total_run = reduce_max(sum(run_ms))
    total_e2e = reduce_max(sum(e2e_ms))
    per_rank = None
    if dist is not None:   # every rank's own mean step time and iteration count (images differ by rank)
        mine = torch.tensor([sum(run_ms) / len(run_ms), float(last_st["num_iterations"]), float(last_st["be_steps"] - last_st["be_prefix_steps"])],
                            dtype=torch.float64, device="cuda")
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = [{"rank": r, "ms_per_step": float(t[0]), "iterations": int(t[1]), "walked_steps": int(t[2])} for r, t in enumerate(allr)]
    launches_all = int(reduce_sum(launches))
    mpix = w * h / 1e6
    value = world * mpix * a.steps / (total_run / 1e3)
    e2e_value = world * mpix * a.steps / (total_e2e / 1e3)

    # ---- extra: BASELINE configs[3], a fixed batch of 64 images of 1920x1080 at q95 split over the ranks,
    # three encodes in flight per GPU (one encode's host phases overlap the others' kernels)
    batch64 = None
    if not a.no_extras:
        bw_, bh_, NIMG, KC = 1920, 1080, 64, 3
        bt = np.float32(gz.ButteraugliScoreForQuality(95))
        distinct = [workload_image(bw_, bh_, 1234 + 10 * k) for k in range(8)]   # image i of the batch = distinct[i % 8]
        mine = [distinct[i % 8] for i in range(NIMG) if i % world == rank]
        ht = max(1, host_threads // KC)
        gz.ProcessBatch(mine[:KC], bt, device=local, inflight=KC, host_threads_per_encode=ht)   # warm-up: slabs, pools
        flush.fill_(7)
        barrier()
        t0 = time.perf_counter()
        res = gz.ProcessBatch(mine, bt, device=local, inflight=KC, host_threads_per_encode=ht)
        torch.cuda.synchronize()
        bdt = reduce_max(time.perf_counter() - t0)
        batch64 = {"workload": "BASELINE configs[3]: 64 synthetic 1920x1080 images (8 distinct, seeds 1234+10k, cycled), quality 95, "
                               "split round-robin over %d GPU(s), gzb_encode_rgb_batch from host buffers" % world,
                   "value": NIMG * bw_ * bh_ / 1e6 / bdt, "unit": "MPix/s", "seconds": bdt, "inflight_per_gpu": KC,
                   "host_threads_per_encode": ht, "images_per_gpu": len(mine), "scaling": "strong",
                   "bytes_first_image": len(res[0][0])}

    # ---- extra (N > 1): BASELINE configs[2], ONE image encoded by all ranks together
    group = None
    if world > 1 and not a.no_extras:
        gimg = workload_image(w, h, 1234) if rank != 0 else img
        ght = min(16, max(1, cores - 2 * (world - 1))) if rank == 0 else 2   # the sequential back end runs on rank 0
        gtimes = []
        gj = b""
        for it in range(3):
            flush.fill_(it)
            barrier()
            t0 = time.perf_counter()
            gj, gst, _ = gz.ProcessGroup(gimg, target, dist, device=local, host_threads=ght)
            torch.cuda.synchronize()
            gdt = reduce_max(time.perf_counter() - t0)
            if it >= 1:
                gtimes.append(gdt)
        group = {"workload": "BASELINE configs[2]: the %dx%d q%g image encoded ONCE by %d GPUs (SelectQuantMatrix candidates and "
                             "zeroing blocks sharded, NCCL all-gather per round; back end on rank 0)" % (w, h, a.quality, world),
                 "value": mpix / (sum(gtimes) / len(gtimes)), "unit": "MPix/s", "seconds": sum(gtimes) / len(gtimes), "scaling": "strong",
                 "bytes_equal_single_gpu": (gj == jpg) if rank == 0 else None,
                 "rank0_phases_ms": {k: round(float(gst[k]), 2) for k in ("search_wall_ms", "zeroing_wall_ms", "device_zeroing_ms", "backend_wall_ms",
                                                                         "compare_wall_ms", "be_walk_ms", "be_order_ms", "create_ms", "prepare_ms", "run_ms", "total_wall_ms") if k in gst}}

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    nblocks = ((w + 7) // 8) * ((h + 7) // 8)
    z_ms, z_n = z_ms_sum, a.steps          # one zeroing launch per encode
    step_ms = total_run / a.steps
    roof = None
    if z_n:
        achieved = ALGO_BYTES_PER_BLOCK * nblocks / (z_ms / z_n / 1e3) / 1e9
        ev, ncu_src = ncu_evidence("k_zeroing_order", w, h)
        fp64 = None
        try:
            pk = gz.MeasureFp64Peak(local)
            fp64 = {"peak_gflops_non_fma": pk, "peak_source": "measured in this run: dependent-free DADD/DMUL streams, all SMs "
                                                              "(gzb_measure_fp64_peak)"}
            if ev and ev.get("fp64_flop_per_launch"):
                fp64["achieved_gflops"] = ev["fp64_flop_per_launch"] / (z_ms / z_n / 1e3) / 1e9
                fp64["frac"] = fp64["achieved_gflops"] / pk
        except Exception as ex:   # noqa
            fp64 = {"error": str(ex)}
        roof = {"bound": "hbm", "kernel": "k_zeroing_order", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": ev.get("dram_bytes_per_launch") if ev else None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_BLOCK * nblocks,
                "avg_launch_ms": z_ms / z_n, "share_of_step": (z_ms / z_n) / step_ms,
                "fp64_pipe_pct_of_peak": ev.get("fp64_pipe_pct") if ev else None, "fp64": fp64, "ncu_source": ncu_src,
                "note": "FP64-pipe/latency-bound search kernel (SURVEY 8d): ~126 CompareBlock trials per 8x8 block in "
                        "double precision; its U2 bytes are touched once, so the HBM fraction is small by construction "
                        "and the FP64 rate against the measured non-FMA peak is the meaningful ceiling"}
    cmp_ms = cmp_ms_sum
    st = last_st
    phases = {k: st[k] for k in ("search_wall_ms", "zeroing_wall_ms", "backend_wall_ms", "compare_wall_ms", "device_compare_ms",
                                 "device_zeroing_ms", "be_order_ms", "be_walk_ms", "be_select_ms", "be_gather_ms", "be_codes_ms",
                                 "be_pool_ms", "be_update_ms", "device_write_ms", "run_ms")}
    phases["note"] = ("wall-clock partition of the last timed step inside gzb_encoder_run: run = search + zeroing + backend (+ small "
                      "rest); backend = compare_wall + be_order + be_walk + be_update + coding waits; be_walk contains be_select, "
                      "be_gather and be_codes (which contains be_pool); device_* are CUDA-event times on the launching stream")
    gold = stored_full_reference(w, h, a.quality, seed)
    parity = {"sha256": hashlib.sha256(jpg).hexdigest(), "bytes": len(jpg)}
    if gold:
        parity["equals_reference_bytes"] = parity["sha256"] == gold["sha256"]
        parity["reference"] = "tests/golden/full_encodes.json (unmodified reference, single thread, same image)"
    line = {
        "metric": "end-to-end encode MPix/s", "value": value, "unit": "MPix/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %dx%d sRGB (SURVEY 8d generator, seed %s), quality %g, one image per GPU per "
                               "step, guetzli::Process" % (w, h, "1234+10*rank" if a.distinct_images else "1234 on every rank", a.quality),
                   "l2": "256 MiB buffer written between timed steps (L2 flush); the same image at every step",
                   "host_threads_per_encode": host_threads, "compares_per_step": n_cmp / a.steps,
                   "iterations_per_step": st["num_iterations"], "be_steps_per_step": st["be_steps"],
                   "be_prefix_steps_per_step": st["be_prefix_steps"]},
        "e2e": {"value": e2e_value, "unit": "MPix/s", "ms_per_step": total_e2e / a.steps,
                "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps},
        "gpu_launches": launches_all,
        "clocks": clocks,
        "roofline": roof,
        "butteraugli": {"note": "averaged over the Compares of an encode: %d of %d recompute BlockDiffMap only around the blocks flipped since "
                                "the previous Compare; a full Compare is what bench.py --mode butteraugli times" % (st["num_fine_bdm_compares"], st["num_compares"]),
                        "compare_device_ms_per_call": cmp_ms / max(1, n_cmp), "mpix_per_s": mpix / (cmp_ms / max(1, n_cmp) / 1e3) if cmp_ms else None,
                        "hbm_frac_U1": (ALGO_BYTES_COMPARE_PER_PX * w * h / (cmp_ms / max(1, n_cmp) / 1e3) / 1e9 / peak) if cmp_ms else None},
        "phases_ms": phases,
        "per_rank": per_rank,
        "parity": parity,
        "batch64": batch64,
        "group": group,
        "cpu_baseline": cpu_baseline_single_core(img, w, h, a.quality) if world == 1 and not a.no_cpu_baseline else None,
    }
    emit(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


def run_butteraugli_sweep(a, local):
    """BASELINE configs[4]: standalone butteraugli Compare over 0.25-24 MPix (M2, SURVEY 8d: the
    candidate is the image after ApplyGlobalQuantization with the all-3 matrix, resident as
    coefficients; diffmap + distance produced on the device) and an encode quality sweep 84-100 on a
    3840x2160 image. One GPU. Prints one JSON line."""
    import torch
    import __graft_entry__ as ge
    import _libs
    gz = ge.load_package()
    if not torch.cuda.is_available() or gz.device_count() == 0:
        raise SystemExit("bench.py: no CUDA device; the product has no CPU fallback")
    torch.cuda.set_device(local)
    peak, peak_src = peaks()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    sweep = []
    for (w, h) in [(512, 512), (1024, 1024), (2048, 2048), (4000, 3000), (6000, 4000)]:
        img = _libs.synth_image(w, h, 1234)
        c = gz.ButteraugliComparator(w, h, img, 1.0, device=local)
        c.SetJpegCoeffs(gz.RgbToJpegCoeffs(img))
        c.CopyFromJpegData()
        c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
        dev_ms, wall_ms = [], []
        for it in range(a.warmup + a.steps):
            # a Compare that follows another one of the SAME candidate would only recompute what changed (nothing);
            # re-rendering the candidate makes the next Compare a full one, as for a new candidate
            c.CopyFromJpegData()
            c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
            flush.fill_(it & 0xff)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            d = c.Compare()
            wall = (time.perf_counter() - t0) * 1e3
            if it >= a.warmup:
                dev_ms.append(c.last_device_ms()); wall_ms.append(wall)
        ms = sum(dev_ms) / len(dev_ms)
        sweep.append({"size": "%dx%d" % (w, h), "mpix": w * h / 1e6, "compare_device_ms": ms,
                      "compare_wall_ms": sum(wall_ms) / len(wall_ms), "mpix_per_s": w * h / 1e6 / (ms / 1e3),
                      "hbm_frac_U1": ALGO_BYTES_COMPARE_PER_PX * w * h / (ms / 1e3) / 1e9 / peak, "distance": float(d)})
        c.close()
    qsweep = []
    w, h = 3840, 2160
    img = _libs.synth_image(w, h, 1234)
    for q in range(84, 101, 2):
        target = np.float32(gz.ButteraugliScoreForQuality(q))
        t0 = time.perf_counter()
        jpg, st, _ = gz.Process(img, target, device=local, host_threads=min(16, os.cpu_count() or 1))
        dt = time.perf_counter() - t0
        qsweep.append({"quality": q, "target": float(target), "encode_s": dt, "mpix_per_s": w * h / 1e6 / dt,
                       "bytes": len(jpg), "iterations": st["num_iterations"], "compares": st["num_compares"],
                       "device_compare_ms": st["device_compare_ms"], "device_zeroing_ms": st["device_zeroing_ms"]})
    big = sweep[3]
    emit(json.dumps({"metric": "butteraugli MPix/s", "value": big["mpix_per_s"], "unit": "MPix/s", "n_gpus": 1,
                      "steps": a.steps, "warmup": a.warmup, "ms_per_step": big["compare_device_ms"],
                      "higher_is_better": True, "dtype": "f64", "data": "synthetic",
                      "config": {"workload": "standalone butteraugli Compare of the all-3-quantised candidate vs the original, "
                                             "value quoted at 4000x3000; l2 flushed between calls", "mode": "butteraugli"},
                      "peak_hbm_gbs": peak, "peak_source": peak_src, "algorithmic_bytes_per_px": ALGO_BYTES_COMPARE_PER_PX,
                      "sweep": sweep, "quality_sweep_3840x2160": qsweep}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference", "reference-cuda"])
    ap.add_argument("--size", default="4000x3000", type=lambda s: tuple(int(v) for v in s.lower().split("x")))
    ap.add_argument("--quality", type=float, default=95.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--distinct-images", action="store_true", help="N > 1: a different image per rank (seed 1234 + 10 * rank)")
    ap.add_argument("--no-extras", action="store_true", help="skip the batch64 (configs[3]) and group (configs[2]) measurements")
    ap.add_argument("--mode", default="batch", choices=["batch", "butteraugli"],
                    help="batch: one image per GPU per step (the headline); butteraugli: standalone Compare sweep + "
                         "quality sweep (configs[4])")
    a = ap.parse_args()
    if os.environ.get("GZB_BENCH_WATCHDOG"):   # debugging aid: dump the Python stacks if the run takes this many seconds
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["GZB_BENCH_WATCHDOG"]), repeat=False, file=sys.stderr)
    guard_stdout()
    rank, world, local = dist_env()
    if a.impl == "reference":
        run_reference(a, rank, world)
    elif a.impl == "reference-cuda":
        run_reference_cuda(a, rank, world)
    elif a.mode == "butteraugli":
        if rank == 0:
            run_butteraugli_sweep(a, local)
    else:
        run_ours(a, rank, world, local)


if __name__ == "__main__":
    main()
