#!/usr/bin/env python
"""bench.py -- headline benchmark of the B200 Guetzli hot path (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--size WxH] [--quality Q]

A "step" is one full guetzli::Process of one synthetic image per GPU (RGB -> q=1 coefficients ->
SelectQuantMatrix -> block-zeroing search -> back-end iterations -> best JPEG). Defaults: N=1,
1024x1024 at quality 90 (BASELINE.json configs[1]). With N>1 (launched by torchrun, one rank per
GPU) every rank encodes its own image of the batch: weak scaling, no data-path collective.

  value  : MPix/s of gzb_encoder_run with the image, its XYB and its q=1 coefficients already
           resident in HBM (gzb_encoder_create is outside the timed region)
  e2e    : MPix/s of gzb_encode_rgb from a host RGB buffer to host JPEG bytes (context creation,
           H2D of the image, every per-iteration copy and the D2H results inside the timed region)
  roofline : dominant kernel (k_zeroing_order), algorithmic bytes U2 = 3084 B per 8x8 block
            / its CUDA-event duration on the launching stream, vs MEASURED_PEAKS.json
  cpu_baseline : the unmodified reference (oracle/_ref) on one host core on a bounded sample

--impl reference runs the reference's own CPU encoder on all host cores (one image per core, the
reference's own batching method, tests/golden_test.sh:25) and prints the same JSON shape.
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
    """dram bytes per launch / FP64-pipe utilisation of `kernel` from the committed ncu --set full
    capture (profiles/r1_ncu_traffic.json, taken at 1024x1024). Only returned for that size."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "r1_ncu_traffic.json")))
        k = d["kernels"][kernel]
        if (w, h) != (1024, 1024):
            return None, None, d["source"]
        return k["dram_bytes_per_launch"], k["fp64_pipe_pct"], d["source"]
    except Exception:
        return None, None, None


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
# reference arm: the unmodified reference CPU encoder on all host cores
# ------------------------------------------------------------------------------------------------
def _ref_worker(args):
    w, h, seed, target = args
    import _libs
    img = _libs.synth_image(w, h, seed)
    t0 = time.perf_counter()
    jpg, iters, _ = _libs.ref_process(img, target)
    return time.perf_counter() - t0, len(jpg), iters


def run_reference(a, rank, world):
    if rank != 0:
        return
    import multiprocessing as mp
    import _libs
    if not _libs.have_ref():
        emit(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libgzref.so not built"}))
        return
    cores = os.cpu_count() or 1
    # bounded sample of the workload: one 256x256 crop-sized synthetic image per core per step
    sw, sh = 256, 256
    target = float(np.float32(_libs.ref().ref_butteraugli_score_for_quality(float(a.quality))))
    ctx = mp.get_context("fork")
    times = []
    with ctx.Pool(cores) as pool:
        for step in range(a.warmup + a.steps):
            jobs = [(sw, sh, 1234 + 10 * (step * cores + k), target) for k in range(cores)]
            t0 = time.perf_counter()
            pool.map(_ref_worker, jobs)
            dt = time.perf_counter() - t0
            if step >= a.warmup:
                times.append(dt)
    mpix_step = cores * sw * sh / 1e6
    total = sum(times)
    value = mpix_step * len(times) / total
    line = {
        "impl": "reference", "metric": "end-to-end encode MPix/s", "value": value, "unit": "MPix/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "reference CPU guetzli::Process, quality %g, %d images of %dx%d per step (one per host core)"
                               % (a.quality, cores, sw, sh)},
        "cpu_baseline": {"value": value, "unit": "MPix/s", "cores": cores, "kind": "reference",
                         "sample": "%d synthetic %dx%d images per step, one per core, quality %g" % (cores, sw, sh, a.quality)},
        "e2e": {"value": value, "unit": "MPix/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(json.dumps(line))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def cpu_baseline_single_core(quality):
    """The unmodified reference on ONE core on a bounded sample (384x384, same generator)."""
    import _libs
    if not _libs.have_ref():
        return None
    sw, sh = 384, 384
    img = _libs.synth_image(sw, sh, 1234)
    target = float(np.float32(_libs.ref().ref_butteraugli_score_for_quality(float(quality))))
    t0 = time.perf_counter()
    jpg, iters, _ = _libs.ref_process(img, target)
    dt = time.perf_counter() - t0
    return {"value": sw * sh / 1e6 / dt, "unit": "MPix/s", "cores": 1, "kind": "reference",
            "sample": "one synthetic %dx%d image, quality %g, guetzli::Process single-threaded: %.1f s, %d iterations"
                      % (sw, sh, quality, dt, iters)}


def run_ours(a, rank, world, local):
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
    if a.mode == "group" and world > 1:
        # one image: the sequential back end runs on rank 0 alone, the other ranks only drive their GPU
        host_threads = min(16, max(1, cores - 2 * (world - 1))) if rank == 0 else 2
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    group = a.mode == "group"

    def one_image(step):
        # batch mode: every rank its own image; group mode: all ranks share the step's image
        return _libs.synth_image(w, h, 1234 + 10 * (step if group else step * world + rank))

    allgather = None
    if group and world > 1:
        allgather = gz.torch_allgather(dist, torch.device("cuda", local))   # NCCL over NVLink

    K = max(1, a.inflight) if not group else 1
    if K > 1:
        host_threads = max(1, host_threads // K)
    images = [one_image(s) for s in range((a.warmup + a.steps) * K)]
    sampler = ClockSampler(local)
    run_ms, e2e_ms, launches, h2d, d2h = [], [], 0, 0, 0
    z_ms_sum = cmp_ms_sum = 0.0
    rounds, trials = [], []
    kt = {}
    n_cmp = 0
    for step in range(a.warmup + a.steps):
        timed = step >= a.warmup
        if step == 0:
            sampler.start()
        if timed and step == a.warmup:
            sampler.mark()
        img = images[step * K]
        # ---- device-resident arm: create outside, run inside the timed region
        encs = [gz.Encoder(images[step * K + k], target, device=local, host_threads=host_threads, profile=False) for k in range(K)]
        enc = encs[0]
        if allgather is not None:
            enc.set_group(rank, world, allgather)
        flush.fill_(step & 0xff)
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t0 = time.perf_counter()
        if K == 1:
            jpg, st, _ = enc.run()
        else:
            # the encodes of the batch run concurrently: one's host phases overlap the others' kernels
            res = [None] * K
            th = [threading.Thread(target=lambda k=k: res.__setitem__(k, encs[k].run())) for k in range(K)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            jpg, st, _ = res[0]
            for k in range(1, K):
                st = dict(st, launches=st["launches"] + res[k][1]["launches"], num_compares=st["num_compares"] + res[k][1]["num_compares"])
        e1.record()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) * 1e3
        if timed:
            run_ms.append(max(dt, e0.elapsed_time(e1)))
            launches += st["launches"]
            n_cmp += st["num_compares"]
            # the dominant kernel and the Compare pipeline are bracketed by CUDA events on their own
            # stream inside the library on every step (no per-launch profiling in the timed region)
            z_ms_sum += st["device_zeroing_ms"]
            cmp_ms_sum += st["device_compare_ms"]
        for en in encs:
            en.close()
        # ---- end-to-end arm: host RGB buffer -> host JPEG bytes
        flush.fill_((step + 1) & 0xff)
        barrier()
        t0 = time.perf_counter()
        if allgather is not None:
            jpg2, st2, _ = gz.ProcessGroup(img, target, dist, device=local, host_threads=host_threads)
        elif K == 1:
            jpg2, st2, _ = gz.Process(img, target, device=local, host_threads=host_threads)
        else:
            res2 = [None] * K
            th = [threading.Thread(target=lambda k=k: res2.__setitem__(k, gz.Process(images[step * K + k], target, device=local,
                                                                                     host_threads=host_threads))) for k in range(K)]
            for t in th:
                t.start()
            for t in th:
                t.join()
            jpg2, st2, _ = res2[0]
            for k in range(1, K):
                st2 = dict(st2, h2d_bytes=st2["h2d_bytes"] + res2[k][1]["h2d_bytes"], d2h_bytes=st2["d2h_bytes"] + res2[k][1]["d2h_bytes"])
        dt2 = (time.perf_counter() - t0) * 1e3
        assert jpg2 == jpg
        if timed:
            rounds.append(st["search_rounds"]); trials.append(st["search_trials"])
        if timed:
            e2e_ms.append(dt2)
            h2d += st2["h2d_bytes"]; d2h += st2["d2h_bytes"] + len(jpg2) * 0
    clocks = sampler.stop()
    # Throughput with several encodes in flight on the one GPU (their host phases overlap each other's
    # kernels): an extra, separately timed measurement next to the single-image headline.
    concurrent = None
    if world == 1 and K == 1 and not group and not a.no_concurrent:
        KC, NB = 3, 6
        ht = max(1, host_threads // KC)
        cimgs = [_libs.synth_image(w, h, 5000 + 10 * i) for i in range(NB * 2)]
        ctimes = []
        for step in range(1 + a.steps):
            flush.fill_(step & 0xff)
            torch.cuda.synchronize()
            batch = [cimgs[(step * NB + k) % len(cimgs)] for k in range(NB)]
            t0 = time.perf_counter()
            gz.ProcessBatch(batch, target, device=local, inflight=KC, host_threads_per_encode=ht)
            if step >= 1:
                ctimes.append((time.perf_counter() - t0) * 1e3)
        concurrent = {"inflight_per_gpu": KC, "host_threads_per_encode": ht, "images_per_batch": NB, "e2e_value": NB * w * h / 1e6 * len(ctimes) / (sum(ctimes) / 1e3),
                      "unit": "MPix/s", "ms_per_batch": sum(ctimes) / len(ctimes),
                      "note": "gzb_encode_rgb_batch: %d images per call from host buffers, %d encodes in flight on %d host threads "
                              "each; the headline value / e2e above are one image at a time" % (NB, KC, ht)}
    # per-kernel breakdown: ONE extra, untimed step with an event pair around every launch
    if rank == 0:
        enc = gz.Encoder(images[-1], target, device=local, host_threads=host_threads, profile=True)
        if allgather is None:
            enc.run()
            for k, (ms, n) in enc.kernel_times().items():
                kt[k] = (ms, n)
        enc.close()

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

    total_run = reduce_max(sum(run_ms))
    total_e2e = reduce_max(sum(e2e_ms))
    launches_all = int(reduce_sum(launches))
    mpix = w * h / 1e6
    images_per_step = 1 if group else world * K
    value = images_per_step * mpix * a.steps / (total_run / 1e3)
    e2e_value = images_per_step * mpix * a.steps / (total_e2e / 1e3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return
    peak, peak_src = peaks()
    nblocks = ((w + 7) // 8) * ((h + 7) // 8)
    z_ms, z_n = z_ms_sum, a.steps          # one zeroing launch per encode
    gpu_ms_total = sum(v[0] for v in kt.values())
    kz_ms = kt.get("k_zeroing_order", (0.0, 0))[0]
    roof = None
    if z_n:
        achieved = ALGO_BYTES_PER_BLOCK * nblocks / (z_ms / z_n / 1e3) / 1e9
        traffic, fp64_pct, ncu_src = ncu_evidence("k_zeroing_order", w, h)
        roof = {"bound": "hbm", "kernel": "k_zeroing_order", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": ALGO_BYTES_PER_BLOCK * nblocks,
                "avg_launch_ms": z_ms / z_n, "share_of_gpu_time": kz_ms / gpu_ms_total if gpu_ms_total else None,
                "fp64_pipe_pct_of_peak": fp64_pct, "ncu_source": ncu_src,
                "note": "FP64-pipe/latency-bound search kernel (SURVEY 8d): ~126 CompareBlock trials per 8x8 block in "
                        "double precision; its U2 bytes are touched once, so the HBM fraction is small by construction "
                        "and the FP64-pipe utilisation from ncu is the meaningful ceiling"}
    cmp_ms = cmp_ms_sum
    kernels = {k: {"ms_per_step": ms, "launches_per_step": n} for k, (ms, n) in sorted(kt.items(), key=lambda kv: -kv[1][0])}
    line = {
        "metric": "end-to-end encode MPix/s", "value": value, "unit": "MPix/s", "n_gpus": world, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": total_run / a.steps, "higher_is_better": True, "scaling": "strong" if group else "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "synthetic %dx%d sRGB (SURVEY 8d generator, seed 1234+10k), quality %g, %s, "
                               "guetzli::Process" % (w, h, a.quality,
                                                     "ONE image per step shared by all GPUs: SelectQuantMatrix candidates and "
                                                     "zeroing blocks sharded, NCCL all-gather per round" if group
                                                     else ("one image per GPU per step" if K == 1 else
                                                           "%d images per GPU per step, encoded concurrently" % K)),
                   "mode": a.mode, "search_rounds_per_step": sum(rounds) / max(1, len(rounds)),
                   "search_trials_per_step": sum(trials) / max(1, len(trials)),
                   "l2": "256 MiB buffer written between timed steps (L2 flush)", "host_threads_per_encode": host_threads, "inflight_per_gpu": K,
                   "compares_per_step": n_cmp / a.steps},
        "e2e": {"value": e2e_value, "unit": "MPix/s", "ms_per_step": total_e2e / a.steps,
                "h2d_bytes_per_step": h2d // a.steps, "d2h_bytes_per_step": d2h // a.steps},
        "gpu_launches": launches_all,
        "clocks": clocks,
        "roofline": roof,
        "butteraugli": {"compare_device_ms_per_call": cmp_ms / max(1, n_cmp), "mpix_per_s": mpix / (cmp_ms / max(1, n_cmp) / 1e3) if cmp_ms else None,
                        "hbm_frac_U1": (ALGO_BYTES_COMPARE_PER_PX * w * h / (cmp_ms / max(1, n_cmp) / 1e3) / 1e9 / peak) if cmp_ms else None},
        "concurrent": concurrent,
        "kernels": kernels,
        "kernels_note": "one extra untimed step with a CUDA-event pair around every launch; the timed steps carry no per-launch profiling",
        "cpu_baseline": cpu_baseline_single_core(a.quality) if world == 1 and not a.no_cpu_baseline else None,
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
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--size", default="1024x1024", type=lambda s: tuple(int(v) for v in s.lower().split("x")))
    ap.add_argument("--quality", type=float, default=90.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-concurrent", action="store_true", help="skip the extra 3-images-in-flight throughput measurement")
    ap.add_argument("--inflight", type=int, default=1,
                    help="batch mode: images encoded concurrently per GPU (one host thread group + one device context "
                         "each); a step is then a batch of that many images per GPU. Default 1 = one image per step")
    ap.add_argument("--mode", default="batch", choices=["batch", "group", "butteraugli"],
                    help="batch: one image per GPU (BASELINE configs[1]/[3]); group: one image shared by all GPUs "
                         "(configs[2]); butteraugli: standalone Compare sweep + quality sweep (configs[4])")
    a = ap.parse_args()
    guard_stdout()
    rank, world, local = dist_env()
    if a.impl == "reference":
        run_reference(a, rank, world)
    elif a.mode == "butteraugli":
        if rank == 0:
            run_butteraugli_sweep(a, local)
    else:
        run_ours(a, rank, world, local)


if __name__ == "__main__":
    main()
