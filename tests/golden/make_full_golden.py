#!/usr/bin/env python
"""Golden results of the UNMODIFIED reference (oracle/_ref) on the FULL bench workloads of
BASELINE.json: synthetic 1024x1024 at q90 (configs[1]), 4000x3000 at q95 (configs[2]) and two images of the
64-image batch of configs[3] (1920x1080 at q95, seeds 1234 and 1244), same generator and seeds as bench.py.
Single-threaded CPU Guetzli: about two minutes, half an hour and five minutes each. Writes tests/golden/full_encodes.json (sha256, size, iterations, seconds, trace)."""
import hashlib, json, os, sys, time
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _libs import ref, ref_process, synth_image


def main():
    L = ref()
    path = os.path.join(HERE, "full_encodes.json")
    out = json.load(open(path)) if os.path.exists(path) else {}
    for (w, h, q, seed) in [(1024, 1024, 90, 1234), (4000, 3000, 95, 1234), (1920, 1080, 95, 1234), (1920, 1080, 95, 1244)]:
        key = "%dx%d_q%d_s%d" % (w, h, q, seed)
        if key in out:
            continue
        im = synth_image(w, h, seed)
        t = L.ref_butteraugli_score_for_quality(float(q))
        t0 = time.time()
        jpg, iters, trace = ref_process(im, t, want_trace=True)
        dt = time.time() - t0
        out[key] = {"sha256": hashlib.sha256(jpg).hexdigest(), "size": len(jpg), "iterations": iters, "target": t,
                    "reference_seconds_one_core": dt, "trace": [l for l in trace.splitlines() if "Out[" in l]}
        json.dump(out, open(path, "w"), indent=1)
        print(key, len(jpg), iters, "%.1f s" % dt, flush=True)


if __name__ == "__main__":
    main()
