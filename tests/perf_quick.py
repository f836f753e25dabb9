"""Ad-hoc probe (not a test): python tests/perf_quick.py W H QUALITY [reps] -- one-line timing + back-end breakdown."""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h, q = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
img = synth_image(w, h, int(os.environ.get("PQ_SEED", "1234")))
t = np.float32(gz.ButteraugliScoreForQuality(q))
best = None
for rep in range(reps + 1):
    t0 = time.time(); jpg, st, _ = gz.Process(img, t); dt = time.time() - t0
    if rep and (best is None or dt < best[0]): best = (dt, st)
dt, st = best
keys = ["search_wall_ms", "zeroing_wall_ms", "device_zeroing_ms", "backend_wall_ms", "compare_wall_ms", "device_compare_ms", "device_write_ms",
        "be_order_ms", "be_walk_ms", "be_sort_ms", "be_select_ms", "be_lazy_ms", "be_gather_ms", "be_codes_ms", "be_pool_ms", "be_update_ms", "be_selects",
        "be_levels", "be_host_ranges", "num_iterations", "num_entropy_code_builds", "be_steps", "be_prefix_steps", "d2h_bytes", "h2d_bytes", "launches"]
print("%dx%d q%g best of %d: %.1f ms  %.2f MPix/s  %d bytes  env=%s" % (w, h, q, reps, dt * 1e3, w * h / 1e6 / dt, len(jpg),
      {k: v for k, v in os.environ.items() if k.startswith("GZB_")}))
print("   " + " ".join("%s=%s" % (k.replace("_ms", ""), round(st[k], 1) if isinstance(st[k], float) else st[k]) for k in keys))
