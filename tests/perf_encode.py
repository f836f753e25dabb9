"""Ad-hoc e2e probe (not a test): python tests/perf_encode.py W H QUALITY [threads] [try|force]"""
import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h, q = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
nt = int(sys.argv[4]) if len(sys.argv) > 4 else 0
mode = sys.argv[5] if len(sys.argv) > 5 else ""
kw = dict(try_420=mode == "try", force_420=mode == "force")
img = synth_image(w, h)
t = np.float32(gz.ButteraugliScoreForQuality(q))
for rep in range(2):
    t0 = time.time(); jpg, st, _ = gz.Process(img, t, host_threads=nt, **kw); dt = time.time() - t0
    print("encode %dx%d q%g: %.3f s -> %.3f MPix/s, %d bytes, iters %d" % (w, h, q, dt, w * h / 1e6 / dt, len(jpg), st["num_iterations"]))
    print(json.dumps({k: (round(v, 2) if isinstance(v, float) else v) for k, v in st.items()}))
os.system("nproc; grep -m1 'model name' /proc/cpuinfo")
# two-step: where does the wall time outside gzb_encoder_run go?
for rep in range(2):
    t0 = time.time(); enc = gz.Encoder(img, t, host_threads=nt, **kw); t1 = time.time()
    jpg, st, _ = enc.run(); t2 = time.time(); enc.close(); t3 = time.time()
    print("two-step: create %.1f ms (lib prepare %.1f)  run %.1f ms (lib run %.1f)  close %.1f ms" %
          ((t1 - t0) * 1e3, st["prepare_ms"], (t2 - t1) * 1e3, st["run_ms"], (t3 - t2) * 1e3))
