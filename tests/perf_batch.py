"""Ad-hoc probe (not a test): python tests/perf_batch.py W H QUALITY NIMG INFLIGHT -- gzb_encode_rgb_batch throughput."""
import sys, time, os, faulthandler
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image
import __graft_entry__ as ge
faulthandler.dump_traceback_later(int(os.environ.get("WATCHDOG", "100")), file=sys.stderr)
gz = ge.load_package()
w, h, q, n, k = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3]), int(sys.argv[4]), int(sys.argv[5])
imgs = [synth_image(w, h, 1234 + 10 * i) for i in range(min(n, 8))]
batch = [imgs[i % len(imgs)] for i in range(n)]
t = np.float32(gz.ButteraugliScoreForQuality(q))
ht = max(1, min(16, os.cpu_count() or 1) // k)
gz.ProcessBatch(batch[:k], t, inflight=k, host_threads_per_encode=ht)
t0 = time.time(); res = gz.ProcessBatch(batch, t, inflight=k, host_threads_per_encode=ht); dt = time.time() - t0
print("batch %d x %dx%d q%g inflight %d (%d host threads each): %.2f s  %.2f MPix/s" % (n, w, h, q, k, ht, dt, n * w * h / 1e6 / dt))
