"""Ad-hoc leak / stability probe (not a test): many encodes of varying size and mode in one process;
prints device and host memory before and after. python tests/stress_probe.py [N]"""
import os, sys, resource
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
n = int(sys.argv[1]) if len(sys.argv) > 1 else 120
rng = np.random.default_rng(0)
t = np.float32(gz.ButteraugliScoreForQuality(90))
def mem():
    free, total = torch.cuda.mem_get_info()
    return (total - free) / 2**20, resource.getrusage(resource.RUSAGE_SELF).ru_maxrss / 1024
gz.Process(synth_image(256, 256), t)
print("start: device used %.0f MiB, host maxrss %.0f MiB" % mem())
for i in range(n):
    w, h = int(rng.integers(32, 700)), int(rng.integers(32, 500))
    img = synth_image(w, h, int(rng.integers(0, 1000)))
    mode = i % 4
    if mode == 0: gz.Process(img, t)
    elif mode == 1: gz.Process(img, t, try_420=True)
    elif mode == 2: gz.ProcessBatch([img, img[:max(32, h // 2), :max(32, w // 2)].copy()], t, inflight=2)
    else: gz.Process(img, t, force_420=True)
    if i % 30 == 29: print("after %d: device used %.0f MiB, host maxrss %.0f MiB" % ((i + 1,) + mem()))
print("done")
