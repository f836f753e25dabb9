"""One whole encode (gzb_encode_rgb) of the bench workload between cudaProfilerStart/Stop, after a warm-up
encode, for `ncu --profile-from-start off ...` launch lists: python encode_probe.py [W H QUALITY]."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h, q = (int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])) if len(sys.argv) > 3 else (4000, 3000, 95.0)
img = synth_image(w, h)
t = np.float32(gz.ButteraugliScoreForQuality(q))
gz.Process(img, t)
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
jpg, st, _ = gz.Process(img, t)
rt.cudaProfilerStop()
print("probe done: %d bytes, %d iterations, %d launches, %.1f ms" % (len(jpg), st["num_iterations"], st["launches"], st["total_wall_ms"]))
