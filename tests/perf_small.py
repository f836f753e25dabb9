"""Ad-hoc: one small encode with trace (debug aid)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
img = synth_image(128, 96, 1234)
jpg, st, tr = gz.Process(img, np.float32(gz.ButteraugliScoreForQuality(90)), want_trace=True, host_threads=4)
print(tr)
print(len(jpg), st["num_iterations"])
