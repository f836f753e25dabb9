import sys, time, os, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
torch.cuda.set_device(0)
img = synth_image(1024, 1024)
t = np.float32(gz.ButteraugliScoreForQuality(90))
keys = ["total_wall_ms","prepare_ms","run_ms","create_ms","host_frontend_ms","host_quant_ms","host_write_ms","compare_wall_ms","zeroing_wall_ms","backend_wall_ms","be_walk_ms","be_order_ms","be_weights_ms","be_update_ms"]
for rep in range(3):
    t0 = time.time(); jpg, st, _ = gz.Process(img, t, host_threads=16); dt = time.time() - t0
    print("Process: %.3f s" % dt, {k: round(st[k], 1) for k in keys})
for rep in range(3):
    t0 = time.time(); enc = gz.Encoder(img, t, host_threads=16); t1 = time.time(); jpg, st, _ = enc.run(); t2 = time.time(); enc.close(); t3 = time.time()
    print("Encoder: create %.3f run %.3f close %.3f" % (t1 - t0, t2 - t1, t3 - t2), {k: round(st[k], 1) for k in keys})
