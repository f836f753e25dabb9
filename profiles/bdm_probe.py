"""One full Compare between cudaProfilerStart/Stop (for ncu -k regex:k_block_diff): python profiles/bdm_probe.py W H"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h = int(sys.argv[1]), int(sys.argv[2])
img = synth_image(w, h)
c = gz.ButteraugliComparator(w, h, img, np.float32(gz.ButteraugliScoreForQuality(95)))
c.SetJpegCoeffs(gz.RgbToJpegCoeffs(img))
c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
c.Compare()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
c.Compare()
rt.cudaProfilerStop()
