"""One Compare + one zeroing-order pass of the bench workload (1024x1024) between
cudaProfilerStart/Stop, for `ncu --profile-from-start off --set full` (see profiles/README.md).
python ncu_probe.py [W H [420]]: with 420 the image is downsampled first (gzb_downsample_420) and both
zeroing passes of the YUV420 branch run (luma blocks, then 16x16 macro-blocks)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1024, 1024)
img = synth_image(w, h)
c = gz.ButteraugliComparator(w, h, img, np.float32(gz.ButteraugliScoreForQuality(90)))
yuv420 = len(sys.argv) > 3 and sys.argv[3] == "420"
c.SetJpegCoeffs(gz.RgbToJpegCoeffs(img))
if yuv420:
    c.Downsample420()
c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
c.Compare(); c.StartBlockComparisons()
rt = torch.cuda.cudart()
rt.cudaProfilerStart()
c.Compare()
if yuv420:
    c.ComputeBlockZeroingCandidates(1)
    c.ComputeBlockZeroingCandidates(6)
else:
    c.ComputeBlockZeroingCandidates(7)
rt.cudaProfilerStop()
print("probe done: distance %.4f" % c.distance)
