"""Small end-to-end run for `compute-sanitizer --tool memcheck` (not a test): 4:4:4 and YUV420 encodes on
sizes that are no multiple of 8 / 16, stage calls, the batch entry point."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image, image_420
import __graft_entry__ as ge
gz = ge.load_package()
t = np.float32(gz.ButteraugliScoreForQuality(90))
a = gz.Process(synth_image(97, 61, 1254), t)[0]
b = gz.Process(image_420("red", 100, 75), t, try_420=True)[0]
c = gz.Process(image_420("red", 33, 47), t, force_420=True)[0]
d = gz.ProcessBatch([synth_image(64, 48), synth_image(40, 33)], t, inflight=2)
cmp_ = gz.ButteraugliComparator(67, 45, synth_image(67, 45), 0.97)
cmp_.SetJpegCoeffs(gz.RgbToJpegCoeffs(synth_image(67, 45)))
cmp_.CopyFromJpegData(); cmp_.ApplyGlobalQuantization(np.full(192, 3, np.int32))
cmp_.Compare(); cmp_.StartBlockComparisons(); cmp_.ComputeBlockZeroingOrder(7); cmp_.CompareBlocks()
cmp_.ComputeBlockErrorAdjustmentWeights(1, 2, 1.0); cmp_.ComputeBlockErrorAdjustmentWeights(-1, 3, 1.0, factor=2)
cmp_.WriteJpeg(np.full(192, 3, np.int32))
print("sanitize probe ok", len(a), len(b), len(c), [len(x[0]) for x in d])
