"""Ad-hoc timing probe (not a test): python tests/perf_probe.py W H"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import oracle, p, synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h = int(sys.argv[1]), int(sys.argv[2])
t0 = time.time(); img = synth_image(w, h); print("synth %.2fs" % (time.time() - t0))
nb = ((w + 7) // 8) * ((h + 7) // 8)
orig = np.zeros((3, nb, 64), np.int16)
t0 = time.time(); oracle().gzo_rgb_to_jpeg_coeffs(p(img), w, h, p(orig[0]), p(orig[1]), p(orig[2])); print("fdct(cpu oracle) %.2fs" % (time.time() - t0))
t0 = time.time(); c = gz.ButteraugliComparator(w, h, img, 0.971769); print("create %.3fs" % (time.time() - t0))
c.SetJpegCoeffs(orig); c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
c.Compare()
c.profile(True)
for i in range(4):
    t0 = time.time(); d = c.Compare(); t1 = time.time()
    print("compare: wall %.3f ms, device %.3f ms, dist %.4f, %.1f MPix/s" % ((t1 - t0) * 1e3, c.last_device_ms(), d, w * h / 1e6 / (c.last_device_ms() / 1e3)))
t0 = time.time(); c.StartBlockComparisons(); print("start_block_cmp wall %.3f ms dev %.3f" % ((time.time() - t0) * 1e3, c.last_device_ms()))
for i in range(2):
    t0 = time.time(); zo = c.ComputeBlockZeroingOrder(7); t1 = time.time()
    print("zeroing: wall %.3f ms, device %.3f ms, candidates %d" % ((t1 - t0) * 1e3, c.last_device_ms(), (zo["err"] > 0).sum()))

kt = c.kernel_times()
tot = sum(v[0] for v in kt.values())
for k, (ms, n) in sorted(kt.items(), key=lambda kv: -kv[1][0]):
    print("  %-32s %8.3f ms total  %5d launches  %8.1f us avg  %5.1f%%" % (k, ms, n, 1e3 * ms / n, 100 * ms / tot))
