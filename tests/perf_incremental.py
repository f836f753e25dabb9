"""Ad-hoc probe (not a test): how many Compares of an encode were incremental, and Compare time after a
single-block update versus a full Compare. python tests/perf_incremental.py W H QUALITY"""
import sys, os, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h, q = int(sys.argv[1]), int(sys.argv[2]), float(sys.argv[3])
img = synth_image(w, h)
t = np.float32(gz.ButteraugliScoreForQuality(q))
L = gz.lib()
L.gzb_incremental_compare_count.restype = C.c_ulonglong
L.gzb_incremental_compare_count.argtypes = [C.c_void_p]
enc = gz.Encoder(img, t)
jpg, st, _ = enc.run()
print("encode: %d compares, %d incremental, device_compare_ms %.1f" % (st["num_compares"], L.gzb_incremental_compare_count(enc._ctx), st["device_compare_ms"]))
enc.close()
c = gz.ButteraugliComparator(w, h, img, t)
c.SetJpegCoeffs(gz.RgbToJpegCoeffs(img)); c.CopyFromJpegData(); c.ApplyGlobalQuantization(np.full(192, 3, np.int32))
c.Compare(); c.Compare()
print("full compare device ms %.3f" % c.last_device_ms())
rng = np.random.default_rng(1)
for n in (1, 4, 10, 40):
    ms = []
    for rep in range(5):
        blocks = rng.integers(0, c.num_blocks, n)
        c.UpdateCoeffs(blocks, np.full(n, 5, np.uint8), np.zeros(n, np.int16))
        c.Compare()
        ms.append(c.last_device_ms())
    print("after %d changed blocks: device ms %s (incremental so far %d)" % (n, ["%.3f" % m for m in ms], c.incremental_compare_count()))
