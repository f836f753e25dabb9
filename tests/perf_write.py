import sys, os, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import synth_image
import __graft_entry__ as ge
gz = ge.load_package()
w, h = int(sys.argv[1]), int(sys.argv[2])
img = synth_image(w, h)
co = gz.RgbToJpegCoeffs(img)
idx = (co // 3).astype(np.int16)   # q=3-like sparsity
L = gz.lib()
L.gzb_bench_write_jpeg.restype = C.c_double
L.gzb_bench_write_jpeg.argtypes = [C.c_void_p] * 3 + [C.c_int] * 5 + [C.c_void_p]
for nt in (1, 2, 4, 8, 16):
    for wh in (0, 1):
        parts = np.zeros(4)
        ms = L.gzb_bench_write_jpeg(idx[0].ctypes.data, idx[1].ctypes.data, idx[2].ctypes.data, w, h, nt, 10, wh, parts.ctypes.data)
        print("threads %2d hist_given %d: %.2f ms/write  hist %.2f code %.2f encode %.2f stitch %.2f" % (nt, wh, ms, *parts))
