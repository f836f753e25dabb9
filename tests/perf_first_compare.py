import sys, os, time
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, __graft_entry__ as ge, _libs
gz=ge.load_package()
img=_libs.synth_image(1024,1024)
co=gz.RgbToJpegCoeffs(img)
for rep in range(4):
    t0=time.time(); c=gz.ButteraugliComparator(1024,1024,img,0.97); t1=time.time()
    c.SetJpegCoeffs(co); c.CopyFromJpegData(); t2=time.time()
    c.Compare(); t3=time.time(); c.Compare(); t4=time.time(); c.close(); t5=time.time()
    print("create %.2f ms, coeffs %.2f, first compare %.2f, second %.2f, close %.2f" % ((t1-t0)*1e3,(t2-t1)*1e3,(t3-t2)*1e3,(t4-t3)*1e3,(t5-t4)*1e3))
