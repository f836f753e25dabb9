"""Ad-hoc A/B timing of two builds of the library (not a test): python tests/perf_ab.py W H Q other.so"""
import os, sys, subprocess, json
w, h, q, other = sys.argv[1:5]
here = os.path.dirname(os.path.abspath(__file__))
code = ("import sys,os,time; sys.path.insert(0,%r); sys.path.insert(0,os.path.dirname(%r)); import numpy as np; "
        "from _libs import synth_image; import __graft_entry__ as ge; gz=ge.load_package(); img=synth_image(%s,%s); "
        "t=np.float32(gz.ButteraugliScoreForQuality(%s)); gz.Process(img,t); gz.Process(img,t); r=[]\n"
        "for i in range(6):\n t0=time.time(); gz.Process(img,t); r.append((time.time()-t0)*1e3)\n"
        "print(min(r), sorted(r)[len(r)//2])") % (here, here, w, h, q)
for rep in range(3):
    for name, lib in (("new", ""), ("old", other)):
        env = dict(os.environ)
        if lib: env["GZB200_LIB"] = lib
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True).stdout.strip().splitlines()[-1]
        print(name, out)
