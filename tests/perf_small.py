"""Ad-hoc: group encode of a special image vs golden (debug aid)."""
import sys, os, json, hashlib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from _libs import special_image, ROOT
import __graft_entry__ as ge
import test_gpu_encode as T
gz = ge.load_package()
name = sys.argv[1] if len(sys.argv) > 1 else "gray_80x64_q90"
world = int(sys.argv[2]) if len(sys.argv) > 2 else 3
gold = json.load(open(os.path.join(ROOT, "tests", "golden", "edge_encodes.json")))[name]
img, _ = special_image(name)
res = T.run_thread_group(gz, img, np.float32(gold["target"]), world)
jpg, st, tr = res[0]
got = [l for l in tr.splitlines() if "Out[" in l]
for a, b in list(zip(got, gold["trace"]))[:14]:
    print("GOT ", a); print("WANT", b)
