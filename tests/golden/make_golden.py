#!/usr/bin/env python
"""Regenerates the committed golden fixtures from the UNMODIFIED reference compiled into
oracle/_ref/libgzref.so (run where /root/reference is mounted: `python tests/golden/make_golden.py`).

  bees_q95.json        sha256/size/iterations + the verbose trace of guetzli::Process on
                       tests/bees.png at quality 95 -- must equal tests/golden_checksums.txt:3
  synth_encodes.json   sha256/size/iterations of guetzli::Process on seeded synthetic images
  yuv420_encodes.json  the same with Params::try_420 / force_420 (the YUV420 passes)
  stage_vectors.npz    small stage-level vectors (diffmap, distance, zeroing order) on a 96x64 image
"""
import hashlib, json, os, sys
import numpy as np
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from _libs import ref, ref_process, RefSession, synth_image, bees

GOLDEN_BEES = "39cfc110d3d389ddf3b9c47adece290ef077b78467f1ed39679b691c6fb7c96d"  # tests/golden_checksums.txt:3


def iter_lines(trace):
    return [l for l in trace.splitlines() if "Out[" in l]


def edge_cases():
    """edge_encodes.json: the reference's bytes and traces on the special images of tests/_libs.py."""
    from _libs import SPECIAL_IMAGES, special_image
    L = ref()
    enc = {}
    for name in SPECIAL_IMAGES:
        im, q = special_image(name)
        t = L.ref_butteraugli_score_for_quality(float(q))
        jpg, iters, trace = ref_process(im, t, want_trace=True)
        enc[name] = {"sha256": hashlib.sha256(jpg).hexdigest(), "size": len(jpg), "iterations": iters, "target": t,
                     "trace": iter_lines(trace)}
        print(name, len(jpg), iters)
    json.dump(enc, open(os.path.join(HERE, "edge_encodes.json"), "w"), indent=1)


CASES_420 = [("red", 96, 80, 95, "force"), ("red", 100, 75, 90, "force"), ("red", 129, 66, 92, "try"),
             ("synth", 96, 80, 95, "try"), ("red", 200, 136, 88, "try"), ("synth", 33, 47, 84, "force"),
             ("bees", 0, 0, 95, "try"), ("bees", 0, 0, 90, "force"),
             ("gray", 80, 64, 90, "force"), ("gray", 80, 64, 90, "try"), ("red", 32, 32, 95, "force")]


def yuv420_cases():
    """yuv420_encodes.json: guetzli::Process with Params::try_420 / force_420 (processor.h:34-42):
    bytes and iteration traces of the reference."""
    from _libs import image_420, ref_process_params
    L = ref()
    enc = {}
    for (kind, w, h, q, mode) in CASES_420:
        im = image_420(kind, w, h)
        t = L.ref_butteraugli_score_for_quality(float(q))
        jpg, iters, trace = ref_process_params(im, t, try_420=mode == "try", force_420=mode == "force", want_trace=True)
        name = "%s_%dx%d_q%d_%s" % (kind, im.shape[1], im.shape[0], q, mode)
        enc[name] = {"sha256": hashlib.sha256(jpg).hexdigest(), "size": len(jpg), "iterations": iters, "target": t,
                     "trace": iter_lines(trace)}
        print(name, len(jpg), iters)
    json.dump(enc, open(os.path.join(HERE, "yuv420_encodes.json"), "w"), indent=1)


def main():
    L = ref()
    out = {}
    img = bees()
    t = L.ref_butteraugli_score_for_quality(95.0)
    jpg, iters, trace = ref_process(img, t, want_trace=True)
    sha = hashlib.sha256(jpg).hexdigest()
    assert sha == GOLDEN_BEES, sha
    json.dump({"sha256": sha, "size": len(jpg), "iterations": iters, "target": t, "trace": iter_lines(trace)},
              open(os.path.join(HERE, "bees_q95.json"), "w"), indent=1)
    enc = {}
    for (w, h, q, seed) in [(128, 96, 90, 1234), (160, 120, 95, 1244), (97, 61, 84, 1254), (256, 256, 90, 1234)]:
        im = synth_image(w, h, seed)
        t = L.ref_butteraugli_score_for_quality(float(q))
        jpg, iters, trace = ref_process(im, t, want_trace=True)
        enc["%dx%d_q%d_s%d" % (w, h, q, seed)] = {"sha256": hashlib.sha256(jpg).hexdigest(), "size": len(jpg),
                                                   "iterations": iters, "target": t, "trace": iter_lines(trace)}
    json.dump(enc, open(os.path.join(HERE, "synth_encodes.json"), "w"), indent=1)
    edge_cases()
    w, h, target = 96, 64, 0.971769
    im = synth_image(w, h)
    s = RefSession(im, target)
    s.apply_quant(np.full(192, 3, np.int32))
    d, dm = s.compare()
    s.start_block_comparisons()
    zo = s.zeroing_order(7)
    np.savez_compressed(os.path.join(HERE, "stage_vectors.npz"), jpg_coeffs=s.jpg_coeffs(), coeffs=s.coeffs(),
                        srgb=s.to_srgb(), distance=np.float32(d), diffmap=dm, zo_idx=zo["idx"], zo_err=zo["err"])
    print("ok")


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "--edge":
        edge_cases()
    elif len(sys.argv) > 1 and sys.argv[1] == "--420":
        yuv420_cases()
    else:
        main()
