#!/bin/bash
# Same-box A/B of two builds of libgzb200.so (not a test): the working tree's library against
# guetzli-cuda-opencl_b200/build/prev/libgzb200.so (build the older revision there first).
# Runs the GPU parity suite on the new build, then bench.py new / prev / new at 1 MPix and prev / new at
# 12 MPix, printing value, ms per step, Compare ms, zeroing-kernel ms and the block-diff kernel class.
# usage (on the GPU box, from the repo root): bash tests/probes/ab.sh TAG
tag=${1:-ab}
L=guetzli-cuda-opencl_b200
timeout 100 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?"; tail -2 gpurun_out/${tag}_tests.log
P='import json,sys; d=json.load(open(sys.argv[1])); print(sys.argv[1], round(d["value"],2), round(d["ms_per_step"],2), "cmp", round(d["butteraugli"]["compare_device_ms_per_call"],4), "zero", round(d["roofline"]["avg_launch_ms"],3), "bdm", round(d["kernels"]["k_block_diff_map"]["ms_per_step"],2))'
cp $L/libgzb200.so /tmp/new.so
use() { if [ "$1" = prev ]; then cp $L/build/prev/libgzb200.so $L/libgzb200.so; else cp /tmp/new.so $L/libgzb200.so; fi; }
for v in new prev new2; do
  use $v
  timeout 40 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-concurrent > gpurun_out/${tag}_$v.json 2>>gpurun_out/${tag}_err.log
  python -c "$P" gpurun_out/${tag}_$v.json
done
for v in prev new; do
  use $v
  timeout 40 python bench.py --size 4000x3000 --quality 95 --steps 2 --warmup 3 --no-cpu-baseline --no-concurrent > gpurun_out/${tag}_12_$v.json 2>>gpurun_out/${tag}_err.log
  python -c "$P" gpurun_out/${tag}_12_$v.json
done
use new
