#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_encode.py tests/test_gpu_420.py tests/test_gpu_backend.py -m gpu -x -q 2>&1 | tail -3
for cfg in "1024 1024 90" "1920 1080 95" "4000 3000 95"; do
  timeout 200 python tests/perf_quick.py $cfg 3 2>&1 | tail -2 | cut -c1-400
done
