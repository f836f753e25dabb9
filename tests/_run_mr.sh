#!/bin/bash
timeout 800 python -m pytest tests/test_gpu_encode.py tests/test_gpu_420.py -m gpu -x -q 2>&1 | tail -2
PQ_SEED=1284 timeout 300 python tests/perf_quick.py 4000 3000 95 2 2>&1 | tail -2 | cut -c1-560
timeout 300 python tests/perf_quick.py 4000 3000 95 3 2>&1 | tail -2 | cut -c1-560
timeout 300 python tests/perf_quick.py 1920 1080 95 3 2>&1 | tail -2 | cut -c1-300
timeout 300 python tests/perf_quick.py 1024 1024 90 3 2>&1 | tail -2 | cut -c1-300
