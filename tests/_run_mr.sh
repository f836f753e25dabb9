#!/bin/bash
timeout 800 python -m pytest tests/test_gpu_encode.py tests/test_gpu_420.py -m gpu -x -q 2>&1 | tail -8
