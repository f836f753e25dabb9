#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python bench.py --mode butteraugli --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print([(x['size'], round(x['compare_device_ms'],3)) for x in d['sweep']])"
