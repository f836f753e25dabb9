#!/bin/bash
# final measurement set of round 2 (one B200)
mkdir -p gpurun_out
O=gpurun_out
(timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4) > $O/r2_gpu_tests_final.log 2>&1
cat $O/r2_gpu_tests_final.log
(timeout 120 python __graft_entry__.py smoke 2>&1 | tail -2) > $O/r2_smoke.log; cat $O/r2_smoke.log
timeout 600 python bench.py > $O/r2_bench_12mpix_q95.json 2> $O/bench.err; echo bench rc=$?; tail -2 $O/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_reference_arm.json 2>> $O/bench.err; echo ref rc=$?
timeout 300 python bench.py --impl reference-cuda --steps 2 --warmup 1 > $O/r2_bench_reference_cuda.json 2>> $O/bench.err; echo refcuda rc=$?
timeout 600 python bench.py --mode butteraugli --steps 5 --warmup 3 > $O/r2_bench_butteraugli_sweep.json 2>> $O/bench.err; echo ba rc=$?
timeout 300 python bench.py --size 1024x1024 --quality 90 --no-extras > $O/r2_bench_1mpix_q90.json 2>> $O/bench.err; echo 1mpix rc=$?
timeout 300 python bench.py --size 1920x1080 --quality 95 --no-extras --no-cpu-baseline > $O/r2_bench_2mpix_q95.json 2>> $O/bench.err; echo 2mpix rc=$?
python - <<'PY'
import json
for f in ["r2_bench_12mpix_q95","r2_bench_reference_arm","r2_bench_reference_cuda","r2_bench_1mpix_q90","r2_bench_2mpix_q95"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%f)); print(f, d.get("value"), d.get("e2e",{}).get("value"), d.get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
# launch list of one whole encode of the bench workload (after the plain runs above exited 0)
TMP=/tmp/ncu_r2; mkdir -p $TMP
timeout 600 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $TMP/r2_launches_12mpix.csv python profiles/encode_probe.py 4000 3000 95 > $O/r2_launches.log 2>&1
python profiles/summarize.py launches $TMP/r2_launches_12mpix.csv \
    "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python profiles/encode_probe.py 4000 3000 95" > $O/r2_launches_12mpix.md
gzip -c $TMP/r2_launches_12mpix.csv > $O/r2_launches_12mpix.csv.gz
head -20 $O/r2_launches_12mpix.md
