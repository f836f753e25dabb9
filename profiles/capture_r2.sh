#!/bin/bash
# Round-2 ncu captures on the GPU box (run with gpurun from the repo root). Only text summaries are left in
# gpurun_out/ (the .ncu-rep files of a 12 MPix context are ~50 MB each and gpurun returns at most 64 MiB).
# Each ncu pass follows a plain run of the same command that exited 0.
set -x
mkdir -p gpurun_out
OUT=gpurun_out
TMP=/tmp/ncu_r2
mkdir -p $TMP
FP64="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
python profiles/encode_probe.py 4000 3000 95 || exit 1
# (1) launch list of one whole encode of the bench workload
timeout 900 ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
    --log-file $TMP/r2_launches_12mpix.csv python profiles/encode_probe.py 4000 3000 95 > $OUT/r2_launches.log 2>&1
python profiles/summarize.py launches $TMP/r2_launches_12mpix.csv \
    "ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none python profiles/encode_probe.py 4000 3000 95" > $OUT/r2_launches_12mpix.md
gzip -c $TMP/r2_launches_12mpix.csv > $OUT/r2_launches_12mpix.csv.gz
# (2) --set full of the zeroing search INSIDE the encode (the bench's candidate: the matrix SelectQuantMatrix chose)
timeout 900 ncu --profile-from-start off --set full --metrics $FP64 --clock-control none --import-source on \
    -k regex:k_zeroing_order --launch-count 1 -o $TMP/r2_zeroing_12mpix -f python profiles/encode_probe.py 4000 3000 95 > $OUT/r2_ncu_zeroing.log 2>&1
# (3) --set full of the kernels of the first (full) Compare of the same encode
timeout 900 ncu --profile-from-start off --set full --metrics $FP64 --clock-control none --import-source on \
    -k regex:"k_block_diff_map|k_block_dc|k_blur|k_opsin|k_mask|k_edge|k_combine|k_diffmap|k_coeffs_to_rgb8|k_max_u32" --launch-count 28 \
    -o $TMP/r2_compare_12mpix -f python profiles/encode_probe.py 4000 3000 95 > $OUT/r2_ncu_compare.log 2>&1
python profiles/summarize.py report $TMP/r2_zeroing_12mpix.ncu-rep $TMP/r2_compare_12mpix.ncu-rep > $OUT/r2_ncu_full_zeroing_compare_12mpix.md
cat $TMP/r2_zeroing_12mpix.ncu-rep > /dev/null
python - <<PY
import sys, json
sys.path.insert(0, "profiles")
import summarize
note = ("ncu --set full --clock-control none of python profiles/encode_probe.py 4000 3000 95 (one whole encode of the bench "
        "workload; k_zeroing_order on the candidate SelectQuantMatrix chose, the Compare kernels from the first Compare)")
a = summarize.traffic("$TMP/r2_zeroing_12mpix.ncu-rep", note)
b = summarize.traffic("$TMP/r2_compare_12mpix.ncu-rep", note)
a["kernels"].update(b["kernels"])
json.dump({"sizes": {"4000x3000": a}}, open("$OUT/r2_ncu_traffic.json", "w"), indent=1)
PY
ncu -i $TMP/r2_zeroing_12mpix.ncu-rep --page source --csv > $TMP/zeroing_source.csv 2>/dev/null && gzip -c $TMP/zeroing_source.csv > $OUT/r2_zeroing_source.csv.gz
ls -la $OUT
