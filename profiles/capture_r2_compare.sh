#!/bin/bash
# Re-capture (final build of round 2) of the kernels of one full Compare inside a 12 MPix encode: the same pass as
# step (3) of capture_r2.sh, after BlockDiffMap became k_block_diff_strip. Only text leaves the box.
set -x
mkdir -p gpurun_out
OUT=gpurun_out
TMP=/tmp/ncu_r2
mkdir -p $TMP
FP64="smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active"
python profiles/encode_probe.py 4000 3000 95 || exit 1
timeout 900 ncu --profile-from-start off --set full --metrics $FP64 --clock-control none --import-source on \
    -k regex:"k_block_diff_strip|k_block_dc|k_blur|k_opsin|k_mask|k_edge|k_combine|k_diffmap|k_coeffs_to_rgb8|k_max_u32" --launch-count 28 \
    -o $TMP/r2_compare_12mpix -f python profiles/encode_probe.py 4000 3000 95 > $OUT/r2_ncu_compare.log 2>&1
python profiles/summarize.py report $TMP/r2_compare_12mpix.ncu-rep > $OUT/r2_ncu_full_compare_12mpix_final.md
python - <<PY
import sys, json
sys.path.insert(0, "profiles")
import summarize
note = ("ncu --set full --clock-control none of python profiles/encode_probe.py 4000 3000 95 (final build of round 2: the Compare kernels "
        "from the first Compare of one whole encode of the bench workload)")
b = summarize.traffic("$TMP/r2_compare_12mpix.ncu-rep", note)
json.dump(b, open("$OUT/r2_ncu_traffic_compare_final.json", "w"), indent=1)
PY
ls -la $OUT | tail -5
