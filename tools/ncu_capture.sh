#!/bin/bash
# ncu captures (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ and are summarised into profiles/.
#   usage: bash tools/ncu_capture.sh <tag> [full]
#   launch list of the whole bench command (summarised per steady-state step by tools/summarize_launches.py --last-step)
#   + with `full`: `--set full` captures of the top kernels (source-level, -lineinfo build).  The .ncu-rep files are
#   exported to CSV on the box (raw page + SASS source page of one launch) and deleted: gpurun_out/ is capped at 64 MiB.
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_launches_${TAG}.log 2>&1
# DRAM traffic of every launch of OUR kernels in a steady-state step (single-pass counters, same metrics the full set
# reports): bench.py's roofline.traffic is the per-launch mean over exactly the launches its `achieved` averages over.
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'tc_|la_|layernorm|fine_|match_|stem_|eightpt|pose_|emm_|linear_simt|split|lse_|upsample2x|scale_shift|pos_flatten' \
    -c 3000 --csv --log-file gpurun_out/traffic_${TAG}.csv $B > gpurun_out/ncu_traffic_${TAG}.log 2>&1
if [ "$2" == "full" ]; then
  NCU="ncu --set full --clock-control none --import-source on"
  B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
  export_rep() {  # <name> <launch index for the source page>
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass --launch-skip $2 --launch-count 1 > gpurun_out/$1_src.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
  }
  $NCU -k regex:tc_gemm_kernel -s 60 -c 10 -f -o gpurun_out/${TAG}_tc_gemm $B1 > gpurun_out/ncu_gemm.log 2>&1
  export_rep ${TAG}_tc_gemm 2
  $NCU -k regex:'tc_emm_pv_kernel' -s 0 -c 1 -f -o gpurun_out/${TAG}_tc_emm_pv $B1 > gpurun_out/ncu_emm.log 2>&1
  export_rep ${TAG}_tc_emm_pv 0
  $NCU -k regex:'tc_score_kernel' -s 1 -c 3 -f -o gpurun_out/${TAG}_tc_score $B1 > gpurun_out/ncu_score.log 2>&1
  export_rep ${TAG}_tc_score 0
  $NCU -k regex:'la_reduce_allheads|la_small|layernorm_vec_kernel|la_fold_merge|fine_window_gather|fine_match_kernel|match_decide|eightpt|pose_solve' -s 20 -c 14 -f -o gpurun_out/${TAG}_hbm_kernels $B1 > gpurun_out/ncu_hbm.log 2>&1
  export_rep ${TAG}_hbm_kernels 0
fi
ls -la gpurun_out
du -sh gpurun_out
