#!/bin/bash
# ncu captures (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ and are summarised into profiles/.
#   usage: bash tools/ncu_capture.sh <tag> [full]
#   launch list of the whole bench command (summarised per steady-state step by tools/summarize_launches.py --last-step)
#   + with `full`: `--set full` captures of the top kernels (source-level, -lineinfo build).
set -x
TAG=${1:-r1}
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_launches_${TAG}.log 2>&1
if [ "$2" == "full" ]; then
  NCU="ncu --set full --clock-control none --import-source on"
  B1="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
  $NCU -k regex:tc_gemm_kernel -s 60 -c 12 -f -o gpurun_out/${TAG}_tc_gemm $B1 > gpurun_out/ncu_gemm.log 2>&1
  $NCU -k regex:'tc_emm_pv_kernel|tc_score_kernel' -s 3 -c 4 -f -o gpurun_out/${TAG}_tc_emm $B1 > gpurun_out/ncu_emm.log 2>&1
  $NCU -k regex:'la_reduce_allheads|la_small_kernel|layernorm_vec_kernel|la_fold_merge|fine_window_gather|linear_simt' -s 30 -c 12 -f -o gpurun_out/${TAG}_hbm_kernels $B1 > gpurun_out/ncu_hbm.log 2>&1
fi
ls -la gpurun_out
