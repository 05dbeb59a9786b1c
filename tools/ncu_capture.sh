#!/bin/bash
# Round-1 ncu captures (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ and are summarised into profiles/.
#   launch list of one steady-state step  +  `--set full` captures of the top kernels (source-level, -lineinfo build).
set -x
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
NCU="ncu --set full --clock-control none --import-source on"
$NCU -k regex:tc_gemm_kernel -s 150 -c 9 -f -o gpurun_out/r1_tc_gemm $B > gpurun_out/ncu_gemm.log 2>&1
$NCU -k regex:'tc_emm_pv_kernel|tc_score_kernel|emm_vt_kernel' -s 6 -c 5 -f -o gpurun_out/r1_tc_emm $B > gpurun_out/ncu_emm.log 2>&1
$NCU -k regex:'la_reduce_allheads|la_apply_allheads|la_small_kernel|layernorm_vec_kernel|fine_window_gather|split_groups|fine_match_kernel|upsample2x|scale_shift' -s 40 -c 14 -f -o gpurun_out/r1_hbm_kernels $B > gpurun_out/ncu_hbm.log 2>&1
ls -la gpurun_out
