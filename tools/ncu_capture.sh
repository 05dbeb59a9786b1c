#!/bin/bash
# ncu captures (run under gpurun, ONE GPU).  Outputs land in gpurun_out/ and are summarised into profiles/.
#   usage: bash tools/ncu_capture.sh <tag> [full]
#   launch lists (gpu__time_duration.sum) of one bench command per workload, summarised per steady-state step by
#   tools/summarize_launches.py --last-step; DRAM bytes of every launch of OUR kernels over one step (feeds
#   bench.py's roofline.traffic through profiles/ncu_traffic.json);
#   + with `full`: `--set full` captures of the top kernels (source-level, -lineinfo build).  The .ncu-rep files are
#   exported to CSV on the box (raw page + SASS source page of one launch) and deleted: gpurun_out/ is capped at 64 MiB.
# The bench runs WITHOUT the CUDA graph (--no-graph) so every kernel is an ordinary launch for the profiler, and without
# the baselines; numbers printed by a bench run under ncu are never bench values.
set -x
TAG=${1:-r3}
mkdir -p gpurun_out
F="--no-cpu-baseline --no-gpu-eager-baseline --no-graph"
B="python bench.py --steps 1 --warmup 3 $F"
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_launches_${TAG}.log 2>&1
for W in vit8pt_b64 mapfree_6dreg; do
  ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/launches_${TAG}_${W}.csv \
      python bench.py --workload $W --steps 1 --warmup 3 $F > gpurun_out/ncu_launches_${TAG}_${W}.log 2>&1
done
# DRAM traffic of every launch of OUR kernels in a steady-state step (single-pass counters, same metrics the full set
# reports): bench.py's roofline.traffic is the per-launch mean over exactly the launches its `achieved` averages over.
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none \
    -k regex:'tc_|la_|layernorm|fine_|match_|stem_|eightpt|pose_|emm_|linear_simt|split|lse_|upsample2x|scale_shift|pos_flatten|ransac_|corrvol|attn_prep' \
    -c 4000 --csv --log-file gpurun_out/traffic_${TAG}.csv $B > gpurun_out/ncu_traffic_${TAG}.log 2>&1
if [ "$2" == "full" ]; then
  NCU="ncu --set full --clock-control none --import-source on"
  B1="python bench.py --steps 1 --warmup 1 $F"
  export_rep() {  # <name> <launch index for the source page>
    ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
    ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass --launch-skip $2 --launch-count 1 > gpurun_out/$1_src.csv 2>/dev/null
    rm -f gpurun_out/$1.ncu-rep
  }
  $NCU -k regex:'tc_gemm' -s 60 -c 10 -f -o gpurun_out/${TAG}_tc_gemm $B1 > gpurun_out/ncu_gemm.log 2>&1
  export_rep ${TAG}_tc_gemm 2
  $NCU -k regex:'tc_emm_pv_kernel|tc_lse64_kernel' -s 0 -c 2 -f -o gpurun_out/${TAG}_tc_emm $B1 > gpurun_out/ncu_emm.log 2>&1
  export_rep ${TAG}_tc_emm 0
  $NCU -k regex:'tc_score_kernel' -s 0 -c 2 -f -o gpurun_out/${TAG}_tc_score $B1 > gpurun_out/ncu_score.log 2>&1
  export_rep ${TAG}_tc_score 0
  $NCU -k regex:'la_reduce_kv_async|la_small|layernorm_vec_kernel|la_fold_merge|fine_window_gather|fine_match_kernel|match_decide|ransac_score|ransac_sample|upsample2x|stem_conv' -s 20 -c 16 -f -o gpurun_out/${TAG}_hbm_kernels $B1 > gpurun_out/ncu_hbm.log 2>&1
  export_rep ${TAG}_hbm_kernels 0
  $NCU -k regex:'tc_flash_kernel' -s 0 -c 2 -f -o gpurun_out/${TAG}_tc_flash python bench.py --workload vit8pt_b64 --steps 1 --warmup 1 $F > gpurun_out/ncu_flash.log 2>&1
  export_rep ${TAG}_tc_flash 0
  $NCU -k regex:'tc_flash_kernel|corrvol_prep' -s 0 -c 3 -f -o gpurun_out/${TAG}_tc_corrvol python bench.py --workload mapfree_6dreg --steps 1 --warmup 1 $F > gpurun_out/ncu_corrvol.log 2>&1
  export_rep ${TAG}_tc_corrvol 0
  $NCU -k regex:'eightpt|essential_decompose|tc_score_kernel' -s 0 -c 6 -f -o gpurun_out/${TAG}_micro python bench.py --workload micro_4096x2048 --pairs 512 --steps 1 --warmup 1 $F > gpurun_out/ncu_micro.log 2>&1
  export_rep ${TAG}_micro 0
  rm -f gpurun_out/*_src.csv.bak
fi
ls -la gpurun_out | head -50
du -sh gpurun_out
