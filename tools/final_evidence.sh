#!/bin/bash
# Final-build evidence in ONE gpurun call (1 GPU): the --set full captures of the top tcgen05 kernels (raw + SASS source
# pages exported to CSV on the box), then the bench lines of the four workloads and the reference arm.
#   usage: bash tools/final_evidence.sh <tag>
TAG=${1:-r3}
mkdir -p gpurun_out
F="--no-cpu-baseline --no-gpu-eager-baseline --no-graph"
NCU="ncu --set full --clock-control none --import-source on"
B1="python bench.py --steps 1 --warmup 1 $F"
export_rep() {  # <name> <launch index for the source page>
  ncu -i gpurun_out/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
  ncu -i gpurun_out/$1.ncu-rep --page source --csv --print-source sass --launch-skip $2 --launch-count 1 > gpurun_out/$1_src.csv 2>/dev/null
  rm -f gpurun_out/$1.ncu-rep
}
timeout 240 $NCU -k regex:'tc_gemm' -s 60 -c 8 -f -o gpurun_out/${TAG}_tc_gemm $B1 > gpurun_out/ncu_gemm.log 2>&1
export_rep ${TAG}_tc_gemm 2
timeout 240 $NCU -k regex:'tc_emm_pv_kernel|tc_lse64_kernel' -s 0 -c 2 -f -o gpurun_out/${TAG}_tc_emm $B1 > gpurun_out/ncu_emm.log 2>&1
export_rep ${TAG}_tc_emm 0
timeout 240 $NCU -k regex:'tc_score_kernel' -s 0 -c 2 -f -o gpurun_out/${TAG}_tc_score $B1 > gpurun_out/ncu_score.log 2>&1
export_rep ${TAG}_tc_score 0
timeout 240 $NCU -k regex:'la_reduce_kv_async|la_small|layernorm_vec_kernel|fine_window_gather|ransac_score' -s 20 -c 10 -f -o gpurun_out/${TAG}_hbm_kernels $B1 > gpurun_out/ncu_hbm.log 2>&1
export_rep ${TAG}_hbm_kernels 0
# bench lines (clean, no profiler)
timeout 400 python bench.py > gpurun_out/${TAG}_bench_mp3d_loftr_far.json 2> gpurun_out/${TAG}_bench_mp3d.err
for W in vit8pt_b64 mapfree_6dreg micro_4096x2048; do
  timeout 400 python bench.py --workload $W > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
done
timeout 400 python bench.py --impl reference > gpurun_out/${TAG}_bench_reference_arm.json 2> gpurun_out/${TAG}_bench_reference_arm.err
ls -la gpurun_out | grep ${TAG}_ | head -30
du -sh gpurun_out
