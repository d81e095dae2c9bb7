#!/bin/bash
# Round-2 evidence run on the GPU box (one call): sanitizer probes, launch list, ncu --set full captures.
# Usage: bash tools/r02_profile.sh [san|ncu|all]
mode=${1:-all}
mkdir -p gpurun_out
if [ "$mode" = "san" ] || [ "$mode" = "all" ]; then
  for tool in memcheck racecheck; do
    for part in plain pair cluster multi; do
      timeout 300 compute-sanitizer --tool $tool --print-limit 20 python tools/san_small.py $part > gpurun_out/r02_sanitizer_${tool}_${part}.txt 2>&1
      echo "$tool $part rc=$? $(grep -c 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/r02_sanitizer_${tool}_${part}.txt) $(grep 'SUMMARY' gpurun_out/r02_sanitizer_${tool}_${part}.txt | tail -1)"
    done
  done
fi
if [ "$mode" = "ncu" ] || [ "$mode" = "all" ]; then
  NCU="ncu --clock-control none"
  $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file gpurun_out/r02a_launches_bench_C2.csv \
      python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_launch.log 2>&1
  $NCU --set full --import-source on -k regex:k_run_cta_win --launch-skip 3 --launch-count 1 -o gpurun_out/r02a_k_run_cta_win_C2 -f \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-extras > gpurun_out/ncu_c2.log 2>&1
  $NCU --set full --import-source on -k regex:k_run_cta_pair --launch-skip 2 --launch-count 1 -o gpurun_out/r02a_k_run_cta_pair_C5 -f \
      python bench.py --workload C5 --trials-per-step 50 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c5.log 2>&1
  $NCU --set full --import-source on -k regex:k_run_warp --launch-skip 2 --launch-count 1 -o gpurun_out/r02a_k_run_warp_C4 -f \
      python bench.py --workload C4 --trials-per-step 20000 --steps 1 --warmup 2 --no-cpu-baseline --no-e2e > gpurun_out/ncu_c4.log 2>&1
  ls -la gpurun_out/*.ncu-rep
fi
