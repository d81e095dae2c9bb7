B="--steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras"
L=polymer-stats_b200
python bench.py --workload K1 $B > gpurun_out/r02i_K1_base.json 2>/dev/null
PMC_LIB_PATH=$PWD/$L/libpolymc_b200_rsqk.so python bench.py --workload K1 $B > gpurun_out/r02i_K1_rsqpar.json 2>/dev/null
python bench.py --workload K5 $B > gpurun_out/r02i_K5_base.json 2>/dev/null
PMC_LIB_PATH=$PWD/$L/libpolymc_b200_rsqk.so python bench.py --workload K5 $B > gpurun_out/r02i_K5_rsqpar.json 2>/dev/null
python bench.py --workload C2 $B > gpurun_out/r02i_C2_base.json 2>/dev/null
PMC_LIB_PATH=$PWD/$L/libpolymc_b200_rsqc.so python bench.py --workload C2 $B > gpurun_out/r02i_C2_rsqpar.json 2>/dev/null
PMC_LIB_PATH=$PWD/$L/libpolymc_b200_rsqk.so timeout 60 python -m pytest tests/test_gpu_cluster.py -m gpu -x -q > gpurun_out/r02i_cluster_tests_rsqpar.log 2>&1; echo rc=$? >> gpurun_out/r02i_cluster_tests_rsqpar.log
for f in gpurun_out/r02i_*.json; do python - "$f" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], d['value'], d['roofline']['kernel'], d['roofline']['kernel_ms_avg'])
P
done; tail -2 gpurun_out/r02i_cluster_tests_rsqpar.log
