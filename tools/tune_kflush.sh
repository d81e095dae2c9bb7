#!/bin/bash
# FP32 rectangle: terms per FP32 chunk sum (PMC_KFLUSH) against throughput and error (developer tool; experiment builds)
for f in _a0f128 _a0f32 _a0f16 _f128; do
  lib=polymer-stats_b200/libpolymc_b200$f.so
  echo "== $lib"
  PMC_LIB_PATH=$(pwd)/$lib python bench.py --workload C2f32 --steps 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3f M updates/s' % (d['value']/1e6), d['roofline']['kernel'])"
  PMC_LIB_PATH=$(pwd)/$lib python -m pytest tests/test_gpu_fp32.py -m gpu -q -s 2>&1 | grep -o "worst error = [0-9.]* of\|drift.*\|[0-9]* passed.*\|[0-9]* failed.*"
done
