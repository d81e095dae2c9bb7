#!/bin/bash
# A/B of an experiment build of the library against the product build on the composite-trial workloads (developer tool).
# Usage: bash tools/tune_sh.sh polymer-stats_b200/libpolymc_b200_sh.so [workloads...]
alt=$1; shift
wl=${@:-K1 K5 K3}
for w in $wl; do
  for lib in "" "$alt"; do
    v=$(PMC_LIB_PATH=$lib python bench.py --workload $w --steps 5 --warmup 3 --no-cpu-baseline --no-e2e --no-extras 2>/dev/null |
        python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.4g %s frac=%.3f' % (d['value'], d['roofline']['kernel'], d['roofline']['frac']))")
    echo "$w ${lib:-product}: $v"
  done
done
