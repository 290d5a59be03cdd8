#!/bin/bash
# A/B of library builds through bench.py itself (the step as the driver times it): every build under build/variants/
# and the in-tree one, REPS times interleaved; prints ms_per_step and the trace-kernel time of each run.
#   tools/ab_bench.sh "cfg3 cfg2" 2
WLS=${1:-"cfg3 cfg2"}
REPS=${2:-2}
for rep in $(seq $REPS); do
  for lib in build/variants/*.so attosecondraytracing_b200/libart_b200.so; do
    for wl in $WLS; do
      ART_B200_LIB=$PWD/$lib python bench.py --workload $wl --sub none --steps 30 --warmup 5 --no-cpu-baseline --no-port-baseline 2>/dev/null |
        python -c "
import json,sys
d=json.loads(sys.stdin.readline())
print('$(basename $lib)', '$wl', 'step_ms', round(d['ms_per_step'],4), 'k1_ms', round(d['roofline']['kernel_ms'],4), 'e2e_ms', round(d['e2e']['ms_per_step'],4))"
    done
  done
done
