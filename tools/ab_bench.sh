#!/bin/bash
# A/B helper: runs bench.py for combinations of env/flags and prints a compact summary line each.
# usage: tools/ab_bench.sh "ENV1=a ENV2=b" "--symmetric 0" ...   (pairs of env-string, flag-string)
while [ $# -gt 0 ]; do
  envs="$1"; flags="$2"; shift 2
  env $envs python bench.py --steps 4 --warmup 3 --no-cpu-baseline $flags 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('[$envs] [$flags]', round(d['frames_per_s'],1), 'fps  eval %.3e'%d['evaluated_pairs_per_s'], 'k3_ms', round(d['kernel_ms']['k3_rdf'],2), 'frac', round(d['roofline']['frac'],3), 'counted', d['check']['counted_pairs_merged'], 'fp64', d['check']['fp64_reevaluated'], 'e2e %.3e'%d['e2e']['value'], 'MHz', d['clocks']['sm_mhz'])
"
done
