#!/bin/bash
# usage: ab_env.sh "ENV=VAL ENV2=VAL" "..." : one short bench per environment setting
for spec in "$@"; do
  env $spec python bench.py --steps 10 --warmup 3 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.load(sys.stdin)
print('$spec |', round(d['value']), 'inst-iter/s', round(d['ms_per_step'],2), 'ms/step', d['workload_stats']['status_histogram_rank0'], round(d['workload_stats']['mean_rollouts_per_iteration'],3), 'e2e', round(d['e2e']['value']))"
done
