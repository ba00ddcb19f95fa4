#!/bin/bash
# run the per-stage profile for a set of env-var kernel variants:  tools/tune.sh "A=1 B=2" "A=2" ...
for v in "$@"; do
  echo "== $v"
  env $v python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), {k:round(x,4) for k,x in d['stage_ms_per_slice'].items()}, d['gpu_launches'])"
done
