#!/bin/bash
# run the per-stage profile for a set of option variants:  tools/tune.sh "order=9" "fuse=0 side_stream=0" ...
# (each word becomes a bench.py --opt KEY=VALUE; "-" is the default configuration)
for v in "$@"; do
  echo "== $v"
  opts=""
  if [ "$v" != "-" ]; then for kv in $v; do opts="$opts --opt $kv"; done; fi
  python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e $opts $TUNE_ARGS 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value'],1), {k:round(x,4) for k,x in d['stage_ms_per_slice'].items()}, d['gpu_launches'])"
done
