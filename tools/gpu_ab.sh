#!/bin/bash
# usage: tools/gpu_ab.sh <streams> "<ENV=.. ENV=..>" ["<ENV..>" ...]  -> one line per configuration (ms/step, RTFx, e2e)
streams=$1; shift
for cfg in "$@"; do
  env $cfg timeout 300 python bench.py --streams $streams --steps ${STEPS:-60} --warmup 10 --no-cpu-baseline 2>/dev/null | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$streams', '$cfg', round(d['ms_per_step'],4), round(d['value']), round(d['e2e']['value']))"
done
