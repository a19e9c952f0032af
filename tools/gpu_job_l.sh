#!/bin/bash
mkdir -p gpurun_out/r2p
cd /root/repo
for c in 0 400 200 1600; do
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu --chunk $c > gpurun_out/r2p/bench_chunk$c.json 2> gpurun_out/r2p/bench_chunk$c.err
  python - gpurun_out/r2p/bench_chunk$c.json $c <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('chunk', sys.argv[2], d['value'], d['ms_per_step'], d['e2e']['value'], {k:v for k,v in d['stage_ms_per_step'].items() if k in ('conv1b','conv2a','conv3b','warp','aggregate')}, d['clocks']['sm_mhz'])
PY
done
