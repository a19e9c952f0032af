#!/bin/bash
# A/B of two builds of the library on the bench step (alternating, two rounds)
mkdir -p gpurun_out/r2l
cd /root/repo
for r in 1 2; do
for v in base hint; do
  SPB200_LIB=/root/repo/pytorch_superpoint_b200/libspb200_$v.so timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/r2l/bench_${v}_$r.json 2> gpurun_out/r2l/bench_${v}_$r.err
  python - gpurun_out/r2l/bench_${v}_$r.json $v <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(sys.argv[2], d['value'], d['ms_per_step'], {k:v for k,v in d['stage_ms_per_step'].items() if k in ('conv1b','conv2a','conv3b')}, d['clocks']['sm_mhz'])
PY
done; done
