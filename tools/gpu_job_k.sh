#!/bin/bash
mkdir -p gpurun_out/r2n
cd /root/repo
timeout 600 python tools/check_heat_error.py > gpurun_out/r2n/heat_error.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/r2n/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n/pytest.log
for r in 1 2; do for t in 0 1; do
  timeout 300 python bench.py --steps 12 --warmup 3 --no-cpu --agg-tex $t > gpurun_out/r2n/bench_tex${t}_$r.json 2> gpurun_out/r2n/bench_tex${t}_$r.err
  python - gpurun_out/r2n/bench_tex${t}_$r.json $t <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('tex', sys.argv[2], d['value'], d['ms_per_step'], {k:v for k,v in d['stage_ms_per_step'].items() if k in ('conv1b','convPb','warp','aggregate')}, d['clocks']['sm_mhz'])
PY
done; done
grep -v Warning gpurun_out/r2n/heat_error.txt | tail -3; tail -5 gpurun_out/r2n/pytest.log
