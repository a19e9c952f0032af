#!/bin/bash
mkdir -p gpurun_out/r2j
cd /root/repo
P=29700
for cfg in "default" "NCCL_MAX_CTAS=4" "NCCL_MAX_CTAS=1"; do
  P=$((P+1))
  if [ "$cfg" = "default" ]; then E=""; else E="$cfg"; fi
  env $E timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r2j/bench_n2_$P.json 2> gpurun_out/r2j/bench_n2_$P.err
  echo "== $cfg"; python - gpurun_out/r2j/bench_n2_$P.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e']['value'], {k:v for k,v in d['stage_ms_per_step'].items() if k in ('conv1b','conv2a','conv2b','warp','aggregate')}, d['clocks']['sm_mhz'])
PY
done
