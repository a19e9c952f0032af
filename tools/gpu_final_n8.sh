#!/bin/bash
# final multi-GPU evidence of the round: bench lines at N ranks for c2 (with NCCL's INIT lines), c4, c5
N=${1:-8}
cd /root/repo; mkdir -p gpurun_out/final$N
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus $N "$@"; }
NCCL_DEBUG=INFO run > gpurun_out/final$N/bench_n$N.json 2> gpurun_out/final$N/bench_n$N.err; echo "c2 rc=$?"
run --impl reference --steps 1 --warmup 0 > gpurun_out/final$N/bench_ref_n$N.json 2> gpurun_out/final$N/bench_ref_n$N.err; echo "ref rc=$?"
for wl in c4 c5; do run --workload $wl --no-cpu > gpurun_out/final$N/bench_n${N}_$wl.json 2> gpurun_out/final$N/bench_n${N}_$wl.err; echo "$wl rc=$?"; done
grep nranks gpurun_out/final$N/bench_n$N.err | head -8 > gpurun_out/final$N/nccl_init.txt
rm -f gpurun_out/final$N/*.err.big
for f in gpurun_out/final$N/bench_n$N.json gpurun_out/final$N/bench_n${N}_c4.json gpurun_out/final$N/bench_n${N}_c5.json; do python -c "
import json;d=json.load(open('$f'));print('$f', d['value'], d['e2e']['value'], d['config'].get('exchange'), d.get('shard_check',{}).get('heat_max_abs'), d.get('image_sharded',{}).get('value'), d.get('matches_per_pair'))"; done
wc -l gpurun_out/final$N/bench_n$N.json; wc -l gpurun_out/final$N/nccl_init.txt; head -c 300 gpurun_out/final$N/bench_ref_n$N.json
# keep the merged output small
for f in gpurun_out/final$N/*.err; do tail -c 20000 $f > $f.tail; rm $f; done
