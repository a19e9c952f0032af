#!/bin/bash
# peer-memory exchange vs NCCL reduce-scatter: bench line at N ranks, both ways + the .npz export check
N=${1:-2}
mkdir -p gpurun_out/peer
cd /root/repo
for ex in peer nccl peer nccl; do
  SPB200_EXCHANGE=$ex timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --no-cpu > gpurun_out/peer/bench_n${N}_$ex.json 2> gpurun_out/peer/bench_n${N}_$ex.err
  echo "== $ex rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/peer/bench_n${N}_$ex.json'));print('$ex', d['value'], d['e2e']['value'], d.get('shard_check'), d.get('image_sharded',{}).get('value'))" || tail -20 gpurun_out/peer/bench_n${N}_$ex.err
done
SPB200_EXCHANGE=peer timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29514 tools/check_export_multigpu.py /tmp/peer_out_$N > gpurun_out/peer/check_export_n$N.txt 2>&1; echo "check rc=$?"; tail -5 gpurun_out/peer/check_export_n$N.txt
