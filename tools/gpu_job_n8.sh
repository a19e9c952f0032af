#!/bin/bash
# 8-GPU runs: the scaling workload (c2) + the other HA configs, with NCCL's communicator lines kept in the .err files
mkdir -p gpurun_out/r2h
cd /root/repo
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
P=29600
for wl in c2 c4 c5; do
  P=$((P+1))
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 8 --steps 12 --warmup 3 --workload $wl > gpurun_out/r2h/bench_n8_$wl.json 2> gpurun_out/r2h/bench_n8_$wl.err; echo "rc=$?" >> gpurun_out/r2h/bench_n8_$wl.err
done
P=$((P+1))
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $P bench.py --gpus 4 --steps 12 --warmup 3 > gpurun_out/r2h/bench_n4_c2.json 2> gpurun_out/r2h/bench_n4_c2.err; echo "rc=$?" >> gpurun_out/r2h/bench_n4_c2.err
for f in gpurun_out/r2h/*.json; do echo $f; cut -c1-160 $f; done; grep -h "nranks 8" gpurun_out/r2h/bench_n8_c2.err | head -2; tail -1 gpurun_out/r2h/*.err
