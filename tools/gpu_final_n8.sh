#!/bin/bash
mkdir -p gpurun_out/final8
cd /root/repo
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus 8 --steps 12 --warmup 3 > gpurun_out/final8/bench_n8.json 2> gpurun_out/final8/bench_n8.err; echo "rc=$?" >> gpurun_out/final8/bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29802 bench.py --gpus 8 --steps 2 --warmup 1 --impl reference > gpurun_out/final8/bench_ref_n8.json 2> gpurun_out/final8/bench_ref_n8.err; echo "rc=$?" >> gpurun_out/final8/bench_ref_n8.err
cut -c1-200 gpurun_out/final8/bench_n8.json; cut -c1-160 gpurun_out/final8/bench_ref_n8.json; grep -c "nranks 8" gpurun_out/final8/bench_n8.err; wc -l gpurun_out/final8/bench_n8.json
