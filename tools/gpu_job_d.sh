#!/bin/bash
# 2-GPU functional check of the sharded path (reduce-scatter, async exchange, owned images) + bench at N=2
mkdir -p gpurun_out/r2e
cd /root/repo
export NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_export_multigpu.py /tmp/mg_out > gpurun_out/r2e/check.log 2>&1; echo "check rc=$?" >> gpurun_out/r2e/check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 12 --warmup 3 > gpurun_out/r2e/bench_n2.out 2> gpurun_out/r2e/bench_n2.err; echo "bench rc=$?" >> gpurun_out/r2e/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/r2e/bench_ref_n2.out 2> gpurun_out/r2e/bench_ref_n2.err; echo "ref rc=$?" >> gpurun_out/r2e/bench_ref_n2.err
grep -v NCCL gpurun_out/r2e/check.log | tail -5; grep "^{" gpurun_out/r2e/bench_n2.out | cut -c1-300; tail -2 gpurun_out/r2e/bench_n2.err; grep -c "NCCL INFO" gpurun_out/r2e/bench_n2.out gpurun_out/r2e/bench_n2.err; grep "^{" gpurun_out/r2e/bench_ref_n2.out | cut -c1-200
