#!/bin/bash
# round-2 first GPU job: baseline sanity + probes + ncu of the post-network kernels
mkdir -p gpurun_out/r2a
cd /root/repo
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > gpurun_out/r2a/smi.txt 2>&1
nproc > gpurun_out/r2a/nproc.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2a/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a/pytest.log
timeout 300 python bench.py --steps 12 --warmup 3 > gpurun_out/r2a/bench.json 2> gpurun_out/r2a/bench.err
timeout 120 python tools/probe_fused.py > gpurun_out/r2a/probe_fused.txt 2>&1
timeout 60 tools/ubench_mma_pair.bin > gpurun_out/r2a/ubench_mma_pair.txt 2>&1
timeout 120 python tools/bench_post.py > gpurun_out/r2a/bench_post.txt 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"nms_kernel|rank_scatter|sample_desc|match_tile|match_select|sample_homographies" -c 12 -o gpurun_out/r2a/post_kernels python tools/bench_post.py > gpurun_out/r2a/ncu_post.log 2>&1
ls -la gpurun_out/r2a
