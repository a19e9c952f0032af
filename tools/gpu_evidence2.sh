#!/bin/bash
# evidence refresh at the end of the round: ncu --set full of the post-network kernels (incl. nms2_kernel), launch list of a bench step
mkdir -p gpurun_out/r2k2 /tmp/ncu
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -o /tmp/ncu/post_kernels python tools/bench_post.py --once > gpurun_out/r2k2/ncu_post.log 2>&1
python tools/ncu_summary.py /tmp/ncu/post_kernels.ncu-rep tools/alg_bytes.json > gpurun_out/r2k2/ncu_post_kernels.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k2/launches.csv python bench.py --images 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2k2/ncu_launch.log 2>&1
timeout 300 python tools/bench_post.py > gpurun_out/r2k2/bench_post.txt 2>&1
grep -c "duration" gpurun_out/r2k2/ncu_post_kernels.txt; wc -l gpurun_out/r2k2/launches.csv; cat gpurun_out/r2k2/bench_post.txt
