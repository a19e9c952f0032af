#!/bin/bash
# evidence job (bash tools/gpu_evidence.sh under gpurun): ncu --set full of one launch of every kernel on the path, summarised ON the box (the .ncu-rep files
# stay in /tmp: gpurun_out/ is limited to 64 MiB), + the launch list of a bench step
mkdir -p gpurun_out/r2k /tmp/ncu
cd /root/repo
timeout 900 ncu --set full --clock-control none --import-source on -o /tmp/ncu/post_kernels python tools/bench_post.py --once > gpurun_out/r2k/ncu_post.log 2>&1
python tools/ncu_summary.py /tmp/ncu/post_kernels.ncu-rep tools/alg_bytes.json > gpurun_out/r2k/ncu_post_kernels.txt 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused|conv_s3|conv_tc_kernel|conv_tc2|warp_views|agg_heat|agg_reduce|ha_prep|finalize" --launch-skip 45 -c 16 -o /tmp/ncu/ha_step python bench.py --images 8 --steps 1 --warmup 3 --no-cpu > gpurun_out/r2k/ncu_step.log 2>&1
python tools/ncu_summary.py /tmp/ncu/ha_step.ncu-rep tools/alg_bytes.json > gpurun_out/r2k/ncu_ha_step.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2k/launches.csv python bench.py --images 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2k/ncu_launch.log 2>&1
timeout 300 python tools/bench_post.py > gpurun_out/r2k/bench_post.txt 2>&1
grep -c "duration" gpurun_out/r2k/ncu_post_kernels.txt gpurun_out/r2k/ncu_ha_step.txt; ls -la gpurun_out/r2k
