#!/bin/bash
mkdir -p gpurun_out/r2d
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d/pytest.log
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/r2d/bench.json 2> gpurun_out/r2d/bench.err; echo "bench rc=$?" >> gpurun_out/r2d/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2d/launches.csv python bench.py --images 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2d/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_views|agg_heat|ha_prep|conv_tc_kernel.*256.*80" -s 8 -c 8 -o gpurun_out/r2d/ha_fast python bench.py --images 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2d/ncu_full.log 2>&1
tail -4 gpurun_out/r2d/pytest.log; cat gpurun_out/r2d/bench.json | tail -1 | cut -c1-600; tail -3 gpurun_out/r2d/bench.err
