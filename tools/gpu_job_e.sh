#!/bin/bash
mkdir -p gpurun_out/r2f
cd /root/repo
timeout 900 python -m pytest tests -m gpu -q -x tests/test_gpu_kernels.py tests/test_gpu_net.py -k "fast_warp or ha_ or headline" > gpurun_out/r2f/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f/pytest.log
timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu > gpurun_out/r2f/bench.json 2> gpurun_out/r2f/bench.err; echo "bench rc=$?" >> gpurun_out/r2f/bench.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"warp_views|agg_heat" -s 4 -c 4 -o gpurun_out/r2f/ha_fast python bench.py --images 8 --steps 2 --warmup 3 --no-cpu > gpurun_out/r2f/ncu_full.log 2>&1
tail -3 gpurun_out/r2f/pytest.log; cat gpurun_out/r2f/bench.json | cut -c1-200; tail -2 gpurun_out/r2f/bench.err
