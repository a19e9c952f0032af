#!/bin/bash
mkdir -p gpurun_out/r2i
cd /root/repo
timeout 600 python -m pytest tests/test_gpu_net.py -m gpu -q -x -k "conv1_ws or conv1_fused" > gpurun_out/r2i/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2i/pytest.log
timeout 300 python tools/ab_fuse1.py 2 800 > gpurun_out/r2i/ab.txt 2>&1
timeout 300 python tools/ab_fuse1.py 2 800 >> gpurun_out/r2i/ab.txt 2>&1
tail -5 gpurun_out/r2i/pytest.log; cat gpurun_out/r2i/ab.txt
