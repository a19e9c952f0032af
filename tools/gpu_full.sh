#!/bin/bash
# the whole GPU suite + smoke + the timing of the train-mode BatchNorm path
mkdir -p gpurun_out/full
cd /root/repo
python -m pytest tests -m gpu -q -x 2>&1 | tail -8 | tee gpurun_out/full/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/full/smoke.txt
python tools/bench_bn_batch.py 2>&1 | tee gpurun_out/full/bn_batch.txt
python tools/bench_bn_batch.py 100 480 640 2>&1 | tee -a gpurun_out/full/bn_batch.txt
