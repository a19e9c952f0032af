#!/bin/bash
# NMS A/B: scan kernel (SPB_NMS_ALGO=0) vs separable window maxima (1): exactness tests + extraction timing pieces
cd /root/repo; mkdir -p gpurun_out/nms
SPB_FUZZ_ITERS=100 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py tests/test_gpu_frontend.py -q -x -k "pts or nms or topk or export or extract or frontend" 2>&1 | tail -3
SPB_NMS_THREADS=512 SPB_FUZZ_ITERS=30 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -q -x -k "pts or nms or topk" 2>&1 | tail -2
for a in 0 1; do
  for b in 64 32 8 1; do
    echo "== SPB_NMS_ALGO=$a B=$b"; SPB_NMS_ALGO=$a python tools/time_extract.py $b 2>&1 | grep -E "whole|NMS"
  done
done | tee gpurun_out/nms/ab_algo.txt
SPB_NMS_ALGO=1 python tools/bench_post.py 2>&1 | head -2 | tee -a gpurun_out/nms/ab_algo.txt
SPB_NMS_ALGO=0 python tools/bench_post.py 2>&1 | head -2 | tee -a gpurun_out/nms/ab_algo.txt
