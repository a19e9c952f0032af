#!/bin/bash
# NMS: exactness tests + extraction timing pieces (SPB_NMS_ALGO=0: scan kernel, 1: separable window maxima)
cd /root/repo; mkdir -p gpurun_out/nms
SPB_FUZZ_ITERS=150 timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py tests/test_gpu_frontend.py -q -x -k "pts or nms or topk or export or extract or frontend" 2>&1 | tail -3
SPB_NMS_THREADS=512 SPB_FUZZ_ITERS=40 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -q -x -k "pts or nms or topk" 2>&1 | tail -2
SPB_NMS_THREADS=256 SPB_FUZZ_ITERS=40 timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -q -x -k "pts or nms or topk" 2>&1 | tail -2
for b in 64 32 8 1; do
  echo "== B=$b"; python tools/time_extract.py $b 2>&1 | grep -E "whole|NMS"
done | tee gpurun_out/nms/ab_shfl.txt
python tools/bench_post.py 2>&1 | head -2 | tee -a gpurun_out/nms/ab_shfl.txt
