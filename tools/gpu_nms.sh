#!/bin/bash
# NMS CTA size A/B: extraction timing pieces at several batch sizes + the exactness tests with every size
cd /root/repo; mkdir -p gpurun_out/nms
for t in 1024 512 256; do
  for b in 64 32 8 1; do
    echo "== SPB_NMS_THREADS=$t B=$b"; SPB_NMS_THREADS=$t python tools/time_extract.py $b 2>&1 | grep -E "whole|NMS"
  done
done | tee gpurun_out/nms/ab.txt
for t in 512 256; do
  SPB_NMS_THREADS=$t python -m pytest tests/test_gpu_kernels.py tests/test_gpu_fuzz.py -q -k "pts or nms or topk" 2>&1 | tail -2
done
echo "== default selection"; for b in 64 32 19 18 8 1; do python tools/time_extract.py $b 2>&1 | grep -E "whole|NMS" | tr '\n' ' '; echo " (B=$b)"; done | tee -a gpurun_out/nms/ab.txt
