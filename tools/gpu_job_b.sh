#!/bin/bash
mkdir -p gpurun_out/r2c
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_frontend.py::test_export_headline_config_vs_oracle > gpurun_out/r2c/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest.log
timeout 600 python -m pytest tests -m gpu -q tests/test_gpu_frontend.py::test_export_headline_config_vs_oracle tests/test_gpu_kernels.py -k "fast_warp or headline" > gpurun_out/r2c/pytest2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c/pytest2.log
timeout 600 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_net.py -q -x -k "ha_end_to_end or ragged or other_config" > gpurun_out/r2c/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2c/memcheck.log
tail -5 gpurun_out/r2c/pytest.log gpurun_out/r2c/pytest2.log; tail -15 gpurun_out/r2c/memcheck.log
