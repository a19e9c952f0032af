#!/bin/bash
mkdir -p gpurun_out/r2m
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2m/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2m/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r2m/smoke.log
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/r2m/bench.json 2> gpurun_out/r2m/bench.err; echo "bench rc=$?" >> gpurun_out/r2m/bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2m/bench_ref.json 2> gpurun_out/r2m/bench_ref.err; echo "ref rc=$?" >> gpurun_out/r2m/bench_ref.err
tail -4 gpurun_out/r2m/pytest.log; tail -2 gpurun_out/r2m/smoke.log; cut -c1-120 gpurun_out/r2m/bench.json; cut -c1-160 gpurun_out/r2m/bench_ref.json; tail -1 gpurun_out/r2m/bench.err gpurun_out/r2m/bench_ref.err
