#!/bin/bash
mkdir -p gpurun_out/r2g
cd /root/repo
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2g/pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g/pytest.log
for wl in c3 c1 c4 c5; do
  timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 > gpurun_out/r2g/bench_$wl.json 2> gpurun_out/r2g/bench_$wl.err; echo "rc=$?" >> gpurun_out/r2g/bench_$wl.err
done
tail -30 gpurun_out/r2g/pytest.log; for wl in c3 c1 c4 c5; do cut -c1-150 gpurun_out/r2g/bench_$wl.json; tail -2 gpurun_out/r2g/bench_$wl.err; done
