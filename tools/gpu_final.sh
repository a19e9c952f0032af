#!/bin/bash
# final evidence of the round: bench lines (all workloads, N = 1) + reference arm + ncu summaries
mkdir -p gpurun_out/final /tmp/ncu
cd /root/repo
timeout 600 python bench.py --steps 12 --warmup 3 > gpurun_out/final/bench_n1.json 2> gpurun_out/final/bench_n1.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final/bench_ref_n1.json 2> gpurun_out/final/bench_ref_n1.err
for wl in c1 c3 c4 c5; do timeout 600 python bench.py --workload $wl --steps 8 --warmup 3 > gpurun_out/final/bench_n1_$wl.json 2> gpurun_out/final/bench_n1_$wl.err; done
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"conv1_fused|conv_s3|conv_tc_kernel|conv_tc2|warp_views|agg_heat|agg_reduce|ha_prep|finalize" --launch-skip 45 -c 16 -o /tmp/ncu/ha_step python bench.py --images 8 --steps 1 --warmup 3 --no-cpu > gpurun_out/final/ncu_step.log 2>&1
python tools/ncu_summary.py /tmp/ncu/ha_step.ncu-rep tools/alg_bytes.json > gpurun_out/final/ncu_ha_step.txt 2>&1
for f in gpurun_out/final/bench_*.json; do echo $f; cut -c1-110 $f; done
