#!/bin/bash
mkdir -p gpurun_out/pair
cd /root/repo
rm -f gpurun_out/pair/*
SPB_C1P_VARIANT=7 timeout 300 python tools/probe_fused.py 2 > gpurun_out/pair/probe_follower_v3.txt 2>&1
SPB_C1P_VARIANT=3 timeout 300 python tools/probe_fused.py 2 > gpurun_out/pair/probe_leader_v3.txt 2>&1
cat gpurun_out/pair/probe_follower_v3.txt | sed 's/s1.acc1_full=-[0-9]* //'
