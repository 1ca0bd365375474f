#!/bin/bash
mkdir -p gpurun_out/r2l
timeout 600 python tools/shard_e2e_probe.py 8 > gpurun_out/r2l/shard8.txt 2>&1; cat gpurun_out/r2l/shard8.txt | grep shard
timeout 600 python tools/shard_e2e_probe.py 4 6:24 1:1000 3:24 > gpurun_out/r2l/shard4.txt 2>&1; cat gpurun_out/r2l/shard4.txt | grep shard
PSB_DEBUG_TIMING=1 PSB_SCAN_HOST_FIRST_DIV=1 PSB_SCAN_HOST_FIRST_MB=1000 timeout 300 python tools/shard_e2e_probe.py 8 1:1000 2>&1 | grep -E "scan_host|scan job|sw16" | tail -12
