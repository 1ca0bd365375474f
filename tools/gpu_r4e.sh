#!/bin/bash
# round-2 late check: all-lanes-active (single basic block) form of the wavefront step
mkdir -p gpurun_out/r4e
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_pair or long_local" > gpurun_out/r4e/pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r4e/pytest.txt
tail -4 gpurun_out/r4e/pytest.txt
timeout 200 python tools/long_trace_probe.py 100000 > gpurun_out/r4e/probe.log 2> gpurun_out/r4e/probe.err
echo "probe rc $?"
cut -c1-420 gpurun_out/r4e/probe.log
PSB_DEBUG_TIMING=1 timeout 100 python tools/long_trace_probe.py 50000 2>&1 | grep "walk32\|timed region" | head -12 > gpurun_out/r4e/split.txt
cat gpurun_out/r4e/split.txt
