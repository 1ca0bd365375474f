#!/bin/bash
# round-2 late check: long pairs with traceback / statistics on the wavefront kernel (tests + probe)
mkdir -p gpurun_out/r4a
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_pair" > gpurun_out/r4a/pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r4a/pytest.txt
tail -15 gpurun_out/r4a/pytest.txt
timeout 200 python tools/long_trace_probe.py 50000 100000 > gpurun_out/r4a/probe.log 2> gpurun_out/r4a/probe.err
echo "probe rc $?"
cat gpurun_out/r4a/probe.log | cut -c1-900
tail -5 gpurun_out/r4a/probe.err
