#!/bin/bash
# round-2 late check: local end cell of the wavefront kernel found after the tile (tile maximum only inside it)
mkdir -p gpurun_out/r4f
timeout 200 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "long_pair or long_local" > gpurun_out/r4f/pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r4f/pytest.txt
tail -n 4 gpurun_out/r4f/pytest.txt
timeout 200 python tools/long_trace_probe.py 100000 > gpurun_out/r4f/probe.log 2> gpurun_out/r4f/probe.err
echo "probe rc $?"
cut -c1-420 gpurun_out/r4f/probe.log
