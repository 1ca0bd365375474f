#!/bin/bash
mkdir -p gpurun_out/r2h
timeout 600 python tools/c5_probe.py 8:4 16:4 4:4 > gpurun_out/r2h/c5_probe.txt 2>&1
cat gpurun_out/r2h/c5_probe.txt
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_pair or wavefront or column_blocked" > gpurun_out/r2h/pytest_wave.txt 2>&1; tail -2 gpurun_out/r2h/pytest_wave.txt
