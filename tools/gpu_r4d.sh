#!/bin/bash
# round-2 late check: lazy row-major trace table of the single-pair API; single-pair latency
mkdir -p gpurun_out/r4d
timeout 200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py -x -q -m gpu -k "single_pair or reference or trace or ssw" > gpurun_out/r4d/pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r4d/pytest.txt
tail -4 gpurun_out/r4d/pytest.txt
timeout 200 python tests/bench_configs.py --quick --only latency --out gpurun_out/r4d/latency.json 2>&1 | tail -3
