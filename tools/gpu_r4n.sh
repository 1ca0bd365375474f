#!/bin/bash
# round-2 late check: walk_trace_kernel asks for the path's line six diagonal steps ahead: parity + single-pair latency
mkdir -p gpurun_out/r4n
timeout 100 python -m pytest tests/test_gpu_parity.py tests/test_gpu_reference_suite.py -x -q -m gpu -k "single_long_pair_trace_api or single_pair_api or reference or ssw or trace_cigar" > gpurun_out/r4n/pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r4n/pytest.txt
tail -n 6 gpurun_out/r4n/pytest.txt
timeout 60 python tests/bench_configs.py --quick --only latency --out gpurun_out/r4n/latency.json 2>&1 | tail -n 3 | cut -c1-600
