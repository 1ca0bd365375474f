#!/bin/bash
# round 2, sixth GPU call: reworked wavefront cell (C5), scan_box timing, tests
mkdir -p gpurun_out/r2g
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2g/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2g/pytest_gpu.txt
tail -4 gpurun_out/r2g/pytest_gpu.txt
timeout 900 python tests/bench_configs.py --only C5,latency --out gpurun_out/r2g/configs_c5.json > gpurun_out/r2g/configs_c5.txt 2>&1
grep -E "^(C5|latency) " gpurun_out/r2g/configs_c5.txt | cut -c1-330
timeout 600 python tools/c5_probe.py 8:4 8:8 4:4 4:8 16:4 16:2 > gpurun_out/r2g/c5_probe.txt 2>&1
cat gpurun_out/r2g/c5_probe.txt
