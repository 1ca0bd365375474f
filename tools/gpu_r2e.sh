#!/bin/bash
# round 2, fifth GPU call: strip-wise long-query scan, biased nw/sg kernels, headline bench
mkdir -p gpurun_out/r2e
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2e/pytest_gpu.txt
tail -5 gpurun_out/r2e/pytest_gpu.txt
PSB_DEBUG_TIMING=1 timeout 600 python tests/bench_configs.py --quick --only C1,C3,C4 --out gpurun_out/r2e/configs_quick.json > gpurun_out/r2e/configs_quick.txt 2>&1
grep -E "^\[psb\]" gpurun_out/r2e/configs_quick.txt | sort | uniq -c | sort -rn | head -12
grep -o '"kernel_gcups": [0-9.]*' gpurun_out/r2e/configs_quick.txt | tr '\n' ' '; echo " (quick: C1 C3 C4)"
timeout 600 python tools/scan_lq_probe.py 100 256 400 401 448 512 700 1000 1500 > gpurun_out/r2e/scan_lq_probe.txt 2>&1
cat gpurun_out/r2e/scan_lq_probe.txt
timeout 600 python bench.py > gpurun_out/r2e/bench.json 2> gpurun_out/r2e/bench.err
python -c "
import json; d=json.load(open('gpurun_out/r2e/bench.json')); print(round(d['value']), round(d['e2e']['value']), d['roofline']['frac'], d['config'].get('verified'))"
tail -2 gpurun_out/r2e/bench.err
