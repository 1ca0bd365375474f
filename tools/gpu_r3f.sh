#!/bin/bash
# round 2, last battery: GPU test suite, smoke, C1/C3/C4 at config size, pairs probe, bench.py
mkdir -p gpurun_out/r3f
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3f/pytest_gpu.txt 2>&1; tail -3 gpurun_out/r3f/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python tests/bench_configs.py --only C1,C3,C4 --out gpurun_out/r3f/configs.json > gpurun_out/r3f/configs.log 2>&1; grep -E "^C[0-9]|Traceback|Error" gpurun_out/r3f/configs.log | cut -c1-330
python tools/pairs_host_probe.py 1000000 2>/dev/null | grep -v "one lane"
timeout 600 python bench.py > gpurun_out/r3f/bench.json 2> gpurun_out/r3f/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r3f/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value']), d['clocks'], 'launches', d['gpu_launches'])"
