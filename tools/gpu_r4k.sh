#!/bin/bash
# round 2, final battery of the last build: GPU test suite, smoke, bench.py
mkdir -p gpurun_out/r4k
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r4k/pytest_gpu.txt 2>&1; tail -n 3 gpurun_out/r4k/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -n 2
timeout 400 python bench.py > gpurun_out/r4k/bench.json 2> gpurun_out/r4k/bench.err; python -c "
import json; d=json.load(open('gpurun_out/r4k/bench.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value']), d['clocks'], 'launches', d['gpu_launches'])"
