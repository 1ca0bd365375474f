#!/bin/bash
# round 2 final battery on one GPU: the GPU test suite, smoke, the config-size benchmark (C1, C3, C4, C5 + latency), bench.py
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z/pytest_gpu.txt 2>&1; tail -3 gpurun_out/r2z/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python tests/bench_configs.py --out gpurun_out/r2z/configs.json > gpurun_out/r2z/configs.log 2>&1; grep -E "^C[0-9]|^latency|Traceback|Error" gpurun_out/r2z/configs.log | cut -c1-420
timeout 600 python bench.py > gpurun_out/r2z/bench.json 2> gpurun_out/r2z/bench.err; cat gpurun_out/r2z/bench.json | cut -c1-1500
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z/bench_ref.json 2>> gpurun_out/r2z/bench.err; cat gpurun_out/r2z/bench_ref.json | cut -c1-600
