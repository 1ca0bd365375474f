#!/bin/bash
# round 2 final battery on one GPU: the GPU test suite, smoke, the config-size benchmark (C1, C3, C4, C5 + latency),
# bench.py both arms, the ncu launch list of the bench command
mkdir -p gpurun_out/r3c
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r3c/pytest_gpu.txt 2>&1; tail -3 gpurun_out/r3c/pytest_gpu.txt
timeout 200 python -c "import __graft_entry__ as g; g.build(); g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python tests/bench_configs.py --out gpurun_out/r3c/configs.json > gpurun_out/r3c/configs.log 2>&1; grep -E "^C[0-9]|^latency|Traceback|Error" gpurun_out/r3c/configs.log | cut -c1-420
timeout 600 python bench.py > gpurun_out/r3c/bench.json 2> gpurun_out/r3c/bench.err; cut -c1-1100 gpurun_out/r3c/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r3c/bench_ref.json 2>> gpurun_out/r3c/bench.err; cut -c1-300 gpurun_out/r3c/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r3c/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r3c/ncu_launches.log 2>&1
echo "ncu launches exit $?"
