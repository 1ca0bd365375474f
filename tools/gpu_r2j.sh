#!/bin/bash
# round 2: full GPU suite, configs at config size, headline bench, ncu launch list + full capture of the scan kernel
mkdir -p gpurun_out/r2j
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2j/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2j/pytest_gpu.txt
tail -4 gpurun_out/r2j/pytest_gpu.txt
PSB_DEBUG_TIMING=1 timeout 1500 python tests/bench_configs.py --only C1,C3,C4,latency --out gpurun_out/r2j/r2_configs.json > gpurun_out/r2j/configs_full.txt 2>&1
echo "configs exit $?"
grep -E "^(C[1-5]|latency) " gpurun_out/r2j/configs_full.txt | cut -c1-300
grep -E "^\[psb\] (pairs16|walk16|pass)" gpurun_out/r2j/configs_full.txt | sort | uniq -c | sort -rn | head -14
timeout 600 python bench.py > gpurun_out/r2j/bench_1gpu.json 2> gpurun_out/r2j/bench_1gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r2j/bench_1gpu.json')); print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'frac', round(d['roofline']['frac'],3), 'cpu', round(d['cpu_baseline']['value']), d['cpu_baseline']['cores'], d['clocks'])"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2j/bench_ref.json 2> gpurun_out/r2j/bench_ref.err
cut -c1-300 gpurun_out/r2j/bench_ref.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2j/ncu_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --e2e-steps 1 > gpurun_out/r2j/ncu_launches.log 2>&1
echo "ncu launches exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sw16_scan_kernel -c 2 -f -o gpurun_out/r2j/ncu_sw16_C2 python bench.py --steps 1 --warmup 3 --no-cpu --e2e-steps 1 --db 200000 > gpurun_out/r2j/ncu_sw16.log 2>&1
echo "ncu sw16 exit $?"
ls -la gpurun_out/r2j/
