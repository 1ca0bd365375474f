#!/bin/bash
# round 2, third GPU call: H-byte trace kernels -- tests, per-kernel timing, full-size configs, headline bench
mkdir -p gpurun_out/r2c
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2c/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2c/pytest_gpu.txt
tail -5 gpurun_out/r2c/pytest_gpu.txt
PSB_DEBUG_TIMING=1 timeout 600 python tests/bench_configs.py --quick --only C1,C3,C4 --out gpurun_out/r2c/configs_quick.json > gpurun_out/r2c/configs_quick.txt 2>&1
grep -E "^\[psb\]" gpurun_out/r2c/configs_quick.txt | sort | uniq -c | sort -rn | head -30
grep -o '"kernel_gcups": [0-9.]*' gpurun_out/r2c/configs_quick.txt | tr '\n' ' '; echo " (quick: C1 C3 C4)"
for w in 0 1; do
  PSB_P16_WIDE=$w timeout 300 python tests/bench_configs.py --quick --only C1,C4 --out gpurun_out/r2c/configs_wide$w.json > gpurun_out/r2c/configs_wide$w.txt 2>&1
  grep -o '"kernel_gcups": [0-9.]*' gpurun_out/r2c/configs_wide$w.txt | tr '\n' ' '; echo " (wide=$w: C1 C4)"
done
timeout 1500 python tests/bench_configs.py --out gpurun_out/r2c/configs_full.json > gpurun_out/r2c/configs_full.txt 2>&1
echo "full configs exit $?" >> gpurun_out/r2c/configs_full.txt
grep -E "^(C[1-5]|latency) " gpurun_out/r2c/configs_full.txt | cut -c1-420
tail -3 gpurun_out/r2c/configs_full.txt | cut -c1-300
timeout 600 python bench.py > gpurun_out/r2c/bench.json 2> gpurun_out/r2c/bench.err
cut -c1-600 gpurun_out/r2c/bench.json; tail -2 gpurun_out/r2c/bench.err
timeout 300 python tools/diff_vs_parasail.py --lib parasail_rs_b200/libparasail_b200.so --pairs 300 --out gpurun_out/r2c/diff_vs_self.json > gpurun_out/r2c/diff_vs_self.txt 2>&1
tail -2 gpurun_out/r2c/diff_vs_self.txt | cut -c1-300
