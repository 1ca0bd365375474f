#!/bin/bash
# round 2, second GPU call: the GPU test-suite with the packed many-pairs kernels, the configs at reduced size
mkdir -p gpurun_out/r2b
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2b/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2b/pytest_gpu.txt
tail -15 gpurun_out/r2b/pytest_gpu.txt
timeout 900 python tests/bench_configs.py --quick --out gpurun_out/r2b/configs_quick.json > gpurun_out/r2b/configs_quick.txt 2>&1
echo "configs exit $?" >> gpurun_out/r2b/configs_quick.txt
tail -12 gpurun_out/r2b/configs_quick.txt
for w in 0 1; do
  PSB_P16_WIDE=$w timeout 300 python tests/bench_configs.py --quick --only C1,C4 --out gpurun_out/r2b/configs_wide$w.json > gpurun_out/r2b/configs_wide$w.txt 2>&1
  grep -o '"kernel_gcups": [0-9.]*' gpurun_out/r2b/configs_wide$w.txt | tr '\n' ' '; echo " (wide=$w: C1 C4)"
done
PSB_NO_P16=1 timeout 300 python tests/bench_configs.py --quick --only C1,C3,C4 --out gpurun_out/r2b/configs_nop16.json > gpurun_out/r2b/configs_nop16.txt 2>&1
grep -o '"kernel_gcups": [0-9.]*' gpurun_out/r2b/configs_nop16.txt | tr '\n' ' '; echo " (32-bit path: C1 C3 C4)"
