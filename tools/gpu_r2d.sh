#!/bin/bash
# round 2, fourth GPU call: tests (loader, scan_box, new walk), per-kernel timing, ncu captures of the many-pairs kernels
mkdir -p gpurun_out/r2d
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2d/pytest_gpu.txt 2>&1
echo "pytest exit $?" >> gpurun_out/r2d/pytest_gpu.txt
tail -5 gpurun_out/r2d/pytest_gpu.txt
PSB_DEBUG_TIMING=1 timeout 600 python tests/bench_configs.py --quick --only C1,C3,C4 --out gpurun_out/r2d/configs_quick.json > gpurun_out/r2d/configs_quick.txt 2>&1
grep -E "^\[psb\]" gpurun_out/r2d/configs_quick.txt | sort | uniq -c | sort -rn | head -16
grep -o '"kernel_gcups": [0-9.]*' gpurun_out/r2d/configs_quick.txt | tr '\n' ' '; echo " (quick: C1 C3 C4)"
for cfg in C1 C3 C4; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:pairs16_kernel -c 1 -f -o gpurun_out/r2d/ncu_p16_$cfg \
     python tests/bench_configs.py --quick --only $cfg --out gpurun_out/r2d/ncu_$cfg.json > gpurun_out/r2d/ncu_$cfg.log 2>&1
  echo "ncu $cfg exit $?"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:walk16_kernel -c 1 -f -o gpurun_out/r2d/ncu_walk16_C4 \
   python tests/bench_configs.py --quick --only C4 --out gpurun_out/r2d/ncu_walk.json > gpurun_out/r2d/ncu_walk.log 2>&1
echo "ncu walk exit $?"
ls -la gpurun_out/r2d/*.ncu-rep
