#!/bin/bash
# round-2 late check: 4 rows per lane (twice the strips, shorter steps) against 8 for pairs whose strips all find a scheduler
mkdir -p gpurun_out/r4i
for L in 20000 50000 74000 100000; do
  echo "== local L=$L" >> gpurun_out/r4i/k_probe.log
  C5_LEN=$L timeout 100 python tools/c5_probe.py 8:4 4:4 >> gpurun_out/r4i/k_probe.log 2>&1
done
for L in 50000 100000; do
  for K in 8 4; do
    echo "== nw/sg L=$L K=$K" >> gpurun_out/r4i/k_probe.log
    NW_LEN=$L PSB_WAVE_K=$K timeout 100 python tools/nw_long_probe.py >> gpurun_out/r4i/k_probe.log 2>&1
  done
done
cat gpurun_out/r4i/k_probe.log
