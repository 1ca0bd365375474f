#!/bin/bash
mkdir -p gpurun_out/r2i
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wave32v3 -c 1 -f -o gpurun_out/r2i/ncu_wave32v3_C5 python tools/c5_probe.py 8:4 > gpurun_out/r2i/ncu_wave.log 2>&1
echo "ncu exit $?"; ls -la gpurun_out/r2i/
