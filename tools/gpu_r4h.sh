#!/bin/bash
# round-2 late check: all-lanes-active step for local launches (late / keyed end cell) against the default build;
# ncu --set full of the global score-only kernel (single-block step) and of a traced local launch
mkdir -p gpurun_out/r4h
for lib in parasail_rs_b200/libparasail_b200.so variants/lib_wavefl.so variants/lib_wavefk.so parasail_rs_b200/libparasail_b200.so variants/lib_wavefl.so; do
  echo "== $lib" >> gpurun_out/r4h/probe.log
  PROBE_MODES=local PSB_LIB_PATH=$PWD/$lib timeout 100 python tools/long_trace_probe.py 100000 >> gpurun_out/r4h/probe.log 2>> gpurun_out/r4h/probe.err
done
cut -c1-200 gpurun_out/r4h/probe.log
timeout 150 ncu --set full --clock-control none --import-source on -k regex:wave32v3 -c 1 -f -o gpurun_out/r4h/ncu_wave32v3_nw_100k python tools/wave_ncu_probe.py 100000 global_ score > gpurun_out/r4h/ncu_nw.log 2>&1
echo "ncu nw exit $?"
timeout 150 ncu --set full --clock-control none --import-source on -k regex:wave32v3 -c 1 -f -o gpurun_out/r4h/ncu_wave32v3_sw_trace_50k python tools/wave_ncu_probe.py 50000 local trace > gpurun_out/r4h/ncu_swtrace.log 2>&1
echo "ncu sw trace exit $?"
ls -la gpurun_out/r4h/
