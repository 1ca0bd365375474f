#!/bin/bash
# round-2 late check: local end cell of the wavefront kernel, keyed (branch-free) against late (located after the tile), A/B/A
mkdir -p gpurun_out/r4g
for lib in parasail_rs_b200/libparasail_b200.so variants/lib_wavekey.so parasail_rs_b200/libparasail_b200.so variants/lib_wavekey.so; do
  echo "== $lib" >> gpurun_out/r4g/probe.log
  PROBE_MODES=local PSB_LIB_PATH=$PWD/$lib timeout 100 python tools/long_trace_probe.py 100000 >> gpurun_out/r4g/probe.log 2>> gpurun_out/r4g/probe.err
done
cut -c1-330 gpurun_out/r4g/probe.log
