#!/bin/bash
# round-2 late diagnostic: per-strip timeline of the wavefront kernel (claimed / done), 100 kb local and global
mkdir -p gpurun_out/r4j
for m in local global_; do
  echo "== $m" >> gpurun_out/r4j/timeline.log
  PSB_DEBUG_TIMING=1 timeout 100 python tools/wave_ncu_probe.py 100000 $m score >> gpurun_out/r4j/timeline.log 2>&1
done
grep -v "pass of" gpurun_out/r4j/timeline.log
