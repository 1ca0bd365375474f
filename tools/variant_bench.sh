#!/bin/bash
# usage: tools/variant_bench.sh lib1.so lib2.so ...   -> resident / e2e GCUPS of bench.py per library variant
for lib in "$@"; do
  PSB_LIB_PATH=$PWD/$lib python bench.py --no-cpu --steps 5 --e2e-steps 2 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$lib', round(d['value']), round(d['e2e']['value']), d['roofline']['kernel_ms_per_launch'], d['config'].get('subjects_rerun_at_32bit'))"
done
