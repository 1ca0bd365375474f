#!/bin/bash
# round 2 (after the pipeline rework): whole-box scaling on one 8-GPU box: psb_scan_box timeline at 8, bench.py at N = 8, 4, 2
mkdir -p gpurun_out/r2t2
timeout 300 python tools/box_timeline.py 8 auto > gpurun_out/r2t2/tl8.txt 2> gpurun_out/r2t2/tl8.err; cat gpurun_out/r2t2/tl8.txt
grep -E "pieces \(MB\)|device timeline|scan_box" gpurun_out/r2t2/tl8.err | tail -22
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2961$N bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 10 \
     > gpurun_out/r2t2/bench_${N}gpu.json 2> gpurun_out/r2t2/bench_${N}gpu.err
  echo "bench N=$N exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/r2t2/bench_${N}gpu.json')); print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), (d['config'].get('verified') or '')[:60])"
done
