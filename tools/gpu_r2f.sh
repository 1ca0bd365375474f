#!/bin/bash
# multi-GPU check: bench.py under torchrun at N GPUs (resident scan per rank + e2e through psb_scan_box from rank 0)
N=${1:-2}
mkdir -p gpurun_out/r2f
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
   > gpurun_out/r2f/bench_${N}gpu.json 2> gpurun_out/r2f/bench_${N}gpu.err
echo "bench exit $?"
python -c "
import json; d=json.load(open('gpurun_out/r2f/bench_${N}gpu.json')); print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), d['config'].get('verified'))"
tail -5 gpurun_out/r2f/bench_${N}gpu.err
PSB_DEBUG_TIMING=1 timeout 300 python - <<'PY' > gpurun_out/r2f/scan_box_probe_${N}.txt 2>&1
import sys, time
sys.path.insert(0, 'tests')
import numpy as np, torch, bench, psb_data
import parasail_rs_b200 as ps
query, cat, off = bench.make_inputs(1000000)
pc = torch.empty(len(cat), dtype=torch.uint8, pin_memory=True); pc.numpy()[:] = cat
po = torch.empty(len(off), dtype=torch.int64, pin_memory=True); po.numpy()[:] = off
b62 = ps.Matrix.from_name('blosum62')
a = ps.Aligner.new().local().gap_open(10).gap_extend(1).profile(ps.Profile.new(query, False, b62)).build()
for ng in (1, 2, 4, 8):
    if ng > torch.cuda.device_count(): break
    for _ in range(3): a.scan_box((pc.numpy(), po.numpy()), ng)
    t0 = time.perf_counter()
    for _ in range(5): r = a.scan_box((pc.numpy(), po.numpy()), ng)
    dt = (time.perf_counter() - t0) / 5
    print(f"scan_box n_gpus {ng}: {dt*1e3:.3f} ms per call, {400*float(off[-1])/dt/1e9:.0f} GCUPS, kernel max {ps.kernel_ms():.3f} ms", flush=True)
PY
grep "scan_box n_gpus" gpurun_out/r2f/scan_box_probe_${N}.txt
