#!/bin/bash
# round 2: whole-box scaling on one 8-GPU box: bench.py at N = 8, 4, 2 (torchrun) and psb_scan_box at 1/2/4/8 from one process
mkdir -p gpurun_out/r2k
for N in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 10 \
     > gpurun_out/r2k/bench_${N}gpu.json 2> gpurun_out/r2k/bench_${N}gpu.err
  echo "bench N=$N exit $?"
  python -c "
import json; d=json.load(open('gpurun_out/r2k/bench_${N}gpu.json')); print('N', d['n_gpus'], 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'ms/step', round(d['ms_per_step'],3), (d['config'].get('verified') or '')[:60])"
done
timeout 300 python - <<'PY' > gpurun_out/r2k/scan_box_probe.txt 2>&1
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
    for _ in range(4): a.scan_box((pc.numpy(), po.numpy()), ng)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); r = a.scan_box((pc.numpy(), po.numpy()), ng); ts.append(time.perf_counter() - t0)
    dt = float(np.median(ts))
    print(f"scan_box n_gpus {ng}: median {dt*1e3:.3f} ms per call (min {min(ts)*1e3:.3f}), {400*float(off[-1])/dt/1e9:.0f} GCUPS, slowest device's kernels {ps.kernel_ms():.3f} ms", flush=True)
PY
cat gpurun_out/r2k/scan_box_probe.txt | grep scan_box
