#!/bin/bash
# round-2 late check: single long pair with use_trace() on the wavefront path, lazy flag-byte table; latency of the 20 kb case
mkdir -p gpurun_out/r4l
timeout 120 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "single_long_pair_trace_api or single_pair_api" > gpurun_out/r4l/pytest.txt 2>&1
echo "pytest rc $?" >> gpurun_out/r4l/pytest.txt
tail -n 12 gpurun_out/r4l/pytest.txt
timeout 60 python - > gpurun_out/r4l/single_long.txt 2>&1 <<'PY'
import sys, time
sys.path.insert(0, "tests")
import psb_data, parasail_rs_b200 as ps
L = 20000
r = psb_data.random_seq(5001, 0, L, protein=False)
q = psb_data.mutate(r, 5001, 1, 0.10, 0.01, protein=False)[:L]
a = ps.Aligner.new().local().matrix(ps.Matrix.create(b"ACGT", 2, -3)).gap_open(5).gap_extend(2).use_trace().build()
a.align(q[:3000], r[:3000])
t0 = time.perf_counter(); res = a.align(q, r); t1 = time.perf_counter()
cg = res.get_cigar(q, r); t2 = time.perf_counter()
print("align(20 kb x 20 kb, use_trace) %.1f ms, score %d, get_cigar %.1f ms, %d chars" % ((t1 - t0) * 1e3, res.get_score(), (t2 - t1) * 1e3, len(cg)))
t3 = time.perf_counter(); tt = res.get_trace_table(); t4 = time.perf_counter()
print("get_trace_table (lazy: the pair again on the flag-byte kernel + row-major copy) %.1f ms, shape %s" % ((t4 - t3) * 1e3, tt.shape))
PY
cat gpurun_out/r4l/single_long.txt
