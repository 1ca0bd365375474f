set -e
CMD="python bench.py --no-cpu --db 200000 --steps 2 --warmup 3 --e2e-steps 1"
$CMD > gpurun_out/r1e_small.json 2> gpurun_out/r1e_small.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1e_launches.csv $CMD > gpurun_out/ncu1.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:sw16 -s 3 -c 1 -f -o gpurun_out/r1e_sw16 $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
