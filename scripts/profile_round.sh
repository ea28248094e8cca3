#!/bin/bash
# Round profile capture (run under gpurun): launch list of the bench command, then one
# `ncu --set full` capture of the dominant kernels.  Outputs land in gpurun_out/.
set -x
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu"
$BENCH > gpurun_out/bench_plain.json 2> gpurun_out/bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_bench_launches.csv \
    $BENCH > gpurun_out/bench_ncu.json 2> gpurun_out/bench_ncu.err
python bench.py --steps 1 --warmup 3 --no-cpu --workload pagerank > gpurun_out/pr_plain.json 2>gpurun_out/pr_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 3 -o gpurun_out/r01_pagerank_sweep \
    python bench.py --steps 1 --warmup 3 --no-cpu --workload pagerank > gpurun_out/pr_ncu.json 2>gpurun_out/pr_ncu.err
python bench.py --steps 1 --warmup 3 --no-cpu --workload scoring > gpurun_out/sc_plain.json 2>gpurun_out/sc_plain.err && \
ncu --set full --clock-control none --import-source on -k regex:k_score -s 3 -c 1 -o gpurun_out/r01_score \
    python bench.py --steps 1 --warmup 3 --no-cpu --workload scoring > gpurun_out/sc_ncu.json 2>gpurun_out/sc_ncu.err
tail -2 gpurun_out/*.err | tail -30
