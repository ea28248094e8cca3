#!/bin/bash
# Round-2 ncu evidence (one gpurun call, one GPU): launch list of the bench command, then full captures of the
# PageRank sweep kernels and of k_score.  Each ncu run follows a plain run of the same command that exited 0.
set -u
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-parity --queries 20000 --mixed-queries 20000"
$CMD > gpurun_out/r02_plain.json 2> gpurun_out/r02_plain.err || { echo "plain run failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
    --log-file gpurun_out/r02_bench_launches.csv $CMD > gpurun_out/r02_ncu_launches.log 2>&1
PR="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --workload pagerank"
$PR > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 18 -c 4 \
    -o gpurun_out/r02_sweep $PR > gpurun_out/r02_ncu_sweep.log 2>&1
SC="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --workload scoring --queries 20000 --phrase-fraction 0"
$SC > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_score -s 3 -c 1 \
    -o gpurun_out/r02_score $SC > gpurun_out/r02_ncu_score.log 2>&1
ls -la gpurun_out/*.ncu-rep
