#!/bin/bash
# ncu --set full of k_score on the first 20,000 queries of the configs[2] batch (one GPU, one ncu use)
SC="python bench.py --steps 1 --warmup 3 --no-cpu --no-parity --workload scoring --queries 20000 --phrase-fraction 0"
$SC > gpurun_out/r02b_score_plain.json 2> gpurun_out/r02b_score_plain.err && ncu --set full --clock-control none --import-source on -k regex:k_score -s 3 -c 1 \
    -o gpurun_out/r02b_score $SC > gpurun_out/r02b_ncu_score.log 2>&1
ls -la gpurun_out/r02b_score.ncu-rep
