#!/bin/bash
# Round profile capture (run under gpurun): the bench without a profiler, then its launch list with
# per-launch DRAM bytes (one ncu pass per kernel), then `ncu --set full` captures of the dominant
# kernels on SHORT runs (a full 100K-query k_score replayed ~40 times costs ~15 GPU-minutes).
set -x
TAG=${1:-r01}
BENCH="python bench.py --steps 2 --warmup 3 --no-cpu"
$BENCH > gpurun_out/${TAG}_bench_plain.json 2> gpurun_out/${TAG}_bench_plain.err && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_bench_launches.csv $BENCH > gpurun_out/${TAG}_bench_ncu.json 2> gpurun_out/${TAG}_bench_ncu.err
python scripts/quick_pr.py > gpurun_out/${TAG}_pr_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_sweep -s 9 -c 3 -o gpurun_out/${TAG}_pagerank_sweep \
    python scripts/quick_pr.py > gpurun_out/${TAG}_pr_ncu.log 2>&1
python scripts/quick_score_class.py mix 3000 > gpurun_out/${TAG}_sc_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_score -s 2 -c 1 -o gpurun_out/${TAG}_score_mix3000 \
    python scripts/quick_score_class.py mix 3000 > gpurun_out/${TAG}_sc_ncu.log 2>&1
tail -n 2 gpurun_out/${TAG}_*.err gpurun_out/${TAG}_*.log | tail -40
