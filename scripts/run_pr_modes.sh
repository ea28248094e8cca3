#!/bin/bash
# PageRank exchange modes back to back on N GPUs (dev tool).  usage: run_pr_modes.sh N "mode1 mode2" [bench args]
N=${1:-4}; MODES=${2:-"default nccl fused"}; shift; shift
for mode in $MODES; do
  if [ $mode = default ]; then unset SS_PR_EXCHANGE; else export SS_PR_EXCHANGE=$mode; fi
  python -m torch.distributed.run --nnodes=1 --nproc-per-node=$N --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus $N --steps 3 --warmup 2 --workload pagerank --no-parity "$@" > gpurun_out/pr_mode_${N}_${mode}.json 2> gpurun_out/pr_mode_${N}_${mode}.err
  python - <<PY
import json
b=json.loads(open('gpurun_out/pr_mode_${N}_${mode}.json').read().strip().splitlines()[-1])
r=b['roofline']
print('$mode', 'N=$N', '$*', 'GTEPS', round(b['value'],1), 'ms/step', round(b['ms_per_step'],2), 'sweep', round(r['avg_sweep_ms'],2), 'exposed', round(r['exchange_exposed_ms_per_sweep'],2), 'busy', round(r['exchange_busy_ms_per_sweep'],2), 'e2e', round(b['e2e']['value'],1), 'e2e ms', round(b['e2e']['ms_per_step'],1), 'load ms', round(b['e2e'].get('load_ms',0),1))
PY
done
