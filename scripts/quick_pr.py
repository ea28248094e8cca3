"""Quick C2-scale PageRank timing (dev helper, not the bench)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
m = int(float(sys.argv[2])) if len(sys.argv) > 2 else 150_000_000
t0 = time.time(); g = synth.graph(n, m); print("gen", time.time() - t0, "E", g.n_edges, flush=True)
e = capi.Engine(timing=True)
t0 = time.time(); e.graph_load_csr(g.row_ptr, g.col_idx); print("load s", time.time() - t0, flush=True)
npg = synth.topics(16)
for rep in range(3):
    t0 = time.time()
    _, iters, st = e.pagerank(0.75, 1e-9, npg, want_rank=False)
    dt = time.time() - t0
    s = e.pagerank_stats()
    B = 4 * g.n_edges + 8 * (n + 1) + 8 * n + 16 * 16 * n
    per = s.sweep_ms_total / s.sweeps
    print(f"rep{rep} wall {dt*1e3:.2f} ms sweeps {s.sweeps} iters {iters.tolist()} sweep_ms {per:.3f} gather_ms {s.gather_ms_total/s.sweeps:.3f} "
          f"GTEPS {g.n_edges*16/per/1e6:.1f} algGB/s {B/per/1e6:.0f} frac {B/per/1e6/6551:.3f}", flush=True)
