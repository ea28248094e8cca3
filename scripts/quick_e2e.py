"""PageRank e2e breakdown with pinned host buffers (dev helper)."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
n, m = 10_000_000, 150_000_000
g = synth.graph(n, m)
def pin(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.uint8).reshape(-1)).pin_memory()
hr, hc = pin(g.row_ptr), pin(g.col_idx)
pr, pc = hr.numpy().view(np.uint64), hc.numpy().view(np.uint32)
out = torch.empty(n * 16, dtype=torch.float64).pin_memory().numpy().reshape(n, 16)
e = capi.Engine(timing=True)
npg = synth.topics(16)
for rep in range(4):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e.graph_load_csr(pr, pc); t1 = time.perf_counter()
    e.pagerank(0.75, 1e-9, npg, want_rank=False); t2 = time.perf_counter()
    e.pagerank_fetch(0, n, out=out); t3 = time.perf_counter()
    e.pagerank(0.75, 1e-9, npg, out=out); t4 = time.perf_counter()
    print(f"rep{rep} load {1e3*(t1-t0):.1f} ms (engine {e.pagerank_stats().load_ms:.1f}) pagerank {1e3*(t2-t1):.1f} fetch {1e3*(t3-t2):.1f}  pagerank+out {1e3*(t4-t3):.1f}", flush=True)
