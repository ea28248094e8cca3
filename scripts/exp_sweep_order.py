"""Sweep-time experiments at configs[1] (harness-level graph transformations, engine untouched):
  base      the generated graph
  relabel   node ids reassigned by descending out-degree (hot sources contiguous)
  nohub     the in-edges of the H largest-in-degree rows removed (what a hub tile would save the pull kernels)
Prints avg sweep / gather ms per variant."""
import sys, time, json
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth

N, E, H = 10_000_000, 150_000_000, int(sys.argv[1]) if len(sys.argv) > 1 else 1280
g = synth.graph(N, E, seed=42)
npg = synth.topics(16)
eng = capi.Engine(device=0, timing=True)
res = {}

def run(tag, row_ptr, col):
    eng.graph_load_csr(row_ptr, col)
    sw = ga = sh = 0.0
    n = 0
    for i in range(5):
        eng.pagerank(0.75, 1e-9, npg, want_rank=False)
        s = eng.pagerank_stats()
        if i >= 2:
            sw += s.sweep_ms_total; ga += s.gather_ms_total; sh += s.short_ms_total; n += s.sweeps
    res[tag] = {"sweep_ms": sw / n, "gather_ms": ga / n, "short_ms": sh / n, "long_ms": (ga - sh) / n, "edges": int(row_ptr[-1]), "sweeps": n}
    print(tag, res[tag], flush=True)

rp, ci = g.row_ptr, g.col_idx
run("base", rp, ci)
outd = np.diff(rp.astype(np.int64))
src = np.repeat(np.arange(N, dtype=np.uint32), outd)
ind = np.bincount(ci, minlength=N)

# drop hub in-edges
hubs = np.argsort(-ind, kind="stable")[:H]
is_hub = np.zeros(N, dtype=bool); is_hub[hubs] = True
keep = ~is_hub[ci]
ci2 = ci[keep]
cnt = np.bincount(src[keep], minlength=N)
rp2 = np.zeros(N + 1, dtype=np.uint64); rp2[1:] = np.cumsum(cnt)
run(f"nohub{H}", rp2, ci2)
del ci2, rp2, keep

# relabel by out-degree descending
order = np.argsort(-outd, kind="stable")          # new id -> old id
newid = np.empty(N, dtype=np.uint32); newid[order] = np.arange(N, dtype=np.uint32)
cnt = outd[order]
rp3 = np.zeros(N + 1, dtype=np.uint64); rp3[1:] = np.cumsum(cnt)
# edges of new row i = edges of old row order[i], children relabelled
starts = rp[:-1].astype(np.int64)[order]
idx = np.repeat(starts - rp3[:-1].astype(np.int64), cnt) + np.arange(int(rp3[-1]), dtype=np.int64)
ci3 = newid[ci[idx]]
del idx
run("relabel_outdeg", rp3, ci3)
# relabel by in-degree descending (hub rows contiguous, sources random)
order = np.argsort(-ind, kind="stable")
newid[order] = np.arange(N, dtype=np.uint32)
cnt = outd[order]
rp4 = np.zeros(N + 1, dtype=np.uint64); rp4[1:] = np.cumsum(cnt)
starts = rp[:-1].astype(np.int64)[order]
idx = np.repeat(starts - rp4[:-1].astype(np.int64), cnt) + np.arange(int(rp4[-1]), dtype=np.int64)
ci4 = newid[ci[idx]]
run("relabel_indeg", rp4, ci4)
json.dump(res, open("gpurun_out/exp_sweep_order.json", "w"), indent=1)
