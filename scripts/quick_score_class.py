"""One query class through ss_score_batch (dev helper for ncu captures): hot | mid3 | mix."""
import sys
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
cls = sys.argv[1] if len(sys.argv) > 1 else "hot"
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 300
K = int(sys.argv[3]) if len(sys.argv) > 3 else 10
D, V = 10_000_000, 1_000_000
title = synth.index_table(V, D, 0); body = synth.index_table(V, D, 1)
e = capi.Engine(timing=True)
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf); e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf)
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
rng = np.random.default_rng(7); pr = (rng.random((D, 16)) + 0.5) / D; e.set_pagerank(pr); probs = np.full(16, 1 / 16)
if cls == "hot":
    qs = [[int(rng.integers(0, 14))] for _ in range(Q)]
elif cls == "mid3":
    qs = [[int(rng.integers(100, 1000)) for _ in range(3)] for _ in range(Q)]
elif cls == "warm":
    qs = [[int(rng.integers(14, 132))] for _ in range(Q)]
else:
    q = synth.queries(Q, V); qs = None
if qs is not None:
    kw_ptr = np.zeros(len(qs) + 1, np.uint64); kw_ptr[1:] = np.cumsum([len(x) for x in qs]); kw = np.array([t for x in qs for t in x], np.uint32)
else:
    kw_ptr, kw = q.kw_ptr, q.kw_terms
for _ in range(3):
    e.score_batch(kw_ptr, kw, topic_probs=probs, k=K); s = e.score_stats()
print(f"{cls} Q={Q} score {s.score_kernel_ms:.2f} ms {s.score_kernel_ms*1e3/Q:.1f} us/query postings/q {s.postings_scanned/Q:.0f} ps/posting {s.score_kernel_ms*1e9/max(1,s.postings_scanned):.1f}", flush=True)
