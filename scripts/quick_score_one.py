"""One query class through ss_score_batch (profiling helper)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
D, V = 10_000_000, 1_000_000
title = synth.index_table(V, D, 0); body = synth.index_table(V, D, 1)
e = capi.Engine(timing=True)
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf); e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf)
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
rng = np.random.default_rng(7); pr = (rng.random((D, 16)) + 0.5) / D; e.set_pagerank(pr); probs = np.full(16, 1 / 16)
lo, hi, Q = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
qs = [[int(rng.integers(lo, hi))] for _ in range(Q)]
kw_ptr = np.arange(Q + 1, dtype=np.uint64); kw = np.array([t for x in qs for t in x], np.uint32)
for _ in range(2):
    e.score_batch(kw_ptr, kw, topic_probs=probs, k=10); s = e.score_stats()
print(f"score {s.score_kernel_ms:.2f} ms {s.score_kernel_ms*1e3/Q:.1f} us/query")
