"""Scoring throughput for several k on the configs[2] index (dev helper)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
D, V, Q = 10_000_000, 1_000_000, 20000
title = synth.index_table(V, D, 0); body = synth.index_table(V, D, 1); q = synth.queries(Q, V)
e = capi.Engine(timing=True)
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf); e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf)
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
rng = np.random.default_rng(7); pr = (rng.random((D, 16)) + 0.5) / D; e.set_pagerank(pr); probs = np.full(16, 1 / 16)
for k in (10, 50, 128):
    for _ in range(2):
        e.score_batch(q.kw_ptr, q.kw_terms, topic_probs=probs, k=k); s = e.score_stats()
    print(f"k={k}: kernel {s.kernel_ms:.1f} ms  {Q / (s.kernel_ms * 1e-3):.0f} queries/s", flush=True)
