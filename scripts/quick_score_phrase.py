"""Keyword-only vs 20 %-phrase query batches on a smaller index with positions (dev helper)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
D, V, Q = 2_000_000, 200_000, 20000
title = synth.index_table(V, D, 0, with_positions=True); body = synth.index_table(V, D, 1, with_positions=True)
e = capi.Engine(timing=True)
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf, title.pos_ptr, title.pos)
e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf, body.pos_ptr, body.pos)
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
rng = np.random.default_rng(7); pr = (rng.random((D, 16)) + 0.5) / D; e.set_pagerank(pr); probs = np.full(16, 1 / 16)
for frac in (0.0, 0.2, 1.0):
    q = synth.queries(Q, V, phrase_fraction=frac, seed=44)
    for _ in range(2):
        e.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10); s = e.score_stats()
    print(f"phrase_fraction {frac}: {s.kernel_ms:.1f} ms  {Q / (s.kernel_ms * 1e-3):.0f} queries/s  postings/q {s.postings_scanned / Q:.0f}", flush=True)
