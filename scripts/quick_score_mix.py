"""Scoring cost by query class (dev helper): hot single-term, hot+rare, rare-only."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
D, V = 10_000_000, 1_000_000
title = synth.index_table(V, D, 0); body = synth.index_table(V, D, 1)
e = capi.Engine(timing=True)
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf); e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf)
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
rng = np.random.default_rng(7); pr = (rng.random((D, 16)) + 0.5) / D; e.set_pagerank(pr); probs = np.full(16, 1 / 16)
def run(name, qs):
    kw_ptr = np.zeros(len(qs) + 1, np.uint64); kw_ptr[1:] = np.cumsum([len(x) for x in qs]); kw = np.array([t for x in qs for t in x], np.uint32)
    for _ in range(2):
        e.score_batch(kw_ptr, kw, topic_probs=probs, k=10); s = e.score_stats()
    print(f"{name:28s} Q={len(qs)} score {s.score_kernel_ms:8.2f} ms  {s.score_kernel_ms*1e3/len(qs):8.1f} us/query  postings/q {s.postings_scanned/len(qs):.0f} matched/q {s.docs_matched/len(qs):.0f}", flush=True)
Q = 500
run("hot single [r<14]", [[int(rng.integers(0, 14))] for _ in range(Q)])
run("hot + rare", [[int(rng.integers(0, 14)), int(rng.integers(1000, V))] for _ in range(Q)])
run("two hot", [[int(rng.integers(0, 14)), int(rng.integers(0, 14))] for _ in range(Q)])
run("mid single [100..1000]", [[int(rng.integers(100, 1000))] for _ in range(Q)])
run("mid x3", [[int(rng.integers(100, 1000)) for _ in range(3)] for _ in range(Q)])
run("rare x3 [>10000]", [[int(rng.integers(10000, V)) for _ in range(3)] for _ in range(Q)])
run("rare x1", [[int(rng.integers(10000, V))] for _ in range(Q)])
