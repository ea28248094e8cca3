"""Quick scoring throughput probe (dev helper)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
D = int(float(sys.argv[1])) if len(sys.argv) > 1 else 10_000_000
V = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
Q = int(float(sys.argv[3])) if len(sys.argv) > 3 else 2000
t0 = time.time()
title = synth.index_table(V, D, 0); body = synth.index_table(V, D, 1); q = synth.queries(Q, V)
print("gen", time.time() - t0, "P", title.n_postings, body.n_postings, flush=True)
e = capi.Engine(timing=True)
t0 = time.time()
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf); e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf)
print("load", time.time() - t0, flush=True); t0 = time.time()
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
print("weights", time.time() - t0, flush=True)
rng = np.random.default_rng(7); pr = (rng.random((D, 16)) + 0.5) / D; e.set_pagerank(pr)
probs = np.full(16, 1 / 16)
for rep in range(3):
    t0 = time.time(); out = e.score_batch(q.kw_ptr, q.kw_terms, topic_probs=probs, k=10); dt = time.time() - t0
    s = e.score_stats()
    print(f"rep{rep} wall {dt*1e3:.1f} ms kernel {s.kernel_ms:.1f} score {s.score_kernel_ms:.1f} ms  q/s {Q/(s.kernel_ms*1e-3):.0f} postings {s.postings_scanned} matched {s.docs_matched} "
          f"algGB/s {s.algorithmic_bytes/(s.score_kernel_ms*1e-3)/1e9:.0f}  Gpost/s {s.postings_scanned/(s.score_kernel_ms*1e-3)/1e9:.1f}", flush=True)
