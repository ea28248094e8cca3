"""Small end-to-end run of both hot paths for compute-sanitizer (memcheck)."""
import sys
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
e = capi.Engine()
g = synth.graph(3000, 60000, seed=1)
e.graph_load_csr(g.row_ptr, g.col_idx)
for T in (1, 3, 8, 16):
    r, it, st = e.pagerank(0.75, 1e-9, synth.topics(T))
    assert np.isfinite(r).all()
V, D = 800, 3000
for tid in (0, 1):
    t = synth.index_table(V, D, tid, with_positions=True)
    e.index_load(tid, D, t.term_ptr, t.doc_ids, t.norm_tf, t.pos_ptr, t.pos)
    e.term_weights(tid, float(D), t.n_postings, D)
e.use_pagerank()
q = synth.queries(300, V, phrase_fraction=0.3)
for k in (1, 10, 128):
    e.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=np.full(16, 1 / 16), k=k)
    e.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=np.random.default_rng(0).random((300, 16)), k=k)
e.score_batch(q.kw_ptr, q.kw_terms, k=50)
# the impact-vector path in all its modes (whole slab, staged sparse tokens, sub-ranges, phrase tokens) and the
# async short-row kernel: force them on the small inputs through the knobs
import os
for env in ({"SS_SCORE_SORT_MAX": "0"}, {"SS_SCORE_SORT_MAX": "0", "SS_SCORE_DENSE_MAX": "3"},
            {"SS_SCORE_SORT_MAX": "0", "SS_SCORE_DENSE_FRAC": "100000"}, {"SS_SCORE_OWNER": "0", "SS_SCORE_QTHR": "0"}):
    os.environ.update(env)
    e.use_pagerank()  # drops the cached impact vectors
    for k in (10, 128):
        e.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=np.full(16, 1 / 16), k=k)
        e.score_batch(q.kw_ptr, q.kw_terms, k=k)
    for key in env:
        del os.environ[key]
os.environ["SS_PR_SHORT"] = "async"
e.graph_load_csr(g.row_ptr, g.col_idx)
for T in (1, 3, 8, 16):
    r, it, st = e.pagerank(0.75, 1e-9, synth.topics(T))
    assert np.isfinite(r).all()
del os.environ["SS_PR_SHORT"]
e.close()
print("sanitize run ok")
