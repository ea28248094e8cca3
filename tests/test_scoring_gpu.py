"""HP-2 parity: TF-IDF weighting, doc norms and the batched score/blend/top-k
through the C ABI against the oracle (ranking/term_weighting.go,
retrieval/main_retrieve.go, get_metadata.go, phrase.go).  Bar from
BASELINE.json north_star: identical top-k doc ids and order (ties by doc id),
scores within 1e-6 relative.  Weights and norms are integer-like work in fixed
order and are required bit-exact."""
import itertools
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import loader as O
from spaghettisearch_b200 import capi, synth
from tests.fixtures import queries_csr, tiny_index

pytestmark = pytest.mark.gpu
KATS = json.loads((Path(__file__).parent / "golden" / "kats.json").read_text())
REL_TOL = 1e-6


def assert_same_results(got, ref, rel=REL_TOL):
    gd, gf, gp, gc = got
    rd, rf, rp, rc = ref
    assert np.array_equal(gc, rc), (gc[:10], rc[:10])
    assert np.array_equal(gd, rd), np.argwhere(gd != rd)[:5]
    for g, r in ((gf, rf), (gp, rp)):
        both_nan = np.isnan(g) & np.isnan(r)
        ok = both_nan | (np.abs(g - r) <= rel * np.abs(r)) | (g == r)
        assert ok.all(), (g[~ok][:5], r[~ok][:5])


def load_weighted(engine, title, body, n_docs, total_docs, df_title=None, df_body=None):
    """Loads both tables, runs ss_term_weights, returns oracle tables + norms computed by the oracle."""
    out = {}
    engine.index_clear()
    for name, tb, tid, df in (("title", title, capi.SS_TITLE, df_title), ("body", body, capi.SS_BODY, df_body)):
        term_ptr, doc_ids, tf, pos_ptr, pos = tb
        engine.index_load(tid, n_docs, term_ptr, doc_ids, tf, pos_ptr, pos)
        w, mag = engine.term_weights(tid, total_docs, len(doc_ids), n_docs, df_global=df)
        if df is None:
            ow, omag = O.term_weights(term_ptr, doc_ids, tf, n_docs, total_docs)
            assert np.array_equal(w.view(np.uint32), ow.view(np.uint32)), name  # bit exact
            assert np.array_equal(mag.view(np.uint64), omag.view(np.uint64)), name
        out[name] = (O.Table(term_ptr, doc_ids, w, pos_ptr, pos), mag)
    return out["title"][0], out["body"][0], out["title"][1], out["body"][1]


def test_kat_sc1(engine):
    k = KATS["KAT-SC-1"]

    def tab(t):
        return (np.array(t["term_ptr"], np.uint64), np.array(t["doc_ids"], np.uint32),
                np.array(t["norm_tf"], np.float32), None, None)

    engine.index_clear()
    for name, tid in (("title", capi.SS_TITLE), ("body", capi.SS_BODY)):
        t = tab(k[name])
        engine.index_load(tid, k["n_docs"], *t)
        w, mag = engine.term_weights(tid, k["total_docs"], len(t[1]), k["n_docs"])
        assert w.tolist() == k["w_" + name] and mag.tolist() == k["mag_" + name]
    engine.set_pagerank(None)
    d, f, p, c = engine.score_batch([0, len(k["query"])], k["query"], k=50)
    assert c[0] == 3
    assert d[0, :3].tolist() == [r["doc"] for r in k["result"]]
    assert f[0, :3].tolist() == [r["final"] for r in k["result"]]  # bit exact: same operation order
    assert (d[0, 3:] == 0xFFFFFFFF).all() and (p == 0).all()


def test_tiny_index_all_queries(engine):
    title, body, n_docs = tiny_index()
    ot, ob, tmag, bmag = load_weighted(engine, title, body, n_docs, 6.0)
    engine.set_pagerank(None)
    terms = [0, 1, 2, 3, 99]  # 99 is an unknown term
    kws, phs = [], []
    for n in (0, 1, 2, 3):
        for kw in itertools.product(terms, repeat=n):
            for ph in ([], [0], [2], [3], [0, 1], [1, 2], [0, 2], [2, 2], [0, 1, 2], [99, 0]):
                if n == 3 and len(ph) > 1:
                    continue
                kws.append(list(kw))
                phs.append(ph)
    kw_ptr, kw, ph_ptr, ph = queries_csr(kws, phs)
    for k in (1, 3, 50):
        got = engine.score_batch(kw_ptr, kw, ph_ptr, ph, k=k)
        ref = O.score_batch(ot, ob, n_docs, tmag, bmag, None, kw_ptr, kw, ph_ptr, ph, k=k)
        assert_same_results(got, ref, rel=0.0)
    # queries without any token match nothing
    assert got[3][0] == 0


def _synth_tables(V, D, with_positions=True, doc_lo=0, doc_hi=None):
    t = synth.index_table(V, D, 0, with_positions=with_positions, doc_lo=doc_lo, doc_hi=doc_hi)
    b = synth.index_table(V, D, 1, with_positions=with_positions, doc_lo=doc_lo, doc_hi=doc_hi)
    as_tuple = lambda x: (x.term_ptr, x.doc_ids, x.norm_tf, x.pos_ptr, x.pos)
    return t, b, as_tuple(t), as_tuple(b)


@pytest.fixture(scope="module")
def synth_index(engine):
    V, D = 20000, 60000
    t, b, tt, bt = _synth_tables(V, D)
    ot, ob, tmag, bmag = load_weighted(engine, tt, bt, D, float(D))
    rng = np.random.default_rng(5)
    pr = rng.uniform(0, 2e-5, (D, 16))
    return dict(V=V, D=D, ot=ot, ob=ob, tmag=tmag, bmag=bmag, pr=pr)


@pytest.mark.parametrize("k", [10, 50])
def test_synthetic_keyword_and_phrase(engine, synth_index, k):
    s = synth_index
    q = synth.queries(1500, s["V"], phrase_fraction=0.3, seed=44)
    engine.set_pagerank(None)
    got = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, k=k)
    ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], None, q.kw_ptr, q.kw_terms, q.ph_ptr,
                        q.ph_terms, k=k)
    assert_same_results(got, ref)
    assert (got[3] > 0).mean() > 0.9
    st = engine.score_stats()
    assert st.postings_scanned > 0 and st.docs_matched > 0 and st.launches >= 2


def test_synthetic_blend_shared_and_per_query(engine, synth_index):
    s = synth_index
    q = synth.queries(800, s["V"], phrase_fraction=0.2, seed=45)
    engine.set_pagerank(s["pr"])
    probs = np.full(16, 1.0 / 16)
    got = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10)
    ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], s["pr"], q.kw_ptr, q.kw_terms, q.ph_ptr,
                        q.ph_terms, topic_probs=probs, k=10)
    assert_same_results(got, ref)
    assert (got[2][got[0] != 0xFFFFFFFF] > 0).all()
    rng = np.random.default_rng(9)
    probs_q = rng.dirichlet(np.ones(16), size=800)
    got = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs_q, k=10)
    ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], s["pr"], q.kw_ptr, q.kw_terms, q.ph_ptr,
                        q.ph_terms, topic_probs=probs_q, k=10)
    assert_same_results(got, ref)
    engine.set_pagerank(None)


def test_blend_at_order_one_magnitude(engine, synth_index):
    """with ranks around 1/D the blend term never changes an order.  Here the PageRank rows are O(1)
    (and signed), so 0.33 * sqd * 100 is as large as the text scores and decides most of the top 10: the blend
    bounds of every path (slab maxima, fp16 blend vector, block maxima, global bound) are exercised for real."""
    s = synth_index
    rng = np.random.default_rng(31)
    pr = rng.uniform(-0.5, 1.5, (s["D"], 16))
    pr[rng.random(s["D"]) < 0.01] *= 40.0          # a few outliers the bounds must not lose
    q = synth.queries(600, s["V"], phrase_fraction=0.2, seed=47)
    engine.set_pagerank(pr)
    try:
        for probs in (np.full(16, 1.0 / 16), rng.dirichlet(np.ones(16), size=600), -np.full(16, 1.0 / 16)):
            got = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10)
            ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], pr, q.kw_ptr, q.kw_terms, q.ph_ptr,
                                q.ph_terms, topic_probs=probs, k=10)
            assert_same_results(got, ref)
        # the blend really reorders: without it the top doc of most queries is a different one
        plain = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, k=10)
        both = (got[3] > 0) & (plain[3] > 0)
        assert (got[0][both, 0] != plain[0][both, 0]).mean() > 0.5
    finally:
        engine.set_pagerank(None)


def test_hot_terms_many_ties(engine, synth_index):
    # single hot term: half the docs match, normTF takes few distinct values -> masses of exact ties,
    # so the order is decided by the doc-id rule
    s = synth_index
    kw_ptr, kw, _, _ = queries_csr([[0], [1], [0, 0], [0, 1, 2, 3, 4]])
    engine.set_pagerank(None)
    for k in (10, 128):
        got = engine.score_batch(kw_ptr, kw, k=k)
        ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], None, kw_ptr, kw, k=k)
        assert_same_results(got, ref)


@pytest.mark.parametrize("env", [
    {},                                                   # default: impact vectors + owner path + running bound
    {"SS_SCORE_DENSE": "0"},                              # accumulator paths only
    {"SS_SCORE_SORT_MAX": "0"},                           # every slab with a dense token takes the impact-vector path
    {"SS_SCORE_SORT_MAX": "0", "SS_SCORE_DENSE_MAX": "3"},   # 3 dense terms: mixed dense + scattered sparse tokens
    {"SS_SCORE_SORT_MAX": "0", "SS_SCORE_DENSE_FRAC": "100000"},  # every term with a posting is "dense" (up to 224)
    {"SS_SCORE_QTHR": "0", "SS_SCORE_OWNER": "0"},        # first-version sparse path, no cross-slab bound
    {"SS_SCORE_PHRASE_DENSE": "0"},                       # phrase queries on the accumulator / sort paths only
])
def test_scoring_paths_agree_with_oracle(engine, synth_index, monkeypatch, env):
    # The execution paths of ss_score_batch differ only in HOW they find the docs worth an exact
    # evaluation; every one must return the oracle's top k (ids, order, scores).
    s = synth_index
    q = synth.queries(1200, s["V"], seed=48)
    hot = queries_csr([[0], [1], [0, 1], [0, 5000], [3, 0, 3], [2, 7, 11, 300, 19000], [250], [250, 251, 252]])
    for key, val in env.items():
        monkeypatch.setenv(key, val)
    probs = np.full(16, 1.0 / 16)
    for pr, tp in ((None, None), (s["pr"], probs)):
        engine.set_pagerank(pr)  # also drops the cached impact vectors, so the env settings take effect
        for kw_ptr, kw in ((q.kw_ptr, q.kw_terms), (hot[0], hot[1])):
            for k in (10, 50):
                got = engine.score_batch(kw_ptr, kw, topic_probs=tp, k=k)
                ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], pr, kw_ptr, kw, topic_probs=tp, k=k)
                assert_same_results(got, ref)
    # phrase queries: with a dense keyword token they take the impact-vector path too (the phrase tokens
    # enter the bound, the positions are checked for the survivors); hot terms 0..3 make that common
    qp = synth.queries(600, s["V"], phrase_fraction=0.5, seed=49)
    hp = queries_csr([[0], [1, 0], [0, 2], [5000, 1], [0], [2, 2]], [[1, 2], [2, 3], [0, 1, 2], [0, 1], [7, 300], [2]])
    engine.set_pagerank(None)
    for kw_ptr, kw, ph_ptr, ph in ((qp.kw_ptr, qp.kw_terms, qp.ph_ptr, qp.ph_terms), hp):
        got = engine.score_batch(kw_ptr, kw, ph_ptr, ph, k=10)
        ref = O.score_batch(s["ot"], s["ob"], s["D"], s["tmag"], s["bmag"], None, kw_ptr, kw, ph_ptr, ph, k=10)
        assert_same_results(got, ref)
    for key in env:
        monkeypatch.delenv(key)
    engine.set_pagerank(None)


def test_term_weights_not_idempotent(engine):
    # term_weighting.go:42-47 multiplies the stored weight in place: a second call weighs again
    V, D = 300, 800
    b = synth.index_table(V, D, 1)
    engine.index_clear()
    engine.index_load(capi.SS_BODY, D, b.term_ptr, b.doc_ids, b.norm_tf)
    w1, _ = engine.term_weights(capi.SS_BODY, float(D), b.n_postings, D)
    w2, mag2 = engine.term_weights(capi.SS_BODY, float(D), b.n_postings, D)
    ow1, _ = O.term_weights(b.term_ptr, b.doc_ids, b.norm_tf, D, float(D))
    ow2, omag2 = O.term_weights(b.term_ptr, b.doc_ids, ow1, D, float(D))
    assert np.array_equal(w1, ow1) and np.array_equal(w2, ow2) and np.array_equal(mag2, omag2)


def test_idf_matches_go_log2_bitwise(engine):
    # one term per df value: idf = float32(Log2(N/df)) for every df in 1..N with a non power-of-two N
    N = 3001
    term_ptr = np.zeros(N + 1, np.uint64)
    term_ptr[1:] = np.cumsum(np.arange(1, N + 1))
    doc_ids = np.concatenate([np.arange(df, dtype=np.uint32) for df in range(1, N + 1)])
    tf = np.ones(len(doc_ids), np.float32)
    engine.index_clear()
    engine.index_load(capi.SS_BODY, N, term_ptr, doc_ids, tf)
    w, mag = engine.term_weights(capi.SS_BODY, float(N), len(doc_ids), N)
    ow, omag = O.term_weights(term_ptr, doc_ids, tf, N, float(N))
    assert np.array_equal(w.view(np.uint32), ow.view(np.uint32))
    assert np.array_equal(mag.view(np.uint64), omag.view(np.uint64))
    idf = w[term_ptr[:-1].astype(np.int64)]
    expect = np.array([np.float32(O.go_log2(N / df)) for df in range(1, N + 1)], np.float32)
    assert np.array_equal(idf, expect)
    assert idf[-1] == 0.0  # df == N


def test_doc_sharded_merge(engine):
    # every shard scores the whole batch against its docs with GLOBAL df,
    # local top-k lists are merged with the same comparator
    V, D, k = 5000, 12000, 10
    t, b, tt, bt = _synth_tables(V, D, with_positions=True)
    q = synth.queries(400, V, phrase_fraction=0.25, seed=46)
    ow_t, mag_t = O.term_weights(t.term_ptr, t.doc_ids, t.norm_tf, D, float(D))
    ow_b, mag_b = O.term_weights(b.term_ptr, b.doc_ids, b.norm_tf, D, float(D))
    ref = O.score_batch(O.Table(t.term_ptr, t.doc_ids, ow_t, t.pos_ptr, t.pos),
                        O.Table(b.term_ptr, b.doc_ids, ow_b, b.pos_ptr, b.pos), D, mag_t, mag_b, None, q.kw_ptr,
                        q.kw_terms, q.ph_ptr, q.ph_terms, k=k)
    bounds = [0, 3000, 7000, D]
    parts = []
    for lo, hi in zip(bounds[:-1], bounds[1:]):
        st, sb, stt, sbt = _synth_tables(V, D, doc_lo=lo, doc_hi=hi)
        shard = capi.Engine(device=0)
        try:
            shard.index_load(capi.SS_TITLE, D, *stt)
            shard.index_load(capi.SS_BODY, D, *sbt)
            shard.term_weights(capi.SS_TITLE, float(D), st.n_postings, D, df_global=st.df_global, want=False)
            shard.term_weights(capi.SS_BODY, float(D), sb.n_postings, D, df_global=sb.df_global, want=False)
            parts.append(shard.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, k=k))
        finally:
            shard.close()
    merged = engine.merge_topk(np.stack([p[0] for p in parts]), np.stack([p[1] for p in parts]),
                               np.stack([p[2] for p in parts]), np.stack([p[3] for p in parts]))
    assert_same_results(merged, ref)


def test_pagerank_feeds_the_blend(engine):
    # HP-1 result used in place (ss_use_pagerank): node id == doc id
    N = 4000
    g = synth.graph(N, 50000, seed=3)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    rank, _, _ = engine.pagerank(0.75, 1e-9, synth.topics(16))
    t, b, tt, bt = _synth_tables(1000, N, with_positions=False)
    ot, ob, tmag, bmag = load_weighted(engine, tt, bt, N, float(N))
    engine.use_pagerank()
    q = synth.queries(300, 1000, seed=47)
    probs = np.full(16, 1.0 / 16)
    got = engine.score_batch(q.kw_ptr, q.kw_terms, topic_probs=probs, k=10)
    ref = O.score_batch(ot, ob, N, tmag, bmag, rank, q.kw_ptr, q.kw_terms, topic_probs=probs, k=10)
    assert_same_results(got, ref)
    engine.set_pagerank(None)


def test_invalid_arguments(engine):
    title, body, n_docs = tiny_index()
    engine.index_clear()
    with pytest.raises(capi.SSError):  # doc id out of range
        engine.index_load(capi.SS_BODY, 3, *body)
    with pytest.raises(capi.SSError):  # docs not ascending within a term
        engine.index_load(capi.SS_BODY, 6, np.array([0, 2], np.uint64), np.array([3, 1], np.uint32),
                          np.array([1, 1], np.float32))
    engine.index_load(capi.SS_BODY, n_docs, *body)
    engine.index_load(capi.SS_TITLE, n_docs, *title)
    with pytest.raises(capi.SSError):  # norms missing
        engine.score_batch([0, 1], [0], k=5)
    engine.term_weights(capi.SS_BODY, 6.0, len(body[1]), n_docs)
    engine.term_weights(capi.SS_TITLE, 6.0, len(title[1]), n_docs)
    with pytest.raises(capi.SSError):
        engine.score_batch([0, 1], [0], k=0)
    with pytest.raises(capi.SSError):
        engine.score_batch([0, 1], [0], k=129)
    with pytest.raises(capi.SSError):  # blend without PageRank
        engine.score_batch([0, 1], [0], topic_probs=np.ones(4), k=5)


# ---- round 2: over-long queries, shard-local doc ids, in-engine shard merge --------------------
def _small_index(engine, V=300, D=3000, seed=43):
    tabs = {}
    engine.index_clear()
    engine.set_pagerank(None)
    for tid in (capi.SS_TITLE, capi.SS_BODY):
        t = synth.index_table(V, D, tid, postings_per_doc=30.0 if tid else 6.0, with_positions=True, seed=seed)
        engine.index_load(tid, D, t.term_ptr, t.doc_ids, t.norm_tf, t.pos_ptr, t.pos)
        w, mag = engine.term_weights(tid, float(D), t.n_postings, D)
        tabs[tid] = (O.Table(t.term_ptr, t.doc_ids, w, t.pos_ptr, t.pos), mag, t)
    return tabs


def test_over_long_queries_are_served_per_query(engine):
    """A 33..256-token phrase or > 64 keyword tokens used to fail the whole batch (ADVICE round 1,
    score.cu:1763); the reference evaluates them (phrase.go:111-118).  They take the wide kernel variant,
    everybody else's rows are unchanged."""
    V, D = 300, 3000
    tabs = _small_index(engine, V, D)
    rng = np.random.default_rng(5)
    body = tabs[capi.SS_BODY][2]
    # a phrase that really occurs: consecutive positions do not exist in the synthetic index, so build the
    # long phrase from one doc's terms anyway (it must evaluate, matching or not), plus repeated-token phrases
    kws = [[1, 2], list(rng.integers(0, V, 100)), [3], list(rng.integers(0, V, 65)), [5, 5, 7], list(range(0, 256))]
    phs = [[], [], list(rng.integers(0, V, 40)), [4, 4], [], list(rng.integers(0, 8, 256))]
    kw_ptr, kw, ph_ptr, ph = queries_csr(kws, phs)
    pr = np.random.default_rng(6).random((D, 4)) * 1e-3
    for probs in (None, np.full(4, 0.25)):
        engine.set_pagerank(None if probs is None else pr)
        got = engine.score_batch(kw_ptr, kw, ph_ptr, ph, topic_probs=probs, k=10)
        exp = O.score_batch(tabs[0][0], tabs[1][0], D, tabs[0][1], tabs[1][1], None if probs is None else pr, kw_ptr,
                            kw, ph_ptr, ph, topic_probs=probs, k=10)
        assert_same_results(got, exp)
        assert got[3][1] > 0 and got[3][5] > 0
    # beyond the wide limits the call still fails, naming the query
    kw_ptr, kw, ph_ptr, ph = queries_csr([[1], list(range(257))], [[], []])
    with pytest.raises(capi.SSError) as ei:
        engine.score_batch(kw_ptr, kw, ph_ptr, ph, k=10)
    assert "query 1" in str(ei.value)
    # a phrase of more than 256 tokens can never match (uint8 TermPos): evaluated as "no phrase hit"
    kw_ptr, kw, ph_ptr, ph = queries_csr([[1, 2]], [list(rng.integers(0, V, 300))])
    engine.set_pagerank(None)
    got = engine.score_batch(kw_ptr, kw, ph_ptr, ph, k=10)
    exp = O.score_batch(tabs[0][0], tabs[1][0], D, tabs[0][1], tabs[1][1], None, kw_ptr, kw, ph_ptr, ph, k=10)
    assert_same_results(got, exp)


def test_merge_topk_rejects_k_above_128(engine):
    docs = np.zeros((2, 1, 129), np.uint32)
    with pytest.raises(capi.SSError):
        engine.merge_topk(docs, np.zeros((2, 1, 129)), np.zeros((2, 1, 129)), np.zeros((2, 1), np.uint32))


def test_doc_base_and_local_ids(engine):
    """A shard loaded under local ids + ss_index_set_doc_base returns the same global result rows as the
    same shard loaded under global ids (round 1 layout)."""
    V, D, lo, hi = 400, 6000, 2000, 4500
    rng = np.random.default_rng(11)
    pr = rng.random((D, 8)) * 1e-4
    probs = rng.random(8)
    q = synth.queries(200, V, phrase_fraction=0.3, seed=9)
    res = []
    for local in (False, True):
        engine.index_clear()
        engine.index_set_doc_base(lo if local else 0)
        for tid in (capi.SS_TITLE, capi.SS_BODY):
            t = synth.index_table(V, D, tid, postings_per_doc=30.0 if tid else 6.0, doc_lo=lo, doc_hi=hi,
                                  with_positions=True)
            ids = t.doc_ids - np.uint32(lo) if local else t.doc_ids
            nd = hi - lo if local else D
            engine.index_load(tid, nd, t.term_ptr, ids, t.norm_tf, t.pos_ptr, t.pos)
            engine.term_weights(tid, float(D), t.n_postings, nd, df_global=t.df_global, want=False)
        engine.set_pagerank(pr[lo:hi] if local else pr)
        res.append(engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10))
        # without a communicator the sharded entry is the plain one
        res.append(engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10,
                                      sharded=True))
    engine.index_clear()
    engine.index_set_doc_base(0)
    for r in res[1:]:
        for a, b in zip(res[0], r):
            assert np.array_equal(a, b, equal_nan=True)
    assert (res[0][3] > 0).any() and res[0][0][res[0][0] != 0xFFFFFFFF].min() >= lo


def test_reload_with_more_docs_refreshes_blend_cache(engine):
    """ADVICE round 1 (index.cu:217): the cached blend vector must not survive a reload with a larger D."""
    V = 200
    rng = np.random.default_rng(3)
    probs = np.full(4, 0.25)
    q = synth.queries(64, V, seed=2)
    for D in (1000, 5000):
        engine.index_clear() if D == 1000 else None
        tabs = {}
        for tid in (capi.SS_BODY, capi.SS_TITLE):
            t = synth.index_table(V, D, tid, postings_per_doc=20.0 if tid else 5.0)
            if D == 5000 and tid == capi.SS_BODY:
                engine.index_clear()
            engine.index_load(tid, D, t.term_ptr, t.doc_ids, t.norm_tf)
            w, mag = engine.term_weights(tid, float(D), t.n_postings, D)
            tabs[tid] = (O.Table(t.term_ptr, t.doc_ids, w), mag)
        pr = rng.random((D, 4)) * 1e-3
        engine.set_pagerank(pr)
        got = engine.score_batch(q.kw_ptr, q.kw_terms, topic_probs=probs, k=10)
        exp = O.score_batch(tabs[0][0], tabs[1][0], D, tabs[0][1], tabs[1][1], pr, q.kw_ptr, q.kw_terms,
                            topic_probs=probs, k=10)
        assert_same_results(got, exp)


def test_live_topic_probabilities_extension(engine):
    """SURVEY.md 8(f)-3 (opt-in, beyond the shipped reference): the repaired computeTopicProbs feeding the
    PageRank blend per query.  ss_topic_probs is bit-exact against the oracle's restatement, and a batch scored
    with those per-query vectors matches the oracle end to end."""
    rng = np.random.default_rng(21)
    n_words, T = 500, 8
    rows = [sorted(rng.choice(T, size=int(rng.integers(0, 5)), replace=False)) for _ in range(n_words)]
    term_ptr = np.zeros(n_words + 1, np.uint64)
    term_ptr[1:] = np.cumsum([len(r) for r in rows])
    topic_ids = np.array([t for r in rows for t in r], np.uint32)
    freq = rng.integers(1, 50, len(topic_ids)).astype(np.float64)   # inv[2] stores counts (uint32)
    word_count = rng.integers(1000, 5000, T).astype(np.float64)     # forw[5] "wordCount"
    V, D = 300, 3000
    tabs = _small_index(engine, V, D)  # (ss_index_clear drops a topic table too: load it afterwards)
    engine.topics_load(term_ptr, topic_ids, freq, word_count)
    q = synth.queries(300, V, phrase_fraction=0.2, seed=12)
    # the query tokens' ids in inv[2]'s word space: here the same numbering, one token unknown to inv[2]
    tok = q.kw_terms.copy()
    tok[::17] = 100000
    got = engine.topic_probs(q.kw_ptr, tok)
    exp = O.topic_probs(term_ptr, topic_ids, freq, word_count, q.kw_ptr, tok)
    assert np.array_equal(got, exp) and (got > 0).any() and (got == 0).any()
    # one-token query listing topic t: exactly freq / wordCount / T (computed like Go: (1 * f/wc) / T)
    w = next(i for i, r in enumerate(rows) if len(r))
    one = engine.topic_probs([0, 1], [w])
    x = int(term_ptr[w])
    assert one[0, topic_ids[x]] == (1.0 * (freq[x] / word_count[topic_ids[x]])) / T
    pr = rng.random((D, T)) * 1e-2
    engine.set_pagerank(pr)
    res = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=got * 1e4, k=10)
    ref = O.score_batch(tabs[0][0], tabs[1][0], D, tabs[0][1], tabs[1][1], pr, q.kw_ptr, q.kw_terms, q.ph_ptr,
                        q.ph_terms, topic_probs=exp * 1e4, k=10)
    assert_same_results(res, ref)
    assert (res[2] != 0).any()  # the blend really is live


def test_large_batches_are_sliced(engine, monkeypatch):
    """Batches beyond 200K queries run as slices inside one call (BASELINE configs[4] is a 1M-query batch);
    with the slice size forced down, a 350-query mixed batch returns exactly the unsliced rows and the
    statistics of the whole batch."""
    V, D = 300, 3000
    _small_index(engine, V, D)
    q = synth.queries(350, V, phrase_fraction=0.3, seed=4)
    rng = np.random.default_rng(8)
    pr = rng.random((D, 4)) * 1e-3
    engine.set_pagerank(pr)
    probs = rng.random((350, 4))
    whole = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10)
    st_whole = engine.score_stats()
    monkeypatch.setenv("SS_SCORE_MAX_SLICE", "100")
    sliced = engine.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=10)
    st_sliced = engine.score_stats()
    for a, b in zip(whole, sliced):
        assert np.array_equal(a, b, equal_nan=True)
    assert st_sliced.postings_scanned == st_whole.postings_scanned and st_sliced.launches > st_whole.launches
