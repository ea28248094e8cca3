"""The oracle against the known-answer fixtures (tests/golden/kats.json) and
its own cross-checks.  CPU only."""
import json
import math
from pathlib import Path

import numpy as np
import pytest

from oracle import loader as O
from spaghettisearch_b200 import synth
from tests.fixtures import tiny_index as _tiny_index

KATS = json.loads((Path(__file__).parent / "golden" / "kats.json").read_text())


def test_go_log2_exact_powers_of_two(built):
    for e in range(-20, 40):
        assert O.go_log2(2.0 ** e) == float(e)
    # log.go special cases
    assert math.isnan(O.go_log(-1.0))
    assert O.go_log(0.0) == -math.inf
    assert O.go_log(math.inf) == math.inf
    assert O.go_log(1.0) == 0.0


def test_go_log2_close_to_libm(built):
    rng = np.random.default_rng(0)
    for x in rng.uniform(1.0, 1e8, 2000):
        a, b = O.go_log2(float(x)), math.log2(float(x))
        assert abs(a - b) <= 4 * np.spacing(b)


def test_kat_pr1(built):
    k = KATS["KAT-PR-1"]
    for case in k["cases"]:
        rank, iters = O.pagerank(k["row_ptr"], k["col_idx"], k["damping"], case["eps"], [case["num_pages"]],
                                 case["max_iters"])
        assert iters[0] == case["iters"]
        assert rank[:, 0].tolist() == case["rank"]  # bit exact: same order of operations
    # SURVEY.md §8(c) seed values
    rank, iters = O.pagerank(k["row_ptr"], k["col_idx"], 0.75, 1e-20, [4, 10])
    assert iters.tolist() == [30, 30]
    assert rank[:, 0].tolist() == [0.2693673224993629, 0.24741289482615944, 0.3782054541450765,
                                   0.17621361167778027]
    assert rank[:, 0].tolist() == rank[:, 1].tolist()
    rank, iters = O.pagerank(k["row_ptr"], k["col_idx"], 0.75, 1e-9, [4, 10])
    assert iters.tolist() == [17, 17]
    assert rank[0].tolist() == [0.2693673224920345, 0.26936732249513917]


def test_pagerank_zero_pages_is_nan_after_one_sweep(built):
    k = KATS["KAT-PR-1"]
    rank, iters = O.pagerank(k["row_ptr"], k["col_idx"], 0.75, 1e-9, [0])
    assert iters[0] == 1 and np.isnan(rank).all()


def test_pagerank_variants_agree(built):
    g = synth.graph(3000, 40000, seed=7)
    npg = synth.topics(5)
    ref, it_ref = O.pagerank(g.row_ptr, g.col_idx, 0.75, 1e-9, npg)
    fair, it_fair, _ = O.pagerank_fair(g.row_ptr, g.col_idx, 0.75, 1e-9, npg, n_threads=4)
    assert it_ref.tolist() == it_fair.tolist()
    assert np.abs(ref - fair).sum(axis=0).max() < 1e-13
    faith, it_f, secs = O.pagerank_faithful(g.row_ptr, g.col_idx, 0.75, 1e-9, int(npg[2]))
    assert it_f == it_ref[2] and secs > 0
    assert np.abs(ref[:, 2] - faith).sum() < 1e-13


def test_kat_sc1(built):
    k = KATS["KAT-SC-1"]
    tabs = {}
    for name in ("title", "body"):
        t = k[name]
        w, mag = O.term_weights(t["term_ptr"], t["doc_ids"], t["norm_tf"], k["n_docs"], k["total_docs"])
        assert w.tolist() == k["w_" + name]
        assert mag.tolist() == k["mag_" + name]
        tabs[name] = (O.Table(t["term_ptr"], t["doc_ids"], w), mag)
    docs, final, pr, count = O.score_batch(tabs["title"][0], tabs["body"][0], k["n_docs"], tabs["title"][1],
                                           tabs["body"][1], None, [0, len(k["query"])], k["query"], k=50)
    assert count[0] == 3
    assert docs[0, :3].tolist() == [r["doc"] for r in k["result"]]
    assert final[0, :3].tolist() == [r["final"] for r in k["result"]]
    assert (pr[0] == 0).all() and (docs[0, 3:] == 0xFFFFFFFF).all()
    # SURVEY.md §8(c) seed values
    assert final[0, :3].tolist() == [47.376154339498676, 27.5118156434649, 3.371181523570759]


def _weighted(built, total_docs=6.0):
    (tp, td, ttf, tpp, tpos), (bp, bd, btf, bpp, bpos), n_docs = _tiny_index()
    tw, tmag = O.term_weights(tp, td, ttf, n_docs, total_docs)
    bw, bmag = O.term_weights(bp, bd, btf, n_docs, total_docs)
    return O.Table(tp, td, tw, tpp, tpos), O.Table(bp, bd, bw, bpp, bpos), tmag, bmag, n_docs


def test_intersect_util(built):
    assert O.intersect([3, 1, 2], [2, 3, 4]) == [2, 3]
    assert O.intersect([1, 1, 2], [1, 1, 1]) == [1, 1]  # multiset: util.go advances both cursors
    assert O.intersect(None, [1]) == [] and O.intersect([1], None) == []
    assert O.intersect([], [1]) == []


def test_phrase_semantics(built):
    title, body, tmag, bmag, n_docs = _weighted(built)
    qm2 = math.sqrt(2.0)

    def run(kw, ph):
        kw_ptr, ph_ptr = [0, len(kw)], [0, len(ph)]
        d, f, p, c = O.score_batch(title, body, n_docs, tmag, bmag, None, kw_ptr, kw, ph_ptr, ph, k=10)
        return d[0, :c[0]].tolist(), f[0, :c[0]].tolist()

    # phrase [t0 t1]: body adjacency needs pos(t1) = pos(t0)+1.  doc0: t0@{1,7}, t1@{2,30} -> match;
    # doc1: t0@{2}, t1@{5} -> no.  Title: doc2 has t0@0 and t1@1 -> title match.
    docs, final = run([], [0, 1])
    assert sorted(docs) == [0, 2]
    w_b = float(np.float32(body.w[0] + body.w[3]))  # fp32 sum in phrase order, phrase.go:59,83
    exp0 = (0.38 * 0.0 + 0.29 * (w_b / (bmag[0] * qm2))) * 100.0
    assert final[docs.index(0)] == exp0
    # the -100 sentinel aligns for a ONE word phrase (phrase.go:68-75), never across words
    docs, _ = run([], [2])
    assert sorted(docs) == [0, 1, 3]
    # doc1 has t0@-100 and t2@-100 in its title: both slots filled, but -100 and -100-1 never align,
    # and t2 has no body posting in doc1 -> not a match.  doc0 matches in the body (t0@3? no: t0@{1,7},
    # t2@{3} -> 3-1=2 not in {1,7}) -> nothing at all.
    docs, _ = run([], [0, 2])
    assert docs == []
    docs, _ = run([], [1, 2])  # body doc0: t1@{2,30}, t2@{3} -> 3-1 = 2 aligns
    assert docs == [0]
    # a posting without positions cannot match even a one-word phrase
    docs, _ = run([], [3])
    assert docs == []
    # keyword + phrase: queryLength counts both (main_retrieve.go:90)
    docs, final = run([3], [0, 1])
    assert 5 in docs and 0 in docs
    # unknown term ids are empty rows; a phrase with an unknown term matches nothing
    docs, _ = run([99], [0, 99])
    assert docs == []


def test_duplicate_query_terms_count_twice(built):
    title, body, tmag, bmag, n_docs = _weighted(built)
    d1, f1, _, c1 = O.score_batch(title, body, n_docs, tmag, bmag, None, [0, 1], [0], k=10)
    d2, f2, _, c2 = O.score_batch(title, body, n_docs, tmag, bmag, None, [0, 2], [0, 0], k=10)
    assert c1[0] == c2[0]
    # sums double, |q| goes 1 -> sqrt(2): scores scale by sqrt(2)
    m1 = dict(zip(d1[0, :c1[0]].tolist(), f1[0, :c1[0]].tolist()))
    m2 = dict(zip(d2[0, :c2[0]].tolist(), f2[0, :c2[0]].tolist()))
    for d in m1:
        assert m2[d] == pytest.approx(m1[d] * math.sqrt(2.0), rel=1e-15)


def test_idf_zero_gives_zero_not_nan(built):
    # df == totalDocs -> idf = 0 -> w = 0, mag = 0 -> 0/0 = NaN -> 0 (get_metadata.go:61-66)
    ptr = np.array([0, 3], np.uint64)
    docs = np.array([0, 1, 2], np.uint32)
    tf = np.array([1.0, 0.5, 0.25], np.float32)
    w, mag = O.term_weights(ptr, docs, tf, 3, 3.0)
    assert (w == 0).all() and (mag == 0).all()
    tab = O.Table(ptr, docs, w)
    empty = O.Table(np.array([0, 0], np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.float32))
    d, f, p, c = O.score_batch(empty, tab, 3, mag, mag, None, [0, 1], [0], k=5)
    assert c[0] == 3 and d[0, :3].tolist() == [0, 1, 2] and (f[0, :3] == 0).all()


def test_blend_and_tie_order(built):
    title, body, tmag, bmag, n_docs = _weighted(built)
    rng = np.random.default_rng(3)
    pr = rng.uniform(0, 1e-3, (n_docs, 4))
    probs = np.array([0.1, 0.2, 0.3, 0.4])
    d, f, p, c = O.score_batch(title, body, n_docs, tmag, bmag, pr, [0, 2], [0, 1], topic_probs=probs, k=10)
    for j in range(c[0]):
        sqd = 0.0
        for t in range(4):
            sqd += probs[t] * pr[d[0, j], t]
        assert p[0, j] == sqd
    assert all(f[0, j] >= f[0, j + 1] for j in range(c[0] - 1))
    # per-query probabilities
    probs2 = np.stack([probs, probs[::-1]])
    d2, f2, p2, c2 = O.score_batch(title, body, n_docs, tmag, bmag, pr, [0, 2, 4], [0, 1, 0, 1],
                                   topic_probs=probs2, k=10)
    assert d2[0].tolist() == d[0].tolist() and f2[0].tolist() == f[0].tolist()
    assert not np.array_equal(p2[0], p2[1])


def test_score_batch_fair_flavour_is_identical(built):
    # the CPU baseline's "fair" flavour (dense accumulators, shared blend term, bounded selection) must return
    # exactly what the hash-map flavour returns: keyword + phrase queries, no blend / shared / per-query blend
    from spaghettisearch_b200 import synth
    V, D = 3000, 8000
    t = synth.index_table(V, D, 0, with_positions=True)
    b = synth.index_table(V, D, 1, with_positions=True)
    wt, mt = O.term_weights(t.term_ptr, t.doc_ids, t.norm_tf, D, float(D))
    wb, mb = O.term_weights(b.term_ptr, b.doc_ids, b.norm_tf, D, float(D))
    ot, ob = O.Table(t.term_ptr, t.doc_ids, wt, t.pos_ptr, t.pos), O.Table(b.term_ptr, b.doc_ids, wb, b.pos_ptr, b.pos)
    q = synth.queries(400, V, phrase_fraction=0.3, seed=51)
    rng = np.random.default_rng(3)
    pr = rng.uniform(0, 1e-4, (D, 16))
    for probs in (None, np.full(16, 1 / 16), rng.dirichlet(np.ones(16), size=400)):
        for k in (1, 10, 50):
            a = O.score_batch(ot, ob, D, mt, mb, pr, q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=k)
            f = O.score_batch(ot, ob, D, mt, mb, pr, q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=k,
                              fair=True, n_threads=4)
            assert np.array_equal(a[0], f[0]) and np.array_equal(a[3], f[3])
            assert np.array_equal(a[1].view(np.uint64), f[1].view(np.uint64))
            assert np.array_equal(a[2].view(np.uint64), f[2].view(np.uint64))


def test_extension_kats(built):
    """The two opt-in extensions against the pure-Python restatement in make_kats.py."""
    k = KATS["KAT-PR-2-biased"]
    rank, iters = O.pagerank_biased(np.array(k["row_ptr"], np.uint64), np.array(k["col_idx"], np.uint32), k["damping"],
                                    k["eps"], np.array(k["num_pages"], np.int64), np.array(k["weights"], np.float64),
                                    n_threads=1)
    assert iters.tolist() == k["iters"]
    assert np.abs(rank - np.array(k["rank"])).max() <= 1e-15
    t = KATS["KAT-TP-1"]
    probs = O.topic_probs(t["term_ptr"], t["topic_ids"], t["freq"], t["word_count"], t["tok_ptr"], t["tok_terms"])
    assert probs.tolist() == t["probs"]  # bit exact: same operations in the same order


def _random_go_tables(rng, n_docs, n_terms):
    """Two inverted tables in Go shape ({term: {doc: [w, pos...]}}) + the equivalent CSR arrays (weights are
    final tf-idf weights here: the scorer never looks at how they were made)."""
    docs = [f"{i:04d}" for i in range(n_docs)]     # ascending string order == ascending dense id
    terms = [f"t{i:02d}" for i in range(n_terms)]
    inv, csr = [], []
    for tb in range(2):
        table, ptr, ids, w, pp, pos = {}, [0], [], [], [0], []
        for t in range(n_terms):
            row = {}
            for d in sorted(rng.choice(n_docs, size=int(rng.integers(0, n_docs // 2)), replace=False)):
                weight = np.float32(rng.choice([0.25, 0.5, 1.0, 1.5, 3.0]))
                kind = rng.integers(0, 4)
                if kind == 0:
                    p = []                                         # posting without positions
                elif kind == 1:
                    p = [-100.0]                                   # meta/anchor sentinel (parser.go:195-207)
                else:
                    p = sorted(float(x) for x in rng.choice(12, size=int(rng.integers(1, 4)), replace=False))
                row[docs[d]] = [float(weight)] + p
                ids.append(d); w.append(weight); pos.extend(p); pp.append(len(pos))
            if row:
                table[terms[t]] = row
            ptr.append(len(ids))
        inv.append(table)
        csr.append(O.Table(np.array(ptr, np.uint64), np.array(ids, np.uint32), np.array(w, np.float32),
                           np.array(pp, np.uint64), np.array(pos, np.float32)))
    return docs, terms, inv, csr


def test_oracle_matches_literal_go_translation(built):
    """The oracle against tests/golden/go_literal.py -- a statement-by-statement Python translation of
    getFromInverted, getPosTerm, evalPhraseOccurrence, intersect, genAggrDocsPipeline, computeFinalRank and
    appendSort on Go-shaped maps -- over random small indexes: keyword duplicates, unknown terms, one- to
    four-word phrases with repeated words, postings without positions, the -100 sentinel, zero norms, blend."""
    from tests.golden import go_literal as G
    rng = np.random.default_rng(123)
    n_checked = n_phrase_hits = 0
    for trial in range(6):
        n_docs, n_terms = 40, 10
        docs, terms, inv, (ot, ob) = _random_go_tables(rng, n_docs, n_terms)
        tmag = rng.choice([0.0, 0.5, 1.0, 2.0, 7.5], size=n_docs)
        bmag = rng.choice([0.0, 1.0, 3.0, 4.25], size=n_docs)
        mag = {docs[d]: {"title": float(tmag[d]), "body": float(bmag[d])} for d in range(n_docs)}
        T = 3
        pr = rng.random((n_docs, T))
        topics = [f"c{t}" for t in range(T)]
        pr_map = {docs[d]: {topics[t]: float(pr[d, t]) for t in range(T)} for d in range(n_docs)}
        for use_blend in (False, True):
            probs = rng.random(T) if use_blend else None
            kws, phs = [], []
            for _ in range(60):
                kws.append([int(x) for x in rng.integers(0, n_terms + 1, size=int(rng.integers(0, 4)))])  # n_terms = unknown
                phs.append([int(x) for x in rng.integers(0, n_terms, size=int(rng.integers(0, 5)))])
            kw_ptr = np.zeros(len(kws) + 1, np.uint64); kw_ptr[1:] = np.cumsum([len(x) for x in kws])
            ph_ptr = np.zeros(len(phs) + 1, np.uint64); ph_ptr[1:] = np.cumsum([len(x) for x in phs])
            kw = np.array([t if t < n_terms else 0xFFFFFFFF for x in kws for t in x], np.uint32)
            ph = np.array([t for x in phs for t in x], np.uint32)
            d, f, p, c = O.score_batch(ot, ob, n_docs, tmag, bmag, pr if use_blend else None, kw_ptr, kw, ph_ptr, ph,
                                       topic_probs=probs, k=50)
            for qi in range(len(kws)):
                name = lambda t: terms[t] if t < n_terms else "unknown"
                lit = G.retrieve([name(t) for t in kws[qi]], [name(t) for t in phs[qi]], inv, mag, pr_map,
                                 {topics[t]: float(probs[t]) for t in range(T)} if use_blend else None)
                # the literal code keeps NaN/inf as Go would; the oracle's order puts NaN last: skip those queries
                if any(math.isnan(x[1]) for x in lit):
                    continue
                got = [(docs[d[qi, j]], f[qi, j], p[qi, j]) for j in range(int(c[qi]))]
                assert [g[0] for g in got] == [x[0] for x in lit], (trial, qi, kws[qi], phs[qi])
                assert [g[1] for g in got] == [x[1] for x in lit], (trial, qi)     # bit exact: same operations
                assert [g[2] for g in got] == [x[2] for x in lit], (trial, qi)
                n_checked += 1
                if phs[qi] and G.get_phrase_from_inverted([name(t) for t in phs[qi]], inv):
                    n_phrase_hits += 1
    assert n_checked > 600 and n_phrase_hits > 50
