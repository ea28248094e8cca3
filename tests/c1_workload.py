"""BASELINE.json configs[0] at size: a synthetic 10k-page link graph + 50k-term inverted index written in the
reference's own table encodings (JSON values keyed by md5 hex, database/noschema_schema.go:125-260), run
through the C++ host mirror of the Go API (UpdateTopicSensitivePagerank, UpdateTermWeights x2, Retrieve of
100 queries) and checked against the oracle on the equivalent dense arrays.  Shared by
tests/test_host_mirror.py::test_c1_at_size and `bench.py --workload c1` (test infrastructure: it imports
the oracle)."""
import hashlib
import time

import numpy as np

from oracle import loader as O
from spaghettisearch_b200 import host, synth

N_PAGES, N_EDGES, N_TERMS, N_QUERIES, N_CATS = 10_000, 150_000, 50_000, 100, 16


def md5(s):
    return hashlib.md5(s.encode()).hexdigest()  # indexer/indexer.go:38-39,366-367


def make_tables(tmp_path, n_pages=N_PAGES, n_edges=N_EDGES, n_terms=N_TERMS, n_queries=N_QUERIES, seed=42):
    """Writes forw2 / forw5 / inv0 / inv1 snapshots; returns the dense views the oracle needs."""
    g = synth.graph(n_pages, n_edges, seed=seed)
    doc_key = [md5(f"http://site/{i}") for i in range(n_pages)]
    rows2 = []
    for u in range(n_pages):
        kids = g.col_idx[g.row_ptr[u]:g.row_ptr[u + 1]]
        rows2.append((doc_key[u], [doc_key[c] for c in kids] if len(kids) else None))  # uncrawled child lists are null
    host.write_jsonl(tmp_path / "forw2.jsonl", rows2)
    npg = synth.topics(N_CATS)
    host.write_jsonl(tmp_path / "forw5.jsonl",
                     [(f"Top{c:02d}", {"numPages": int(npg[c]), "wordCount": 1000 + c}) for c in range(N_CATS)])
    term_key = [md5(f"term{t}") for t in range(n_terms)]
    tabs = {}
    for name, tid in (("inv0", 0), ("inv1", 1)):
        t = synth.index_table(n_terms, n_pages, tid, with_positions=True, seed=seed + 1)
        rows = []
        for term in range(n_terms):
            a, b = int(t.term_ptr[term]), int(t.term_ptr[term + 1])
            if a == b:
                continue
            val = {}
            for p in range(a, b):
                pos = t.pos[int(t.pos_ptr[p]):int(t.pos_ptr[p + 1])]
                val[doc_key[t.doc_ids[p]]] = [float(t.norm_tf[p])] + [float(x) for x in pos]
            rows.append((term_key[term], val))
        host.write_jsonl(tmp_path / f"{name}.jsonl", rows)
        tabs[name] = t
    q = synth.queries(n_queries, n_terms, phrase_fraction=0.2, seed=seed + 2)
    return {"graph": g, "doc_key": doc_key, "term_key": term_key, "tabs": tabs, "queries": q, "num_pages": npg}


def run(engine, tmp_path, w, eps=1e-9):
    """The three Go-API calls over the snapshots; returns results and wall-clock seconds per call."""
    db = host.DB()
    times = {}
    try:
        t0 = time.perf_counter()
        for t in ("forw2", "forw5", "inv0", "inv1"):
            db.load(t, tmp_path / f"{t}.jsonl")
        times["load_snapshots"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        db.update_pagerank(engine, 0.75, eps)
        times["UpdateTopicSensitivePagerank"] = time.perf_counter() - t0
        t0 = time.perf_counter()
        db.update_term_weights(engine, "title")
        db.update_term_weights(engine, "body")
        times["UpdateTermWeights x2"] = time.perf_counter() - t0
        q, term_key = w["queries"], w["term_key"]
        qs = []
        for i in range(q.n_queries):
            kw = [term_key[t] for t in q.kw_terms[int(q.kw_ptr[i]):int(q.kw_ptr[i + 1])]]
            ph = [term_key[t] for t in q.ph_terms[int(q.ph_ptr[i]):int(q.ph_ptr[i + 1])]]
            qs.append((kw, ph))
        # the server keeps one retrieval::Index (loaded once, like retrieval.LoadIndex of the Go shim) and serves
        # its requests through the coalescing front-end; db.retrieve() would mirror a cold Retrieve call that
        # exports and uploads the whole index for every query
        t0 = time.perf_counter()
        results, n_calls, n_served = db.retrieve_concurrent(engine, qs, threads=8, window_us=200)
        times["Retrieve x%d" % len(qs)] = time.perf_counter() - t0  # includes building the resident index once
        times["score_batch_calls"] = n_calls
        assert n_served == len(qs)
        t0 = time.perf_counter()
        cold = db.retrieve(engine, qs[0][0], qs[0][1])  # one cold call: same rows as the served one
        times["cold Retrieve x1 (exports the index again)"] = time.perf_counter() - t0
        assert cold == results[0]
        for t in ("forw3", "forw4"):
            db.save(t, tmp_path / f"{t}.out.jsonl")
    finally:
        db.close()
    return results, times


def check(tmp_path, w, results, eps=1e-9):
    """forw[3], forw[4] and the 100 result lists against the oracle on the dense arrays.  Returns a dict of
    what was compared; raises AssertionError on a mismatch."""
    g, doc_key, tabs, q = w["graph"], w["doc_key"], w["tabs"], w["queries"]
    n = len(doc_key)
    f3 = host.read_jsonl(tmp_path / "forw3.out.jsonl")
    f4 = host.read_jsonl(tmp_path / "forw4.out.jsonl")
    assert len(f3) == n
    cats = sorted(f"Top{c:02d}" for c in range(N_CATS))
    t0 = time.perf_counter()
    ref, _ = O.pagerank(g.row_ptr, g.col_idx, 0.75, eps, w["num_pages"])
    cpu_pr_s = time.perf_counter() - t0
    got = np.array([[f3[doc_key[v]][c] for c in cats] for v in range(n)])
    l1 = float(np.abs(got - ref).sum(axis=0).max())
    assert l1 <= 1e-9, l1
    dense = {}
    t0 = time.perf_counter()
    for name, info in (("inv0", "title"), ("inv1", "body")):
        t = tabs[name]
        ow, omag = O.term_weights(t.term_ptr, t.doc_ids, t.norm_tf, n, float(n))
        for d in range(n):
            if omag[d] > 0:
                assert f4[doc_key[d]][info] == omag[d], (d, info)
        dense[name] = (O.Table(t.term_ptr, t.doc_ids, ow, t.pos_ptr, t.pos), omag)
    cpu_tw_s = time.perf_counter() - t0
    t0 = time.perf_counter()
    exp = O.score_batch(dense["inv0"][0], dense["inv1"][0], n, dense["inv0"][1], dense["inv1"][1], None, q.kw_ptr,
                        q.kw_terms, q.ph_ptr, q.ph_terms, k=50)
    cpu_sc_s = time.perf_counter() - t0
    n_hits = 0
    for i, res in enumerate(results):
        cnt = int(exp[3][i])
        assert len(res) == cnt, (i, len(res), cnt)
        want = sorted(((exp[1][i][j], doc_key[exp[0][i][j]]) for j in range(cnt)), key=lambda x: (-x[0], x[1]))
        if cnt < 50:  # complete list: (score desc, hash asc) is fully determined
            assert [r["DocHash"] for r in res] == [h for _, h in want], i
        assert np.allclose([r["FinalRank"] for r in res], [s for s, _ in want], rtol=1e-6, atol=0), i
        n_hits += cnt
    return {"pagerank_max_l1": l1, "queries": len(results), "results_compared": n_hits,
            "cpu_seconds": {"pagerank": cpu_pr_s, "term_weights": cpu_tw_s, "score_100": cpu_sc_s}}
