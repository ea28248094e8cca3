"""Synthetic workload generators: determinism, invariants, shard consistency."""
import numpy as np

from spaghettisearch_b200 import synth


def test_graph_invariants(built):
    g = synth.graph(20000, 300000, seed=42)
    assert g.row_ptr[0] == 0 and g.row_ptr[-1] == len(g.col_idx)
    assert abs(g.n_edges - 300000) < 0.05 * 300000
    od = np.diff(g.row_ptr.astype(np.int64))
    assert 0.15 < (od == 0).mean() < 0.25  # ~20 % dangling
    assert g.col_idx.max() < 20000
    for u in np.flatnonzero(od > 1)[:500]:  # children are a sorted set (crawler/crawler.go:163-170)
        row = g.col_idx[g.row_ptr[u]:g.row_ptr[u + 1]]
        assert (np.diff(row.astype(np.int64)) > 0).all()
    ind = np.bincount(g.col_idx, minlength=20000)
    assert ind.max() > 50 * ind.mean()  # power-law in-degree


def test_graph_deterministic_and_thread_independent(built):
    a = synth.graph(5000, 60000, seed=1, n_threads=1)
    b = synth.graph(5000, 60000, seed=1, n_threads=4)
    c = synth.graph(5000, 60000, seed=2, n_threads=4)
    assert np.array_equal(a.row_ptr, b.row_ptr) and np.array_equal(a.col_idx, b.col_idx)
    assert not np.array_equal(a.col_idx[:1000], c.col_idx[:1000])


def test_graph_row_slices_concatenate(built):
    full = synth.graph(30000, 400000, seed=5)
    parts = [synth.graph_rows(30000, 400000, lo, hi, seed=5, n_threads=2)
             for lo, hi in ((0, 7000), (7000, 7001), (7001, 30000))]
    assert sum(p.n_edges for p in parts) == full.n_edges
    assert np.array_equal(np.concatenate([p.col_idx for p in parts]), full.col_idx)
    assert np.array_equal(np.diff(full.row_ptr.astype(np.int64)),
                          np.concatenate([np.diff(p.row_ptr.astype(np.int64)) for p in parts]))


def test_graph_edge_cases(built):
    g = synth.graph(1, 10)
    assert g.n_nodes == 1 and g.n_edges <= 1
    g = synth.graph(0, 0)
    assert g.n_edges == 0
    g = synth.graph(10, 0)
    assert g.n_edges == 0


def test_index_invariants_and_sharding(built):
    V, D = 2000, 5000
    for table, ppd in ((1, 100.0), (0, 8.0)):
        full = synth.index_table(V, D, table, with_positions=True)
        assert abs(full.n_postings - ppd * D) < 0.1 * ppd * D
        assert np.array_equal(np.diff(full.term_ptr.astype(np.int64)), full.df_global.astype(np.int64))
        assert full.doc_ids.max() < D
        assert (full.norm_tf > 0).all() and (full.norm_tf <= 1).all()
        for t in (0, 1, 17, V - 1):
            row = full.doc_ids[full.term_ptr[t]:full.term_ptr[t + 1]]
            assert (np.diff(row.astype(np.int64)) > 0).all()
        assert full.pos_ptr[-1] == len(full.pos)
        # a doc shard is exactly the slice of the full table
        lo, hi = 1234, 3456
        part = synth.index_table(V, D, table, doc_lo=lo, doc_hi=hi, with_positions=True)
        assert np.array_equal(part.df_global, full.df_global)
        for t in (0, 3, 500, V - 1):
            frow = full.doc_ids[full.term_ptr[t]:full.term_ptr[t + 1]]
            ftf = full.norm_tf[full.term_ptr[t]:full.term_ptr[t + 1]]
            keep = (frow >= lo) & (frow < hi)
            prow = part.doc_ids[part.term_ptr[t]:part.term_ptr[t + 1]]
            assert np.array_equal(prow, frow[keep])
            assert np.array_equal(part.norm_tf[part.term_ptr[t]:part.term_ptr[t + 1]], ftf[keep])
    title = synth.index_table(V, D, 0, with_positions=True)
    assert (title.pos == -100.0).any()  # anchor/meta sentinel present (parser/parser.go:203)


def test_queries(built):
    q = synth.queries(5000, 10000, phrase_fraction=0.2)
    lens = np.diff(q.kw_ptr.astype(np.int64))
    assert lens.min() >= 1 and lens.max() <= 5
    assert abs((lens == 2).mean() - 0.35) < 0.03
    assert q.kw_terms.max() < 10000
    pl = np.diff(q.ph_ptr.astype(np.int64))
    assert set(np.unique(pl)) <= {0, 2, 3}
    assert abs((pl > 0).mean() - 0.2) < 0.03
    # Zipf(0.8): low ranks dominate
    assert (q.kw_terms < 100).mean() > 0.1
    q2 = synth.queries(5000, 10000, phrase_fraction=0.2)
    assert np.array_equal(q.kw_terms, q2.kw_terms) and np.array_equal(q.ph_terms, q2.ph_terms)
