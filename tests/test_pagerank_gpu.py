"""HP-1 parity: the CUDA PageRank through the C ABI against the oracle
(ranking/pagerank.go:14-145).  Tolerance from BASELINE.json north_star:
fp64 within 1e-9 L1 per topic; equal per-topic sweep counts."""
import json
from pathlib import Path

import numpy as np
import pytest

from oracle import loader as O
from spaghettisearch_b200 import capi, synth

pytestmark = pytest.mark.gpu
KATS = json.loads((Path(__file__).parent / "golden" / "kats.json").read_text())
L1_TOL = 1e-9


def _check(engine, row_ptr, col_idx, d, eps, npg, max_iters=0, fair_threads=0):
    engine.graph_load_csr(row_ptr, col_idx)
    got, it_got, status = engine.pagerank(d, eps, npg, max_iters)
    if fair_threads:
        ref, it_ref, _ = O.pagerank_fair(row_ptr, col_idx, d, eps, npg, max_iters, n_threads=fair_threads)
    else:
        ref, it_ref = O.pagerank(row_ptr, col_idx, d, eps, npg, max_iters)
    assert it_got.tolist() == it_ref.tolist()
    l1 = np.abs(got - ref).sum(axis=0)
    assert (l1 <= L1_TOL).all(), l1
    return got, it_got, status


def test_kat_pr1(engine):
    k = KATS["KAT-PR-1"]
    engine.graph_load_csr(k["row_ptr"], k["col_idx"])
    for case in k["cases"]:
        rank, iters, _ = engine.pagerank(k["damping"], case["eps"], [case["num_pages"]], case["max_iters"])
        if case["eps"] >= 1e-12:  # at 1e-20 the stop is an exact-fixed-point question
            assert iters[0] == case["iters"]
        assert np.abs(rank[:, 0] - np.array(case["rank"])).sum() <= 1e-12


@pytest.mark.parametrize("n_topics", [1, 2, 3, 4, 5, 8, 11, 16])
def test_topic_widths(engine, n_topics):
    g = synth.graph(5000, 70000, seed=11)
    _check(engine, g.row_ptr, g.col_idx, 0.75, 1e-9, synth.topics(n_topics))


@pytest.mark.parametrize("n,e,seed", [(300, 3000, 1), (20000, 300000, 2), (200000, 3000000, 3)])
def test_random_graphs(engine, n, e, seed):
    g = synth.graph(n, e, seed=seed)
    got, iters, status = _check(engine, g.row_ptr, g.col_idx, 0.75, 1e-9, synth.topics(16))
    assert status == 0 and (iters >= 2).all()
    # the reference's ranks do not sum to one, but every column is positive and finite
    assert np.isfinite(got).all() and (got > 0).all()


def test_short_row_kernels_agree_bitwise(engine, monkeypatch):
    # The three short-row kernels -- first version (legacy), lean registers (k_sweep_short32, the
    # default) and the opt-in cp.async gather ring (k_sweep_short_async) -- use the same gather order and the same
    # epilogue expressions, so their results must be identical bit for bit.  Frozen topics (different
    # num_pages -> different sweep counts) exercise the masked instantiations; the topic counts cover
    # every lane shape.
    g = synth.graph(60000, 900000, seed=13)
    monkeypatch.setenv("SS_PR_SHORT", "async")  # the ring kernel's short-only CSR is built at load on request
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    cases = [(synth.topics(16), 1e-9), ([3, 50000, 7, 1000000, 11], 1e-13), (synth.topics(8), 1e-9),
             (synth.topics(3), 1e-9), ([50000, 9], 1e-11), ([12345], 1e-9)]
    for npg, eps in cases:
        out = {}
        for mode in ("legacy", "lean", "async"):
            monkeypatch.setenv("SS_PR_SHORT", mode)
            out[mode] = engine.pagerank(0.75, eps, npg)
        monkeypatch.delenv("SS_PR_SHORT")
        for mode in ("lean", "async"):
            assert out[mode][1].tolist() == out["legacy"][1].tolist(), mode
            assert np.array_equal(out[mode][0].view(np.uint64), out["legacy"][0].view(np.uint64)), mode


def test_damping_085_and_tight_eps(engine):
    g = synth.graph(20000, 300000, seed=5)
    _check(engine, g.row_ptr, g.col_idx, 0.85, 1e-12, synth.topics(16))


def test_long_rows_and_fixup(engine):
    # a star: every node links to node 0 (one row with N-1 in-edges -> many tasks + fix-up),
    # plus a ring so that other rows are short
    n = 40000
    src = np.arange(1, n, dtype=np.uint32)
    rows = [[] for _ in range(n)]
    row_ptr = np.zeros(n + 1, dtype=np.uint64)
    col = []
    for u in range(n):
        kids = sorted({0, (u + 1) % n}) if u else [1]
        col.extend(kids)
        row_ptr[u + 1] = len(col)
    _check(engine, row_ptr, np.array(col, dtype=np.uint32), 0.75, 1e-9, synth.topics(16))


def test_edge_cases(engine):
    npg = synth.topics(4)
    # no edges at all: everything dangling
    row_ptr = np.zeros(11, dtype=np.uint64)
    _check(engine, row_ptr, np.zeros(0, dtype=np.uint32), 0.75, 1e-9, npg)
    # one node with a self loop
    _check(engine, np.array([0, 1], np.uint64), np.array([0], np.uint32), 0.75, 1e-9, npg)
    # duplicate children count twice (pagerank.go:140-142 iterates list entries)
    _check(engine, np.array([0, 3, 4, 4], np.uint64), np.array([1, 1, 2, 0], np.uint32), 0.75, 1e-9, npg)
    # numPages = 0 -> 1/0 = +Inf -> NaN ranks, loop ends after one sweep (NaN > eps is false)
    g = synth.graph(500, 5000, seed=9)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    rank, iters, _ = engine.pagerank(0.75, 1e-9, [0, 50000])
    ref, it_ref = O.pagerank(g.row_ptr, g.col_idx, 0.75, 1e-9, [0, 50000])
    assert iters.tolist() == it_ref.tolist() and iters[0] == 1
    assert np.isnan(rank[:, 0]).all() and np.isnan(ref[:, 0]).all()
    assert np.abs(rank[:, 1] - ref[:, 1]).sum() <= L1_TOL
    # empty forw[5]: no topics, nothing to write
    rank, iters, status = engine.pagerank(0.75, 1e-9, [])
    assert status == 0 and rank.shape == (500, 0)


def test_max_iters_status_and_fetch(engine):
    g = synth.graph(3000, 40000, seed=4)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    rank, iters, status = engine.pagerank(0.75, 1e-30, synth.topics(3), max_iters=2)
    assert status == 1 and iters.tolist() == [2, 2, 2]  # SS_NOT_CONVERGED
    ref, _ = O.pagerank(g.row_ptr, g.col_idx, 0.75, 1e-30, synth.topics(3), 2)
    assert np.abs(rank - ref).sum(axis=0).max() <= L1_TOL
    part = engine.pagerank_fetch(100, 200)
    assert np.array_equal(part, rank[100:200])
    # result left on the device, fetched later
    _, iters2, _ = engine.pagerank(0.75, 1e-9, synth.topics(3), want_rank=False)
    assert np.abs(engine.pagerank_fetch(0, 3000) - O.pagerank(g.row_ptr, g.col_idx, 0.75, 1e-9,
                                                               synth.topics(3))[0]).sum() <= 3e-9


def test_reference_call_site_eps(engine):
    # cmd/crawl/start_crawl.go:175 calls with eps = 1e-20: run-until-fixed-point.  The oracle reaches an
    # exact fixed point; the engine stops on its bit-for-bit fixed-point guard.  Ranks must agree.
    g = synth.graph(2000, 30000, seed=6)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    rank, iters, status = engine.pagerank(0.75, 1e-20, synth.topics(16), max_iters=200)
    ref, it_ref = O.pagerank(g.row_ptr, g.col_idx, 0.75, 1e-20, synth.topics(16), 200)
    assert status == 0 and iters.max() < 200
    assert np.abs(rank - ref).sum(axis=0).max() <= 1e-12


def test_invalid_arguments(engine):
    from spaghettisearch_b200 import capi
    with pytest.raises(capi.SSError):  # child id out of range
        engine.graph_load_csr(np.array([0, 1], np.uint64), np.array([7], np.uint32))
    with pytest.raises(capi.SSError):  # row_ptr not monotone
        engine.graph_load_csr(np.array([0, 2, 1, 2], np.uint64), np.array([0, 1], np.uint32))
    g = synth.graph(100, 1000, seed=1)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    with pytest.raises(capi.SSError):
        engine.pagerank(0.75, 1e-9, synth.topics(17))


def test_million_node_graph_against_fair_oracle(engine):
    g = synth.graph(1_000_000, 15_000_000, seed=42)
    _check(engine, g.row_ptr, g.col_idx, 0.75, 1e-9, synth.topics(16), fair_threads=8)
    st = engine.pagerank_stats()
    assert st.sweeps >= 3 and st.launches > 0 and st.sweep_ms_total > 0


def test_topic_biased_teleport_extension(engine):
    """SURVEY.md 8(f)-4 (opt-in, beyond the shipped reference): per-topic teleport vectors.  All-ones weights are
    the reference's uniform teleport bit for bit; a genuinely biased vector matches the oracle's extension."""
    n = 30000
    g = synth.graph(n, 400000, seed=5)
    npg = synth.topics(6)
    engine.graph_load_csr(g.row_ptr, g.col_idx)
    base, it0, _ = engine.pagerank(0.75, 1e-9, npg)
    engine.pagerank_set_teleport(np.ones((n, 6)))
    same, it1, _ = engine.pagerank(0.75, 1e-9, npg)
    assert np.array_equal(base, same) and it0.tolist() == it1.tolist()
    rng = np.random.default_rng(2)
    w = np.zeros((n, 6))
    for t in range(6):  # topic t teleports only to its own page set (Haveliwala), weights sum to n per topic
        pages = rng.choice(n, size=500 + 100 * t, replace=False)
        w[pages, t] = n / len(pages)
    engine.pagerank_set_teleport(w)
    got, iters, status = engine.pagerank(0.75, 1e-9, npg)
    ref, it_ref = O.pagerank_biased(g.row_ptr, g.col_idx, 0.75, 1e-9, npg, w)
    assert status == 0 and iters.tolist() == it_ref.tolist()
    assert np.abs(got - ref).sum(axis=0).max() <= 1e-9
    assert np.abs(got - base).sum() > 1e-3  # it really is a different ranking
    with pytest.raises(capi.SSError):
        engine.pagerank(0.75, 1e-9, npg[:4])  # weights were set for 6 topics
    engine.pagerank_set_teleport(None)
    again, _, _ = engine.pagerank(0.75, 1e-9, npg)
    assert np.array_equal(base, again)
