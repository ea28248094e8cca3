"""The N>1 host logic on CPU: two gloo ranks.

* graph slices generated per rank and exchanged reproduce the single-process graph;
* doc-sharded scoring: each rank scores the whole batch against its shard (global df),
  per-shard top-k lists merged with the reference comparator equal the unsharded result.
  The per-shard scorer here is the oracle (no GPU on this box); the same protocol runs
  on GPUs in tests/test_multigpu.py."""
import os
import sys
import tempfile

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

N_NODES, N_EDGES = 20000, 250000
V, D, Q, K = 3000, 8000, 200, 10


def _merge_ref(docs, finals, prs, counts, k):
    """Numpy restatement of the result order: FinalRank desc, doc id asc, NaN last."""
    world, nq, _ = docs.shape
    out_d = np.full((nq, k), 0xFFFFFFFF, np.uint32)
    out_f = np.zeros((nq, k))
    out_c = np.zeros(nq, np.uint32)
    for q in range(nq):
        items = []
        for r in range(world):
            for j in range(int(counts[r, q])):
                f = finals[r, q, j]
                items.append((np.isnan(f), -f if not np.isnan(f) else 0.0, int(docs[r, q, j]), f))
        items.sort(key=lambda x: (x[0], x[1], x[2]))
        out_c[q] = min(k, len(items))
        for j, it in enumerate(items[:k]):
            out_d[q, j], out_f[q, j] = it[2], it[3]
    return out_d, out_f, out_c


def _worker(rank, world, init_file, ret):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import loader as O
    from spaghettisearch_b200 import sharding, synth
    dist.init_process_group("gloo", init_method=f"file://{init_file}", rank=rank, world_size=world)
    try:
        # ---- graph slices -> full graph on every rank
        lo, hi = sharding.row_slice(rank, world, N_NODES)
        part = synth.graph_rows(N_NODES, N_EDGES, lo, hi, seed=42, n_threads=2)
        row_ptr, col_idx = sharding.assemble_graph(N_NODES, part.row_ptr, part.col_idx)
        full = synth.graph(N_NODES, N_EDGES, seed=42, n_threads=2)
        ok_graph = np.array_equal(row_ptr, full.row_ptr) and np.array_equal(col_idx, full.col_idx)
        # ---- the unique id travels as bytes
        uid = sharding.share_unique_id(lambda: bytes(range(128)))
        ok_uid = uid == bytes(range(128))
        # one id per engine group: two groups of one rank each get their own ids, one group of two shares one
        own = sharding.share_group_unique_id(lambda: bytes([rank]) * 128, 0, 1)
        both = sharding.share_group_unique_id(lambda: bytes([7]) * 128, rank, 2)
        ok_uid = ok_uid and own == bytes([rank]) * 128 and both == bytes([7]) * 128
        # ---- doc-sharded scoring + merge
        dlo, dhi = sharding.doc_shard(rank, world, D)
        tabs = []
        for tid in (0, 1):
            t = synth.index_table(V, D, tid, doc_lo=dlo, doc_hi=dhi, with_positions=True, n_threads=2)
            w = t.norm_tf.copy()
            mag2 = np.zeros(D)
            for term in range(V):  # idf from GLOBAL df (ss_term_weights df_global)
                a, b = int(t.term_ptr[term]), int(t.term_ptr[term + 1])
                if a == b:
                    continue
                idf = np.float32(O.go_log2(float(D) / float(t.df_global[term])))
                w[a:b] = t.norm_tf[a:b] * idf
                sq = (w[a:b] * w[a:b]).astype(np.float64)
                np.add.at(mag2, t.doc_ids[a:b], sq)
            tabs.append((O.Table(t.term_ptr, t.doc_ids, w, t.pos_ptr, t.pos), np.sqrt(mag2)))
        q = synth.queries(Q, V, phrase_fraction=0.25, seed=44)
        local = O.score_batch(tabs[0][0], tabs[1][0], D, tabs[0][1], tabs[1][1], None, q.kw_ptr, q.kw_terms,
                              q.ph_ptr, q.ph_terms, k=K, n_threads=2)
        docs, finals, prs, counts = sharding.gather_result_lists(*local)
        ok_merge = True
        if rank == 0:
            md, mf, mc = _merge_ref(docs, finals, prs, counts, K)
            ft = synth.index_table(V, D, 0, with_positions=True, n_threads=2)
            fb = synth.index_table(V, D, 1, with_positions=True, n_threads=2)
            wt, mt = O.term_weights(ft.term_ptr, ft.doc_ids, ft.norm_tf, D, float(D))
            wb, mb = O.term_weights(fb.term_ptr, fb.doc_ids, fb.norm_tf, D, float(D))
            ref = O.score_batch(O.Table(ft.term_ptr, ft.doc_ids, wt, ft.pos_ptr, ft.pos),
                                O.Table(fb.term_ptr, fb.doc_ids, wb, fb.pos_ptr, fb.pos), D, mt, mb, None, q.kw_ptr,
                                q.kw_terms, q.ph_ptr, q.ph_terms, k=K, n_threads=2)
            ok_merge = (np.array_equal(md, ref[0]) and np.array_equal(mc, ref[3]) and
                        np.allclose(mf, ref[1], rtol=1e-12, atol=0))
        ret[rank] = (bool(ok_graph), bool(ok_uid), bool(ok_merge))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo(built):
    world = 2
    with tempfile.TemporaryDirectory() as tmp:
        mgr = mp.Manager()
        ret = mgr.dict()
        mp.spawn(_worker, args=(world, os.path.join(tmp, "init"), ret), nprocs=world, join=True)
        assert len(ret) == world
        for r in range(world):
            assert ret[r] == (True, True, True), (r, ret[r])


def test_engine_grid():
    from spaghettisearch_b200 import sharding
    # 8 ranks as 4 row groups x 2 topic groups over 16 topics: ranks 0-3 run topics 0-7, ranks 4-7 topics 8-15
    grid = [sharding.engine_grid(r, 4, 2, 16) for r in range(8)]
    assert [g[0] for g in grid] == [0, 1, 2, 3, 0, 1, 2, 3]
    assert [g[1] for g in grid] == [0, 0, 0, 0, 1, 1, 1, 1]
    assert [g[2:] for g in grid] == [(0, 8)] * 4 + [(8, 8)] * 4
    assert sharding.engine_grid(0, 1, 1, 16) == (0, 0, 0, 16)
    covered = sorted(t for g in {gr[1:] for gr in grid} for t in range(g[1], g[1] + g[2]))
    assert covered == list(range(16))


def test_shard_bounds():
    from spaghettisearch_b200 import sharding
    for n, w in ((10, 3), (1000003, 8), (5, 8), (0, 2)):
        b = [sharding.doc_shard(r, w, n) for r in range(w)]
        assert b[0][0] == 0 and b[-1][1] == n
        assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
        assert all(hi >= lo for lo, hi in b)
