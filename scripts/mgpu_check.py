"""Multi-GPU parity check, one process per GPU (run under torchrun):
row-partitioned PageRank against the oracle and doc-sharded scoring + merge."""
import os, sys
import numpy as np
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import loader as O
from spaghettisearch_b200 import capi, sharding, synth

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = capi.Engine(device=local, timing=True)
eng.comm_init(sharding.share_unique_id(capi.comm_unique_id), rank, world)

# ---- HP-1: every load / exchange combination against the oracle
N, E = 300000, 4000000
lo, hi = sharding.row_slice(rank, world, N)
part = synth.graph_rows(N, E, lo, hi, seed=42, n_threads=4)
row_ptr, col_idx = sharding.assemble_graph(N, part.row_ptr, part.col_idx, device="cuda")
npg = synth.topics(16)
ref, it_ref, _ = O.pagerank_fair(row_ptr, col_idx, 0.75, 1e-9, npg, n_threads=4)
ok_pr = True
for load, env in (("full", {}), ("rows", {}), ("rows", {"SS_PR_EXCHANGE": "nccl", "SS_PR_CHUNKS": "4"}),
                  ("full", {"SS_PR_EXCHANGE": "nccl", "SS_PR_CHUNKS": "3"}), ("rows", {"SS_PR_EXCHANGE": "nccl"}),
                  ("rows", {"SS_PR_EXCHANGE": "copy", "SS_PR_CHUNKS": "4"}), ("full", {"SS_PR_EXCHANGE": "copy"})):
    os.environ.update(env)
    if load == "full":
        eng.graph_load_csr(row_ptr, col_idx)
    else:  # sharded export: this rank's slice only
        eng.graph_load_csr_rows(N, lo, hi, part.row_ptr, part.col_idx)
    for n_t in (16, 8):
        rank_all, iters, status = eng.pagerank(0.75, 1e-9, npg[:n_t])
        st = eng.pagerank_stats()
        l1 = np.abs(rank_all - ref[:, :n_t]).sum(axis=0).max()
        own = eng.pagerank_fetch(int(st.row_lo), int(st.row_lo + st.local_rows))
        ok = (status == 0 and iters.tolist() == it_ref[:n_t].tolist() and l1 <= 1e-9 and
              np.array_equal(own, rank_all[int(st.row_lo): int(st.row_lo + st.local_rows)]))
        ok_pr = ok_pr and ok
        if rank == 0:
            print(f"pagerank load={load} env={env} topics={n_t}: sweeps {st.sweeps} L1 {l1:.2e} "
                  f"exposed exchange {st.exchange_ms_total / max(1, st.sweeps):.3f} ms/sweep, busy "
                  f"{st.exchange_busy_ms_total / max(1, st.sweeps):.3f} -> {'ok' if ok else 'FAILED'}", flush=True)
    for k in env:
        del os.environ[k]
rows = torch.tensor([st.local_rows, st.local_edges], dtype=torch.int64, device="cuda")
allrows = [torch.zeros_like(rows) for _ in range(world)]
dist.all_gather(allrows, rows)
if rank == 0:
    print("partition rows/edges:", [a.tolist() for a in allrows], flush=True)
    assert sum(int(a[0]) for a in allrows) == N and sum(int(a[1]) for a in allrows) == int(row_ptr[-1])

# ---- HP-2: doc-sharded index under shard-local doc ids, global df, cross-shard merge inside the engine
V, D, Q, K = 5000, 40000, 500, 10
dlo, dhi = sharding.doc_shard(rank, world, D)
rng = np.random.default_rng(17)
pr = rng.random((D, 16)) * 1e-4          # forw[3] rows, same on every rank
probs = np.full(16, 1.0 / 16)
eng.index_set_doc_base(dlo)
for tid in (capi.SS_TITLE, capi.SS_BODY):
    t = synth.index_table(V, D, tid, doc_lo=dlo, doc_hi=dhi, with_positions=True, n_threads=4)
    eng.index_load(tid, dhi - dlo, t.term_ptr, t.doc_ids - np.uint32(dlo), t.norm_tf, t.pos_ptr, t.pos)
    eng.term_weights(tid, float(D), t.n_postings, dhi - dlo, df_global=t.df_global, want=False)
eng.set_pagerank(pr[dlo:dhi])
q = synth.queries(Q, V, phrase_fraction=0.25, seed=44)
merged = eng.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=K, sharded=True)
# the round-1 route (per-shard lists gathered by the harness, ss_merge_topk on rank 0) must agree
local_res = eng.score_batch(q.kw_ptr, q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=K)
docs, finals, prs, counts = sharding.gather_result_lists(*local_res, device="cuda")
ok_sc = True
ft = synth.index_table(V, D, 0, with_positions=True, n_threads=4)
fb = synth.index_table(V, D, 1, with_positions=True, n_threads=4)
wt, mt = O.term_weights(ft.term_ptr, ft.doc_ids, ft.norm_tf, D, float(D))
wb, mb = O.term_weights(fb.term_ptr, fb.doc_ids, fb.norm_tf, D, float(D))
exp = O.score_batch(O.Table(ft.term_ptr, ft.doc_ids, wt, ft.pos_ptr, ft.pos),
                    O.Table(fb.term_ptr, fb.doc_ids, wb, fb.pos_ptr, fb.pos), D, mt, mb, pr, q.kw_ptr,
                    q.kw_terms, q.ph_ptr, q.ph_terms, topic_probs=probs, k=K, n_threads=4)
ok_sc = (np.array_equal(merged[0], exp[0]) and np.array_equal(merged[3], exp[3]) and
         np.allclose(merged[1], exp[1], rtol=1e-6, atol=0) and np.allclose(merged[2], exp[2], rtol=1e-6, atol=0))
if rank == 0:
    m2 = eng.merge_topk(docs, finals, prs, counts)
    ok_sc = ok_sc and all(np.array_equal(a, b) for a, b in zip(m2, merged))
    print("sharded scoring: merged top-k", "identical" if ok_sc else "DIFFERS", "shard_merge_ms",
          eng.score_stats().shard_merge_ms, flush=True)
flag = torch.tensor([int(ok_pr and ok_sc)], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
eng.close()
dist.destroy_process_group()
if rank == 0:
    print("MGPU_CHECK", "OK" if int(flag.item()) == 1 else "FAILED", flush=True)
sys.exit(0 if int(flag.item()) == 1 else 1)
