"""Where the time of the benchmark's keyword batch goes, by query class (dev tool, one GPU):
the batch is split by (#dense tokens, #sparse tokens, postings) and every class is scored as its own batch."""
import sys
import numpy as np
sys.path.insert(0, ".")
from spaghettisearch_b200 import capi, synth
Q = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
D, V = 10_000_000, 1_000_000
title = synth.index_table(V, D, 0); body = synth.index_table(V, D, 1)
e = capi.Engine(timing=True)
e.index_load(0, D, title.term_ptr, title.doc_ids, title.norm_tf); e.index_load(1, D, body.term_ptr, body.doc_ids, body.norm_tf)
e.term_weights(0, float(D), title.n_postings, D, want=False); e.term_weights(1, float(D), body.n_postings, D, want=False)
rng = np.random.default_rng(7); e.set_pagerank((rng.random((D, 16)) + 0.5) / D); probs = np.full(16, 1 / 16)
df = np.diff(title.term_ptr.astype(np.int64)) + np.diff(body.term_ptr.astype(np.int64))
order = np.argsort(-df, kind="stable")
dense = np.zeros(V, bool); cand = order[df[order] >= D // 32][:224]; dense[cand] = True
q = synth.queries(Q, V, seed=44)
kp, kt = q.kw_ptr.astype(np.int64), q.kw_terms
classes = {}
for i in range(Q):
    t = kt[kp[i]:kp[i + 1]]
    nd = int(dense[t].sum()); ns = len(t) - nd
    post = int(df[t[~dense[t]]].sum())
    if nd == 0:
        c = "sparse<=3K" if post <= 3072 else ("sparse<=100K" if post <= 100000 else "sparse>100K")
    elif nd == 1:
        c = "1 dense" if ns == 0 else ("1 dense + sparse<=100K" if post <= 100000 else "1 dense + sparse>100K")
    else:
        c = "2+ dense" if ns == 0 else "2+ dense + sparse"
    classes.setdefault(c, []).append(i)
def run(idx):
    ptr = np.zeros(len(idx) + 1, np.uint64); toks = []
    for j, i in enumerate(idx):
        toks.append(kt[kp[i]:kp[i + 1]]); ptr[j + 1] = ptr[j] + (kp[i + 1] - kp[i])
    kw = np.concatenate(toks).astype(np.uint32)
    for _ in range(2):
        e.score_batch(ptr, kw, topic_probs=probs, k=10); s = e.score_stats()
    return s
import os
configs = [("block maxima off, slab-kth bound", {"SS_SCORE_BLOCKMAX": "0", "SS_SCORE_GTOP": "0"}),
           ("block maxima on,  slab-kth bound", {"SS_SCORE_BLOCKMAX": "1", "SS_SCORE_GTOP": "0"}),
           ("block maxima on,  global top-k bound (default)", {})]
ref = None
for name, env in configs:
    os.environ.update(env)
    print("==", name, flush=True)
    ptr0 = np.asarray(q.kw_ptr, np.uint64)
    for _ in range(2):
        res = e.score_batch(ptr0, kt, topic_probs=probs, k=10); s_all = e.score_stats()
    if ref is None:
        ref = [r.copy() for r in res]
    else:
        assert all(np.array_equal(a, b) for a, b in zip(ref, res)), "results differ between configurations"
    print(f"whole batch: {s_all.kernel_ms:.1f} ms ({Q / s_all.kernel_ms * 1e3:.0f} q/s), k_score {s_all.score_kernel_ms:.1f} ms", flush=True)
    tot = 0.0
    for c, idx in sorted(classes.items()):
        s = run(idx); tot += s.score_kernel_ms
        print(f"{c:28s} {len(idx):6d} queries  k_score {s.score_kernel_ms:8.2f} ms  {s.score_kernel_ms * 1e3 / len(idx):7.2f} us/query  "
              f"model GB {s.model_bytes / 1e9:8.1f}  -> {s.model_bytes / s.score_kernel_ms / 1e6:6.0f} GB/s", flush=True)
    print(f"sum of classes {tot:.1f} ms")
    for k_ in env:
        del os.environ[k_]
