"""Generates tests/golden/kats.json from a pure-Python literal restatement of
the reference lines (independent of oracle/ and of the CUDA engine).

    python tests/golden/make_kats.py

ranking/pagerank.go:85-145, ranking/term_weighting.go:29-50,
retrieval/get_metadata.go:53-69.
"""
import json
import math
import struct
from pathlib import Path


def f32(x):
    return struct.unpack("f", struct.pack("f", x))[0]


def pagerank(children, nodes, d, eps, n, max_iters=0):
    cur, last = {}, {}
    it, change = 1, float("inf")
    while change > eps:
        cur, last = last, cur
        if it > 1:
            for k in nodes:
                cur[k] = 0.0
        else:
            for k in nodes:
                cur[k] = 1.0 / n
                last[k] = 1.0 / n
        tot = 0.0
        for p in nodes:  # ascending id; Go iterates in random map order
            kids = children.get(p, [])
            if not kids:
                continue
            w = d * last[p] / len(kids)
            tot += w
            for c in kids:
                cur[c] += w
        tot += (1 - d) * len(cur)
        change = 0.0
        for k in nodes:
            cur[k] = (cur[k] + (1 - d)) / tot
            change += abs(cur[k] - last[k])
        if max_iters and it >= max_iters:
            it += 1
            break
        it += 1
    return [cur[k] for k in nodes], it - 1


def kat_pr1():
    # A,B,C,D = 0..3 ; forw[2] = {A:[B,C], B:[C], D:[A]}  (C dangling, D has no parent)
    children = {0: [1, 2], 1: [2], 3: [0]}
    nodes = [0, 1, 2, 3]
    out = {"row_ptr": [0, 2, 3, 3, 4], "col_idx": [1, 2, 2, 0], "damping": 0.75, "cases": []}
    for n, eps, mi in [(10, 1e-20, 1), (4, 1e-20, 0), (10, 1e-20, 0), (4, 1e-9, 0), (10, 1e-9, 0)]:
        r, it = pagerank(children, nodes, 0.75, eps, n, mi)
        out["cases"].append({"num_pages": n, "eps": eps, "max_iters": mi, "rank": r, "iters": it})
    return out


def kat_sc1():
    # totalDocs = 8; docs d1,d2,d3 = ids 1,2,3 (id 0 unused); terms t1,t2,t3 = ids 0,1,2
    total = 8.0
    body = {0: {1: 0.5, 2: 1.0}, 1: {1: 1.0, 3: 0.25}, 2: {3: 1.0}}
    title = {0: {2: 1.0}, 2: {1: 1.0, 3: 0.5}}

    def weigh(tab):
        w, mag = {}, {}
        for t in sorted(tab):
            idf = f32(math.log2(total / len(tab[t])))  # df 1,2 -> exact powers of two
            w[t] = {}
            for dd, tf in tab[t].items():
                x = f32(f32(tf) * idf)
                w[t][dd] = x
                mag[dd] = mag.get(dd, 0.0) + float(f32(x * x))
        return w, {dd: math.sqrt(m) for dd, m in mag.items()}

    wb, mb = weigh(body)
    wt, mt = weigh(title)
    query = [0, 1]
    qm = math.sqrt(len(query))
    agg = {}
    for t in query:
        for dd, x in wb.get(t, {}).items():
            agg.setdefault(dd, [0.0, 0.0])[1] += float(x)
        for dd, x in wt.get(t, {}).items():
            agg.setdefault(dd, [0.0, 0.0])[0] += float(x)
    res = []
    for dd, (tr, br) in agg.items():
        b = br / (mb.get(dd, 0.0) * qm) if mb.get(dd, 0.0) * qm != 0 or br != 0 else float("nan")
        t = tr / (mt.get(dd, 0.0) * qm) if mt.get(dd, 0.0) * qm != 0 or tr != 0 else float("nan")
        b = 0.0 if math.isnan(b) else b
        t = 0.0 if math.isnan(t) else t
        res.append({"doc": dd, "title": t, "body": b, "final": (0.33 * 0.0 + 0.38 * t + 0.29 * b) * 100.0})
    res.sort(key=lambda r: (-r["final"], r["doc"]))

    def csc(tab, n_terms):
        ptr, docs, tfs = [0], [], []
        for t in range(n_terms):
            for dd in sorted(tab.get(t, {})):
                docs.append(dd)
                tfs.append(tab[t][dd])
            ptr.append(len(docs))
        return {"term_ptr": ptr, "doc_ids": docs, "norm_tf": tfs}

    return {"total_docs": total, "n_docs": 4, "n_terms": 3, "body": csc(body, 3), "title": csc(title, 3),
            "w_body": [wb[t][dd] for t in range(3) for dd in sorted(body.get(t, {}))],
            "w_title": [wt[t][dd] for t in range(3) for dd in sorted(title.get(t, {}))],
            "mag_body": [mb.get(dd, 0.0) for dd in range(4)], "mag_title": [mt.get(dd, 0.0) for dd in range(4)],
            "query": query, "result": res}


def pagerank_biased(children, nodes, d, eps, n, w):
    """Extension: the same loop with the teleport term (1-d) * w[k] per node, Tot unchanged."""
    cur, last = {}, {}
    it, change = 1, float("inf")
    while change > eps:
        cur, last = last, cur
        for k in nodes:
            if it > 1:
                cur[k] = 0.0
            else:
                cur[k] = last[k] = 1.0 / n
        tot = 0.0
        for p in nodes:
            kids = children.get(p, [])
            if not kids:
                continue
            x = d * last[p] / len(kids)
            tot += x
            for c in kids:
                cur[c] += x
        tot += (1 - d) * len(cur)
        change = 0.0
        for k in nodes:
            cur[k] = (cur[k] + (1 - d) * w[k]) / tot
            change += abs(cur[k] - last[k])
        it += 1
    return [cur[k] for k in nodes], it - 1


def kat_pr2():
    # KAT-PR-1's graph, topic 0 teleports to {A, B} only, topic 1 uniformly (weights sum to N = 4 per topic)
    children = {0: [1, 2], 1: [2], 3: [0]}
    nodes = [0, 1, 2, 3]
    w0 = {0: 2.0, 1: 2.0, 2: 0.0, 3: 0.0}
    w1 = {k: 1.0 for k in nodes}
    r0, i0 = pagerank_biased(children, nodes, 0.75, 1e-12, 7, w0)
    r1, i1 = pagerank_biased(children, nodes, 0.75, 1e-12, 9, w1)
    u1, j1 = pagerank(children, nodes, 0.75, 1e-12, 9)
    assert r1 == u1 and i1 == j1  # all-ones weights are the reference's loop
    return {"row_ptr": [0, 2, 3, 3, 4], "col_idx": [1, 2, 2, 0], "damping": 0.75, "eps": 1e-12, "num_pages": [7, 9],
            "weights": [[w0[k], w1[k]] for k in nodes], "rank": [[r0[k], r1[k]] for k in nodes], "iters": [i0, i1]}


def kat_tp1():
    # Extension: computeTopicProbs (main_retrieve.go:106-159) with probs starting at 1.
    # inv[2]: word -> {topic: count}; forw[5] wordCount per topic; 3 topics, 4 words
    inv2 = {0: {0: 3, 2: 1}, 1: {1: 5}, 2: {0: 2, 1: 2, 2: 2}, 3: {}}
    word_count = [100.0, 50.0, 40.0]
    queries = [[0], [0, 2], [1, 3], [3], [], [2, 2], [0, 99]]
    T = 3
    out = []
    for q in queries:
        row = []
        for t in range(T):
            probs, any_ = 1.0, False
            for word in q:
                f = inv2.get(word, {}).get(t)
                if f is None:
                    continue
                probs *= (float(f) / word_count[t])
                any_ = True
            row.append(probs / float(T) if any_ else 0.0)
        out.append(row)
    ptr, ids, freq = [0], [], []
    for word in range(4):
        for t in sorted(inv2[word]):
            ids.append(t)
            freq.append(float(inv2[word][t]))
        ptr.append(len(ids))
    tok_ptr = [0]
    toks = []
    for q in queries:
        toks += q
        tok_ptr.append(len(toks))
    return {"term_ptr": ptr, "topic_ids": ids, "freq": freq, "word_count": word_count, "tok_ptr": tok_ptr,
            "tok_terms": toks, "probs": out}


if __name__ == "__main__":
    out = {"KAT-PR-1": kat_pr1(), "KAT-SC-1": kat_sc1(), "KAT-PR-2-biased": kat_pr2(), "KAT-TP-1": kat_tp1()}
    Path(__file__).with_name("kats.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out["KAT-SC-1"]["result"], indent=1))
