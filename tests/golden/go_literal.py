"""A second, independent restatement of the retrieval core: a LITERAL translation of the Go functions into
Python, statement by statement, on Go-shaped data (maps keyed by hash strings, []float32 posting values) -- the
oracle (oracle/oracle.cpp) works on dense CSR arrays instead.  Used by tests/test_oracle.py to cross-check the
oracle's phrase and ranking logic; test infrastructure only.

Translated functions (reference line ranges):
  getFromInverted        retrieval/main_retrieve.go:204-247
  getPosTerm             retrieval/phrase.go:122-170
  evalPhraseOccurrence   retrieval/phrase.go:53-109
  intersect/sortFloat32  retrieval/util.go:162-203
  genAggrDocsPipeline    retrieval/main_retrieve.go:170-187
  computeFinalRank       retrieval/get_metadata.go:39-69 (arithmetic only)
  appendSort + [:50]     retrieval/util.go:48-54, retrieval/main_retrieve.go:94-103
Go's random map iteration is replaced by ascending key order where the order is observable (the same pinning
the oracle documents): fan-in of keyword terms in query order, final docs in ascending hash order.
"""
import math

import numpy as np

f32 = np.float32


def sort_float32(sl):
    if len(sl) == 0:
        return None  # util.go:163-165: returns nil
    as64 = sorted(float(x) for x in sl)
    return [f32(x) for x in as64]


def intersect(s1, s2):
    if s1 is None or s2 is None:
        return None
    ret = None
    s1 = sort_float32(s1)
    s2 = sort_float32(s2)
    if s1 is None:
        s1 = []
    if s2 is None:
        s2 = []
    i = j = 0
    while i != len(s1) and j != len(s2):
        if s1[i] == s2[j]:
            ret = (ret or []) + [s1[i]]
            i += 1
            j += 1
        elif s1[i] > s2[j]:
            j += 1
        else:
            i += 1
    return ret


def go_len(sl):
    return 0 if sl is None else len(sl)


def get_from_inverted(term, inv):
    """-> {docHash: {"T": [w] | None, "B": [w] | None}}"""
    body = inv[1].get(term)   # nil map when the key is missing (ErrKeyNotFound tolerated)
    title = inv[0].get(term)
    ret = {}
    for doc, list_pos in (body or {}).items():
        ret[doc] = {"T": None, "B": [f32(list_pos[0])]}
    for doc, list_pos in (title or {}).items():
        v = ret.get(doc, {"T": None, "B": None})
        v["T"] = [f32(list_pos[0])]
        ret[doc] = v
    return ret


def get_pos_term(term, pos, inv):
    body = inv[1].get(term)
    title = inv[0].get(term)
    ret = {}
    for doc, list_pos in (body or {}).items():
        lp = [f32(x) for x in list_pos]          # every Get decodes a fresh copy
        for i in range(1, len(lp)):
            lp[i] = f32(lp[i] - f32(pos))
        ret[doc] = {"T": None, "B": lp, "pos": pos}
    for doc, list_pos in (title or {}).items():
        lp = [f32(x) for x in list_pos]
        for i in range(1, len(lp)):
            lp[i] = f32(lp[i] - f32(pos))
        v = ret.get(doc, {"T": None, "B": None, "pos": 0})
        v["T"] = lp
        v["pos"] = pos
        ret[doc] = v
    return ret


def eval_phrase_occurrence(agg, length_phrase):
    ret = {}
    for doc in sorted(agg):
        tw = agg[doc]
        sum_b, sum_t = f32(0), f32(0)
        b_int = t_int = None
        if len(tw) != length_phrase:
            b_int = t_int = None
        else:
            if go_len(tw[0]["B"]) != 0:
                sum_b = f32(sum_b + tw[0]["B"][0])
                b_int = tw[0]["B"][1:]
            if go_len(tw[0]["T"]) != 0:
                sum_t = f32(sum_t + tw[0]["T"][0])
                t_int = tw[0]["T"][1:]
            for idx in range(1, len(tw)):
                i = idx & 0xFF                      # uint8(idx)
                e = tw.get(i, {"T": None, "B": None})   # a missing map key reads as the zero Rank_term
                if go_len(e["B"]) == 0:
                    b_int = None
                else:
                    sum_b = f32(sum_b + e["B"][0])
                    b_int = intersect(b_int, e["B"][1:])
                if go_len(e["T"]) == 0:
                    t_int = None
                else:
                    sum_t = f32(sum_t + e["T"][0])
                    t_int = intersect(t_int, e["T"][1:])
        if go_len(b_int) != 0 or go_len(t_int) != 0:
            v = ret.get(doc, {"T": None, "B": None})
            if go_len(b_int) != 0:
                v["B"] = (v["B"] or []) + [sum_b]
            if go_len(t_int) != 0:
                v["T"] = (v["T"] or []) + [sum_t]
            ret[doc] = v
    return ret


def get_phrase_from_inverted(phrase, inv):
    agg = {}
    for pos, term in enumerate(phrase):
        tp = pos & 0xFF                             # termPhrase.Pos is uint8 (phrase.go:115)
        for doc, ranks in get_pos_term(term, tp, inv).items():
            val_ = agg.get(doc)
            if val_ is None:
                val_ = {}
            val = val_.get(ranks["pos"], {"T": None, "B": None})
            val = {"T": ranks["T"], "B": ranks["B"]}
            val_[ranks["pos"]] = val
            agg[doc] = val_
    return eval_phrase_occurrence(agg, len(phrase))


def retrieve(query_tokens, phrase_tokens, inv, mag, pagerank=None, topic_probs=None, limit=50):
    """inv = [title, body] with {term: {doc: [w, pos...]}}; mag = {doc: {"title": m, "body": m}};
    pagerank = {doc: {topic: rank}}; topic_probs = {topic: p} or None (nil map as shipped).
    -> [(doc, FinalRank, PageRank)], at most `limit`."""
    doc_phrase = get_phrase_from_inverted(phrase_tokens, inv)
    aggregated = {}
    for term in query_tokens:                       # fan-in pinned to query order
        for doc, ranks in get_from_inverted(term, inv).items():
            val = aggregated.get(doc, {"T": None, "B": None})
            val = {"T": (val["T"] or []) + (ranks["T"] or []), "B": (val["B"] or []) + (ranks["B"] or [])}
            aggregated[doc] = val
    for doc, ranks in doc_phrase.items():
        val = aggregated.get(doc, {"T": None, "B": None})
        val = {"T": (val["T"] or []) + (ranks["T"] or []), "B": (val["B"] or []) + (ranks["B"] or [])}
        aggregated[doc] = val
    query_length = len(query_tokens) + len(phrase_tokens)
    final = []
    for doc in sorted(aggregated):                  # arrival order pinned to ascending hash
        rank = aggregated[doc]
        title_rank = 0.0
        for w in rank["T"] or []:
            title_rank += float(w)
        body_rank = 0.0
        for w in rank["B"] or []:
            body_rank += float(w)
        sqd = 0.0
        for topic in sorted(topic_probs or {}):
            sqd += topic_probs[topic] * pagerank[doc].get(topic, 0.0)
        page_mag = mag[doc]                         # the reference panics if the doc is missing
        qm = math.sqrt(float(query_length))
        with np.errstate(divide="ignore", invalid="ignore"):
            body_rank = float(np.float64(body_rank) / np.float64(page_mag.get("body", 0.0) * qm))
            title_rank = float(np.float64(title_rank) / np.float64(page_mag.get("title", 0.0) * qm))
        if math.isnan(body_rank):
            body_rank = 0.0
        if math.isnan(title_rank):
            title_rank = 0.0
        fr = (0.33 * sqd + 0.38 * title_rank + 0.29 * body_rank) * 100.0
        # appendSort: first index whose FinalRank < the new one; equal scores keep arrival order
        lo, hi = 0, len(final)
        while lo < hi:
            mid = (lo + hi) // 2
            if not (final[mid][1] < fr):
                lo = mid + 1
            else:
                hi = mid
        final.insert(lo, (doc, fr, sqd))
    return final[:limit]
