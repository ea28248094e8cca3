"""ctypes loader for liboracle.so -- TEST INFRASTRUCTURE ONLY.

Importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs; nothing under spaghettisearch_b200/ may import it.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path
from typing import Optional

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "liboracle.so"
_lib = None


class _Table(C.Structure):
    _fields_ = [("n_terms", C.c_uint64), ("term_ptr", C.c_void_p), ("doc_ids", C.c_void_p),
                ("w", C.c_void_p), ("pos_ptr", C.c_void_p), ("pos", C.c_void_p)]


def build(force: bool = False):
    """Compile the checker with oracle/Makefile (building it is not using it)."""
    src_m = max((HERE / "oracle.cpp").stat().st_mtime, (HERE / "oracle.h").stat().st_mtime)
    if force or not LIB.exists() or LIB.stat().st_mtime < src_m:
        subprocess.run(["make", "-s", "-C", str(HERE)] + (["-B"] if force else []) + ["liboracle.so"],
                       check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB))
        L.oracle_go_log.argtypes = [C.c_double]
        L.oracle_go_log.restype = C.c_double
        L.oracle_go_log2.argtypes = [C.c_double]
        L.oracle_go_log2.restype = C.c_double
        L.oracle_pagerank.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                      C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
        L.oracle_pagerank.restype = C.c_int
        L.oracle_pagerank_faithful.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_double,
                                               C.c_double, C.c_int64, C.c_uint32, C.c_void_p,
                                               C.POINTER(C.c_uint32), C.POINTER(C.c_double)]
        L.oracle_pagerank_faithful.restype = C.c_int
        L.oracle_pagerank_fair.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_double,
                                           C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int,
                                           C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_pagerank_fair.restype = C.c_int
        L.oracle_csc_build.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_csc_build.restype = C.c_int
        L.oracle_pagerank_fair_csc.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double,
                                               C.c_double, C.c_uint32, C.c_void_p, C.c_uint32, C.c_uint32,
                                               C.c_int, C.c_void_p, C.c_void_p, C.POINTER(C.c_double)]
        L.oracle_pagerank_fair_csc.restype = C.c_int
        L.oracle_pagerank_biased.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_double, C.c_double, C.c_uint32,
                                             C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.oracle_pagerank_biased.restype = C.c_int
        L.oracle_term_weights.argtypes = [C.c_uint64, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p,
                                          C.c_double, C.c_void_p, C.c_void_p]
        L.oracle_term_weights.restype = C.c_int
        L.oracle_score_batch.argtypes = [C.POINTER(_Table), C.POINTER(_Table), C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p,
                                         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int,
                                         C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_int]
        L.oracle_score_batch.restype = C.c_int
        L.oracle_score_batch_fair.argtypes = L.oracle_score_batch.argtypes
        L.oracle_score_batch_fair.restype = C.c_int
        L.oracle_intersect.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_void_p]
        L.oracle_intersect.restype = C.c_uint64
        _lib = L
    return _lib


def _p(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def go_log2(x: float) -> float:
    return lib().oracle_go_log2(float(x))


def go_log(x: float) -> float:
    return lib().oracle_go_log(float(x))


def pagerank(row_ptr, col_idx, damping, eps, num_pages, max_iters=0):
    """ranking/pagerank.go:14-145 -> (rank [N][T] float64, iters [T])."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    num_pages = np.ascontiguousarray(num_pages, dtype=np.int64)
    n, t = len(row_ptr) - 1, len(num_pages)
    rank = np.zeros((n, t), dtype=np.float64)
    iters = np.zeros(t, dtype=np.uint32)
    rc = lib().oracle_pagerank(n, _p(row_ptr), _p(col_idx), damping, eps, t, _p(num_pages), max_iters,
                               _p(rank), _p(iters))
    assert rc == 0
    return rank, iters


def pagerank_faithful(row_ptr, col_idx, damping, eps, num_pages_t, fixed_iters=0, want_rank=True):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    n = len(row_ptr) - 1
    rank = np.zeros(n, dtype=np.float64) if want_rank else None
    iters = C.c_uint32(0)
    secs = C.c_double(0)
    rc = lib().oracle_pagerank_faithful(n, _p(row_ptr), _p(col_idx), damping, eps, int(num_pages_t),
                                        fixed_iters, _p(rank), C.byref(iters), C.byref(secs))
    assert rc == 0
    return rank, iters.value, secs.value


def pagerank_fair(row_ptr, col_idx, damping, eps, num_pages, max_iters=0, fixed_iters=0, n_threads=0,
                  want_rank=True):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    num_pages = np.ascontiguousarray(num_pages, dtype=np.int64)
    n, t = len(row_ptr) - 1, len(num_pages)
    rank = np.zeros((n, t), dtype=np.float64) if want_rank else None
    iters = np.zeros(t, dtype=np.uint32)
    secs = C.c_double(0)
    rc = lib().oracle_pagerank_fair(n, _p(row_ptr), _p(col_idx), damping, eps, t, _p(num_pages),
                                    max_iters, fixed_iters, n_threads, _p(rank), _p(iters),
                                    C.byref(secs))
    assert rc == 0
    return rank, iters, secs.value


def csc_build(row_ptr, col_idx):
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    n = len(row_ptr) - 1
    in_ptr = np.zeros(n + 1, dtype=np.uint64)
    in_src = np.zeros(max(1, len(col_idx)), dtype=np.uint32)
    assert lib().oracle_csc_build(n, _p(row_ptr), _p(col_idx), _p(in_ptr), _p(in_src)) == 0
    return in_ptr, in_src


def pagerank_biased(row_ptr, col_idx, damping, eps, num_pages, tele_w, max_iters=0, n_threads=0):
    """Extension: per-topic teleport weights tele_w [N][T] = N * v_t[v] -> (rank, iters)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    col_idx = np.ascontiguousarray(col_idx, dtype=np.uint32)
    num_pages = np.ascontiguousarray(num_pages, dtype=np.int64)
    tele_w = np.ascontiguousarray(tele_w, dtype=np.float64)
    n, t = len(row_ptr) - 1, len(num_pages)
    assert tele_w.shape == (n, t)
    rank = np.zeros((n, t), dtype=np.float64)
    iters = np.zeros(t, dtype=np.uint32)
    rc = lib().oracle_pagerank_biased(n, _p(row_ptr), _p(col_idx), damping, eps, t, _p(num_pages), max_iters,
                                      n_threads, _p(tele_w), _p(rank), _p(iters))
    if rc != 0:
        raise RuntimeError(f"oracle_pagerank_biased failed: {rc}")
    return rank, iters


def topic_probs(term_ptr, topic_ids, freq, word_count, tok_ptr, tok_terms):
    """Extension: the repaired computeTopicProbs -> [Q][T]."""
    term_ptr = np.ascontiguousarray(term_ptr, dtype=np.uint64)
    topic_ids = np.ascontiguousarray(topic_ids, dtype=np.uint32)
    freq = np.ascontiguousarray(freq, dtype=np.float64)
    word_count = np.ascontiguousarray(word_count, dtype=np.float64)
    tok_ptr = np.ascontiguousarray(tok_ptr, dtype=np.uint64)
    tok_terms = np.ascontiguousarray(tok_terms, dtype=np.uint32)
    nq, t = len(tok_ptr) - 1, len(word_count)
    out = np.zeros((nq, t), dtype=np.float64)
    L = lib()
    L.oracle_topic_probs.argtypes = [C.c_uint64, C.c_uint32] + [C.c_void_p] * 4 + [C.c_uint64] + [C.c_void_p] * 3
    L.oracle_topic_probs.restype = C.c_int
    L.oracle_topic_probs(len(term_ptr) - 1, t, _p(term_ptr), _p(topic_ids), _p(freq), _p(word_count), nq, _p(tok_ptr),
                         _p(tok_terms), _p(out))
    return out


def pagerank_fair_csc(row_ptr, in_ptr, in_src, damping, eps, num_pages, max_iters=0, fixed_iters=0,
                      n_threads=0, want_rank=True):
    """Fair CPU arm on a prebuilt in-edge CSC -> (rank or None, iters, seconds in the sweeps)."""
    row_ptr = np.ascontiguousarray(row_ptr, dtype=np.uint64)
    num_pages = np.ascontiguousarray(num_pages, dtype=np.int64)
    n, t = len(row_ptr) - 1, len(num_pages)
    rank = np.zeros((n, t), dtype=np.float64) if want_rank else None
    iters = np.zeros(t, dtype=np.uint32)
    secs = C.c_double(0)
    rc = lib().oracle_pagerank_fair_csc(n, _p(row_ptr), _p(in_ptr), _p(in_src), damping, eps, t,
                                        _p(num_pages), max_iters, fixed_iters, n_threads, _p(rank),
                                        _p(iters), C.byref(secs))
    assert rc == 0
    return rank, iters, secs.value


def term_weights(term_ptr, doc_ids, norm_tf, n_docs, total_docs):
    """ranking/term_weighting.go:10-123 -> (w float32 [P], mag float64 [D])."""
    term_ptr = np.ascontiguousarray(term_ptr, dtype=np.uint64)
    doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
    norm_tf = np.ascontiguousarray(norm_tf, dtype=np.float32)
    w = np.zeros(len(doc_ids), dtype=np.float32)
    mag = np.zeros(n_docs, dtype=np.float64)
    rc = lib().oracle_term_weights(len(term_ptr) - 1, n_docs, _p(term_ptr), _p(doc_ids), _p(norm_tf),
                                   float(total_docs), _p(w), _p(mag))
    assert rc == 0
    return w, mag


class Table:
    """Weighted postings of one inverted table, kept alive for the C struct."""

    def __init__(self, term_ptr, doc_ids, w, pos_ptr=None, pos=None):
        self.term_ptr = np.ascontiguousarray(term_ptr, dtype=np.uint64)
        self.doc_ids = np.ascontiguousarray(doc_ids, dtype=np.uint32)
        self.w = np.ascontiguousarray(w, dtype=np.float32)
        self.pos_ptr = None if pos_ptr is None else np.ascontiguousarray(pos_ptr, dtype=np.uint64)
        self.pos = None if pos is None else np.ascontiguousarray(pos, dtype=np.float32)
        self.c = _Table(len(self.term_ptr) - 1, _p(self.term_ptr), _p(self.doc_ids), _p(self.w),
                        _p(self.pos_ptr), _p(self.pos))


def score_batch(title: Table, body: Table, n_docs, mag_title, mag_body, pagerank_m, kw_ptr, kw_terms,
                ph_ptr=None, ph_terms=None, topic_probs=None, k=50, n_threads=0, fair=False):
    """retrieval.Retrieve's score/blend/top-k core -> (doc [Q][k], final, pr, count [Q]).
    fair=True: the same results from the CPU-friendly flavour (dense accumulators, all cores)."""
    mag_title = np.ascontiguousarray(mag_title, dtype=np.float64)
    mag_body = np.ascontiguousarray(mag_body, dtype=np.float64)
    kw_ptr = np.ascontiguousarray(kw_ptr, dtype=np.uint64)
    kw_terms = np.ascontiguousarray(kw_terms, dtype=np.uint32)
    nq = len(kw_ptr) - 1
    if ph_ptr is not None:
        ph_ptr = np.ascontiguousarray(ph_ptr, dtype=np.uint64)
        ph_terms = np.ascontiguousarray(ph_terms, dtype=np.uint32)
    n_topics, per_q = 0, 0
    if pagerank_m is not None:
        pagerank_m = np.ascontiguousarray(pagerank_m, dtype=np.float64)
        n_topics = pagerank_m.shape[1]
    if topic_probs is not None:
        topic_probs = np.ascontiguousarray(topic_probs, dtype=np.float64)
        per_q = 1 if topic_probs.ndim == 2 else 0
        n_topics = topic_probs.shape[-1]
    out_doc = np.zeros((nq, k), dtype=np.uint32)
    out_final = np.zeros((nq, k), dtype=np.float64)
    out_pr = np.zeros((nq, k), dtype=np.float64)
    out_count = np.zeros(nq, dtype=np.uint32)
    fn = lib().oracle_score_batch_fair if fair else lib().oracle_score_batch
    rc = fn(C.byref(title.c), C.byref(body.c), n_docs, _p(mag_title), _p(mag_body), _p(pagerank_m), n_topics, nq,
            _p(kw_ptr), _p(kw_terms), _p(ph_ptr), _p(ph_terms), _p(topic_probs), per_q, k, _p(out_doc),
            _p(out_final), _p(out_pr), _p(out_count), n_threads)
    assert rc == 0
    return out_doc, out_final, out_pr, out_count


def intersect(a, b):
    """retrieval/util.go:179-203; None models a nil slice."""
    aa = None if a is None else np.array(a, dtype=np.float32)
    bb = None if b is None else np.array(b, dtype=np.float32)
    out = np.zeros(max(1, min(len(a) if a is not None else 0, len(b) if b is not None else 0)),
                   dtype=np.float32)
    n = lib().oracle_intersect(_p(aa), 0 if aa is None else len(aa), _p(bb),
                               0 if bb is None else len(bb), _p(out))
    return out[:n].tolist()
