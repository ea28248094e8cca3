"""ctypes binding of libss_synth.so: deterministic synthetic link graphs,
inverted indexes and query batches.  Returns numpy arrays."""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional

import numpy as np

from . import _build

_lib = None


class _CIndex(C.Structure):
    _fields_ = [("n_terms", C.c_uint64), ("n_docs", C.c_uint64), ("n_postings", C.c_uint64),
                ("term_ptr", C.POINTER(C.c_uint64)), ("doc_ids", C.POINTER(C.c_uint32)),
                ("norm_tf", C.POINTER(C.c_float)), ("pos_ptr", C.POINTER(C.c_uint64)),
                ("pos", C.POINTER(C.c_float)), ("df_global", C.POINTER(C.c_uint64))]


def lib():
    global _lib
    if _lib is None:
        path = _build.SYNTH_LIB
        if not path.exists():
            _build.build_synth()
        _lib = C.CDLL(str(path))
        _lib.ss_synth_graph.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p,
                                        C.POINTER(C.POINTER(C.c_uint32)), C.POINTER(C.c_uint64)]
        _lib.ss_synth_graph.restype = C.c_int
        _lib.ss_synth_graph_rows.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_int, C.c_uint64,
                                             C.c_uint64, C.c_void_p, C.POINTER(C.POINTER(C.c_uint32)),
                                             C.POINTER(C.c_uint64)]
        _lib.ss_synth_graph_rows.restype = C.c_int
        _lib.ss_synth_topics.argtypes = [C.c_uint32, C.c_void_p]
        _lib.ss_synth_index_make.argtypes = [C.c_uint64, C.c_uint64, C.c_int, C.c_double, C.c_uint64,
                                             C.c_uint64, C.c_int, C.c_uint64, C.c_int, C.POINTER(_CIndex)]
        _lib.ss_synth_index_make.restype = C.c_int
        _lib.ss_synth_index_free.argtypes = [C.POINTER(_CIndex)]
        _lib.ss_synth_queries.argtypes = [C.c_uint64, C.c_uint64, C.c_double, C.c_uint64, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.c_void_p]
        _lib.ss_synth_queries.restype = C.c_int
        _lib.ss_synth_free.argtypes = [C.c_void_p]
    return _lib


def _copy(ptr, n, dtype):
    if n == 0:
        return np.zeros(0, dtype=dtype)
    return np.ctypeslib.as_array(ptr, shape=(int(n),)).astype(dtype, copy=True)


@dataclass
class Graph:
    """forw[2] on dense ids: children of u = col_idx[row_ptr[u]:row_ptr[u+1]]."""
    n_nodes: int
    row_ptr: np.ndarray  # uint64 [N+1]
    col_idx: np.ndarray  # uint32 [E]

    @property
    def n_edges(self) -> int:
        return int(self.row_ptr[-1])


def graph(n_nodes: int, target_edges: int, seed: int = 42, n_threads: int = 0) -> Graph:
    L = lib()
    row_ptr = np.zeros(n_nodes + 1, dtype=np.uint64)
    col = C.POINTER(C.c_uint32)()
    ne = C.c_uint64(0)
    rc = L.ss_synth_graph(n_nodes, target_edges, seed, n_threads, row_ptr.ctypes.data, C.byref(col),
                          C.byref(ne))
    if rc != 0:
        raise RuntimeError(f"ss_synth_graph failed: {rc}")
    try:
        col_idx = _copy(col, ne.value, np.uint32)
    finally:
        L.ss_synth_free(col)
    return Graph(n_nodes, row_ptr, col_idx)


def graph_rows(n_nodes: int, target_edges: int, u_lo: int, u_hi: int, seed: int = 42,
               n_threads: int = 0) -> Graph:
    """Rows [u_lo, u_hi) of graph(n_nodes, target_edges, seed); row_ptr is local to the slice."""
    L = lib()
    row_ptr = np.zeros(u_hi - u_lo + 1, dtype=np.uint64)
    col = C.POINTER(C.c_uint32)()
    ne = C.c_uint64(0)
    rc = L.ss_synth_graph_rows(n_nodes, target_edges, seed, n_threads, u_lo, u_hi, row_ptr.ctypes.data,
                               C.byref(col), C.byref(ne))
    if rc != 0:
        raise RuntimeError(f"ss_synth_graph_rows failed: {rc}")
    try:
        col_idx = _copy(col, ne.value, np.uint32)
    finally:
        L.ss_synth_free(col)
    return Graph(u_hi - u_lo, row_ptr, col_idx)


def topics(n_topics: int = 16) -> np.ndarray:
    out = np.zeros(n_topics, dtype=np.int64)
    lib().ss_synth_topics(n_topics, out.ctypes.data)
    return out


@dataclass
class IndexTable:
    """inv[0] (title) or inv[1] (body), term major, on dense ids."""
    table: int
    n_terms: int
    n_docs: int
    term_ptr: np.ndarray            # uint64 [V+1]
    doc_ids: np.ndarray             # uint32 [P]
    norm_tf: np.ndarray             # float32 [P]
    pos_ptr: Optional[np.ndarray]   # uint64 [P+1]
    pos: Optional[np.ndarray]       # float32
    df_global: np.ndarray           # uint64 [V]

    @property
    def n_postings(self) -> int:
        return int(self.term_ptr[-1])


def index_table(n_terms: int, n_docs: int, table: int, postings_per_doc: float = 0.0,
                doc_lo: int = 0, doc_hi: Optional[int] = None, with_positions: bool = False,
                seed: int = 43, n_threads: int = 0) -> IndexTable:
    L = lib()
    ci = _CIndex()
    hi = n_docs if doc_hi is None else doc_hi
    rc = L.ss_synth_index_make(n_terms, n_docs, table, postings_per_doc, doc_lo, hi,
                               1 if with_positions else 0, seed, n_threads, C.byref(ci))
    if rc != 0:
        L.ss_synth_index_free(C.byref(ci))
        raise RuntimeError(f"ss_synth_index_make failed: {rc}")
    try:
        P = ci.n_postings
        term_ptr = _copy(ci.term_ptr, n_terms + 1, np.uint64)
        doc_ids = _copy(ci.doc_ids, P, np.uint32)
        norm_tf = _copy(ci.norm_tf, P, np.float32)
        df = _copy(ci.df_global, n_terms, np.uint64)
        pos_ptr = pos = None
        if with_positions:
            pos_ptr = _copy(ci.pos_ptr, P + 1, np.uint64)
            pos = _copy(ci.pos, int(pos_ptr[-1]), np.float32)
    finally:
        L.ss_synth_index_free(C.byref(ci))
    return IndexTable(table, n_terms, n_docs, term_ptr, doc_ids, norm_tf, pos_ptr, pos, df)


@dataclass
class QueryBatch:
    kw_ptr: np.ndarray    # uint64 [Q+1]
    kw_terms: np.ndarray  # uint32
    ph_ptr: np.ndarray    # uint64 [Q+1]
    ph_terms: np.ndarray  # uint32

    @property
    def n_queries(self) -> int:
        return len(self.kw_ptr) - 1


def queries(n_queries: int, n_terms: int, phrase_fraction: float = 0.0, seed: int = 44) -> QueryBatch:
    kw_ptr = np.zeros(n_queries + 1, dtype=np.uint64)
    ph_ptr = np.zeros(n_queries + 1, dtype=np.uint64)
    kw = np.zeros(max(1, 5 * n_queries), dtype=np.uint32)
    ph = np.zeros(max(1, 3 * n_queries), dtype=np.uint32)
    rc = lib().ss_synth_queries(n_queries, n_terms, phrase_fraction, seed, kw_ptr.ctypes.data,
                                kw.ctypes.data, ph_ptr.ctypes.data, ph.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"ss_synth_queries failed: {rc}")
    return QueryBatch(kw_ptr, kw[: int(kw_ptr[-1])].copy(), ph_ptr, ph[: int(ph_ptr[-1])].copy())
