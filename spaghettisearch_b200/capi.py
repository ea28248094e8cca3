"""ctypes binding of libspaghetti_gpu.so (include/spaghetti.h).

This is the only way Python reaches the engine, and it goes through exactly
the C ABI a cgo shim would bind.  There is no fallback: if the CUDA library
is missing or no sm_100 device exists, calls raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _build

SS_OK = 0
SS_NOT_CONVERGED = 1
SS_TITLE, SS_BODY = 0, 1
SS_FLAG_TIMING = 1

EXPORTS = [
    "ss_version", "ss_create", "ss_destroy", "ss_last_error", "ss_stream_handle", "ss_comm_unique_id", "ss_comm_init",
    "ss_graph_load_csr", "ss_graph_load_csr_rows", "ss_pagerank", "ss_pagerank_fetch", "ss_pagerank_set_teleport", "ss_pagerank_get_stats", "ss_index_load",
    "ss_index_clear", "ss_index_set_doc_base", "ss_score_batch_sharded",
    "ss_term_weights", "ss_set_doc_norms", "ss_set_pagerank", "ss_use_pagerank", "ss_score_batch",
    "ss_merge_topk", "ss_score_get_stats", "ss_topics_load", "ss_topic_probs",
]


class SSError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libspaghetti_gpu error {code}: {msg}")
        self.code = code


class Config(C.Structure):
    _fields_ = [("device", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_uint32 * 6)]


class PagerankStats(C.Structure):
    _fields_ = [("n_nodes", C.c_uint64), ("n_edges", C.c_uint64), ("row_lo", C.c_uint64), ("local_rows", C.c_uint64),
                ("local_edges", C.c_uint64), ("sweeps", C.c_uint32), ("launches", C.c_uint32),
                ("sweep_ms_total", C.c_double), ("gather_ms_total", C.c_double),
                ("exchange_ms_total", C.c_double), ("load_ms", C.c_double), ("short_ms_total", C.c_double),
                ("exchange_busy_ms_total", C.c_double)]


class ScoreStats(C.Structure):
    _fields_ = [("postings_scanned", C.c_uint64), ("docs_matched", C.c_uint64),
                ("algorithmic_bytes", C.c_uint64), ("launches", C.c_uint32), ("kernel_ms", C.c_double),
                ("score_kernel_ms", C.c_double), ("shard_merge_ms", C.c_double), ("model_bytes", C.c_uint64)]


_lib = None


def lib_path():
    return _build.GPU_LIB


def load():
    """dlopen the engine.  Raises if it was not built (no silent fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    path = lib_path()
    if not path.exists():
        raise FileNotFoundError(
            f"{path} is missing: build it with `python -m spaghettisearch_b200._build` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(str(path))
    vp, u64, u32, i32, dbl = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int32, C.c_double
    L.ss_version.restype = C.c_int
    L.ss_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.ss_destroy.argtypes = [vp]
    L.ss_destroy.restype = None
    L.ss_last_error.restype = C.c_char_p
    L.ss_stream_handle.argtypes = [vp]
    L.ss_stream_handle.restype = vp
    L.ss_comm_unique_id.argtypes = [vp]
    L.ss_comm_init.argtypes = [vp, vp, i32, i32]
    L.ss_graph_load_csr.argtypes = [vp, u64, u64, vp, vp]
    L.ss_graph_load_csr_rows.argtypes = [vp, u64, u64, u64, vp, vp]
    L.ss_pagerank.argtypes = [vp, dbl, dbl, u32, vp, u32, vp, vp]
    L.ss_pagerank_set_teleport.argtypes = [vp, u64, u32, vp]
    L.ss_pagerank_fetch.argtypes = [vp, u64, u64, vp]
    L.ss_pagerank_get_stats.argtypes = [vp, C.POINTER(PagerankStats)]
    L.ss_index_load.argtypes = [vp, C.c_int, u64, u64, vp, vp, vp, vp, vp]
    L.ss_index_clear.argtypes = [vp]
    L.ss_term_weights.argtypes = [vp, C.c_int, dbl, vp, vp, vp]
    L.ss_set_doc_norms.argtypes = [vp, C.c_int, u64, vp]
    L.ss_set_pagerank.argtypes = [vp, u64, u32, vp]
    L.ss_use_pagerank.argtypes = [vp]
    L.ss_score_batch.argtypes = [vp, u64, vp, vp, vp, vp, vp, i32, u32, vp, vp, vp, vp]
    L.ss_score_batch_sharded.argtypes = [vp, u64, vp, vp, vp, vp, vp, i32, u32, vp, vp, vp, vp]
    L.ss_index_set_doc_base.argtypes = [vp, u64]
    L.ss_topics_load.argtypes = [vp, u64, u32, vp, vp, vp, vp]
    L.ss_topic_probs.argtypes = [vp, u64, vp, vp, vp]
    L.ss_merge_topk.argtypes = [vp, u32, u64, u32, vp, vp, vp, vp, vp, vp, vp, vp]
    L.ss_score_get_stats.argtypes = [vp, C.POINTER(ScoreStats)]
    for name in EXPORTS:
        if name not in ("ss_destroy", "ss_last_error", "ss_stream_handle"):
            getattr(L, name).restype = C.c_int
    _lib = L
    return L


def _ptr(a):
    """Host pointer of a numpy array / torch CPU tensor (pinned or not); None -> NULL."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        assert a.flags["C_CONTIGUOUS"]
        return a.ctypes.data
    if hasattr(a, "data_ptr"):  # torch tensor
        assert not a.is_cuda and a.is_contiguous()
        return a.data_ptr()
    raise TypeError(type(a))


def _as(a, dtype):
    if a is None or hasattr(a, "data_ptr"):
        return a
    return np.ascontiguousarray(a, dtype=dtype)


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    rc = load().ss_comm_unique_id(buf)
    if rc < 0:
        raise SSError(rc, load().ss_last_error().decode())
    return buf.raw


class Engine:
    """One engine per GPU; mirrors the C ABI one to one."""

    def __init__(self, device: int = 0, timing: bool = False):
        self.L = load()
        self.h = C.c_void_p()
        cfg = Config(device=device, flags=SS_FLAG_TIMING if timing else 0)
        self._check(self.L.ss_create(C.byref(cfg), C.byref(self.h)))
        self.n_nodes = 0
        self.n_topics = 0

    def _check(self, rc: int) -> int:
        if rc < 0:
            raise SSError(rc, self.L.ss_last_error().decode())
        return rc

    def close(self):
        if self.h:
            self.L.ss_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def stream_handle(self) -> int:
        return int(self.L.ss_stream_handle(self.h) or 0)

    # ---- multi-GPU
    def comm_init(self, unique_id: bytes, rank: int, world: int):
        self._check(self.L.ss_comm_init(self.h, unique_id, rank, world))

    # ---- HP-1
    def graph_load_csr(self, row_ptr, col_idx):
        row_ptr, col_idx = _as(row_ptr, np.uint64), _as(col_idx, np.uint32)
        n = len(row_ptr) - 1
        self._check(self.L.ss_graph_load_csr(self.h, n, len(col_idx), _ptr(row_ptr), _ptr(col_idx)))
        self.n_nodes = n

    def graph_load_csr_rows(self, n_nodes, row_lo, row_hi, row_ptr, col_idx):
        """Sharded export: this rank's slice [row_lo, row_hi) of the CSR (row_ptr local to the slice)."""
        row_ptr, col_idx = _as(row_ptr, np.uint64), _as(col_idx, np.uint32)
        assert len(row_ptr) == row_hi - row_lo + 1
        self._check(self.L.ss_graph_load_csr_rows(self.h, n_nodes, row_lo, row_hi, _ptr(row_ptr), _ptr(col_idx)))

    def pagerank(self, damping, eps, num_pages, max_iters=0, out=None, want_rank=True):
        """-> (rank [N][T] or None, iters [T], status)."""
        num_pages = np.ascontiguousarray(num_pages, dtype=np.int64)
        t = len(num_pages)
        if out is None and want_rank:
            out = np.zeros((self.n_nodes, t), dtype=np.float64)
        iters = np.zeros(max(t, 1), dtype=np.uint32)
        rc = self._check(self.L.ss_pagerank(self.h, damping, eps, t, _ptr(num_pages), max_iters, _ptr(out),
                                            _ptr(iters)))
        self.n_topics = t
        return out, iters[:t], rc

    def pagerank_set_teleport(self, weight):
        """Topic-biased teleport (extension): weight [N][T] = N * v_t[v]; None resets to the reference's uniform."""
        if weight is None:
            self._check(self.L.ss_pagerank_set_teleport(self.h, 0, 0, None))
            return
        weight = _as(weight, np.float64)
        self._check(self.L.ss_pagerank_set_teleport(self.h, weight.shape[0], weight.shape[1], _ptr(weight)))

    def pagerank_fetch(self, row_lo, row_hi, out=None):
        if out is None:
            out = np.zeros((row_hi - row_lo, self.n_topics), dtype=np.float64)
        self._check(self.L.ss_pagerank_fetch(self.h, row_lo, row_hi, _ptr(out)))
        return out

    def pagerank_stats(self) -> PagerankStats:
        st = PagerankStats()
        self._check(self.L.ss_pagerank_get_stats(self.h, C.byref(st)))
        return st

    # ---- HP-2
    def index_load(self, table, n_docs, term_ptr, doc_ids, norm_tf, pos_ptr=None, pos=None):
        term_ptr, doc_ids, norm_tf = _as(term_ptr, np.uint64), _as(doc_ids, np.uint32), _as(norm_tf, np.float32)
        pos_ptr, pos = _as(pos_ptr, np.uint64), _as(pos, np.float32)
        self._check(self.L.ss_index_load(self.h, table, len(term_ptr) - 1, n_docs, _ptr(term_ptr),
                                         _ptr(doc_ids), _ptr(norm_tf), _ptr(pos_ptr), _ptr(pos)))

    def index_clear(self):
        self._check(self.L.ss_index_clear(self.h))

    def index_set_doc_base(self, doc_base: int):
        self._check(self.L.ss_index_set_doc_base(self.h, int(doc_base)))

    def term_weights(self, table, total_docs, n_postings, n_docs, df_global=None, want=True):
        df_global = _as(df_global, np.uint64)
        w = np.zeros(n_postings, dtype=np.float32) if want else None
        mag = np.zeros(n_docs, dtype=np.float64) if want else None
        self._check(self.L.ss_term_weights(self.h, table, float(total_docs), _ptr(df_global), _ptr(w), _ptr(mag)))
        return w, mag

    def set_doc_norms(self, table, mag):
        mag = _as(mag, np.float64)
        self._check(self.L.ss_set_doc_norms(self.h, table, len(mag), _ptr(mag)))

    def set_pagerank(self, rank: Optional[np.ndarray]):
        if rank is None:
            self._check(self.L.ss_set_pagerank(self.h, 0, 0, None))
            return
        rank = _as(rank, np.float64)
        self._check(self.L.ss_set_pagerank(self.h, rank.shape[0], rank.shape[1], _ptr(rank)))

    def use_pagerank(self):
        self._check(self.L.ss_use_pagerank(self.h))

    def score_batch(self, kw_ptr, kw_terms, ph_ptr=None, ph_terms=None, topic_probs=None, k=50, out=None,
                    sharded=False):
        """-> (doc [Q][k] uint32, final [Q][k], pr [Q][k], count [Q]).  sharded: ss_score_batch_sharded
        (every rank of the communicator calls it with the same batch and gets the global top-k)."""
        kw_ptr, kw_terms = _as(kw_ptr, np.uint64), _as(kw_terms, np.uint32)
        ph_ptr, ph_terms = _as(ph_ptr, np.uint64), _as(ph_terms, np.uint32)
        nq = len(kw_ptr) - 1
        per_q = 0
        if topic_probs is not None:
            topic_probs = _as(topic_probs, np.float64)
            per_q = 1 if topic_probs.ndim == 2 else 0
        if out is None:
            out = (np.zeros((nq, k), dtype=np.uint32), np.zeros((nq, k), dtype=np.float64),
                   np.zeros((nq, k), dtype=np.float64), np.zeros(nq, dtype=np.uint32))
        fn = self.L.ss_score_batch_sharded if sharded else self.L.ss_score_batch
        self._check(fn(self.h, nq, _ptr(kw_ptr), _ptr(kw_terms), _ptr(ph_ptr), _ptr(ph_terms),
                _ptr(topic_probs), per_q, k, _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3])))
        return out

    def topics_load(self, term_ptr, topic_ids, freq, word_count):
        """inv[2] as CSR (term -> (topic id, frequency)) + forw[5] wordCount per topic (extension, 8(f)-3)."""
        term_ptr, topic_ids = _as(term_ptr, np.uint64), _as(topic_ids, np.uint32)
        freq, word_count = _as(freq, np.float64), _as(word_count, np.float64)
        self._check(self.L.ss_topics_load(self.h, len(term_ptr) - 1, len(word_count), _ptr(term_ptr), _ptr(topic_ids),
                                          _ptr(freq), _ptr(word_count)))
        self._n_topics_nb = len(word_count)

    def topic_probs(self, tok_ptr, tok_terms):
        """-> [Q][T] naive-Bayes topic probabilities of the queries' keyword tokens."""
        tok_ptr, tok_terms = _as(tok_ptr, np.uint64), _as(tok_terms, np.uint32)
        out = np.zeros((len(tok_ptr) - 1, self._n_topics_nb), dtype=np.float64)
        self._check(self.L.ss_topic_probs(self.h, len(tok_ptr) - 1, _ptr(tok_ptr), _ptr(tok_terms), _ptr(out)))
        return out

    def merge_topk(self, docs, finals, prs, counts):
        """Inputs [n_lists][Q][k] / [n_lists][Q]."""
        docs, finals, prs = _as(docs, np.uint32), _as(finals, np.float64), _as(prs, np.float64)
        counts = _as(counts, np.uint32)
        n_lists, nq, k = docs.shape
        out = (np.zeros((nq, k), dtype=np.uint32), np.zeros((nq, k), dtype=np.float64),
               np.zeros((nq, k), dtype=np.float64), np.zeros(nq, dtype=np.uint32))
        self._check(self.L.ss_merge_topk(self.h, n_lists, nq, k, _ptr(docs), _ptr(finals), _ptr(prs), _ptr(counts),
                                         _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3])))
        return out

    def score_stats(self) -> ScoreStats:
        st = ScoreStats()
        self._check(self.L.ss_score_get_stats(self.h, C.byref(st)))
        return st
