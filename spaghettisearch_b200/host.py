"""ctypes binding of libspaghetti_host.so: the C++ mirror of the reference's Go API
(ranking.UpdateTopicSensitivePagerank, ranking.UpdateTermWeights, retrieval.Retrieve)
over JSON-lines snapshots of the Badger tables.  Used by tests and tools."""
from __future__ import annotations

import ctypes as C
import json
from pathlib import Path

from . import _build

_lib = None


class HostError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        path = _build.HOST_LIB
        if not path.exists():
            _build.build_host()
        L = C.CDLL(str(path))
        L.ssh_last_error.restype = C.c_char_p
        L.ssh_db_new.restype = C.c_void_p
        L.ssh_db_free.argtypes = [C.c_void_p]
        L.ssh_db_load_jsonl.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.ssh_db_save_jsonl.argtypes = [C.c_void_p, C.c_char_p, C.c_char_p]
        L.ssh_db_rows.argtypes = [C.c_void_p, C.c_char_p]
        L.ssh_db_rows.restype = C.c_longlong
        L.ssh_export_graph.argtypes = [C.c_void_p, C.c_char_p]
        L.ssh_update_pagerank.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_double]
        L.ssh_update_term_weights.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p]
        L.ssh_retrieve.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_size_t]
        L.ssh_retrieve_concurrent.argtypes = [C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_int, C.c_char_p,
                                              C.c_size_t, C.POINTER(C.c_ulonglong)]
        _lib = L
    return _lib


class DB:
    """database.DB_init's nine tables as in-memory snapshots (inv0..2, forw0..5)."""

    def __init__(self):
        self.L = lib()
        self.h = C.c_void_p(self.L.ssh_db_new())

    def _ok(self, rc):
        if rc != 0:
            raise HostError(self.L.ssh_last_error().decode())

    def close(self):
        if self.h:
            self.L.ssh_db_free(self.h)
            self.h = None

    def load(self, table: str, path):
        self._ok(self.L.ssh_db_load_jsonl(self.h, table.encode(), str(path).encode()))

    def save(self, table: str, path):
        self._ok(self.L.ssh_db_save_jsonl(self.h, table.encode(), str(path).encode()))

    def rows(self, table: str) -> int:
        return int(self.L.ssh_db_rows(self.h, table.encode()))

    def export_graph(self, path):
        self._ok(self.L.ssh_export_graph(self.h, str(path).encode()))

    def update_pagerank(self, engine, damping=0.75, eps=1e-20):
        """cmd/crawl/start_crawl.go:175 calls with d = 0.75, eps = 1e-20."""
        self._ok(self.L.ssh_update_pagerank(self.h, engine.h, damping, eps))

    def update_term_weights(self, engine, info: str):
        self._ok(self.L.ssh_update_term_weights(self.h, engine.h, info.encode()))

    def retrieve(self, engine, kw_hashes, ph_hashes=()):
        buf = C.create_string_buffer(1 << 16)
        self._ok(self.L.ssh_retrieve(self.h, engine.h, " ".join(kw_hashes).encode(), " ".join(ph_hashes).encode(), buf,
                                     len(buf)))
        return json.loads(buf.value.decode())


def _retrieve_concurrent(self, engine, queries, threads=8, window_us=200):
    """queries: list of (kw_hashes, ph_hashes); served through the C++ BatchingRetriever from `threads`
    concurrent callers -> (results per query, ss_score_batch calls made)."""
    text = "\n".join(" ".join(kw) + "|" + " ".join(ph) for kw, ph in queries)
    buf = C.create_string_buffer(1 << 24)
    stats = (C.c_ulonglong * 2)()
    self._ok(self.L.ssh_retrieve_concurrent(self.h, engine.h, text.encode(), threads, window_us, buf, len(buf), stats))
    lines = buf.value.decode().splitlines()
    return [json.loads(l) for l in lines], int(stats[0]), int(stats[1])


DB.retrieve_concurrent = _retrieve_concurrent


def write_jsonl(path, rows):
    """rows: iterable of (key, python value) -> {"k": key, "v": value} lines."""
    with open(path, "w") as f:
        for k, v in rows:
            f.write(json.dumps({"k": k, "v": v}) + "\n")


def read_jsonl(path):
    out = {}
    for line in Path(path).read_text().splitlines():
        if line.strip():
            r = json.loads(line)
            out[r["k"]] = r["v"]
    return out
