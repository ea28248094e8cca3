/* libspaghetti_gpu -- C ABI of the B200 (sm_100a) ranking engine that replaces
 * SpaghettiSearch's two data-parallel hot paths behind the reference's Go API.
 *
 * A cgo shim keeps these Go signatures unchanged and calls in
 * here after exporting the Badger tables to dense-id CSR/CSC arrays:
 *   ranking.UpdateTopicSensitivePagerank   ranking/pagerank.go:14
 *   ranking.UpdateTermWeights              ranking/term_weighting.go:10
 *   retrieval.Retrieve                     retrieval/main_retrieve.go:15
 *
 * Conventions
 *  - Every pointer argument is caller-owned HOST memory, valid only for the
 *    duration of the call (cgo rule: C must not retain Go pointers).  Loads
 *    copy to the device before returning; outputs are caller allocated.
 *  - Every entry point returns an ss_status; nothing aborts, exits or throws.
 *    ss_last_error() gives the message of the calling thread's last failure;
 *    the Go shim panics with it, which is the reference's error convention
 *    (ranking/pagerank.go:20,29,49; retrieval/get_metadata.go:33,47).
 *  - There is NO CPU fallback: ss_create fails without an sm_100 device.
 *  - Dense ids: node/doc id = rank of the 32-hex md5 key in ascending order,
 *    term id likewise, so "ties by docID" is well defined.
 */
#ifndef SPAGHETTI_H_
#define SPAGHETTI_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define SS_API
#else
#define SS_API __attribute__((visibility("default")))
#endif

typedef struct ss_engine ss_engine;

typedef enum ss_status {
  SS_OK = 0,
  SS_NOT_CONVERGED = 1,     /* status, not an error: max_iters reached first */
  SS_ERR_INVALID = -1,      /* bad argument */
  SS_ERR_OOM = -2,          /* device or host allocation failed */
  SS_ERR_CUDA = -3,         /* CUDA runtime error (message has the detail) */
  SS_ERR_NCCL = -4,         /* NCCL error or libnccl not loadable */
  SS_ERR_STATE = -5,        /* call order violated, e.g. pagerank before load */
  SS_ERR_NO_DEVICE = -6     /* no sm_100 device: there is no CPU path */
} ss_status;

typedef enum ss_table { SS_TITLE = 0, SS_BODY = 1 } ss_table; /* inv[0], inv[1] */

enum {
  SS_FLAG_TIMING = 1u << 0 /* record CUDA-event timings of the hot kernels */
};

typedef struct ss_config {
  int32_t device;   /* CUDA device ordinal */
  uint32_t flags;   /* SS_FLAG_* */
  uint32_t reserved[6];
} ss_config;

SS_API int ss_version(void);
SS_API int ss_create(const ss_config* cfg, ss_engine** out);
SS_API void ss_destroy(ss_engine* e);
SS_API const char* ss_last_error(void);
/* The engine's CUDA stream (cudaStream_t) so that a harness can bracket calls
 * with its own events; all of the engine's kernels and copies run on it. */
SS_API void* ss_stream_handle(ss_engine* e);

/* ---- multi-GPU (one engine per GPU) -----------------------
 * ss_comm_unique_id: 128-byte NCCL id created by rank 0 and passed to every
 * rank out of band.  After ss_comm_init the engine is rank `rank` of `world`:
 * ss_graph_load_csr keeps the in-edges of this rank's destination rows
 * (edge-balanced 1-D partition) and ss_pagerank exchanges rank blocks over
 * NVLink once per sweep; ss_index_load callers pass only their doc shard and
 * merge local top-k lists with ss_merge_topk. */
SS_API int ss_comm_unique_id(void* id128);
SS_API int ss_comm_init(ss_engine* e, const void* id128, int32_t rank, int32_t world);

/* ---- HP-1: ranking/pagerank.go:14-145 --------------------------------------
 * Graph = forw[2] on dense ids, out-edge CSR: children of u are
 * col_idx[row_ptr[u] .. row_ptr[u+1]); node set = parents U children
 * (pagerank.go:24-44) = [0, n_nodes). */
SS_API int ss_graph_load_csr(ss_engine* e, uint64_t n_nodes, uint64_t n_edges,
                             const uint64_t* row_ptr, const uint32_t* col_idx);

/* Sharded export for a multi-GPU engine group: every rank of the ss_comm_init group passes the slice
 * [row_lo, row_hi) of the same CSR that it exported (row_ptr [row_hi - row_lo + 1] local to the slice,
 * starting at 0; col_idx holds global child ids).  The slices tile [0, n_nodes); for the best gather order
 * they should ascend with the rank.  The ranks exchange edges once over NVLink so that each keeps exactly
 * the in-edges of the rows it owns: host-to-device bytes and sort work per rank are 1/world of
 * ss_graph_load_csr, which every rank would have to call with the whole graph.  Collective: every rank of
 * the group must call it.  With no communicator it is ss_graph_load_csr. */
SS_API int ss_graph_load_csr_rows(ss_engine* e, uint64_t n_nodes, uint64_t row_lo, uint64_t row_hi,
                                  const uint64_t* row_ptr, const uint32_t* col_idx);

/* One power-iteration run per topic (pagerank.go:54-63), all topics advanced
 * together as columns of one SpMM; topic t starts from 1/num_pages[t]
 * (pagerank.go:61,104-105) and stops on its own when its L1 change <= eps
 * (pagerank.go:93).  max_iters = 0 is unbounded like the reference, except
 * that a run whose rank vector stopped changing bit-for-bit is ended.
 * out_rank: [n_nodes][n_topics] row major (forw[3] values), may be NULL to
 * leave the result on the device (ss_pagerank_fetch / ss_use_pagerank).
 * out_iters: [n_topics] sweeps executed, may be NULL.
 * Returns SS_OK, or SS_NOT_CONVERGED if some topic hit max_iters. */
SS_API int ss_pagerank(ss_engine* e, double damping, double eps, uint32_t n_topics,
                       const int64_t* num_pages, uint32_t max_iters, double* out_rank,
                       uint32_t* out_iters);
/* Extension beyond the reference as shipped (README.md:9 promises it, ranking/pagerank.go:90,117
 * teleports uniformly): per-topic teleport vectors.  weight is [n_nodes][n_topics] row major with
 * weight[v][t] = n_nodes * v_t[v] for a teleport distribution v_t over the nodes (sum_v v_t[v] = 1), so that
 * all ones is the reference's uniform teleport: cur[v] = (inherited + (1-d) * weight[v][t]) / Tot, Tot unchanged.
 * Applies to the following ss_pagerank calls with the same n_topics until the next graph load; NULL resets.
 * Call after the graph load (a rank keeps the weights of the rows it owns). */
SS_API int ss_pagerank_set_teleport(ss_engine* e, uint64_t n_nodes, uint32_t n_topics, const double* weight);
/* Copy the last result, rows [row_lo, row_hi), to host. */
SS_API int ss_pagerank_fetch(ss_engine* e, uint64_t row_lo, uint64_t row_hi, double* out_rank);

typedef struct ss_pagerank_stats {
  uint64_t n_nodes, n_edges;   /* global */
  uint64_t row_lo;             /* first destination row owned by this rank */
  uint64_t local_rows, local_edges; /* this rank's partition */
  uint32_t sweeps;             /* sweeps of the last ss_pagerank */
  uint32_t launches;           /* kernels launched by the last ss_pagerank */
  double sweep_ms_total;       /* device time in sweep kernels (SS_FLAG_TIMING) */
  double gather_ms_total;      /* device time of the two gather kernels only */
  double exchange_ms_total;    /* device time the sweep loop waits for the NVLink exchange after its own
                                  kernels (the exposed part; the rest overlaps the sweep) */
  double load_ms;              /* last ss_graph_load_csr, host wall clock */
  double short_ms_total;       /* device time of the short-row kernel only (SS_FLAG_TIMING) */
  double exchange_busy_ms_total; /* exchange stream: first chunk broadcast .. end of the all-reduce, per sweep summed */
} ss_pagerank_stats;
SS_API int ss_pagerank_get_stats(ss_engine* e, ss_pagerank_stats* out);

/* ---- HP-2 offline: ranking/term_weighting.go:10-123 ------------------------
 * One inverted table (inv[0] title / inv[1] body), term major: postings of
 * term t are [term_ptr[t], term_ptr[t+1]), doc_ids ascending within a term,
 * norm_tf = listPos[0] (indexer/indexer.go:362-363), positions = listPos[1:]
 * as f32 (pos_ptr NULL => table without positions; phrases then never match).
 * n_docs is the size of the doc id space (>= max doc id + 1). */
SS_API int ss_index_load(ss_engine* e, int table, uint64_t n_terms, uint64_t n_docs,
                         const uint64_t* term_ptr, const uint32_t* doc_ids,
                         const float* norm_tf, const uint64_t* pos_ptr, const float* pos);

/* Doc-sharded index: a shard loads its docs under shard-LOCAL ids 0 .. n_docs-1
 * (doc_ids, norms and PageRank rows all local) and declares the global id of its local doc 0 here;
 * ss_score_batch adds it to every doc id it returns.  Work per shard then depends on the shard's own
 * size only.  Default 0. */
SS_API int ss_index_set_doc_base(ss_engine* e, uint64_t doc_base);

/* Drop both tables, their norms, the blend input and the topic table (before loading another index). */
SS_API int ss_index_clear(ss_engine* e);

/* idf = float32(log2(total_docs / df)) with Go's Log2, w = norm_tf * idf in
 * fp32, mag[doc] = sqrt(sum float64(float32(w*w))), ascending term order.
 * df_global: NULL => df = this table's row length; else [n_terms] (doc-sharded
 * index: global df).  The table keeps the weights on the device for scoring.
 * out_w [P] and out_mag [n_docs] may be NULL.  Not idempotent, like the
 * reference (term_weighting.go:42-47): each call multiplies the stored
 * weights again. */
SS_API int ss_term_weights(ss_engine* e, int table, double total_docs, const uint64_t* df_global,
                           float* out_w, double* out_mag);
/* Load already-weighted postings' doc norms (forw[4]) instead of computing them. */
SS_API int ss_set_doc_norms(ss_engine* e, int table, uint64_t n_docs, const double* mag);

/* ---- HP-2 online: retrieval.Retrieve score/blend/top-k core -----------------
 * forw[3] rows for the blend: rank is [n_docs][n_topics]; NULL clears it. */
SS_API int ss_set_pagerank(ss_engine* e, uint64_t n_docs, uint32_t n_topics, const double* rank);
/* Use the device-resident result of the last ss_pagerank (node id == doc id). */
SS_API int ss_use_pagerank(ss_engine* e);

/* Extension beyond the reference as shipped: live topic probabilities, i.e. the dead
 * computeTopicProbs (retrieval/main_retrieve.go:106-159) with its defects repaired (see csrc/topics.cu).
 * ss_topics_load: inv[2] as CSR over its own dense term ids -- row of term w = (topic id, frequency) pairs --
 * and forw[5]'s wordCount per topic.  ss_topic_probs: for each query's keyword tokens (ids in inv[2]'s term
 * space, unknown word = id >= n_terms) the naive-Bayes value prod_i(freq_i[t] / wordCount[t]) / n_topics over
 * the tokens that list topic t, 0 if none does; out_probs [n_q][n_topics] is what ss_score_batch takes as
 * topic_probs with probs_per_query = 1. */
SS_API int ss_topics_load(ss_engine* e, uint64_t n_terms, uint32_t n_topics, const uint64_t* term_ptr,
                          const uint32_t* topic_ids, const double* freq, const double* word_count);
SS_API int ss_topic_probs(ss_engine* e, uint64_t n_q, const uint64_t* tok_ptr, const uint32_t* tok_terms,
                          double* out_probs);

/* Query q: keyword tokens kw_terms[kw_ptr[q] .. kw_ptr[q+1]) (duplicates kept,
 * main_retrieve.go:61-69) and one concatenated phrase ph_terms[ph_ptr[q] ..)
 * (main_retrieve.go:26; ph_ptr NULL => no phrases).  A term id >= n_terms is
 * an unknown term => empty postings (main_retrieve.go:193,218).
 * topic_probs: NULL => sqd = 0 as shipped (main_retrieve.go:87-88); else
 * [n_topics] (probs_per_query = 0) or [n_q][n_topics] (probs_per_query = 1).
 * FinalRank = (0.33*sqd + 0.38*Title + 0.29*Body) * 100, get_metadata.go:53-69.
 * Order: FinalRank descending, ties by ascending doc id, NaN last
 * (util.go:48-54 with the arrival-order tie pinned).  k <= 128; the reference
 * uses 50 (main_retrieve.go:99).
 * Outputs [n_q][k]: out_doc (0xFFFFFFFF in unused slots), out_final, out_pr
 * (Rank_combined.FinalRank / .PageRank, util.go:25-36); out_count [n_q]. */
SS_API int ss_score_batch(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr,
                          const uint32_t* kw_terms, const uint64_t* ph_ptr,
                          const uint32_t* ph_terms, const double* topic_probs,
                          int32_t probs_per_query, uint32_t k, uint32_t* out_doc,
                          double* out_final, double* out_pr, uint32_t* out_count);

/* The same call on a doc-sharded index (retrieval/main_retrieve.go:15-104 over all shards): every rank of
 * the ss_comm_init group passes the SAME query batch, scores it against its shard, all-gathers the local
 * top-k lists over NCCL and merges them on its own stream with the same comparator, so every rank
 * returns the global result (doc ids global via ss_index_set_doc_base).  With no communicator it is
 * ss_score_batch.  world * k <= 16384. */
SS_API int ss_score_batch_sharded(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr,
                                  const uint32_t* kw_terms, const uint64_t* ph_ptr,
                                  const uint32_t* ph_terms, const double* topic_probs,
                                  int32_t probs_per_query, uint32_t k, uint32_t* out_doc,
                                  double* out_final, double* out_pr, uint32_t* out_count);

/* Merge `n_lists` per-shard results of the same query batch ([n_q][k] each,
 * concatenated shard-major) into one [n_q][k] with the same comparator.  Each
 * input list must be in ss_score_batch's output order (best first): this is a
 * k-way merge.  k <= 128 and n_lists * k <= 16384. */
SS_API int ss_merge_topk(ss_engine* e, uint32_t n_lists, uint64_t n_q, uint32_t k,
                         const uint32_t* docs, const double* finals, const double* prs,
                         const uint32_t* counts, uint32_t* out_doc, double* out_final,
                         double* out_pr, uint32_t* out_count);

typedef struct ss_score_stats {
  uint64_t postings_scanned;   /* sum over queries of the postings of their lists (nominal: the impact-vector
                                  path covers a dense term's list with 2 bytes per doc instead of reading it) */
  uint64_t docs_matched;       /* sum over queries of matched docs (phrase queries on the impact-vector path:
                                  docs holding a posting of any query token, an upper bound) */
  uint64_t algorithmic_bytes;  /* SURVEY.md §8(d) B_q summed over the batch */
  uint32_t launches;           /* kernels launched by the last ss_score_batch */
  double kernel_ms;            /* device time of the last batch (SS_FLAG_TIMING) */
  double score_kernel_ms;      /* device time of the dominant scoring kernel */
  double shard_merge_ms;       /* device time from the end of the local merge to the end of the result copies
                                  (ss_score_batch_sharded: NCCL all-gather + cross-shard merge + D2H) */
  uint64_t model_bytes;        /* bytes the batch has to move on the path it takes: a query with a dense keyword
                                  streams 2 B per doc per dense token and reads 8 B per posting
                                  of its other tokens; other queries read 8 B per posting; + 12 B per result */
} ss_score_stats;
SS_API int ss_score_get_stats(ss_engine* e, ss_score_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* SPAGHETTI_H_ */
