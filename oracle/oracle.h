/* CPU oracle for the two SpaghettiSearch hot paths -- TEST INFRASTRUCTURE ONLY.
 *
 * A literal restatement, on dense ids, of the reference's Go arithmetic:
 *   ranking/pagerank.go:85-145, ranking/term_weighting.go:10-57,
 *   retrieval/main_retrieve.go:15-104,161-247, retrieval/get_metadata.go:16-77,
 *   retrieval/phrase.go:11-170, retrieval/util.go:48-54,162-203.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library; the product (libspaghetti_gpu)
 * never links or calls it.
 *
 * PARITY UNPINNED: the reference ships no golden vector, known-answer test or
 * fixture for these paths and no Go toolchain exists in
 * this image, so the oracle cannot be checked against the reference itself.
 * It is pinned instead against hand-derived known answers (tests/golden/) that
 * follow from the cited lines.
 *
 * Where Go's map iteration makes the reference's floating-point summation
 * order random, the oracle fixes ascending-id order (documented per function).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Go's math.Log (FDLIBM e_log.c port, go/src/math/log.go) and math.Log2
 * (go/src/math/log10.go: Frexp, exact for powers of two). */
double oracle_go_log(double x);
double oracle_go_log2(double x);

/* ---- HP-1: ranking/pagerank.go -------------------------------------------
 * Out-edge CSR of forw[2] on dense ids: node u's children are
 * col_idx[row_ptr[u] .. row_ptr[u+1]).  One power iteration run per topic,
 * sequentially (pagerank.go:54-63), n = num_pages[t].
 * out_rank is [n_nodes][n_topics] row major; out_iters[t] = sweeps executed.
 * Parents are visited in ascending id order (Go: random map order).
 * max_iters = 0 means unbounded, like the reference. */
int oracle_pagerank(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                    double damping, double eps, uint32_t n_topics, const int64_t* num_pages,
                    uint32_t max_iters, double* out_rank, uint32_t* out_iters);

/* Same arithmetic for ONE topic on string-keyed hash maps (32-hex keys), the
 * data structures the Go code pays for; used for the "faithful" CPU timing.
 * fixed_iters > 0 runs exactly that many sweeps and ignores eps. */
int oracle_pagerank_faithful(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                             double damping, double eps, int64_t num_pages, uint32_t fixed_iters,
                             double* out_rank /* [n_nodes] or NULL */, uint32_t* out_iters,
                             double* sweep_seconds /* total time inside the sweep loop */);

/* "Fair" CPU arm: same update as a dense-id pull over an in-edge CSC with
 * OpenMP over rows and all T topics per row.  fixed_iters > 0 runs exactly
 * that many sweeps for every topic.  sweep_seconds excludes the transpose. */
int oracle_pagerank_fair(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                         double damping, double eps, uint32_t n_topics, const int64_t* num_pages,
                         uint32_t max_iters, uint32_t fixed_iters, int n_threads,
                         double* out_rank, uint32_t* out_iters, double* sweep_seconds);

/* The transpose on its own (so that a timing loop pays for it once) and the fair
 * arm on a prebuilt CSC. in_ptr is [n_nodes+1], in_src is [n_edges]. */
int oracle_csc_build(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                     uint64_t* in_ptr, uint32_t* in_src);
/* Extension beyond the shipped reference (README.md:9): per-topic teleport vectors.
 * tele_w [N][T] = N * v_t[v]; all ones reproduces oracle_pagerank_fair. */
int oracle_pagerank_biased(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                           double damping, double eps, uint32_t n_topics, const int64_t* num_pages,
                           uint32_t max_iters, int n_threads, const double* tele_w, double* out_rank,
                           uint32_t* out_iters);
int oracle_pagerank_fair_csc(uint64_t n_nodes, const uint64_t* row_ptr, const uint64_t* in_ptr,
                             const uint32_t* in_src, double damping, double eps, uint32_t n_topics,
                             const int64_t* num_pages, uint32_t max_iters, uint32_t fixed_iters,
                             int n_threads, double* out_rank, uint32_t* out_iters,
                             double* sweep_seconds);

/* ---- HP-2 offline: ranking/term_weighting.go:10-57 + saveMagnitude ---------
 * Term-major postings.  idf = float32(Log2(total_docs / df)), df = postings of
 * the term in THIS table; w = normTF * idf in fp32; mag[doc] += float64(w*w)
 * with the square rounded to fp32 first; out_mag[doc] = sqrt(sum)
 * (term_weighting.go:59-123).  Terms visited in ascending id order. */
int oracle_term_weights(uint64_t n_terms, uint64_t n_docs, const uint64_t* term_ptr,
                        const uint32_t* doc_ids, const float* norm_tf, double total_docs,
                        float* out_w, double* out_mag);

/* ---- HP-2 online: retrieval.Retrieve score / blend / top-k core ------------ */
typedef struct oracle_table {
  uint64_t n_terms;
  const uint64_t* term_ptr; /* [n_terms+1] */
  const uint32_t* doc_ids;  /* ascending within a term */
  const float* w;           /* tf-idf weights (listPos[0]) */
  const uint64_t* pos_ptr;  /* [P+1] or NULL (then phrases never match) */
  const float* pos;         /* listPos[1:] */
} oracle_table;

/* For each query: keyword tokens kw_terms[kw_ptr[q]..kw_ptr[q+1]) (duplicates
 * kept, main_retrieve.go:61-69), one concatenated phrase ph_terms[...]
 * (main_retrieve.go:26; ph_ptr may be NULL).  Term ids >= n_terms are unknown
 * terms (ErrKeyNotFound => empty, main_retrieve.go:193,218).
 * topic_probs: NULL reproduces the shipped behaviour (nil map => sqd = 0,
 * main_retrieve.go:87-88); else [n_topics] shared or [n_q][n_topics].
 * Result order: FinalRank descending, ties by ascending doc id, NaN last.
 * Outputs are [n_q][k]; unused slots get doc 0xFFFFFFFF and score 0. */
int oracle_score_batch(const oracle_table* title, const oracle_table* body, uint64_t n_docs,
                       const double* mag_title, const double* mag_body, const double* pagerank,
                       uint32_t n_topics, uint64_t n_q, const uint64_t* kw_ptr,
                       const uint32_t* kw_terms, const uint64_t* ph_ptr, const uint32_t* ph_terms,
                       const double* topic_probs, int probs_per_query, uint32_t k,
                       uint32_t* out_doc, double* out_final, double* out_pr, uint32_t* out_count,
                       int n_threads);
/* same results, CPU-friendly data structures (dense accumulators, shared blend term, bounded selection) */
int oracle_score_batch_fair(const oracle_table* title, const oracle_table* body, uint64_t n_docs,
                       const double* mag_title, const double* mag_body, const double* pagerank,
                       uint32_t n_topics, uint64_t n_q, const uint64_t* kw_ptr,
                       const uint32_t* kw_terms, const uint64_t* ph_ptr, const uint32_t* ph_terms,
                       const double* topic_probs, int probs_per_query, uint32_t k,
                       uint32_t* out_doc, double* out_final, double* out_pr, uint32_t* out_count,
                       int n_threads);

/* retrieval/util.go:179-203 on its own (sorted multiset intersection). */
uint64_t oracle_intersect(float* a, uint64_t na, float* b, uint64_t nb, float* out);

/* Extension: computeTopicProbs (retrieval/main_retrieve.go:106-159) with probs starting at
 * 1 instead of 0; inv[2] as CSR over its own term ids. */
int oracle_topic_probs(uint64_t n_terms, uint32_t n_topics, const uint64_t* term_ptr,
                       const uint32_t* topic_ids, const double* freq, const double* word_count,
                       uint64_t n_q, const uint64_t* tok_ptr, const uint32_t* tok_terms, double* out_probs);

#ifdef __cplusplus
}
#endif
