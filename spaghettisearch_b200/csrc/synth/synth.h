// Deterministic synthetic workloads for the two hot paths.
//
// One generator feeds the CPU oracle, the CPU baseline and the GPU engine, so
// every arm of a parity test or a bench run sees bit-identical inputs.  All
// sampling is counter based -- value = mix(seed, index) -- so any slice
// (a row range of the graph, a doc range of the index) can be produced
// independently by a rank without generating the rest.
//
// Shapes follow the reference's tables:
//   graph  = forw[2]  docHash -> [childHash...]      (ranking/pagerank.go:18-39)
//   index  = inv[0|1] wordHash -> {docHash: [normTF, pos...]}
//                                                   (indexer/indexer.go:350-408)
//   query  = token lists after parser.Laundry + md5  (retrieval/main_retrieve.go:25-36)
// with hashes replaced by dense ids (ascending-hash rank).
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

// ---- link graph -----------------------------------------------------------
// Out-degree: 20 % of nodes dangling; the rest truncated power law
// (alpha 2.1, max 1e4) rescaled so that the de-duplicated edge count lands on
// target_edges.  Children: rank-Zipf(1.0) through a fixed bijection of node
// ids, de-duplicated per parent, self-loops allowed, ascending within a row.
//
// row_ptr is caller allocated [n_nodes+1]; *col_idx_out is malloc'ed by the
// generator (free with ss_synth_free).  Returns 0 on success.
int ss_synth_graph(uint64_t n_nodes, uint64_t target_edges, uint64_t seed,
                   int n_threads, uint64_t* row_ptr, uint32_t** col_idx_out,
                   uint64_t* n_edges_out);

// Rows [u_lo, u_hi) of the same graph (row_ptr is [u_hi-u_lo+1], offsets local
// to the slice); slices of one (n_nodes, target_edges, seed) concatenate to the
// full graph, so ranks can generate disjoint slices independently.
int ss_synth_graph_rows(uint64_t n_nodes, uint64_t target_edges, uint64_t seed, int n_threads,
                        uint64_t u_lo, uint64_t u_hi, uint64_t* row_ptr, uint32_t** col_idx_out,
                        uint64_t* n_edges_out);

// numPages[t] = 50000 + 12345*t  (forw[5] "numPages", crawler/ODP-scraper.go:104-107)
void ss_synth_topics(uint32_t n_topics, int64_t* num_pages);

// ---- inverted index -------------------------------------------------------
// Term r (1-based rank) has df(r) = min(D/2, C/r) postings; C is solved so the
// table holds about postings_per_doc * D postings.  A term's docs are one per
// stratum of [0, D), so they are distinct and ascending with O(1) work each,
// and the slice that falls in [doc_lo, doc_hi) is computable on its own.
typedef struct ss_synth_index {
  uint64_t n_terms;
  uint64_t n_docs;       // global D (ids in doc_ids are global)
  uint64_t n_postings;   // postings in [doc_lo, doc_hi)
  uint64_t* term_ptr;    // [n_terms+1]
  uint32_t* doc_ids;     // [n_postings] ascending within a term
  float* norm_tf;        // [n_postings] count/maxFreq in (0,1]
  uint64_t* pos_ptr;     // [n_postings+1] or NULL
  float* pos;            // positions as f32 (parser/parser.go:195-207), -100 sentinel
  uint64_t* df_global;   // [n_terms] df over the whole doc space
} ss_synth_index;

// table: 0 = title (about 8 postings/doc), 1 = body (about 100 postings/doc)
// unless postings_per_doc > 0 overrides it.
int ss_synth_index_make(uint64_t n_terms, uint64_t n_docs, int table,
                        double postings_per_doc, uint64_t doc_lo, uint64_t doc_hi,
                        int with_positions, uint64_t seed, int n_threads,
                        ss_synth_index* out);
void ss_synth_index_free(ss_synth_index* idx);

// ---- queries ---------------------------------------------------------------
// Lengths 1..5 with P = {.25,.35,.2,.12,.08}; terms rank-Zipf(0.8) over V.
// phrase_fraction of the queries also carry one 2-3 token phrase made of
// neighbouring term ids (so that positional matches exist).
// kw_ptr/ph_ptr are caller allocated [n_queries+1]; kw_terms/ph_terms are
// caller allocated with capacity 5*n_queries / 3*n_queries.
int ss_synth_queries(uint64_t n_queries, uint64_t n_terms, double phrase_fraction,
                     uint64_t seed, uint64_t* kw_ptr, uint32_t* kw_terms,
                     uint64_t* ph_ptr, uint32_t* ph_terms);

void ss_synth_free(void* p);

#ifdef __cplusplus
}
#endif
