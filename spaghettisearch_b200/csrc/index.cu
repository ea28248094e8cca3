// HP-2 (placeholder until the index/scoring kernels land in the next commit).
#include "common.cuh"

struct IndexState {};
void index_state_free(IndexState* s) { delete s; }

#define SS_TODO(name) ss::set_error(name ": not implemented yet"); return SS_ERR_STATE

extern "C" {
SS_API int ss_index_load(ss_engine*, int, uint64_t, uint64_t, const uint64_t*, const uint32_t*, const float*,
                         const uint64_t*, const float*) { SS_TODO("ss_index_load"); }
SS_API int ss_term_weights(ss_engine*, int, double, const uint64_t*, float*, double*) { SS_TODO("ss_term_weights"); }
SS_API int ss_set_doc_norms(ss_engine*, int, uint64_t, const double*) { SS_TODO("ss_set_doc_norms"); }
SS_API int ss_set_pagerank(ss_engine*, uint64_t, uint32_t, const double*) { SS_TODO("ss_set_pagerank"); }
SS_API int ss_use_pagerank(ss_engine*) { SS_TODO("ss_use_pagerank"); }
SS_API int ss_score_batch(ss_engine*, uint64_t, const uint64_t*, const uint32_t*, const uint64_t*, const uint32_t*,
                          const double*, int32_t, uint32_t, uint32_t*, double*, double*, uint32_t*) { SS_TODO("ss_score_batch"); }
SS_API int ss_merge_topk(ss_engine*, uint32_t, uint64_t, uint32_t, const uint32_t*, const double*, const double*,
                         const uint32_t*, uint32_t*, double*, double*, uint32_t*) { SS_TODO("ss_merge_topk"); }
SS_API int ss_score_get_stats(ss_engine*, ss_score_stats*) { SS_TODO("ss_score_get_stats"); }
}
