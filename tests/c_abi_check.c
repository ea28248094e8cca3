/* The boundary is a C ABI: this file is compiled as C99 (what cgo does with the header), references every
 * entry point, and on a box without a GPU checks that ss_create fails loudly instead of falling back. */
#include <stdio.h>
#include <string.h>

#include "spaghetti.h"

int main(void) {
  typedef void (*fn_t)(void);
  fn_t fns[] = {(fn_t)ss_version, (fn_t)ss_create, (fn_t)ss_destroy, (fn_t)ss_last_error, (fn_t)ss_stream_handle,
                 (fn_t)ss_comm_unique_id, (fn_t)ss_comm_init, (fn_t)ss_graph_load_csr, (fn_t)ss_graph_load_csr_rows,
                 (fn_t)ss_pagerank, (fn_t)ss_pagerank_set_teleport, (fn_t)ss_pagerank_fetch,
                 (fn_t)ss_pagerank_get_stats, (fn_t)ss_index_load, (fn_t)ss_index_set_doc_base, (fn_t)ss_index_clear,
                 (fn_t)ss_term_weights, (fn_t)ss_set_doc_norms, (fn_t)ss_set_pagerank, (fn_t)ss_use_pagerank,
                 (fn_t)ss_topics_load, (fn_t)ss_topic_probs, (fn_t)ss_score_batch, (fn_t)ss_score_batch_sharded,
                 (fn_t)ss_merge_topk, (fn_t)ss_score_get_stats};
  size_t i, n = sizeof(fns) / sizeof(fns[0]);
  for (i = 0; i < n; ++i)
    if (!fns[i]) return 2;
  if (ss_version() <= 0) return 3;
  {
    ss_config cfg;
    ss_engine* e = NULL;
    int rc;
    memset(&cfg, 0, sizeof(cfg));
    rc = ss_create(&cfg, &e);
    if (rc == SS_OK) { /* a GPU is present: fine, just clean up */
      ss_destroy(e);
      printf("engine created (GPU present), %u entry points\n", (unsigned)n);
      return 0;
    }
    if (rc != SS_ERR_NO_DEVICE || e != NULL || strstr(ss_last_error(), "no CPU path") == NULL) return 4;
    /* error paths that need no device */
    if (ss_graph_load_csr(NULL, 0, 0, NULL, NULL) != SS_ERR_INVALID) return 5;
    if (ss_score_batch(NULL, 0, NULL, NULL, NULL, NULL, NULL, 0, 10, NULL, NULL, NULL, NULL) != SS_ERR_INVALID) return 6;
  }
  printf("no device: ss_create -> SS_ERR_NO_DEVICE, %u entry points\n", (unsigned)n);
  return 0;
}
