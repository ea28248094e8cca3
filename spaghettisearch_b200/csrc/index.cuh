// Device-resident inverted index shared by the offline weighting (index.cu)
// and the online scoring (score.cu).
#pragma once
#include "common.cuh"

struct TableState {
  bool loaded = false;
  uint64_t V = 0, P = 0;
  ss::DevBuf<uint64_t> term_ptr;  // [V + 1]
  ss::DevBuf<uint32_t> doc_ids;   // [P] ascending within a term
  ss::DevBuf<float> w;            // [P] listPos[0]: normTF, then tf-idf after ss_term_weights
  bool has_pos = false;
  ss::DevBuf<uint64_t> pos_ptr;   // [P + 1]
  ss::DevBuf<float> pos;          // listPos[1:]
  bool has_mag = false;
  ss::DevBuf<double> mag;         // [D] doc norms (forw[4])
  void clear() {
    loaded = has_pos = has_mag = false;
    V = P = 0;
    term_ptr.reset();
    doc_ids.reset();
    w.reset();
    pos_ptr.reset();
    pos.reset();
    mag.reset();
  }
};

struct IndexState {
  uint64_t D = 0;  // doc id space
  TableState tab[2];
  // forw[3] rows for the blend
  uint32_t T = 0;
  uint64_t pr_docs = 0;
  ss::DevBuf<double> pr;          // [pr_docs][T]
  // sqd[d] = sum_t probs[t] * pr[d][t] cached for the last shared topic vector
  ss::DevBuf<double> sqd;
  std::vector<double> sqd_probs;
  bool sqd_valid = false;
  // packed fp32 screening record per doc (score.cu), rebuilt when norms / blend inputs change
  ss::DevBuf<float4> meta32;
  bool meta32_valid = false;
  int meta32_mode = -1;
  // all weights finite and >= 0 (checked lazily by score.cu; enables the fp32 screened path)
  bool wcheck_valid = false;
  bool weights_nonneg = false;
  ss_score_stats stats{};
  // grow-only device workspace of ss_score_batch (cudaMalloc/cudaFree per batch would
  // synchronise the device and dominate small batches)
  struct Workspace {
    ss::DevBuf<uint64_t> kw_ptr, ph_ptr;
    ss::DevBuf<uint32_t> kw, ph, part_doc, part_count, out_doc, out_count, narrow;
    ss::DevBuf<double> probs, part_final, part_pr, out_final, out_pr, zero_mag;
    ss::DevBuf<unsigned long long> stats, qthr;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
    ~Workspace() {
      for (auto& e : ev)
        if (e) cudaEventDestroy(e);
    }
  } ws;
};

template <class T>
inline int ws_reserve(ss::DevBuf<T>& b, size_t n) {
  if (b.p && b.n >= n) return SS_OK;
  return b.alloc(n + n / 4);
}

IndexState* index_state(ss_engine* e);  // creates on first use; nullptr on OOM
