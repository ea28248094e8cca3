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
  std::vector<uint32_t> df_host;  // [V] postings per term, host copy (byte model of ss_score_stats)
  void clear() {
    loaded = has_pos = has_mag = false;
    df_host.clear();
    V = P = 0;
    term_ptr.reset();
    doc_ids.reset();
    w.reset();
    pos_ptr.reset();
    pos.reset();
    mag.reset();
  }
};

// inv[2] (word -> {topic: frequency}) and forw[5]'s wordCount column, for ss_topic_probs (topics.cu)
struct TopicTable {
  bool loaded = false;
  uint64_t n_terms = 0;
  uint32_t T = 0;
  ss::DevBuf<uint64_t> term_ptr;
  ss::DevBuf<uint32_t> topic_ids;
  ss::DevBuf<double> freq, word_count;
};

struct IndexState {
  uint64_t D = 0;  // doc id space (shard-local ids 0 .. D-1)
  uint64_t doc_base = 0;  // global id of local doc 0 (doc-sharded index): added to the ids ss_score_batch returns
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
  // Impact vectors of the densest terms (score.cu): for each of the n_dense terms with the most
  // postings, one fp16 value per doc = an upper bound of 100 * (0.38 w_title / |title| + 0.29 w_body /
  // |body|) (0 = the doc has no posting of the term), plus zvec = upper bound of 33 * blend input.
  // Rebuilt when weights, norms or blend inputs change.
  bool dense_valid = false;     // uvec / dense_map match the current weights and norms
  bool zvec_valid = false;      // zvec matches meta32
  uint32_t n_dense = 0;
  uint64_t d_pad = 0;           // padded doc count of one vector
  ss::DevBuf<uint16_t> uvec;    // [n_dense][d_pad] fp16 bits
  ss::DevBuf<float2> ublk;      // [n_dense][d_pad / 4096] {largest impact, postings} per block of a vector
  ss::DevBuf<uint16_t> zvec;    // [d_pad]
  ss::DevBuf<float> zblk;       // [d_pad / 4096] largest blend term per doc block
  ss::DevBuf<uint8_t> dense_map;  // [V] dense slot of a term, 255 = none
  uint64_t dense_map_V = 0;
  std::vector<uint8_t> dense_host;  // [dense_map_V] 1 = the term has an impact vector (host copy for the byte model)
  TopicTable topics;
  ss_score_stats stats{};
  // grow-only device workspace of ss_score_batch (cudaMalloc/cudaFree per batch would
  // synchronise the device and dominate small batches)
  struct Workspace {
    ss::DevBuf<uint64_t> kw_ptr, ph_ptr;
    ss::DevBuf<uint32_t> kw, ph, part_doc, part_count, out_doc, out_count, narrow;
    ss::DevBuf<double> probs, part_final, part_pr, out_final, out_pr, zero_mag;
    ss::DevBuf<unsigned long long> stats, qthr, gtop;
    ss::DevBuf<uint32_t> glock;
    ss::DevBuf<uint8_t> group_len;
    // cross-shard gather of ss_score_batch_sharded: [world][n_q][k] lists and [world][n_q] counts
    ss::DevBuf<uint32_t> all_doc, all_count, loc_doc, loc_count;
    ss::DevBuf<double> all_final, all_pr, loc_final, loc_pr;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
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
