// live topic probabilities for the PageRank blend -- the reference's computeTopicProbs
// (retrieval/main_retrieve.go:106-159) with its three defects repaired, as an opt-in extension:
//   * `probs` starts at 0 and is only ever multiplied (:142-145), so every probability is 0 -> starts at 1;
//   * inv[2] is looked up by the md5 of the token although the scraper keys it by the plain word
//     (crawler/ODP-scraper.go:132-135) -> the caller passes dense term ids of inv[2]'s own key space;
//   * the value is asserted as map[string]float64 although the table stores map[string]uint32
//     (:123 vs database/database.go:112) -> frequencies arrive as numbers.
// What is kept exactly: multinomial naive Bayes with maximum-likelihood estimates, only the keyword tokens
// (not the phrase) enter, a token contributes to a topic only if its inv[2] row lists that topic, factors are
// multiplied in token order as (freq / wordCount[topic]) in fp64, a topic no token lists gets 0, the uniform
// prior divides by the number of topics at the end (:147).  A tiny dense [Q x T] operation: one thread per
// (query, topic).  The result feeds ss_score_batch(topic_probs, probs_per_query = 1).
#include "index.cuh"

namespace {

__global__ void k_topic_probs(const uint64_t* __restrict__ term_ptr, const uint32_t* __restrict__ topic_ids,
                              const double* __restrict__ freq, const double* __restrict__ word_count,
                              uint64_t n_terms, uint32_t T, const uint64_t* __restrict__ tok_ptr,
                              const uint32_t* __restrict__ tok_terms, uint64_t n_q, double* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_q * T) return;
  const uint64_t q = i / T;
  const uint32_t t = (uint32_t)(i % T);
  double p = 1.0;
  bool any = false;
  const double wc = word_count[t];
  for (uint64_t k = tok_ptr[q]; k < tok_ptr[q + 1]; ++k) {
    const uint32_t term = tok_terms[k];
    if (term >= n_terms) continue;  // a word inv[2] does not know contributes to no topic
    for (uint64_t x = term_ptr[term]; x < term_ptr[term + 1]; ++x)
      if (topic_ids[x] == t) {
        p = __dmul_rn(p, __ddiv_rn(freq[x], wc));
        any = true;
        break;
      }
  }
  out[i] = any ? __ddiv_rn(p, (double)T) : 0.0;
}

}  // namespace

extern "C" {

SS_API int ss_topics_load(ss_engine* e, uint64_t n_terms, uint32_t n_topics, const uint64_t* term_ptr,
                          const uint32_t* topic_ids, const double* freq, const double* word_count) {
  SS_REQUIRE(e && term_ptr && word_count, SS_ERR_INVALID, "ss_topics_load: NULL argument");
  SS_REQUIRE(n_topics >= 1 && n_topics <= 4096, SS_ERR_INVALID, "ss_topics_load: %u topics", n_topics);
  const uint64_t n = term_ptr[n_terms];
  SS_REQUIRE(n == 0 || (topic_ids && freq), SS_ERR_INVALID, "ss_topics_load: NULL rows");
  for (uint64_t t = 0; t < n_terms; ++t)
    SS_REQUIRE(term_ptr[t] <= term_ptr[t + 1], SS_ERR_INVALID, "ss_topics_load: term_ptr not monotone");
  for (uint64_t x = 0; x < n; ++x)
    SS_REQUIRE(topic_ids[x] < n_topics, SS_ERR_INVALID, "ss_topics_load: topic id %u out of range", topic_ids[x]);
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = index_state(e);
  SS_REQUIRE(ix, SS_ERR_OOM, "host allocation failed");
  TopicTable& tt = ix->topics;
  tt.loaded = false;
  SS_TRY(tt.term_ptr.alloc(n_terms + 1));
  SS_TRY(tt.topic_ids.alloc(n));
  SS_TRY(tt.freq.alloc(n));
  SS_TRY(tt.word_count.alloc(n_topics));
  cudaStream_t st = e->stream;
  SS_CUDA(cudaMemcpyAsync(tt.term_ptr.p, term_ptr, (n_terms + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n) {
    SS_CUDA(cudaMemcpyAsync(tt.topic_ids.p, topic_ids, n * 4, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemcpyAsync(tt.freq.p, freq, n * 8, cudaMemcpyHostToDevice, st));
  }
  SS_CUDA(cudaMemcpyAsync(tt.word_count.p, word_count, n_topics * 8, cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaStreamSynchronize(st));
  tt.n_terms = n_terms;
  tt.T = n_topics;
  tt.loaded = true;
  return SS_OK;
}

SS_API int ss_topic_probs(ss_engine* e, uint64_t n_q, const uint64_t* tok_ptr, const uint32_t* tok_terms,
                          double* out_probs) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_topic_probs: engine is NULL");
  if (n_q == 0) return SS_OK;
  SS_REQUIRE(tok_ptr && out_probs, SS_ERR_INVALID, "ss_topic_probs: NULL argument");
  const uint64_t n_tok = tok_ptr[n_q];
  SS_REQUIRE(n_tok == 0 || tok_terms, SS_ERR_INVALID, "ss_topic_probs: NULL tokens");
  for (uint64_t q = 0; q < n_q; ++q)
    SS_REQUIRE(tok_ptr[q] <= tok_ptr[q + 1], SS_ERR_INVALID, "ss_topic_probs: tok_ptr not monotone");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = e->idx;
  SS_REQUIRE(ix && ix->topics.loaded, SS_ERR_STATE, "ss_topic_probs: no topic table loaded (ss_topics_load)");
  TopicTable& tt = ix->topics;
  cudaStream_t st = e->stream;
  ss::DevBuf<uint64_t> d_ptr;
  ss::DevBuf<uint32_t> d_tok;
  ss::DevBuf<double> d_out;
  SS_TRY(d_ptr.alloc(n_q + 1));
  SS_TRY(d_tok.alloc(n_tok));
  SS_TRY(d_out.alloc(n_q * tt.T));
  SS_CUDA(cudaMemcpyAsync(d_ptr.p, tok_ptr, (n_q + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n_tok) SS_CUDA(cudaMemcpyAsync(d_tok.p, tok_terms, n_tok * 4, cudaMemcpyHostToDevice, st));
  k_topic_probs<<<ss::div_up(n_q * tt.T, 256), 256, 0, st>>>(tt.term_ptr.p, tt.topic_ids.p, tt.freq.p, tt.word_count.p,
                                                             tt.n_terms, tt.T, d_ptr.p, d_tok.p, n_q, d_out.p);
  SS_CUDA(cudaMemcpyAsync(out_probs, d_out.p, n_q * tt.T * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

}  // extern "C"
