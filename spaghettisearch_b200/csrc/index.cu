// HP-2 offline: ranking/term_weighting.go:10-123 on the device.
//
//   idf[t]   = float32(Log2(totalDocs / df_t))          term_weighting.go:37
//   w[p]     = normTF[p] * idf[term(p)]   (fp32 multiply) :42
//   mag[doc] = sqrt(sum float64(float32(w*w)))           :44, saveMagnitude :72
//
// Go's math.Log2 (Frexp + FDLIBM log, exact for powers of two) is restated
// with explicitly rounded operations -- nvcc would otherwise contract a*b+c
// into an FMA and a handful of idf values would round differently.
// The reference sums a document's squares in Go map order (random); here and
// in the oracle the order is ascending term id, obtained by a stable sort of
// (doc, weight) pairs so that the norms are bit-reproducible.
#include <cub/device/device_radix_sort.cuh>

#include <cmath>
#include <cstring>

#include "index.cuh"

void index_state_free(IndexState* s) { delete s; }

IndexState* index_state(ss_engine* e) {
  if (!e->idx) e->idx = new (std::nothrow) IndexState();
  return e->idx;
}

namespace {

__device__ __forceinline__ double go_log(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01;
  const double Ln2Lo = 1.90821492927058770002e-10;
  const double L1 = 6.666666666666735130e-01;
  const double L2 = 3.999999999940941908e-01;
  const double L3 = 2.857142874366239149e-01;
  const double L4 = 2.222219843214978396e-01;
  const double L5 = 1.818357216161805012e-01;
  const double L6 = 1.531383769920937332e-01;
  const double L7 = 1.479819860511658591e-01;
  if (isnan(x) || (isinf(x) && x > 0)) return x;
  if (x < 0) return __longlong_as_double(0x7FF8000000000000ll);
  if (x == 0) return -__longlong_as_double(0x7FF0000000000000ll);
  int ki;
  double f1 = frexp(x, &ki);
  if (f1 < 0.70710678118654752440) {
    f1 = __dmul_rn(f1, 2.0);
    ki--;
  }
  const double f = __dadd_rn(f1, -1.0);
  const double k = (double)ki;
  const double s = __ddiv_rn(f, __dadd_rn(2.0, f));
  const double s2 = __dmul_rn(s, s);
  const double s4 = __dmul_rn(s2, s2);
  const double t1 = __dmul_rn(
      s2, __dadd_rn(L1, __dmul_rn(s4, __dadd_rn(L3, __dmul_rn(s4, __dadd_rn(L5, __dmul_rn(s4, L7)))))));
  const double t2 = __dmul_rn(s4, __dadd_rn(L2, __dmul_rn(s4, __dadd_rn(L4, __dmul_rn(s4, L6)))));
  const double R = __dadd_rn(t1, t2);
  const double hfsq = __dmul_rn(__dmul_rn(0.5, f), f);
  // k*Ln2Hi - ((hfsq - (s*(hfsq+R) + k*Ln2Lo)) - f)
  const double inner = __dadd_rn(__dmul_rn(s, __dadd_rn(hfsq, R)), __dmul_rn(k, Ln2Lo));
  return __dadd_rn(__dmul_rn(k, Ln2Hi), -__dadd_rn(__dadd_rn(hfsq, -inner), -f));
}

__device__ __forceinline__ double go_log2(double x) {
  int e;
  const double frac = frexp(x, &e);
  if (frac == 0.5) return (double)(e - 1);
  return __dadd_rn(__dmul_rn(go_log(frac), 1.0 / 0.693147180559945309417232121458176568), (double)e);
}

__global__ void k_idf(const uint64_t* __restrict__ term_ptr, const uint64_t* __restrict__ df_global, uint64_t V,
                      double total_docs, float* __restrict__ idf) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  const uint64_t local = term_ptr[t + 1] - term_ptr[t];
  const uint64_t df = df_global ? df_global[t] : local;
  idf[t] = local ? (float)go_log2(__ddiv_rn(total_docs, (double)df)) : 0.0f;
}

// Each thread weighs kPer consecutive postings: one binary search for the term
// of the first, then a linear walk over term boundaries.
constexpr int kPer = 8;
__global__ void k_weigh(const uint64_t* __restrict__ term_ptr, uint64_t V, uint64_t P,
                        const float* __restrict__ idf, float* __restrict__ w) {
  const uint64_t p0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * kPer;
  if (p0 >= P) return;
  uint64_t lo = 0, hi = V;  // last t with term_ptr[t] <= p0
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (term_ptr[mid] <= p0) lo = mid; else hi = mid;
  }
  uint64_t t = lo;
  uint64_t next = term_ptr[t + 1];
  float f = idf[t];
  const uint64_t pe = min(P, p0 + kPer);
  for (uint64_t p = p0; p < pe; ++p) {
    while (p >= next) {
      ++t;
      next = term_ptr[t + 1];
      f = idf[t];
    }
    w[p] = __fmul_rn(w[p], f);
  }
}

// (doc, w) pairs stably sorted by doc => a doc's weights are in ascending term
// order; one thread per doc folds its run.
__global__ void k_doc_norm(const uint32_t* __restrict__ doc_sorted, const float* __restrict__ w_sorted, uint64_t P,
                           uint64_t D, double* __restrict__ mag) {
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  uint64_t lo = 0, hi = P;  // first j with doc_sorted[j] >= d
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (doc_sorted[mid] >= d) hi = mid; else lo = mid + 1;
  }
  double acc = 0.0;
  for (uint64_t j = lo; j < P && doc_sorted[j] == d; ++j) {
    const float wv = w_sorted[j];
    acc = __dadd_rn(acc, (double)__fmul_rn(wv, wv));  // square rounded to fp32 first
  }
  mag[d] = sqrt(acc);
}

__global__ void k_check_docs(const uint32_t* __restrict__ doc_ids, uint64_t P, uint64_t D, int* __restrict__ bad) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < P && doc_ids[p] >= D) *bad = 1;
}
__global__ void k_check_ptr(const uint64_t* __restrict__ ptr, uint64_t n, uint64_t total, int* __restrict__ bad) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && ptr[i] > ptr[i + 1]) *bad = 1;
  if (i == 0 && (ptr[0] != 0 || ptr[n] != total)) *bad = 1;
}
// doc ids strictly ascending inside each term row (a Go map has unique keys)
__global__ void k_check_sorted(const uint64_t* __restrict__ term_ptr, uint64_t V, const uint32_t* __restrict__ doc_ids,
                               uint64_t P, int* __restrict__ bad) {
  const uint64_t p = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p + 1 >= P) return;
  if (doc_ids[p] < doc_ids[p + 1]) return;
  // allowed only at a term boundary: find whether p+1 starts a row
  uint64_t lo = 0, hi = V;
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (term_ptr[mid] <= p + 1) lo = mid; else hi = mid;
  }
  // several empty rows may share the offset; any row starting exactly at p+1 makes it a boundary
  if (term_ptr[lo] != p + 1) *bad = 1;
}

}  // namespace

// implemented in pagerank.cu: copies the last PageRank result, unscaled, [N][T]
int pagerank_export_device(ss_engine* e, ss::DevBuf<double>* out, uint64_t* n_rows, uint32_t* n_topics);

extern "C" {

SS_API int ss_index_load(ss_engine* e, int table, uint64_t n_terms, uint64_t n_docs, const uint64_t* term_ptr,
                         const uint32_t* doc_ids, const float* norm_tf, const uint64_t* pos_ptr, const float* pos) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_index_load: engine is NULL");
  SS_REQUIRE(table == SS_TITLE || table == SS_BODY, SS_ERR_INVALID, "ss_index_load: table %d", table);
  SS_REQUIRE(term_ptr, SS_ERR_INVALID, "ss_index_load: term_ptr is NULL");
  SS_REQUIRE(n_docs < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_index_load: doc ids are 32 bit");
  const uint64_t V = n_terms, P = term_ptr[V];
  SS_REQUIRE(P == 0 || (doc_ids && norm_tf), SS_ERR_INVALID, "ss_index_load: NULL postings");
  SS_REQUIRE(P < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_index_load: %llu postings; shard the table below 2^32",
             (unsigned long long)P);
  SS_REQUIRE(!pos_ptr || pos || pos_ptr[P] == 0, SS_ERR_INVALID, "ss_index_load: pos is NULL");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = index_state(e);
  SS_REQUIRE(ix, SS_ERR_OOM, "host allocation failed");
  const int other = table ^ 1;
  SS_REQUIRE(!ix->tab[other].loaded || ix->D == n_docs, SS_ERR_INVALID,
             "ss_index_load: n_docs %llu differs from the other table's %llu", (unsigned long long)n_docs,
             (unsigned long long)ix->D);
  cudaStream_t st = e->stream;
  TableState& tb = ix->tab[table];
  tb.clear();
  ix->D = n_docs;
  tb.V = V;
  tb.P = P;
  SS_TRY(tb.term_ptr.alloc(V + 1));
  SS_TRY(tb.doc_ids.alloc(P));
  SS_TRY(tb.w.alloc(P));
  SS_CUDA(cudaMemcpyAsync(tb.term_ptr.p, term_ptr, (V + 1) * 8, cudaMemcpyHostToDevice, st));
  if (P) {
    SS_CUDA(cudaMemcpyAsync(tb.doc_ids.p, doc_ids, P * 4, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemcpyAsync(tb.w.p, norm_tf, P * 4, cudaMemcpyHostToDevice, st));
  }
  if (pos_ptr) {
    const uint64_t n_pos = pos_ptr[P];
    tb.has_pos = true;
    SS_TRY(tb.pos_ptr.alloc(P + 1));
    SS_TRY(tb.pos.alloc(n_pos));
    SS_CUDA(cudaMemcpyAsync(tb.pos_ptr.p, pos_ptr, (P + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_pos) SS_CUDA(cudaMemcpyAsync(tb.pos.p, pos, n_pos * 4, cudaMemcpyHostToDevice, st));
  }
  // validate: monotone pointers, doc ids in range and ascending inside a row
  ss::DevBuf<int> d_bad;
  SS_TRY(d_bad.alloc(1));
  SS_CUDA(cudaMemsetAsync(d_bad.p, 0, sizeof(int), st));
  if (V) k_check_ptr<<<ss::div_up(V, 256), 256, 0, st>>>(tb.term_ptr.p, V, P, d_bad.p);
  if (P) {
    k_check_docs<<<ss::div_up(P, 256), 256, 0, st>>>(tb.doc_ids.p, P, n_docs, d_bad.p);
    if (V) k_check_sorted<<<ss::div_up(P, 256), 256, 0, st>>>(tb.term_ptr.p, V, tb.doc_ids.p, P, d_bad.p);
    if (pos_ptr) k_check_ptr<<<ss::div_up(P, 256), 256, 0, st>>>(tb.pos_ptr.p, P, pos_ptr[P], d_bad.p);
  }
  int bad = 0;
  SS_CUDA(cudaMemcpyAsync(&bad, d_bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  if (bad) {
    tb.clear();
    ss::set_error("ss_index_load: pointers not monotone, doc id >= n_docs, or docs not ascending within a term");
    return SS_ERR_INVALID;
  }
  tb.df_host.resize(V);
  for (uint64_t t = 0; t < V; ++t) tb.df_host[t] = (uint32_t)(term_ptr[t + 1] - term_ptr[t]);
  tb.loaded = true;
  // everything derived from the doc id space, the weights or the norms is stale now
  ix->dense_valid = false;
  ix->sqd_valid = false;
  ix->meta32_valid = false;
  ix->zvec_valid = false;
  return SS_OK;
}

SS_API int ss_index_set_doc_base(ss_engine* e, uint64_t doc_base) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_index_set_doc_base: engine is NULL");
  SS_REQUIRE(doc_base < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_index_set_doc_base: doc ids are 32 bit");
  std::lock_guard<std::mutex> lock(e->mu);
  IndexState* ix = index_state(e);
  SS_REQUIRE(ix, SS_ERR_OOM, "host allocation failed");
  ix->doc_base = doc_base;
  return SS_OK;
}

SS_API int ss_index_clear(ss_engine* e) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_index_clear: engine is NULL");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  index_state_free(e->idx);
  e->idx = nullptr;
  return SS_OK;
}

SS_API int ss_term_weights(ss_engine* e, int table, double total_docs, const uint64_t* df_global, float* out_w,
                           double* out_mag) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_term_weights: engine is NULL");
  SS_REQUIRE(table == SS_TITLE || table == SS_BODY, SS_ERR_INVALID, "ss_term_weights: table %d", table);
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = e->idx;
  SS_REQUIRE(ix && ix->tab[table].loaded, SS_ERR_STATE, "ss_term_weights: table %d not loaded", table);
  TableState& tb = ix->tab[table];
  cudaStream_t st = e->stream;
  const uint64_t V = tb.V, P = tb.P, D = ix->D;

  ss::DevBuf<float> idf;
  ss::DevBuf<uint64_t> d_df;
  SS_TRY(idf.alloc(V));
  if (df_global) {
    SS_TRY(d_df.alloc(V));
    SS_CUDA(cudaMemcpyAsync(d_df.p, df_global, V * 8, cudaMemcpyHostToDevice, st));
  }
  if (V) k_idf<<<ss::div_up(V, 256), 256, 0, st>>>(tb.term_ptr.p, df_global ? d_df.p : nullptr, V, total_docs, idf.p);
  if (P) k_weigh<<<ss::div_up(ss::div_up(P, kPer), 256), 256, 0, st>>>(tb.term_ptr.p, V, P, idf.p, tb.w.p);

  SS_TRY(tb.mag.alloc(D));
  {
    ss::DevBuf<uint32_t> doc_sorted;
    ss::DevBuf<float> w_sorted;
    ss::DevBuf<char> tmp;
    SS_TRY(doc_sorted.alloc(P));
    SS_TRY(w_sorted.alloc(P));
    if (P) {
      int end_bit = 1;
      while ((1ull << end_bit) < D) ++end_bit;
      size_t tmp_bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, tb.doc_ids.p, doc_sorted.p, tb.w.p, w_sorted.p, (int64_t)P,
                                      0, end_bit, st);
      SS_TRY(tmp.alloc(tmp_bytes));
      SS_CUDA(cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, tb.doc_ids.p, doc_sorted.p, tb.w.p, w_sorted.p,
                                              (int64_t)P, 0, end_bit, st));
    }
    if (D) k_doc_norm<<<ss::div_up(D, 256), 256, 0, st>>>(doc_sorted.p, w_sorted.p, P, D, tb.mag.p);
    SS_CUDA(cudaStreamSynchronize(st));
  }
  tb.has_mag = true;
  ix->meta32_valid = false;
  ix->dense_valid = false;
  if (out_w && P) SS_CUDA(cudaMemcpyAsync(out_w, tb.w.p, P * 4, cudaMemcpyDeviceToHost, st));
  if (out_mag && D) SS_CUDA(cudaMemcpyAsync(out_mag, tb.mag.p, D * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

SS_API int ss_set_doc_norms(ss_engine* e, int table, uint64_t n_docs, const double* mag) {
  SS_REQUIRE(e && mag, SS_ERR_INVALID, "ss_set_doc_norms: NULL argument");
  SS_REQUIRE(table == SS_TITLE || table == SS_BODY, SS_ERR_INVALID, "ss_set_doc_norms: table %d", table);
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = e->idx;
  SS_REQUIRE(ix && ix->tab[table].loaded, SS_ERR_STATE, "ss_set_doc_norms: table %d not loaded", table);
  SS_REQUIRE(n_docs == ix->D, SS_ERR_INVALID, "ss_set_doc_norms: n_docs %llu != %llu", (unsigned long long)n_docs,
             (unsigned long long)ix->D);
  TableState& tb = ix->tab[table];
  SS_TRY(tb.mag.alloc(n_docs));
  if (n_docs) SS_CUDA(cudaMemcpyAsync(tb.mag.p, mag, n_docs * 8, cudaMemcpyHostToDevice, e->stream));
  SS_CUDA(cudaStreamSynchronize(e->stream));
  tb.has_mag = true;
  ix->meta32_valid = false;
  ix->dense_valid = false;
  return SS_OK;
}

SS_API int ss_set_pagerank(ss_engine* e, uint64_t n_docs, uint32_t n_topics, const double* rank) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_set_pagerank: engine is NULL");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = index_state(e);
  SS_REQUIRE(ix, SS_ERR_OOM, "host allocation failed");
  ix->sqd_valid = false;
  ix->meta32_valid = false;
  ix->dense_valid = false;
  if (!rank || n_docs == 0 || n_topics == 0) {
    ix->pr.reset();
    ix->T = 0;
    ix->pr_docs = 0;
    return SS_OK;
  }
  SS_TRY(ix->pr.alloc(n_docs * n_topics));
  SS_CUDA(cudaMemcpyAsync(ix->pr.p, rank, n_docs * n_topics * 8, cudaMemcpyHostToDevice, e->stream));
  SS_CUDA(cudaStreamSynchronize(e->stream));
  ix->T = n_topics;
  ix->pr_docs = n_docs;
  return SS_OK;
}

SS_API int ss_use_pagerank(ss_engine* e) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_use_pagerank: engine is NULL");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = index_state(e);
  SS_REQUIRE(ix, SS_ERR_OOM, "host allocation failed");
  ix->sqd_valid = false;
  ix->meta32_valid = false;
  ix->dense_valid = false;
  uint64_t rows = 0;
  uint32_t topics = 0;
  SS_TRY(pagerank_export_device(e, &ix->pr, &rows, &topics));
  ix->pr_docs = rows;
  ix->T = topics;
  return SS_OK;
}

}  // extern "C"
