// CPU oracle -- TEST INFRASTRUCTURE ONLY (see oracle.h).  PARITY UNPINNED by
// the reference's own tests; pinned by hand-derived known answers instead.
//
// Build with -ffp-contract=off: Go on amd64 never fuses a*b+c, and the
// reference's results are defined by separately rounded operations.
#include "oracle.h"

#include <omp.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <unordered_map>
#include <vector>

// ------------------------------------------------------------------ log ----
// go/src/math/log.go (a port of FreeBSD e_log.c): argument reduction
// x = 2^k (1+f), s = f/(2+f), 14th degree polynomial in s.
extern "C" double oracle_go_log(double x) {
  const double Ln2Hi = 6.93147180369123816490e-01;
  const double Ln2Lo = 1.90821492927058770002e-10;
  const double L1 = 6.666666666666735130e-01;
  const double L2 = 3.999999999940941908e-01;
  const double L3 = 2.857142874366239149e-01;
  const double L4 = 2.222219843214978396e-01;
  const double L5 = 1.818357216161805012e-01;
  const double L6 = 1.531383769920937332e-01;
  const double L7 = 1.479819860511658591e-01;
  if (std::isnan(x) || (std::isinf(x) && x > 0)) return x;
  if (x < 0) return std::numeric_limits<double>::quiet_NaN();
  if (x == 0) return -std::numeric_limits<double>::infinity();
  int ki;
  double f1 = std::frexp(x, &ki);
  if (f1 < 0.70710678118654752440 /* Sqrt2/2 */) {
    f1 *= 2;
    ki--;
  }
  double f = f1 - 1;
  double k = (double)ki;
  double s = f / (2 + f);
  double s2 = s * s;
  double s4 = s2 * s2;
  double t1 = s2 * (L1 + s4 * (L3 + s4 * (L5 + s4 * L7)));
  double t2 = s4 * (L2 + s4 * (L4 + s4 * L6));
  double R = t1 + t2;
  double hfsq = 0.5 * f * f;
  return k * Ln2Hi - ((hfsq - (s * (hfsq + R) + k * Ln2Lo)) - f);
}

// go/src/math/log10.go log2(): Frexp, exact answer for powers of two.
extern "C" double oracle_go_log2(double x) {
  int e;
  double frac = std::frexp(x, &e);
  if (frac == 0.5) return (double)(e - 1);
  // frexp passes Inf/NaN/0 through; Log handles them
  return oracle_go_log(frac) * (1.0 / 0.693147180559945309417232121458176568) + (double)e;
}

// -------------------------------------------------------------- PageRank ----
// ranking/pagerank.go:126-145 on dense ids, parents ascending.
static double rank_inherited(std::vector<double>& cur, const std::vector<double>& last, double d,
                             uint64_t n, const uint64_t* row_ptr, const uint32_t* col_idx) {
  double total = 0.0;
  for (uint64_t p = 0; p < n; ++p) {
    uint64_t b = row_ptr[p], e = row_ptr[p + 1];
    if (e == b) continue;  // pagerank.go:132-134
    double w = d * last[p] / (double)(e - b);
    total += w;  // once per parent, pagerank.go:137
    for (uint64_t i = b; i < e; ++i) cur[col_idx[i]] += w;
  }
  return total;
}

extern "C" int oracle_pagerank(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                               double damping, double eps, uint32_t n_topics,
                               const int64_t* num_pages, uint32_t max_iters, double* out_rank,
                               uint32_t* out_iters) {
  const uint64_t N = n_nodes;
  const double teleport = 1.0 - damping;  // pagerank.go:90
  std::vector<double> cur(N), last(N);
  for (uint32_t t = 0; t < n_topics; ++t) {
    const double init = 1.0 / (double)num_pages[t];  // pagerank.go:104-105
    uint32_t iteration = 1;
    double change = std::numeric_limits<double>::max();
    for (; change > eps; ++iteration) {  // pagerank.go:93
      cur.swap(last);
      if (iteration > 1) {
        std::fill(cur.begin(), cur.end(), 0.0);
      } else {
        std::fill(cur.begin(), cur.end(), init);
        std::fill(last.begin(), last.end(), init);
      }
      double total = rank_inherited(cur, last, damping, N, row_ptr, col_idx);
      total += teleport * (double)N;  // pagerank.go:112
      change = 0.0;
      for (uint64_t v = 0; v < N; ++v) {
        cur[v] = (cur[v] + teleport) / total;
        change += std::fabs(cur[v] - last[v]);
      }
      if (max_iters && iteration >= max_iters) {
        ++iteration;
        break;
      }
    }
    if (out_iters) out_iters[t] = iteration - 1;
    for (uint64_t v = 0; v < N; ++v) out_rank[v * n_topics + t] = cur[v];
  }
  return 0;
}

static std::string hex_key(uint64_t u) {
  static const char* digits = "0123456789abcdef";
  auto mix = [](uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
  };
  uint64_t a = mix(u), b = mix(u ^ 0x5555555555555555ull);
  std::string s(32, '0');
  for (int i = 0; i < 16; ++i) {
    s[i] = digits[(a >> (60 - 4 * i)) & 15];
    s[16 + i] = digits[(b >> (60 - 4 * i)) & 15];
  }
  return s;
}

extern "C" int oracle_pagerank_faithful(uint64_t n_nodes, const uint64_t* row_ptr,
                                        const uint32_t* col_idx, double damping, double eps,
                                        int64_t num_pages, uint32_t fixed_iters, double* out_rank,
                                        uint32_t* out_iters, double* sweep_seconds) {
  // pagerank.go:24-44: webNodes map[string][]string, setWebNodes []string
  std::vector<std::string> keys(n_nodes);
  for (uint64_t u = 0; u < n_nodes; ++u) keys[u] = hex_key(u);
  std::unordered_map<std::string, std::vector<std::string>> web;
  web.reserve(n_nodes);
  for (uint64_t u = 0; u < n_nodes; ++u) {
    if (row_ptr[u + 1] == row_ptr[u]) continue;  // never crawled: not a forw[2] key
    auto& kids = web[keys[u]];
    kids.reserve(row_ptr[u + 1] - row_ptr[u]);
    for (uint64_t i = row_ptr[u]; i < row_ptr[u + 1]; ++i) kids.push_back(keys[col_idx[i]]);
  }
  std::unordered_map<std::string, double> cur, last;
  cur.reserve(n_nodes);
  last.reserve(n_nodes);
  const double teleport = 1.0 - damping;
  const double init = 1.0 / (double)num_pages;
  auto t0 = std::chrono::steady_clock::now();
  uint32_t iteration = 1;
  double change = std::numeric_limits<double>::max();
  for (; fixed_iters ? iteration <= fixed_iters : change > eps; ++iteration) {
    cur.swap(last);
    if (iteration > 1) {
      for (auto& k : keys) cur[k] = 0.0;
    } else {
      for (auto& k : keys) {
        cur[k] = init;
        last[k] = init;
      }
    }
    double total = 0.0;
    for (auto& kv : cur) {  // pagerank.go:130, map order
      auto it = web.find(kv.first);
      if (it == web.end() || it->second.empty()) continue;
      double w = damping * last[kv.first] / (double)it->second.size();
      total += w;
      for (auto& c : it->second) cur[c] += w;
    }
    total += teleport * (double)cur.size();
    change = 0.0;
    for (auto& kv : cur) {
      kv.second = (kv.second + teleport) / total;
      change += std::fabs(kv.second - last[kv.first]);
    }
  }
  auto t1 = std::chrono::steady_clock::now();
  if (sweep_seconds) *sweep_seconds = std::chrono::duration<double>(t1 - t0).count();
  if (out_iters) *out_iters = iteration - 1;
  if (out_rank)
    for (uint64_t u = 0; u < n_nodes; ++u) out_rank[u] = cur[keys[u]];
  return 0;
}

// in-edge CSC of the out-edge CSR (sources ascending within a row)
extern "C" int oracle_csc_build(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                                uint64_t* in_ptr, uint32_t* in_src) {
  const uint64_t N = n_nodes, E = row_ptr[N];
  for (uint64_t v = 0; v <= N; ++v) in_ptr[v] = 0;
  for (uint64_t i = 0; i < E; ++i) in_ptr[col_idx[i] + 1]++;
  for (uint64_t v = 0; v < N; ++v) in_ptr[v + 1] += in_ptr[v];
  std::vector<uint64_t> fill(in_ptr, in_ptr + N);
  for (uint64_t u = 0; u < N; ++u)
    for (uint64_t i = row_ptr[u]; i < row_ptr[u + 1]; ++i) in_src[fill[col_idx[i]]++] = (uint32_t)u;
  return 0;
}

// tele_w: NULL = the reference's uniform teleport (pagerank.go:90,117: every node gets 1-d); else [N][T]
// weights N * v_t[v] of a per-topic teleport vector v_t (sum_v v_t[v] = 1), SURVEY.md 8(f)-4: node v of topic
// t gets (1-d) * tele_w[v][t] instead, Tot is unchanged because the weights of a topic sum to N.
static int pagerank_fair_csc_impl(uint64_t n_nodes, const uint64_t* row_ptr, const uint64_t* in_ptr,
                                  const uint32_t* in_src, double damping, double eps,
                                  uint32_t n_topics, const int64_t* num_pages,
                                  uint32_t max_iters, uint32_t fixed_iters, int n_threads,
                                  const double* tele_w, double* out_rank, uint32_t* out_iters,
                                  double* sweep_seconds) {
  const uint64_t N = n_nodes;
  const uint32_t T = n_topics;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  const double teleport = 1.0 - damping;
  std::vector<double> last(N * T), cur(N * T), contrib(N * T);
  std::vector<double> tot(T), change(T, std::numeric_limits<double>::max());
  std::vector<char> active(T, 1);
  std::vector<uint32_t> iters(T, 0);
#pragma omp parallel for num_threads(n_threads) schedule(static)
  for (uint64_t v = 0; v < N; ++v)
    for (uint32_t t = 0; t < T; ++t) last[v * T + t] = 1.0 / (double)num_pages[t];
  auto t0 = std::chrono::steady_clock::now();
  for (uint32_t iteration = 1;; ++iteration) {
    bool any = false;
    for (uint32_t t = 0; t < T; ++t) {
      active[t] = fixed_iters ? (iteration <= fixed_iters)
                              : (change[t] > eps && !(max_iters && iteration > max_iters));
      any |= active[t];
    }
    if (!any) break;
    std::vector<double> S(T, 0.0);
#pragma omp parallel num_threads(n_threads)
    {
      std::vector<double> s_loc(T, 0.0);
#pragma omp for schedule(static)
      for (uint64_t u = 0; u < N; ++u) {
        uint64_t od = row_ptr[u + 1] - row_ptr[u];
        for (uint32_t t = 0; t < T; ++t) {
          double w = od ? damping * last[u * T + t] / (double)od : 0.0;
          contrib[u * T + t] = w;
          s_loc[t] += w;
        }
      }
#pragma omp critical
      for (uint32_t t = 0; t < T; ++t) S[t] += s_loc[t];
    }
    for (uint32_t t = 0; t < T; ++t) tot[t] = S[t] + teleport * (double)N;
    std::vector<double> delta(T, 0.0);
#pragma omp parallel num_threads(n_threads)
    {
      std::vector<double> d_loc(T, 0.0), acc(T);
#pragma omp for schedule(dynamic, 2048)
      for (uint64_t v = 0; v < N; ++v) {
        for (uint32_t t = 0; t < T; ++t) acc[t] = iteration == 1 ? last[v * T + t] : 0.0;
        for (uint64_t i = in_ptr[v]; i < in_ptr[v + 1]; ++i) {
          const double* c = &contrib[(uint64_t)in_src[i] * T];
          for (uint32_t t = 0; t < T; ++t) acc[t] += c[t];
        }
        for (uint32_t t = 0; t < T; ++t) {
          if (!active[t]) {
            cur[v * T + t] = last[v * T + t];
            continue;
          }
          double r = (acc[t] + (tele_w ? teleport * tele_w[v * T + t] : teleport)) / tot[t];
          cur[v * T + t] = r;
          d_loc[t] += std::fabs(r - last[v * T + t]);
        }
      }
#pragma omp critical
      for (uint32_t t = 0; t < T; ++t) delta[t] += d_loc[t];
    }
    for (uint32_t t = 0; t < T; ++t)
      if (active[t]) {
        change[t] = delta[t];
        iters[t] = iteration;
      }
    cur.swap(last);
  }
  auto t1 = std::chrono::steady_clock::now();
  if (sweep_seconds) *sweep_seconds = std::chrono::duration<double>(t1 - t0).count();
  if (out_iters) memcpy(out_iters, iters.data(), T * sizeof(uint32_t));
  if (out_rank) memcpy(out_rank, last.data(), N * T * sizeof(double));
  return 0;
}

extern "C" int oracle_pagerank_fair_csc(uint64_t n_nodes, const uint64_t* row_ptr, const uint64_t* in_ptr,
                                        const uint32_t* in_src, double damping, double eps,
                                        uint32_t n_topics, const int64_t* num_pages,
                                        uint32_t max_iters, uint32_t fixed_iters, int n_threads,
                                        double* out_rank, uint32_t* out_iters,
                                        double* sweep_seconds) {
  return pagerank_fair_csc_impl(n_nodes, row_ptr, in_ptr, in_src, damping, eps, n_topics, num_pages, max_iters,
                                fixed_iters, n_threads, nullptr, out_rank, out_iters, sweep_seconds);
}
// Extension (not in the reference as shipped): genuinely topic-biased teleport, README.md:9 / SURVEY.md 8(f)-4.
extern "C" int oracle_pagerank_biased(uint64_t n_nodes, const uint64_t* row_ptr, const uint32_t* col_idx,
                                      double damping, double eps, uint32_t n_topics,
                                      const int64_t* num_pages, uint32_t max_iters, int n_threads,
                                      const double* tele_w, double* out_rank, uint32_t* out_iters) {
  const uint64_t N = n_nodes, E = row_ptr[N];
  std::vector<uint64_t> in_ptr(N + 1);
  std::vector<uint32_t> in_src(E ? E : 1);
  oracle_csc_build(N, row_ptr, col_idx, in_ptr.data(), in_src.data());
  return pagerank_fair_csc_impl(N, row_ptr, in_ptr.data(), in_src.data(), damping, eps, n_topics, num_pages,
                                max_iters, 0, n_threads, tele_w, out_rank, out_iters, nullptr);
}
extern "C" int oracle_pagerank_fair(uint64_t n_nodes, const uint64_t* row_ptr,
                                    const uint32_t* col_idx, double damping, double eps,
                                    uint32_t n_topics, const int64_t* num_pages,
                                    uint32_t max_iters, uint32_t fixed_iters, int n_threads,
                                    double* out_rank, uint32_t* out_iters,
                                    double* sweep_seconds) {
  const uint64_t N = n_nodes, E = row_ptr[N];
  std::vector<uint64_t> in_ptr(N + 1);
  std::vector<uint32_t> in_src(E ? E : 1);
  oracle_csc_build(N, row_ptr, col_idx, in_ptr.data(), in_src.data());
  return oracle_pagerank_fair_csc(N, row_ptr, in_ptr.data(), in_src.data(), damping, eps, n_topics,
                                  num_pages, max_iters, fixed_iters, n_threads, out_rank, out_iters,
                                  sweep_seconds);
}

// ---------------------------------------------------------- term weights ----
extern "C" int oracle_term_weights(uint64_t n_terms, uint64_t n_docs, const uint64_t* term_ptr,
                                   const uint32_t* doc_ids, const float* norm_tf,
                                   double total_docs, float* out_w, double* out_mag) {
  std::vector<double> mag(n_docs, 0.0);  // pageMagnitude, term_weighting.go:27
  for (uint64_t t = 0; t < n_terms; ++t) {
    uint64_t b = term_ptr[t], e = term_ptr[t + 1];
    if (e == b) continue;  // a term with no row does not exist in the table
    float idf = (float)oracle_go_log2(total_docs / (double)(e - b));  // :37
    for (uint64_t p = b; p < e; ++p) {
      float w = norm_tf[p] * idf;  // :42, fp32 multiply
      out_w[p] = w;
      float sq = w * w;  // :44, rounded to fp32 before widening
      mag[doc_ids[p]] += (double)sq;
    }
  }
  for (uint64_t d = 0; d < n_docs; ++d) out_mag[d] = std::sqrt(mag[d]);  // :72,97,105
  return 0;
}

// ------------------------------------------------------------- retrieval ----
// retrieval/util.go:162-177
static void sort_float32(std::vector<float>& s) { std::sort(s.begin(), s.end()); }

// retrieval/util.go:179-203.  nil is modelled by `valid == false`.
struct PosList {
  bool valid = false;
  std::vector<float> v;
};
static PosList intersect(PosList a, PosList b) {
  PosList r;
  if (!a.valid || !b.valid) return r;  // nil in => nil out
  sort_float32(a.v);
  sort_float32(b.v);
  size_t i = 0, j = 0;
  while (i != a.v.size() && j != b.v.size()) {
    if (a.v[i] == b.v[j]) {
      r.v.push_back(a.v[i]);
      ++i;
      ++j;
    } else if (a.v[i] > b.v[j]) {
      ++j;
    } else {
      ++i;
    }
  }
  r.valid = !r.v.empty();  // `var ret []float32` stays nil when nothing is appended
  return r;
}

extern "C" uint64_t oracle_intersect(float* a, uint64_t na, float* b, uint64_t nb, float* out) {
  PosList x, y;
  x.valid = a != nullptr;
  y.valid = b != nullptr;
  if (a) x.v.assign(a, a + na);
  if (b) y.v.assign(b, b + nb);
  PosList r = intersect(x, y);
  for (size_t i = 0; i < r.v.size(); ++i) out[i] = r.v[i];
  return r.v.size();
}

namespace {

struct Posting {
  bool present = false;
  float w = 0;
  std::vector<float> pos;  // already shifted by the term's phrase position
};
struct PhraseEntry {  // Rank_term as used by phrase.go
  Posting title, body;
};
struct DocWeights {  // Rank_term in Retrieve: weight lists, summed in arrival order
  double title = 0.0, body = 0.0;
};

inline void table_range(const oracle_table* tb, uint32_t term, uint64_t* b, uint64_t* e) {
  if (!tb || term >= tb->n_terms) {
    *b = *e = 0;  // ErrKeyNotFound => empty
    return;
  }
  *b = tb->term_ptr[term];
  *e = tb->term_ptr[term + 1];
}

// retrieval/phrase.go:11-109
void eval_phrase(const oracle_table* title, const oracle_table* body, const uint32_t* ph,
                 uint64_t L, std::unordered_map<uint32_t, std::pair<bool, float>>& out_title,
                 std::unordered_map<uint32_t, std::pair<bool, float>>& out_body) {
  if (L == 0 || L > 256) return;  // TermPos is uint8: >256 tokens can never fill every slot
  // aggregatedResult: doc -> TermPos -> entry (phrase.go:26-44)
  std::unordered_map<uint32_t, std::unordered_map<uint8_t, PhraseEntry>> agg;
  for (uint64_t i = 0; i < L; ++i) {
    const uint8_t tp = (uint8_t)i;
    const float shift = (float)tp;
    uint64_t b, e;
    // getPosTerm, phrase.go:120-170: one map per term, then overwrite into agg[doc][tp]
    std::unordered_map<uint32_t, PhraseEntry> ret;
    table_range(body, ph[i], &b, &e);
    for (uint64_t p = b; p < e; ++p) {
      PhraseEntry& en = ret[body->doc_ids[p]];
      en.body.present = true;
      en.body.w = body->w[p];
      if (body->pos_ptr)
        for (uint64_t k = body->pos_ptr[p]; k < body->pos_ptr[p + 1]; ++k)
          en.body.pos.push_back(body->pos[k] - shift);  // :144-146
    }
    table_range(title, ph[i], &b, &e);
    for (uint64_t p = b; p < e; ++p) {
      PhraseEntry& en = ret[title->doc_ids[p]];
      en.title.present = true;
      en.title.w = title->w[p];
      if (title->pos_ptr)
        for (uint64_t k = title->pos_ptr[p]; k < title->pos_ptr[p + 1]; ++k)
          en.title.pos.push_back(title->pos[k] - shift);  // :156-158
    }
    for (auto& kv : ret) agg[kv.first][tp] = std::move(kv.second);
  }
  // evalPhraseOccurrence, phrase.go:53-109
  for (auto& dv : agg) {
    auto& tw = dv.second;
    float sum_body = 0, sum_title = 0;
    PosList bi, ti;
    if (tw.size() == L) {
      PhraseEntry& e0 = tw[0];
      if (e0.body.present) {
        sum_body += e0.body.w;
        bi.valid = true;
        bi.v = e0.body.pos;
      }
      if (e0.title.present) {
        sum_title += e0.title.w;
        ti.valid = true;
        ti.v = e0.title.pos;
      }
      for (uint64_t idx = 1; idx < tw.size(); ++idx) {
        PhraseEntry& en = tw[(uint8_t)idx];
        if (!en.body.present) {
          bi = PosList();
        } else {
          sum_body += en.body.w;
          PosList o;
          o.valid = true;
          o.v = en.body.pos;
          bi = intersect(bi, o);
        }
        if (!en.title.present) {
          ti = PosList();
        } else {
          sum_title += en.title.w;
          PosList o;
          o.valid = true;
          o.v = en.title.pos;
          ti = intersect(ti, o);
        }
      }
    }
    bool hb = bi.valid && !bi.v.empty(), ht = ti.valid && !ti.v.empty();
    if (hb) out_body[dv.first] = {true, sum_body};
    if (ht) out_title[dv.first] = {true, sum_title};
  }
}

struct Hit {
  uint32_t doc;
  double final_rank, pr;
};
inline bool hit_before(const Hit& a, const Hit& b) {
  bool an = std::isnan(a.final_rank), bn = std::isnan(b.final_rank);
  if (an != bn) return bn;  // NaN sorts last
  if (!an && a.final_rank != b.final_rank) return a.final_rank > b.final_rank;
  return a.doc < b.doc;
}

}  // namespace

extern "C" int oracle_score_batch(const oracle_table* title, const oracle_table* body,
                                  uint64_t n_docs, const double* mag_title,
                                  const double* mag_body, const double* pagerank,
                                  uint32_t n_topics, uint64_t n_q, const uint64_t* kw_ptr,
                                  const uint32_t* kw_terms, const uint64_t* ph_ptr,
                                  const uint32_t* ph_terms, const double* topic_probs,
                                  int probs_per_query, uint32_t k, uint32_t* out_doc,
                                  double* out_final, double* out_pr, uint32_t* out_count,
                                  int n_threads) {
  (void)n_docs;
  if (n_threads <= 0) n_threads = omp_get_max_threads();
#pragma omp parallel for num_threads(n_threads) schedule(dynamic, 1)
  for (uint64_t q = 0; q < n_q; ++q) {
    const uint64_t kb = kw_ptr[q], ke = kw_ptr[q + 1];
    const uint64_t pb = ph_ptr ? ph_ptr[q] : 0, pe = ph_ptr ? ph_ptr[q + 1] : 0;
    // keyword terms: main_retrieve.go:50-69,204-247 (token order fixed here)
    std::unordered_map<uint32_t, DocWeights> agg;
    for (uint64_t i = kb; i < ke; ++i) {
      uint64_t b, e;
      table_range(body, kw_terms[i], &b, &e);
      for (uint64_t p = b; p < e; ++p) agg[body->doc_ids[p]].body += (double)body->w[p];
      table_range(title, kw_terms[i], &b, &e);
      for (uint64_t p = b; p < e; ++p) agg[title->doc_ids[p]].title += (double)title->w[p];
    }
    // phrase docs appended last: main_retrieve.go:73-78
    if (pe > pb) {
      std::unordered_map<uint32_t, std::pair<bool, float>> pt, pbod;
      eval_phrase(title, body, ph_terms + pb, pe - pb, pt, pbod);
      for (auto& kv : pbod) agg[kv.first].body += (double)kv.second.second;
      for (auto& kv : pt) agg[kv.first].title += (double)kv.second.second;
    }
    // computeFinalRank: get_metadata.go:38-69
    const double* probs = topic_probs ? topic_probs + (probs_per_query ? q * n_topics : 0) : nullptr;
    const double qmag = std::sqrt((double)((ke - kb) + (pe - pb)));  // :53, main_retrieve.go:90
    std::vector<Hit> hits;
    hits.reserve(agg.size());
    for (auto& kv : agg) {
      const uint32_t doc = kv.first;
      double sqd = 0.0;
      if (probs && pagerank)
        for (uint32_t t = 0; t < n_topics; ++t) sqd += probs[t] * pagerank[(uint64_t)doc * n_topics + t];
      double br = kv.second.body / (mag_body[doc] * qmag);    // :57
      double tr = kv.second.title / (mag_title[doc] * qmag);  // :58
      if (std::isnan(br)) br = 0;
      if (std::isnan(tr)) tr = 0;
      Hit h;
      h.doc = doc;
      h.pr = sqd;
      h.final_rank = (0.33 * sqd + 0.38 * tr + 0.29 * br) * 100.0;  // :69
      hits.push_back(h);
    }
    // appendSort + truncation: util.go:48-54, main_retrieve.go:99-103
    size_t keep = std::min<size_t>(k, hits.size());
    std::partial_sort(hits.begin(), hits.begin() + keep, hits.end(), hit_before);
    out_count[q] = (uint32_t)keep;
    for (uint32_t j = 0; j < k; ++j) {
      bool ok = j < keep;
      out_doc[q * k + j] = ok ? hits[j].doc : 0xFFFFFFFFu;
      out_final[q * k + j] = ok ? hits[j].final_rank : 0.0;
      out_pr[q * k + j] = ok ? hits[j].pr : 0.0;
    }
  }
  return 0;
}

// "Fair" CPU flavour of oracle_score_batch (the same arithmetic on dense ids with the
// data structures a CPU implementation would choose, all cores): per-thread dense accumulators instead of
// a hash map per query, the blend term of a shared topic vector computed once per batch, a bounded
// selection instead of sorting every hit.  Sums are formed in the same order (query tokens, body then
// title postings, phrase last), so the results are the oracle's bit for bit.
extern "C" int oracle_score_batch_fair(const oracle_table* title, const oracle_table* body, uint64_t n_docs,
                                       const double* mag_title, const double* mag_body, const double* pagerank,
                                       uint32_t n_topics, uint64_t n_q, const uint64_t* kw_ptr,
                                       const uint32_t* kw_terms, const uint64_t* ph_ptr, const uint32_t* ph_terms,
                                       const double* topic_probs, int probs_per_query, uint32_t k, uint32_t* out_doc,
                                       double* out_final, double* out_pr, uint32_t* out_count, int n_threads) {
  if (n_threads <= 0) n_threads = omp_get_max_threads();
  std::vector<double> sqd_shared;
  if (topic_probs && pagerank && !probs_per_query) {
    sqd_shared.resize(n_docs);
#pragma omp parallel for num_threads(n_threads) schedule(static)
    for (uint64_t d = 0; d < n_docs; ++d) {
      double s = 0.0;
      for (uint32_t t = 0; t < n_topics; ++t) s += topic_probs[t] * pagerank[d * n_topics + t];
      sqd_shared[d] = s;
    }
  }
#pragma omp parallel num_threads(n_threads)
  {
    std::vector<double> acc_t(n_docs, 0.0), acc_b(n_docs, 0.0);
    std::vector<uint8_t> seen(n_docs, 0);
    std::vector<uint32_t> touched;
    std::vector<Hit> best;
#pragma omp for schedule(dynamic, 1)
    for (uint64_t q = 0; q < n_q; ++q) {
      const uint64_t kb = kw_ptr[q], ke = kw_ptr[q + 1];
      const uint64_t pb = ph_ptr ? ph_ptr[q] : 0, pe = ph_ptr ? ph_ptr[q + 1] : 0;
      touched.clear();
      auto touch = [&](uint32_t d) {
        if (!seen[d]) {
          seen[d] = 1;
          touched.push_back(d);
        }
      };
      for (uint64_t i = kb; i < ke; ++i) {
        uint64_t b, e;
        table_range(body, kw_terms[i], &b, &e);
        for (uint64_t p = b; p < e; ++p) {
          const uint32_t d = body->doc_ids[p];
          touch(d);
          acc_b[d] += (double)body->w[p];
        }
        table_range(title, kw_terms[i], &b, &e);
        for (uint64_t p = b; p < e; ++p) {
          const uint32_t d = title->doc_ids[p];
          touch(d);
          acc_t[d] += (double)title->w[p];
        }
      }
      if (pe > pb) {
        std::unordered_map<uint32_t, std::pair<bool, float>> pt, pbod;
        eval_phrase(title, body, ph_terms + pb, pe - pb, pt, pbod);
        for (auto& kv : pbod) {
          touch(kv.first);
          acc_b[kv.first] += (double)kv.second.second;
        }
        for (auto& kv : pt) {
          touch(kv.first);
          acc_t[kv.first] += (double)kv.second.second;
        }
      }
      const double* probs = topic_probs ? topic_probs + (probs_per_query ? q * n_topics : 0) : nullptr;
      const double qmag = std::sqrt((double)((ke - kb) + (pe - pb)));
      // bounded selection: keep the k best seen so far, worst of them at the heap's top
      best.clear();
      auto worse_first = [](const Hit& a, const Hit& b) { return hit_before(a, b); };  // heap top = last in order
      for (uint32_t doc : touched) {
        double sqd = 0.0;
        if (probs && pagerank) {
          if (!probs_per_query) sqd = sqd_shared[doc];
          else
            for (uint32_t t = 0; t < n_topics; ++t) sqd += probs[t] * pagerank[(uint64_t)doc * n_topics + t];
        }
        double br = acc_b[doc] / (mag_body[doc] * qmag);
        double tr = acc_t[doc] / (mag_title[doc] * qmag);
        if (std::isnan(br)) br = 0;
        if (std::isnan(tr)) tr = 0;
        acc_b[doc] = 0.0;
        acc_t[doc] = 0.0;
        seen[doc] = 0;
        Hit h;
        h.doc = doc;
        h.pr = sqd;
        h.final_rank = (0.33 * sqd + 0.38 * tr + 0.29 * br) * 100.0;
        if (best.size() < k) {
          best.push_back(h);
          std::push_heap(best.begin(), best.end(), worse_first);
        } else if (k && hit_before(h, best.front())) {
          std::pop_heap(best.begin(), best.end(), worse_first);
          best.back() = h;
          std::push_heap(best.begin(), best.end(), worse_first);
        }
      }
      std::sort(best.begin(), best.end(), hit_before);
      const size_t keep = best.size();
      out_count[q] = (uint32_t)keep;
      for (uint32_t j = 0; j < k; ++j) {
        const bool ok = j < keep;
        out_doc[q * k + j] = ok ? best[j].doc : 0xFFFFFFFFu;
        out_final[q * k + j] = ok ? best[j].final_rank : 0.0;
        out_pr[q * k + j] = ok ? best[j].pr : 0.0;
      }
    }
  }
  return 0;
}

// ---- extension: live topic probabilities -------------------------------------------------
// computeTopicProbs (retrieval/main_retrieve.go:106-159) restated with its defects repaired: `probs` starts at
// 1 (as shipped it starts at 0 and stays 0, :142-145); inv[2] rows are found by the word's own id.  Everything
// else as written: only tokens whose row lists the topic contribute a factor freq / wordCount[topic] (:116-134,
// :144), token order, a topic nobody lists gets 0 (:149-151), uniform prior 1 / #topics applied last (:147).
extern "C" int oracle_topic_probs(uint64_t n_terms, uint32_t n_topics, const uint64_t* term_ptr,
                                  const uint32_t* topic_ids, const double* freq, const double* word_count,
                                  uint64_t n_q, const uint64_t* tok_ptr, const uint32_t* tok_terms,
                                  double* out_probs) {
  for (uint64_t q = 0; q < n_q; ++q)
    for (uint32_t t = 0; t < n_topics; ++t) {
      double probs = 1.0;
      bool any = false;
      for (uint64_t k = tok_ptr[q]; k < tok_ptr[q + 1]; ++k) {
        const uint32_t term = tok_terms[k];
        if (term >= n_terms) continue;
        for (uint64_t x = term_ptr[term]; x < term_ptr[term + 1]; ++x)
          if (topic_ids[x] == t) {
            probs *= (freq[x] / word_count[t]);
            any = true;
            break;
          }
      }
      out_probs[q * n_topics + t] = any ? probs / (double)n_topics : 0.0;
    }
  return 0;
}
