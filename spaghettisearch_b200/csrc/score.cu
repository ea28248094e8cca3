// HP-2 online: the score / blend / top-k core of retrieval.Retrieve
// (retrieval/main_retrieve.go:15-104, get_metadata.go:16-77, phrase.go:11-170,
// util.go:48-54,179-203) for a whole query batch.
//
// One CTA scores one (query, group of doc slabs) pair.  Every path does the same two things: find
// the docs whose score could reach the running k-th best with cheap, conservative fp32/fp16 bounds,
// then evaluate those docs exactly (finish_exact: fp64 sums of the fp32 weights in query-token
// order, cosine, NaN -> 0, PageRank blend, get_metadata.go:53-69).  Paths, by (query, range):
//   owner_path  few postings, no phrase: lists staged in shared memory, lookups between the sorted
//               lists, the first list holding a doc folds its weights in list order;
//   dtiv_path   keyword query with a dense term: the term's contribution to the bound is one fp16
//               impact per doc, streamed; sparse tokens staged or scattered; survivors looked up;
//   accumulator path (in k_score): sub-ranges of kRange docs with two fp64 accumulators in shared
//               memory, tokens applied in query order, phrase (phrase.go) last, matched-doc list;
//   sort_path   sparse ranges of phrase queries (bitonic sort of tagged postings).
// k_narrow finds every list's offsets at the slab boundaries once per batch, k_plan groups the slabs
// per query, CTAs are ordered slab-major (co-resident CTAs read the same slice of the index), a running
// per-query bound is shared across slabs, k_merge merges the per-slab lists; the same kernel merges
// per-shard lists for the doc-sharded multi-GPU path.  DESIGN.md section 5 has the reasoning and numbers.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "index.cuh"

namespace {

constexpr int kT = 256;        // threads per CTA
constexpr int kRange = 2048;   // docs per sub-range
constexpr int kCand = 1024;    // candidate buffer = doc slots finished per round
constexpr int kMaxK = 128;
constexpr int kMaxKw = 64;     // keyword tokens per query handled in-kernel
constexpr int kMaxPh = 32;     // phrase tokens per query handled in-kernel
constexpr int kMaxLists = 2 * (kMaxKw + kMaxPh);
constexpr int kBounds = 4096;   // sub-range boundary table entries per pass
constexpr uint32_t kNoDoc = 0xFFFFFFFFu;
constexpr int kDenseRange = 4096;  // docs per sub-range of the impact-vector path
constexpr int kMaxDense = 254;     // dense slots (uint8 map, 255 = none)
constexpr int kDensePadDocs = 2 * kDenseRange;

struct TableView {
  const uint64_t* term_ptr;
  const uint32_t* doc_ids;
  const float* w;
  const uint64_t* pos_ptr;
  const float* pos;
  uint64_t V;
};

struct ScoreParams {
  TableView tab[2];      // 0 = title, 1 = body
  const double* mag[2];
  const float4* meta32;  // [D] {1/mag_title, 1/mag_body, blend bound, 0} in fp32 for the screening pass
  int prefetch_meta;       // issue L1 prefetches of the screening records per sub-range
  const uint32_t* narrow;  // [list][n_slabs+1] posting offset (from the term's row start) of every slab boundary
  uint32_t sort_max;     // slabs with at most this many postings take the sort path (<= kSortMax)
  int owner_path;        // phrase-free sparse slabs: lookups in the sorted lists instead of a sort
  int phrase_dense;      // phrase queries may take the impact-vector path (SS_SCORE_PHRASE_DENSE)
  const double* sqd;     // [D] blend term for a shared topic vector, or NULL
  const double* pr;      // [D][T] for per-query topic vectors
  const double* probs;   // [n_q][T] when per-query
  uint32_t T;
  uint64_t D;
  const uint64_t* kw_ptr;
  const uint32_t* kw_terms;
  const uint64_t* ph_ptr;
  const uint32_t* ph_terms;
  uint32_t n_q, n_slabs, k;
  uint64_t slab_docs;    // multiple of kRange
  uint32_t* part_doc;    // [n_q][n_slabs][k]
  double* part_final;
  double* part_pr;
  uint32_t* part_count;  // [n_q][n_slabs]
  unsigned long long* stats;  // [0] postings scanned, [1] docs matched
  unsigned long long* qthr;   // [n_q] running per-query bound (score key), zeroed per batch
  int use_qthr;
  // impact vectors of the densest terms (see IndexState)
  const uint16_t* uvec;       // [n_dense][d_pad] fp16 bits, NULL = path disabled
  const uint16_t* zvec;       // [d_pad]
  const float* zblk;          // [d_pad / kRange] largest blend term of a doc block
  const uint8_t* dense_map;   // [dense_map_V]
  uint64_t d_pad, dense_map_V;
  // slab groups (k_plan): group_len[q][slab] = number of consecutive slabs the CTA of (q, slab) scores as
  // one range, 0 = the slab belongs to an earlier CTA's group (that CTA exits at once)
  const uint8_t* group_len;
};

// Total order of results: FinalRank descending, ties by ascending doc id, NaN
// last (util.go:48-54 with the arrival-order tie pinned).  Scores map to
// unsigned keys whose integer order is the score order.
__device__ __forceinline__ uint64_t score_key(double f) {
  if (isnan(f)) return 0ull;
  if (f == 0.0) f = 0.0;  // -0 == +0
  const uint64_t b = (uint64_t)__double_as_longlong(f);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_score(uint64_t key) {
  if (key == 0ull) return __longlong_as_double(0x7FF8000000000000ll);
  const uint64_t b = (key >> 63) ? (key & 0x7FFFFFFFFFFFFFFFull) : ~key;
  return __longlong_as_double((long long)b);
}
__device__ __forceinline__ bool beats(uint64_t ka, uint32_t da, uint64_t kb, uint32_t db) {
  return ka > kb || (ka == kb && da < db);
}

// first index in [lo, hi) with docs[i] >= target
__device__ __forceinline__ uint64_t lower_bound_doc(const uint32_t* __restrict__ docs, uint64_t lo, uint64_t hi,
                                                    uint64_t target) {
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (docs[mid] < target) lo = mid + 1; else hi = mid;
  }
  return lo;
}

struct Smem {
  double acc[2][kRange];        // [0] TitleRank, [1] BodyRank sums of the sub-range
  unsigned long long cand_key[kCand];
  unsigned long long top_key[2][kMaxK];
  unsigned long long cur[kMaxLists], hi[kMaxLists];  // current sub-range of every list
  unsigned long long base[kMaxLists];               // start of the list inside the slab
  uint32_t len[kMaxLists];                          // postings of the list inside the slab
  uint32_t bounds[kBounds];                         // [list][sub-range] offsets from base
  uint32_t cand_doc[kCand];
  uint32_t top_doc[2][kMaxK];
  uint32_t bits[kRange / 32];
  uint32_t dbits[kDenseRange / 32];  // impact-vector path: docs with a sparse-token posting
  uint16_t mlist[kRange];  // slots of the sub-range's matched docs, in first-touch order
  uint32_t n_cand, n_ent, n_list, top_n, top_buf;
  unsigned long long thr_key, piv_key;
  uint32_t thr_doc, piv_doc;
  float thr_f;  // fp32 lower bound of the score a doc needs: max(local k-th best, the query's running bound)
  unsigned long long gkey;  // the query's running bound (score key of some slab's k-th best), 0 = none
  float gthr_f;             // its score rounded down to fp32 (-inf when none)
  uint8_t tok_dense[kMaxKw + kMaxPh];   // dense slot of every token, keyword then phrase (255 = sparse)
  uint8_t dense_slots[kMaxKw + kMaxPh]; // the dense tokens' slots, in token order
  uint8_t sparse_toks[kMaxKw + kMaxPh]; // indices of the sparse tokens
  uint32_t n_dense_tok, n_sparse_tok, n_dense_kw;
  float zred[kT / 32];
};

// Union of the running top-k and the candidate buffer -> new running top-k by
// counting, for each element, how many others beat it (elements are distinct
// in (key, doc), so ranks are a permutation).  A large candidate set is first
// thinned with a pivot: the k-th best of a kSample-element sample is beaten by
// at most k-1 sample members, so every candidate the pivot beats is outside
// the top k and can be dropped without changing the result.
constexpr uint32_t kSample = 128;
__device__ void merge_candidates(Smem& s, uint32_t k) {
  if (s.n_cand > 2 * kSample && 4 * k <= kSample) {
    const uint32_t nc0 = s.n_cand;
    if (threadIdx.x < kSample) {
      const unsigned long long ki = s.cand_key[threadIdx.x];
      const uint32_t di = s.cand_doc[threadIdx.x];
      uint32_t rank = 0;
      for (uint32_t j = 0; j < kSample; ++j) rank += beats(s.cand_key[j], s.cand_doc[j], ki, di) ? 1u : 0u;
      if (rank == k - 1) {
        s.piv_key = ki;
        s.piv_doc = di;
      }
    }
    // survivors are re-packed: read into registers, barrier, write
    constexpr int kPerThread = kCand / kT;
    unsigned long long rk[kPerThread];
    uint32_t rd[kPerThread];
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      const uint32_t i = threadIdx.x + c * kT;
      rk[c] = i < nc0 ? s.cand_key[i] : 0ull;
      rd[c] = i < nc0 ? s.cand_doc[i] : kNoDoc;
    }
    __syncthreads();
    if (threadIdx.x == 0) s.n_cand = 0;
    __syncthreads();
#pragma unroll
    for (int c = 0; c < kPerThread; ++c) {
      const uint32_t i = threadIdx.x + c * kT;
      if (i < nc0 && !beats(s.piv_key, s.piv_doc, rk[c], rd[c])) {
        const uint32_t j = atomicAdd(&s.n_cand, 1u);
        s.cand_key[j] = rk[c];
        s.cand_doc[j] = rd[c];
      }
    }
    __syncthreads();
  }
  const uint32_t nt = s.top_n, nc = s.n_cand, n = nt + nc;
  const uint32_t ob = s.top_buf, nb = ob ^ 1;
  for (uint32_t i = threadIdx.x; i < n; i += kT) {
    const unsigned long long ki = i < nt ? s.top_key[ob][i] : s.cand_key[i - nt];
    const uint32_t di = i < nt ? s.top_doc[ob][i] : s.cand_doc[i - nt];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nt; ++j) rank += beats(s.top_key[ob][j], s.top_doc[ob][j], ki, di) ? 1u : 0u;
    for (uint32_t j = 0; j < nc; ++j) rank += beats(s.cand_key[j], s.cand_doc[j], ki, di) ? 1u : 0u;
    if (rank < k) {
      s.top_key[nb][rank] = ki;
      s.top_doc[nb][rank] = di;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    s.top_n = min(k, n);
    s.top_buf = nb;
    s.n_cand = 0;
    if (s.top_n == k) {
      s.thr_key = s.top_key[nb][k - 1];
      s.thr_doc = s.top_doc[nb][k - 1];
      const double thr = key_score(s.thr_key);
      // NaN as k-th best (only NaN scores so far) must not filter anything
      s.thr_f = isnan(thr) ? s.gthr_f : fmaxf(s.gthr_f, __double2float_rd(thr));
    }
  }
  __syncthreads();
}

// Sort-path entry: doc offset in the slab (24 bits) | list sequence (8) | weight bits (32).
// Sorting the 64-bit keys groups a doc's weights in query-token order.
constexpr uint32_t kSortMax = 2 * kRange;  // entries; the buffer aliases the two accumulator arrays
__device__ __forceinline__ unsigned long long make_entry(uint32_t off, uint32_t seq, float w) {
  return ((unsigned long long)off << 40) | ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(w);
}

// First posting to touch a doc slot of the sub-range appends it to the matched list, so that the
// finalize step costs what the matches cost (the first version walked the whole bitmap: ~760
// warp instructions per warp and sub-range whatever the density, the dominant cost for all but
// the densest lists).
__device__ __forceinline__ void mark_matched(Smem& s, uint32_t slot) {
  const uint32_t bit = 1u << (slot & 31);
  const uint32_t old = atomicOr(&s.bits[slot >> 5], bit);
  if (!(old & bit)) s.mlist[atomicAdd(&s.n_list, 1u) & (kRange - 1)] = (uint16_t)slot;
}

// phrase.go:53-109 for one table over the lists' current ranges [cur, hi).  Lists
// l0+2*i+tb hold token i.  A doc gets ONE weight = fp32 sum of the tokens' weights
// in phrase order iff every token has a posting in this table and some position a
// of token 0 has a + i among token i's positions (compared as
// (pos_i - float32(i)) == pos_0, phrase.go:144-146, util.go:185).
// EMIT = false: add into the sub-range accumulators; EMIT = true: append a sort entry.
template <int EMIT>
__device__ void apply_phrase(const ScoreParams& p, Smem& s, int tb, uint32_t l0, uint32_t L, uint64_t d0,
                             uint32_t seq, unsigned long long& n_postings) {
  const TableView& tv = p.tab[tb];
  uint32_t drv = 0;
  unsigned long long best = ~0ull;
  for (uint32_t i = 0; i < L; ++i) {
    const uint32_t l = l0 + 2 * i + tb;
    const unsigned long long len = s.hi[l] - s.cur[l];
    if (len == 0) return;  // a token without postings here: nothing can match (uniform across the CTA)
    if (len < best) {
      best = len;
      drv = i;
    }
  }
  const uint32_t ld = l0 + 2 * drv + tb;
  for (unsigned long long q = s.cur[ld] + threadIdx.x; q < s.hi[ld]; q += kT) {
    const uint32_t doc = tv.doc_ids[q];
    unsigned long long pi[kMaxPh];
    bool all = true;
    for (uint32_t i = 0; i < L && all; ++i) {
      if (i == drv) {
        pi[i] = q;
        continue;
      }
      const uint32_t l = l0 + 2 * i + tb;
      unsigned long long a = s.cur[l], b = s.hi[l];
      while (a < b) {
        const unsigned long long mid = (a + b) >> 1;
        if (tv.doc_ids[mid] < doc) a = mid + 1; else b = mid;
      }
      if (a < s.hi[l] && tv.doc_ids[a] == doc) pi[i] = a; else all = false;
    }
    n_postings += L;
    if (!all) continue;
    bool hit = false;
    if (tv.pos_ptr) {
      const unsigned long long a0 = tv.pos_ptr[pi[0]], a1 = tv.pos_ptr[pi[0] + 1];
      for (unsigned long long x = a0; x < a1 && !hit; ++x) {
        const float a = __fadd_rn(tv.pos[x], -0.0f);
        bool ok = true;
        for (uint32_t i = 1; i < L && ok; ++i) {
          const float shift = (float)(uint8_t)i;
          bool found = false;
          for (unsigned long long y = tv.pos_ptr[pi[i]]; y < tv.pos_ptr[pi[i] + 1] && !found; ++y)
            found = __fadd_rn(tv.pos[y], -shift) == a;
          ok = found;
        }
        hit = ok;
      }
    }
    if (!hit) continue;
    float sum = 0.0f;  // phrase.go:59,69,83: float32 running sum in token order
    for (uint32_t i = 0; i < L; ++i) sum = __fadd_rn(sum, tv.w[pi[i]]);
    const uint32_t slot = (uint32_t)(doc - d0);
    if (EMIT == 1) {
      unsigned long long* ent = reinterpret_cast<unsigned long long*>(&s.acc[0][0]);
      ent[atomicAdd(&s.n_ent, 1u)] = make_entry(slot, seq, sum);
    } else if (EMIT == 2) {  // lookup path: (doc offset, weight) appended from position `seq` of its arrays
      const uint32_t at = seq + atomicAdd(&s.n_ent, 1u);
      reinterpret_cast<uint32_t*>(&s.acc[0][0])[at] = slot;
      reinterpret_cast<float*>(&s.acc[1][0])[at] = sum;
    } else {
      s.acc[tb][slot] = __dadd_rn(s.acc[tb][slot], (double)sum);
      mark_matched(s, slot);
    }
  }
}

// Exact per-doc inputs of the final rank (only docs that survive the screening need them).
struct DocMeta {
  double mag_t, mag_b, sqd;
};
__device__ __forceinline__ DocMeta load_meta(const ScoreParams& p, uint32_t q, uint64_t doc) {
  DocMeta m;
  m.mag_b = p.mag[1][doc];
  m.mag_t = p.mag[0][doc];
  m.sqd = 0.0;  // get_metadata.go:39-42
  if (p.sqd) {
    m.sqd = p.sqd[doc];
  } else if (p.probs) {
    const double* pr = p.pr + doc * p.T;
    const double* pq = p.probs + (uint64_t)q * p.T;
    for (uint32_t t = 0; t < p.T; ++t) m.sqd = __dadd_rn(m.sqd, __dmul_rn(pq[t], pr[t]));
  }
  return m;
}

// cosine, NaN -> 0, PageRank blend (get_metadata.go:53-69); a doc that can still
// make the top k goes to the candidate buffer.
//
// Screening first: once k results exist, a doc whose score -- bounded from one packed fp32
// record (reciprocal norms, blend bound) with a margin far above fp32 rounding -- stays below
// the running k-th best cannot enter the top k, so its exact fp64 inputs are never fetched.
// blend_scale = 1 for a shared topic vector (the record holds sqd itself) or sum |p_t| for a
// per-query vector (the record holds max_t |PR[doc][t]|).  NaN/Inf fall through to the exact path.
__device__ __forceinline__ bool screened_out(const Smem& s, float tr, float br, const float4& m32, float qf_inv,
                                             float blend_scale) {
  const float a = 0.33f * m32.z * blend_scale;
  const float b = tr != 0.0f ? 0.38f * (tr * m32.x * qf_inv) : 0.0f;
  const float c = br != 0.0f ? 0.29f * (br * m32.y * qf_inv) : 0.0f;
  const float approx = (a + b + c) * 100.0f;
  const float slack = (fabsf(a) + fabsf(b) + fabsf(c)) * 1e-2f + 1e-30f;  // 1e-4 relative, x100
  return approx + slack < s.thr_f;  // false for NaN: those go to the exact path
}
// exact final rank from the fp64 sums (get_metadata.go:53-69) and the candidate test
__device__ __forceinline__ void finish_exact(const ScoreParams& p, Smem& s, uint32_t q, uint64_t doc, double tr,
                                             double br, double qm, uint32_t k) {
  const DocMeta m = load_meta(p, q, doc);
  // get_metadata.go:57-66.  x/y with x == 0 is 0 or NaN, and NaN becomes 0: skip the divide
  double body = 0.0, title = 0.0;
  if (br != 0.0) {
    body = __ddiv_rn(br, __dmul_rn(m.mag_b, qm));
    if (isnan(body)) body = 0.0;
  }
  if (tr != 0.0) {
    title = __ddiv_rn(tr, __dmul_rn(m.mag_t, qm));
    if (isnan(title)) title = 0.0;
  }
  // (0.33*sqd + 0.38*Title + 0.29*Body) * 100.0, left to right, no fusing (:69)
  const double fin =
      __dmul_rn(__dadd_rn(__dadd_rn(__dmul_rn(0.33, m.sqd), __dmul_rn(0.38, title)), __dmul_rn(0.29, body)), 100.0);
  const unsigned long long key = score_key(fin);
  if (key < s.gkey) return;  // below another slab's k-th best: cannot be in the query's top k
  if (s.top_n < k || beats(key, (uint32_t)doc, s.thr_key, s.thr_doc)) {
    const uint32_t j = atomicAdd(&s.n_cand, 1u);
    s.cand_key[j] = key;
    s.cand_doc[j] = (uint32_t)doc;
  }
}
__device__ __forceinline__ void finish_doc(const ScoreParams& p, Smem& s, uint32_t q, uint64_t doc, double tr,
                                           double br, const float4& m32, float qf_inv, float blend_scale,
                                           double qm, uint32_t k) {
  if (screened_out(s, (float)tr, (float)br, m32, qf_inv, blend_scale)) return;
  finish_exact(p, s, q, doc, tr, br, qm, k);
}

// One list's postings of the sub-range into the accumulators; a list holds a doc at
// most once, so plain read-modify-write is race free.  Four postings per thread are
// fetched before the first is applied.
__device__ __forceinline__ void accumulate_list(const TableView& tv, Smem& s, int tb, unsigned long long x0,
                                                unsigned long long x1, uint64_t d0) {
  constexpr int U = 4;
  for (unsigned long long x = x0 + threadIdx.x; x < x1; x += (unsigned long long)U * kT) {
    uint32_t d[U];
    float w[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const unsigned long long xx = x + (unsigned long long)u * kT;
      const bool ok = xx < x1;
      d[u] = ok ? tv.doc_ids[xx] : 0xFFFFFFFFu;
      w[u] = ok ? tv.w[xx] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      if (d[u] == 0xFFFFFFFFu) continue;
      const uint32_t slot = (uint32_t)(d[u] - d0);
      s.acc[tb][slot] = __dadd_rn(s.acc[tb][slot], (double)w[u]);
      mark_matched(s, slot);
    }
  }
}

// Sparse (query, slab) pairs: all postings of the slab fit in shared memory.  They are
// tagged (doc offset, list sequence, weight), sorted, and every doc's run is folded in
// sequence order -- the same sums as the dense path without touching empty sub-ranges.
__device__ void sort_path(const ScoreParams& p, Smem& s, uint32_t q, uint64_t slab_lo, uint32_t n_kw, uint32_t n_ph,
                          double qm, float qf_inv, float blend_scale, uint32_t k, unsigned long long& n_postings,
                          unsigned long long& n_matched) {
  unsigned long long* ent = reinterpret_cast<unsigned long long*>(&s.acc[0][0]);
  const uint32_t tid = threadIdx.x, n_kw_lists = 2 * n_kw;
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t l = 0; l < n_kw_lists; ++l) {
      s.bounds[l] = run;
      run += s.len[l];
    }
    s.bounds[n_kw_lists] = run;
    s.n_ent = run;
  }
  for (uint32_t l = n_kw_lists + tid; l < n_kw_lists + 2 * n_ph; l += kT) {  // phrase lists: whole slab range
    s.cur[l] = s.base[l];
    s.hi[l] = s.base[l] + s.len[l];
  }
  __syncthreads();
  const uint32_t total_kw = s.bounds[n_kw_lists];
  for (uint32_t idx = tid; idx < total_kw; idx += kT) {
    uint32_t lo = 0, hi = n_kw_lists;  // last l with bounds[l] <= idx
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s.bounds[mid] <= idx) lo = mid; else hi = mid;
    }
    const TableView& tv = p.tab[lo & 1];
    const unsigned long long at = s.base[lo] + (idx - s.bounds[lo]);
    ent[idx] = make_entry((uint32_t)(tv.doc_ids[at] - slab_lo), lo, tv.w[at]);
  }
  if (tid == 0) n_postings += total_kw;
  if (n_ph) {
    apply_phrase<1>(p, s, 1, n_kw_lists, n_ph, slab_lo, n_kw_lists + 1, n_postings);
    apply_phrase<1>(p, s, 0, n_kw_lists, n_ph, slab_lo, n_kw_lists, n_postings);
  }
  __syncthreads();
  const uint32_t n = s.n_ent;
  uint32_t n2 = 2;
  while (n2 < n) n2 <<= 1;
  for (uint32_t i = n + tid; i < n2; i += kT) ent[i] = ~0ull;
  __syncthreads();
  for (uint32_t kk = 2; kk <= n2; kk <<= 1) {
    for (uint32_t j = kk >> 1; j > 0; j >>= 1) {
      for (uint32_t i = tid; i < n2; i += kT) {
        const uint32_t x = i ^ j;
        if (x > i) {
          const unsigned long long a = ent[i], b = ent[x];
          if ((a > b) == ((i & kk) == 0)) {
            ent[i] = b;
            ent[x] = a;
          }
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t r0 = 0; r0 < n; r0 += kCand) {
    const uint32_t r1 = min(n, r0 + kCand);
    for (uint32_t i = r0 + tid; i < r1; i += kT) {
      const unsigned long long e0 = ent[i];
      const uint32_t off = (uint32_t)(e0 >> 40);
      if (i > 0 && (uint32_t)(ent[i - 1] >> 40) == off) continue;  // not the head of a doc's run
      double tr = 0.0, br = 0.0;
      for (uint32_t j = i; j < n; ++j) {
        const unsigned long long e = ent[j];
        if ((uint32_t)(e >> 40) != off) break;
        const double w = (double)__uint_as_float((uint32_t)e);
        if ((e >> 32) & 1ull) br = __dadd_rn(br, w); else tr = __dadd_rn(tr, w);
      }
      ++n_matched;
      finish_doc(p, s, q, slab_lo + off, tr, br, p.meta32[slab_lo + off], qf_inv, blend_scale, qm, k);
    }
    __syncthreads();
    if (s.n_cand) merge_candidates(s, k);
  }
}

// Sparse (query, slab) pairs without a phrase: the keyword lists are staged in shared memory as
// they are -- every list is already sorted by doc -- and each posting looks its doc up in the
// other lists by binary search.  The posting of the FIRST list that holds the doc owns it: it
// folds the doc's weights in list (= query-token) order, which is the order of the reference's
// sums, and finishes the doc.  No sort, no barriers inside the loop: the bitonic sort this
// replaces cost ~100 ps per posting against ~10 ps in the dense path (78 block barriers per
// 4096 entries), and 44 % of the benchmark's query tokens take this path.
__device__ void owner_path(const ScoreParams& p, Smem& s, uint32_t q, uint64_t slab_lo, uint32_t n_kw, uint32_t n_ph,
                           double qm, float qf_inv, float blend_scale, uint32_t k, unsigned long long& n_postings,
                           unsigned long long& n_matched) {
  uint32_t* soff = reinterpret_cast<uint32_t*>(&s.acc[0][0]);  // [kSortMax] doc offset in the slab
  float* sw = reinterpret_cast<float*>(&s.acc[1][0]);          // [kSortMax] weight
  const uint32_t tid = threadIdx.x, n_kw_lists = 2 * n_kw;
  // With a phrase, its hits form two more "lists" behind the keyword lists: title hits (even index, so
  // their weight joins TitleRank) then body hits (odd), each weight appended after the keyword weights as
  // in main_retrieve.go:73-78.  They fit: a table's hits are at most its shortest phrase list, which is
  // what `work` counted.
  const uint32_t n_lists = n_kw_lists + (n_ph ? 2u : 0u);
  if (tid == 0) {
    uint32_t run = 0;
    for (uint32_t l = 0; l < n_kw_lists; ++l) {
      s.bounds[l] = run;
      run += s.len[l];
    }
    s.bounds[n_kw_lists] = run;
    s.n_ent = 0;
  }
  for (uint32_t l = n_kw_lists + tid; l < n_kw_lists + 2 * n_ph; l += kT) {  // phrase lists: whole range
    s.cur[l] = s.base[l];
    s.hi[l] = s.base[l] + s.len[l];
  }
  __syncthreads();
  const uint32_t kw_total = s.bounds[n_kw_lists];
  // stage: list by list, coalesced
  for (uint32_t l = 0; l < n_kw_lists; ++l) {
    const uint32_t len = s.len[l];
    if (!len) continue;
    const TableView& tv = p.tab[l & 1];
    const unsigned long long at = s.base[l];
    const uint32_t o = s.bounds[l];
    for (uint32_t i = tid; i < len; i += kT) {
      soff[o + i] = (uint32_t)(tv.doc_ids[at + i] - slab_lo);
      sw[o + i] = tv.w[at + i];
    }
  }
  if (tid == 0) n_postings += kw_total;
  if (n_ph) {
    // phrase hits of the title table, then of the body table, each sorted by doc (ranking by counting:
    // a table holds a doc once, so the ranks are a permutation)
    for (int tb = 0; tb < 2; ++tb) {
      const uint32_t start = tb == 0 ? kw_total : s.bounds[n_kw_lists + 1];
      apply_phrase<2>(p, s, tb, n_kw_lists, n_ph, slab_lo, start, n_postings);
      __syncthreads();
      const uint32_t cnt = s.n_ent;
      uint32_t my_off[kSortMax / kT], my_rank[kSortMax / kT];
      float my_w[kSortMax / kT];
#pragma unroll 1
      for (uint32_t c = 0, i = tid; i < cnt; i += kT, ++c) {
        my_off[c] = soff[start + i];
        my_w[c] = sw[start + i];
        uint32_t r = 0;
        for (uint32_t j = 0; j < cnt; ++j) r += soff[start + j] < my_off[c] ? 1u : 0u;
        my_rank[c] = r;
      }
      __syncthreads();
#pragma unroll 1
      for (uint32_t c = 0, i = tid; i < cnt; i += kT, ++c) {
        soff[start + my_rank[c]] = my_off[c];
        sw[start + my_rank[c]] = my_w[c];
      }
      if (tid == 0) {
        s.bounds[n_kw_lists + tb + 1] = start + cnt;
        s.n_ent = 0;
      }
      __syncthreads();
    }
  } else {
    __syncthreads();
  }
  const uint32_t total = s.bounds[n_lists];
  // position of doc offset `off` in list j, or kNoDoc
  auto find = [&](uint32_t j, uint32_t off) -> uint32_t {
    uint32_t lo = s.bounds[j], hi = s.bounds[j + 1];
    const uint32_t end = hi;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (soff[mid] < off) lo = mid + 1; else hi = mid;
    }
    return (lo < end && soff[lo] == off) ? lo : kNoDoc;
  };
  for (uint32_t r0 = 0; r0 < total; r0 += kCand) {
    const uint32_t r1 = min(total, r0 + kCand);
    for (uint32_t idx = r0 + tid; idx < r1; idx += kT) {
      uint32_t lo = 0, hi = n_lists;  // last l with bounds[l] <= idx
      while (hi - lo > 1) {
        const uint32_t mid = (lo + hi) >> 1;
        if (s.bounds[mid] <= idx) lo = mid; else hi = mid;
      }
      const uint32_t l = lo, off = soff[idx];
      bool owner = true;
      for (uint32_t j = 0; j < l && owner; ++j) owner = find(j, off) == kNoDoc;
      if (!owner) continue;
      double tr = 0.0, br = 0.0;
      {
        const double w = (double)sw[idx];
        if (l & 1u) br = w; else tr = w;
      }
      for (uint32_t j = l + 1; j < n_lists; ++j) {
        const uint32_t at = find(j, off);
        if (at == kNoDoc) continue;
        const double w = (double)sw[at];
        if (j & 1u) br = __dadd_rn(br, w); else tr = __dadd_rn(tr, w);
      }
      ++n_matched;
      finish_doc(p, s, q, slab_lo + off, tr, br, p.meta32[slab_lo + off], qf_inv, blend_scale, qm, k);
    }
    __syncthreads();
    if (s.n_cand) merge_candidates(s, k);
  }
}

// ---- impact-vector path ----------------------------------------------------------------------
// Keyword queries that contain one of the densest terms (>= 1/32 of the docs; 96 % of the
// benchmark's postings belong to ~220 such terms).  Walking a 5M-posting list costs 8 B and ~150
// thread instructions per posting in the accumulator path; here the term's whole contribution to
// the SCREENING bound of a doc is one precomputed fp16 value (2 B, coalesced, no doc ids, no
// atomics): U_t[d] >= 100 * (0.38 w_title/|title| + 0.29 w_body/|body|), rounded up, 0 = no
// posting.  A sub-range of kDenseRange docs is streamed: bound(d) = blend_scale * Z[d] +
// (sum of the dense tokens' U[d] + the sparse tokens' scattered fp32 impacts) / |q|, all terms
// non-negative and rounded up, so bound(d) >= FinalRank(d).  Only docs whose bound reaches the
// running threshold are evaluated exactly: their weights are looked up in the posting lists by
// binary search and folded in query-token order -- the same fp64 sums, cosine and blend as the
// other paths, so the top k is identical (tests/test_scoring_gpu.py compares the paths).
__device__ __forceinline__ float half_bits_to_float(uint32_t h) {
  return __half2float(__ushort_as_half((unsigned short)h));
}

// The doc's weight in list l of the query (slab range), if it has one: interpolation start (docs are
// spread over the slab), gallop to bracket the doc, bisect.
__device__ __forceinline__ bool find_index_in_list(const ScoreParams& p, const Smem& s, uint32_t l, uint32_t doc,
                                                   uint64_t slab_lo, uint64_t slab_docs, unsigned long long& at) {
  const uint32_t len = s.len[l];
  if (!len) return false;
  const uint32_t* __restrict__ docs = p.tab[l & 1].doc_ids + s.base[l];
  uint32_t pos = (uint32_t)min((uint64_t)len - 1, (uint64_t)len * (doc - slab_lo) / slab_docs);
  uint32_t lo, hi;  // invariant: docs[lo - 1] < doc (or lo == 0), docs[hi] >= doc (or hi == len)
  if (docs[pos] < doc) {
    lo = pos + 1;
    uint32_t step = 16;
    while (true) {
      hi = min(len, lo + step);
      if (hi == len || docs[hi] >= doc) break;
      lo = hi + 1;
      step <<= 1;
    }
  } else {
    hi = pos;
    uint32_t step = 16;
    while (true) {
      lo = hi > step ? hi - step : 0u;
      if (lo == 0 || docs[lo - 1] < doc) break;
      hi = lo - 1;
      step <<= 1;
    }
  }
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (docs[mid] < doc) lo = mid + 1; else hi = mid;
  }
  if (lo == len || docs[lo] != doc) return false;
  at = s.base[l] + lo;
  return true;
}
__device__ __forceinline__ bool find_in_list(const ScoreParams& p, const Smem& s, uint32_t l, uint32_t doc,
                                             uint64_t slab_lo, uint64_t slab_docs, float& w_out) {
  unsigned long long at;
  if (!find_index_in_list(p, s, l, doc, slab_lo, slab_docs, at)) return false;
  w_out = p.tab[l & 1].w[at];
  return true;
}

// phrase.go:53-109 for ONE doc and one table (the per-doc form of apply_phrase): lists l0+2*i+tb hold
// token i.  The doc gets one weight = fp32 sum of the tokens' weights in phrase order iff every token has
// a posting of the doc in this table and some position a of token 0 has a + i among token i's positions
// (compared as (pos_i - float32(i)) == pos_0, phrase.go:144-146, util.go:185).
__device__ bool phrase_hit(const ScoreParams& p, const Smem& s, int tb, uint32_t l0, uint32_t L, uint32_t doc,
                           uint64_t slab_lo, uint64_t slab_docs, float& sum_out) {
  const TableView& tv = p.tab[tb];
  if (!tv.pos_ptr) return false;
  unsigned long long pi[kMaxPh];
  for (uint32_t i = 0; i < L; ++i)
    if (!find_index_in_list(p, s, l0 + 2 * i + tb, doc, slab_lo, slab_docs, pi[i])) return false;
  bool hit = false;
  const unsigned long long a0 = tv.pos_ptr[pi[0]], a1 = tv.pos_ptr[pi[0] + 1];
  for (unsigned long long x = a0; x < a1 && !hit; ++x) {
    const float a = __fadd_rn(tv.pos[x], -0.0f);
    bool ok = true;
    for (uint32_t i = 1; i < L && ok; ++i) {
      const float shift = (float)(uint8_t)i;
      bool found = false;
      for (unsigned long long y = tv.pos_ptr[pi[i]]; y < tv.pos_ptr[pi[i] + 1] && !found; ++y)
        found = __fadd_rn(tv.pos[y], -shift) == a;
      ok = found;
    }
    hit = ok;
  }
  if (!hit) return false;
  float sum = 0.0f;  // phrase.go:59,69,83: float32 running sum in token order
  for (uint32_t i = 0; i < L; ++i) sum = __fadd_rn(sum, tv.w[pi[i]]);
  sum_out = sum;
  return true;
}

// Exact evaluation of n survivors (slab offsets in the ring from position r0): four lanes per doc look it
// up in four lists at a time, the weights are folded in list (= query-token) order -- the same fp64 sums
// as the accumulator paths -- and the group's first lane finishes the doc.  Every thread of the CTA calls.
__device__ __forceinline__ void evaluate_survivors(const ScoreParams& p, Smem& s, uint32_t q, const uint16_t* surv,
                                                   uint32_t ring_mask, uint32_t r0, uint32_t n, uint64_t rel_base,
                                                   uint32_t n_kw, uint32_t n_ph, uint64_t slab_lo, uint64_t slab_docs,
                                                   double qm, uint32_t k) {
  const uint32_t n_lists = 2 * n_kw;  // keyword lists; the phrase is evaluated after them
  const uint32_t sub = threadIdx.x >> 2, gl = threadIdx.x & 3;
  for (uint32_t base = 0; base < n; base += kT / 4) {
    const uint32_t i = base + sub;
    const bool active = i < n;
    const uint32_t doc = active ? (uint32_t)(rel_base + surv[(r0 + i) & ring_mask]) : 0u;
    double tr = 0.0, br = 0.0;
    bool first_t = true, first_b = true;
    for (uint32_t l0 = 0; l0 < n_lists; l0 += 4) {
      const uint32_t l = l0 + gl;
      float w = 0.0f;
      const bool found = active && l < n_lists && find_in_list(p, s, l, doc, slab_lo, slab_docs, w);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const bool fj = __shfl_sync(0xFFFFFFFFu, found ? 1 : 0, j, 4) != 0;
        const double wj = (double)__shfl_sync(0xFFFFFFFFu, w, j, 4);
        if (!fj) continue;
        if ((l0 + j) & 1u) {
          br = first_b ? wj : __dadd_rn(br, wj);
          first_b = false;
        } else {
          tr = first_t ? wj : __dadd_rn(tr, wj);
          first_t = false;
        }
      }
    }
    if (active && gl == 0) {
      bool matched = !(first_t && first_b);  // a keyword posting
      if (n_ph) {  // the phrase's weight is appended after the keyword weights (main_retrieve.go:73-78)
        float ws = 0.0f;
        if (phrase_hit(p, s, 1, 2 * n_kw, n_ph, doc, slab_lo, slab_docs, ws)) {
          br = first_b ? (double)ws : __dadd_rn(br, (double)ws);
          matched = true;
        }
        if (phrase_hit(p, s, 0, 2 * n_kw, n_ph, doc, slab_lo, slab_docs, ws)) {
          tr = first_t ? (double)ws : __dadd_rn(tr, (double)ws);
          matched = true;
        }
      }
      if (matched) finish_exact(p, s, q, doc, tr, br, qm, k);
    }
  }
}

// Whole-slab stream of the impact-vector path for keyword queries whose tokens in this slab are all
// dense (no scattered sparse impacts) once a threshold exists: no barriers inside, two 8-doc groups
// per thread and step with all loads issued first.  Survivors go to the ring as 16-bit slab offsets;
// returns false (uniformly, after a barrier) if the ring overflowed -- the caller then redoes the slab
// sub-range by sub-range.  ND = number of dense tokens when it is 1 or 2 (vector bases in registers),
// 0 = any.
template <int ND>
__device__ __forceinline__ bool dtiv_stream_slab(const ScoreParams& p, Smem& s, uint16_t* surv, uint32_t ring_mask,
                                                 uint32_t surv_done, uint64_t slab_lo, uint64_t slab_hi, uint32_t nd,
                                                 float za, float blend_scale, float qf_inv, float thr_f,
                                                 const uint32_t* excl, uint32_t& my_matched) {
  // excl: bitmap over the slab of docs that also have a sparse-token posting; those are scored by the
  // caller (their bound needs the sparse impacts), so they are masked out here.  NULL = none.
  const uint32_t tid = threadIdx.x;
  const uint32_t n_docs = (uint32_t)(slab_hi - slab_lo);
  const uint16_t* __restrict__ base[ND ? ND : 1];
#pragma unroll
  for (int i = 0; i < (ND ? ND : 1); ++i) base[i] = p.uvec + (size_t)s.dense_slots[i] * p.d_pad + slab_lo;
  const uint16_t* __restrict__ zv = p.zvec + slab_lo;
  bool ovf = false;
  auto process = [&](const float (&sum)[8], uint32_t off) {
    uint32_t present = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) present |= (sum[j] > 0.0f ? 1u : 0u) << j;  // impacts are >= 0, > 0 for a posting
    if (excl) present &= ~((excl[off >> 5] >> (off & 31)) & 0xFFu);
    if (off + 8 > n_docs) present &= (1u << (n_docs - off)) - 1u;
    if (!present) return;
    my_matched += __popc(present);
    const float gmax = fmaxf(fmaxf(fmaxf(sum[0], sum[1]), fmaxf(sum[2], sum[3])),
                             fmaxf(fmaxf(sum[4], sum[5]), fmaxf(sum[6], sum[7])));
    {
      const float b = qf_inv * gmax;
      if ((za + b) + (fabsf(za) + fabsf(b)) * 1e-4f + 1e-30f < thr_f) return;
    }
    const uint4 zz = __ldg(reinterpret_cast<const uint4*>(zv + off));
    const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (!((present >> j) & 1u)) continue;
      const float z = half_bits_to_float(j & 1 ? zw[j >> 1] >> 16 : zw[j >> 1] & 0xFFFFu);
      const float a = blend_scale * z, b = qf_inv * sum[j];
      if ((a + b) + (fabsf(a) + fabsf(b)) * 1e-4f + 1e-30f < thr_f) continue;  // NaN and +inf stay in
      const uint32_t pos = atomicAdd(&s.n_list, 1u);
      if (pos - surv_done > ring_mask) ovf = true; else surv[pos & ring_mask] = (uint16_t)(off + j);
    }
  };
  auto add8 = [](float (&sum)[8], const uint4& u) {
    const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&uw[j]));
      sum[2 * j] += f.x;
      sum[2 * j + 1] += f.y;
    }
  };
#pragma unroll 1
  for (uint32_t off = 8 * tid; off < n_docs; off += 16 * kT) {
    const uint32_t off2 = off + 8 * kT;
    const bool has2 = off2 < n_docs;
    float sa[8], sb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) sa[j] = sb[j] = 0.0f;
    if (ND) {
      uint4 ua[ND ? ND : 1], ub[ND ? ND : 1];
#pragma unroll
      for (int i = 0; i < (ND ? ND : 1); ++i) {
        ua[i] = __ldg(reinterpret_cast<const uint4*>(base[i] + off));
        ub[i] = has2 ? __ldg(reinterpret_cast<const uint4*>(base[i] + off2)) : make_uint4(0u, 0u, 0u, 0u);
      }
#pragma unroll
      for (int i = 0; i < (ND ? ND : 1); ++i) {
        add8(sa, ua[i]);
        add8(sb, ub[i]);
      }
    } else {
      for (uint32_t i = 0; i < nd; ++i) {
        const uint16_t* b = p.uvec + (size_t)s.dense_slots[i] * p.d_pad + slab_lo;
        const uint4 ua = __ldg(reinterpret_cast<const uint4*>(b + off));
        const uint4 ub = has2 ? __ldg(reinterpret_cast<const uint4*>(b + off2)) : make_uint4(0u, 0u, 0u, 0u);
        add8(sa, ua);
        add8(sb, ub);
      }
    }
    process(sa, off);
    if (has2) process(sb, off2);
  }
  return !__syncthreads_or(ovf ? 1 : 0);
}

__device__ void dtiv_path(const ScoreParams& p, Smem& s, uint32_t q, uint64_t slab_lo, uint64_t slab_hi, uint32_t n_kw,
                          uint32_t n_ph, double qm, float qf_inv, float blend_scale, uint32_t k,
                          unsigned long long& n_postings, unsigned long long& n_matched) {
  // Phrase tokens take part in the bound like keyword tokens: the phrase adds at most the sum of its
  // tokens' weights per table (phrase.go:97-106), and only if the positions line up, which the exact
  // evaluation of the survivors decides.  "Present" then means "has a posting of some query token".
  constexpr int RD = kDenseRange;
  constexpr uint32_t kRing = 2 * RD;                               // survivor ring capacity
  float* sacc = reinterpret_cast<float*>(&s.acc[0][0]);            // [RD] sparse tokens' impact sums
  uint16_t* surv = reinterpret_cast<uint16_t*>(&s.acc[1][0]);      // [kRing] slab-relative slots >> 0 of survivors
  uint32_t* dbits = s.dbits;                                       // [RD / 32] presence from sparse tokens
  const uint32_t tid = threadIdx.x, n_lists = 2 * (n_kw + n_ph);
  const uint32_t nd = s.n_dense_tok, nsp = s.n_sparse_tok;
  const uint32_t n_sub = (uint32_t)((slab_hi - slab_lo + RD - 1) / RD);
  if (tid == 0) {
    unsigned long long tot = 0;
    for (uint32_t l = 0; l < n_lists; ++l) tot += s.len[l];
    n_postings += tot;  // nominal: the posting lists this (query, slab) pair covers
  }
  // sparse tokens: posting offsets at every sub-range boundary
  bool has_sparse = false;
  for (uint32_t i = 0; i < nsp; ++i) has_sparse |= (s.len[2 * s.sparse_toks[i]] | s.len[2 * s.sparse_toks[i] + 1]) != 0;
  // whole-slab mode (below) applies when the sparse tokens have few postings here and a threshold exists
  constexpr uint32_t kSparseStage = 2048;
  const bool flush_each = (slab_hi - slab_lo) > 65536u;
  uint32_t sp_total = 0;
  if (has_sparse)
    for (uint32_t i = 0; i < nsp; ++i) sp_total += s.len[2 * s.sparse_toks[i]] + s.len[2 * s.sparse_toks[i] + 1];
  const bool whole_slab = !flush_each && sp_total <= kSparseStage && s.thr_f != -__int_as_float(0x7f800000);
  auto setup_subranges = [&]() {  // sub-range mode: boundary table of the sparse lists, clean scratch
    for (uint32_t idx = tid; idx < 2 * nsp * (n_sub + 1); idx += kT) {
      const uint32_t li = idx / (n_sub + 1), j = idx % (n_sub + 1);
      const uint32_t l = 2 * s.sparse_toks[li >> 1] + (li & 1);
      const uint64_t target = min(slab_hi, slab_lo + (uint64_t)j * RD);
      s.bounds[idx] = s.len[l] ? (uint32_t)(lower_bound_doc(p.tab[l & 1].doc_ids, s.base[l], s.base[l] + s.len[l], target) - s.base[l]) : 0u;
    }
    for (uint32_t i = tid; i < RD; i += kT) sacc[i] = 0.0f;
    for (uint32_t i = tid; i < RD / 32; i += kT) dbits[i] = 0;
  };
  if (has_sparse && !whole_slab) setup_subranges();
  // largest blend term of the slab (group-level rejection test): one zblk entry per thread, warp max,
  // combined through shared memory; a NaN entry disables the rejection
  {
    const uint64_t b0 = slab_lo / kRange, nb = (slab_hi - slab_lo + kRange - 1) / kRange;
    float zm = -__int_as_float(0x7f800000);
    bool nan = false;
    for (uint64_t b = tid; b < nb; b += kT) {
      const float z = p.zblk[b0 + b];
      nan |= z != z;
      zm = fmaxf(zm, z);
    }
    for (int o = 16; o; o >>= 1) zm = fmaxf(zm, __shfl_xor_sync(0xFFFFFFFFu, zm, o));
    nan = __any_sync(0xFFFFFFFFu, nan);
    if ((tid & 31) == 0) s.zred[tid >> 5] = nan ? __int_as_float(0x7fc00000) : zm;
  }
  __syncthreads();
  float za;
  {
    float zm = s.zred[0];
    bool nan = zm != zm;
    for (int w = 1; w < kT / 32; ++w) {
      const float z = s.zred[w];
      nan |= z != z;
      zm = fmaxf(zm, z);
    }
    za = nan ? __int_as_float(0x7fc00000) : blend_scale * zm;
  }
  // survivors carry (sub-range, slot) as a slab-relative doc offset in 16 bits when the slab allows it,
  // else they are flushed every sub-range (flush_each)
  uint32_t surv_done = s.n_list;
  uint32_t my_matched = 0;
  // Whole-slab mode: every token of the slab is dense, or the sparse tokens have few postings here.
  if (whole_slab) {
    const float thr_f = s.thr_f;
    // mixed layout: acc[0] = staged sparse postings (slab offset, impact), acc[1] = survivor ring (4096) + bitmap
    uint32_t* soff = reinterpret_cast<uint32_t*>(&s.acc[0][0]);
    float* simp = reinterpret_cast<float*>(soff + kSparseStage);
    uint32_t* excl = reinterpret_cast<uint32_t*>(surv + RD);  // [65536 / 32] words = 8 KB
    const uint32_t ring_mask = has_sparse ? (uint32_t)RD - 1u : kRing - 1u;
    if (has_sparse) {
      __syncthreads();  // the sparse scratch set up above (sacc / bounds) is not used in this mode
      for (uint32_t i = tid; i < 65536 / 32; i += kT) excl[i] = 0;
      if (tid == 0) {
        uint32_t run = 0;
        for (uint32_t li = 0; li < 2 * nsp; ++li) {
          s.bounds[li] = run;
          run += s.len[2 * s.sparse_toks[li >> 1] + (li & 1)];
        }
        s.bounds[2 * nsp] = run;
      }
      __syncthreads();
      for (uint32_t li = 0; li < 2 * nsp; ++li) {
        const uint32_t l = 2 * s.sparse_toks[li >> 1] + (li & 1), len = s.len[l];
        const TableView& tv = p.tab[l & 1];
        for (uint32_t i = tid; i < len; i += kT) {
          const uint32_t doc = tv.doc_ids[s.base[l] + i];
          const float w = tv.w[s.base[l] + i];
          const float4 m = p.meta32[doc];
          float v = (l & 1) ? 29.0f * (w * m.y) : 38.0f * (w * m.x);
          v = fmaxf(v, 0.0f) * 1.00001f;
          const uint32_t off = (uint32_t)(doc - slab_lo);
          soff[s.bounds[li] + i] = off;
          simp[s.bounds[li] + i] = v;
          atomicOr(&excl[off >> 5], 1u << (off & 31));
        }
      }
      __syncthreads();
    }
    uint32_t cnt = 0;
    bool ovf_forced = false;
    if (has_sparse) {
      // docs with a sparse posting: the first sparse list that holds the doc owns it and bounds it with the
      // doc's sparse impacts plus its entries of the dense vectors
      auto find = [&](uint32_t lj, uint32_t off) -> uint32_t {
        uint32_t lo = s.bounds[lj], hi = s.bounds[lj + 1];
        const uint32_t end = hi;
        while (lo < hi) {
          const uint32_t mid = (lo + hi) >> 1;
          if (soff[mid] < off) lo = mid + 1; else hi = mid;
        }
        return (lo < end && soff[lo] == off) ? lo : kNoDoc;
      };
      for (uint32_t idx = tid; idx < sp_total; idx += kT) {
        uint32_t lo = 0, hi = 2 * nsp;  // last li with bounds[li] <= idx
        while (hi - lo > 1) {
          const uint32_t mid = (lo + hi) >> 1;
          if (s.bounds[mid] <= idx) lo = mid; else hi = mid;
        }
        const uint32_t li = lo, off = soff[idx];
        bool owner = true;
        for (uint32_t j = 0; j < li && owner; ++j) owner = find(j, off) == kNoDoc;
        if (!owner) continue;
        ++cnt;
        float sum = simp[idx];
        for (uint32_t j = li + 1; j < 2 * nsp; ++j) {
          const uint32_t at = find(j, off);
          if (at != kNoDoc) sum += simp[at];
        }
        for (uint32_t i = 0; i < nd; ++i)
          sum += half_bits_to_float(p.uvec[(size_t)s.dense_slots[i] * p.d_pad + slab_lo + off]);
        const float z = half_bits_to_float(p.zvec[slab_lo + off]);
        const float a = blend_scale * z, b = qf_inv * sum;
        if ((a + b) + (fabsf(a) + fabsf(b)) * 1e-4f + 1e-30f < thr_f) continue;
        const uint32_t pos = atomicAdd(&s.n_list, 1u);
        if (pos - surv_done > ring_mask) ovf_forced = true; else surv[pos & ring_mask] = (uint16_t)off;
      }
    }
    const uint32_t* ex = has_sparse ? excl : nullptr;
    bool ok;
    if (nd == 1) ok = dtiv_stream_slab<1>(p, s, surv, ring_mask, surv_done, slab_lo, slab_hi, nd, za, blend_scale, qf_inv, thr_f, ex, cnt);
    else if (nd == 2) ok = dtiv_stream_slab<2>(p, s, surv, ring_mask, surv_done, slab_lo, slab_hi, nd, za, blend_scale, qf_inv, thr_f, ex, cnt);
    else ok = dtiv_stream_slab<0>(p, s, surv, ring_mask, surv_done, slab_lo, slab_hi, nd, za, blend_scale, qf_inv, thr_f, ex, cnt);
    ok = !__syncthreads_or(ovf_forced ? 1 : 0) && ok;
    if (ok) {
      n_matched += cnt;
      const uint32_t surv_end = s.n_list;
      for (uint32_t r0 = surv_done; r0 != surv_end; r0 += min((uint32_t)kCand, surv_end - r0)) {
        const uint32_t n_round = min((uint32_t)kCand, surv_end - r0);
        evaluate_survivors(p, s, q, surv, ring_mask, r0, n_round, slab_lo, n_kw, n_ph, slab_lo, slab_hi - slab_lo, qm, k);
        __syncthreads();
        if (s.n_cand) merge_candidates(s, k);
      }
      return;
    }
    // the survivor ring overflowed (masses of ties at the threshold): forget the pass, go sub-range by sub-range
    if (tid == 0) s.n_list = surv_done;
    if (has_sparse) setup_subranges();  // the staged postings used the sub-range scratch
    __syncthreads();
  }
  for (uint32_t sj = 0; sj < n_sub; ++sj) {
    const uint64_t d0 = slab_lo + (uint64_t)sj * RD, d1 = min(slab_hi, d0 + RD);
    const uint32_t rel0 = flush_each ? 0u : (uint32_t)(d0 - slab_lo);
    bool any_sp = false;
    if (has_sparse) {
      for (uint32_t li = 0; li < 2 * nsp; ++li)
        any_sp |= s.bounds[li * (n_sub + 1) + sj] != s.bounds[li * (n_sub + 1) + sj + 1];
    }
    if (any_sp) {
      for (uint32_t li = 0; li < 2 * nsp; ++li) {
        const uint32_t l = 2 * s.sparse_toks[li >> 1] + (li & 1);
        const unsigned long long x0 = s.base[l] + s.bounds[li * (n_sub + 1) + sj];
        const unsigned long long x1 = s.base[l] + s.bounds[li * (n_sub + 1) + sj + 1];
        const TableView& tv = p.tab[l & 1];
        for (unsigned long long x = x0 + tid; x < x1; x += kT) {
          const uint32_t doc = tv.doc_ids[x];
          const float w = tv.w[x];
          const float4 m = p.meta32[doc];
          float v = (l & 1) ? 29.0f * (w * m.y) : 38.0f * (w * m.x);
          v = fmaxf(v, 0.0f) * 1.00001f;  // NaN -> 0: a NaN component counts as 0 in the reference too
          const uint32_t slot = (uint32_t)(doc - d0);
          atomicAdd(&sacc[slot], v);
          atomicOr(&dbits[slot >> 5], 1u << (slot & 31));
        }
      }
      __syncthreads();
    }
    // stream the sub-range: 8 consecutive docs per thread and step.  A group whose best bound (largest
    // impact sum, the block's largest blend term) stays below the threshold is dropped with one test.
    const float thr_f = s.thr_f;
#pragma unroll 1
    for (uint32_t g0 = 0; g0 < RD; g0 += 8 * kT) {
      const uint32_t slot0 = g0 + 8 * tid;
      const uint64_t doc0 = d0 + slot0;
      if (doc0 >= d1) continue;
      float sum[8];
      uint32_t nzw[4] = {0u, 0u, 0u, 0u};  // 0xFFFF per half-word: some dense token has a posting of that doc
#pragma unroll
      for (int j = 0; j < 8; ++j) sum[j] = 0.0f;
      for (uint32_t i = 0; i < nd; ++i) {
        const uint4 u = __ldg(reinterpret_cast<const uint4*>(p.uvec + (size_t)s.dense_slots[i] * p.d_pad + doc0));
        const uint32_t uw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&uw[j]));
          sum[2 * j] += f.x;
          sum[2 * j + 1] += f.y;
          nzw[j] |= __vcmpne2(uw[j], 0u);
        }
      }
      uint32_t sp_bits = 0;
      if (any_sp) {
        sp_bits = (dbits[slot0 >> 5] >> (slot0 & 31)) & 0xFFu;
        const float4 a0 = *reinterpret_cast<const float4*>(sacc + slot0);
        const float4 a1 = *reinterpret_cast<const float4*>(sacc + slot0 + 4);
        sum[0] += a0.x; sum[1] += a0.y; sum[2] += a0.z; sum[3] += a0.w;
        sum[4] += a1.x; sum[5] += a1.y; sum[6] += a1.z; sum[7] += a1.w;
      }
      // present docs (dense posting or sparse posting), docs past the slab end masked out
      uint32_t present = sp_bits;
#pragma unroll
      for (int j = 0; j < 4; ++j) present |= ((nzw[j] & 1u) << (2 * j)) | (((nzw[j] >> 16) & 1u) << (2 * j + 1));
      if (doc0 + 8 > d1) present &= (1u << (uint32_t)(d1 - doc0)) - 1u;
      my_matched += __popc(present);
      if (!present) continue;
      float gmax = fmaxf(fmaxf(fmaxf(sum[0], sum[1]), fmaxf(sum[2], sum[3])), fmaxf(fmaxf(sum[4], sum[5]), fmaxf(sum[6], sum[7])));
      {
        const float b = qf_inv * gmax;
        // fmaxf drops NaNs: a NaN sum can only come from inf - inf, impossible here (all terms >= 0)
        if ((za + b) + (fabsf(za) + fabsf(b)) * 1e-4f + 1e-30f < thr_f) continue;
      }
      const uint4 zz = __ldg(reinterpret_cast<const uint4*>(p.zvec + doc0));
      const uint32_t zw[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!((present >> j) & 1u)) continue;
        const float z = half_bits_to_float(j & 1 ? zw[j >> 1] >> 16 : zw[j >> 1] & 0xFFFFu);
        const float a = blend_scale * z, b = qf_inv * sum[j];
        if ((a + b) + (fabsf(a) + fabsf(b)) * 1e-4f + 1e-30f < thr_f) continue;  // NaN and +inf stay in
        surv[atomicAdd(&s.n_list, 1u) & (kRing - 1)] = (uint16_t)(rel0 + slot0 + j);
      }
    }
    __syncthreads();
    // clean the sparse scratch for the next sub-range (not read again before the next barrier)
    if (any_sp) {
      for (uint32_t i = tid; i < RD; i += kT) sacc[i] = 0.0f;
      for (uint32_t i = tid; i < RD / 32; i += kT) dbits[i] = 0;
    }
    // Exact evaluation of the survivors stalls the block on a few threads' dependent lookups, so it
    // is deferred until the slab ends, the ring could overflow, or no threshold exists yet.
    const uint32_t surv_end = s.n_list;
    const bool flush = flush_each || sj + 1 == n_sub || surv_end - surv_done > (uint32_t)RD ||
                       (surv_end != surv_done && thr_f == -__int_as_float(0x7f800000));
    if (flush) {
      const uint64_t rel_base = flush_each ? d0 : slab_lo;
      for (uint32_t r0 = surv_done; r0 != surv_end; r0 += min((uint32_t)kCand, surv_end - r0)) {
        const uint32_t cnt = min((uint32_t)kCand, surv_end - r0);
        evaluate_survivors(p, s, q, surv, kRing - 1, r0, cnt, rel_base, n_kw, n_ph, slab_lo, slab_hi - slab_lo, qm, k);
        __syncthreads();
        if (s.n_cand) merge_candidates(s, k);
      }
      surv_done = surv_end;
    }
    if (any_sp) __syncthreads();  // the scratch is clean before the next scatter
  }
  n_matched += my_matched;
}

// PHRASE = false: the batch has no phrase token; that instantiation carries none of the phrase code (the
// keyword paths lose ~9 % to register pressure and code size otherwise).
template <bool PHRASE>
__global__ void __launch_bounds__(kT, 3) k_score(ScoreParams p) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  Smem& s = *reinterpret_cast<Smem*>(smem_raw);
  const uint32_t q = blockIdx.x % p.n_q, slab = blockIdx.x / p.n_q;  // slab-major launch order
  const uint32_t glen = p.group_len[(size_t)q * p.n_slabs + slab];
  if (glen == 0) return;  // scored by the CTA that leads this slab's group (its part_count stays 0)
  const uint32_t tid = threadIdx.x;
  const uint64_t kb = p.kw_ptr[q], ke = p.kw_ptr[q + 1];
  const uint64_t pb = (PHRASE && p.ph_ptr) ? p.ph_ptr[q] : 0, pe = (PHRASE && p.ph_ptr) ? p.ph_ptr[q + 1] : 0;
  const uint32_t n_kw = (uint32_t)(ke - kb);
  uint32_t n_ph = PHRASE ? (uint32_t)(pe - pb) : 0u;
  const uint32_t q_len = n_kw + n_ph;            // main_retrieve.go:90
  if (n_ph > kMaxPh) n_ph = 0;                   // host rejects 33..256; >256 can never match (uint8 TermPos)
  const uint32_t n_tok = n_kw + n_ph, n_lists = 2 * n_tok;
  const uint64_t slab_lo = (uint64_t)slab * p.slab_docs;
  const uint64_t slab_hi = min(p.D, slab_lo + (uint64_t)glen * p.slab_docs);
  const uint32_t k = p.k;

  if (tid == 0) {
    s.n_cand = 0;
    s.n_ent = 0;
    s.n_list = 0;
    s.top_n = 0;
    s.top_buf = 0;
    s.thr_key = 0;
    s.thr_doc = kNoDoc;
    // Running bound of the query: CTAs are launched slab-major, so by the time slab j of a query
    // starts, earlier slabs have usually published their k-th best score.  A doc scoring strictly
    // below ANY slab's k-th best cannot be among the query's k best, so it is dropped here exactly
    // as the per-slab threshold drops it; ties are kept (the doc id decides them in k_merge).  The
    // result does not depend on which bound a CTA happens to see.
    const unsigned long long g = p.use_qthr ? *reinterpret_cast<volatile unsigned long long*>(p.qthr + q) : 0ull;
    const double gs = key_score(g);
    s.gkey = g;
    s.gthr_f = (g == 0ull || isnan(gs)) ? -__int_as_float(0x7f800000) : __double2float_rd(gs);
    s.thr_f = s.gthr_f;
  }
  // which keyword tokens have an impact vector
  if (tid == 0) {
    uint32_t ndt = 0, nst = 0, ndk = 0;
    for (uint32_t i = 0; i < n_tok; ++i) {
      const uint32_t term = i < n_kw ? p.kw_terms[kb + i] : p.ph_terms[pb + (i - n_kw)];
      const uint8_t slot = (p.uvec && term < p.dense_map_V) ? p.dense_map[term] : (uint8_t)255;
      s.tok_dense[i] = slot;
      if (slot != 255) s.dense_slots[ndt++] = slot; else s.sparse_toks[nst++] = (uint8_t)i;
      if (slot != 255 && i < n_kw) ++ndk;
    }
    s.n_dense_kw = ndk;
    s.n_dense_tok = ndt;
    s.n_sparse_tok = nst;
  }
  // narrow every list to the slab
  for (uint32_t l = tid; l < n_lists; l += kT) {
    const uint32_t tok = l >> 1, tb = l & 1;
    const uint32_t term = tok < n_kw ? p.kw_terms[kb + tok] : p.ph_terms[pb + (tok - n_kw)];
    const TableView& tv = p.tab[tb];
    unsigned long long a = 0;
    if (tv.term_ptr && term < tv.V) a = tv.term_ptr[term];  // unknown term => empty row (main_retrieve.go:193,218)
    // slab boundaries of every list were located once for the whole batch (k_narrow)
    const uint32_t* nar = p.narrow + ((size_t)2 * (kb + pb) + l) * (p.n_slabs + 1) + slab;
    const uint32_t o0 = nar[0], o1 = nar[glen];
    s.base[l] = a + o0;
    s.len[l] = o1 - o0;
  }
  __syncthreads();

  // how much work is there in this slab?  (uniform: every thread reads the same shared values)
  unsigned long long work = 0;
  for (uint32_t l = 0; l < 2 * n_kw; ++l) work += s.len[l];
  for (int tb = 0; tb < 2 && n_ph; ++tb) {
    uint32_t mn = 0xFFFFFFFFu;
    for (uint32_t i = 0; i < n_ph; ++i) mn = min(mn, s.len[2 * n_kw + 2 * i + tb]);
    work += mn;  // a phrase can hit at most once per posting of its shortest list
  }
  const double qm = sqrt((double)q_len);  // get_metadata.go:53
  const float qf_inv = 1.0f / (float)qm;
  float blend_scale = 1.0f;  // screening: |sum_t p_t PR_t| <= (sum_t |p_t|) * max_t |PR_t|
  if (p.probs) {
    blend_scale = 0.0f;
    for (uint32_t t = 0; t < p.T; ++t) blend_scale += fabsf((float)p.probs[(uint64_t)q * p.T + t]);
    blend_scale *= 1.0001f;
  }
  // screening coefficients: 100 * (0.33 * blend, 0.38 / |q|, 0.29 / |q|)
  const float sc_a = 33.0f * blend_scale, sc_t = 38.0f * qf_inv, sc_b = 29.0f * qf_inv;
  unsigned long long n_postings = 0, n_matched = 0;

  if (work == 0) {
    // nothing of this query lives in this slab
  } else if (work > p.sort_max &&
             // a phrase query needs a dense KEYWORD token: its many matches set the threshold the phrase
             // docs are screened against (phrase hits alone are too rare to ever establish one)
             (n_ph ? (p.phrase_dense && s.n_dense_kw > 0) : s.n_dense_tok > 0) &&
             2 * s.n_sparse_tok * ((slab_hi - slab_lo + kDenseRange - 1) / kDenseRange + 1) <= (uint64_t)kBounds) {
    dtiv_path(p, s, q, slab_lo, slab_hi, n_kw, n_ph, qm, qf_inv, blend_scale, k, n_postings, n_matched);
  } else if (work <= p.sort_max && p.owner_path && (n_ph == 0 || slab_hi - slab_lo <= 0xFFFFFFFFull)) {
    owner_path(p, s, q, slab_lo, n_kw, n_ph, qm, qf_inv, blend_scale, k, n_postings, n_matched);
  } else if (work <= p.sort_max && slab_hi - slab_lo <= (1ull << 24)) {
    sort_path(p, s, q, slab_lo, n_kw, n_ph, qm, qf_inv, blend_scale, k, n_postings, n_matched);
  } else {
  for (uint32_t i = tid; i < kRange; i += kT) {
    s.acc[0][i] = 0.0;
    s.acc[1][i] = 0.0;
  }
  for (uint32_t i = tid; i < kRange / 32; i += kT) s.bits[i] = 0;
  uint32_t list_done = 0;  // n_list at the end of the previous sub-range (uniform)
  // Sub-range boundaries of every list are found up front, all threads searching
  // in parallel (one dependent-load chain per CTA pass instead of one per sub-range).
  const uint32_t n_sub = (uint32_t)((slab_hi - slab_lo + kRange - 1) / kRange);
  const uint32_t pass_sub = n_lists ? max(1u, min(n_sub, (uint32_t)kBounds / n_lists - 1u)) : n_sub;
  for (uint32_t sub0 = 0; sub0 < n_sub; sub0 += pass_sub) {
  const uint32_t nb = min(pass_sub, n_sub - sub0);
  __syncthreads();
  for (uint32_t idx = tid; idx < n_lists * (nb + 1); idx += kT) {
    const uint32_t l = idx / (nb + 1), j = idx % (nb + 1);
    const uint64_t target = min(slab_hi, slab_lo + (uint64_t)(sub0 + j) * kRange);
    const uint32_t* docs = p.tab[l & 1].doc_ids;
    s.bounds[idx] = s.len[l] ? (uint32_t)(lower_bound_doc(docs, s.base[l], s.base[l] + s.len[l], target) - s.base[l]) : 0u;
  }
  __syncthreads();
  for (uint32_t sj = 0; sj < nb; ++sj) {
    const uint64_t d0 = slab_lo + (uint64_t)(sub0 + sj) * kRange;
    // list l covers postings [base + bounds[l][sj], base + bounds[l][sj+1]) in this sub-range
    auto lo_of = [&](uint32_t l) { return s.base[l] + s.bounds[l * (nb + 1) + sj]; };
    auto hi_of = [&](uint32_t l) { return s.base[l] + s.bounds[l * (nb + 1) + sj + 1]; };

    // the screening records of this sub-range (kRange x 16 B, one 128-byte line per thread) are
    // requested now so that they arrive in L1 while the postings are being applied
    if (p.prefetch_meta) {
      const uint64_t first = d0 + (uint64_t)tid * 8;
      if (first < p.D) asm volatile("prefetch.global.L1 [%0];" ::"l"(p.meta32 + first));
    }
    // keyword tokens in query order (duplicates count again); a barrier only after a token
    // that touched the accumulators
    bool any = false;
    for (uint32_t i = 0; i < n_kw; ++i) {
      const unsigned long long t0 = lo_of(2 * i), t1 = hi_of(2 * i), b0 = lo_of(2 * i + 1), b1 = hi_of(2 * i + 1);
      if (t1 == t0 && b1 == b0) continue;
      if (any) __syncthreads();  // the previous token's updates are complete
      any = true;
      accumulate_list(p.tab[1], s, 1, b0, b1, d0);
      accumulate_list(p.tab[0], s, 0, t0, t1, d0);
      if (tid == 0) n_postings += (t1 - t0) + (b1 - b0);
    }
    // the phrase's weight is appended after the keyword weights (main_retrieve.go:73-78)
    if (n_ph) {
      bool ph_any = false;
      for (uint32_t i = 0; i < 2 * n_ph; ++i) ph_any |= hi_of(2 * n_kw + i) != lo_of(2 * n_kw + i);
      if (ph_any) {
        if (any) __syncthreads();
        for (uint32_t l = 2 * n_kw + tid; l < n_lists; l += kT) {
          s.cur[l] = lo_of(l);
          s.hi[l] = hi_of(l);
        }
        __syncthreads();
        any = true;
        apply_phrase<0>(p, s, 1, 2 * n_kw, n_ph, d0, 0, n_postings);
        apply_phrase<0>(p, s, 0, 2 * n_kw, n_ph, d0, 0, n_postings);
      }
    }
    if (!any) continue;  // uniform
    __syncthreads();

    // finish the matched docs from the first-touch list, kCand (the candidate buffer's capacity)
    // per round; a thread's docs of a round are fetched together before any of them is scored.
    // n_list only grows (ring index), so nothing but the bitmap needs a reset.
    const uint32_t list_end = s.n_list;
    if (tid == 0) n_matched += list_end - list_done;
    if (tid < kRange / 32) s.bits[tid] = 0;  // not read again before the barrier below
    const float4* meta_base = p.meta32 + d0;
    constexpr int kPer = kCand / kT;
    for (uint32_t r0 = list_done; r0 != list_end; r0 += min((uint32_t)kCand, list_end - r0)) {
      bool has[kPer];
      uint32_t slot[kPer];
      float4 meta[kPer];
      double tr[kPer], br[kPer];
#pragma unroll
      for (int c = 0; c < kPer; ++c) {
        const uint32_t i = tid + c * kT;
        has[c] = i < list_end - r0;
        slot[c] = has[c] ? s.mlist[(r0 + i) & (kRange - 1)] : 0u;
      }
#pragma unroll
      for (int c = 0; c < kPer; ++c) {
        if (!has[c]) continue;
        meta[c] = meta_base[slot[c]];
        tr[c] = s.acc[0][slot[c]];
        br[c] = s.acc[1][slot[c]];
      }
      // the running k-th best only changes in merge_candidates, i.e. between rounds
      const float thr_f = s.thr_f;
#pragma unroll
      for (int c = 0; c < kPer; ++c) {
        if (!has[c]) continue;
        s.acc[0][slot[c]] = 0.0;
        s.acc[1][slot[c]] = 0.0;
        // screening (see finish_doc): sc_a/sc_t/sc_b fold the blend weights, 1/|q| and the x100
        const float a = sc_a * meta[c].z;
        const float b = tr[c] != 0.0 ? sc_t * ((float)tr[c] * meta[c].x) : 0.0f;
        const float cc = br[c] != 0.0 ? sc_b * ((float)br[c] * meta[c].y) : 0.0f;
        if ((a + b + cc) + (fabsf(a) + fabsf(b) + fabsf(cc)) * 1e-4f + 1e-30f < thr_f) continue;
        finish_exact(p, s, q, d0 + slot[c], tr[c], br[c], qm, k);
      }
      __syncthreads();  // accumulators of this round are clean, candidates are complete
      if (s.n_cand) merge_candidates(s, k);
    }
    list_done = list_end;
  }
  }
  }

  // this slab's list
  __syncthreads();
  const size_t base = ((size_t)q * p.n_slabs + slab) * k;
  const uint32_t tb = s.top_buf, tn = s.top_n;
  for (uint32_t j = tid; j < k; j += kT) {
    if (j < tn) {
      const uint32_t doc = s.top_doc[tb][j];
      double sqd = 0.0;
      if (p.sqd) {
        sqd = p.sqd[doc];
      } else if (p.probs) {
        const double* pr = p.pr + (uint64_t)doc * p.T;
        const double* pq = p.probs + (uint64_t)q * p.T;
        for (uint32_t t = 0; t < p.T; ++t) sqd = __dadd_rn(sqd, __dmul_rn(pq[t], pr[t]));
      }
      p.part_doc[base + j] = doc;
      p.part_final[base + j] = key_score(s.top_key[tb][j]);
      p.part_pr[base + j] = sqd;
    } else {
      p.part_doc[base + j] = kNoDoc;
      p.part_final[base + j] = 0.0;
      p.part_pr[base + j] = 0.0;
    }
  }
  if (tid == 0) {
    p.part_count[(size_t)q * p.n_slabs + slab] = tn;
    if (p.use_qthr && tn == k && s.top_key[tb][k - 1] > s.gkey) atomicMax(p.qthr + q, s.top_key[tb][k - 1]);
  }
  // stats: warp-reduce then one atomic per warp
  for (int o = 16; o; o >>= 1) {
    n_postings += __shfl_xor_sync(0xFFFFFFFFu, n_postings, o);
    n_matched += __shfl_xor_sync(0xFFFFFFFFu, n_matched, o);
  }
  if ((tid & 31) == 0) {
    if (n_postings) atomicAdd(p.stats, n_postings);
    if (n_matched) atomicAdd(p.stats + 1, n_matched);
  }
}

// Merge n_lists lists of up to k results per query into one, same total order.
// in_*: [n_q][n_lists][k] (list_major == 0) or [n_lists][n_q][k] (list_major == 1).
// Every input list is sorted best first (k_score writes its running top k in rank order, and so does this
// kernel), so this is a k-way merge: k rounds, each picks the best list head with a block-wide argmax over
// the total order (score key descending, doc id ascending).  Cost k * (n_lists / 256 + log 256), whatever
// the lists hold -- the first version ranked all n_lists * k entries against each other, which at k = 50
// (half-full lists: the running bound is only a slab-local k-th best) took twice as long as the scoring.
constexpr int kMergeMax = 16384;  // n_lists * k accepted (two bytes of shared memory per list)
__global__ void __launch_bounds__(kT) k_merge(uint32_t n_lists, uint32_t k, uint32_t n_q, int list_major,
                                              const uint32_t* __restrict__ in_doc, const double* __restrict__ in_final,
                                              const double* __restrict__ in_pr, const uint32_t* __restrict__ in_count,
                                              uint32_t* __restrict__ out_doc, double* __restrict__ out_final,
                                              double* __restrict__ out_pr, uint32_t* __restrict__ out_count) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint8_t* head = smem_raw;            // [n_lists] next unread entry of every list
  uint8_t* cnt = smem_raw + n_lists;   // [n_lists] entries of every list (<= k <= 128)
  __shared__ unsigned long long w_key[kT / 32];
  __shared__ uint32_t w_doc[kT / 32], w_list[kT / 32], win_list;
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  auto src_of = [&](uint32_t l, uint32_t j) -> size_t {
    return list_major ? ((size_t)l * n_q + q) * k + j : ((size_t)q * n_lists + l) * k + j;
  };
  for (uint32_t l = tid; l < n_lists; l += kT) {
    head[l] = 0;
    cnt[l] = (uint8_t)min(k, in_count[list_major ? (size_t)l * n_q + q : (size_t)q * n_lists + l]);
  }
  __syncthreads();
  // the head of the thread's first list is cached in registers (n_lists <= 256 is the common shape)
  unsigned long long c_key = 0;
  uint32_t c_doc = kNoDoc;
  bool c_valid = false;
  auto load_head = [&](uint32_t l, unsigned long long& key, uint32_t& doc) -> bool {
    if (head[l] >= cnt[l]) return false;
    const size_t src = src_of(l, head[l]);
    doc = in_doc[src];
    if (doc == kNoDoc) return false;  // padding: the list ends here
    key = score_key(in_final[src]);
    return true;
  };
  if (tid < n_lists) c_valid = load_head(tid, c_key, c_doc);
  uint32_t n_out = 0;
  for (uint32_t step = 0; step < k; ++step) {
    unsigned long long b_key = c_key;
    uint32_t b_doc = c_doc, b_list = c_valid ? tid : kNoDoc;
    for (uint32_t l = tid + kT; l < n_lists; l += kT) {  // further lists of this thread: read every round
      unsigned long long key;
      uint32_t doc;
      if (!load_head(l, key, doc)) continue;
      if (b_list == kNoDoc || beats(key, doc, b_key, b_doc)) {
        b_key = key;
        b_doc = doc;
        b_list = l;
      }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      const unsigned long long o_key = __shfl_xor_sync(0xFFFFFFFFu, b_key, o);
      const uint32_t o_doc = __shfl_xor_sync(0xFFFFFFFFu, b_doc, o), o_list = __shfl_xor_sync(0xFFFFFFFFu, b_list, o);
      if (o_list != kNoDoc && (b_list == kNoDoc || beats(o_key, o_doc, b_key, b_doc))) {
        b_key = o_key;
        b_doc = o_doc;
        b_list = o_list;
      }
    }
    if (lane == 0) {
      w_key[warp] = b_key;
      w_doc[warp] = b_doc;
      w_list[warp] = b_list;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned long long key = w_key[0];
      uint32_t doc = w_doc[0], lst = w_list[0];
      for (int w = 1; w < kT / 32; ++w)
        if (w_list[w] != kNoDoc && (lst == kNoDoc || beats(w_key[w], w_doc[w], key, doc))) {
          key = w_key[w];
          doc = w_doc[w];
          lst = w_list[w];
        }
      win_list = lst;
      if (lst != kNoDoc) {
        const size_t src = src_of(lst, head[lst]);
        out_doc[(size_t)q * k + step] = doc;
        out_final[(size_t)q * k + step] = in_final[src];
        out_pr[(size_t)q * k + step] = in_pr[src];
        head[lst] = head[lst] + 1;
      }
    }
    __syncthreads();
    const uint32_t wl = win_list;
    if (wl == kNoDoc) break;  // every list is exhausted (uniform)
    ++n_out;
    if (wl == tid) c_valid = load_head(tid, c_key, c_doc);  // the owner refreshes its cached head
  }
  for (uint32_t j = n_out + tid; j < k; j += kT) {
    out_doc[(size_t)q * k + j] = kNoDoc;
    out_final[(size_t)q * k + j] = 0.0;
    out_pr[(size_t)q * k + j] = 0.0;
  }
  if (tid == 0) out_count[q] = n_out;
}

// Slab groups of every query: consecutive slabs are merged while their postings (all lists of the query)
// stay within merge_max, so that a query with short lists costs a few CTAs instead of one nearly empty
// CTA per slab (each pays the same prologue, barriers and output writes).  Merged groups are small enough
// for the sparse paths, which do not depend on the width of the doc range.
__global__ void k_plan(ScoreParams p, uint32_t merge_max, uint8_t* __restrict__ group_len) {
  const uint32_t q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= p.n_q) return;
  const uint64_t kb = p.kw_ptr[q], pb = p.ph_ptr ? p.ph_ptr[q] : 0;
  const uint32_t n_lists = 2 * (uint32_t)((p.kw_ptr[q + 1] - kb) + (p.ph_ptr ? p.ph_ptr[q + 1] - pb : 0));
  const uint32_t per = p.n_slabs + 1;
  const uint32_t* nar = p.narrow + (size_t)2 * (kb + pb) * per;
  auto work = [&](uint32_t sl) {
    unsigned long long w = 0;
    for (uint32_t l = 0; l < n_lists; ++l) w += nar[(size_t)l * per + sl + 1] - nar[(size_t)l * per + sl];
    return w;
  };
  // keyword queries with a dense term: the impact-vector path streams up to 65536 docs per CTA
  bool dense_q = false;
  if (p.uvec) {
    for (uint64_t i = kb; i < p.kw_ptr[q + 1] && !dense_q; ++i) {
      const uint32_t term = p.kw_terms[i];
      dense_q = term < p.dense_map_V && p.dense_map[term] != 255;
    }
    if (p.ph_ptr && p.ph_ptr[q + 1] > pb && !p.phrase_dense) dense_q = false;
  }
  const uint32_t dense_len = dense_q ? (uint32_t)max((uint64_t)1, min((uint64_t)255, 65536 / p.slab_docs)) : 1u;
  uint8_t* out = group_len + (size_t)q * p.n_slabs;
  uint32_t sl = 0;
  while (sl < p.n_slabs) {
    unsigned long long w = work(sl);
    if (dense_len > 1) {
      const uint32_t len = min(dense_len, p.n_slabs - sl);
      unsigned long long wg = w;
      for (uint32_t i = 1; i < len; ++i) wg += work(sl + i);
      if (wg > p.sort_max) {
        out[sl] = (uint8_t)len;
        for (uint32_t i = 1; i < len; ++i) out[sl + i] = 0;
        sl += len;
        continue;
      }
    }
    uint32_t len = 1;
    while (w <= merge_max && sl + len < p.n_slabs && len < 255) {
      const unsigned long long w2 = work(sl + len);
      if (w + w2 > merge_max) break;
      w += w2;
      ++len;
    }
    out[sl] = (uint8_t)len;
    for (uint32_t i = 1; i < len; ++i) out[sl + i] = 0;
    sl += len;
  }
}

// Slab boundaries of every posting list of the batch: narrow[list][j] = offset, inside the
// term's row, of the first posting with doc >= j * slab_docs.  One independent binary search
// per (list, boundary) instead of a dependent chain at the start of every (query, slab) CTA.
__global__ void k_narrow(ScoreParams p, uint32_t* __restrict__ narrow, uint64_t n_entries) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_entries) return;
  const uint32_t per = p.n_slabs + 1;
  const uint64_t list = i / per;
  const uint32_t j = (uint32_t)(i % per);
  // list -> (query, token, table): lists are laid out query by query, keyword tokens then phrase tokens
  uint64_t lo = 0, hi = p.n_q;  // last q with 2*(kw_ptr[q] + ph_ptr[q]) <= list
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    const uint64_t first = 2 * (p.kw_ptr[mid] + (p.ph_ptr ? p.ph_ptr[mid] : 0));
    if (first <= list) lo = mid; else hi = mid;
  }
  const uint64_t q = lo;
  const uint64_t kb = p.kw_ptr[q], n_kw = p.kw_ptr[q + 1] - kb, pb = p.ph_ptr ? p.ph_ptr[q] : 0;
  const uint32_t l = (uint32_t)(list - 2 * (kb + pb));
  const uint32_t tok = l >> 1, tb = l & 1;
  const uint32_t term = tok < n_kw ? p.kw_terms[kb + tok] : p.ph_terms[pb + (tok - n_kw)];
  const TableView& tv = p.tab[tb];
  uint32_t off = 0;
  if (tv.term_ptr && term < tv.V) {
    const uint64_t a = tv.term_ptr[term], b = tv.term_ptr[term + 1];
    const uint64_t target = min(p.D, (uint64_t)j * p.slab_docs);
    off = (uint32_t)(lower_bound_doc(tv.doc_ids, a, b, j == p.n_slabs ? p.D : target) - a);
  }
  narrow[i] = off;
}

// sqd[d] = sum_t probs[t] * pr[d][t], ascending t, separately rounded (get_metadata.go:39-42)
__global__ void k_sqd(const double* __restrict__ pr, const double* __restrict__ probs, uint32_t T, uint64_t D,
                      double* __restrict__ sqd) {
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  double acc = 0.0;
  for (uint32_t t = 0; t < T; ++t) acc = __dadd_rn(acc, __dmul_rn(probs[t], pr[d * T + t]));
  sqd[d] = acc;
}

// Screening record per doc: fp32 reciprocals of the two norms and the blend term (shared
// topic vector: sqd itself; per-query vectors: max_t |PR[doc][t]|; no blend: 0).
__global__ void k_meta32(const double* __restrict__ mag_t, const double* __restrict__ mag_b,
                         const double* __restrict__ sqd, const double* __restrict__ pr, uint32_t T, uint64_t D,
                         float4* __restrict__ out) {
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float4 m;
  m.x = 1.0f / (float)mag_t[d];
  m.y = 1.0f / (float)mag_b[d];
  m.z = 0.0f;
  if (sqd) {
    m.z = (float)sqd[d];
  } else if (pr) {
    float mx = 0.0f;
    for (uint32_t t = 0; t < T; ++t) mx = fmaxf(mx, fabsf((float)pr[d * T + t]) * 1.0001f);
    m.z = mx;
  }
  m.w = 0.0f;
  out[d] = m;
}

// ---- impact vectors: build ---------------------------------------------------------------------
// candidates: terms whose title + body postings reach min_df
__global__ void k_dense_candidates(TableView t0, TableView t1, uint64_t V, uint64_t min_df, uint32_t cap,
                                   uint32_t* __restrict__ n_out, uint32_t* __restrict__ out_term,
                                   unsigned long long* __restrict__ out_df) {
  const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= V) return;
  unsigned long long df = 0;
  if (t0.term_ptr && t < t0.V) df += t0.term_ptr[t + 1] - t0.term_ptr[t];
  if (t1.term_ptr && t < t1.V) df += t1.term_ptr[t + 1] - t1.term_ptr[t];
  if (df < min_df || df == 0) return;
  const uint32_t i = atomicAdd(n_out, 1u);
  if (i < cap) {
    out_term[i] = (uint32_t)t;
    out_df[i] = df;
  }
}
__global__ void k_dense_map(const uint32_t* __restrict__ terms, uint32_t n, uint8_t* __restrict__ map) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) map[terms[i]] = (uint8_t)i;
}
// One table's postings of the dense terms into the vectors: U[slot][doc] (+)= coef * w / norm, rounded up,
// never 0 for a posting.  blockIdx.y = dense slot.  The two tables run one after the other (a table holds a
// doc at most once per term, so there are no races).
__global__ void k_dense_fill(TableView tv, int table, const uint32_t* __restrict__ terms, const float4* __restrict__ meta32,
                             uint64_t d_pad, uint16_t* __restrict__ uvec) {
  const uint32_t term = terms[blockIdx.y];
  if (!tv.term_ptr || term >= tv.V) return;
  const uint64_t a = tv.term_ptr[term], b = tv.term_ptr[term + 1];
  uint16_t* u = uvec + (size_t)blockIdx.y * d_pad;
  for (uint64_t x = a + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; x < b; x += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t doc = tv.doc_ids[x];
    const float w = tv.w[x];
    const float4 m = meta32[doc];
    float v = table ? 29.0f * (w * m.y) : 38.0f * (w * m.x);
    v = fmaxf(v, 0.0f) * 1.00001f;  // NaN -> 0 (a NaN component counts as 0, get_metadata.go:61-66)
    const float old = __half2float(__ushort_as_half(u[doc]));
    unsigned short h = __half_as_ushort(__float2half_ru(old + v * 1.00001f));
    if (h == 0) h = 1;  // smallest positive value: "has a posting"
    u[doc] = h;
  }
}
// largest 33 * blend input of every kRange-doc block (one CTA per block)
__global__ void k_zblk(const float4* __restrict__ meta32, uint64_t D, float* __restrict__ zblk) {
  __shared__ float sm[256];
  const uint64_t d0 = (uint64_t)blockIdx.x * kRange;
  float m = -__int_as_float(0x7f800000);
  for (uint64_t d = d0 + threadIdx.x; d < d0 + kRange && d < D; d += blockDim.x) {
    float z = 33.0f * meta32[d].z;
    z += fabsf(z) * 1e-5f;
    m = z > m || z != z ? z : m;  // a NaN blend input poisons the block: its groups are never dropped
    if (z != z) break;
  }
  sm[threadIdx.x] = m;
  __syncthreads();
  if (threadIdx.x == 0) {
    float r = sm[0];
    for (int i = 1; i < 256; ++i) r = (sm[i] > r || sm[i] != sm[i]) && r == r ? sm[i] : r;
    zblk[blockIdx.x] = r == -__int_as_float(0x7f800000) ? 0.0f : r;
  }
}
__global__ void k_zvec(const float4* __restrict__ meta32, uint64_t D, uint64_t d_pad, uint16_t* __restrict__ zvec) {
  const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= d_pad) return;
  float z = 0.0f;
  if (d < D) {
    z = 33.0f * meta32[d].z;
    z += fabsf(z) * 1e-5f;
  }
  zvec[d] = __half_as_ushort(__float2half_ru(z));
}

TableView view_of(const TableState& tb) {
  TableView v{};
  if (!tb.loaded) return v;
  v.term_ptr = tb.term_ptr.p;
  v.doc_ids = tb.doc_ids.p;
  v.w = tb.w.p;
  v.pos_ptr = tb.has_pos ? tb.pos_ptr.p : nullptr;
  v.pos = tb.has_pos ? tb.pos.p : nullptr;
  v.V = tb.V;
  return v;
}

}  // namespace

// Build (or refresh) the impact vectors; see IndexState.  Called with the engine lock held and meta32 fresh.
static int build_dense_vectors(ss_engine* e, IndexState* ix, cudaStream_t st, uint32_t* launches) {
  const uint64_t D = ix->D;
  uint32_t frac = 32, max_dense = 224;
  if (const char* env = getenv("SS_SCORE_DENSE_FRAC")) frac = (uint32_t)std::max(1, atoi(env));
  if (const char* env = getenv("SS_SCORE_DENSE_MAX")) max_dense = (uint32_t)std::min(kMaxDense, std::max(0, atoi(env)));
  const uint64_t d_pad = (D + kDenseRange - 1) / kDenseRange * kDenseRange + kDensePadDocs;
  if (!ix->dense_valid) {
    ix->n_dense = 0;
    const uint64_t V = std::max(ix->tab[0].loaded ? ix->tab[0].V : 0, ix->tab[1].loaded ? ix->tab[1].V : 0);
    if (V && max_dense) {
      constexpr uint32_t cap = 4096;
      ss::DevBuf<uint32_t> d_n, d_term;
      ss::DevBuf<unsigned long long> d_df;
      SS_TRY(d_n.alloc(1));
      SS_TRY(d_term.alloc(cap));
      SS_TRY(d_df.alloc(cap));
      SS_CUDA(cudaMemsetAsync(d_n.p, 0, 4, st));
      const uint64_t min_df = std::max<uint64_t>(1, D / frac);
      k_dense_candidates<<<ss::div_up(V, 256), 256, 0, st>>>(view_of(ix->tab[0]), view_of(ix->tab[1]), V, min_df, cap,
                                                           d_n.p, d_term.p, d_df.p);
      uint32_t n = 0;
      SS_CUDA(cudaMemcpyAsync(&n, d_n.p, 4, cudaMemcpyDeviceToHost, st));
      SS_CUDA(cudaStreamSynchronize(st));
      n = std::min(n, cap);
      std::vector<uint32_t> terms(n);
      std::vector<unsigned long long> dfs(n);
      if (n) {
        SS_CUDA(cudaMemcpy(terms.data(), d_term.p, n * 4, cudaMemcpyDeviceToHost));
        SS_CUDA(cudaMemcpy(dfs.data(), d_df.p, n * 8, cudaMemcpyDeviceToHost));
      }
      std::vector<uint32_t> order(n);
      for (uint32_t i = 0; i < n; ++i) order[i] = i;
      std::sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
        return dfs[a] != dfs[b] ? dfs[a] > dfs[b] : terms[a] < terms[b];
      });
      // the vectors may take at most a quarter of the device memory that is free now (2 bytes per doc and term)
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) == cudaSuccess)
        max_dense = (uint32_t)std::min<uint64_t>(max_dense, (free_b / 4) / (d_pad * 2));
      const uint32_t nd = std::min(n, max_dense);
      std::vector<uint32_t> chosen(nd);
      for (uint32_t i = 0; i < nd; ++i) chosen[i] = terms[order[i]];
      if (nd) {
        SS_TRY(ws_reserve(ix->uvec, (size_t)nd * d_pad));
        SS_TRY(ws_reserve(ix->dense_map, V));
        SS_CUDA(cudaMemsetAsync(ix->uvec.p, 0, (size_t)nd * d_pad * 2, st));
        SS_CUDA(cudaMemsetAsync(ix->dense_map.p, 0xFF, V, st));
        SS_CUDA(cudaMemcpyAsync(d_term.p, chosen.data(), nd * 4, cudaMemcpyHostToDevice, st));
        k_dense_map<<<ss::div_up(nd, 256), 256, 0, st>>>(d_term.p, nd, ix->dense_map.p);
        const dim3 grid(64, nd);
        for (int tb = 1; tb >= 0; --tb)
          if (ix->tab[tb].loaded)
            k_dense_fill<<<grid, 256, 0, st>>>(view_of(ix->tab[tb]), tb, d_term.p, ix->meta32.p, d_pad, ix->uvec.p);
        SS_CUDA(cudaStreamSynchronize(st));  // d_term goes out of scope
        SS_CUDA(cudaGetLastError());
        *launches += 4;
      }
      ix->n_dense = nd;
      ix->dense_map_V = V;
    }
    ix->d_pad = d_pad;
    ix->dense_valid = true;
    ix->zvec_valid = false;
  }
  if (ix->n_dense && !ix->zvec_valid) {
    SS_TRY(ws_reserve(ix->zvec, d_pad));
    SS_TRY(ws_reserve(ix->zblk, d_pad / kRange));
    k_zvec<<<ss::div_up(d_pad, 256), 256, 0, st>>>(ix->meta32.p, D, d_pad, ix->zvec.p);
    k_zblk<<<(unsigned)(d_pad / kRange), 256, 0, st>>>(ix->meta32.p, D, ix->zblk.p);
    *launches += 2;
    ix->zvec_valid = true;
  }
  (void)e;
  return SS_OK;
}

extern "C" {

SS_API int ss_score_batch(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr, const uint32_t* kw_terms,
                          const uint64_t* ph_ptr, const uint32_t* ph_terms, const double* topic_probs,
                          int32_t probs_per_query, uint32_t k, uint32_t* out_doc, double* out_final, double* out_pr,
                          uint32_t* out_count) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_score_batch: engine is NULL");
  SS_REQUIRE(n_q == 0 || (kw_ptr && out_doc && out_final && out_pr && out_count), SS_ERR_INVALID,
             "ss_score_batch: NULL argument");
  SS_REQUIRE(k >= 1 && k <= (uint32_t)kMaxK, SS_ERR_INVALID, "ss_score_batch: k = %u, supported 1..%d", k, kMaxK);
  SS_REQUIRE(n_q < 0x7FFFFFFFull, SS_ERR_INVALID, "ss_score_batch: batch too large");
  if (n_q == 0) return SS_OK;
  const uint64_t n_kw = kw_ptr[n_q], n_ph = ph_ptr ? ph_ptr[n_q] : 0;
  SS_REQUIRE((n_kw == 0 || kw_terms) && (n_ph == 0 || ph_terms), SS_ERR_INVALID, "ss_score_batch: NULL terms");
  for (uint64_t q = 0; q < n_q; ++q) {
    SS_REQUIRE(kw_ptr[q] <= kw_ptr[q + 1] && kw_ptr[q + 1] - kw_ptr[q] <= (uint64_t)kMaxKw, SS_ERR_INVALID,
               "ss_score_batch: query %llu has a bad keyword range (max %d tokens)", (unsigned long long)q, kMaxKw);
    if (ph_ptr) {
      SS_REQUIRE(ph_ptr[q] <= ph_ptr[q + 1], SS_ERR_INVALID, "ss_score_batch: ph_ptr not monotone");
      const uint64_t L = ph_ptr[q + 1] - ph_ptr[q];
      // > 256 tokens can never match (uint8 TermPos, phrase.go:115); 33..256 would, but is not supported
      SS_REQUIRE(L <= (uint64_t)kMaxPh || L > 256, SS_ERR_INVALID,
                 "ss_score_batch: query %llu has a %llu-token phrase (max %d)", (unsigned long long)q,
                 (unsigned long long)L, kMaxPh);
    }
  }
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  IndexState* ix = e->idx;
  SS_REQUIRE(ix && (ix->tab[0].loaded || ix->tab[1].loaded), SS_ERR_STATE, "ss_score_batch: no index loaded");
  for (int tb = 0; tb < 2; ++tb)
    SS_REQUIRE(!ix->tab[tb].loaded || ix->tab[tb].has_mag, SS_ERR_STATE,
               "ss_score_batch: table %d has no doc norms (ss_term_weights / ss_set_doc_norms)", tb);
  const uint64_t D = ix->D;
  const bool blend = topic_probs != nullptr;
  if (blend) {
    SS_REQUIRE(ix->pr.p && ix->T > 0, SS_ERR_STATE, "ss_score_batch: topic_probs given but no PageRank set");
    SS_REQUIRE(ix->pr_docs >= D, SS_ERR_STATE, "ss_score_batch: PageRank covers %llu docs, index has %llu",
               (unsigned long long)ix->pr_docs, (unsigned long long)D);
  }
  cudaStream_t st = e->stream;
  const bool timing = (e->flags & SS_FLAG_TIMING) != 0;
  uint32_t launches = 0;
  IndexState::Workspace& ws = ix->ws;

  // slabs: enough CTAs to fill the machine, index slice per slab around the L2 size,
  // and n_slabs * k small enough for the merge kernel
  const uint64_t n_sub = std::max<uint64_t>(1, (D + kRange - 1) / kRange);
  const uint64_t index_bytes = (ix->tab[0].P + ix->tab[1].P) * 8;
  uint64_t slab_bytes = 64ull << 20;
  if (const char* env = getenv("SS_SCORE_SLAB_MB")) slab_bytes = std::max(1ull, strtoull(env, nullptr, 10)) << 20;
  uint64_t n_slabs = std::max<uint64_t>(1, (index_bytes + slab_bytes - 1) / slab_bytes);
  const uint64_t want_ctas = (uint64_t)e->sm_count * 16;
  n_slabs = std::max(n_slabs, (want_ctas + n_q - 1) / n_q);
  n_slabs = std::min<uint64_t>(n_slabs, n_sub);
  // per-slab partial lists: n_q * n_slabs * k entries of 20 bytes, kept under 16 GB
  const uint64_t max_entries = std::max<uint64_t>(k, std::min<uint64_t>(kMergeMax, (16ull << 30) / (n_q * 20)));
  n_slabs = std::min<uint64_t>(n_slabs, std::max<uint64_t>(1, max_entries / k));
  uint64_t sub_per_slab = (n_sub + n_slabs - 1) / n_slabs;
  // the impact-vector path keeps survivors as 16-bit slab offsets: slabs of <= 65536 docs when the merge allows
  // (with smaller slabs k_plan merges them back into 65536-doc ranges for that path).  Measured: 32768-doc
  // slabs help queries with several mid-frequency lists but cost more than that elsewhere (216K vs 206K
  // queries/s on the benchmark mix), so 65536 is the default; SS_SCORE_SLAB_DOCS overrides.
  uint64_t slab_target = 65536;
  if (const char* env = getenv("SS_SCORE_SLAB_DOCS")) slab_target = std::max<uint64_t>(kRange, strtoull(env, nullptr, 10));
  for (uint64_t t : {slab_target, (uint64_t)65536}) {
    const uint64_t spt = std::max<uint64_t>(1, t / kRange);
    if (sub_per_slab > spt && (n_sub + spt - 1) / spt <= std::max<uint64_t>(1, max_entries / k)) {
      sub_per_slab = spt;
      break;
    }
  }
  n_slabs = (n_sub + sub_per_slab - 1) / sub_per_slab;
  SS_REQUIRE(n_q * n_slabs < 0x7FFFFFFFull, SS_ERR_INVALID, "ss_score_batch: batch too large; split it");

  // workspace (grow-only) and events: everything allocated before the timed region
  SS_TRY(ws_reserve(ws.kw_ptr, n_q + 1));
  SS_TRY(ws_reserve(ws.kw, n_kw));
  if (ph_ptr) {
    SS_TRY(ws_reserve(ws.ph_ptr, n_q + 1));
    SS_TRY(ws_reserve(ws.ph, n_ph));
  }
  if (blend) SS_TRY(ws_reserve(ws.probs, probs_per_query ? n_q * ix->T : ix->T));
  SS_TRY(ws_reserve(ws.part_doc, n_q * n_slabs * k));
  SS_TRY(ws_reserve(ws.part_final, n_q * n_slabs * k));
  SS_TRY(ws_reserve(ws.part_pr, n_q * n_slabs * k));
  SS_TRY(ws_reserve(ws.part_count, n_q * n_slabs));
  SS_TRY(ws_reserve(ws.out_doc, n_q * k));
  SS_TRY(ws_reserve(ws.out_final, n_q * k));
  SS_TRY(ws_reserve(ws.out_pr, n_q * k));
  SS_TRY(ws_reserve(ws.out_count, n_q));
  SS_TRY(ws_reserve(ws.stats, 2));
  SS_TRY(ws_reserve(ws.qthr, n_q));
  const uint64_t n_narrow = 2 * (n_kw + n_ph) * (n_slabs + 1);
  SS_TRY(ws_reserve(ws.narrow, n_narrow));
  // a table that was never loaded behaves as an empty one with zero norms
  if ((!ix->tab[0].loaded || !ix->tab[1].loaded) && ws.zero_mag.n < std::max<uint64_t>(D, 1)) {
    SS_TRY(ws.zero_mag.alloc(D));
    SS_CUDA(cudaMemsetAsync(ws.zero_mag.p, 0, std::max<uint64_t>(D, 1) * 8, st));
  }
  const bool shared_blend = blend && !probs_per_query;
  bool sqd_fresh = false;
  if (shared_blend) {
    sqd_fresh = ix->sqd_valid && ix->sqd_probs.size() == ix->T &&
                memcmp(ix->sqd_probs.data(), topic_probs, ix->T * 8) == 0;
    if (!sqd_fresh && ix->sqd.n < std::max<uint64_t>(D, 1)) SS_TRY(ix->sqd.alloc(D));
  }
  if (ix->meta32.n < std::max<uint64_t>(D, 1)) {
    SS_TRY(ix->meta32.alloc(D));
    ix->meta32_valid = false;
  }
  if (timing)
    for (auto& x : ws.ev)
      if (!x) SS_CUDA(cudaEventCreate(&x));
  SS_CUDA(cudaFuncSetAttribute(k_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));
  SS_CUDA(cudaFuncSetAttribute(k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(Smem)));

  if (timing) SS_CUDA(cudaEventRecord(ws.ev[0], st));
  SS_CUDA(cudaMemcpyAsync(ws.kw_ptr.p, kw_ptr, (n_q + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n_kw) SS_CUDA(cudaMemcpyAsync(ws.kw.p, kw_terms, n_kw * 4, cudaMemcpyHostToDevice, st));
  if (ph_ptr) {
    SS_CUDA(cudaMemcpyAsync(ws.ph_ptr.p, ph_ptr, (n_q + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_ph) SS_CUDA(cudaMemcpyAsync(ws.ph.p, ph_terms, n_ph * 4, cudaMemcpyHostToDevice, st));
  }
  SS_CUDA(cudaMemsetAsync(ws.stats.p, 0, 16, st));
  SS_CUDA(cudaMemsetAsync(ws.qthr.p, 0, n_q * 8, st));
  // blend term: one pass over forw[3] for a shared topic vector, cached across batches
  const double* sqd_ptr = nullptr;
  if (shared_blend) {
    if (!sqd_fresh) {
      SS_CUDA(cudaMemcpyAsync(ws.probs.p, topic_probs, ix->T * 8, cudaMemcpyHostToDevice, st));
      if (D) k_sqd<<<ss::div_up(D, 256), 256, 0, st>>>(ix->pr.p, ws.probs.p, ix->T, D, ix->sqd.p);
      ++launches;
      ix->sqd_probs.assign(topic_probs, topic_probs + ix->T);
      ix->sqd_valid = true;
    }
    sqd_ptr = ix->sqd.p;
  } else if (blend) {
    SS_CUDA(cudaMemcpyAsync(ws.probs.p, topic_probs, n_q * ix->T * 8, cudaMemcpyHostToDevice, st));
  }

  const double* mag0 = ix->tab[0].loaded ? ix->tab[0].mag.p : ws.zero_mag.p;
  const double* mag1 = ix->tab[1].loaded ? ix->tab[1].mag.p : ws.zero_mag.p;
  {
    // rebuilt whenever its inputs may have changed (cheap: one pass over D docs)
    const int mode = !blend ? 0 : (probs_per_query ? 2 : 1);
    const bool fresh = ix->meta32_valid && ix->meta32_mode == mode && (mode != 1 || sqd_fresh);
    if (!fresh) {
      if (D) k_meta32<<<ss::div_up(D, 256), 256, 0, st>>>(mag0, mag1, mode == 1 ? ix->sqd.p : nullptr,
                                                          mode == 2 ? ix->pr.p : nullptr, ix->T, D, ix->meta32.p);
      ++launches;
      ix->meta32_valid = true;
      ix->meta32_mode = mode;
      ix->zvec_valid = false;
    }
  }
  // impact vectors of the densest terms (SS_SCORE_DENSE=0 disables the path)
  bool use_dense = true;
  if (const char* env = getenv("SS_SCORE_DENSE")) use_dense = atoi(env) != 0;
  if (use_dense && D) SS_TRY(build_dense_vectors(e, ix, st, &launches));

  ScoreParams p{};
  p.sort_max = kSortMax;
  p.owner_path = 1;
  p.phrase_dense = 1;
  if (const char* env = getenv("SS_SCORE_PHRASE_DENSE")) p.phrase_dense = atoi(env);
  if (const char* env = getenv("SS_SCORE_OWNER")) p.owner_path = atoi(env);
  if (const char* env = getenv("SS_SCORE_SORT_MAX")) p.sort_max = std::min<uint32_t>(kSortMax, (uint32_t)atoi(env));
  p.meta32 = ix->meta32.p;
  p.tab[0] = view_of(ix->tab[0]);
  p.tab[1] = view_of(ix->tab[1]);
  p.mag[0] = mag0;
  p.mag[1] = mag1;
  p.sqd = sqd_ptr;
  p.pr = (blend && probs_per_query) ? ix->pr.p : nullptr;
  p.probs = (blend && probs_per_query) ? ws.probs.p : nullptr;
  p.T = ix->T;
  p.D = D;
  p.kw_ptr = ws.kw_ptr.p;
  p.kw_terms = ws.kw.p;
  p.ph_ptr = ph_ptr ? ws.ph_ptr.p : nullptr;
  p.ph_terms = ws.ph.p;
  p.n_q = (uint32_t)n_q;
  p.n_slabs = (uint32_t)n_slabs;
  p.k = k;
  p.slab_docs = sub_per_slab * kRange;
  p.part_doc = ws.part_doc.p;
  p.part_final = ws.part_final.p;
  p.part_pr = ws.part_pr.p;
  p.part_count = ws.part_count.p;
  p.stats = ws.stats.p;
  p.qthr = ws.qthr.p;
  if (use_dense && ix->dense_valid && ix->n_dense) {
    p.uvec = ix->uvec.p;
    p.zvec = ix->zvec.p;
    p.zblk = ix->zblk.p;
    p.dense_map = ix->dense_map.p;
    p.d_pad = ix->d_pad;
    p.dense_map_V = ix->dense_map_V;
  }
  p.use_qthr = 1;
  if (const char* env = getenv("SS_SCORE_QTHR")) p.use_qthr = atoi(env);

  p.narrow = ws.narrow.p;
  p.prefetch_meta = 0;  // measured: no effect (the finalize step is issue bound, not latency bound)
  if (const char* env = getenv("SS_SCORE_PREFETCH")) p.prefetch_meta = atoi(env);
  if (n_narrow) k_narrow<<<ss::div_up(n_narrow, 256), 256, 0, st>>>(p, ws.narrow.p, n_narrow);
  ++launches;
  {
    uint32_t merge_max = 3072;  // <= sort_max: merged groups take the sparse paths
    if (const char* env = getenv("SS_SCORE_MERGE_MAX")) merge_max = (uint32_t)std::max(0, atoi(env));
    merge_max = std::min(merge_max, p.sort_max);
    SS_TRY(ws_reserve(ws.group_len, n_q * n_slabs));
    SS_CUDA(cudaMemsetAsync(ws.part_count.p, 0, n_q * n_slabs * 4, st));
    k_plan<<<ss::div_up(n_q, 128), 128, 0, st>>>(p, merge_max, ws.group_len.p);
    p.group_len = ws.group_len.p;
    ++launches;
  }
  if (timing) SS_CUDA(cudaEventRecord(ws.ev[1], st));
  if (n_ph) k_score<true><<<(unsigned)(n_q * n_slabs), kT, sizeof(Smem), st>>>(p);
  else k_score<false><<<(unsigned)(n_q * n_slabs), kT, sizeof(Smem), st>>>(p);
  if (timing) SS_CUDA(cudaEventRecord(ws.ev[2], st));
  k_merge<<<(unsigned)n_q, kT, (size_t)n_slabs * 2, st>>>((uint32_t)n_slabs, k, (uint32_t)n_q, 0, ws.part_doc.p,
                                                              ws.part_final.p, ws.part_pr.p, ws.part_count.p,
                                                              ws.out_doc.p, ws.out_final.p, ws.out_pr.p,
                                                              ws.out_count.p);
  launches += 2;
  SS_CUDA(cudaMemcpyAsync(out_doc, ws.out_doc.p, n_q * k * 4, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_final, ws.out_final.p, n_q * k * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_pr, ws.out_pr.p, n_q * k * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_count, ws.out_count.p, n_q * 4, cudaMemcpyDeviceToHost, st));
  unsigned long long h_stats[2] = {0, 0};
  SS_CUDA(cudaMemcpyAsync(h_stats, ws.stats.p, 16, cudaMemcpyDeviceToHost, st));
  if (timing) SS_CUDA(cudaEventRecord(ws.ev[3], st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  ix->stats.postings_scanned = h_stats[0];
  ix->stats.docs_matched = h_stats[1];
  // SURVEY.md §8(d) B_q summed: 8 B per posting, per matched doc two norms + the blend
  // input this implementation reads (8 B cached sqd or the 8T-byte forw[3] row), 12 B per result
  const uint64_t per_doc = 16 + (blend ? (probs_per_query ? 8ull * ix->T : 8ull) : 0ull);
  ix->stats.algorithmic_bytes = 8ull * h_stats[0] + per_doc * h_stats[1] + 12ull * k * n_q;
  ix->stats.launches = launches;
  if (timing) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ws.ev[0], ws.ev[3]);  // H2D of the queries .. D2H of the results
    ix->stats.kernel_ms = ms;
    cudaEventElapsedTime(&ms, ws.ev[1], ws.ev[2]);
    ix->stats.score_kernel_ms = ms;
  }
  return SS_OK;
}

SS_API int ss_merge_topk(ss_engine* e, uint32_t n_lists, uint64_t n_q, uint32_t k, const uint32_t* docs,
                         const double* finals, const double* prs, const uint32_t* counts, uint32_t* out_doc,
                         double* out_final, double* out_pr, uint32_t* out_count) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_merge_topk: engine is NULL");
  SS_REQUIRE(n_q == 0 || (docs && finals && prs && counts && out_doc && out_final && out_pr && out_count),
             SS_ERR_INVALID, "ss_merge_topk: NULL argument");
  SS_REQUIRE(n_lists >= 1 && k >= 1 && (uint64_t)n_lists * k <= (uint64_t)kMergeMax, SS_ERR_INVALID,
             "ss_merge_topk: n_lists * k = %llu, max %d", (unsigned long long)n_lists * k, kMergeMax);
  if (n_q == 0) return SS_OK;
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  cudaStream_t st = e->stream;
  const size_t n = (size_t)n_lists * n_q * k;
  ss::DevBuf<uint32_t> d_doc, d_cnt, o_doc, o_cnt;
  ss::DevBuf<double> d_fin, d_pr, o_fin, o_pr;
  SS_TRY(d_doc.alloc(n));
  SS_TRY(d_fin.alloc(n));
  SS_TRY(d_pr.alloc(n));
  SS_TRY(d_cnt.alloc((size_t)n_lists * n_q));
  SS_TRY(o_doc.alloc(n_q * k));
  SS_TRY(o_fin.alloc(n_q * k));
  SS_TRY(o_pr.alloc(n_q * k));
  SS_TRY(o_cnt.alloc(n_q));
  SS_CUDA(cudaMemcpyAsync(d_doc.p, docs, n * 4, cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(d_fin.p, finals, n * 8, cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(d_pr.p, prs, n * 8, cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemcpyAsync(d_cnt.p, counts, (size_t)n_lists * n_q * 4, cudaMemcpyHostToDevice, st));
  k_merge<<<(unsigned)n_q, kT, (size_t)n_lists * 2, st>>>(n_lists, k, (uint32_t)n_q, 1, d_doc.p, d_fin.p, d_pr.p,
                                                              d_cnt.p, o_doc.p, o_fin.p, o_pr.p, o_cnt.p);
  SS_CUDA(cudaMemcpyAsync(out_doc, o_doc.p, n_q * k * 4, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_final, o_fin.p, n_q * k * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_pr, o_pr.p, n_q * k * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_count, o_cnt.p, n_q * 4, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

SS_API int ss_score_get_stats(ss_engine* e, ss_score_stats* out) {
  SS_REQUIRE(e && out, SS_ERR_INVALID, "ss_score_get_stats: NULL argument");
  std::lock_guard<std::mutex> lock(e->mu);
  SS_REQUIRE(e->idx, SS_ERR_STATE, "ss_score_get_stats: no index loaded");
  *out = e->idx->stats;
  return SS_OK;
}

}  // extern "C"
