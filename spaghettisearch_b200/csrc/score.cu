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

#include "comm.cuh"
#include "index.cuh"

namespace {

constexpr int kT = 256;        // threads per CTA
constexpr int kRange = 2048;   // docs per sub-range
constexpr int kCand = 1024;    // candidate buffer = doc slots finished per round
constexpr int kMaxK = 128;
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
  unsigned long long* gtop;   // [n_q][k] the k best keys of the slabs finished so far (k <= 32), or NULL
  uint32_t* glock;            // [n_q] lock of gtop[q]
  // impact vectors of the densest terms (see IndexState)
  const uint16_t* uvec;       // [n_dense][d_pad] fp16 bits, NULL = path disabled
  const uint16_t* zvec;       // [d_pad]
  const float* zblk;          // [d_pad / kRange] largest blend term of a doc block
  const uint8_t* dense_map;   // [dense_map_V]
  uint64_t d_pad, dense_map_V;
  // block maxima of the impact vectors: ublk[slot][block] = {largest impact, number of postings} of the
  // kDenseRange-doc block, NULL = no block skipping (SS_SCORE_BLOCKMAX=0 or slabs not block aligned)
  const float2* ublk;
  uint32_t n_blk;
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

// The kernel itself lives in score_kernels.inc and is compiled once per token-limit variant.
#define SS_LIM_NS lim_std
#define SS_LIM_KW 64
#define SS_LIM_PH 32
#define SS_LIM_OCC 3
#include "score_kernels.inc"
#undef SS_LIM_NS
#undef SS_LIM_KW
#undef SS_LIM_PH
#undef SS_LIM_OCC
#define SS_LIM_NS lim_wide
#define SS_LIM_KW 256
#define SS_LIM_PH 256
#define SS_LIM_OCC 2
#define SS_LIM_WIDE 1
#include "score_kernels.inc"
#undef SS_LIM_NS
#undef SS_LIM_KW
#undef SS_LIM_PH
#undef SS_LIM_OCC
#undef SS_LIM_WIDE
constexpr int kMaxKw = lim_std::kMaxKw, kMaxPh = lim_std::kMaxPh;
constexpr int kWideKw = lim_wide::kMaxKw, kWidePh = lim_wide::kMaxPh;

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
                                              double* __restrict__ out_pr, uint32_t* __restrict__ out_count,
                                              uint32_t doc_add) {
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
        out_doc[(size_t)q * k + step] = doc + doc_add;  // shard-local id -> global id (ss_index_set_doc_base)
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
// {largest impact, number of postings} of every kDenseRange-doc block of every impact vector
// (grid: blocks x dense slots): what lets the stream skip a block it cannot find a candidate in
__global__ void k_dense_blk(const uint16_t* __restrict__ uvec, uint64_t d_pad, uint32_t n_blk, float2* __restrict__ ublk) {
  __shared__ float smax[256 / 32];
  __shared__ uint32_t scnt[256 / 32];
  const uint16_t* u = uvec + (size_t)blockIdx.y * d_pad + (size_t)blockIdx.x * kDenseRange;
  float m = 0.0f;
  uint32_t c = 0;
  for (uint32_t i = threadIdx.x; i < (uint32_t)kDenseRange; i += blockDim.x) {
    const uint16_t h = u[i];
    c += h != 0;
    m = fmaxf(m, __half2float(__ushort_as_half(h)));  // impacts are >= 0, never NaN (k_dense_fill)
  }
  for (int o = 16; o; o >>= 1) {
    m = fmaxf(m, __shfl_xor_sync(0xFFFFFFFFu, m, o));
    c += __shfl_xor_sync(0xFFFFFFFFu, c, o);
  }
  if ((threadIdx.x & 31) == 0) {
    smax[threadIdx.x >> 5] = m;
    scnt[threadIdx.x >> 5] = c;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 256 / 32; ++w) {
      m = fmaxf(m, smax[w]);
      c += scnt[w];
    }
    ublk[(size_t)blockIdx.y * n_blk + blockIdx.x] = make_float2(m, (float)c);
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
    ix->dense_host.clear();
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
        const uint32_t n_blk = (uint32_t)(d_pad / kDenseRange);
        SS_TRY(ws_reserve(ix->ublk, (size_t)nd * n_blk));
        k_dense_blk<<<dim3(n_blk, nd), 256, 0, st>>>(ix->uvec.p, d_pad, n_blk, ix->ublk.p);
        SS_CUDA(cudaStreamSynchronize(st));  // d_term goes out of scope
        SS_CUDA(cudaGetLastError());
        *launches += 5;
      }
      ix->n_dense = nd;
      ix->dense_map_V = V;
      ix->dense_host.assign(V, 0);
      for (uint32_t i = 0; i < nd; ++i) ix->dense_host[chosen[i]] = 1;
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

// ss_score_batch / ss_score_batch_sharded.  sharded: after the local top-k every rank all-gathers the
// [n_q][k] lists over NCCL and merges them with k_merge on its own stream, so
// every rank returns the global top-k.
// One batch whose queries all fit the chosen kernel variant (`wide`: lim_wide, accumulator path only).
// Called with the engine lock held.
static int score_batch_core(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr, const uint32_t* kw_terms,
                            const uint64_t* ph_ptr, const uint32_t* ph_terms, const double* topic_probs,
                            int32_t probs_per_query, uint32_t k, uint32_t* out_doc, double* out_final, double* out_pr,
                            uint32_t* out_count, bool sharded, bool wide) {
  const uint64_t n_kw = kw_ptr[n_q], n_ph = ph_ptr ? ph_ptr[n_q] : 0;
  IndexState* ix = e->idx;
  SS_REQUIRE(ix && (ix->tab[0].loaded || ix->tab[1].loaded), SS_ERR_STATE, "ss_score_batch: no index loaded");
  for (int tb = 0; tb < 2; ++tb)
    SS_REQUIRE(!ix->tab[tb].loaded || ix->tab[tb].has_mag, SS_ERR_STATE,
               "ss_score_batch: table %d has no doc norms (ss_term_weights / ss_set_doc_norms)", tb);
  const uint64_t D = ix->D;
  SS_REQUIRE(D + ix->doc_base <= 0xFFFFFFFFull, SS_ERR_INVALID, "ss_score_batch: doc base %llu + %llu docs exceed 32 bit",
             (unsigned long long)ix->doc_base, (unsigned long long)D);
  const int world = sharded ? comm_world(e) : 1;
  SS_REQUIRE((uint64_t)world * k <= (uint64_t)kMergeMax, SS_ERR_INVALID, "ss_score_batch_sharded: world * k too large");
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
  SS_TRY(ws_reserve(ws.gtop, n_q * std::min<uint32_t>(k, 32)));
  SS_TRY(ws_reserve(ws.glock, n_q));
  if (world > 1) {
    SS_TRY(ws_reserve(ws.all_doc, (size_t)world * n_q * k));
    SS_TRY(ws_reserve(ws.all_final, (size_t)world * n_q * k));
    SS_TRY(ws_reserve(ws.all_pr, (size_t)world * n_q * k));
    SS_TRY(ws_reserve(ws.all_count, (size_t)world * n_q));
    SS_TRY(ws_reserve(ws.loc_doc, n_q * k));
    SS_TRY(ws_reserve(ws.loc_final, n_q * k));
    SS_TRY(ws_reserve(ws.loc_pr, n_q * k));
    SS_TRY(ws_reserve(ws.loc_count, n_q));
  }
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
  SS_CUDA(cudaFuncSetAttribute(lim_std::k_score<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(lim_std::Smem)));
  SS_CUDA(cudaFuncSetAttribute(lim_std::k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(lim_std::Smem)));
  SS_CUDA(cudaFuncSetAttribute(lim_wide::k_score<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)sizeof(lim_wide::Smem)));

  if (timing) SS_CUDA(cudaEventRecord(ws.ev[0], st));
  SS_CUDA(cudaMemcpyAsync(ws.kw_ptr.p, kw_ptr, (n_q + 1) * 8, cudaMemcpyHostToDevice, st));
  if (n_kw) SS_CUDA(cudaMemcpyAsync(ws.kw.p, kw_terms, n_kw * 4, cudaMemcpyHostToDevice, st));
  if (ph_ptr) {
    SS_CUDA(cudaMemcpyAsync(ws.ph_ptr.p, ph_ptr, (n_q + 1) * 8, cudaMemcpyHostToDevice, st));
    if (n_ph) SS_CUDA(cudaMemcpyAsync(ws.ph.p, ph_terms, n_ph * 4, cudaMemcpyHostToDevice, st));
  }
  SS_CUDA(cudaMemsetAsync(ws.stats.p, 0, 16, st));
  SS_CUDA(cudaMemsetAsync(ws.qthr.p, 0, n_q * 8, st));
  bool global_topk = k <= 32;
  if (const char* env = getenv("SS_SCORE_GTOP")) global_topk = global_topk && atoi(env) != 0;
  if (global_topk) {
    SS_CUDA(cudaMemsetAsync(ws.gtop.p, 0, n_q * k * 8, st));
    SS_CUDA(cudaMemsetAsync(ws.glock.p, 0, n_q * 4, st));
  }
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
  bool use_dense = !wide;
  if (const char* env = getenv("SS_SCORE_DENSE")) use_dense = use_dense && atoi(env) != 0;
  if (use_dense && D) SS_TRY(build_dense_vectors(e, ix, st, &launches));

  ScoreParams p{};
  p.sort_max = lim_std::kSortMax;
  p.owner_path = 1;
  p.phrase_dense = 1;
  if (const char* env = getenv("SS_SCORE_PHRASE_DENSE")) p.phrase_dense = atoi(env);
  if (const char* env = getenv("SS_SCORE_OWNER")) p.owner_path = atoi(env);
  if (const char* env = getenv("SS_SCORE_SORT_MAX"))
    p.sort_max = std::min<uint32_t>(lim_std::kSortMax, (uint32_t)atoi(env));
  if (wide) p.sort_max = 0;  // the sparse paths tag postings with an 8-bit list number: accumulator path only
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
    bool block_max = true;
    if (const char* env = getenv("SS_SCORE_BLOCKMAX")) block_max = atoi(env) != 0;
    if (block_max && (sub_per_slab * kRange) % kDenseRange == 0) {  // slab starts are block boundaries
      p.ublk = ix->ublk.p;
      p.n_blk = (uint32_t)(ix->d_pad / kDenseRange);
    }
  }
  p.use_qthr = 1;
  if (const char* env = getenv("SS_SCORE_QTHR")) p.use_qthr = atoi(env);
  if (global_topk && p.use_qthr) {
    p.gtop = ws.gtop.p;
    p.glock = ws.glock.p;
  }

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
  if (wide) lim_wide::k_score<true><<<(unsigned)(n_q * n_slabs), kT, sizeof(lim_wide::Smem), st>>>(p);
  else if (n_ph) lim_std::k_score<true><<<(unsigned)(n_q * n_slabs), kT, sizeof(lim_std::Smem), st>>>(p);
  else lim_std::k_score<false><<<(unsigned)(n_q * n_slabs), kT, sizeof(lim_std::Smem), st>>>(p);
  if (timing) SS_CUDA(cudaEventRecord(ws.ev[2], st));
  k_merge<<<(unsigned)n_q, kT, (size_t)n_slabs * 2, st>>>((uint32_t)n_slabs, k, (uint32_t)n_q, 0, ws.part_doc.p,
                                                              ws.part_final.p, ws.part_pr.p, ws.part_count.p,
                                                              ws.out_doc.p, ws.out_final.p, ws.out_pr.p,
                                                              ws.out_count.p, (uint32_t)ix->doc_base);
  launches += 2;
  const uint32_t* r_doc = ws.out_doc.p;
  const double *r_final = ws.out_final.p, *r_pr = ws.out_pr.p;
  const uint32_t* r_count = ws.out_count.p;
  if (timing) SS_CUDA(cudaEventRecord(ws.ev[4], st));
  if (world > 1) {
    // doc-sharded index: every rank contributes its local top-k (already global doc ids) and merges all
    // of them with the same comparator
    const void* in[4] = {ws.out_doc.p, ws.out_final.p, ws.out_pr.p, ws.out_count.p};
    void* all[4] = {ws.all_doc.p, ws.all_final.p, ws.all_pr.p, ws.all_count.p};
    const size_t bytes[4] = {n_q * k * 4, n_q * k * 8, n_q * k * 8, n_q * 4};
    SS_TRY(comm_allgather_dev(e, 4, in, all, bytes));
    k_merge<<<(unsigned)n_q, kT, (size_t)world * 2, st>>>((uint32_t)world, k, (uint32_t)n_q, 1, ws.all_doc.p,
                                                             ws.all_final.p, ws.all_pr.p, ws.all_count.p, ws.loc_doc.p,
                                                             ws.loc_final.p, ws.loc_pr.p, ws.loc_count.p, 0u);
    ++launches;
    r_doc = ws.loc_doc.p;
    r_final = ws.loc_final.p;
    r_pr = ws.loc_pr.p;
    r_count = ws.loc_count.p;
  }
  SS_CUDA(cudaMemcpyAsync(out_doc, r_doc, n_q * k * 4, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_final, r_final, n_q * k * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_pr, r_pr, n_q * k * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(out_count, r_count, n_q * 4, cudaMemcpyDeviceToHost, st));
  unsigned long long h_stats[2] = {0, 0};
  SS_CUDA(cudaMemcpyAsync(h_stats, ws.stats.p, 16, cudaMemcpyDeviceToHost, st));
  if (timing) SS_CUDA(cudaEventRecord(ws.ev[3], st));
  // Byte model of what this batch has to move on the path it takes (host arithmetic while the kernels run):
  // a query with a dense keyword streams 2 B per doc for each of its dense tokens (the blend bound is only read
  // for the few 8-doc groups that pass the first test) and reads 8 B per posting of its other tokens; any other
  // query reads 8 B per posting of all its lists.
  uint64_t model_bytes = 12ull * k * n_q;
  {
    const bool dense_on = use_dense && ix->dense_valid && ix->n_dense && !ix->dense_host.empty();
    auto df_of = [&](uint32_t term) -> uint64_t {
      uint64_t d = 0;
      for (int tb = 0; tb < 2; ++tb)
        if (ix->tab[tb].loaded && term < ix->tab[tb].df_host.size()) d += ix->tab[tb].df_host[term];
      return d;
    };
    for (uint64_t q = 0; q < n_q; ++q) {
      uint64_t n_dense_tok = 0, sparse_postings = 0, all_postings = 0;
      bool dense_kw = false;
      auto visit = [&](uint32_t term, bool is_kw) {
        const uint64_t df = df_of(term);
        all_postings += df;
        if (dense_on && term < ix->dense_host.size() && ix->dense_host[term]) {
          ++n_dense_tok;
          dense_kw = dense_kw || is_kw;
        } else {
          sparse_postings += df;
        }
      };
      for (uint64_t i = kw_ptr[q]; i < kw_ptr[q + 1]; ++i) visit(kw_terms[i], true);
      if (ph_ptr)
        for (uint64_t i = ph_ptr[q]; i < ph_ptr[q + 1]; ++i) visit(ph_terms[i], false);
      model_bytes += dense_kw ? 2ull * D * n_dense_tok + 8ull * sparse_postings : 8ull * all_postings;
    }
  }
  ix->stats.model_bytes = model_bytes;
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
    cudaEventElapsedTime(&ms, ws.ev[4], ws.ev[3]);  // cross-shard all-gather + merge + D2H of the results
    ix->stats.shard_merge_ms = ms;
  }
  return SS_OK;
}

// Validates the batch and routes it: queries within the standard limits (64 keyword tokens, 32 phrase tokens)
// go through lim_std in one launch; the rare over-long ones (up to 256 / 256; the reference evaluates phrases
// of up to 256 tokens, retrieval/phrase.go:111-118) are scored as a second sub-batch by lim_wide and their
// rows scattered back, so one such query no longer fails everybody else's.
static int score_batch_entry(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr, const uint32_t* kw_terms,
                             const uint64_t* ph_ptr, const uint32_t* ph_terms, const double* topic_probs,
                             int32_t probs_per_query, uint32_t k, uint32_t* out_doc, double* out_final,
                             double* out_pr, uint32_t* out_count, bool sharded) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_score_batch: engine is NULL");
  SS_REQUIRE(n_q == 0 || (kw_ptr && out_doc && out_final && out_pr && out_count), SS_ERR_INVALID,
             "ss_score_batch: NULL argument");
  SS_REQUIRE(k >= 1 && k <= (uint32_t)kMaxK, SS_ERR_INVALID, "ss_score_batch: k = %u, supported 1..%d", k, kMaxK);
  SS_REQUIRE(n_q < 0x7FFFFFFFull, SS_ERR_INVALID, "ss_score_batch: batch too large");
  if (n_q == 0) return SS_OK;
  const uint64_t n_kw = kw_ptr[n_q], n_ph = ph_ptr ? ph_ptr[n_q] : 0;
  SS_REQUIRE((n_kw == 0 || kw_terms) && (n_ph == 0 || ph_terms), SS_ERR_INVALID, "ss_score_batch: NULL terms");
  std::vector<uint64_t> wide_q;
  for (uint64_t q = 0; q < n_q; ++q) {
    SS_REQUIRE(kw_ptr[q] <= kw_ptr[q + 1], SS_ERR_INVALID, "ss_score_batch: kw_ptr not monotone at query %llu",
               (unsigned long long)q);
    const uint64_t nk = kw_ptr[q + 1] - kw_ptr[q];
    SS_REQUIRE(nk <= (uint64_t)kWideKw, SS_ERR_INVALID, "ss_score_batch: query %llu has %llu keyword tokens (max %d)",
               (unsigned long long)q, (unsigned long long)nk, kWideKw);
    uint64_t L = 0;
    if (ph_ptr) {
      SS_REQUIRE(ph_ptr[q] <= ph_ptr[q + 1], SS_ERR_INVALID, "ss_score_batch: ph_ptr not monotone");
      L = ph_ptr[q + 1] - ph_ptr[q];
    }
    // a phrase of more than 256 tokens can never match (uint8 TermPos, phrase.go:115): the kernel drops it
    // (its tokens still count in queryLength, main_retrieve.go:90)
    if (nk > (uint64_t)kMaxKw || (L > (uint64_t)kMaxPh && L <= (uint64_t)kWidePh)) wide_q.push_back(q);
  }
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  if (wide_q.empty())
    return score_batch_core(e, n_q, kw_ptr, kw_terms, ph_ptr, ph_terms, topic_probs, probs_per_query, k, out_doc,
                            out_final, out_pr, out_count, sharded, false);
  // split: sub-batch 0 = standard queries, sub-batch 1 = wide queries
  IndexState* ix = e->idx;
  const uint32_t T = ix ? ix->T : 0;
  ss_score_stats total{};
  size_t wi = 0;
  std::vector<uint64_t> members[2];
  for (uint64_t q = 0; q < n_q; ++q) {
    const bool w = wi < wide_q.size() && wide_q[wi] == q;
    if (w) ++wi;
    members[w ? 1 : 0].push_back(q);
  }
  for (int part = 0; part < 2; ++part) {
    const std::vector<uint64_t>& m = members[part];
    if (m.empty()) continue;
    const uint64_t nq = m.size();
    std::vector<uint64_t> s_kw_ptr(nq + 1, 0), s_ph_ptr(nq + 1, 0);
    std::vector<uint32_t> s_kw, s_ph;
    std::vector<double> s_probs;
    for (uint64_t i = 0; i < nq; ++i) {
      const uint64_t q = m[i];
      s_kw.insert(s_kw.end(), kw_terms + kw_ptr[q], kw_terms + kw_ptr[q + 1]);
      s_kw_ptr[i + 1] = s_kw.size();
      if (ph_ptr) s_ph.insert(s_ph.end(), ph_terms + ph_ptr[q], ph_terms + ph_ptr[q + 1]);
      s_ph_ptr[i + 1] = s_ph.size();
      if (topic_probs && probs_per_query) s_probs.insert(s_probs.end(), topic_probs + q * T, topic_probs + (q + 1) * T);
    }
    std::vector<uint32_t> o_doc(nq * k), o_cnt(nq);
    std::vector<double> o_fin(nq * k), o_pr(nq * k);
    const double* probs = topic_probs ? (probs_per_query ? s_probs.data() : topic_probs) : nullptr;
    SS_TRY(score_batch_core(e, nq, s_kw_ptr.data(), s_kw.data(), ph_ptr ? s_ph_ptr.data() : nullptr, s_ph.data(), probs,
                            probs_per_query, k, o_doc.data(), o_fin.data(), o_pr.data(), o_cnt.data(), sharded,
                            part == 1));
    for (uint64_t i = 0; i < nq; ++i) {
      const uint64_t q = m[i];
      memcpy(out_doc + q * k, o_doc.data() + i * k, k * 4);
      memcpy(out_final + q * k, o_fin.data() + i * k, k * 8);
      memcpy(out_pr + q * k, o_pr.data() + i * k, k * 8);
      out_count[q] = o_cnt[i];
    }
    const ss_score_stats& st = e->idx->stats;
    total.postings_scanned += st.postings_scanned;
    total.docs_matched += st.docs_matched;
    total.algorithmic_bytes += st.algorithmic_bytes;
    total.launches += st.launches;
    total.kernel_ms += st.kernel_ms;
    total.score_kernel_ms += st.score_kernel_ms;
    total.shard_merge_ms += st.shard_merge_ms;
    total.model_bytes += st.model_bytes;
  }
  e->idx->stats = total;
  return SS_OK;
}

// Very large batches are scored in slices: the per-slab partial lists (n_q * n_slabs * k entries) are capped, and
// beyond ~250K queries the cap would force slabs wider than the 65536 docs the impact-vector path is built for.
// BASELINE configs[4] (a 1M-query batch) therefore runs as five slices inside one call; results and statistics
// are those of the whole batch.
static uint64_t max_slice() {
  uint64_t n = 200000;
  if (const char* env = getenv("SS_SCORE_MAX_SLICE")) n = std::max<uint64_t>(1, strtoull(env, nullptr, 10));
  return n;
}
static int score_batch_sliced(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr, const uint32_t* kw_terms,
                              const uint64_t* ph_ptr, const uint32_t* ph_terms, const double* topic_probs,
                              int32_t probs_per_query, uint32_t k, uint32_t* out_doc, double* out_final,
                              double* out_pr, uint32_t* out_count, bool sharded) {
  const uint64_t kMaxSlice = max_slice();
  if (n_q <= kMaxSlice || !kw_ptr || !out_doc || !out_final || !out_pr || !out_count || !e)
    return score_batch_entry(e, n_q, kw_ptr, kw_terms, ph_ptr, ph_terms, topic_probs, probs_per_query, k, out_doc,
                             out_final, out_pr, out_count, sharded);
  for (uint64_t q = 0; q < n_q; ++q) {  // the slices rebase the offsets: check them first
    SS_REQUIRE(kw_ptr[q] <= kw_ptr[q + 1], SS_ERR_INVALID, "ss_score_batch: kw_ptr not monotone at query %llu",
               (unsigned long long)q);
    SS_REQUIRE(!ph_ptr || ph_ptr[q] <= ph_ptr[q + 1], SS_ERR_INVALID, "ss_score_batch: ph_ptr not monotone");
  }
  ss_score_stats total{};
  const uint32_t T = e->idx ? e->idx->T : 0;
  std::vector<uint64_t> s_kw(kMaxSlice + 1), s_ph(kMaxSlice + 1);
  for (uint64_t lo = 0; lo < n_q; lo += kMaxSlice) {
    const uint64_t n = std::min(kMaxSlice, n_q - lo);
    for (uint64_t i = 0; i <= n; ++i) {
      s_kw[i] = kw_ptr[lo + i] - kw_ptr[lo];
      if (ph_ptr) s_ph[i] = ph_ptr[lo + i] - ph_ptr[lo];
    }
    const double* probs = topic_probs ? (probs_per_query ? topic_probs + lo * T : topic_probs) : nullptr;
    SS_TRY(score_batch_entry(e, n, s_kw.data(), kw_terms ? kw_terms + kw_ptr[lo] : nullptr, ph_ptr ? s_ph.data() : nullptr,
                             (ph_ptr && ph_terms) ? ph_terms + ph_ptr[lo] : nullptr, probs, probs_per_query, k,
                             out_doc + lo * k, out_final + lo * k, out_pr + lo * k, out_count + lo, sharded));
    std::lock_guard<std::mutex> lock(e->mu);
    const ss_score_stats& st = e->idx->stats;
    total.postings_scanned += st.postings_scanned;
    total.docs_matched += st.docs_matched;
    total.algorithmic_bytes += st.algorithmic_bytes;
    total.model_bytes += st.model_bytes;
    total.launches += st.launches;
    total.kernel_ms += st.kernel_ms;
    total.score_kernel_ms += st.score_kernel_ms;
    total.shard_merge_ms += st.shard_merge_ms;
  }
  std::lock_guard<std::mutex> lock(e->mu);
  e->idx->stats = total;
  return SS_OK;
}

extern "C" {

SS_API int ss_score_batch(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr, const uint32_t* kw_terms,
                          const uint64_t* ph_ptr, const uint32_t* ph_terms, const double* topic_probs,
                          int32_t probs_per_query, uint32_t k, uint32_t* out_doc, double* out_final, double* out_pr,
                          uint32_t* out_count) {
  return score_batch_sliced(e, n_q, kw_ptr, kw_terms, ph_ptr, ph_terms, topic_probs, probs_per_query, k, out_doc,
                            out_final, out_pr, out_count, false);
}

SS_API int ss_score_batch_sharded(ss_engine* e, uint64_t n_q, const uint64_t* kw_ptr, const uint32_t* kw_terms,
                                  const uint64_t* ph_ptr, const uint32_t* ph_terms, const double* topic_probs,
                                  int32_t probs_per_query, uint32_t k, uint32_t* out_doc, double* out_final,
                                  double* out_pr, uint32_t* out_count) {
  return score_batch_sliced(e, n_q, kw_ptr, kw_terms, ph_ptr, ph_terms, topic_probs, probs_per_query, k, out_doc,
                            out_final, out_pr, out_count, true);
}

SS_API int ss_merge_topk(ss_engine* e, uint32_t n_lists, uint64_t n_q, uint32_t k, const uint32_t* docs,
                         const double* finals, const double* prs, const uint32_t* counts, uint32_t* out_doc,
                         double* out_final, double* out_pr, uint32_t* out_count) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_merge_topk: engine is NULL");
  SS_REQUIRE(n_q == 0 || (docs && finals && prs && counts && out_doc && out_final && out_pr && out_count),
             SS_ERR_INVALID, "ss_merge_topk: NULL argument");
  // k_merge keeps each list's head and length in one byte: k <= kMaxK (128), as in ss_score_batch
  SS_REQUIRE(k >= 1 && k <= (uint32_t)kMaxK, SS_ERR_INVALID, "ss_merge_topk: k = %u, supported 1..%d", k, kMaxK);
  SS_REQUIRE(n_lists >= 1 && (uint64_t)n_lists * k <= (uint64_t)kMergeMax, SS_ERR_INVALID,
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
                                                              d_cnt.p, o_doc.p, o_fin.p, o_pr.p, o_cnt.p, 0u);
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
