// HP-1: the reference's "topic-sensitive" PageRank (ranking/pagerank.go:14-145)
// as a pull SpMM over the in-edge lists, all topics advanced together.
//
// Reference update for one topic (pagerank.go:93-119), last = previous ranks:
//   w_p      = d * last[p] / out(p)            for every parent with out(p) > 0
//   inh[c]   = sum of w_p over in-edges p -> c  (+ 1/n in the very first sweep,
//              because iteration 1 accumulates on top of the initial value)
//   Tot      = sum_p w_p + (1-d) * N            (each parent counted ONCE)
//   cur[v]   = (inh[v] + (1-d)) / Tot ;  delta = sum_v |cur[v] - last[v]|
// There is no dangling-mass term and teleport is uniform for every topic; the
// topic enters only through the start value 1/numPages[t].
//
// Device layout (per rank): the state is kept PRE-SCALED,
//   y[v][t] = rank[v][t] * m(v),  m(v) = d/out(v) if out(v) > 0 else 1,
// row major [N][TP] fp64 (TP = topics padded to 2/4/8/16; 128-byte rows at
// T = 16), so that an in-edge gather is one aligned row load with no per-edge
// scale and the normaliser is S_t = sum over non-dangling rows of y[.][t].
// One sweep reads col-idx once, gathers E rows of y_last, streams y_last and
// y_next once: the algorithmic bytes of SURVEY.md §8(d).
//
// Work split: rows with in-degree <= kShortMax are handled one per sub-warp
// (LPR lanes, 16 bytes per lane) by one of three bit-identical kernels --
// k_sweep_short32 (default: 32-bit padded row pointers, packed scale record, no
// spills), k_sweep_short (first version; 64-bit-pointer fallback) and
// k_sweep_short_async (opt-in: gathered rows land in a cp.async shared-memory
// ring); longer rows are cut into kChunk-edge tasks,
// one warp each, ordered by first source id so that concurrently running tasks
// gather from neighbouring source ranges (L2 reuse); rows spanning several
// tasks are finished by a fix-up pass that sums their partial rows in task
// order.  Every reduction has a fixed shape, so results are deterministic.
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "comm.cuh"
#include "common.cuh"

namespace {

#ifndef SS_GATHER_EVICT_LAST
#define SS_GATHER_EVICT_LAST 0
#endif
constexpr uint32_t kShortMax = 16;   // longest row handled by one sub-warp (the async ring holds 32/LPR such rows + 31)
constexpr uint32_t kChunk = 512;     // edges per long-row task
constexpr int kThreads = 256;
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct LongTask {
  uint64_t e_begin;  // offset into in_src
  uint32_t row;      // local destination row
  uint32_t n;        // edges in this task
  int32_t slot;      // partial-row slot, or -1 when the task covers the whole row
  uint32_t pad;
};
struct FixRow {
  uint32_t row, first_slot, n_slots, pad;
};

constexpr int kMaxPeers = 7;  // fused exchange up to 8 GPUs
struct SweepParams {
  const double* y_last;
  double* y_next;
  double* peer_next[kMaxPeers];  // the same y_next buffer on the other ranks (peer memory), fused exchange
  int n_peers;
  const uint64_t* in_ptr;   // [rows_loc + 1], offsets into in_src
  const uint32_t* in_src;   // sources ascending within a row
  const double* mul;        // [rows_loc] d/out, or 0 for dangling rows
  const uint32_t* in_ptr32; // in_ptr as 32-bit offsets, padded with E_loc (fast short-row kernel), or NULL
  const double2* mul2;      // [rows_loc] {scale m (1 for dangling rows; 0 marks a long row), 1/m (-1: dangling)}
  const uint32_t* sptr;     // [rows_loc + 1 (+pad)] row pointers into ssrc: short rows only, long rows are empty
  const uint32_t* ssrc;     // sources of the short rows, compacted in row order
  const uint32_t* warp_rows;  // [n_warps + 1] row range of every warp of the async short-row kernel
  const double* inv_tot;    // [TP] 1 / (S_t + (1-d) N)
  const double* init;       // [TP] 1/num_pages[t]
  const double* tele_w;     // [rows_loc][TP] teleport weights N * v_t[v] of this rank's rows, or NULL = uniform
  double* red;              // [slots][3*TP] per-CTA partial sums (delta, S, changed)
  uint64_t row_lo;          // global id of local row 0
  uint32_t row_begin;       // first local row of this launch's chunk (lean short-row kernel)
  uint32_t rows_loc;        // one past the last local row of this launch's chunk
  uint32_t active_mask;     // bit t set: topic t still iterating
  double tele;              // 1 - d
  int first;                // sweep 1: add 1/n, compare against 1/n
};

// ---- row access -------------------------------------------------------------
// A lane owns VEC consecutive topics of a row: VEC = 2 is a 128-bit access,
// VEC = 4 the 256-bit LDG/STG that sm_100 adds (LDG.E.ENL2.256), which halves
// the lanes per 128-byte row and doubles the rows a warp keeps in flight.
template <int VEC>
struct Vec {
  double v[VEC];
};
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_row_gather(const double* p);
template <>
__device__ __forceinline__ Vec<2> ld_row_gather<2>(const double* p) {
  const double2 t = __ldg(reinterpret_cast<const double2*>(p));
  return Vec<2>{{t.x, t.y}};
}
template <>
__device__ __forceinline__ Vec<4> ld_row_gather<4>(const double* p) {
  Vec<4> r;
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
               : "l"(p));
  return r;
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_row_stream(const double* p);
template <>
__device__ __forceinline__ Vec<2> ld_row_stream<2>(const double* p) {
  const double2 t = __ldcs(reinterpret_cast<const double2*>(p));
  return Vec<2>{{t.x, t.y}};
}
template <>
__device__ __forceinline__ Vec<4> ld_row_stream<4>(const double* p) {
  Vec<4> r;
  asm volatile("ld.global.cs.v4.f64 {%0, %1, %2, %3}, [%4];"
               : "=d"(r.v[0]), "=d"(r.v[1]), "=d"(r.v[2]), "=d"(r.v[3])
               : "l"(p));
  return r;
}
template <int VEC>
__device__ __forceinline__ void st_row_stream(double* p, const Vec<VEC>& v);
template <>
__device__ __forceinline__ void st_row_stream<2>(double* p, const Vec<2>& v) {
  __stcs(reinterpret_cast<double2*>(p), make_double2(v.v[0], v.v[1]));
}
template <>
__device__ __forceinline__ void st_row_stream<4>(double* p, const Vec<4>& v) {
  asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(p), "d"(v.v[0]), "d"(v.v[1]), "d"(v.v[2]),
               "d"(v.v[3])
               : "memory");
}
template <int VEC>
__device__ __forceinline__ void st_row_plain(double* p, const Vec<VEC>& v) {
#pragma unroll
  for (int j = 0; j < VEC; j += 2) *reinterpret_cast<double2*>(p + j) = make_double2(v.v[j], v.v[j + 1]);
}
template <int VEC>
__device__ __forceinline__ Vec<VEC> ld_row_plain(const double* p) {
  Vec<VEC> r;
#pragma unroll
  for (int j = 0; j < VEC; j += 2) {
    const double2 t = *reinterpret_cast<const double2*>(p + j);
    r.v[j] = t.x;
    r.v[j + 1] = t.y;
  }
  return r;
}

template <int VEC>
struct Acc {  // per-thread running sums for its VEC topic columns
  double d[VEC], s[VEC], c[VEC];
  __device__ Acc() {
#pragma unroll
    for (int j = 0; j < VEC; ++j) d[j] = s[j] = c[j] = 0.0;
  }
};

// Fused normalise / teleport / residual for one finished row; executed by the
// LPR lanes that own the row, lane l8 holding topics VEC*l8 .. VEC*l8+VEC-1.  The
// row's own previous value yl and scale m are loaded by the caller (early, so
// that they overlap the gathers).
template <int LPR, int VEC>
__device__ __forceinline__ void epilogue(const SweepParams& p, uint32_t r, int l8, Vec<VEC> a, const Vec<VEC>& yl,
                                         double m, Acc<VEC>& acc) {
  constexpr int TP = LPR * VEC;
  const uint64_t v = p.row_lo + r;
  const bool has_out = m > 0.0;
  const double mul = has_out ? m : 1.0;
  // One reciprocal per row instead of two fp64 divides per topic: the previous rank
  // (only used for the residual) is y * (1/m), and the new rank multiplies by the
  // per-topic 1/Tot computed once per sweep.  Both differ from the divide by <= 1 ulp,
  // five orders of magnitude inside the 1e-9 L1 budget.
  const double inv_mul = has_out ? 1.0 / m : 1.0;
  const Vec<VEC> inv_tot = ld_row_plain<VEC>(p.inv_tot + VEC * l8);
  Vec<VEC> tw;  // teleport term per topic: (1-d), or (1-d) * N * v_t[v] for a topic-biased run
#pragma unroll
  for (int j = 0; j < VEC; ++j) tw.v[j] = p.tele;
  if (p.tele_w) {
    const Vec<VEC> w = ld_row_stream<VEC>(p.tele_w + (uint64_t)r * TP + VEC * l8);
#pragma unroll
    for (int j = 0; j < VEC; ++j) tw.v[j] = __dmul_rn(p.tele, w.v[j]);
  }
  Vec<VEC> yn;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    const int t = VEC * l8 + j;
    double last_rank = yl.v[j] * inv_mul;
    if (p.first) {
      const double i0 = p.init[t];
      a.v[j] += i0;
      last_rank = i0;
    }
    if ((p.active_mask >> t) & 1u) {
      const double nr = (a.v[j] + tw.v[j]) * inv_tot.v[j];
      acc.d[j] += fabs(nr - last_rank);
      yn.v[j] = nr * mul;
      acc.c[j] += (__double_as_longlong(yn.v[j]) != __double_as_longlong(yl.v[j])) ? 1.0 : 0.0;
    } else {
      yn.v[j] = yl.v[j];
    }
    if (has_out) acc.s[j] += yn.v[j];
  }
  st_row_stream<VEC>(p.y_next + v * TP + VEC * l8, yn);
  // Fused exchange: the finished row goes straight into every peer's copy of y_next over
  // NVLink (posted stores), overlapping the transfer with the rest of the sweep instead of
  // an all-gather afterwards.
  for (int q = 0; q < p.n_peers; ++q) st_row_plain<VEC>(p.peer_next[q] + v * TP + VEC * l8, yn);
}

// CTA-wide fixed-shape reduction of the per-thread sums into red[blockIdx.x].
template <int LPR, int VEC>
__device__ __forceinline__ void block_reduce(const SweepParams& p, const Acc<VEC>& acc, bool owner) {
  constexpr int TP = LPR * VEC;
  constexpr int kWarps = kThreads / 32;
  __shared__ double sm[kWarps][3 * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane % LPR;
  double v[3 * VEC];
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    v[j] = owner ? acc.d[j] : 0.0;
    v[VEC + j] = owner ? acc.s[j] : 0.0;
    v[2 * VEC + j] = owner ? acc.c[j] : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 3 * VEC; ++i)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], o);
  if (lane < LPR) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) {
      sm[warp][VEC * l8 + j] = v[j];
      sm[warp][TP + VEC * l8 + j] = v[VEC + j];
      sm[warp][2 * TP + VEC * l8 + j] = v[2 * VEC + j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 * TP) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s += sm[w][threadIdx.x];
    p.red[(size_t)blockIdx.x * 3 * TP + threadIdx.x] = s;
  }
}

// One gather step of an LPR-lane group: NJ source ids come from the `idx`
// registers of lanes first_lane, first_lane + STRIDE, ...; all NJ row loads are
// issued before any add so that they are in flight together.  Invalid slots
// (kNone) load row 0 and add nothing.
template <int LPR, int VEC, int NJ, int STRIDE>
__device__ __forceinline__ void gather_step(const double* __restrict__ y, unsigned mask, uint32_t idx, int first_lane,
                                            int l8, Vec<VEC>& a) {
  constexpr int TP = LPR * VEC;
  Vec<VEC> rows[NJ];
  bool ok[NJ];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const uint32_t u = __shfl_sync(mask, idx, first_lane + j * STRIDE);
    ok[j] = u != kNone;
    rows[j] = ld_row_gather<VEC>(y + (uint64_t)(ok[j] ? u : 0u) * TP + VEC * l8);
  }
#pragma unroll
  for (int j = 0; j < NJ; ++j)
#pragma unroll
    for (int c = 0; c < VEC; ++c) a.v[c] += ok[j] ? rows[j].v[c] : 0.0;
}

// Rows with in-degree <= kShortMax: one row per LPR-lane group, 32/LPR rows
// per warp, consecutive rows in consecutive groups (coalesced y rows).
// Software pipeline over row blocks: the row pointers are fetched two blocks
// ahead and the first LPR source ids one block ahead, and the row's own y/scale
// before the gathers, so that only the row loads themselves are a dependent
// round trip per block.
template <int LPR, int VEC>
__global__ void __launch_bounds__(kThreads, 4) k_sweep_short(SweepParams p, uint32_t n_row_blocks) {
  constexpr int TP = LPR * VEC, GPW = 32 / LPR, GPC = GPW * (kThreads / 32);
  const int lane = threadIdx.x & 31, l8 = lane % LPR, g = lane / LPR;
  const unsigned gmask = (LPR == 32 ? 0xFFFFFFFFu : ((1u << LPR) - 1u)) << (g * LPR);
  const int group_in_cta = (threadIdx.x >> 5) * GPW + g;
  Acc<VEC> acc;
  auto load_ptr = [&](uint32_t rb, uint64_t& b, uint64_t& e) {
    const uint32_t r = rb * GPC + group_in_cta;
    b = e = 0;
    if (rb < n_row_blocks && r < p.rows_loc) {
      b = p.in_ptr[r];
      e = p.in_ptr[r + 1];
    }
  };
  auto first_ids = [&](uint64_t b, uint64_t e) -> uint32_t {
    return (e - b <= kShortMax && b + l8 < e) ? __ldg(p.in_src + b + l8) : kNone;
  };
  uint32_t rb = blockIdx.x;
  uint64_t b_c, e_c, b_n, e_n;
  load_ptr(rb, b_c, e_c);
  load_ptr(rb + gridDim.x, b_n, e_n);
  uint32_t idx_c = first_ids(b_c, e_c);
  for (; rb < n_row_blocks; rb += gridDim.x) {
    const uint32_t r = rb * GPC + group_in_cta;
    const uint64_t b = b_c, e = e_c;
    uint32_t idx = idx_c;
    b_c = b_n;
    e_c = e_n;
    load_ptr(rb + 2 * gridDim.x, b_n, e_n);
    idx_c = first_ids(b_c, e_c);
    if (r >= p.rows_loc) continue;
    if (e - b > kShortMax) continue;  // a long row: k_sweep_long / k_sweep_fix own it
    const Vec<VEC> yl = ld_row_stream<VEC>(p.y_last + (p.row_lo + r) * TP + VEC * l8);
    const double m = p.mul[r];
    Vec<VEC> a;
#pragma unroll
    for (int c = 0; c < VEC; ++c) a.v[c] = 0.0;
    for (uint64_t i = b; i < e; i += LPR) {
      if (i != b) idx = (i + l8 < e) ? __ldg(p.in_src + i + l8) : kNone;
      gather_step<LPR, VEC, LPR, 1>(p.y_last, gmask, idx, g * LPR, l8, a);
    }
    epilogue<LPR, VEC>(p, r, l8, a, yl, m, acc);
  }
  if (p.n_peers) __threadfence_system();  // pushed rows are performed before the kernel retires
  block_reduce<LPR, VEC>(p, acc, true);
}

// ---- short rows, lean variant ---------------------------------------------------
// Same arithmetic, bit for bit, as k_sweep_short (same gather order, same epilogue
// expressions), restructured after the round-1 ncu source profile: the first version
// spent 72 % of its instructions outside the gathers and spilled its prefetched row
// pointers, which serialised three memory round trips per row block.  Here
//   * row pointers are 32-bit (local edge count < 2^32) and the array is padded, so the
//     two-block-ahead pointer prefetch needs no bounds checks and two registers;
//   * the row scale and its reciprocal come from one packed 16-byte record (no fp64
//     divide per row), the per-topic 1/Tot is loaded once per kernel;
//   * "first sweep" and "some topic frozen" are template parameters, the changed-bits
//     test is an integer OR, shuffles run under the full mask with a warp-uniform trip
//     count (no MATCH/BRA.DIV), and absent edges are predicated off instead of loading
//     row 0.
template <int VEC>
struct Acc2 {
  double d[VEC], s[VEC];
  uint32_t chg;  // bit j: some row changed topic column j of this lane
  __device__ Acc2() : chg(0) {
#pragma unroll
    for (int j = 0; j < VEC; ++j) d[j] = s[j] = 0.0;
  }
};

template <int LPR, int VEC, bool FIRST, bool ALL>
__device__ __forceinline__ void epilogue2(const SweepParams& p, uint32_t r, int l8, Vec<VEC> a, const Vec<VEC>& yl,
                                          double2 m2, const Vec<VEC>& inv_tot, const Vec<VEC>& init,
                                          Acc2<VEC>& acc) {
  constexpr int TP = LPR * VEC;
  const uint64_t v = p.row_lo + r;
  const bool has_out = m2.y > 0.0;
  const double mul = m2.x, inv_mul = fabs(m2.y);
  Vec<VEC> tw;  // see epilogue()
#pragma unroll
  for (int j = 0; j < VEC; ++j) tw.v[j] = p.tele;
  if (p.tele_w) {
    const Vec<VEC> w = ld_row_stream<VEC>(p.tele_w + (uint64_t)r * TP + VEC * l8);
#pragma unroll
    for (int j = 0; j < VEC; ++j) tw.v[j] = __dmul_rn(p.tele, w.v[j]);
  }
  Vec<VEC> yn;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    double last_rank = yl.v[j] * inv_mul;
    if (FIRST) {
      a.v[j] += init.v[j];
      last_rank = init.v[j];
    }
    const double nr = (a.v[j] + tw.v[j]) * inv_tot.v[j];
    const double cand = nr * mul;
    const bool live = ALL || ((p.active_mask >> (VEC * l8 + j)) & 1u);
    if (live) {
      acc.d[j] += fabs(nr - last_rank);
      yn.v[j] = cand;
      acc.chg |= (__double_as_longlong(cand) != __double_as_longlong(yl.v[j])) ? (1u << j) : 0u;
    } else {
      yn.v[j] = yl.v[j];
    }
    acc.s[j] += has_out ? yn.v[j] : 0.0;
  }
  st_row_stream<VEC>(p.y_next + v * TP + VEC * l8, yn);
  for (int q = 0; q < p.n_peers; ++q) st_row_plain<VEC>(p.peer_next[q] + v * TP + VEC * l8, yn);
}

template <int LPR, int VEC>
__device__ __forceinline__ void block_reduce2(const SweepParams& p, const Acc2<VEC>& a2) {
  Acc<VEC> acc;
#pragma unroll
  for (int j = 0; j < VEC; ++j) {
    acc.d[j] = a2.d[j];
    acc.s[j] = a2.s[j];
    acc.c[j] = ((a2.chg >> j) & 1u) ? 1.0 : 0.0;
  }
  block_reduce<LPR, VEC>(p, acc, true);
}

template <int LPR, int VEC, bool FIRST, bool ALL>
__global__ void __launch_bounds__(kThreads, 4) k_sweep_short32(SweepParams p, uint32_t n_row_blocks) {
  constexpr int TP = LPR * VEC, GPW = 32 / LPR, GPC = GPW * (kThreads / 32);
  const int lane = threadIdx.x & 31, l8 = lane % LPR, gbase = lane - l8;
  const uint32_t group_in_cta = (threadIdx.x >> 5) * GPW + (lane / LPR);
  const uint32_t* __restrict__ ptr = p.in_ptr32;
  const uint32_t* __restrict__ src = p.in_src;
  const double* __restrict__ y = p.y_last;
  Acc2<VEC> acc;
  const Vec<VEC> inv_tot = ld_row_plain<VEC>(p.inv_tot + VEC * l8);
  Vec<VEC> init;
#pragma unroll
  for (int c = 0; c < VEC; ++c) init.v[c] = FIRST ? p.init[VEC * l8 + c] : 0.0;
  const uint32_t step = gridDim.x * GPC;
  uint32_t r = p.row_begin + blockIdx.x * GPC + group_in_cta;
  // the pointer array is padded past rows_loc by three grid strides (ss_graph_load_csr)
  uint32_t b_c = ptr[r], e_c = ptr[r + 1];
  uint32_t b_n = ptr[r + step], e_n = ptr[r + step + 1];
  uint32_t idx_c = (e_c - b_c <= kShortMax && b_c + l8 < e_c) ? __ldg(src + b_c + l8) : 0u;
  for (uint32_t rb = blockIdx.x; rb < n_row_blocks; rb += gridDim.x, r += step) {
    const uint32_t b = b_c, e = e_c;
    uint32_t idx = idx_c;
    b_c = b_n;
    e_c = e_n;
    b_n = ptr[r + 2 * step];
    e_n = ptr[r + 2 * step + 1];
    idx_c = (e_c - b_c <= kShortMax && b_c + l8 < e_c) ? __ldg(src + b_c + l8) : 0u;
    const bool mine = r < p.rows_loc && e - b <= kShortMax;  // long rows: k_sweep_long / k_sweep_fix
    const uint32_t deg = mine ? e - b : 0u;
    Vec<VEC> yl;
    double2 m2 = make_double2(1.0, -1.0);
    if (mine) {
      yl = ld_row_stream<VEC>(y + (p.row_lo + r) * TP + VEC * l8);
      m2 = __ldg(p.mul2 + r);
    }
    Vec<VEC> a;
#pragma unroll
    for (int c = 0; c < VEC; ++c) a.v[c] = 0.0;
    const uint32_t max_deg = __reduce_max_sync(0xFFFFFFFFu, deg);
    for (uint32_t i = 0; i < max_deg; i += LPR) {
      if (i) idx = (b + i + l8 < e) ? __ldg(src + b + i + l8) : 0u;
      Vec<VEC> rows[LPR];
#pragma unroll
      for (int j = 0; j < LPR; ++j) {
        const uint32_t u = __shfl_sync(0xFFFFFFFFu, idx, gbase + j);
        if (i + j < deg) {
          rows[j] = ld_row_gather<VEC>(y + (uint64_t)u * TP + VEC * l8);
        } else {
#pragma unroll
          for (int c = 0; c < VEC; ++c) rows[j].v[c] = 0.0;
        }
      }
#pragma unroll
      for (int j = 0; j < LPR; ++j)
#pragma unroll
        for (int c = 0; c < VEC; ++c) a.v[c] += rows[j].v[c];
    }
    if (mine) epilogue2<LPR, VEC, FIRST, ALL>(p, r, l8, a, yl, m2, inv_tot, init, acc);
  }
  if (p.n_peers) __threadfence_system();
  block_reduce2<LPR, VEC>(p, acc);
}

// ---- short rows, asynchronous gather ring -----------------------------------------
// The lean kernel above is bound by how many row loads fit in registers: 8 x 16 B per lane,
// half of them empty for the average short row, against a loaded memory latency of ~2.5 us
// (ncu: 79 % long-scoreboard stalls at 35 % DRAM utilisation).  Here the gathered rows land in
// shared memory instead (cp.async, 16 B per lane, LPR lanes per row): every warp owns a ring
// of kRingBytes / (TP * 8) row slots, streams the compacted short-row source list `ssrc` in
// 32-edge windows, and keeps up to four windows (128 rows at T = 16) in flight, all of them
// real edges.  The consumer half of the same warp walks its rows in order, one row per LPR-lane
// group, adds the landed rows in ascending source order (the same sums, bit for bit, as the
// register kernels) and runs the fused epilogue.  Each warp owns one contiguous row range
// (edge- and row-balanced, k_short_partition), so the pipeline never drains inside a sweep.
constexpr int kAsyncThreads = 128;
constexpr int kRingBytes = 16384;  // per warp

__device__ __forceinline__ void cp_async16(uint32_t smem_addr, const void* gptr) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

template <int LPR, int NT>
__device__ __forceinline__ void block_reduce_nt(const SweepParams& p, const Acc2<2>& acc) {
  constexpr int TP = LPR * 2, kWarps = NT / 32;
  __shared__ double sm[kWarps][3 * 16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane % LPR;
  double v[6];
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    v[j] = acc.d[j];
    v[2 + j] = acc.s[j];
    v[4 + j] = ((acc.chg >> j) & 1u) ? 1.0 : 0.0;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) v[i] += __shfl_xor_sync(0xFFFFFFFFu, v[i], o);
  if (lane < LPR) {
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      sm[warp][2 * l8 + j] = v[j];
      sm[warp][TP + 2 * l8 + j] = v[2 + j];
      sm[warp][2 * TP + 2 * l8 + j] = v[4 + j];
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 * TP) {
    double t = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) t += sm[w][threadIdx.x];
    p.red[(size_t)blockIdx.x * 3 * TP + threadIdx.x] = t;
  }
}

template <int LPR, bool FIRST, bool ALL>
__global__ void __launch_bounds__(kAsyncThreads, 3) k_sweep_short_async(SweepParams p) {
  constexpr int VEC = 2, TP = LPR * VEC, GPW = 32 / LPR, ROWB = TP * 8;
  constexpr uint32_t S = kRingBytes / ROWB;  // ring slots: >= GPW * kShortMax + 32, power of two
  static_assert(S >= GPW * kShortMax + 32 && (S & (S - 1)) == 0, "ring too small");
  constexpr uint32_t kMaxAhead = S / 32;     // windows that fit in the ring
  extern __shared__ __align__(128) unsigned char ring_raw[];
  const int lane = threadIdx.x & 31, warp_in_cta = threadIdx.x >> 5, l8 = lane % LPR, g = lane / LPR;
  const uint32_t gw = blockIdx.x * (kAsyncThreads / 32) + warp_in_cta;
  unsigned char* ring = ring_raw + (size_t)warp_in_cta * kRingBytes;
  const uint32_t ring_s = (uint32_t)__cvta_generic_to_shared(ring);
  const uint32_t* __restrict__ sptr = p.sptr;
  const uint32_t* __restrict__ ssrc = p.ssrc;
  const double* __restrict__ y = p.y_last;
  const uint32_t r_lo = p.warp_rows[gw], r_hi = p.warp_rows[gw + 1];
  const uint32_t e_lo = sptr[r_lo], e_hi = sptr[r_hi];
  Acc2<VEC> acc;
  const Vec<VEC> inv_tot = ld_row_plain<VEC>(p.inv_tot + VEC * l8);
  Vec<VEC> init;
#pragma unroll
  for (int c = 0; c < VEC; ++c) init.v[c] = FIRST ? p.init[VEC * l8 + c] : 0.0;

  // producer state: windows are absolute 32-edge blocks of ssrc; pw = next window to issue
  uint32_t pw = e_lo >> 5;
  const uint32_t w_end = (e_hi + 31) >> 5;  // one past the last window holding an edge of this warp
  auto load_idx = [&](uint32_t w) -> uint32_t {
    const uint32_t j = w * 32 + lane;
    return (w < w_end && j >= e_lo && j < e_hi) ? __ldg(ssrc + j) : kNone;
  };
  uint32_t idx_next = load_idx(pw);
  auto issue_window = [&]() {
    const uint32_t idx = idx_next;
    idx_next = load_idx(pw + 1);
    const uint32_t slot0 = (pw * 32) & (S - 1);
#pragma unroll
    for (int k = 0; k < LPR; ++k) {
      const int t = k * GPW + g;  // edge of the window handled by this lane group in step k
      const uint32_t u = __shfl_sync(0xFFFFFFFFu, idx, t);
      if (u != kNone) cp_async16(ring_s + (slot0 + t) * ROWB + l8 * 16, y + (uint64_t)u * TP + VEC * l8);
    }
    cp_async_commit();
    ++pw;
  };
  // fill the ring
  for (uint32_t k = 0; k < kMaxAhead && pw < w_end; ++k) issue_window();

  uint32_t r = r_lo + g;
  uint32_t b_n = 0, e_n = 0;
  Vec<VEC> yl_n;
  double2 m2_n = make_double2(0.0, -1.0);
#pragma unroll
  for (int c = 0; c < VEC; ++c) yl_n.v[c] = 0.0;
  auto prefetch_row = [&](uint32_t rr) {
    b_n = e_n = e_hi;
    m2_n = make_double2(0.0, -1.0);
    if (rr < r_hi) {
      b_n = sptr[rr];
      e_n = sptr[rr + 1];
      m2_n = __ldg(p.mul2 + rr);
      yl_n = ld_row_stream<VEC>(y + (p.row_lo + rr) * TP + VEC * l8);
    }
  };
  prefetch_row(r);
  for (uint32_t r0 = r_lo; r0 < r_hi; r0 += GPW, r += GPW) {
    const uint32_t b = b_n, e = e_n;
    const Vec<VEC> yl = yl_n;
    const double2 m2 = m2_n;
    prefetch_row(r + GPW);
    const bool mine = r < r_hi && m2.x != 0.0;  // m2.x == 0 marks a long row (k_sweep_long / k_sweep_fix)
    const uint32_t deg = e - b;                // 0 for long rows and rows past the range
    // every edge below e_need must have landed
    const uint32_t e_need = __reduce_max_sync(0xFFFFFFFFu, e);
    const uint32_t w_need = (e_need + 31) >> 5;
    while (pw < w_need) issue_window();  // cannot happen while the ring invariant holds; kept for safety
    const uint32_t pending = pw - w_need;
    if (pending == 0) cp_async_wait<0>();
    else if (pending == 1) cp_async_wait<1>();
    else if (pending == 2) cp_async_wait<2>();
    else if (pending == 3) cp_async_wait<3>();
    else if (pending <= 7) cp_async_wait<4>();
    else if (pending <= 15) cp_async_wait<8>();
    else cp_async_wait<16>();
    __syncwarp();
    Vec<VEC> a;
#pragma unroll
    for (int c = 0; c < VEC; ++c) a.v[c] = 0.0;
    const uint32_t max_deg = __reduce_max_sync(0xFFFFFFFFu, deg);
    for (uint32_t i = 0; i < max_deg; i += 4) {
      double2 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        v[j] = make_double2(0.0, 0.0);
        if (i + j < deg)
          v[j] = *reinterpret_cast<const double2*>(ring + (size_t)((b + i + j) & (S - 1)) * ROWB + l8 * 16);
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        a.v[0] += v[j].x;
        a.v[1] += v[j].y;
      }
    }
    if (mine) epilogue2<LPR, VEC, FIRST, ALL>(p, r, l8, a, yl, m2, inv_tot, init, acc);
    // refill: window w may be issued once every edge below (w + 1) * 32 - S has been consumed
    __syncwarp();
    const uint32_t consumed = min(__reduce_min_sync(0xFFFFFFFFu, b_n), e_hi);  // next block's first edge
    while (pw < w_end && (pw + 1) * 32 <= consumed + S) issue_window();
  }
  cp_async_wait<0>();
  if (p.n_peers) __threadfence_system();
  block_reduce_nt<LPR, kAsyncThreads>(p, acc);
}

// Long-row tasks: one warp per task, 32 edges per step, group g takes edges
// j*GPW+g; the GPW partial rows are combined with shuffles.  The next step's
// indices are fetched before this step's rows.
template <int LPR, int VEC>
__global__ void __launch_bounds__(kThreads, 4) k_sweep_long(SweepParams p, const LongTask* __restrict__ tasks,
                                                           uint32_t n_tasks, double* __restrict__ partials) {
  constexpr int TP = LPR * VEC, GPW = 32 / LPR;
  const int lane = threadIdx.x & 31, l8 = lane % LPR, g = lane / LPR;
  const uint32_t warp = (blockIdx.x * kThreads + threadIdx.x) >> 5;
  const uint32_t n_warps = (gridDim.x * kThreads) >> 5;
  Acc<VEC> acc;
  for (uint32_t ti = warp; ti < n_tasks; ti += n_warps) {
    const LongTask t = tasks[ti];
    const uint32_t* src = p.in_src + t.e_begin;
    Vec<VEC> yl;
#pragma unroll
    for (int c = 0; c < VEC; ++c) yl.v[c] = 0.0;
    double m = 0;
    if (t.slot < 0 && g == 0) {
      yl = ld_row_stream<VEC>(p.y_last + (p.row_lo + t.row) * TP + VEC * l8);
      m = p.mul[t.row];
    }
    Vec<VEC> a;
#pragma unroll
    for (int c = 0; c < VEC; ++c) a.v[c] = 0.0;
    uint32_t idx_next = lane < t.n ? __ldg(src + lane) : kNone;
    for (uint32_t i = 0; i < t.n; i += 32) {
      const uint32_t idx = idx_next;
      idx_next = (i + 32 + lane < t.n) ? __ldg(src + i + 32 + lane) : kNone;
      gather_step<LPR, VEC, LPR, GPW>(p.y_last, 0xFFFFFFFFu, idx, g, l8, a);
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
      for (int c = 0; c < VEC; ++c) a.v[c] += __shfl_xor_sync(0xFFFFFFFFu, a.v[c], o);
    if (g == 0) {
      if (t.slot < 0) {
        epilogue<LPR, VEC>(p, t.row, l8, a, yl, m, acc);
      } else {
        st_row_plain<VEC>(partials + (size_t)t.slot * TP + VEC * l8, a);
      }
    }
  }
  if (p.n_peers) __threadfence_system();
  block_reduce<LPR, VEC>(p, acc, g == 0);
}

// Rows that span several tasks: one CTA per row sums the partial rows in slot
// order (fixed tree) and runs the epilogue.
template <int LPR, int VEC>
__global__ void __launch_bounds__(kThreads) k_sweep_fix(SweepParams p, const FixRow* __restrict__ rows,
                                                       uint32_t n_fix, const double* __restrict__ partials) {
  constexpr int TP = LPR * VEC, GPW = 32 / LPR, GPC = GPW * (kThreads / 32);
  __shared__ double sm[kThreads / 32][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane % LPR, g = lane / LPR;
  const int group_in_cta = warp * GPW + g;
  Acc<VEC> acc;
  for (uint32_t fi = blockIdx.x; fi < n_fix; fi += gridDim.x) {
    const FixRow fr = rows[fi];
    Vec<VEC> yl;
#pragma unroll
    for (int c = 0; c < VEC; ++c) yl.v[c] = 0.0;
    double m = 0;
    if (warp == 0 && g == 0) {
      yl = ld_row_stream<VEC>(p.y_last + (p.row_lo + fr.row) * TP + VEC * l8);
      m = p.mul[fr.row];
    }
    Vec<VEC> a;
#pragma unroll
    for (int c = 0; c < VEC; ++c) a.v[c] = 0.0;
    for (uint32_t s = group_in_cta; s < fr.n_slots; s += GPC) {
      const Vec<VEC> v = ld_row_plain<VEC>(partials + (size_t)(fr.first_slot + s) * TP + VEC * l8);
#pragma unroll
      for (int c = 0; c < VEC; ++c) a.v[c] += v.v[c];
    }
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
      for (int c = 0; c < VEC; ++c) a.v[c] += __shfl_xor_sync(0xFFFFFFFFu, a.v[c], o);
    __syncthreads();  // sm reuse across iterations
    if (g == 0) {
#pragma unroll
      for (int c = 0; c < VEC; ++c) sm[warp][VEC * l8 + c] = a.v[c];
    }
    __syncthreads();
    if (warp == 0 && g == 0) {
      Vec<VEC> b;
#pragma unroll
      for (int c = 0; c < VEC; ++c) {
        double t = 0;
#pragma unroll
        for (int w = 0; w < kThreads / 32; ++w) t += sm[w][VEC * l8 + c];
        b.v[c] = t;
      }
      epilogue<LPR, VEC>(p, fr.row, l8, b, yl, m, acc);
    }
  }
  if (p.n_peers) __threadfence_system();
  block_reduce<LPR, VEC>(p, acc, warp == 0 && g == 0);
}

// Two-stage fixed-order sum of the per-CTA partial rows: kReduceCtas CTAs each
// fold a strided subset of the slots, then k_fold_stage folds those.
constexpr int kReduceCtas = 32;
__global__ void __launch_bounds__(kThreads) k_reduce_partials(const double* __restrict__ red, uint32_t n_slots,
                                                              int width, double* __restrict__ stage) {
  __shared__ double sm[kThreads];
  const int stripes = kThreads / width;
  const int c = threadIdx.x % width, s = threadIdx.x / width;
  double v = 0;
  if (s < stripes) {
    const uint32_t step = gridDim.x * stripes;
    uint32_t i = blockIdx.x * stripes + s;
    for (; i + 3 * step < n_slots; i += 4 * step) {
      const double x0 = red[(size_t)i * width + c], x1 = red[(size_t)(i + step) * width + c];
      const double x2 = red[(size_t)(i + 2 * step) * width + c], x3 = red[(size_t)(i + 3 * step) * width + c];
      v += x0;
      v += x1;
      v += x2;
      v += x3;
    }
    for (; i < n_slots; i += step) v += red[(size_t)i * width + c];
  }
  sm[threadIdx.x] = (s < stripes) ? v : 0.0;
  __syncthreads();
  if (threadIdx.x < width) {
    double t = 0;
    for (int k = 0; k < stripes; ++k) t += sm[k * width + threadIdx.x];
    stage[(size_t)blockIdx.x * width + threadIdx.x] = t;
  }
}

// sums = fold of the staged rows; tot[t] = S_t + (1-d) N (pagerank.go:111-112).
// With several ranks the fold and tot are split around the cross-rank sum.
__global__ void k_fold_stage(const double* __restrict__ stage, int n_rows, int width, double* __restrict__ sums) {
  const int c = threadIdx.x;
  if (c >= width) return;
  double t = 0;
  for (int k = 0; k < n_rows; ++k) t += stage[(size_t)k * width + c];
  sums[c] = t;
}
__global__ void k_finish_tot(const double* __restrict__ sums, int TP, double tele, double n_nodes,
                             double* __restrict__ tot) {
  const int t = threadIdx.x;
  if (t < TP) tot[t] = 1.0 / (sums[TP + t] + tele * n_nodes);  // the sweep multiplies by 1/Tot
}

// y0[v][t] = m(v) / n_t for every node (pagerank.go:101-107) and the partial
// S_0 of this rank's rows.
template <int LPR, int VEC>
__global__ void __launch_bounds__(kThreads) k_init(double* __restrict__ y, const uint32_t* __restrict__ outdeg,
                                                  uint64_t n_nodes, double damping, const double* __restrict__ init,
                                                  uint64_t row_lo, uint32_t rows_loc, double* __restrict__ mul_loc,
                                                  double2* __restrict__ mul2_loc,
                                                  const uint64_t* __restrict__ in_ptr_loc, double* __restrict__ red) {
  constexpr int TP = LPR * VEC;
  __shared__ double sm[kThreads / 32][16];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, l8 = lane % LPR;
  const uint64_t n_groups = (uint64_t)gridDim.x * kThreads / LPR;
  double s[VEC];
#pragma unroll
  for (int c = 0; c < VEC; ++c) s[c] = 0.0;
  for (uint64_t v = ((uint64_t)blockIdx.x * kThreads + threadIdx.x) / LPR; v < n_nodes; v += n_groups) {
    const uint32_t od = outdeg[v];
    const double m = od ? damping / (double)od : 1.0;
    Vec<VEC> val;
#pragma unroll
    for (int c = 0; c < VEC; ++c) val.v[c] = m * init[VEC * l8 + c];
    st_row_plain<VEC>(y + v * TP + VEC * l8, val);
    if (v >= row_lo && v < row_lo + rows_loc) {
      if (l8 == 0) {
        mul_loc[v - row_lo] = od ? m : 0.0;
        const bool has_out = od && m > 0.0;  // the sweeps' test (mul > 0)
        double2 rec = has_out ? make_double2(m, 1.0 / m) : make_double2(1.0, -1.0);
        if (in_ptr_loc[v - row_lo + 1] - in_ptr_loc[v - row_lo] > kShortMax) rec.x = 0.0;  // long row marker
        mul2_loc[v - row_lo] = rec;
      }
      if (od) {
#pragma unroll
        for (int c = 0; c < VEC; ++c) s[c] += val.v[c];
      }
    }
  }
#pragma unroll
  for (int o = LPR; o < 32; o <<= 1)
#pragma unroll
    for (int c = 0; c < VEC; ++c) s[c] += __shfl_xor_sync(0xFFFFFFFFu, s[c], o);
  if (lane < LPR) {
#pragma unroll
    for (int c = 0; c < VEC; ++c) sm[warp][VEC * l8 + c] = s[c];
  }
  __syncthreads();
  if (threadIdx.x < 3 * TP) {
    double t = 0;
    if (threadIdx.x >= TP && threadIdx.x < 2 * TP)
      for (int w = 0; w < kThreads / 32; ++w) t += sm[w][threadIdx.x - TP];
    red[(size_t)blockIdx.x * 3 * TP + threadIdx.x] = t;
  }
}

// teleport weights [rows][T] -> [rows][TP] (padded columns get 1: they are frozen anyway)
__global__ void k_pad_rows(const double* __restrict__ in, uint64_t rows, int T, int TP, double* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * (uint64_t)TP) return;
  const uint64_t r = i / TP;
  const int t = (int)(i % TP);
  out[i] = t < T ? in[r * T + t] : 1.0;
}

// rank[v][t] = y[v][t] / m(v) for rows [lo, hi) into a dense [rows][T] buffer.
__global__ void k_unscale(const double* __restrict__ y, const uint32_t* __restrict__ outdeg, double damping,
                          int TP, int T, uint64_t lo, uint64_t hi, double* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint64_t total = (hi - lo) * (uint64_t)T;
  if (i >= total) return;
  const uint64_t v = lo + i / T;
  const int t = (int)(i % T);
  const uint32_t od = outdeg[v];
  const double m = od ? damping / (double)od : 1.0;
  out[i] = y[v * TP + t] / m;
}

// ---- load-time kernels ------------------------------------------------------
__global__ void k_outdeg(const uint64_t* __restrict__ row_ptr, uint64_t n, uint32_t* __restrict__ outdeg) {
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n) outdeg[u] = (uint32_t)(row_ptr[u + 1] - row_ptr[u]);
}
// src[e] = the row that owns out-edge e, four consecutive edges per thread (one
// binary search, then a walk over the row boundaries); children validated.
__global__ void k_expand_src4(const uint64_t* __restrict__ row_ptr, uint64_t n, uint64_t n_edges,
                              const uint32_t* __restrict__ col_idx, uint32_t* __restrict__ src,
                              int* __restrict__ bad, uint64_t n_child) {
  const uint64_t e0 = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (e0 >= n_edges) return;
  uint64_t lo = 0, hi = n;  // last u with row_ptr[u] <= e0
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (row_ptr[mid] <= e0) lo = mid; else hi = mid;
  }
  uint64_t u = lo, next = row_ptr[u + 1];
  const uint64_t e1 = min(n_edges, e0 + 4);
  for (uint64_t e = e0; e < e1; ++e) {
    while (e >= next) next = row_ptr[++u + 1];
    src[e] = (uint32_t)u;
    if (col_idx[e] >= n_child) *bad = 1;
  }
}
// sharded load: cnt[v] = in-edges of v inside this rank's slice, from the slice's child-sorted list
__global__ void k_count_from_ptr(const unsigned long long* __restrict__ ptr, uint64_t n, uint32_t* __restrict__ cnt) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v < n) cnt[v] = (uint32_t)(ptr[v + 1] - ptr[v]);
}
__global__ void k_widen_counts(const uint32_t* __restrict__ cnt, uint64_t n, unsigned long long* __restrict__ out) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v <= n) out[v] = v < n ? cnt[v] : 0ull;
}
__global__ void k_add_u32(uint32_t* __restrict__ a, uint64_t n, uint32_t add) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] += add;
}
// in_ptr[v] = first position of child v in the child-sorted edge list
__global__ void k_in_ptr_from_sorted(const uint32_t* __restrict__ dst_sorted, uint64_t n_edges, uint64_t n,
                                     unsigned long long* __restrict__ in_ptr) {
  const uint64_t v = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v > n) return;
  uint64_t lo = 0, hi = n_edges;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (dst_sorted[mid] < v) lo = mid + 1; else hi = mid;
  }
  in_ptr[v] = lo;
}
// sharded load: in_ptr_loc[i] = first position of child row_lo + i in this rank's child-sorted in-edge list
__global__ void k_in_ptr_local(const uint32_t* __restrict__ dst_sorted, uint64_t n_edges, uint64_t row_lo,
                               uint32_t rows_loc, uint64_t* __restrict__ in_ptr_loc) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i > rows_loc) return;
  const uint64_t v = row_lo + i;
  uint64_t lo = 0, hi = n_edges;
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (dst_sorted[mid] < v) lo = mid + 1; else hi = mid;
  }
  in_ptr_loc[i] = lo;
}
__global__ void k_check_row_ptr(const uint64_t* __restrict__ row_ptr, uint64_t n, uint64_t n_edges,
                                int* __restrict__ bad) {
  const uint64_t u = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (u < n && row_ptr[u] > row_ptr[u + 1]) *bad = 1;
  if (u == 0 && (row_ptr[0] != 0 || row_ptr[n] != n_edges)) *bad = 1;
}
// Edge- and row-balanced 1-D partition: cost(v) = in_ptr[v] + 2 v is monotone;
// boundary r = first v with cost(v) >= r * cost(N) / world.
__global__ void k_partition(const unsigned long long* __restrict__ in_ptr, uint64_t n, int world,
                            unsigned long long* __restrict__ bounds) {
  const int r = threadIdx.x;
  if (r > world) return;
  if (r == world) {
    bounds[r] = n;
    return;
  }
  const unsigned long long total = in_ptr[n] + 2ull * n;
  const unsigned long long want = (unsigned long long)(((unsigned __int128)total * (unsigned)r) / (unsigned)world);
  uint64_t lo = 0, hi = n;  // first v with cost(v) >= want
  while (lo < hi) {
    const uint64_t mid = (lo + hi) >> 1;
    if (in_ptr[mid] + 2ull * mid >= want) hi = mid; else lo = mid + 1;
  }
  bounds[r] = lo;
}
__global__ void k_local_ptr(const unsigned long long* __restrict__ in_ptr_full, uint64_t row_lo, uint32_t rows_loc,
                            uint64_t* __restrict__ in_ptr_loc) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r <= rows_loc) in_ptr_loc[r] = in_ptr_full[row_lo + r] - in_ptr_full[row_lo];
}
// 32-bit copy of the local row pointers, padded with E_loc (rows of degree 0) so that the lean
// short-row kernel can prefetch pointers two grid strides ahead without bounds checks
__global__ void k_local_ptr32(const uint64_t* __restrict__ in_ptr_loc, uint32_t rows_loc, uint64_t n_padded,
                              uint32_t* __restrict__ out) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n_padded) out[r] = (uint32_t)in_ptr_loc[r <= rows_loc ? r : rows_loc];
}
// Short-row structures of the async kernel: sdeg[r] = in-degree if the row is short, else 0;
// its exclusive scan is sptr; ssrc = the short rows' sources, compacted in row order.
__global__ void k_short_deg(const uint64_t* __restrict__ in_ptr, uint32_t rows_loc, uint32_t* __restrict__ sdeg) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r > rows_loc) return;
  uint32_t d = 0;
  if (r < rows_loc) {
    const uint64_t deg = in_ptr[r + 1] - in_ptr[r];
    d = deg <= kShortMax ? (uint32_t)deg : 0u;
  }
  sdeg[r] = d;
}
__global__ void k_short_compact(const uint64_t* __restrict__ in_ptr, const uint32_t* __restrict__ in_src,
                                const uint32_t* sptr, uint32_t rows_loc, uint32_t pad, uint32_t* sptr_pad,
                                uint32_t* __restrict__ ssrc) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < rows_loc) {
    const uint32_t b = sptr[r], n = sptr[r + 1] - b;
    const uint64_t from = in_ptr[r];
    for (uint32_t i = 0; i < n; ++i) ssrc[b + i] = in_src[from + i];
  } else if (r < (uint64_t)rows_loc + pad) {
    sptr_pad[r + 1] = sptr[rows_loc];  // entries past rows_loc repeat the total
  }
}
// Row range of every warp of the async kernel: cost(r) = sptr[r] + 3 r (edges + rows) is monotone;
// boundary i = first r with cost(r) >= i * cost(R) / n_warps.
__global__ void k_short_partition(const uint32_t* __restrict__ sptr, uint32_t rows_loc, uint32_t n_warps,
                                  uint32_t* __restrict__ warp_rows) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i > n_warps) return;
  if (i == n_warps) {
    warp_rows[i] = rows_loc;
    return;
  }
  const uint64_t total = (uint64_t)sptr[rows_loc] + 3ull * rows_loc;
  const uint64_t want = total * i / n_warps;
  uint32_t lo = 0, hi = rows_loc;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if ((uint64_t)sptr[mid] + 3ull * mid >= want) hi = mid; else lo = mid + 1;
  }
  warp_rows[i] = lo;
}
__global__ void k_count_tasks(const uint64_t* __restrict__ in_ptr, uint32_t rows_loc, uint32_t* __restrict__ nt,
                              uint32_t* __restrict__ nf) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows_loc) return;
  const uint64_t deg = in_ptr[r + 1] - in_ptr[r];
  const uint32_t t = deg > kShortMax ? (uint32_t)((deg + kChunk - 1) / kChunk) : 0u;
  nt[r] = t;
  nf[r] = t > 1 ? 1u : 0u;
}
__global__ void k_fill_tasks(const uint64_t* __restrict__ in_ptr, const uint32_t* __restrict__ in_src,
                             uint32_t rows_loc, const uint32_t* __restrict__ nt, const uint32_t* __restrict__ toff,
                             const uint32_t* __restrict__ foff, LongTask* __restrict__ tasks,
                             unsigned long long* __restrict__ keys, uint32_t* __restrict__ order,
                             FixRow* __restrict__ fix, const unsigned long long* __restrict__ chunk_rows, int n_chunks) {
  const uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows_loc) return;
  const uint32_t t = nt[r];
  if (t == 0) return;
  unsigned long long chunk = 0;  // row chunk of the overlapped exchange (0 when there is one chunk)
  while ((int)chunk + 1 < n_chunks && r >= chunk_rows[chunk + 1]) ++chunk;
  const uint64_t b = in_ptr[r], deg = in_ptr[r + 1] - b;
  const uint32_t base = toff[r];
  for (uint32_t c = 0; c < t; ++c) {
    LongTask k;
    k.e_begin = b + (uint64_t)c * kChunk;
    k.row = (uint32_t)r;
    k.n = (uint32_t)min((uint64_t)kChunk, deg - (uint64_t)c * kChunk);
    k.slot = t > 1 ? (int32_t)(base + c) : -1;
    k.pad = 0;
    tasks[base + c] = k;
    keys[base + c] = (chunk << 32) | in_src[k.e_begin];  // grouped by chunk, then by first source
    order[base + c] = base + c;
  }
  if (t > 1) {
    FixRow f;
    f.row = (uint32_t)r;
    f.first_slot = base;
    f.n_slots = t;
    f.pad = 0;
    fix[foff[r]] = f;
  }
}
__global__ void k_gather_tasks(const LongTask* __restrict__ in, const uint32_t* __restrict__ order, uint32_t n,
                               LongTask* __restrict__ out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[order[i]];
}

}  // namespace

struct PagerankState {
  bool loaded = false;
  uint64_t N = 0, E = 0;
  uint64_t row_lo = 0;
  uint32_t rows_loc = 0;
  uint64_t E_loc = 0;
  std::vector<uint64_t> bounds;  // [world + 1]
  ss::DevBuf<uint32_t> outdeg;   // [N]
  ss::DevBuf<uint64_t> in_ptr;   // [rows_loc + 1]
  ss::DevBuf<uint32_t> in_ptr32; // [rows_loc + 1 + ptr32_pad], only when E_loc < 2^32
  uint64_t ptr32_pad = 0;        // 0: no 32-bit copy
  ss::DevBuf<double2> mul2;      // [rows_loc] {m, 1/m} or {1, -1}; m = 0 marks a long row
  ss::DevBuf<uint32_t> sptr, ssrc, warp_rows;  // async short-row kernel: short-only CSR and warp row ranges
  bool have_short = false;
  uint32_t E_short = 0;
  ss::DevBuf<uint32_t> in_src;   // [E_loc]
  ss::DevBuf<LongTask> tasks;
  ss::DevBuf<FixRow> fix;
  uint32_t n_tasks = 0, n_fix = 0;
  // Row chunks of the overlapped exchange (> 2 ranks over NCCL): the rank's rows are cut into n_chunks
  // edge-balanced ranges; a sweep finishes chunk c (short rows, long-row tasks, fix rows) and broadcasts
  // it on the exchange stream while chunk c + 1 computes.  One chunk otherwise.
  int n_chunks = 1;
  std::vector<uint32_t> chunk_rows, chunk_task, chunk_fix;  // [n_chunks + 1] local row / task / fix-row boundaries
  std::vector<uint64_t> all_chunk_rows;                     // [world][n_chunks + 1] global row boundaries of every rank
  cudaStream_t xstream = nullptr;                           // exchange stream (NCCL calls of the sweep loop)
  cudaStream_t pstream[kMaxPeers] = {};                     // one copy stream per peer: the peer copies of a chunk run
  cudaEvent_t pdone[kMaxPeers] = {};                        // on different copy engines at the same time
  cudaEvent_t chunk_ev[8] = {}, x_done = nullptr, red_done = nullptr;
  cudaEvent_t x_t0 = nullptr, x_t1 = nullptr;               // timing: exchange stream busy interval of a sweep
  // per-run state
  int TP = 0, T = 0;
  double damping = 0;
  ss::DevBuf<double> y[2];
  int cur = 0;  // y[cur] holds the latest ranks
  ss::DevBuf<double> mul, partials, red, stage, sums, tot, init;
  bool have_result = false;
  // topic-biased teleport (ss_pagerank_set_teleport): weights of this rank's rows, [rows_loc][tele_T] as given
  ss::DevBuf<double> tele_raw, tele_w;  // raw [rows_loc][tele_T]; padded [rows_loc][TP] built per run
  uint32_t tele_T = 0;                  // 0: uniform teleport (the reference)
  ss_pagerank_stats stats{};
  cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};  // [4]: end of the short-row kernel
  // exchange over peer memory (CUDA IPC): peers' y[0]/y[1] mapped into this process.
  //   fused: the peers are mapped (every rank agreed);  push_from_epilogue: 2 ranks, the sweep epilogue stores
  //   each finished row into the peer's state;  otherwise (3+ ranks) finished row chunks are pushed to every
  //   peer with copy-engine peer copies on the exchange stream, overlapped with the next chunk's sweep.
  bool fused = false;
  bool push_from_epilogue = false;
  int peer_rank[kMaxPeers] = {};
  int n_peers = 0;
  double* peer_y[2][kMaxPeers] = {};
  void* y_base_exported[2] = {nullptr, nullptr};
  // load-time scratch, kept between loads (grow-only)
  struct Scratch {
    ss::DevBuf<uint64_t> row_ptr;
    ss::DevBuf<uint32_t> col, src, col_sorted, src_sorted, nt, nf, toff, foff, order, order_out, sdeg;
    ss::DevBuf<unsigned long long> in_ptr_full, bounds, keys, keys_out, chunk_rows, slice_ptr;
    ss::DevBuf<uint32_t> cnt, recv_col, recv_src;
    ss::DevBuf<int> bad;
    ss::DevBuf<char> tmp;
    ss::DevBuf<LongTask> unsorted;
  } sc;
};

static void close_peers(PagerankState* s) {
  for (int b = 0; b < 2; ++b)
    for (int q = 0; q < kMaxPeers; ++q)
      if (s->peer_y[b][q]) {
        cudaIpcCloseMemHandle(s->peer_y[b][q]);
        s->peer_y[b][q] = nullptr;
      }
  s->fused = false;
  s->n_peers = 0;
}

// Map every peer's y buffers into this process (CUDA IPC) so that the sweep epilogue can
// push finished rows directly.  Handles travel through the engine's own NCCL communicator.
static int open_peers(ss_engine* e, PagerankState* s) {
  const int world = comm_world(e), rank = comm_rank(e);
  if (s->fused && s->y_base_exported[0] == s->y[0].p && s->y_base_exported[1] == s->y[1].p) return SS_OK;
  close_peers(s);
  if (world == 1 || world > kMaxPeers + 1) return SS_OK;
  // Measured on 8 B200s (profiles/r01_multigpu.txt): with one peer the pushed rows ride along
  // for free (2 GPUs: 658 -> 870 GTEPS), but seven 128-byte remote stores per row throttle the
  // whole sweep (70 ms vs 5 ms + 14.5 ms NCCL).  Default: fused for 2 ranks, NCCL beyond;
  // SS_PR_EXCHANGE=fused|nccl overrides.
  // Round 2: with 3+ ranks the rows no longer leave from the epilogue; finished chunks are pushed by the copy
  // engines instead (see PagerankState::fused), which costs the sweep no issue slots at all.
  // SS_PR_EXCHANGE=nccl keeps everything on NCCL broadcasts; =fused forces the epilogue push, =copy the
  // copy-engine push, at any rank count.
  const char* env = getenv("SS_PR_EXCHANGE");
  if (env && !strcmp(env, "nccl")) return SS_OK;
  s->push_from_epilogue = (world == 2 && !(env && !strcmp(env, "copy"))) || (env && !strcmp(env, "fused"));
  struct Pack {
    cudaIpcMemHandle_t h[2];
    int device;
  } mine{}, all[kMaxPeers + 1];
  SS_CUDA(cudaIpcGetMemHandle(&mine.h[0], s->y[0].p));
  SS_CUDA(cudaIpcGetMemHandle(&mine.h[1], s->y[1].p));
  mine.device = e->device;
  SS_TRY(comm_allgather_host_bytes(e, &mine, sizeof(Pack), all));
  int q = 0;
  for (int r = 0; r < world; ++r) {
    if (r == rank) continue;
    for (int b = 0; b < 2; ++b) {
      void* ptr = nullptr;
      cudaError_t err = cudaIpcOpenMemHandle(&ptr, all[r].h[b], cudaIpcMemLazyEnablePeerAccess);
      if (err != cudaSuccess) {  // no peer path: fall back to the NCCL exchange
        cudaGetLastError();
        close_peers(s);
        return SS_OK;
      }
      s->peer_y[b][q] = (double*)ptr;
    }
    s->peer_rank[q] = r;
    ++q;
  }
  s->n_peers = q;
  s->fused = true;
  s->y_base_exported[0] = s->y[0].p;
  s->y_base_exported[1] = s->y[1].p;
  return SS_OK;
}

// Every rank must take the same exchange path: fused only if all ranks mapped all peers.
static int agree_on_fused(ss_engine* e, PagerankState* s) {
  const int world = comm_world(e);
  if (world == 1) return SS_OK;
  SS_TRY(s->sums.reserve(1));
  const double mine = s->fused ? 1.0 : 0.0;
  double total = 0;
  SS_CUDA(cudaMemcpyAsync(s->sums.p, &mine, 8, cudaMemcpyHostToDevice, e->stream));
  SS_TRY(comm_allreduce_sum_f64(e, s->sums.p, 1));
  SS_CUDA(cudaMemcpyAsync(&total, s->sums.p, 8, cudaMemcpyDeviceToHost, e->stream));
  SS_CUDA(cudaStreamSynchronize(e->stream));
  if (total != (double)world && s->fused) {
    close_peers(s);
    s->y_base_exported[0] = s->y_base_exported[1] = nullptr;
  }
  return SS_OK;
}

void pagerank_state_free(PagerankState* s) {
  if (!s) return;
  close_peers(s);
  for (auto& e : s->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : s->chunk_ev)
    if (e) cudaEventDestroy(e);
  if (s->x_done) cudaEventDestroy(s->x_done);
  if (s->red_done) cudaEventDestroy(s->red_done);
  if (s->x_t0) cudaEventDestroy(s->x_t0);
  if (s->x_t1) cudaEventDestroy(s->x_t1);
  if (s->xstream) cudaStreamDestroy(s->xstream);
  for (auto& st : s->pstream)
    if (st) cudaStreamDestroy(st);
  for (auto& ev : s->pdone)
    if (ev) cudaEventDestroy(ev);
  delete s;
}

// Lane shape per padded topic count: (lanes per row, doubles per lane).
// 128-bit accesses (vec 2) are the default: the sweep sits at the chip's row-gather
// ceiling either way (scripts/bench_gather.cu) and the 256-bit shape spills its
// twelve reduction accumulators.  SS_PR_VEC=4 selects the 256-bit shape for TP >= 8.
struct Shape {
  int lpr, vec;
};
static Shape shape_for_topics(uint32_t T) {
  const char* env = getenv("SS_PR_VEC");
  const bool wide = env && atoi(env) == 4;
  if (T <= 2) return {1, 2};
  if (T <= 4) return {2, 2};
  if (T <= 8) return wide ? Shape{2, 4} : Shape{4, 2};
  return wide ? Shape{4, 4} : Shape{8, 2};
}

template <class F>
static int dispatch_shape(Shape sh, F&& f) {
  using std::integral_constant;
  if (sh.vec == 2) {
    switch (sh.lpr) {
      case 1: return f(integral_constant<int, 1>(), integral_constant<int, 2>());
      case 2: return f(integral_constant<int, 2>(), integral_constant<int, 2>());
      case 4: return f(integral_constant<int, 4>(), integral_constant<int, 2>());
      default: return f(integral_constant<int, 8>(), integral_constant<int, 2>());
    }
  }
  if (sh.lpr == 2) return f(integral_constant<int, 2>(), integral_constant<int, 4>());
  return f(integral_constant<int, 4>(), integral_constant<int, 4>());
}

static int create_sync_objects(PagerankState* s) {
  for (auto& ev : s->ev) SS_CUDA(cudaEventCreate(&ev));
  for (auto& ev : s->chunk_ev) SS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  SS_CUDA(cudaEventCreateWithFlags(&s->x_done, cudaEventDisableTiming));
  SS_CUDA(cudaEventCreateWithFlags(&s->red_done, cudaEventDisableTiming));
  SS_CUDA(cudaEventCreate(&s->x_t0));
  SS_CUDA(cudaEventCreate(&s->x_t1));
  SS_CUDA(cudaStreamCreateWithFlags(&s->xstream, cudaStreamNonBlocking));
  for (auto& st : s->pstream) SS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  for (auto& ev : s->pdone) SS_CUDA(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  return SS_OK;
}

// Second half of a graph load, shared by ss_graph_load_csr and ss_graph_load_csr_rows: everything that only
// depends on this rank's in-edge lists (s->in_ptr, s->in_src), the global out-degrees and the partition.
static int finish_load(ss_engine* e, PagerankState* s, std::chrono::steady_clock::time_point t_begin) {
  PagerankState::Scratch& sc = s->sc;
  cudaStream_t st = e->stream;
  const uint64_t N = s->N, E = s->E;
  const int world = comm_world(e);
  s->ptr32_pad = 0;
  if (s->E_loc < 0xFFFFFFFFull) {
    // padding: GPC <= 256 rows per CTA and step, persistent grid <= 8 CTAs per SM, two strides ahead
    const uint64_t pad = 2ull * (uint64_t)e->sm_count * 8 * 256 + 1024;
    const uint64_t n_padded = (uint64_t)s->rows_loc + 1 + pad;
    SS_TRY(s->in_ptr32.reserve(n_padded));
    k_local_ptr32<<<ss::div_up(n_padded, 256), 256, 0, st>>>(s->in_ptr.p, s->rows_loc, n_padded, s->in_ptr32.p);
    s->ptr32_pad = pad;
  }

  // short-only CSR of the cp.async ring kernel: measured equal to the lean register kernel at
  // configs[1] (1.39 vs 1.38 ms: both sit at the chip's random 128-byte row rate), so it is opt-in
  // and its structures are only built on request
  s->have_short = false;
  const char* short_env = getenv("SS_PR_SHORT");
  if (short_env && !strcmp(short_env, "async") && s->E_loc < 0xFFFFFFFFull && s->rows_loc) {
    const uint32_t R = s->rows_loc, pad = 64;
    SS_TRY(sc.sdeg.reserve((size_t)R + 1));
    SS_TRY(s->sptr.reserve((size_t)R + 2 + pad));
    k_short_deg<<<ss::div_up((uint64_t)R + 1, 256), 256, 0, st>>>(s->in_ptr.p, R, sc.sdeg.p);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sc.sdeg.p, s->sptr.p, (int)(R + 1), st);
    SS_TRY(sc.tmp.reserve(tmp_bytes));
    tmp_bytes = sc.tmp.n;
    SS_CUDA(cub::DeviceScan::ExclusiveSum(sc.tmp.p, tmp_bytes, sc.sdeg.p, s->sptr.p, (int)(R + 1), st));
    uint32_t e_short = 0;
    SS_CUDA(cudaMemcpyAsync(&e_short, s->sptr.p + R, 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    s->E_short = e_short;
    SS_TRY(s->ssrc.reserve((size_t)e_short + 64));
    k_short_compact<<<ss::div_up((uint64_t)R + pad, 256), 256, 0, st>>>(s->in_ptr.p, s->in_src.p, s->sptr.p, R, pad,
                                                                       s->sptr.p, s->ssrc.p);
    s->have_short = true;
  }

  // row chunks of the overlapped exchange
  {
    int C = world > 2 ? 4 : 1;
    if (const char* env = getenv("SS_PR_CHUNKS")) C = std::max(1, std::min(8, atoi(env)));
    if (world == 1 || s->rows_loc < 4096u * C || s->ptr32_pad == 0) C = 1;
    if (world > 1) {  // every rank must cut its block into the same number of chunks
      int32_t mine = C, all[64];
      SS_TRY(comm_allgather_host_bytes(e, &mine, sizeof(mine), all));
      for (int r = 0; r < world; ++r) C = std::min<int>(C, all[r]);
    }
    s->n_chunks = C;
    SS_TRY(sc.chunk_rows.reserve(C + 1));
    std::vector<unsigned long long> hc(C + 1, 0);
    hc[C] = s->rows_loc;
    if (C > 1) {
      k_partition<<<1, 64, 0, st>>>(reinterpret_cast<const unsigned long long*>(s->in_ptr.p), s->rows_loc, C,
                                    sc.chunk_rows.p);
      SS_CUDA(cudaMemcpyAsync(hc.data(), sc.chunk_rows.p, (C + 1) * 8, cudaMemcpyDeviceToHost, st));
      SS_CUDA(cudaStreamSynchronize(st));
    } else {
      SS_CUDA(cudaMemcpyAsync(sc.chunk_rows.p, hc.data(), (C + 1) * 8, cudaMemcpyHostToDevice, st));
      SS_CUDA(cudaStreamSynchronize(st));
    }
    s->chunk_rows.assign(hc.begin(), hc.end());
    s->chunk_task.assign(C + 1, 0);
    s->chunk_fix.assign(C + 1, 0);
  }

  // long-row tasks and fix rows
  s->n_tasks = s->n_fix = 0;
  if (s->rows_loc) {
    const uint32_t R = s->rows_loc;
    SS_TRY(sc.nt.reserve((size_t)R + 1));
    SS_TRY(sc.nf.reserve((size_t)R + 1));
    SS_TRY(sc.toff.reserve((size_t)R + 1));
    SS_TRY(sc.foff.reserve((size_t)R + 1));
    SS_CUDA(cudaMemsetAsync(sc.nt.p + R, 0, 4, st));
    SS_CUDA(cudaMemsetAsync(sc.nf.p + R, 0, 4, st));
    k_count_tasks<<<ss::div_up(R, 256), 256, 0, st>>>(s->in_ptr.p, R, sc.nt.p, sc.nf.p);
    size_t tmp_bytes = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tmp_bytes, sc.nt.p, sc.toff.p, (int)(R + 1), st);
    SS_TRY(sc.tmp.reserve(tmp_bytes));
    tmp_bytes = sc.tmp.n;
    SS_CUDA(cub::DeviceScan::ExclusiveSum(sc.tmp.p, tmp_bytes, sc.nt.p, sc.toff.p, (int)(R + 1), st));
    tmp_bytes = sc.tmp.n;
    SS_CUDA(cub::DeviceScan::ExclusiveSum(sc.tmp.p, tmp_bytes, sc.nf.p, sc.foff.p, (int)(R + 1), st));
    uint32_t n_tasks = 0, n_fix = 0;
    SS_CUDA(cudaMemcpyAsync(&n_tasks, sc.toff.p + R, 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaMemcpyAsync(&n_fix, sc.foff.p + R, 4, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    s->n_tasks = n_tasks;
    s->n_fix = n_fix;
    for (int c = 0; c <= s->n_chunks; ++c) {  // task / fix-row boundaries of the chunks (scans are in row order)
      SS_CUDA(cudaMemcpyAsync(&s->chunk_task[c], sc.toff.p + s->chunk_rows[c], 4, cudaMemcpyDeviceToHost, st));
      SS_CUDA(cudaMemcpyAsync(&s->chunk_fix[c], sc.foff.p + s->chunk_rows[c], 4, cudaMemcpyDeviceToHost, st));
    }
    SS_CUDA(cudaStreamSynchronize(st));
    if (n_tasks) {
      SS_TRY(sc.unsorted.reserve(n_tasks));
      SS_TRY(s->tasks.reserve(n_tasks));
      SS_TRY(s->fix.reserve(n_fix));
      SS_TRY(sc.keys.reserve(n_tasks));
      SS_TRY(sc.keys_out.reserve(n_tasks));
      SS_TRY(sc.order.reserve(n_tasks));
      SS_TRY(sc.order_out.reserve(n_tasks));
      k_fill_tasks<<<ss::div_up(R, 256), 256, 0, st>>>(s->in_ptr.p, s->in_src.p, R, sc.nt.p, sc.toff.p, sc.foff.p,
                                                       sc.unsorted.p, sc.keys.p, sc.order.p, s->fix.p,
                                                       sc.chunk_rows.p, s->n_chunks);
      size_t sort_bytes = 0;
      cub::DeviceRadixSort::SortPairs(nullptr, sort_bytes, sc.keys.p, sc.keys_out.p, sc.order.p, sc.order_out.p,
                                      (int)n_tasks, 0, 36, st);
      SS_TRY(sc.tmp.reserve(sort_bytes));
      sort_bytes = sc.tmp.n;
      SS_CUDA(cub::DeviceRadixSort::SortPairs(sc.tmp.p, sort_bytes, sc.keys.p, sc.keys_out.p, sc.order.p,
                                              sc.order_out.p, (int)n_tasks, 0, 36, st));
      k_gather_tasks<<<ss::div_up(n_tasks, 256), 256, 0, st>>>(sc.unsorted.p, sc.order_out.p, n_tasks, s->tasks.p);
    }
  }
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  s->tele_T = 0;  // teleport weights belong to the previous partition
  s->stats = ss_pagerank_stats{};
  s->stats.n_nodes = N;
  s->stats.n_edges = E;
  s->stats.row_lo = s->row_lo;
  s->stats.local_rows = s->rows_loc;
  s->stats.local_edges = s->E_loc;
  s->stats.load_ms =
      std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
  s->loaded = true;
  return SS_OK;
}

// rank rows [lo, hi) of the last result -> host.  The idle state buffer y[cur ^ 1] is the staging area.
static int unscale_and_copy(ss_engine* e, PagerankState* s, uint64_t lo, uint64_t hi, double* out_rank) {
  if (hi <= lo || s->T == 0) return SS_OK;
  cudaStream_t st = e->stream;
  const uint64_t total = (hi - lo) * (uint64_t)s->T;
  double* stage = s->y[s->cur ^ 1].p;
  k_unscale<<<ss::div_up(total, 256), 256, 0, st>>>(s->y[s->cur].p, s->outdeg.p, s->damping, s->TP, s->T, lo, hi, stage);
  SS_CUDA(cudaMemcpyAsync(out_rank, stage, total * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_CUDA(cudaGetLastError());
  s->stats.launches += 1;
  return SS_OK;
}

extern "C" {

SS_API int ss_graph_load_csr(ss_engine* e, uint64_t n_nodes, uint64_t n_edges, const uint64_t* row_ptr,
                             const uint32_t* col_idx) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_graph_load_csr: engine is NULL");
  SS_REQUIRE(row_ptr && (col_idx || n_edges == 0), SS_ERR_INVALID, "ss_graph_load_csr: NULL array");
  SS_REQUIRE(n_nodes < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_graph_load_csr: node ids are 32 bit");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  auto t_begin = std::chrono::steady_clock::now();
  if (!e->pr) {
    e->pr = new (std::nothrow) PagerankState();
    SS_REQUIRE(e->pr, SS_ERR_OOM, "host allocation failed");
    SS_TRY(create_sync_objects(e->pr));
  }
  PagerankState* s = e->pr;  // device blocks are reused across loads
  PagerankState::Scratch& sc = s->sc;
  s->loaded = false;
  s->have_result = false;

  cudaStream_t st = e->stream;
  const uint64_t N = n_nodes, E = n_edges;
  const int world = comm_world(e), rank = comm_rank(e);
  SS_REQUIRE(world < 64, SS_ERR_INVALID, "world size %d too large", world);
  s->N = N;
  s->E = E;
  s->bounds.assign(world + 1, 0);

  SS_TRY(sc.row_ptr.reserve(N + 1));
  SS_TRY(sc.col.reserve(E));
  SS_TRY(sc.src.reserve(E));
  SS_TRY(sc.col_sorted.reserve(E));
  SS_TRY(sc.src_sorted.reserve(E));
  SS_TRY(sc.in_ptr_full.reserve(N + 2));
  SS_TRY(sc.bounds.reserve(world + 1));
  SS_TRY(sc.bad.reserve(1));
  SS_TRY(s->outdeg.reserve(N));
  SS_CUDA(cudaMemcpyAsync(sc.row_ptr.p, row_ptr, (N + 1) * 8, cudaMemcpyHostToDevice, st));
  if (E) SS_CUDA(cudaMemcpyAsync(sc.col.p, col_idx, E * 4, cudaMemcpyHostToDevice, st));
  SS_CUDA(cudaMemsetAsync(sc.bad.p, 0, sizeof(int), st));
  if (N) {
    k_check_row_ptr<<<ss::div_up(N, 256), 256, 0, st>>>(sc.row_ptr.p, N, E, sc.bad.p);
    k_outdeg<<<ss::div_up(N, 256), 256, 0, st>>>(sc.row_ptr.p, N, s->outdeg.p);
  }
  int bad = 0;
  SS_CUDA(cudaMemcpyAsync(&bad, sc.bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_REQUIRE(!bad, SS_ERR_INVALID, "ss_graph_load_csr: row_ptr not monotone or row_ptr[n] != n_edges");
  if (E) k_expand_src4<<<ss::div_up(ss::div_up(E, 4), 256), 256, 0, st>>>(sc.row_ptr.p, N, E, sc.col.p, sc.src.p, sc.bad.p, N);
  SS_CUDA(cudaMemcpyAsync(&bad, sc.bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));

  // sort edges by child (stable: parents stay ascending inside a row)
  if (E) {
    int end_bit = 1;
    while ((1ull << end_bit) < N) ++end_bit;
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, sc.col.p, sc.col_sorted.p, sc.src.p, sc.src_sorted.p,
                                    (int64_t)E, 0, end_bit, st);
    SS_TRY(sc.tmp.reserve(tmp_bytes));
    tmp_bytes = sc.tmp.n;
    SS_CUDA(cub::DeviceRadixSort::SortPairs(sc.tmp.p, tmp_bytes, sc.col.p, sc.col_sorted.p, sc.src.p,
                                            sc.src_sorted.p, (int64_t)E, 0, end_bit, st));
  }
  k_in_ptr_from_sorted<<<ss::div_up(N + 1, 256), 256, 0, st>>>(sc.col_sorted.p, E, N, sc.in_ptr_full.p);
  k_partition<<<1, 64, 0, st>>>(sc.in_ptr_full.p, N, world, sc.bounds.p);
  std::vector<unsigned long long> hb(world + 1);
  SS_CUDA(cudaMemcpyAsync(hb.data(), sc.bounds.p, (world + 1) * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  SS_REQUIRE(!bad, SS_ERR_INVALID, "ss_graph_load_csr: child id out of range");
  for (int r = 0; r <= world; ++r) s->bounds[r] = hb[r];
  s->row_lo = s->bounds[rank];
  s->rows_loc = (uint32_t)(s->bounds[rank + 1] - s->bounds[rank]);

  unsigned long long e_lo = 0, e_hi = 0;
  SS_CUDA(cudaMemcpyAsync(&e_lo, sc.in_ptr_full.p + s->row_lo, 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaMemcpyAsync(&e_hi, sc.in_ptr_full.p + s->row_lo + s->rows_loc, 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  s->E_loc = e_hi - e_lo;
  SS_TRY(s->in_src.reserve(s->E_loc));
  if (s->E_loc)
    SS_CUDA(cudaMemcpyAsync(s->in_src.p, sc.src_sorted.p + e_lo, s->E_loc * 4, cudaMemcpyDeviceToDevice, st));
  SS_TRY(s->in_ptr.reserve((size_t)s->rows_loc + 1));
  k_local_ptr<<<ss::div_up((uint64_t)s->rows_loc + 1, 256), 256, 0, st>>>(sc.in_ptr_full.p, s->row_lo, s->rows_loc,
                                                                         s->in_ptr.p);

  return finish_load(e, s, t_begin);
}

/* Sharded export: every rank passes the slice [row_lo, row_hi) of the out-edge CSR it exported (spaghetti.h).
 * Each slice is sorted by child on its own GPU, in-degrees are summed over the ranks, the partition is cut,
 * and every rank receives exactly the in-edges of the rows it owns (one all-to-all over NVLink): H2D bytes,
 * sort work and scratch memory per rank are 1/world of the replicated load. */
SS_API int ss_graph_load_csr_rows(ss_engine* e, uint64_t n_nodes, uint64_t row_lo, uint64_t row_hi,
                                  const uint64_t* row_ptr, const uint32_t* col_idx) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_graph_load_csr_rows: engine is NULL");
  SS_REQUIRE(row_ptr && row_lo <= row_hi && row_hi <= n_nodes, SS_ERR_INVALID, "ss_graph_load_csr_rows: bad slice");
  SS_REQUIRE(n_nodes < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_graph_load_csr_rows: node ids are 32 bit");
  SS_REQUIRE(row_ptr[0] == 0, SS_ERR_INVALID, "ss_graph_load_csr_rows: row_ptr is local to the slice (starts at 0)");
  const uint64_t n_slice = row_hi - row_lo, E_slice = row_ptr[n_slice];
  SS_REQUIRE(col_idx || E_slice == 0, SS_ERR_INVALID, "ss_graph_load_csr_rows: col_idx is NULL");
  SS_REQUIRE(E_slice < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_graph_load_csr_rows: slice has >= 2^32 edges");
  if (comm_world(e) == 1) {
    SS_REQUIRE(row_lo == 0 && row_hi == n_nodes, SS_ERR_INVALID,
               "ss_graph_load_csr_rows: a single rank must pass every row");
    return ss_graph_load_csr(e, n_nodes, E_slice, row_ptr, col_idx);
  }
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  auto t_begin = std::chrono::steady_clock::now();
  if (!e->pr) {
    e->pr = new (std::nothrow) PagerankState();
    SS_REQUIRE(e->pr, SS_ERR_OOM, "host allocation failed");
    SS_TRY(create_sync_objects(e->pr));
  }
  PagerankState* s = e->pr;
  PagerankState::Scratch& sc = s->sc;
  s->loaded = false;
  s->have_result = false;
  cudaStream_t st = e->stream;
  const uint64_t N = n_nodes;
  const int world = comm_world(e), rank = comm_rank(e);
  SS_REQUIRE(world < 64, SS_ERR_INVALID, "world size %d too large", world);
  s->N = N;
  s->bounds.assign(world + 1, 0);

  // Local phase: no collective in here.  Its status is agreed on below, so that a rank with a bad slice
  // or no memory makes every rank return instead of leaving the others inside a collective.
  int end_bit = 1;
  while ((1ull << end_bit) < N) ++end_bit;
  auto local_phase = [&]() -> int {
    SS_TRY(sc.row_ptr.reserve(n_slice + 1));
    SS_TRY(sc.col.reserve(E_slice));
    SS_TRY(sc.src.reserve(E_slice));
    SS_TRY(sc.col_sorted.reserve(E_slice));
    SS_TRY(sc.src_sorted.reserve(E_slice));
    SS_TRY(sc.slice_ptr.reserve(N + 2));
    SS_TRY(sc.cnt.reserve(N + 1));
    SS_TRY(sc.keys.reserve(N + 2));
    SS_TRY(sc.in_ptr_full.reserve(N + 2));
    SS_TRY(sc.bounds.reserve(world + 1));
    SS_TRY(sc.bad.reserve(1));
    SS_TRY(s->outdeg.reserve(N));
    size_t tmp_bytes = 0, scan_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, sc.col.p, sc.col_sorted.p, sc.src.p, sc.src_sorted.p,
                                    (int64_t)std::max<uint64_t>(E_slice, 1), 0, end_bit, st);
    cub::DeviceScan::ExclusiveSum(nullptr, scan_bytes, sc.keys.p, sc.in_ptr_full.p, (int64_t)(N + 1), st);
    SS_TRY(sc.tmp.reserve(std::max(tmp_bytes, scan_bytes)));
    SS_CUDA(cudaMemcpyAsync(sc.row_ptr.p, row_ptr, (n_slice + 1) * 8, cudaMemcpyHostToDevice, st));
    if (E_slice) SS_CUDA(cudaMemcpyAsync(sc.col.p, col_idx, E_slice * 4, cudaMemcpyHostToDevice, st));
    SS_CUDA(cudaMemsetAsync(sc.bad.p, 0, sizeof(int), st));
    if (n_slice) {
      k_check_row_ptr<<<ss::div_up(n_slice, 256), 256, 0, st>>>(sc.row_ptr.p, n_slice, E_slice, sc.bad.p);
      k_outdeg<<<ss::div_up(n_slice, 256), 256, 0, st>>>(sc.row_ptr.p, n_slice, s->outdeg.p + row_lo);
    }
    if (E_slice) {
      k_expand_src4<<<ss::div_up(ss::div_up(E_slice, 4), 256), 256, 0, st>>>(sc.row_ptr.p, n_slice, E_slice, sc.col.p,
                                                                            sc.src.p, sc.bad.p, N);
      if (row_lo) k_add_u32<<<ss::div_up(E_slice, 256), 256, 0, st>>>(sc.src.p, E_slice, (uint32_t)row_lo);
      tmp_bytes = sc.tmp.n;
      SS_CUDA(cub::DeviceRadixSort::SortPairs(sc.tmp.p, tmp_bytes, sc.col.p, sc.col_sorted.p, sc.src.p,
                                              sc.src_sorted.p, (int64_t)E_slice, 0, end_bit, st));
    }
    // slice_ptr[v] = first position of child v in the slice's sorted list; cnt[v] = its in-edges from this slice
    k_in_ptr_from_sorted<<<ss::div_up(N + 1, 256), 256, 0, st>>>(sc.col_sorted.p, E_slice, N, sc.slice_ptr.p);
    if (N) k_count_from_ptr<<<ss::div_up(N, 256), 256, 0, st>>>(sc.slice_ptr.p, N, sc.cnt.p);
    int bad = 0;
    SS_CUDA(cudaMemcpyAsync(&bad, sc.bad.p, sizeof(int), cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    SS_CUDA(cudaGetLastError());
    SS_REQUIRE(!bad, SS_ERR_INVALID, "ss_graph_load_csr_rows: row_ptr not monotone or child id out of range");
    return SS_OK;
  };
  struct Slice {
    uint64_t lo, hi, edges;
    int64_t status;
  } mine{row_lo, row_hi, E_slice, 0}, all[64];
  mine.status = local_phase();
  SS_TRY(comm_allgather_host_bytes(e, &mine, sizeof(mine), all));
  uint64_t covered = 0, E = 0;
  for (int r = 0; r < world; ++r) {
    if (all[r].status < 0) {
      if (mine.status >= 0) ss::set_error("ss_graph_load_csr_rows: rank %d failed its local phase (%lld)", r,
                                          (long long)all[r].status);
      return mine.status < 0 ? (int)mine.status : SS_ERR_STATE;
    }
    covered += all[r].hi - all[r].lo;
    E += all[r].edges;
  }
  SS_REQUIRE(covered == N, SS_ERR_INVALID, "ss_graph_load_csr_rows: the ranks' slices cover %llu of %llu rows",
             (unsigned long long)covered, (unsigned long long)N);
  s->E = E;

  // out-degree of every node (each rank contributes its slice); in-degrees summed over the slices
  {
    std::vector<size_t> off(world), cnt(world);
    for (int r = 0; r < world; ++r) {
      off[r] = all[r].lo * 4;
      cnt[r] = (all[r].hi - all[r].lo) * 4;
    }
    SS_TRY(comm_allgatherv_bytes(e, s->outdeg.p, off.data(), cnt.data()));
  }
  SS_TRY(comm_allreduce_sum_u32(e, sc.cnt.p, N));
  k_widen_counts<<<ss::div_up(N + 1, 256), 256, 0, st>>>(sc.cnt.p, N, sc.keys.p);
  {
    size_t scan_bytes = sc.tmp.n;
    SS_CUDA(cub::DeviceScan::ExclusiveSum(sc.tmp.p, scan_bytes, sc.keys.p, sc.in_ptr_full.p, (int64_t)(N + 1), st));
  }
  k_partition<<<1, 64, 0, st>>>(sc.in_ptr_full.p, N, world, sc.bounds.p);
  std::vector<unsigned long long> hb(world + 1);
  SS_CUDA(cudaMemcpyAsync(hb.data(), sc.bounds.p, (world + 1) * 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  for (int r = 0; r <= world; ++r) s->bounds[r] = hb[r];
  s->row_lo = s->bounds[rank];
  s->rows_loc = (uint32_t)(s->bounds[rank + 1] - s->bounds[rank]);

  // all-to-all of the edges: rank q owns children [bounds[q], bounds[q+1]) = a contiguous piece of every
  // slice's child-sorted list
  std::vector<unsigned long long> cut(world + 1);
  for (int r = 0; r <= world; ++r)
    SS_CUDA(cudaMemcpyAsync(&cut[r], sc.slice_ptr.p + s->bounds[r], 8, cudaMemcpyDeviceToHost, st));
  SS_CUDA(cudaStreamSynchronize(st));
  std::vector<uint64_t> send_cnt64(world), matrix((size_t)world * world);
  for (int r = 0; r < world; ++r) send_cnt64[r] = cut[r + 1] - cut[r];
  SS_TRY(comm_allgather_host_bytes(e, send_cnt64.data(), world * 8, matrix.data()));
  std::vector<size_t> send_off(world), send_cnt(world), recv_off(world), recv_cnt(world);
  uint64_t E_loc = 0;
  for (int r = 0; r < world; ++r) {
    send_off[r] = cut[r];
    send_cnt[r] = send_cnt64[r];
    recv_off[r] = E_loc;
    recv_cnt[r] = matrix[(size_t)r * world + rank];
    E_loc += recv_cnt[r];
  }
  // second allocation round, status agreed on again
  auto alloc2 = [&]() -> int {
    SS_REQUIRE(E_loc < 0xFFFFFFFFull, SS_ERR_INVALID, "ss_graph_load_csr_rows: %llu local edges; use more ranks",
               (unsigned long long)E_loc);
    SS_TRY(sc.recv_col.reserve(E_loc));
    SS_TRY(sc.recv_src.reserve(E_loc));
    SS_TRY(sc.col.reserve(E_loc));
    SS_TRY(s->in_src.reserve(E_loc));
    SS_TRY(s->in_ptr.reserve((size_t)s->rows_loc + 1));
    size_t tmp_bytes = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, sc.recv_col.p, sc.col.p, sc.recv_src.p, s->in_src.p,
                                    (int64_t)std::max<uint64_t>(E_loc, 1), 0, end_bit, st);
    SS_TRY(sc.tmp.reserve(tmp_bytes));
    return SS_OK;
  };
  int64_t st2 = alloc2(), all2[64];
  SS_TRY(comm_allgather_host_bytes(e, &st2, sizeof(st2), all2));
  for (int r = 0; r < world; ++r)
    if (all2[r] < 0) {
      if (st2 >= 0) ss::set_error("ss_graph_load_csr_rows: rank %d could not allocate its partition", r);
      return st2 < 0 ? (int)st2 : SS_ERR_STATE;
    }
  {
    const uint32_t* send[2] = {sc.col_sorted.p, sc.src_sorted.p};
    uint32_t* recv[2] = {sc.recv_col.p, sc.recv_src.p};
    SS_TRY(comm_alltoallv_u32(e, 2, send, send_off.data(), send_cnt.data(), recv, recv_off.data(), recv_cnt.data()));
  }
  // pieces arrive in rank order, each sorted by (child, parent): a stable sort by child leaves the parents of a
  // row ascending when the slices ascend with the rank
  if (E_loc) {
    size_t tmp_bytes = sc.tmp.n;
    SS_CUDA(cub::DeviceRadixSort::SortPairs(sc.tmp.p, tmp_bytes, sc.recv_col.p, sc.col.p, sc.recv_src.p, s->in_src.p,
                                            (int64_t)E_loc, 0, end_bit, st));
  }
  s->E_loc = E_loc;
  k_in_ptr_local<<<ss::div_up((uint64_t)s->rows_loc + 1, 256), 256, 0, st>>>(sc.col.p, E_loc, s->row_lo, s->rows_loc,
                                                                            s->in_ptr.p);
  return finish_load(e, s, t_begin);
}

SS_API int ss_pagerank(ss_engine* e, double damping, double eps, uint32_t n_topics, const int64_t* num_pages,
                       uint32_t max_iters, double* out_rank, uint32_t* out_iters) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_pagerank: engine is NULL");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  PagerankState* s = e->pr;
  SS_REQUIRE(s && s->loaded, SS_ERR_STATE, "ss_pagerank: no graph loaded");
  SS_REQUIRE(n_topics == 0 || num_pages, SS_ERR_INVALID, "ss_pagerank: num_pages is NULL");
  SS_REQUIRE(n_topics <= 16, SS_ERR_INVALID,
             "ss_pagerank: %u topics; run topics in slabs of <= 16 (they are independent)", n_topics);
  s->have_result = false;
  s->stats.sweeps = 0;
  s->stats.launches = 0;
  s->stats.sweep_ms_total = s->stats.gather_ms_total = s->stats.exchange_ms_total = s->stats.short_ms_total = 0;
  s->stats.exchange_busy_ms_total = 0;
  if (n_topics == 0) {  // empty forw[5]: every node gets {} (pagerank.go:53-63)
    s->T = 0;
    s->have_result = true;
    return SS_OK;
  }
  cudaStream_t st = e->stream;
  const int world = comm_world(e);
  const Shape shape = shape_for_topics(n_topics);
  const int LPR = shape.lpr, TP = shape.lpr * shape.vec, T = (int)n_topics;
  const uint64_t N = s->N;
  const uint32_t R = s->rows_loc;
  const bool timing = (e->flags & SS_FLAG_TIMING) != 0;
  s->TP = TP;
  s->T = T;
  s->damping = damping;

  const int GPC = (32 / LPR) * (kThreads / 32);
  const uint32_t n_row_blocks = ss::div_up(R, GPC);
  int occ_short = 4, occ_long = 4, occ_short32 = 4;
  dispatch_shape(shape, [&](auto lpr, auto vec) {
    constexpr int L = decltype(lpr)::value, V = decltype(vec)::value;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_short, k_sweep_short<L, V>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_short32, k_sweep_short32<L, V, false, true>, kThreads, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_long, k_sweep_long<L, V>, kThreads, 0);
    return SS_OK;
  });
  // persistent grids: exactly one resident wave, static block-cyclic work split
  // lean short-row kernel: needs the 32-bit pointer copy and its padding to cover two grid strides
  const char* short_env = getenv("SS_PR_SHORT");
  bool lean = s->ptr32_pad > 0 && !(short_env && !strcmp(short_env, "legacy"));
  if (lean) {
    const uint64_t need = 2ull * (uint64_t)e->sm_count * std::max(1, occ_short32) * GPC + GPC + 2;
    if (need > s->ptr32_pad) lean = false;
  }
  if (lean) occ_short = occ_short32;
  // async gather ring (SS_PR_SHORT=async at load and run time): needs the short-only CSR and a
  // 16-byte-per-lane shape
  const bool want_async = short_env && !strcmp(short_env, "async");
  const bool use_async = want_async && s->have_short && shape.vec == 2 && R > 0;
  const size_t async_smem = (size_t)(kAsyncThreads / 32) * kRingBytes;
  uint32_t grid_short =
      std::max(1u, std::min<uint32_t>(n_row_blocks, (uint32_t)(e->sm_count * std::max(1, occ_short))));
  if (use_async) {
    int occ_async = 1;
    dispatch_shape(shape, [&](auto lpr, auto) {
      constexpr int L = decltype(lpr)::value;
      auto set = [&](auto kern) {
        cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)async_smem);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_async, kern, kAsyncThreads, async_smem);
      };
      set(k_sweep_short_async<L, true, true>);
      set(k_sweep_short_async<L, true, false>);
      set(k_sweep_short_async<L, false, true>);
      set(k_sweep_short_async<L, false, false>);
      return SS_OK;
    });
    SS_CUDA(cudaGetLastError());
    grid_short = std::max(1u, std::min<uint32_t>(ss::div_up(R, 64), (uint32_t)(e->sm_count * std::max(1, occ_async))));
    const uint32_t n_warps = grid_short * (kAsyncThreads / 32);
    SS_TRY(s->warp_rows.reserve((size_t)n_warps + 1));
    k_short_partition<<<ss::div_up((uint64_t)n_warps + 1, 256), 256, 0, st>>>(s->sptr.p, R, n_warps, s->warp_rows.p);
  }
  // Per-chunk launch shapes (one chunk unless the exchange is overlapped, see PagerankState::n_chunks).
  // Chunking needs the lean short-row kernel (the others walk the whole row range).
  const bool chunked = s->n_chunks > 1 && lean && !use_async && world > 1;
  const int C = chunked ? s->n_chunks : 1;
  struct ChunkShape {
    uint32_t row_begin, row_end, n_row_blocks, grid_short, task_begin, n_tasks, grid_long, fix_begin, n_fix, grid_fix;
    uint32_t slot0;  // first reduction slot of the chunk's three launches
  };
  std::vector<ChunkShape> chunks(C);
  uint32_t sweep_slots = 0;
  for (int c = 0; c < C; ++c) {
    ChunkShape& k = chunks[c];
    k.row_begin = chunked ? s->chunk_rows[c] : 0u;
    k.row_end = chunked ? s->chunk_rows[c + 1] : R;
    k.n_row_blocks = ss::div_up(k.row_end - k.row_begin, GPC);
    k.grid_short = chunked ? std::max(1u, std::min<uint32_t>(k.n_row_blocks, (uint32_t)(e->sm_count * std::max(1, occ_short))))
                           : grid_short;
    k.task_begin = chunked ? s->chunk_task[c] : 0u;
    k.n_tasks = chunked ? s->chunk_task[c + 1] - s->chunk_task[c] : s->n_tasks;
    k.grid_long = std::max(1u, std::min<uint32_t>(ss::div_up(k.n_tasks, kThreads / 32),
                                                  (uint32_t)(e->sm_count * std::max(1, occ_long))));
    k.fix_begin = chunked ? s->chunk_fix[c] : 0u;
    k.n_fix = chunked ? s->chunk_fix[c + 1] - s->chunk_fix[c] : s->n_fix;
    k.grid_fix = std::max(1u, std::min<uint32_t>(k.n_fix, (uint32_t)e->sm_count * 8));
    k.slot0 = sweep_slots;
    sweep_slots += k.grid_short + k.grid_long + k.grid_fix;
  }
  const uint32_t grid_init = (uint32_t)e->sm_count * 8;
  const uint32_t red_slots = std::max(sweep_slots, grid_init);
  const int W = 3 * TP;

  // Every allocation of the run happens here, before the first collective, and the ranks agree on the outcome:
  // a rank that runs out of memory makes all of them return instead of leaving the others inside NCCL
  // (ADVICE round 1).
  auto allocate = [&]() -> int {
    // multi-rank: fixed-size state so that the peer mappings survive topic-count changes
    SS_TRY(s->y[0].reserve(world > 1 ? N * 16 : N * TP));
    SS_TRY(s->y[1].reserve(world > 1 ? N * 16 : N * TP));
    SS_TRY(s->mul.reserve(R));
    SS_TRY(s->mul2.reserve(R));
    SS_TRY(s->partials.reserve((size_t)s->n_tasks * TP));
    SS_TRY(s->red.reserve((size_t)red_slots * W));
    SS_TRY(s->sums.reserve(W));
    SS_TRY(s->stage.reserve((size_t)kReduceCtas * W));
    SS_TRY(s->tot.reserve(TP));
    SS_TRY(s->init.reserve(TP));
    if (s->tele_T) SS_TRY(s->tele_w.reserve((size_t)R * TP));
    return SS_OK;
  };
  const int64_t alloc_status = allocate();
  if (world > 1) {
    int64_t all_status[64];
    SS_TRY(comm_allgather_host_bytes(e, &alloc_status, sizeof(alloc_status), all_status));
    for (int r = 0; r < world; ++r)
      if (all_status[r] < 0) {
        if (alloc_status >= 0) ss::set_error("ss_pagerank: rank %d could not allocate its state", r);
        return alloc_status < 0 ? (int)alloc_status : SS_ERR_OOM;
      }
    SS_TRY(open_peers(e, s));
    SS_TRY(agree_on_fused(e, s));
  } else if (alloc_status < 0) {
    return (int)alloc_status;
  }

  const bool biased = s->tele_T != 0;
  if (biased) {
    SS_REQUIRE(s->tele_T == n_topics, SS_ERR_INVALID, "ss_pagerank: teleport weights were set for %u topics, run has %u",
               s->tele_T, n_topics);
    if (R) k_pad_rows<<<ss::div_up((uint64_t)R * TP, 256), 256, 0, st>>>(s->tele_raw.p, R, T, TP, s->tele_w.p);
  }
  double h_init[16];
  for (int t = 0; t < TP; ++t) h_init[t] = t < T ? 1.0 / (double)num_pages[t] : 0.0;  // pagerank.go:104
  SS_CUDA(cudaMemcpyAsync(s->init.p, h_init, TP * 8, cudaMemcpyHostToDevice, st));

  const double tele = 1.0 - damping;  // pagerank.go:90
  // exchange layout: every rank's row block, cut into the same number of chunks on every rank
  std::vector<uint64_t> all_rows((size_t)world * (C + 1));
  if (world > 1) {
    std::vector<uint64_t> mine(C + 1);
    for (int c = 0; c <= C; ++c) mine[c] = s->row_lo + (chunked ? s->chunk_rows[c] : (c == C ? R : 0u));
    SS_TRY(comm_allgather_host_bytes(e, mine.data(), (C + 1) * 8, all_rows.data()));
  }
  std::vector<size_t> byte_off(world), byte_cnt(world);
  auto chunk_bytes = [&](int c) {
    for (int r = 0; r < world; ++r) {
      const uint64_t lo = all_rows[(size_t)r * (C + 1) + c], hi = all_rows[(size_t)r * (C + 1) + c + 1];
      byte_off[r] = (size_t)lo * TP * 8;
      byte_cnt[r] = (size_t)(hi - lo) * TP * 8;
    }
  };

  // y0, mul, S_0
  int rc = dispatch_shape(shape, [&](auto lpr, auto vec) {
    constexpr int L = decltype(lpr)::value, V = decltype(vec)::value;
    k_init<L, V><<<grid_init, kThreads, 0, st>>>(s->y[0].p, s->outdeg.p, N, damping, s->init.p, s->row_lo, R, s->mul.p,
                                             s->mul2.p, s->in_ptr.p, s->red.p);
    return SS_OK;
  });
  SS_TRY(rc);
  k_reduce_partials<<<kReduceCtas, kThreads, 0, st>>>(s->red.p, grid_init, W, s->stage.p);
  k_fold_stage<<<1, 64, 0, st>>>(s->stage.p, kReduceCtas, W, s->sums.p);
  SS_TRY(comm_allreduce_sum_f64(e, s->sums.p, W));
  k_finish_tot<<<1, 32, 0, st>>>(s->sums.p, TP, tele, (double)N, s->tot.p);
  s->stats.launches += 4;
  s->cur = 0;

  const uint32_t all_topics = T >= 32 ? 0xFFFFFFFFu : ((1u << T) - 1u);
  uint32_t active = all_topics;
  std::vector<uint32_t> iters(T, 0);
  std::vector<double> h_sums(W);
  bool hit_max = false;
  for (uint32_t sweep = 1; active; ++sweep) {
    SweepParams p;
    p.y_last = s->y[s->cur].p;
    p.y_next = s->y[s->cur ^ 1].p;
    p.in_ptr = s->in_ptr.p;
    p.in_src = s->in_src.p;
    p.mul = s->mul.p;
    p.in_ptr32 = s->in_ptr32.p;
    p.mul2 = s->mul2.p;
    p.sptr = s->sptr.p;
    p.ssrc = s->ssrc.p;
    p.warp_rows = s->warp_rows.p;
    p.inv_tot = s->tot.p;
    p.init = s->init.p;
    p.tele_w = biased ? s->tele_w.p : nullptr;
    p.row_lo = s->row_lo;
    p.row_begin = 0;
    p.rows_loc = R;
    p.active_mask = active;
    p.tele = tele;
    p.first = sweep == 1;
    const bool epi_push = s->fused && s->push_from_epilogue;
    p.n_peers = epi_push ? s->n_peers : 0;
    for (int q = 0; q < kMaxPeers; ++q) p.peer_next[q] = epi_push ? s->peer_y[s->cur ^ 1][q] : nullptr;
    if (timing) SS_CUDA(cudaEventRecord(s->ev[0], st));
    const bool overlap = world > 1 && !epi_push;  // exchange on its own stream, chunk by chunk
    for (int c = 0; c < C; ++c) {
      const ChunkShape& k = chunks[c];
      rc = dispatch_shape(shape, [&](auto lpr, auto vec) {
        constexpr int L = decltype(lpr)::value, V = decltype(vec)::value;
        SweepParams q = p;
        q.row_begin = k.row_begin;
        q.rows_loc = k.row_end;
        q.red = s->red.p + (size_t)k.slot0 * W;
        const bool all = active == all_topics && T == TP;  // padded columns stay frozen through the mask
        if (use_async) {
          if (p.first) {
            if (all) k_sweep_short_async<L, true, true><<<k.grid_short, kAsyncThreads, async_smem, st>>>(q);
            else k_sweep_short_async<L, true, false><<<k.grid_short, kAsyncThreads, async_smem, st>>>(q);
          } else {
            if (all) k_sweep_short_async<L, false, true><<<k.grid_short, kAsyncThreads, async_smem, st>>>(q);
            else k_sweep_short_async<L, false, false><<<k.grid_short, kAsyncThreads, async_smem, st>>>(q);
          }
        } else if (lean) {
          if (p.first) {
            if (all) k_sweep_short32<L, V, true, true><<<k.grid_short, kThreads, 0, st>>>(q, k.n_row_blocks);
            else k_sweep_short32<L, V, true, false><<<k.grid_short, kThreads, 0, st>>>(q, k.n_row_blocks);
          } else {
            if (all) k_sweep_short32<L, V, false, true><<<k.grid_short, kThreads, 0, st>>>(q, k.n_row_blocks);
            else k_sweep_short32<L, V, false, false><<<k.grid_short, kThreads, 0, st>>>(q, k.n_row_blocks);
          }
        } else {
          k_sweep_short<L, V><<<k.grid_short, kThreads, 0, st>>>(q, k.n_row_blocks);
        }
        if (timing && c == C - 1) cudaEventRecord(s->ev[4], st);
        q.rows_loc = R;
        q.red = s->red.p + (size_t)(k.slot0 + k.grid_short) * W;
        k_sweep_long<L, V><<<k.grid_long, kThreads, 0, st>>>(q, s->tasks.p + k.task_begin, k.n_tasks, s->partials.p);
        if (timing && c == C - 1) cudaEventRecord(s->ev[1], st);
        q.red = s->red.p + (size_t)(k.slot0 + k.grid_short + k.grid_long) * W;
        k_sweep_fix<L, V><<<k.grid_fix, kThreads, 0, st>>>(q, s->fix.p + k.fix_begin, k.n_fix, s->partials.p);
        return SS_OK;
      });
      SS_TRY(rc);
      s->stats.launches += 3;
      if (overlap) {
        // chunk c of every rank is final: ship it while the next chunk computes
        SS_CUDA(cudaEventRecord(s->chunk_ev[c], st));
        SS_CUDA(cudaStreamWaitEvent(s->xstream, s->chunk_ev[c], 0));
        if (timing && c == 0) SS_CUDA(cudaEventRecord(s->x_t0, s->xstream));
        chunk_bytes(c);
        if (s->fused) {
          // copy-engine push: my chunk into every peer's copy of the state, starting with a different peer
          // on every rank so that the receivers' links fill evenly
          const int rank = comm_rank(e);
          for (int i = 0; i < s->n_peers && byte_cnt[rank]; ++i) {
            int q = 0;
            for (int j = 0; j < s->n_peers; ++j)
              if (s->peer_rank[j] == (rank + 1 + i) % world) q = j;
            char* dst = reinterpret_cast<char*>(s->peer_y[s->cur ^ 1][q]) + byte_off[rank];
            const char* src = reinterpret_cast<const char*>(p.y_next) + byte_off[rank];
            SS_CUDA(cudaStreamWaitEvent(s->pstream[i], s->chunk_ev[c], 0));
            SS_CUDA(cudaMemcpyAsync(dst, src, byte_cnt[rank], cudaMemcpyDefault, s->pstream[i]));
          }
        } else {
          SS_TRY(comm_allgatherv_bytes_on(e, s->xstream, p.y_next, byte_off.data(), byte_cnt.data()));
        }
      }
    }
    k_reduce_partials<<<kReduceCtas, kThreads, 0, st>>>(s->red.p, sweep_slots, W, s->stage.p);
    k_fold_stage<<<1, 64, 0, st>>>(s->stage.p, kReduceCtas, W, s->sums.p);
    if (timing) SS_CUDA(cudaEventRecord(s->ev[2], st));
    if (world > 1) {
      // every NCCL call of the loop goes to the exchange stream, in the same order on every rank; the
      // all-reduce of the 3*TP sums closes the sweep (fused exchange: it is also the barrier that orders
      // every rank's pushed rows before anybody's next sweep)
      SS_CUDA(cudaEventRecord(s->red_done, st));
      SS_CUDA(cudaStreamWaitEvent(s->xstream, s->red_done, 0));
      if (overlap && s->fused)  // the peer copies of every chunk precede the all-reduce
        for (int i = 0; i < s->n_peers; ++i) {
          SS_CUDA(cudaEventRecord(s->pdone[i], s->pstream[i]));
          SS_CUDA(cudaStreamWaitEvent(s->xstream, s->pdone[i], 0));
        }
      if (timing && !overlap) SS_CUDA(cudaEventRecord(s->x_t0, s->xstream));
      SS_TRY(comm_allreduce_sum_f64_on(e, s->xstream, s->sums.p, W));
      if (timing) SS_CUDA(cudaEventRecord(s->x_t1, s->xstream));
      SS_CUDA(cudaEventRecord(s->x_done, s->xstream));
      SS_CUDA(cudaStreamWaitEvent(st, s->x_done, 0));
    }
    k_finish_tot<<<1, 32, 0, st>>>(s->sums.p, TP, tele, (double)N, s->tot.p);
    if (timing) SS_CUDA(cudaEventRecord(s->ev[3], st));
    SS_CUDA(cudaMemcpyAsync(h_sums.data(), s->sums.p, W * 8, cudaMemcpyDeviceToHost, st));
    SS_CUDA(cudaStreamSynchronize(st));
    SS_CUDA(cudaGetLastError());
    s->stats.launches += 3;
    s->stats.sweeps = sweep;
    if (timing) {
      float ms = 0;
      cudaEventElapsedTime(&ms, s->ev[0], s->ev[2]);
      s->stats.sweep_ms_total += ms;
      cudaEventElapsedTime(&ms, s->ev[0], s->ev[1]);
      s->stats.gather_ms_total += ms;
      cudaEventElapsedTime(&ms, s->ev[0], s->ev[4]);
      s->stats.short_ms_total += ms;
      cudaEventElapsedTime(&ms, s->ev[2], s->ev[3]);
      s->stats.exchange_ms_total += ms;  // what the sweep loop waits for after its own kernels (exposed)
      if (world > 1) {
        cudaEventElapsedTime(&ms, s->x_t0, s->x_t1);
        s->stats.exchange_busy_ms_total += ms;  // first chunk broadcast .. end of the all-reduce
      }
    }
    s->cur ^= 1;
    for (int t = 0; t < T; ++t) {
      if (!((active >> t) & 1u)) continue;
      iters[t] = sweep;
      const double delta = h_sums[t], changed = h_sums[2 * TP + t];
      // `for ...; lastChange > convergenceCriterion; ...` (pagerank.go:93): NaN ends the loop too
      bool go_on = delta > eps;
      if (go_on && changed == 0.0) go_on = false;  // bit-for-bit fixed point
      if (go_on && max_iters && sweep >= max_iters) {
        go_on = false;
        hit_max = true;
      }
      if (!go_on) active &= ~(1u << t);
    }
  }
  s->have_result = true;
  if (out_iters) memcpy(out_iters, iters.data(), T * sizeof(uint32_t));
  if (out_rank && N) {
    // the state is replicated on every rank after the exchange: unscale every row into the idle state
    // buffer (same or smaller footprint: [N][T] vs [N][TP]) and copy it out in one piece -- no staging
    // chunks, no host round trip per chunk
    SS_TRY(unscale_and_copy(e, s, 0, N, out_rank));
  }
  return hit_max ? SS_NOT_CONVERGED : SS_OK;
}

SS_API int ss_pagerank_set_teleport(ss_engine* e, uint64_t n_nodes, uint32_t n_topics, const double* weight) {
  SS_REQUIRE(e, SS_ERR_INVALID, "ss_pagerank_set_teleport: engine is NULL");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  PagerankState* s = e->pr;
  SS_REQUIRE(s && s->loaded, SS_ERR_STATE, "ss_pagerank_set_teleport: load the graph first");
  if (!weight || n_topics == 0) {
    s->tele_T = 0;
    return SS_OK;
  }
  SS_REQUIRE(n_nodes == s->N, SS_ERR_INVALID, "ss_pagerank_set_teleport: %llu rows, graph has %llu",
             (unsigned long long)n_nodes, (unsigned long long)s->N);
  SS_REQUIRE(n_topics <= 16, SS_ERR_INVALID, "ss_pagerank_set_teleport: at most 16 topics per run");
  const size_t n = (size_t)s->rows_loc * n_topics;
  SS_TRY(s->tele_raw.reserve(n));
  if (n)  // this rank's rows are one contiguous piece of the caller's [N][T] array
    SS_CUDA(cudaMemcpyAsync(s->tele_raw.p, weight + s->row_lo * n_topics, n * 8, cudaMemcpyHostToDevice, e->stream));
  SS_CUDA(cudaStreamSynchronize(e->stream));
  s->tele_T = n_topics;
  return SS_OK;
}

SS_API int ss_pagerank_fetch(ss_engine* e, uint64_t row_lo, uint64_t row_hi, double* out_rank) {
  SS_REQUIRE(e && out_rank, SS_ERR_INVALID, "ss_pagerank_fetch: NULL argument");
  std::lock_guard<std::mutex> lock(e->mu);
  DeviceGuard guard(e->device);
  PagerankState* s = e->pr;
  SS_REQUIRE(s && s->have_result, SS_ERR_STATE, "ss_pagerank_fetch: no result");
  SS_REQUIRE(row_lo <= row_hi && row_hi <= s->N, SS_ERR_INVALID, "ss_pagerank_fetch: bad row range");
  if (s->T == 0 || row_lo == row_hi) return SS_OK;
  SS_TRY(unscale_and_copy(e, s, row_lo, row_hi, out_rank));
  return SS_OK;
}

}  // extern "C"

// Unscaled copy of the last result, [N][T], for the scoring blend (index.cu).
int pagerank_export_device(ss_engine* e, ss::DevBuf<double>* out, uint64_t* n_rows, uint32_t* n_topics) {
  PagerankState* s = e->pr;
  SS_REQUIRE(s && s->have_result, SS_ERR_STATE, "ss_use_pagerank: no PageRank result on the device");
  *n_rows = s->N;
  *n_topics = (uint32_t)s->T;
  if (s->T == 0 || s->N == 0) {
    out->reset();
    return SS_OK;
  }
  SS_TRY(out->alloc(s->N * (uint64_t)s->T));
  const uint64_t total = s->N * (uint64_t)s->T;
  k_unscale<<<ss::div_up(total, 256), 256, 0, e->stream>>>(s->y[s->cur].p, s->outdeg.p, s->damping, s->TP, s->T, 0,
                                                           s->N, out->p);
  SS_CUDA(cudaStreamSynchronize(e->stream));
  SS_CUDA(cudaGetLastError());
  return SS_OK;
}

extern "C" {

SS_API int ss_pagerank_get_stats(ss_engine* e, ss_pagerank_stats* out) {
  SS_REQUIRE(e && out, SS_ERR_INVALID, "ss_pagerank_get_stats: NULL argument");
  std::lock_guard<std::mutex> lock(e->mu);
  SS_REQUIRE(e->pr, SS_ERR_STATE, "ss_pagerank_get_stats: no graph loaded");
  *out = e->pr->stats;
  return SS_OK;
}

}  // extern "C"
