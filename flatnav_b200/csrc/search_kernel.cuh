// Warp-per-query greedy beam search over the flat navigable-small-world graph (sm_100a).
//
// Replaces, for a whole batch of queries in one launch,
//   Index::search / initializeSearch / beamSearch / processCandidateNode
//       (include/flatnav/index/Index.h:387-409, 845-870, 606-659, 661-707 of the reference)
//   the two std::priority_queue heaps (Index.h:47-53, 623-624)  -> one sorted list in shared memory
//   VisitedSet / VisitedSetPool (util/VisitedSetPool.h:16-197)   -> per-warp bucketed tag set in shared
//                                                                  memory (never a false positive, may forget)
//   the AVX-512/AVX/SSE distance kernels (util/SquaredL2SimdExtensions.h, InnerProductSimdExtensions.h)
//                                                                -> 128-bit gathers + shuffle reduction
//   executeInParallel (util/Multithreading.h:18-48)              -> persistent warps pulling query ids
//
// Formulation (oracle/flatnav_oracle.cpp `search_list` is its CPU twin, bit for bit):
//   L = list of at most B = max(ef, K) entries sorted by (distance, node id), each with an "expanded" bit.
//   repeat: pick the first unexpanded entry (stop if none), mark it; read its M links; for each link not
//   yet visited: mark visited, evaluate the distance; accept it iff |L| < B or d < worst(L) (strict);
//   merge the accepted ones into L and truncate to B.
// This equals the reference's two-heap loop except on exact distance ties (heap order among equal keys
// is unspecified there).
//
// Mapping to the machine:
//   * one warp = one query; a CTA is a bundle of independent warps sharing nothing but the SM.
//   * query vector lives in registers (CH 16-byte chunks per lane).
//   * a row is read by G lanes with ld.global.nc.L1::no_allocate.v4 (G*16 contiguous bytes per row per
//     instruction, 32/G rows per warp-wide instruction, U instructions unrolled => up to U*CH*512 B in
//     flight per warp with no shared-memory staging cost; the register file is the staging buffer).
//   * visited test happens BEFORE the gather (as the reference does, Index.h:679-685), so only fresh
//     rows are fetched.
//   * the adjacency row of the runner-up candidate is prefetched into L2 while the current one is expanded.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "fnb_layout.h"

namespace fnb {

enum { DT_F32 = 0, DT_U8 = 1, DT_I8 = 2 };
enum { M_L2 = 0, M_IP = 1 };

#define FNB_WARPS_PER_CTA 4
// Occupancy plan.  The traversal is latency-bound per warp (a hop is a chain of dependent steps), so throughput
// comes from resident warps: 24 per SM for rows up to 512 B.  Registers are kept low by holding only ONE
// warp-wide batch of row loads in registers (CH loads per lane) — every further fresh row of the expansion is
// already on its way into L2 via prefetch.global.L2, which costs no registers.  Longer rows need more registers
// per lane (query + one batch), so fewer CTAs are planned for them.
#ifndef FNB_CTAS_SHORT_ROWS
#define FNB_CTAS_SHORT_ROWS 6
#endif
#ifndef FNB_CTAS_TINY_ROWS
#define FNB_CTAS_TINY_ROWS FNB_CTAS_SHORT_ROWS
#endif
__host__ __device__ constexpr int fnb_min_ctas(int ch) {
  return ch <= 1 ? FNB_CTAS_TINY_ROWS : (ch <= 4 ? FNB_CTAS_SHORT_ROWS : (ch <= 8 ? 4 : 3));
}
// Dense plan for LARGE batches of short rows: 7 CTAs = 28 warps per SM (72 registers per thread).  Measured on 100k-query
// batches: +6 ... +11 % QPS over the 24-warp plan on every short-row configuration.  On a batch of a few waves (10k
// queries are 2.8 waves of 3552 warps) the last, partly filled wave costs more with more slots than the extra
// residency earns, so the host picks the plan per launch from the batch size (choose_dense_plan).
#define FNB_CTAS_DENSE 7
#ifndef FNB_U_CH4
#define FNB_U_CH4 1
#endif
#ifndef FNB_U_CH2
#define FNB_U_CH2 2
#endif
__host__ __device__ constexpr int fnb_batches_in_flight(int ch) { return ch >= 4 ? FNB_U_CH4 : (ch == 2 ? FNB_U_CH2 : 4 / ch); }
// Latency variant (few queries, one warp per CTA, registers are free): hold as many warp-wide load batches in
// registers as ~96 staging registers allow, up to all 32 rows of an expansion, so that one hop costs ONE HBM round
// trip for its rows instead of one per batch.
__host__ __device__ constexpr int fnb_batches_in_flight_lat(int g, int ch) {
  return (24 / ch < 1) ? 1 : ((24 / ch > g) ? g : 24 / ch);  // g = 32 / rows-per-instruction = batches per 32 rows
}
#define FNB_FULL 0xffffffffu
#define FNB_EMPTY 0xffffffffu

struct SearchParams {
  const uint4* __restrict__ vec;       // [N][stride]
  const uint32_t* __restrict__ adj;    // [N][M]
  const int32_t* __restrict__ labels;  // [N]
  const void* __restrict__ queries;    // [Q][dim] elements of the index data type, dense rows
  float* __restrict__ out_dist;        // [Q][K]
  int32_t* __restrict__ out_label;     // [Q][K]
  uint32_t* __restrict__ out_ndist;    // [Q] or null
  uint32_t* __restrict__ out_nhops;    // [Q] or null
  uint32_t* __restrict__ out_len;      // [Q] or null: results found per query (< K = short result)
  unsigned int* counter;               // persistent-warp work counter (zeroed before launch)
  unsigned long long* totals;          // [3]: sum n_dist, sum n_hops, #short results
  uint32_t N, M, dim, nchunks, stride;
  uint32_t Q, K, B, Bcap, Bpow2;  // Bpow2: largest power of two <= Bcap (fixed-step lower_bound)
  uint32_t nprobe, step;
  uint32_t vs_buckets;   // visited set: number of 16-byte buckets per warp
  uint32_t vs_shift;     // 32 - nbits, nbits = ceil(log2(N)): ids are hashed by a bijection of [0, 2^nbits)
  uint32_t vs_tag_mask;  // low bits of the hash kept as the tag (unique within a bucket)
  uint32_t vs_wide;      // 1 => 4 x 32-bit tags per bucket (huge N), 0 => 8 x 16-bit tags
  uint32_t lines_per_row;  // 128-byte lines per padded row (for the L2 prefetch)
  uint32_t warp_smem;  // bytes of shared memory per warp
  uint32_t query_vec_ok;  // 1 => query rows are 16-byte aligned and a whole number of chunks
  uint32_t query_pitch_chunks;  // != 0 => queries are rows of a padded vector array with this pitch (construction:
                                // the new nodes' own rows); chunks beyond the data are zero there
  uint32_t dense;  // 1 => the 28-warps-per-SM instantiation (large batches of short rows; never with lat)
  uint32_t pf2;  // CTA latency kernel: two-hop prefetch (search_cta_kernel.cuh, cta_rows); set by the host for batches of a few queries
  uint32_t lat;  // latency variants (few queries; query = blockIdx.x, grid-stride, `counter` unused): 1 => one warp per
                 // CTA (fnb_search_kernel<.., LAT>), 2 => one CTA of four warps per query (search_cta_kernel.cuh)
  // Self-cleaning launch state (fnb_search_device): != null => the last warp of the grid to finish copies `totals` to
  // `last_totals`, zeroes counter / totals / done for the slot's next user and reports `seq` to the host — no memset
  // between launches, so consecutive launches are adjacent in the stream and can overlap (pdl).
  unsigned int* done;
  unsigned long long* last_totals;  // [3]
  volatile unsigned int* done_seq;  // host-mapped: sequence number of the last launch that completed (a hint), or null
  uint32_t seq;
  uint32_t pdl;  // launch with programmatic stream serialization: the next launch of the stream may start its CTAs as
                 // this one's exit (the kernel triggers at its start and waits for its predecessor before it writes)
  // != null => the queries are still being copied into device memory while the kernel runs (fnb_search with pageable
  // caller buffers): *q_ready = number of leading queries already in place; a warp waits for it to pass its query index
  const unsigned int* q_ready;
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
// predicated form: when `pred` is false nothing is loaded and the destination keeps garbage that the
// caller never uses (saves the zero-fill moves a C++ `if` would need)
__device__ __forceinline__ uint4 ldg_stream_if(const uint4* p, bool pred) {
  uint4 r;
  asm volatile(
      "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %5, 0;\n\t"
      "@pp ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p), "r"((int)pred));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// Host-fed batches (fnb_search with pageable caller buffers): the queries are still being copied into device memory
// while the kernel runs; *q_ready is the number of leading queries already in place (device memory, written by the
// copy engine after each chunk, in stream order).
//
// Wait until *q_ready > qi.  Written as ONE block of PTX on purpose: as a C++ loop (any flavour: acquire load in inline
// asm + __nanosleep, out-of-line call) it made ptxas spill 72 ... 152 bytes in the traversal loop of the 512-byte-row
// instantiations and grow their code by half; as an opaque block it costs nothing.  Bounded (~10 s of 500 ns naps): if
// the feeding thread died the warp goes on with whatever is there and the call fails on the host side.
__device__ __forceinline__ void feed_wait(const unsigned int* p, uint32_t qi) {
  asm volatile(
      "{\n\t.reg .pred pw;\n\t.reg .u32 rv, rc;\n\t"
      "mov.u32 rc, 0;\n\t"
      "FNB_FEED_WAIT_%=:\n\t"
      "ld.acquire.gpu.global.u32 rv, [%0];\n\t"
      "add.u32 rc, rc, 1;\n\t"
      "setp.ls.u32 pw, rv, %1;\n\t"
      "setp.lt.and.u32 pw, rc, 20000000, pw;\n\t"
      "@pw nanosleep.u32 500;\n\t"
      "@pw bra FNB_FEED_WAIT_%=;\n\t}" ::"l"(p),
      "r"(qi)
      : "memory");
}

__device__ __forceinline__ uint32_t ord_f32(float f) {
  uint32_t u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float unord_f32(uint32_t o) {
  return __uint_as_float((o & 0x80000000u) ? (o ^ 0x80000000u) : ~o);
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t id) {
  return ((uint64_t)ord_f32(d) << 32) | ((uint64_t)id << 1);
}
__device__ __forceinline__ uint64_t shfl64(uint64_t v, int src) {
  uint32_t lo = __shfl_sync(FNB_FULL, (uint32_t)v, src);
  uint32_t hi = __shfl_sync(FNB_FULL, (uint32_t)(v >> 32), src);
  return ((uint64_t)hi << 32) | lo;
}

// ---- per-chunk accumulation -------------------------------------------------------------------
template <int DT, int METRIC>
struct Arith;

template <>
struct Arith<DT_F32, M_L2> {
  typedef float acc_t;
  static __device__ __forceinline__ void step(float& a, const uint4& q, const uint4& x) {
    float d;
    d = __fsub_rn(__uint_as_float(q.x), __uint_as_float(x.x)); a = __fmaf_rn(d, d, a);
    d = __fsub_rn(__uint_as_float(q.y), __uint_as_float(x.y)); a = __fmaf_rn(d, d, a);
    d = __fsub_rn(__uint_as_float(q.z), __uint_as_float(x.z)); a = __fmaf_rn(d, d, a);
    d = __fsub_rn(__uint_as_float(q.w), __uint_as_float(x.w)); a = __fmaf_rn(d, d, a);
  }
  static __device__ __forceinline__ float combine(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float finish(float a) { return a; }
};
template <>
struct Arith<DT_F32, M_IP> {
  typedef float acc_t;
  static __device__ __forceinline__ void step(float& a, const uint4& q, const uint4& x) {
    a = __fmaf_rn(__uint_as_float(q.x), __uint_as_float(x.x), a);
    a = __fmaf_rn(__uint_as_float(q.y), __uint_as_float(x.y), a);
    a = __fmaf_rn(__uint_as_float(q.z), __uint_as_float(x.z), a);
    a = __fmaf_rn(__uint_as_float(q.w), __uint_as_float(x.w), a);
  }
  static __device__ __forceinline__ float combine(float a, float b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ float finish(float a) { return __fsub_rn(1.0f, a); }
};
template <>
struct Arith<DT_U8, M_L2> {
  typedef unsigned acc_t;
  static __device__ __forceinline__ void step(unsigned& a, const uint4& q, const uint4& x) {
    unsigned d;
    d = __vabsdiffu4(q.x, x.x); a = __dp4a(d, d, a);
    d = __vabsdiffu4(q.y, x.y); a = __dp4a(d, d, a);
    d = __vabsdiffu4(q.z, x.z); a = __dp4a(d, d, a);
    d = __vabsdiffu4(q.w, x.w); a = __dp4a(d, d, a);
  }
  static __device__ __forceinline__ unsigned combine(unsigned a, unsigned b) { return a + b; }
  static __device__ __forceinline__ float finish(unsigned a) { return __uint2float_rn(a); }
};
template <>
struct Arith<DT_U8, M_IP> {
  typedef unsigned acc_t;
  static __device__ __forceinline__ void step(unsigned& a, const uint4& q, const uint4& x) {
    a = __dp4a(q.x, x.x, a); a = __dp4a(q.y, x.y, a); a = __dp4a(q.z, x.z, a); a = __dp4a(q.w, x.w, a);
  }
  static __device__ __forceinline__ unsigned combine(unsigned a, unsigned b) { return a + b; }
  static __device__ __forceinline__ float finish(unsigned a) { return __fsub_rn(1.0f, __uint2float_rn(a)); }
};
template <>
struct Arith<DT_I8, M_L2> {
  typedef unsigned acc_t;
  static __device__ __forceinline__ void step(unsigned& a, const uint4& q, const uint4& x) {
    unsigned d;  // per-byte |a-b| of signed bytes, 0..255, as unsigned bytes
    d = __vabsdiffs4(q.x, x.x); a = __dp4a(d, d, a);
    d = __vabsdiffs4(q.y, x.y); a = __dp4a(d, d, a);
    d = __vabsdiffs4(q.z, x.z); a = __dp4a(d, d, a);
    d = __vabsdiffs4(q.w, x.w); a = __dp4a(d, d, a);
  }
  static __device__ __forceinline__ unsigned combine(unsigned a, unsigned b) { return a + b; }
  static __device__ __forceinline__ float finish(unsigned a) { return __uint2float_rn(a); }
};
template <>
struct Arith<DT_I8, M_IP> {
  typedef int acc_t;
  static __device__ __forceinline__ void step(int& a, const uint4& q, const uint4& x) {
    a = __dp4a((int)q.x, (int)x.x, a); a = __dp4a((int)q.y, (int)x.y, a);
    a = __dp4a((int)q.z, (int)x.z, a); a = __dp4a((int)q.w, (int)x.w, a);
  }
  static __device__ __forceinline__ int combine(int a, int b) { return a + b; }
  static __device__ __forceinline__ float finish(int a) { return __fsub_rn(1.0f, __int2float_rn(a)); }
};

template <typename T>
__device__ __forceinline__ T shfl_xor_t(T v, int off) {
  return __shfl_xor_sync(FNB_FULL, v, off);
}

// ---- query load -----------------------------------------------------------------------------------
template <int DT>
__device__ __forceinline__ uint4 load_query_chunk(const SearchParams& p, uint32_t qi, uint32_t chunk) {
  uint4 r = make_uint4(0, 0, 0, 0);
  if (chunk >= p.nchunks) return r;
  if (p.query_pitch_chunks) return __ldg(reinterpret_cast<const uint4*>(p.queries) + (size_t)qi * p.query_pitch_chunks + chunk);
  if (p.query_vec_ok) {
    const uint4* row = reinterpret_cast<const uint4*>(p.queries) + (size_t)qi * p.nchunks;
    return __ldcg(row + chunk);  // L2 only: a query row is read once, and a host-fed batch changes under the kernel
  }
  if (DT == DT_F32) {
    const float* row = reinterpret_cast<const float*>(p.queries) + (size_t)qi * p.dim;
    uint32_t e = chunk * 4;
    if (e + 0 < p.dim) r.x = __float_as_uint(row[e + 0]);
    if (e + 1 < p.dim) r.y = __float_as_uint(row[e + 1]);
    if (e + 2 < p.dim) r.z = __float_as_uint(row[e + 2]);
    if (e + 3 < p.dim) r.w = __float_as_uint(row[e + 3]);
  } else {
    const uint8_t* row = reinterpret_cast<const uint8_t*>(p.queries) + (size_t)qi * p.dim;
    uint32_t w[4] = {0, 0, 0, 0};
#pragma unroll
    for (int b = 0; b < 16; b++) {
      uint32_t e = chunk * 16 + b;
      if (e < p.dim) w[b >> 2] |= (uint32_t)row[e] << (8 * (b & 3));
    }
    r = make_uint4(w[0], w[1], w[2], w[3]);
  }
  return r;
}

// ---- distances of up to 32 rows (one id per lane), cooperatively ------------------------------------
// Each valid lane gets back the distance of ITS row.  s_ids: 32-entry per-warp scratch.
// EXACT: the row is exactly G*CH chunks (e.g. D=128 f32 with G=8, CH=4): no per-chunk bounds test.
// prefetch: rows beyond the first register batch are pulled into L2 right away (prefetch.global.L2 costs no
// registers), so the later register batches pay an L2 hit instead of another HBM round trip each.
template <int DT, int METRIC, int G, int CH, bool EXACT, int UU = 0>
__device__ __forceinline__ float batch_distance(const SearchParams& p, const uint4 (&q)[CH], uint32_t my_id,
                                                bool valid, uint32_t* s_ids, int lane, bool prefetch) {
  typedef Arith<DT, METRIC> A;
  constexpr int RPI = 32 / G;                                   // rows per warp-wide load instruction
  constexpr int U = UU > 0 ? UU : fnb_batches_in_flight(CH);  // warp-wide load batches held in registers
  const unsigned mask = __ballot_sync(FNB_FULL, valid);
  const int n = __popc(mask);
  const int myrank = __popc(mask & ((1u << lane) - 1u));
  if (valid) s_ids[myrank] = my_id;
  __syncwarp();
  const int g = lane / G, pos = lane % G;
  if (prefetch && n > RPI * U) {
    if (p.lines_per_row <= (uint32_t)G) {  // one 128-byte line per lane of the group: a single predicated prefetch
      const bool mine_line = (uint32_t)pos < p.lines_per_row;
      for (int c0 = RPI * U; c0 < n; c0 += RPI) {
        const int c = c0 + g;
        if (mine_line && c < n) prefetch_l2(p.vec + (size_t)s_ids[c] * p.stride + pos * 8);
      }
    } else {
      for (int c0 = RPI * U; c0 < n; c0 += RPI) {
        const int c = c0 + g;
        if (c < n) {
          const uint4* row = p.vec + (size_t)s_ids[c] * p.stride;
          for (uint32_t line = (uint32_t)pos; line < p.lines_per_row; line += G) prefetch_l2(row + line * 8u);
        }
      }
    }
  }
  float mine = 0.f;
  for (int r0 = 0; r0 < n; r0 += RPI * U) {
    uint4 x[U][CH];
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (r0 + u * RPI < n) {  // warp-uniform
        const int c = r0 + u * RPI + g;
        const bool ok = c < n;
        const uint32_t rid = s_ids[ok ? c : 0];
        const uint4* row = p.vec + (size_t)rid * p.stride + pos;
#pragma unroll
        for (int k = 0; k < CH; k++) {
          if (EXACT)
            x[u][k] = ldg_stream_if(row + k * G, ok);
          else
            x[u][k] = ldg_stream_if(row + k * G, ok && (uint32_t)(k * G + pos) < p.nchunks);
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (r0 + u * RPI < n) {  // warp-uniform
        typename A::acc_t acc = 0;
#pragma unroll
        for (int k = 0; k < CH; k++) {
          if (EXACT || (uint32_t)(k * G + pos) < p.nchunks) A::step(acc, q[k], x[u][k]);
        }
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) acc = A::combine(acc, shfl_xor_t(acc, off));
        const float d = A::finish(acc);
        const int rel = myrank - (r0 + u * RPI);  // which group of this instruction handled my row
        const float v = __shfl_sync(FNB_FULL, d, (rel * G) & 31);
        if (rel >= 0 && rel < RPI) mine = v;
      }
    }
  }
  __syncwarp();
  return mine;
}

// ---- visited set ------------------------------------------------------------------------------------
// Replaces VisitedSet (util/VisitedSetPool.h:16-89: an N-byte table per thread).  Per warp, in shared memory:
// `vs_buckets` buckets of 16 bytes, each holding 8 16-bit tags (or 4 32-bit tags when N is so large that a tag
// does not fit 15 bits).  A node id is mapped by a bijection h of [0, 2^nbits) (odd multiplier), the bucket is
// floor(h * buckets / 2^nbits) and the tag is the low bits of h, chosen wide enough that (bucket, tag)
// identifies the id exactly: the set can never report an unvisited node as visited.  It CAN forget (a full
// bucket overwrites a slot; two lanes racing for one empty slot lose one insert).  Forgetting is harmless for the
// result: a forgotten node that is met again is re-evaluated and either rejected again (its distance is >= the
// current worst, which only decreases) or, if it still sits in the list, dropped by the duplicate test of the
// merge.  It only costs the extra row fetch; at the default sizing (>= 2 slots per expected visit) the measured
// re-evaluation rate is below 1 %.  One 128-bit load, straight-line code, no atomics, no probing loop.
__device__ __forceinline__ bool visited_test_and_set(uint32_t* tab, const SearchParams& p, uint32_t id) {
  const uint32_t h = (id * 0x9E3779B1u) << p.vs_shift;  // bijective hash, left-aligned
  const uint32_t bucket = __umulhi(h, p.vs_buckets);
  const uint32_t tag = (h >> p.vs_shift) & p.vs_tag_mask;
  uint4* bp = reinterpret_cast<uint4*>(tab) + bucket;
  uint4 w;  // one 128-bit shared-memory load, never cached in registers across calls
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
               : "r"((uint32_t)__cvta_generic_to_shared(bp)));
  if (!p.vs_wide) {
    const uint32_t t2 = tag | (tag << 16);
    // halfword-equals test: zero halfword of (w ^ t2)
    uint32_t x, hit = 0, e0, e1, e2, e3;
    x = w.x ^ t2; hit |= (x - 0x00010001u) & ~x;
    x = w.y ^ t2; hit |= (x - 0x00010001u) & ~x;
    x = w.z ^ t2; hit |= (x - 0x00010001u) & ~x;
    x = w.w ^ t2; hit |= (x - 0x00010001u) & ~x;
    if (hit & 0x80008000u) return false;
    // first empty (0xFFFF) halfword, else a victim chosen by the hash
    x = ~w.x; e0 = (x - 0x00010001u) & ~x & 0x80008000u;
    x = ~w.y; e1 = (x - 0x00010001u) & ~x & 0x80008000u;
    x = ~w.z; e2 = (x - 0x00010001u) & ~x & 0x80008000u;
    x = ~w.w; e3 = (x - 0x00010001u) & ~x & 0x80008000u;
    uint32_t slot;
    if (e0) slot = (e0 & 0x8000u) ? 0u : 1u;
    else if (e1) slot = (e1 & 0x8000u) ? 2u : 3u;
    else if (e2) slot = (e2 & 0x8000u) ? 4u : 5u;
    else if (e3) slot = (e3 & 0x8000u) ? 6u : 7u;
    else slot = (h >> 13) & 7u;
    reinterpret_cast<volatile uint16_t*>(bp)[slot] = (uint16_t)tag;
    return true;
  } else {
    if (w.x == tag || w.y == tag || w.z == tag || w.w == tag) return false;
    uint32_t slot;
    if (w.x == FNB_EMPTY) slot = 0;
    else if (w.y == FNB_EMPTY) slot = 1;
    else if (w.z == FNB_EMPTY) slot = 2;
    else if (w.w == FNB_EMPTY) slot = 3;
    else slot = (h >> 13) & 3u;
    reinterpret_cast<volatile uint32_t*>(bp)[slot] = tag;
    return true;
  }
}

// Read-only membership test (speculative prefetch of the latency variant): true iff `id` is in the set.
__device__ __forceinline__ bool visited_peek(uint32_t* tab, const SearchParams& p, uint32_t id) {
  const uint32_t h = (id * 0x9E3779B1u) << p.vs_shift;
  const uint32_t bucket = __umulhi(h, p.vs_buckets);
  const uint32_t tag = (h >> p.vs_shift) & p.vs_tag_mask;
  uint4* bp = reinterpret_cast<uint4*>(tab) + bucket;
  uint4 w;
  asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(w.x), "=r"(w.y), "=r"(w.z), "=r"(w.w)
               : "r"((uint32_t)__cvta_generic_to_shared(bp)));
  if (!p.vs_wide) {
    const uint32_t t2 = tag | (tag << 16);
    uint32_t x, hit = 0;
    x = w.x ^ t2; hit |= (x - 0x00010001u) & ~x;
    x = w.y ^ t2; hit |= (x - 0x00010001u) & ~x;
    x = w.z ^ t2; hit |= (x - 0x00010001u) & ~x;
    x = w.w ^ t2; hit |= (x - 0x00010001u) & ~x;
    return (hit & 0x80008000u) != 0u;
  }
  return w.x == tag || w.y == tag || w.z == tag || w.w == tag;
}

__device__ __forceinline__ void visited_clear(uint32_t* tab, uint32_t buckets, int lane) {
  uint4* t4 = reinterpret_cast<uint4*>(tab);
  const uint4 e = make_uint4(FNB_EMPTY, FNB_EMPTY, FNB_EMPTY, FNB_EMPTY);
  for (uint32_t i = lane; i < buckets; i += 32) t4[i] = e;
}

// ---------------------------------------------------------------------------------------------------
// Merge of the accepted entries of one expansion into the sorted list (in place, shared memory).
//   key/acc : this lane's candidate (valid iff acc)
// Every lane runs the same fixed-step lower_bound (no divergent loop).  A candidate whose (distance, id) is
// already in the list (a node the visited set forgot) or that equals another lane's candidate (a link listed
// twice) is dropped.  Each remaining lane's final slot is fin = lower_bound + rank among the accepted.  The
// final slots of the new entries form a bitmask T over [0, len + n_acc); old entries keep their order and fill
// the other slots, so destination t takes old entry t - popc(T[0..t)).  Chunks of 32 destinations are rewritten
// from the top down, which makes the in-place update safe (a source is never above its destination).
__device__ __forceinline__ void merge_accepted(volatile uint64_t* list, uint32_t& len, uint32_t& start,
                                               const uint32_t B, const uint32_t Bpow2, const uint64_t key, bool acc,
                                               const int lane) {
  // lower_bound(list[0..len), key) ignoring the expanded bit.  Binary search on the DISTANCE word alone: 32-bit loads
  // and compares, the load unconditional through a clamped index, so the loop body is straight-line code.  Only when a
  // lane lands on an entry with its own distance word (an exact distance tie, or a node the visited set forgot) is
  // the search repeated with the full (distance, id) keys.
  const uint32_t khi = (uint32_t)(key >> 32);
  const volatile uint32_t* l32 = reinterpret_cast<const volatile uint32_t*>(list);  // word 2i+1 = distance of entry i
  uint32_t lo = 0;
  for (uint32_t step = Bpow2; step; step >>= 1) {
    const uint32_t idx = lo + step;
    const uint32_t j = min(idx, len);  // len >= 1: the entry node is in the list before the first merge
    if ((l32[2u * j - 1u] < khi) & (idx <= len)) lo = idx;
  }
  const bool tie = acc && lo < len && l32[2u * lo + 1u] == khi;
  if (__any_sync(FNB_FULL, tie)) {
    lo = 0;
    for (uint32_t step = Bpow2; step; step >>= 1) {
      const uint32_t idx = lo + step;
      if (idx <= len && (list[idx - 1] & ~1ull) < key) lo = idx;
    }
    if (lo < len && (list[lo] & ~1ull) == key) acc = false;  // already in the list
  }
  unsigned am = __ballot_sync(FNB_FULL, acc);
  uint32_t rank = 0;
  if (am & (am - 1u)) {  // two or more accepted: rank each among the others
    // Candidates of one expansion rarely share a distance word; then 32-bit compares rank them and no link can be
    // listed twice among them.  match.any groups the accepted lanes by distance word in one instruction.
    const unsigned peers = __match_any_sync(FNB_FULL, acc ? khi : ~(uint32_t)lane) & am;
    if (!__any_sync(FNB_FULL, acc && peers != (1u << lane))) {
      for (unsigned m = am; m; m &= m - 1) rank += (__shfl_sync(FNB_FULL, khi, __ffs(m) - 1) < khi) ? 1u : 0u;
    } else {
      bool dup = false;
      for (unsigned m = am; m; m &= m - 1) {
        const int s = __ffs(m) - 1;
        const uint64_t kj = shfl64(key, s);
        rank += (kj < key) ? 1u : 0u;
        dup |= (kj == key) && (s < lane);
      }
      if (__any_sync(FNB_FULL, acc && dup)) {  // a link listed twice: keep the lowest lane, redo the ranks
        acc = acc && !dup;
        am = __ballot_sync(FNB_FULL, acc);
        rank = 0;
        for (unsigned m = am; m; m &= m - 1) rank += (shfl64(key, __ffs(m) - 1) < key) ? 1u : 0u;
      }
    }
  }
  if (!am) return;
  const uint32_t n_acc = (uint32_t)__popc(am);
  const uint32_t fin = acc ? lo + rank : 0xffffffffu;
  const uint32_t fin_min = __reduce_min_sync(FNB_FULL, fin);
  const uint32_t new_len = min(B, len + n_acc);
  if (fin_min < new_len) {
    const int c_lo = (int)(fin_min >> 5), c_hi = (int)((new_len - 1) >> 5);
    uint32_t below = (uint32_t)__popc(__ballot_sync(FNB_FULL, acc && (fin >> 5) < (uint32_t)c_hi));
    for (int c = c_hi; c >= c_lo; c--) {
      const uint32_t T = __reduce_or_sync(FNB_FULL, (acc && (fin >> 5) == (uint32_t)c) ? (1u << (fin & 31)) : 0u);
      const uint32_t t = (uint32_t)c * 32 + lane;
      const bool is_new = (T >> lane) & 1u;
      const uint32_t src = t - below - (uint32_t)__popc(T & ((1u << lane) - 1u));
      const bool mv = t < new_len && !is_new && src != t;
      const uint64_t y = mv ? list[src] : 0ull;
      __syncwarp();
      if (mv) list[t] = y;
      if (c > c_lo) below -= (uint32_t)__popc(__ballot_sync(FNB_FULL, acc && (fin >> 5) == (uint32_t)(c - 1)));
      __syncwarp();
    }
    if (acc && fin < B) list[fin] = key;
    __syncwarp();
  }
  len = new_len;
  start = min(start, fin_min);
}

// ---------------------------------------------------------------------------------------------------
// LAT = latency variant, chosen by the host when the batch is too small to fill the machine (search_single, small
// batches): the same algorithm and arithmetic, bit for bit, but tuned for the length of ONE query's dependency chain
// instead of for resident warps —
//   * one warp per CTA, no register cap: every row of an expansion is loaded in one go (fnb_batches_in_flight_lat),
//     so a hop pays one HBM round trip for its rows rather than one per register batch;
//   * the adjacency row of every fresh neighbour is prefetched into L2 together with its vector, so whichever of
//     them is expanded later finds its links in L2;
//   * speculation that cannot change the result: while the current node's rows are in flight, the links of the
//     runner-up candidate (the most likely next expansion) are read, filtered through the visited set READ-ONLY, and
//     their vector rows prefetched into L2.
template <int DT, int METRIC, int G, int CH, bool EXACT, bool LAT, int OCC = 0>
__global__ void __launch_bounds__(LAT ? 32 : FNB_WARPS_PER_CTA * 32, LAT ? 1 : (OCC > 0 ? OCC : fnb_min_ctas(CH)))
    fnb_search_kernel(const SearchParams p) {
  extern __shared__ __align__(16) unsigned char fnb_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  unsigned char* wbase = fnb_smem + (size_t)warp * p.warp_smem;
  volatile uint64_t* list = reinterpret_cast<volatile uint64_t*>(wbase);
  uint32_t* tab = reinterpret_cast<uint32_t*>(wbase + (size_t)p.Bcap * 8);
  uint32_t* s_ids = tab + p.vs_buckets * 4;
  const int pos = lane % G;
  constexpr int UL = LAT ? fnb_batches_in_flight_lat(G, CH) : 0;  // 0: the throughput variant's default
  uint32_t qi_next = blockIdx.x;
  // Programmatic dependent launch: let the NEXT launch of this stream place its CTAs as soon as ours retire (the tail
  // of a batch of a few "rounds" of queries leaves SMs idle otherwise).  Nothing here depends on the previous launch
  // except the order of writes to the caller's output buffers, which griddepcontrol.wait restores below.  Both are
  // no-ops for a launch without the attribute.
  asm volatile("griddepcontrol.launch_dependents;");

  for (;;) {
    uint32_t qi = 0;
    if (LAT) {
      qi = qi_next;
      qi_next += gridDim.x;
    } else {
      if (lane == 0) qi = atomicAdd(p.counter, 1u);
      qi = __shfl_sync(FNB_FULL, qi, 0);
    }
    if (qi >= p.Q) break;
    if (!LAT && p.q_ready) feed_wait(p.q_ready, qi);  // host-fed batch: query qi may still be on its way

    uint4 q[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) q[k] = load_query_chunk<DT>(p, qi, (uint32_t)(k * G + pos));

    visited_clear(tab, p.vs_buckets, lane);
    __syncwarp();

    uint32_t ndist = 0, nhops = 0, len = 0;

    if (p.N > 0) {
      // ---- entry selection: strided probes, first strict minimum wins (Index.h:845-870) ----
      uint64_t best = ~0ull;  // (ordered distance, probe index)
      for (uint32_t base = 0; base < p.nprobe; base += 32) {
        const uint32_t pi = base + lane;
        const bool valid = pi < p.nprobe;
        const float d = batch_distance<DT, METRIC, G, CH, EXACT, UL>(p, q, pi * p.step, valid, s_ids, lane, false);
        if (valid) {
          const uint64_t k = ((uint64_t)ord_f32(d) << 32) | pi;
          best = k < best ? k : best;
        }
      }
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        const uint64_t o = ((uint64_t)__shfl_xor_sync(FNB_FULL, (uint32_t)(best >> 32), off) << 32) |
                           __shfl_xor_sync(FNB_FULL, (uint32_t)best, off);
        best = o < best ? o : best;
      }
      ndist = p.nprobe;
      const uint32_t entry = (uint32_t)best * p.step;
      if (lane == 0) {
        list[0] = (best & 0xffffffff00000000ull) | ((uint64_t)entry << 1);
        visited_test_and_set(tab, p, entry);
      }
      len = 1;
      __syncwarp();

      uint32_t start = 0;
      // ---- main loop (Index.h:627-658) ----
      for (;;) {
        uint32_t cur = FNB_EMPTY;
        uint32_t spec = FNB_EMPTY;  // LAT: the runner-up candidate
        for (uint32_t base = start & ~31u; base < len; base += 32) {
          const uint32_t i = base + lane;
          const uint64_t e = (i < len) ? list[i] : 1ull;
          const unsigned b = __ballot_sync(FNB_FULL, !(e & 1ull));
          if (b) {
            const int src = __ffs(b) - 1;
            cur = __shfl_sync(FNB_FULL, (uint32_t)e, src) >> 1;
            if (lane == src) list[i] = e | 1ull;
            start = base + (uint32_t)src;
            const unsigned b2 = b & (b - 1);
            if (b2) {  // runner-up: most likely the next node to be expanded
              const uint32_t id2 = __shfl_sync(FNB_FULL, (uint32_t)e, __ffs(b2) - 1) >> 1;
              if (LAT) spec = id2;
              else if (lane == 0) prefetch_l2(p.adj + (size_t)id2 * p.M);
            }
            break;
          }
        }
        if (cur == FNB_EMPTY) break;
        __syncwarp();
        nhops++;

        // LAT: the runner-up's first 32 links travel together with the current node's (its row was prefetched
        // into L2 when it was a fresh neighbour)
        uint32_t nb2 = spec;
        if (LAT && spec != FNB_EMPTY && p.lines_per_row <= 8u && (uint32_t)lane < p.M)
          nb2 = __ldg(p.adj + (size_t)spec * p.M + lane);

        for (uint32_t l0 = 0; l0 < p.M; l0 += 32) {
          uint32_t nb = cur;
          if (l0 + lane < p.M) nb = __ldg(p.adj + (size_t)cur * p.M + l0 + lane);
          // unused link slots are self-loops (Index.h:270): skip them without touching the visited set
          const bool fresh = (nb != cur) && visited_test_and_set(tab, p, nb);
          const unsigned fm = __ballot_sync(FNB_FULL, fresh);
          if (LAT) {
            if (fresh) {
              const uint32_t* arow = p.adj + (size_t)nb * p.M;
              for (uint32_t o = 0; o < p.M; o += 32) prefetch_l2(arow + o);
            }
            if (l0 == 0 && nb2 != spec && !visited_peek(tab, p, nb2)) {  // speculative: rows of the runner-up's links
              const uint4* row = p.vec + (size_t)nb2 * p.stride;
              for (uint32_t line = 0; line < p.lines_per_row; line++) prefetch_l2(row + line * 8u);
            }
          }
          if (!fm) continue;
          ndist += (uint32_t)__popc(fm);
          const bool full = len >= p.B;
          const uint32_t worst_hi = (uint32_t)(list[len - 1] >> 32);

          const float d = batch_distance<DT, METRIC, G, CH, EXACT, UL>(p, q, nb, fresh, s_ids, lane, !LAT || UL * (32 / G) < 32);
          const uint64_t key = make_key(d, nb);
          const bool acc = fresh && (!full || (uint32_t)(key >> 32) < worst_hi);
          if (!__any_sync(FNB_FULL, acc)) continue;
          merge_accepted(list, len, start, p.B, p.Bpow2, key, acc, lane);
        }
      }
    }

    // ---- output: ascending distance, label field of the node (Index.h:393-406) ----
    asm volatile("griddepcontrol.wait;" ::: "memory");  // the previous launch of the stream has finished writing
    for (uint32_t i = lane; i < p.K; i += 32) {
      float od = __int_as_float(0x7f800000);
      int32_t ol = -1;
      if (i < len) {
        const uint64_t e = list[i];
        od = unord_f32((uint32_t)(e >> 32));
        ol = p.labels ? __ldg(p.labels + ((uint32_t)e >> 1)) : (int32_t)((uint32_t)e >> 1);  // null: node ids
      }
      p.out_dist[(size_t)qi * p.K + i] = od;
      p.out_label[(size_t)qi * p.K + i] = ol;
    }
    if (lane == 0) {
      if (p.out_ndist) p.out_ndist[qi] = ndist;
      if (p.out_nhops) p.out_nhops[qi] = nhops;
      if (p.out_len) p.out_len[qi] = len < p.K ? len : p.K;
      if (p.totals) {
        atomicAdd(p.totals + 0, (unsigned long long)ndist);
        atomicAdd(p.totals + 1, (unsigned long long)nhops);
        if (len < p.K) atomicAdd(p.totals + 2, 1ull);
      }
    }
    __syncwarp();
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");  // also by warps that found no work: launches complete in stream order
  if (p.done && lane == 0) {
    __threadfence_system();  // this warp's results (possibly in pinned host memory) before its count
    const unsigned int n_warps = gridDim.x * (LAT ? 1u : (unsigned int)FNB_WARPS_PER_CTA);
    if (atomicAdd(p.done, 1u) == n_warps - 1u) {  // the last warp of the grid
      __threadfence();
      if (p.totals) {
        p.last_totals[0] = atomicExch(p.totals + 0, 0ull);
        p.last_totals[1] = atomicExch(p.totals + 1, 0ull);
        p.last_totals[2] = atomicExch(p.totals + 2, 0ull);
      }
      if (p.counter) *p.counter = 0u;
      *p.done = 0u;
      __threadfence_system();
      if (p.done_seq) *p.done_seq = p.seq;  // the host may be polling this word instead of synchronising the stream
    }
  }
}

// ---- host-side launcher -----------------------------------------------------------------------------
// Shared memory per warp = list (Bcap*8) + visited set (buckets*16) + 128 B scratch.  The visited set takes what
// the planned occupancy (min_ctas CTAs of 4 warps per SM) leaves of the 227 KB, up to ~48 slots per list entry
// (~20 evaluations per unit of ef were measured on the BASELINE configs, so <= 45 % load).
inline void size_visited(SearchParams& p, int force_buckets, int min_ctas, int queries_per_cta = FNB_WARPS_PER_CTA) {
  const uint32_t list_bytes = p.Bcap * 8u + 128u;
  uint32_t nbits = 1;
  while ((1ull << nbits) < (uint64_t)p.N && nbits < 31) nbits++;
  // 16-bit tags are possible when a bucket's share of the hash range fits 15 bits
  auto tag_bits_for = [&](uint32_t buckets) {
    uint64_t span = ((1ull << nbits) + buckets - 1) / buckets;
    uint32_t tb = 0;
    while ((1ull << tb) < span) tb++;
    return tb;
  };
  uint32_t buckets;
  if (force_buckets > 0) {
    buckets = (uint32_t)force_buckets;
  } else {
    // Full size: ~48 slots per list entry.  Resident warps are worth far more than a roomy set: the traversal is
    // latency-bound per warp, and a forgotten node only costs one extra row fetch (measured at ef=1000 on
    // 1M x 128: 6 CTAs/SM with a set so small that n_dist grows 64 % is still 1.9x faster than 2 CTAs/SM with a
    // full-size set).  So the planned CTA count is lowered only when the LIST no longer fits beside a minimal
    // set of one slot per list entry.
    const uint32_t want = (p.B * 48u + 7u) / 8u;
    const uint32_t floor_b = p.B / 8u > 16u ? p.B / 8u : 16u;
    buckets = 16u;
    for (int c = min_ctas; c >= 1; c--) {
      const uint32_t budget = (227u * 1024u - (uint32_t)c * 1024u) / ((uint32_t)c * (uint32_t)queries_per_cta);  // per query
      const uint32_t room = budget > list_bytes ? (budget - list_bytes) / 16u : 0u;
      buckets = want < room ? want : room;
      if (buckets >= floor_b || c == 1) break;
    }
    if (buckets < 16u) buckets = 16u;
  }
  uint32_t tb = tag_bits_for(buckets);
  p.vs_wide = tb > 15u ? 1u : 0u;
  if (tb > 31u) tb = 31u;
  p.vs_buckets = buckets;
  p.vs_shift = 32u - nbits;
  p.vs_tag_mask = (uint32_t)((1ull << tb) - 1ull);
  p.warp_smem = (list_bytes + buckets * 16u + 15u) & ~15u;
}

// The occupancy query and the shared-memory opt-in are per (device, kernel, shared-memory size): cached, because
// the latency variant is launched once per query by search_single and every host microsecond counts there.
struct LaunchCache {
  size_t smem = ~(size_t)0;
  int ctas_per_sm = 0;
};

template <typename Kern>
static inline cudaError_t launch_maybe_pdl(Kern kern, unsigned grid, unsigned block, size_t smem, cudaStream_t stream,
                                           const SearchParams& p) {
  if (!p.pdl) {
    kern<<<grid, block, smem, stream>>>(p);
    return cudaGetLastError();
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

template <typename Kern>
static inline cudaError_t plan_launch(Kern kern, int threads, size_t smem, LaunchCache* cache, int* ctas_per_sm) {
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return e;
  LaunchCache& c = cache[dev & 15];
  if (c.smem != smem) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, smem);
    if (e != cudaSuccess) return e;
    c.ctas_per_sm = n;
    c.smem = smem;
  }
  *ctas_per_sm = c.ctas_per_sm;
  return c.ctas_per_sm < 1 ? cudaErrorLaunchOutOfResources : cudaSuccess;
}

template <int DT, int METRIC, int G, int CH>
cudaError_t launch_search(const SearchParams& p, int num_sms, cudaStream_t stream) {
  const bool exact = p.nchunks == (uint32_t)(G * CH);
  static LaunchCache cache[6][16];  // [exact][lat | dense] x device; callers serialise launches per index, races only redo the query
  int ctas_per_sm = 0;
  if (p.lat) {
    auto kern = exact ? fnb_search_kernel<DT, METRIC, G, CH, true, true> : fnb_search_kernel<DT, METRIC, G, CH, false, true>;
    const size_t smem = (size_t)p.warp_smem;
    cudaError_t e = plan_launch(kern, 32, smem, cache[exact ? 3 : 2], &ctas_per_sm);
    if (e != cudaSuccess) return e;
    long long grid = (long long)num_sms * ctas_per_sm;
    if (grid > (long long)p.Q) grid = p.Q;
    if (grid < 1) grid = 1;
    return launch_maybe_pdl(kern, (unsigned)grid, 32u, smem, stream, p);
  }
  auto kern = exact ? fnb_search_kernel<DT, METRIC, G, CH, true, false> : fnb_search_kernel<DT, METRIC, G, CH, false, false>;
  LaunchCache* lc = cache[exact ? 1 : 0];
  if constexpr (CH <= 4 && G <= 8) {
    if (p.dense) {
      kern = exact ? fnb_search_kernel<DT, METRIC, G, CH, true, false, FNB_CTAS_DENSE>
                   : fnb_search_kernel<DT, METRIC, G, CH, false, false, FNB_CTAS_DENSE>;
      lc = cache[exact ? 5 : 4];
    }
  }
  const size_t smem = (size_t)p.warp_smem * FNB_WARPS_PER_CTA;
  cudaError_t e = plan_launch(kern, FNB_WARPS_PER_CTA * 32, smem, lc, &ctas_per_sm);
  if (e != cudaSuccess) return e;
  long long grid = (long long)num_sms * ctas_per_sm;
  const long long need = ((long long)p.Q + FNB_WARPS_PER_CTA - 1) / FNB_WARPS_PER_CTA;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  return launch_maybe_pdl(kern, (unsigned)grid, (unsigned)FNB_WARPS_PER_CTA * 32u, smem, stream, p);
}

// Host rule for SearchParams::dense.  Both plans keep persistent warps pulling queries, so a launch lasts about
// waves(c) = Q / slots(c) "rounds" of slots(c) / throughput(c) each, and its last round is only partly filled; the
// measured cost of that tail is about half of what idle slots would suggest (the busy warps of a thin round run faster).
// throughput(7 CTAs) = 1.08 x throughput(6 CTAs) on full waves (tools/ab_probe.py, 100k-query batches).  The dense
// plan is taken when it wins by 2 % or more: never at 10k queries (2.8 vs 2.4 waves), from ~16k queries on always.
// FNB_DENSE=0 / 1 forces the choice (tests, experiments).
inline uint32_t choose_dense_plan(int64_t Q, int num_sms, int lanes_per_row, int chunks_per_lane, uint32_t B) {
  static const int forced = [] {
    const char* e = getenv("FNB_DENSE");
    return e ? atoi(e) : -1;
  }();
  // rows above 512 B keep their own plans (more registers per lane); long lists need the shared memory
  if (lanes_per_row > 8 || chunks_per_lane > 4 || B > 256u) return 0u;
  if (forced >= 0) return forced ? 1u : 0u;
  auto cost = [&](int ctas, double thr) {
    const double slots = (double)num_sms * ctas * FNB_WARPS_PER_CTA;
    const double waves = (double)Q / slots;
    double whole = (double)(long long)waves;
    if (whole < waves) whole += 1.0;
    return (waves + 0.5 * (whole - waves)) * slots / thr;
  };
  return cost(FNB_CTAS_DENSE, 1.08) < 0.98 * cost(fnb_min_ctas(chunks_per_lane), 1.0) ? 1u : 0u;
}

// Host rule for SearchParams::lat: a latency variant when every query of the batch can have an SM quarter of its own
// (<= 4 CTAs per SM): the CTA-per-query kernel (2).  FNB_LAT=0 / 1 / 2 forces the throughput kernel / the one-warp
// latency variant / the CTA kernel (tests, experiments).
inline uint32_t choose_latency_variant(int64_t Q, int num_sms) {
  static const int forced = [] {
    const char* e = getenv("FNB_LAT");
    return e ? atoi(e) : -1;
  }();
  if (forced >= 0) return forced > 2 ? 2u : (uint32_t)forced;
  return Q <= 4ll * num_sms ? 2u : 0u;
}

}  // namespace fnb
