// Warp-per-query greedy beam search over the flat navigable-small-world graph (sm_100a).
//
// Replaces, for a whole batch of queries in one launch,
//   Index::search / initializeSearch / beamSearch / processCandidateNode
//       (include/flatnav/index/Index.h:387-409, 845-870, 606-659, 661-707 of the reference)
//   the two std::priority_queue heaps (Index.h:47-53, 623-624)  -> one sorted list in shared memory
//   VisitedSet / VisitedSetPool (util/VisitedSetPool.h:16-197)   -> per-warp open-addressing hash in
//                                                                  shared memory with a bounded reset
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

#include "fnb_layout.h"

namespace fnb {

enum { DT_F32 = 0, DT_U8 = 1, DT_I8 = 2 };
enum { M_L2 = 0, M_IP = 1 };

#define FNB_WARPS_PER_CTA 4
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
  unsigned int* counter;               // persistent-warp work counter (zeroed before launch)
  unsigned long long* totals;          // [3]: sum n_dist, sum n_hops, #short results
  uint32_t N, M, dim, nchunks, stride;
  uint32_t Q, K, B, Bcap;
  uint32_t nprobe, step;
  uint32_t hash_bits, hash_limit;
  uint32_t warp_smem;  // bytes of shared memory per warp
  uint32_t query_vec_ok;  // 1 => query rows are 16-byte aligned and a whole number of chunks
};

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

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
  if (p.query_vec_ok) {
    const uint4* row = reinterpret_cast<const uint4*>(p.queries) + (size_t)qi * p.nchunks;
    return __ldg(row + chunk);
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
template <int DT, int METRIC, int G, int CH>
__device__ __forceinline__ float batch_distance(const SearchParams& p, const uint4 (&q)[CH], uint32_t my_id,
                                                bool valid, uint32_t* s_ids, int lane) {
  typedef Arith<DT, METRIC> A;
  constexpr int RPI = 32 / G;                                   // rows per warp-wide load instruction
  constexpr int U = (20 / CH) < 1 ? 1 : ((20 / CH) > 8 ? 8 : (20 / CH));  // load instructions in flight / CH
  const unsigned mask = __ballot_sync(FNB_FULL, valid);
  const int n = __popc(mask);
  const int myrank = __popc(mask & ((1u << lane) - 1u));
  if (valid) s_ids[myrank] = my_id;
  __syncwarp();
  const int g = lane / G, pos = lane % G;
  float mine = 0.f;
  for (int r0 = 0; r0 < n; r0 += RPI * U) {
    uint4 x[U][CH];
#pragma unroll
    for (int u = 0; u < U; u++) {
      const int c = r0 + u * RPI + g;
      const bool ok = c < n;
      const uint32_t rid = s_ids[ok ? c : 0];
      const uint4* row = p.vec + (size_t)rid * p.stride;
#pragma unroll
      for (int k = 0; k < CH; k++) {
        const uint32_t chunk = (uint32_t)(k * G + pos);
        if (ok && chunk < p.nchunks)
          x[u][k] = ldg_stream(row + chunk);
        else
          x[u][k] = make_uint4(0, 0, 0, 0);
      }
    }
#pragma unroll
    for (int u = 0; u < U; u++) {
      if (r0 + u * RPI < n) {  // warp-uniform
        typename A::acc_t acc = 0;
#pragma unroll
        for (int k = 0; k < CH; k++) A::step(acc, q[k], x[u][k]);
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) acc = A::combine(acc, shfl_xor_t(acc, off));
        const float d = A::finish(acc);
        const int cbase = r0 + u * RPI;
        const int rel = myrank - cbase;  // which group of this instruction handled my row (if 0 <= rel < RPI)
        const float v = __shfl_sync(FNB_FULL, d, (rel * G) & 31);
        if (valid && rel >= 0 && rel < RPI) mine = v;
      }
    }
  }
  __syncwarp();
  return mine;
}

// ---- visited hash ----------------------------------------------------------------------------------
__device__ __forceinline__ bool hash_test_and_set(uint32_t* tab, uint32_t bits, uint32_t id) {
  const uint32_t cap_mask = (1u << bits) - 1u;
  uint32_t slot = (id * 0x9E3779B1u) >> (32 - bits);
  for (;;) {
    uint32_t v = reinterpret_cast<volatile uint32_t*>(tab)[slot];
    if (v == id) return false;
    if (v == FNB_EMPTY) {
      uint32_t old = atomicCAS(&tab[slot], FNB_EMPTY, id);
      if (old == FNB_EMPTY) return true;
      if (old == id) return false;
    }
    slot = (slot + 1) & cap_mask;
  }
}

__device__ __forceinline__ void hash_clear(uint32_t* tab, uint32_t bits, int lane) {
  uint4* t4 = reinterpret_cast<uint4*>(tab);
  const uint32_t n4 = (1u << bits) / 4;
  const uint4 e = make_uint4(FNB_EMPTY, FNB_EMPTY, FNB_EMPTY, FNB_EMPTY);
  for (uint32_t i = lane; i < n4; i += 32) t4[i] = e;
}

// ---------------------------------------------------------------------------------------------------
template <int DT, int METRIC, int G, int CH>
__global__ void __launch_bounds__(FNB_WARPS_PER_CTA * 32, 3) fnb_search_kernel(const SearchParams p) {
  extern __shared__ __align__(16) unsigned char fnb_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  unsigned char* wbase = fnb_smem + (size_t)warp * p.warp_smem;
  uint64_t* list = reinterpret_cast<uint64_t*>(wbase);
  uint32_t* tab = reinterpret_cast<uint32_t*>(wbase + (size_t)p.Bcap * 8);
  uint32_t* s_ids = tab + (1u << p.hash_bits);
  const int pos = lane % G;

  for (;;) {
    uint32_t qi = 0;
    if (lane == 0) qi = atomicAdd(p.counter, 1u);
    qi = __shfl_sync(FNB_FULL, qi, 0);
    if (qi >= p.Q) break;

    uint4 q[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) q[k] = load_query_chunk<DT>(p, qi, (uint32_t)(k * G + pos));

    hash_clear(tab, p.hash_bits, lane);
    __syncwarp();

    uint32_t ndist = 0, nhops = 0, len = 0;

    if (p.N > 0) {
      // ---- entry selection: strided probes, first strict minimum wins (Index.h:845-870) ----
      uint64_t best = ~0ull;  // (ordered distance, probe index)
      for (uint32_t base = 0; base < p.nprobe; base += 32) {
        const uint32_t pi = base + lane;
        const bool valid = pi < p.nprobe;
        const float d = batch_distance<DT, METRIC, G, CH>(p, q, pi * p.step, valid, s_ids, lane);
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
        hash_test_and_set(tab, p.hash_bits, entry);
      }
      len = 1;
      __syncwarp();

      uint32_t start = 0, n_ins = 1;
      // ---- main loop (Index.h:627-658) ----
      for (;;) {
        uint32_t cur = FNB_EMPTY;
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
              if (lane == 0) prefetch_l2(p.adj + (size_t)id2 * p.M);
            }
            break;
          }
        }
        if (cur == FNB_EMPTY) break;
        __syncwarp();
        nhops++;

        for (uint32_t l0 = 0; l0 < p.M; l0 += 32) {
          if (n_ins + 32 > p.hash_limit) {
            // bounded reset: forget everything except the current list.  A forgotten node that is met
            // again is re-evaluated and rejected again (its distance is >= the current worst), so results
            // do not change; only n_dist grows.
            hash_clear(tab, p.hash_bits, lane);
            __syncwarp();
            for (uint32_t i = lane; i < len; i += 32) hash_test_and_set(tab, p.hash_bits, (uint32_t)list[i] >> 1);
            __syncwarp();
            n_ins = len;
          }
          uint32_t nb = 0;
          bool fresh = false;
          if (l0 + lane < p.M) {
            nb = __ldg(p.adj + (size_t)cur * p.M + l0 + lane);
            fresh = hash_test_and_set(tab, p.hash_bits, nb);
          }
          const unsigned fm = __ballot_sync(FNB_FULL, fresh);
          if (!fm) continue;
          const uint32_t nf = (uint32_t)__popc(fm);
          n_ins += nf;
          ndist += nf;
          const bool full = len >= p.B;
          const uint32_t worst_hi = (uint32_t)(list[len - 1] >> 32);

          const float d = batch_distance<DT, METRIC, G, CH>(p, q, nb, fresh, s_ids, lane);
          const uint64_t key = make_key(d, nb);
          const bool acc = fresh && (!full || (uint32_t)(key >> 32) < worst_hi);
          const unsigned am = __ballot_sync(FNB_FULL, acc);
          if (!am) continue;
          const uint32_t n_acc = (uint32_t)__popc(am);

          // rank among the accepted, position in the list
          uint32_t rank = 0;
          for (unsigned m = am; m; m &= m - 1) {
            const uint64_t kj = shfl64(key, __ffs(m) - 1);
            rank += (kj < key) ? 1u : 0u;
          }
          uint32_t ipos = 0xffffffffu;
          if (acc) {
            uint32_t lo = 0, hi = len;
            while (lo < hi) {
              const uint32_t mid = (lo + hi) >> 1;
              if (list[mid] < key) lo = mid + 1; else hi = mid;
            }
            ipos = lo;
          }
          const uint32_t fin = acc ? ipos + rank : 0xffffffffu;
          const uint32_t pos_min = __reduce_min_sync(FNB_FULL, ipos);
          const uint32_t fin_min = __reduce_min_sync(FNB_FULL, fin);
          __syncwarp();
          // shift the tail of the list, highest chunk first (in place)
          if (pos_min < len) {
            for (int c = (int)((len - 1) >> 5); c >= (int)(pos_min >> 5); c--) {
              const uint32_t i = (uint32_t)c * 32 + lane;
              const bool have = i < len && i >= pos_min;
              const uint64_t y = have ? list[i] : 0ull;
              uint32_t sh = 0;
              for (unsigned m = am; m; m &= m - 1) {
                const uint32_t pj = __shfl_sync(FNB_FULL, ipos, __ffs(m) - 1);
                sh += (pj <= i) ? 1u : 0u;
              }
              __syncwarp();
              if (have && sh > 0 && i + sh < p.B) list[i + sh] = y;
              __syncwarp();
            }
          }
          if (acc && fin < p.B) list[fin] = key;
          __syncwarp();
          len = min(p.B, len + n_acc);
          start = min(start, fin_min);
        }
      }
    }

    // ---- output: ascending distance, label field of the node (Index.h:393-406) ----
    for (uint32_t i = lane; i < p.K; i += 32) {
      float od = __int_as_float(0x7f800000);
      int32_t ol = -1;
      if (i < len) {
        const uint64_t e = list[i];
        od = unord_f32((uint32_t)(e >> 32));
        ol = __ldg(p.labels + ((uint32_t)e >> 1));
      }
      p.out_dist[(size_t)qi * p.K + i] = od;
      p.out_label[(size_t)qi * p.K + i] = ol;
    }
    if (lane == 0) {
      if (p.out_ndist) p.out_ndist[qi] = ndist;
      if (p.out_nhops) p.out_nhops[qi] = nhops;
      atomicAdd(p.totals + 0, (unsigned long long)ndist);
      atomicAdd(p.totals + 1, (unsigned long long)nhops);
      if (len < p.K) atomicAdd(p.totals + 2, 1ull);
    }
    __syncwarp();
  }
}

// ---- host-side launcher -----------------------------------------------------------------------------
template <int DT, int METRIC, int G, int CH>
cudaError_t launch_search(const SearchParams& p, int num_sms, cudaStream_t stream) {
  auto kern = fnb_search_kernel<DT, METRIC, G, CH>;
  const size_t smem = (size_t)p.warp_smem * FNB_WARPS_PER_CTA;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int ctas_per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, FNB_WARPS_PER_CTA * 32, smem);
  if (e != cudaSuccess) return e;
  if (ctas_per_sm < 1) return cudaErrorLaunchOutOfResources;
  long long grid = (long long)num_sms * ctas_per_sm;
  const long long need = ((long long)p.Q + FNB_WARPS_PER_CTA - 1) / FNB_WARPS_PER_CTA;
  if (grid > need) grid = need;
  if (grid < 1) grid = 1;
  kern<<<(unsigned)grid, FNB_WARPS_PER_CTA * 32, smem, stream>>>(p);
  return cudaGetLastError();
}

}  // namespace fnb
