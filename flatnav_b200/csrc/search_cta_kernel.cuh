// CTA-per-query latency kernel (sm_100a): the same traversal as fnb_search_kernel, bit for bit, laid out for the length
// of ONE query's dependency chain instead of for resident queries.  Chosen by the host for batches that cannot fill
// the machine anyway (search_single, Index::search, a handful of queries): see choose_latency_variant.
//
// A lone warp walking the graph is bound by instruction latency, not by HBM: one hop of the single-warp latency variant
// costs ~890 dependent-ish instructions at ~7 cycles each, of which only ~1.8 are memory waits (profiles/r1_lat_*),
// 3.1 us per hop against 1.7 us for one CPU thread of the reference.  Here six warps share one query, in three roles:
//   * the driver (warp 0): pick, adjacency row, visited filter, acceptance test;
//   * four workers (warps 1..4): fetch and evaluate the fresh rows of the expansion, a quarter each, every warp with ALL
//     its rows in flight at once (registers are free at one CTA per query) — one HBM round trip per hop; after handing
//     in their distances they prefetch, into L2, the vectors of the links of every row that will enter the list
//     (SearchParams::pf2, see cta_rows), so that the hop which expands it later is an L2 round trip;
//   * the merge warp (warp 5): owns the sorted list — inserts the accepted candidates, hands back the list length and
//     the first unexpanded entry.
//   The hop is software-pipelined: the next node to expand is min(first unexpanded list entry, smallest accepted
//   candidate), which is known BEFORE the accepted candidates are merged into the list.  So the driver hands the
//   candidates to the merge warp and goes straight on to the next hop (adjacency, filter, row fetch); the merge (~28 % of
//   a hop's instructions in the one-warp variant) runs beside them and is waited for only before the next acceptance test.
// Why this is the same search: a chosen candidate always survives the merge's truncation (it is smaller than the list's
// worst entry, or the list is not full), and an old unexpanded entry that is the overall minimum cannot be truncated
// away (all entries before it would have to be old and expanded, i.e. the list was longer than its capacity).  The
// acceptance test of a hop runs after the previous hop's merge has finished, against the same `worst` as in the
// one-warp kernels; visited marks are set in the same order.  Output bytes, n_dist and n_hops are identical (tested:
// the whole parity suite also runs with this variant forced).
#pragma once
#include "search_kernel.cuh"

namespace fnb {

#define FNB_CTA_WARPS 6    // warp 0 drives, warps 1..4 evaluate rows, warp 5 owns the list merge
#define FNB_CTA_WORKERS 4
// warp-wide load batches one warp holds in registers in the CTA kernel: a quarter of 32 rows if the staging registers
// (NB x CH uint4 per lane) allow — 24 uint4 for rows of up to 512 B, 16 for the whole-warp-per-row shapes, whose query
// alone takes up to 16 uint4 per lane
__host__ __device__ constexpr int fnb_cta_batches(int g, int ch) {
  const int want = 32 / (32 / g) / FNB_CTA_WORKERS < 1 ? 1 : 32 / (32 / g) / FNB_CTA_WORKERS;  // batches of a quarter of 32 rows
  const int cap = (g == 32 ? 16 : 24) / ch < 1 ? 1 : (g == 32 ? 16 : 24) / ch;
  return want < cap ? want : cap;
}

__device__ __forceinline__ void distances_ready_arrive();

// Rows ids[0..n) (shared memory), a quarter per worker warp: batch b (RPI rows) belongs to worker b % 4.  Same arithmetic
// and reduction order as batch_distance.  Distances go to dist[0..n) in shared memory; the worker then ARRIVES at
// "distances ready" itself, because of what may follow:
// Two-hop prefetch (SearchParams::pf2, batches of a few queries on an otherwise idle GPU).  A row whose distance is
// below the list's worst entry (hint_hi: its distance word as of the previous hop; 0 while the list is still filling,
// when every row would qualify) enters the list and is likely to be expanded later.  After handing in its distances the
// worker reads that node's links (an L2 hit: the driver prefetched the adjacency row when it filtered the link) and
// prefetches the vector rows of ALL of them into L2, so that the hop which expands the node — the next one or the
// twentieth — finds its rows in L2 instead of HBM.  A prefetch cannot change a result; the cost is ~16 KB of L2 fills
// and ~130 requests on the SM's memory pipe per accepted candidate, off the critical path.
template <int DT, int METRIC, int G, int CH, bool EXACT>
__device__ __forceinline__ void cta_rows(const SearchParams& p, const uint4 (&q)[CH], const uint32_t* ids, uint32_t n,
                                         float* dist, int worker, int lane, uint32_t hint_hi) {
  typedef Arith<DT, METRIC> A;
  constexpr int RPI = 32 / G;
  constexpr int NB = fnb_cta_batches(G, CH);
  const int g = lane / G, pos = lane % G;
  bool arrived = false;
  for (uint32_t b0 = (uint32_t)worker; b0 * RPI < n; b0 += FNB_CTA_WORKERS * NB) {
    uint4 x[NB][CH];
    uint32_t rid[NB];
    unsigned qual = 0;  // batches of this pass whose row (this lane's group) passes the hint
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const uint32_t c = (b0 + (uint32_t)u * FNB_CTA_WORKERS) * RPI + (uint32_t)g;
      const bool ok = c < n;
      rid[u] = ids[ok ? c : 0];
      const uint4* row = p.vec + (size_t)rid[u] * p.stride + pos;
#pragma unroll
      for (int k = 0; k < CH; k++) x[u][k] = ldg_stream_if(row + k * G, ok && (EXACT || (uint32_t)(k * G + pos) < p.nchunks));
    }
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const uint32_t c = (b0 + (uint32_t)u * FNB_CTA_WORKERS) * RPI + (uint32_t)g;
      if ((b0 + (uint32_t)u * FNB_CTA_WORKERS) * RPI < n) {  // warp-uniform
        typename A::acc_t acc = 0;
#pragma unroll
        for (int k = 0; k < CH; k++)
          if (EXACT || (uint32_t)(k * G + pos) < p.nchunks) A::step(acc, q[k], x[u][k]);
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) acc = A::combine(acc, shfl_xor_t(acc, off));
        const float d = A::finish(acc);  // every lane of the group holds the full sum
        if (pos == 0 && c < n) dist[c] = d;
        if (c < n && ord_f32(d) < hint_hi) qual |= 1u << u;
      }
    }
    if ((b0 + FNB_CTA_WORKERS * NB) * RPI >= n) {  // warp-uniform: this worker's last pass of the round
      distances_ready_arrive();
      arrived = true;
    }
    if (p.pf2 && qual) {
      const size_t row_bytes = (size_t)p.stride * FNB_CHUNK_BYTES;
#pragma unroll
      for (int u = 0; u < NB; u++) {
        if (!((qual >> u) & 1u)) continue;
        for (uint32_t o = (uint32_t)pos * 4u; o < p.M; o += 4u * G) {  // pf2 implies M % 4 == 0: 16-byte slices of the links
          const uint4 a = __ldg(reinterpret_cast<const uint4*>(p.adj + (size_t)rid[u] * p.M + o));
          const uint32_t nbv[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
          for (int j = 0; j < 4; j++) {
            if (nbv[j] == rid[u]) continue;  // unused link slot
            const char* r = reinterpret_cast<const char*>(p.vec) + (size_t)nbv[j] * row_bytes;
            for (uint32_t l = 0; l < p.lines_per_row; l++) prefetch_l2(r + (size_t)l * 128u);
          }
        }
      }
    }
  }
  if (!arrived) distances_ready_arrive();  // no rows for this worker in this round
}

// Hand-off between the driver and the workers: two named barriers used as producer / consumer pairs (PTX bar.arrive /
// bar.sync with an explicit thread count — the arrangement of a warp-specialised kernel; the two sides reach them from
// different places in the code, which __syncthreads() does not promise to support):
//   barrier 1 "rows published":  the driver ARRIVES (it does not wait), the workers wait;
//   barrier 2 "distances ready": the workers ARRIVE, the driver waits.
// A worker cannot run ahead: after arriving at 2 it waits at 1, which needs the driver, who first waits at 2.
// Barrier 3 is the plain all-threads barrier at the end of a query.  bar.* are warp-aligned: reconverge first (inline
// asm does not make the compiler do it).
#define FNB_CTA_THREADS (FNB_CTA_WARPS * 32)
#define FNB_CTA_ROW_THREADS ((1 + FNB_CTA_WORKERS) * 32)  // the driver and the workers: the parties of barriers 1 and 2
__device__ __forceinline__ void rows_published_arrive() {
  __syncwarp();
  asm volatile("bar.arrive 1, %0;" ::"n"(FNB_CTA_ROW_THREADS) : "memory");
}
__device__ __forceinline__ void rows_published_wait() {
  __syncwarp();
  asm volatile("bar.sync 1, %0;" ::"n"(FNB_CTA_ROW_THREADS) : "memory");
}
__device__ __forceinline__ void distances_ready_arrive() {
  __syncwarp();
  asm volatile("bar.arrive 2, %0;" ::"n"(FNB_CTA_ROW_THREADS) : "memory");
}
__device__ __forceinline__ void distances_ready_wait() {
  __syncwarp();
  asm volatile("bar.sync 2, %0;" ::"n"(FNB_CTA_ROW_THREADS) : "memory");
}
__device__ __forceinline__ void cta_sync() {
  __syncwarp();
  asm volatile("bar.sync 3, %0;" ::"n"(FNB_CTA_THREADS) : "memory");
}
//   barrier 4 "candidates published": the driver ARRIVES, the merge warp waits;
//   barrier 5 "list merged":          the merge warp ARRIVES, the driver waits.           (64 threads each)
__device__ __forceinline__ void candidates_published_arrive() {
  __syncwarp();
  asm volatile("bar.arrive 4, 64;" ::: "memory");
}
__device__ __forceinline__ void candidates_published_wait() {
  __syncwarp();
  asm volatile("bar.sync 4, 64;" ::: "memory");
}
__device__ __forceinline__ void list_merged_arrive() {
  __syncwarp();
  asm volatile("bar.arrive 5, 64;" ::: "memory");
}
__device__ __forceinline__ void list_merged_wait() {
  __syncwarp();
  asm volatile("bar.sync 5, 64;" ::: "memory");
}

__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
  const uint32_t hi = __reduce_min_sync(FNB_FULL, (uint32_t)(v >> 32));
  const uint32_t lo = __reduce_min_sync(FNB_FULL, (uint32_t)(v >> 32) == hi ? (uint32_t)v : 0xffffffffu);
  return ((uint64_t)hi << 32) | lo;
}

template <int DT, int METRIC, int G, int CH, bool EXACT>
__global__ void __launch_bounds__(FNB_CTA_WARPS * 32, 2) fnb_search_cta_kernel(const SearchParams p) {
  extern __shared__ __align__(16) unsigned char fnb_smem[];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  volatile uint64_t* list = reinterpret_cast<volatile uint64_t*>(fnb_smem);
  uint32_t* tab = reinterpret_cast<uint32_t*>(fnb_smem + (size_t)p.Bcap * 8);
  uint32_t* ids = tab + p.vs_buckets * 4;                                         // 32 ids
  float* dist = reinterpret_cast<float*>(fnb_smem + p.warp_smem);                 // 32 distances
  volatile uint32_t* ctl = reinterpret_cast<volatile uint32_t*>(dist + 32);       // [0] rows of this round, ~0 = query done
  // driver <-> merge warp: [1] candidate mask of the round (~0 = query done), [2] list length, [3] pick start hint,
  // [4] index of the first unexpanded entry (or ~0), and the entry itself / the candidates' keys
  volatile uint64_t* first_unexp = reinterpret_cast<volatile uint64_t*>(ctl + 8);
  volatile uint64_t* pend = first_unexp + 1;                                      // 32 keys
  const int pos = lane % G;
  asm volatile("griddepcontrol.launch_dependents;");

  for (uint32_t qi = blockIdx.x; qi < p.Q; qi += gridDim.x) {
    uint4 q[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) q[k] = load_query_chunk<DT>(p, qi, (uint32_t)(k * G + pos));
    uint32_t ndist = 0, nhops = 0, len = 0;

    if (warp == FNB_CTA_WARPS - 1) {
      // ---- merge warp: owns the list between "candidates published" and "list merged" ----
      for (;;) {
        candidates_published_wait();
        const uint32_t am = ctl[1];
        if (am == 0xffffffffu && ctl[2] == 0xffffffffu) break;
        uint32_t len = ctl[2], start = ctl[3];
        const uint64_t key = pend[lane];
        const bool acc = (am >> lane) & 1u;
        if (am) merge_accepted(list, len, start, p.B, p.Bpow2, key, acc, lane);
        // the first unexpanded entry of the merged list, for the driver's pick
        uint64_t e_list = ~0ull;
        uint32_t i_list = 0xffffffffu;
        for (uint32_t base = start & ~31u; base < len; base += 32) {
          const uint32_t i = base + lane;
          const uint64_t e = (i < len) ? list[i] : 1ull;
          const unsigned b = __ballot_sync(FNB_FULL, !(e & 1ull));
          if (b) {
            const int src = __ffs(b) - 1;
            e_list = shfl64(e, src);
            i_list = base + (uint32_t)src;
            break;
          }
        }
        __syncwarp();  // every lane has read the control words above
        if (lane == 0) {
          ctl[2] = len;
          ctl[3] = i_list != 0xffffffffu ? i_list : start;
          ctl[4] = i_list;
          first_unexp[0] = e_list;
        }
        list_merged_arrive();
      }
    } else if (warp != 0) {
      // ---- workers: evaluate their quarter of every published round of rows ----
      for (;;) {
        rows_published_wait();
        const uint32_t n = ctl[0];
        if (n == 0xffffffffu) break;
        cta_rows<DT, METRIC, G, CH, EXACT>(p, q, ids, n, dist, warp - 1, lane, ctl[5]);  // (arrives at "distances ready")
      }
    } else {
      visited_clear(tab, p.vs_buckets, lane);
      if (lane == 0) ctl[5] = 0u;  // two-hop prefetch hint: nothing qualifies during entry selection and while the list fills
      __syncwarp();
      if (p.N > 0) {
        // ---- entry selection: strided probes, first strict minimum wins (Index.h:845-870) ----
        uint64_t best = ~0ull;
        for (uint32_t base = 0; base < p.nprobe; base += 32) {
          const uint32_t pi = base + lane;
          const uint32_t n = min(32u, p.nprobe - base);
          if (pi < p.nprobe) ids[lane] = pi * p.step;
          if (lane == 0) ctl[0] = n;
          rows_published_arrive();
          distances_ready_wait();
          if (pi < p.nprobe) {
            const uint64_t k = ((uint64_t)ord_f32(dist[lane]) << 32) | pi;
            best = k < best ? k : best;
          }
          __syncwarp();
        }
        best = warp_min_u64(best);
        ndist = p.nprobe;
        uint32_t cur = (uint32_t)best * p.step;
        if (lane == 0) {
          list[0] = (best & 0xffffffff00000000ull) | ((uint64_t)cur << 1) | 1ull;  // the entry node, already being expanded
          visited_test_and_set(tab, p, cur);
        }
        len = 1;
        __syncwarp();
        uint64_t pkey = 0;  // this lane's accepted candidate of the previous round, not merged yet
        bool pacc = false;
        if (lane == 0) {
          ctl[2] = 1u;  // list length
          ctl[3] = 0u;  // pick start hint
        }

        // ---- main loop (Index.h:627-658), software-pipelined: see the header comment ----
        while (cur != FNB_EMPTY) {
          nhops++;
          for (uint32_t l0 = 0; l0 < p.M; l0 += 32) {
            // hand the previous round's accepted candidates to the merge warp, then start this round
            pend[lane] = pkey;
            {
              const unsigned am = __ballot_sync(FNB_FULL, pacc);
              if (lane == 0) ctl[1] = am;
            }
            candidates_published_arrive();
            pacc = false;
            uint32_t nb = cur;
            if (l0 + lane < p.M) nb = __ldg(p.adj + (size_t)cur * p.M + l0 + lane);
            const bool fresh = (nb != cur) && visited_test_and_set(tab, p, nb);
            const unsigned fm = __ballot_sync(FNB_FULL, fresh);
            const uint32_t n = (uint32_t)__popc(fm);
            const int myrank = __popc(fm & ((1u << lane) - 1u));
            if (fresh) {
              ids[myrank] = nb;
              const uint32_t* arow = p.adj + (size_t)nb * p.M;  // whichever of them is expanded later finds its links in L2
              for (uint32_t o = 0; o < p.M; o += 32) prefetch_l2(arow + o);
            }
            if (n) {
              if (lane == 0) ctl[0] = n;
              rows_published_arrive();  // the workers start fetching
              ndist += n;
              distances_ready_wait();
            }
            list_merged_wait();  // the list now holds every earlier round's candidates
            len = ctl[2];
            if (n) {
              const bool full = len >= p.B;
              const uint32_t worst_hi = (uint32_t)(list[len - 1] >> 32);
              pkey = make_key(fresh ? dist[myrank] : 0.f, nb);
              pacc = fresh && (!full || (uint32_t)(pkey >> 32) < worst_hi);
              if (lane == 0) ctl[5] = full ? worst_hi : 0u;  // the workers' hint for the next round
            }
            __syncwarp();
          }
          // ---- next node: min(first unexpanded list entry, smallest pending candidate) ----
          const uint32_t i_list = ctl[4];
          const uint64_t e_list = i_list != 0xffffffffu ? first_unexp[0] : ~0ull;
          uint64_t kmin;
          for (;;) {
            kmin = warp_min_u64(pacc ? pkey : ~0ull);
            if (!(kmin < e_list)) break;
            // About to expand a candidate that is not in the list yet: make sure it is not a node the visited set forgot
            // (then its key is already in the list, possibly expanded; the merge would drop it, and so must the pick).
            bool known = false;
            for (uint32_t i = lane; i < len; i += 32) known |= (list[i] & ~1ull) == kmin;
            if (!__any_sync(FNB_FULL, known)) break;
            if (pacc && pkey == kmin) pacc = false;
          }
          if (kmin < e_list) {  // a candidate of this round (strict: an equal key is a node the visited set forgot)
            cur = (uint32_t)kmin >> 1;
            if (pacc && pkey == kmin) pkey |= 1ull;  // enters the list as expanded (a duplicated link: both copies, one is dropped)
          } else if (e_list != ~0ull) {
            cur = (uint32_t)e_list >> 1;
            __syncwarp();
            if (lane == 0) list[i_list] = e_list | 1ull;  // the merge warp is idle between "list merged" and the next hand-over
            __syncwarp();
          } else {
            cur = FNB_EMPTY;
          }
        }
        len = ctl[2];
      }
      __syncwarp();  // every lane has read the list length
      if (lane == 0) {
        ctl[1] = 0xffffffffu;
        ctl[2] = 0xffffffffu;
      }
      candidates_published_arrive();  // releases the merge warp from this query
      if (lane == 0) ctl[0] = 0xffffffffu;
      rows_published_arrive();  // releases the workers from this query

      // ---- output: ascending distance, label field of the node (Index.h:393-406) ----
      asm volatile("griddepcontrol.wait;" ::: "memory");
      for (uint32_t i = lane; i < p.K; i += 32) {
        float od = __int_as_float(0x7f800000);
        int32_t ol = -1;
        if (i < len) {
          const uint64_t e = list[i];
          od = unord_f32((uint32_t)(e >> 32));
          ol = p.labels ? __ldg(p.labels + ((uint32_t)e >> 1)) : (int32_t)((uint32_t)e >> 1);
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
    }
    cta_sync();  // the next query of this CTA reuses the shared memory
  }
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.done && threadIdx.x == 0) {
    __threadfence_system();  // this CTA's results (possibly in pinned host memory) before its count
    if (atomicAdd(p.done, 1u) == gridDim.x - 1u) {  // the last CTA of the grid: publish the totals, leave the slot clean
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

template <int DT, int METRIC, int G, int CH>
cudaError_t launch_search_cta(const SearchParams& p, int num_sms, cudaStream_t stream) {
  const bool exact = p.nchunks == (uint32_t)(G * CH);
  static LaunchCache cache[2][16];
  auto kern = exact ? fnb_search_cta_kernel<DT, METRIC, G, CH, true> : fnb_search_cta_kernel<DT, METRIC, G, CH, false>;
  const size_t smem = (size_t)p.warp_smem + 32 * 4 + 32 + 8 + 32 * 8;  // + distances, control words, first unexpanded entry, candidate keys
  int ctas_per_sm = 0;
  cudaError_t e = plan_launch(kern, FNB_CTA_WARPS * 32, smem, cache[exact ? 1 : 0], &ctas_per_sm);
  if (e != cudaSuccess) return e;
  long long grid = (long long)num_sms * ctas_per_sm;
  if (grid > (long long)p.Q) grid = p.Q;
  if (grid < 1) grid = 1;
  return launch_maybe_pdl(kern, (unsigned)grid, (unsigned)FNB_CTA_WARPS * 32u, smem, stream, p);
}

}  // namespace fnb
