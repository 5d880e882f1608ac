// CTA-per-query latency kernel with speculative row evaluation (sm_100a): fnb_search_cta_kernel's roles (driver, four
// workers, merge warp) and its traversal, bit for bit — plus a small cache of distances computed AHEAD of the walk.
//
// In fnb_search_cta_kernel every hop pays, in sequence, the adjacency row of the node (an L2 hit, ~250 cycles), one HBM
// round trip for the neighbours' vectors (~600-900 cycles) and the reduction / hand-off around it.  But most hops
// expand a node that has been sitting in the sorted list, unexpanded, for several hops (measured on cfg1-like graphs:
// 67 % of the hops at ef=100, 79 % at ef=200): the node to expand next is min(first unexpanded list
// entry, best new candidate), and new candidates beat the list's front only a quarter of the time.  So while the walk
// goes on, the driver takes the first of the list's three leading unexpanded entries that it has not handled yet (a
// "target"), reads its links, filters them through the visited set READ-ONLY, and hands the surviving rows to the
// workers together with the rows the current hop needs (one round of up to 64 rows, one HBM round trip).  The
// distances land in one of four slots {node, links, valid mask, distances}.  When the walk later expands that node,
// its links and the distances of its still-unvisited neighbours are already in shared memory: the hop costs the
// visited filter, the acceptance test and the wait for the merge warp — no memory round trip.
//
// Why the results cannot change: a distance is a pure function of (query, row), computed by the same code in the same
// order (cta_rows_spec = cta_rows); the visited set is only MODIFIED by the walk itself, at the same points and in the
// same order as without speculation (targets are filtered with visited_peek); acceptance, merge and pick are
// untouched.  A link that was visited when the target was filtered but is "fresh" when the node is expanded (the
// visited set may forget, see visited_test_and_set) has no cached distance: its row joins the hop's own round, as do
// all rows of a node that was not a target.  n_dist counts, as in the reference, the rows the WALK evaluates — rows
// evaluated ahead of time and never used are extra memory traffic, not distance computations of the algorithm.
//
// Hand-offs: the barriers of search_cta_kernel.cuh.  One round of rows is in flight at a time: the driver posts a
// round (bar.arrive 1), and collects it (bar.sync 2) when it needs one of its distances, before it posts the next
// round, and before the query ends — exactly once per round.
#pragma once
#include "search_cta_kernel.cuh"

namespace fnb {

#define FNB_SPEC_SLOTS 4   // cached targets
#define FNB_SPEC_DEPTH 3   // leading unexpanded list entries considered as targets
#define FNB_SPEC_ROWS 64   // rows per round: up to 32 of the hop + up to 32 of the target
// shared memory beyond SearchParams::warp_smem (see the carve-up in the kernel)
#define FNB_SPEC_EXTRA_SMEM 2048u

// Development build (-DFNB_SPEC_DEBUG, tools/spec_probe.py): the driver counts events and clock cycles of its waits, and
// out_ndist receives counter SearchParams::dbg - 1 instead of n_dist.  Not compiled into the shipped library.
#ifdef FNB_SPEC_DEBUG
#define FNB_DBG(...) __VA_ARGS__
// segment i of the hop ends here: its cycles go to counter 4 + i if the hop is of the class asked for
#define FNB_SEG(i)                      \
  {                                     \
    const uint32_t now = clock();       \
    if (dbg_on) dbgc[4 + (i)] += now - dbg_t; \
    dbg_t = clock();                    \
  }
#define FNB_DBG_END                                                        \
  dbgc[15] = clock() - dbg_t_loop;                                         \
  dbgc[0] = nhops;                                                         \
  if (p.dbg & 0xffu) {                                                     \
    uint32_t v = 0;                                                        \
    _Pragma("unroll") for (int i = 0; i < 16; i++) if ((p.dbg & 0xffu) == (uint32_t)i + 1u) v = dbgc[i]; \
    ndist = v;                                                             \
  }
#else
#define FNB_DBG(...)
#define FNB_SEG(i)
#define FNB_DBG_END
#endif

__host__ __device__ constexpr int fnb_spec_batches(int g, int ch) {
  const int want = FNB_SPEC_ROWS / (32 / g) / FNB_CTA_WORKERS < 1 ? 1 : FNB_SPEC_ROWS / (32 / g) / FNB_CTA_WORKERS;
  const int cap = (g == 32 ? 16 : 24) / ch < 1 ? 1 : (g == 32 ? 16 : 24) / ch;
  return want < cap ? want : cap;
}

// cta_rows with a destination per row: rows[0..n) are evaluated (a quarter per worker warp, same arithmetic and
// reduction order as batch_distance) and the distance of row c goes to dbuf[dst[c]].
template <int DT, int METRIC, int G, int CH, bool EXACT>
__device__ __forceinline__ void cta_rows_spec(const SearchParams& p, const uint4 (&q)[CH], const uint32_t* rows,
                                              const uint32_t* dst, uint32_t n, float* dbuf, int worker, int lane) {
  typedef Arith<DT, METRIC> A;
  constexpr int RPI = 32 / G;
  constexpr int NB = fnb_spec_batches(G, CH);
  const int g = lane / G, pos = lane % G;
  for (uint32_t b0 = (uint32_t)worker; b0 * RPI < n; b0 += FNB_CTA_WORKERS * NB) {
    uint4 x[NB][CH];
#pragma unroll
    for (int u = 0; u < NB; u++) {
      const uint32_t c = (b0 + (uint32_t)u * FNB_CTA_WORKERS) * RPI + (uint32_t)g;
      const bool ok = c < n;
      const uint32_t rid = rows[ok ? c : 0];
      const uint4* row = p.vec + (size_t)rid * p.stride + pos;
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
        if (pos == 0 && c < n) dbuf[dst[c]] = A::finish(acc);
      }
    }
  }
}

template <int DT, int METRIC, int G, int CH, bool EXACT>
__global__ void __launch_bounds__(FNB_CTA_WARPS * 32, 2) fnb_search_cta_spec_kernel(const SearchParams p) {
  extern __shared__ __align__(16) unsigned char fnb_smem[];
  constexpr int S = FNB_SPEC_SLOTS;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const unsigned lt = (1u << lane) - 1u;
  volatile uint64_t* list = reinterpret_cast<volatile uint64_t*>(fnb_smem);
  uint32_t* tab = reinterpret_cast<uint32_t*>(fnb_smem + (size_t)p.Bcap * 8);
  // beyond warp_smem (FNB_SPEC_EXTRA_SMEM bytes):
  float* dbuf = reinterpret_cast<float*>(fnb_smem + p.warp_smem);  // [0,32) the hop's own rows, [32 + 32 s + j] slot s, link j
  uint32_t* rows = reinterpret_cast<uint32_t*>(dbuf + 32 + 32 * S);                // the round: row ids ...
  uint32_t* dst = rows + FNB_SPEC_ROWS;                                            // ... and where their distances go
  volatile uint32_t* slinks = dst + FNB_SPEC_ROWS;                                 // slot s: the node's first 32 links
  volatile uint32_t* snode = slinks + 32 * S;                                      // slot s: node id, FNB_EMPTY = free
  volatile uint32_t* sround = snode + S;                                           // slot s: the round its distances came / come with
  volatile uint32_t* svmask = sround + S;                                          // slot s: links with a distance
  // [0] rows of this round, ~0 = query done; driver <-> merge warp: [1] candidate mask of the round (~0 = query done),
  // [2] list length, [3] pick start hint, [4] index of the first unexpanded entry (or ~0)
  volatile uint32_t* ctl = svmask + S;
  volatile uint64_t* first_unexp = reinterpret_cast<volatile uint64_t*>(ctl + 8);  // the first FNB_SPEC_DEPTH unexpanded entries (~0 = none)
  volatile uint64_t* pend = first_unexp + 4;                                       // 32 candidate keys
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
        // the leading unexpanded entries of the merged list: the first is the driver's pick, all are targets.  Lane j
        // keeps the j-th; the search goes at most one 32-entry window beyond the window of the first.
        uint64_t mine = ~0ull;
        uint32_t i_list = 0xffffffffu, first_base = 0;
        int cnt = 0;
        for (uint32_t base = start & ~31u; base < len && cnt < FNB_SPEC_DEPTH; base += 32) {
          if (cnt > 0 && base > first_base + 32u) break;
          const uint32_t i = base + lane;
          const uint64_t e = (i < len) ? list[i] : 1ull;
          unsigned b = __ballot_sync(FNB_FULL, !(e & 1ull));
          while (b && cnt < FNB_SPEC_DEPTH) {
            const int src = __ffs(b) - 1;
            const uint64_t v = shfl64(e, src);
            if (cnt == 0) {
              i_list = base + (uint32_t)src;
              first_base = base;
            }
            if (lane == cnt) mine = v;
            cnt++;
            b &= b - 1u;
          }
        }
        __syncwarp();  // every lane has read the control words above
        if (lane < FNB_SPEC_DEPTH) first_unexp[lane] = mine;
        if (lane == 0) {
          ctl[2] = len;
          ctl[3] = i_list != 0xffffffffu ? i_list : start;
          ctl[4] = i_list;
        }
        list_merged_arrive();
      }
    } else if (warp != 0) {
      // ---- workers: evaluate their quarter of every published round of rows ----
      for (;;) {
        rows_published_wait();
        const uint32_t n = ctl[0];
        if (n == 0xffffffffu) break;
        cta_rows_spec<DT, METRIC, G, CH, EXACT>(p, q, rows, dst, n, dbuf, warp - 1, lane);
        distances_ready_arrive();
      }
    } else {
      visited_clear(tab, p.vs_buckets, lane);
      if (lane < S) snode[lane] = FNB_EMPTY;
      __syncwarp();
      bool outstanding = false;  // a round has been posted and not collected
      if (p.N > 0) {
        // ---- entry selection: strided probes, first strict minimum wins (Index.h:845-870) ----
        uint64_t best = ~0ull;
        for (uint32_t base = 0; base < p.nprobe; base += 32) {
          const uint32_t pi = base + lane;
          const uint32_t n = min(32u, p.nprobe - base);
          if (pi < p.nprobe) {
            rows[lane] = pi * p.step;
            dst[lane] = (uint32_t)lane;
          }
          if (lane == 0) ctl[0] = n;
          rows_published_arrive();
          distances_ready_wait();
          if (pi < p.nprobe) {
            const uint64_t k = ((uint64_t)ord_f32(dbuf[lane]) << 32) | pi;
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
        uint32_t round = 0;                                // rounds posted in the main loop
        uint32_t t_node = FNB_EMPTY, t_link = FNB_EMPTY;  // target whose links are loaded (this lane's link) but not posted yet
        uint32_t x_node = FNB_EMPTY, x_link = FNB_EMPTY;  // the best new candidate's links, loaded while the pick is decided
        FNB_DBG(uint32_t dbgc[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0}; const uint32_t dbg_t_loop = clock();
                uint32_t dbg_t = 0; bool dbg_on = false; const uint32_t dbg_cls = p.dbg >> 8;)

        // ---- main loop (Index.h:627-658), software-pipelined as in fnb_search_cta_kernel ----
        while (cur != FNB_EMPTY) {
          nhops++;
          FNB_DBG(dbg_t = clock(); dbg_on = false; uint32_t dbg_a = 0;)
          for (uint32_t l0 = 0; l0 < p.M; l0 += 32) {
            // hand the previous round's accepted candidates to the merge warp, then start this round
            pend[lane] = pkey;
            {
              const unsigned am = __ballot_sync(FNB_FULL, pacc);
              if (lane == 0) ctl[1] = am;
            }
            candidates_published_arrive();
            pacc = false;
            FNB_DBG(dbg_a = clock() - dbg_t; dbg_t = clock();)

            // -- the node's links: a slot (with distances), a target not posted yet, the early load of the pick, or memory
            uint32_t nb = cur, vm = 0;
            float cd = 0.f;
            int s = -1;
            if (l0 == 0) {
              const unsigned hit = __ballot_sync(FNB_FULL, lane < S && snode[lane] == cur);
              if (hit) s = __ffs(hit) - 1;
            }
            if (s >= 0) {
              if (outstanding && sround[s] == round) {  // its distances are in the round in flight
                distances_ready_wait();
                outstanding = false;
              }
              nb = slinks[s * 32 + lane];
              vm = svmask[s];
              cd = dbuf[32 + s * 32 + lane];
              __syncwarp();
              if (lane == 0) snode[s] = FNB_EMPTY;  // consumed (links and distances are in registers now)
              __syncwarp();
            } else if (l0 == 0 && t_node == cur) {
              nb = t_link;
              t_node = FNB_EMPTY;
            } else if (l0 == 0 && x_node == cur) {
              nb = x_link;
            } else if (l0 + lane < p.M) {
              nb = __ldg(p.adj + (size_t)cur * p.M + l0 + lane);
            }
            FNB_DBG(dbg_on = dbg_cls == 0u || (dbg_cls == 1u && s >= 0) || (dbg_cls == 2u && s < 0); if (dbg_on) dbgc[4] += dbg_a;
                    if (s >= 0) dbgc[1]++;)
            FNB_SEG(1)
            const bool fresh = (nb != cur) && visited_test_and_set(tab, p, nb);
            const bool have = fresh && ((vm >> lane) & 1u);
            const bool need = fresh && !have;
            const uint32_t n = (uint32_t)__popc(__ballot_sync(FNB_FULL, fresh));
            const unsigned nm = __ballot_sync(FNB_FULL, need);
            const uint32_t nA = (uint32_t)__popc(nm);
            const uint32_t rankA = (uint32_t)__popc(nm & lt);
            if (fresh) {
              const uint32_t* arow = p.adj + (size_t)nb * p.M;  // whichever of them is expanded later finds its links in L2
              for (uint32_t o = 0; o < p.M; o += 32) prefetch_l2(arow + o);
            }
            FNB_SEG(2)
            // -- the target whose links were loaded during the previous hop: rows of its unvisited links join this round
            const bool postB = (l0 == 0) && (t_node != FNB_EMPTY);
            const bool bfresh = postB && (t_link != t_node) && !visited_peek(tab, p, t_link);
            const unsigned bm = __ballot_sync(FNB_FULL, bfresh);
            const uint32_t nB = (uint32_t)__popc(bm);
            FNB_SEG(3)
            if (postB || nA) {
              if ((nA + nB) && outstanding) {  // one round in flight at a time: rows[] / dst[] are the workers' until they are done
                distances_ready_wait();
                outstanding = false;
              }
              FNB_SEG(4)
              uint32_t sB = 0;
              if (postB) {
                // a free slot, else the one filled longest ago
                const uint32_t age = lane < S ? (snode[lane] == FNB_EMPTY ? 0u : ((sround[lane] + 1u) << 5)) | (uint32_t)lane
                                              : 0xffffffffu;
                sB = __reduce_min_sync(FNB_FULL, age) & 31u;
                __syncwarp();  // every lane has read the slot words
                if (lane == 0) {
                  snode[sB] = t_node;
                  sround[sB] = round + ((nA + nB) ? 1u : 0u);
                  svmask[sB] = bm;
                }
                slinks[sB * 32 + lane] = t_link;
              }
              if (need) {
                rows[rankA] = nb;
                dst[rankA] = rankA;
              }
              if (bfresh) {
                const uint32_t r = nA + (uint32_t)__popc(bm & lt);
                rows[r] = t_link;
                dst[r] = 32u + sB * 32u + (uint32_t)lane;
              }
              if (nA + nB) {
                if (lane == 0) ctl[0] = nA + nB;
                rows_published_arrive();  // the workers start fetching
                outstanding = true;
                round++;
                FNB_DBG(dbgc[2]++; dbgc[3] += nA + nB;)
              }
              if (postB) t_node = FNB_EMPTY;
            }
            ndist += n;
            FNB_SEG(5)
            if (nA) {  // the hop waits for its own rows only
              distances_ready_wait();
              outstanding = false;
            }
            FNB_SEG(6)
            list_merged_wait();  // the list now holds every earlier round's candidates
            len = ctl[2];
            FNB_SEG(7)
            // -- next target: the first leading unexpanded entry that has no slot yet; its links are loaded now, used a hop later
            if (l0 + 32 >= p.M && t_node == FNB_EMPTY) {
#pragma unroll
              for (int j = 0; j < FNB_SPEC_DEPTH; j++) {
                const uint64_t e = first_unexp[j];
                if (e == ~0ull || t_node != FNB_EMPTY) continue;
                const uint32_t id = (uint32_t)e >> 1;
                if (!__any_sync(FNB_FULL, lane < S && snode[lane] == id)) t_node = id;
              }
              if (t_node != FNB_EMPTY) t_link = (uint32_t)lane < p.M ? __ldg(p.adj + (size_t)t_node * p.M + lane) : t_node;
            }
            FNB_SEG(8)
            if (n) {
              const bool full = len >= p.B;
              const uint32_t worst_hi = (uint32_t)(list[len - 1] >> 32);
              const float d = have ? cd : (need ? dbuf[rankA] : 0.f);
              pkey = make_key(fresh ? d : 0.f, nb);
              pacc = fresh && (!full || (uint32_t)(pkey >> 32) < worst_hi);
            }
            __syncwarp();
            FNB_SEG(9)
          }
          // ---- next node: min(first unexpanded list entry, smallest pending candidate) ----
          const uint32_t i_list = ctl[4];
          const uint64_t e_list = i_list != 0xffffffffu ? first_unexp[0] : ~0ull;
          uint64_t kmin;
          x_node = FNB_EMPTY;
          for (;;) {
            kmin = warp_min_u64(pacc ? pkey : ~0ull);
            if (!(kmin < e_list)) break;
            if (x_node == FNB_EMPTY) {  // most likely the pick: load its links while the list is searched for it
              x_node = (uint32_t)kmin >> 1;
              x_link = (uint32_t)lane < p.M ? __ldg(p.adj + (size_t)x_node * p.M + lane) : x_node;
            }
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
          FNB_SEG(10)
        }
        len = ctl[2];
        FNB_DBG_END
      }
      if (outstanding) {  // a round of target rows nobody needs any more: still collected, once per round
        distances_ready_wait();
        outstanding = false;
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
cudaError_t launch_search_cta_spec(const SearchParams& p, int num_sms, cudaStream_t stream) {
  const bool exact = p.nchunks == (uint32_t)(G * CH);
  static LaunchCache cache[2][16];
  auto kern = exact ? fnb_search_cta_spec_kernel<DT, METRIC, G, CH, true> : fnb_search_cta_spec_kernel<DT, METRIC, G, CH, false>;
  const size_t smem = (size_t)p.warp_smem + FNB_SPEC_EXTRA_SMEM;
  int ctas_per_sm = 0;
  cudaError_t e = plan_launch(kern, FNB_CTA_WARPS * 32, smem, cache[exact ? 1 : 0], &ctas_per_sm);
  if (e != cudaSuccess) return e;
  long long grid = (long long)num_sms * ctas_per_sm;
  if (grid > (long long)p.Q) grid = p.Q;
  if (grid < 1) grid = 1;
  return launch_maybe_pdl(kern, (unsigned)grid, (unsigned)FNB_CTA_WARPS * 32u, smem, stream, p);
}

}  // namespace fnb
