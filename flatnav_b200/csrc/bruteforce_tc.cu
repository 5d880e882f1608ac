// Exact top-K scan, tensor-core edition (sm_100a): a tcgen05 GEMM FILTERS candidates, CUDA cores re-rank them
// exactly.  This is the only place the engine touches the tensor cores (BASELINE.json north_star: "tensor cores
// used only for the dense brute-force ground-truth / exact-rerank GEMM, never for the irregular traversal").
//
// The reference has no brute force (its recall tests take ground truth from outside,
// python-bindings/unit_tests/test_utils.py:57-91); the semantics — top-K by (distance, node id), distances in the
// traversal kernel's arithmetic — are those of bruteforce.cu / oracle `ora_bruteforce`, and the output of this
// path is bit-identical to theirs.
//
// Phases (all on the replica's stream):
//  1. prep      vectors and queries are split into bf16 hi + lo parts (x = hi + lo + O(2^-18 |x|)), rows padded
//               to a multiple of 64 elements, K-major.  uint8 / int8 values are exact in bf16: hi only.  Queries
//               are pre-scaled by -2 (L2) or -1 (IP) so that the score of (q, x) is  s = q'.x + |x|^2  (L2) or
//               s = q'.x (IP):  distance = s + |q|^2  resp.  1 + s.
//  2. filter    bf_tc_kernel: persistent CTAs, one work unit = 128 queries x one slice of the database.
//               warp 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle), warp 1 = tcgen05.mma issuer
//               (128 x 128 x 16 bf16, three passes hi.hi + hi.lo + lo.hi into one fp32 accumulator in TMEM, two
//               accumulator buffers), warps 4-7 = epilogue: tcgen05.ld the 128 x 128 scores, one query row per
//               thread, keep the Kp = K + 6 smallest in a per-thread list in shared memory.
//  3. re-rank   bf_rerank_kernel: one warp per query evaluates every kept candidate with the exact arithmetic of
//               the traversal kernel and keeps the top K by (distance, id).  A slice whose list was full and whose
//               largest kept score is within the error bound eps of the K-th exact distance might have dropped a
//               true neighbour: the query is flagged and
//  4. re-scan   flagged queries (rare) go through the CUDA-core exact scan of bruteforce.cu.
//
// Error bound.  |s_gemm - s_true| <= eps(q) = 1.5 * [ c * |q'| * max|x| + c2 * (|q| + max|x|)^2 ] with
// c = 3 * 2^-18 (dropped lo.lo term and split residuals) + (3 D / 16) * 2^-21 (fp32 accumulation in the tensor
// core) and c2 = D * 2^-23 (rounding of the fp32 reference arithmetic itself and of the norms).  For integer data
// with D * maxabs^2 * 2 < 2^24 every step is exact and eps = 0.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include "../../include/flatnav_b200.h"
#include "bf_common.cuh"

namespace fnb {
namespace tc {

constexpr int BM = 128;       // queries per tile  = TMEM lanes
constexpr int BN = 128;       // database rows per tile = TMEM columns of one accumulator
constexpr int BK = 64;        // bf16 elements per k-block: one 128-byte swizzled row
constexpr int UMMA_K = 16;    // K of one tcgen05.mma.kind::f16
constexpr int THREADS = 256;  // warp 0 TMA, warp 1 MMA, warp 2 TMEM alloc, warps 4-7 epilogue
constexpr uint32_t TILE_BYTES = 128 * BK * 2;  // 16 KB operand block
constexpr int ACC_BUFS = 2;
constexpr int MAX_STAGES = 6;
constexpr uint32_t TMEM_COLS = ACC_BUFS * BN;  // 256

// ---- PTX wrappers -------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)m) : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[smem] . B[smem]^T, bf16 x bf16 -> fp32, issued by one thread for the whole CTA
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5}, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
// the registers are tied to the wait so that no use of them can be scheduled before the load has landed
__device__ __forceinline__ void tmem_wait_ld(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]),
                 "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]),
                 "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]),
                 "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}
__device__ __forceinline__ void epilogue_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major [rows][64 bf16] block written by TMA with 128-byte swizzle:
// 8-row groups are 1024 B apart (SBO), rows 128 B; bits [46,48) = descriptor version 1, [61,64) = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
// kind::f16 instruction descriptor: D = f32 (bits 4-5 = 1), A = B = bf16 (bits 7-9, 10-12 = 1), both K-major,
// N >> 3 at bit 17, M >> 4 at bit 24.
constexpr uint32_t IDESC = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

struct FilterParams {
  const float* __restrict__ xn;    // [n_tiles * 128]: |x|^2 (L2) or 0 (IP); +inf beyond N
  float* __restrict__ cand_s;      // [Q][S][Kp]
  uint32_t* __restrict__ cand_id;  // [Q][S][Kp]
  uint32_t* __restrict__ cand_cnt; // [Q][S]
  uint32_t Q, n_qtiles, n_tiles, S, tiles_per_split, kblocks, Kp, stages, n_units;
};

__host__ __device__ inline uint32_t smem_lists_offset(uint32_t stages, uint32_t stage_bytes) { return stages * stage_bytes; }

// per-thread unsorted list of the Kp smallest scores seen so far, [k][128 threads] in shared memory
__device__ __forceinline__ void list_insert(float* ls, uint32_t* li, const uint32_t Kp, uint32_t& cnt, uint32_t& maxpos,
                                            float& tau, const float s, const uint32_t id) {
  const uint32_t slot = cnt < Kp ? cnt : maxpos;
  ls[slot * 128] = s;
  li[slot * 128] = id;
  if (cnt < Kp) cnt++;
  if (cnt == Kp) {  // (re)locate the largest kept score: it is the admission threshold from now on
    float m = ls[0];
    uint32_t mp = 0;
#pragma unroll 4
    for (uint32_t k = 1; k < Kp; k++) {
      const float v = ls[k * 128];
      const bool g = v > m;
      m = g ? v : m;
      mp = g ? k : mp;
    }
    tau = m;
    maxpos = mp;
  }
}

// Kp == KREG: the list lives in registers, sorted ascending (+inf = empty): find the position with KREG
// independent compares, shift with static selects.  No shared-memory latency on the (divergent) admission path.
template <int KREG>
__device__ __forceinline__ void reg_insert(float (&rs)[KREG], uint32_t (&ri)[KREG], float& tau, const uint32_t Kp,
                                           const float v, const uint32_t vid) {
  uint32_t pos = 0;
#pragma unroll
  for (int k = 0; k < KREG; k++) pos += (rs[k] <= v) ? 1u : 0u;
#pragma unroll
  for (int k = KREG - 1; k >= 0; k--) {
    const bool keep = (uint32_t)k < pos, here = (uint32_t)k == pos;
    const float below = k > 0 ? rs[k > 0 ? k - 1 : 0] : v;
    const uint32_t below_i = k > 0 ? ri[k > 0 ? k - 1 : 0] : vid;
    rs[k] = keep ? rs[k] : (here ? v : below);
    ri[k] = keep ? ri[k] : (here ? vid : below_i);
  }
  tau = rs[KREG - 1];  // Kp == KREG on this path: the threshold is the last register
}

template <int NSPLIT, int KREG>
__global__ void __launch_bounds__(THREADS, 1)
    bf_tc_kernel(const __grid_constant__ CUtensorMap tm_qh, const __grid_constant__ CUtensorMap tm_ql,
                 const __grid_constant__ CUtensorMap tm_xh, const __grid_constant__ CUtensorMap tm_xl, const FilterParams p) {
  extern __shared__ unsigned char tc_smem_raw[];
  constexpr uint32_t STAGE_BYTES = 2u * NSPLIT * TILE_BYTES;
  const uint32_t raw = smem_u32(tc_smem_raw);
  const uint32_t pad = ((raw + 1023u) & ~1023u) - raw;  // 128-byte swizzle atoms need 1024-byte alignment
  unsigned char* base = tc_smem_raw + pad;
  unsigned char* after = base + p.stages * STAGE_BYTES;
  float* xn_s = reinterpret_cast<float*>(after);                   // [2][128]
  float* ls_all = xn_s + ACC_BUFS * BN;                            // [Kp][128]
  uint32_t* li_all = reinterpret_cast<uint32_t*>(ls_all + p.Kp * 128);  // [Kp][128]
  float* scratch = reinterpret_cast<float*>(li_all + p.Kp * 128);       // [32][128]: one chunk of scores per thread
  uint64_t* bars = reinterpret_cast<uint64_t*>(scratch + 32 * 128);
  uint64_t* full = bars;
  uint64_t* empty = bars + MAX_STAGES;
  uint64_t* tfull = bars + 2 * MAX_STAGES;
  uint64_t* tempty = tfull + ACC_BUFS;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tempty + ACC_BUFS);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tm_qh);
    tma_prefetch_desc(&tm_xh);
    if (NSPLIT == 2) {
      tma_prefetch_desc(&tm_ql);
      tma_prefetch_desc(&tm_xl);
    }
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t i = 0; i < p.stages; i++) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < ACC_BUFS; i++) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);  // one arrival per epilogue warp
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(tmem_ptr, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const uint32_t split = unit / p.n_qtiles, qt = unit - split * p.n_qtiles;
        const uint32_t t0 = split * p.tiles_per_split, t1 = min(p.n_tiles, t0 + p.tiles_per_split);
        for (uint32_t tile = t0; tile < t1; tile++) {
          for (uint32_t kb = 0; kb < p.kblocks; kb++) {
            mbar_wait(&empty[stage], phase ^ 1u);
            mbar_expect_tx(&full[stage], STAGE_BYTES);
            const uint32_t st = smem_u32(base + stage * STAGE_BYTES);
            tma_load_2d(&tm_qh, &full[stage], st, (int)(kb * BK), (int)(qt * BM));
            if (NSPLIT == 2) tma_load_2d(&tm_ql, &full[stage], st + TILE_BYTES, (int)(kb * BK), (int)(qt * BM));
            tma_load_2d(&tm_xh, &full[stage], st + NSPLIT * TILE_BYTES, (int)(kb * BK), (int)(tile * BN));
            if (NSPLIT == 2) tma_load_2d(&tm_xl, &full[stage], st + 3 * TILE_BYTES, (int)(kb * BK), (int)(tile * BN));
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: one thread drives the tensor core of the SM =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, ab = 0, aphase = 0;
      for (uint32_t unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
        const uint32_t split = unit / p.n_qtiles;
        const uint32_t t0 = split * p.tiles_per_split, t1 = min(p.n_tiles, t0 + p.tiles_per_split);
        for (uint32_t tile = t0; tile < t1; tile++) {
          mbar_wait(&tempty[ab], aphase ^ 1u);  // epilogue has drained this accumulator
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + ab * BN;
          for (uint32_t kb = 0; kb < p.kblocks; kb++) {
            mbar_wait(&full[stage], phase);
            tc_fence_after();
            const uint32_t a_h = smem_u32(base + stage * STAGE_BYTES);
            const uint32_t a_l = a_h + TILE_BYTES;
            const uint32_t b_h = a_h + NSPLIT * TILE_BYTES;
            const uint32_t b_l = b_h + TILE_BYTES;
#pragma unroll
            for (int pass = 0; pass < (NSPLIT == 2 ? 3 : 1); pass++) {
              const uint32_t a = pass == 2 ? a_l : a_h;
              const uint32_t b = pass == 1 ? b_l : b_h;
#pragma unroll
              for (int k = 0; k < BK / UMMA_K; k++)
                umma_bf16(d_tmem, make_desc(a + k * UMMA_K * 2), make_desc(b + k * UMMA_K * 2), IDESC,
                          (kb | (uint32_t)pass | (uint32_t)k) != 0u);
            }
            umma_commit(&empty[stage]);  // stage is free once these MMAs have read it
            if (++stage == p.stages) {
              stage = 0;
              phase ^= 1u;
            }
          }
          umma_commit(&tfull[ab]);  // accumulator complete
          ab ^= 1u;
          if (ab == 0) aphase ^= 1u;
        }
      }
    }
  } else if (warp >= 4) {
    // ===== epilogue: thread t owns query row t of the tile (TMEM lane t) =====
    const int t = threadIdx.x - 128;
    float* ls = ls_all + t;
    uint32_t* li = li_all + t;
    const uint32_t lane_base = (uint32_t)(t & ~31) << 16;
    uint32_t ab = 0, aphase = 0;
    for (uint32_t unit = blockIdx.x; unit < p.n_units; unit += gridDim.x) {
      const uint32_t split = unit / p.n_qtiles, qt = unit - split * p.n_qtiles;
      const uint32_t t0 = split * p.tiles_per_split, t1 = min(p.n_tiles, t0 + p.tiles_per_split);
      const uint32_t qrow = qt * BM + (uint32_t)t;
      const bool active = qrow < p.Q;
      uint32_t cnt = 0, maxpos = 0;
      float tau = active ? __int_as_float(0x7f800000) : __int_as_float(0xff800000);
      float rs[KREG > 0 ? KREG : 1];
      uint32_t ri[KREG > 0 ? KREG : 1];
#pragma unroll
      for (int k = 0; k < (KREG > 0 ? KREG : 1); k++) {
        rs[k] = __int_as_float(0x7f800000);
        ri[k] = 0xffffffffu;
      }
      float xn_next = t0 < t1 ? __ldg(p.xn + (size_t)t0 * BN + t) : 0.f;
      for (uint32_t tile = t0; tile < t1; tile++) {
        float* xs = xn_s + ab * BN;
        xs[t] = xn_next;
        if (tile + 1 < t1) xn_next = __ldg(p.xn + (size_t)(tile + 1) * BN + t);  // in flight during this tile
        epilogue_bar_sync();
        mbar_wait(&tfull[ab], aphase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < BN / 32; c++) {
          uint32_t r[32];
          tmem_ld32(tmem_base + lane_base + ab * BN + c * 32, r);
          float4 xv[8];
#pragma unroll
          for (int i = 0; i < 8; i++) xv[i] = reinterpret_cast<const float4*>(xs + c * 32)[i];
          tmem_wait_ld(r);
          // branch-free scan of the 32 scores of this chunk; admission is rare per thread (~Kp ln(n/Kp) / n)
          uint32_t mask = 0;
#pragma unroll
          for (int i = 0; i < 8; i++) {
            const float s0 = __uint_as_float(r[4 * i + 0]) + xv[i].x, s1 = __uint_as_float(r[4 * i + 1]) + xv[i].y;
            const float s2 = __uint_as_float(r[4 * i + 2]) + xv[i].z, s3 = __uint_as_float(r[4 * i + 3]) + xv[i].w;
            r[4 * i + 0] = __float_as_uint(s0);
            r[4 * i + 1] = __float_as_uint(s1);
            r[4 * i + 2] = __float_as_uint(s2);
            r[4 * i + 3] = __float_as_uint(s3);
            mask |= (s0 < tau ? 1u : 0u) << (4 * i + 0);
            mask |= (s1 < tau ? 1u : 0u) << (4 * i + 1);
            mask |= (s2 < tau ? 1u : 0u) << (4 * i + 2);
            mask |= (s3 < tau ? 1u : 0u) << (4 * i + 3);
          }
          if (mask) {  // park the chunk in shared memory so the admitted scores can be picked by position
            float* scr = scratch + t;
#pragma unroll
            for (int j = 0; j < 32; j++) scr[j * 128] = __uint_as_float(r[j]);
            const uint32_t id0 = tile * BN + c * 32;
            do {
              const uint32_t j = (uint32_t)__ffs(mask) - 1u;
              mask &= mask - 1u;
              const float sj = scr[j * 128];
              if (sj < tau) {
                if (KREG > 0) reg_insert<(KREG > 0 ? KREG : 1)>(rs, ri, tau, p.Kp, sj, id0 + j);
                else list_insert(ls, li, p.Kp, cnt, maxpos, tau, sj, id0 + j);
              }
            } while (mask);
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&tempty[ab]);
        ab ^= 1u;
        if (ab == 0) aphase ^= 1u;
      }
      if (active) {
        const size_t o = ((size_t)qrow * p.S + split) * p.Kp;
        if (KREG > 0) {
          cnt = 0;
#pragma unroll
          for (int k = 0; k < (KREG > 0 ? KREG : 1); k++) {
            if ((uint32_t)k < p.Kp && ri[k] != 0xffffffffu) {
              p.cand_s[o + k] = rs[k];
              p.cand_id[o + k] = ri[k];
              cnt++;
            }
          }
        } else {
          for (uint32_t k = 0; k < cnt; k++) {
            p.cand_s[o + k] = ls[k * 128];
            p.cand_id[o + k] = li[k * 128];
          }
        }
        p.cand_cnt[(size_t)qrow * p.S + split] = cnt;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ---- phase 1: bf16 hi / lo split, one warp per row --------------------------------------------------------
template <int DT>
__global__ void bf_prep_kernel(const unsigned char* __restrict__ src, size_t src_row_bytes, uint32_t rows, uint32_t dim,
                               uint32_t Dpad, float scale, __nv_bfloat16* __restrict__ hi, __nv_bfloat16* __restrict__ lo,
                               float* __restrict__ norm2, unsigned int* __restrict__ maxnorm_bits) {
  const int lane = threadIdx.x & 31;
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < rows; row += warps) {
    const unsigned char* s = src + (size_t)row * src_row_bytes;
    float acc = 0.f;
    for (uint32_t e = lane; e < Dpad; e += 32) {
      float v = 0.f;
      if (e < dim) {
        if (DT == DT_F32) v = reinterpret_cast<const float*>(s)[e];
        else if (DT == DT_U8) v = (float)s[e];
        else v = (float)reinterpret_cast<const signed char*>(s)[e];
      }
      acc = fmaf(v, v, acc);
      const float w = v * scale;  // scale is +-2^k: exact
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      hi[(size_t)row * Dpad + e] = h;
      if (lo) lo[(size_t)row * Dpad + e] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(FNB_FULL, acc, off);
    if (lane == 0) {
      norm2[row] = acc;
      if (maxnorm_bits) atomicMax(maxnorm_bits, __float_as_uint(sqrtf(acc)));
    }
  }
}

// xn[j] = |x_j|^2 (L2) or 0 (IP) for j < N, +inf for the padding rows of the last tile
__global__ void bf_xn_kernel(const float* __restrict__ norm2, float* __restrict__ xn, uint32_t N, uint32_t Npad, int is_ip) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < Npad) xn[i] = i < N ? (is_ip ? 0.f : norm2[i]) : __int_as_float(0x7f800000);
}

// ---- phase 3: exact re-rank, one warp per query ---------------------------------------------------------------
struct RerankParams {
  SearchParams sp;  // vec, stride, nchunks, queries, dim, query_vec_ok
  const int32_t* __restrict__ labels;
  const float* __restrict__ cand_s;
  const uint32_t* __restrict__ cand_id;
  const uint32_t* __restrict__ cand_cnt;
  const float* __restrict__ qn2;            // [Q] |q|^2
  const unsigned int* __restrict__ xmax_bits;  // max |x| as float bits
  float* __restrict__ out_dist;
  int32_t* __restrict__ out_label;
  unsigned int* n_unsafe;
  uint32_t* unsafe_list;
  unsigned long long* n_cand;
  uint32_t Q, K, Kcap, S, Kp;
  float c_rel, c_abs, qscale;
  int is_ip;
};

template <int DT, int METRIC, int G, int CH, bool EXACT>
__global__ void __launch_bounds__(128) bf_rerank_kernel(const RerankParams p) {
  extern __shared__ __align__(16) unsigned char rr_smem[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint64_t* list = reinterpret_cast<uint64_t*>(rr_smem) + (size_t)warp * (p.Kcap + 16);
  uint32_t* s_ids = reinterpret_cast<uint32_t*>(list + p.Kcap);
  const uint32_t qi = blockIdx.x * 4 + warp;
  if (qi >= p.Q) return;
  const int pos = lane % G;
  uint4 q[CH];
#pragma unroll
  for (int k = 0; k < CH; k++) q[k] = load_query_chunk<DT>(p.sp, qi, (uint32_t)(k * G + pos));

  uint32_t len = 0, total = 0;
  float min_full_max = __int_as_float(0x7f800000);
  for (uint32_t s = 0; s < p.S; s++) {
    const uint32_t cnt = p.cand_cnt[(size_t)qi * p.S + s];
    const size_t o = ((size_t)qi * p.S + s) * p.Kp;
    float smax = __int_as_float(0xff800000);
    total += cnt;
    for (uint32_t k0 = 0; k0 < cnt; k0 += 32) {
      const uint32_t k = k0 + lane;
      const bool valid = k < cnt;
      const uint32_t id = valid ? p.cand_id[o + k] : 0u;
      if (valid) smax = fmaxf(smax, p.cand_s[o + k]);
      const float d = batch_distance<DT, METRIC, G, CH, EXACT>(p.sp, q, id, valid, s_ids, lane, false);
      const uint64_t key = ((uint64_t)ord_f32(d) << 32) | (uint64_t)id;
      const uint64_t worst = len ? list[len - 1] : 0ull;
      const bool cand = valid && (len < p.K || key < worst);
      for (unsigned cm = __ballot_sync(FNB_FULL, cand); cm; cm &= cm - 1)
        warp_topk_insert(list, len, p.K, shfl64(key, __ffs(cm) - 1), lane);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) smax = fmaxf(smax, __shfl_xor_sync(FNB_FULL, smax, off));
    if (cnt == p.Kp) min_full_max = fminf(min_full_max, smax);
  }
  __syncwarp();
  // completeness proof: every dropped row of a full slice has score >= that slice's largest kept score
  const float qn2 = p.qn2[qi];
  const float qn = sqrtf(qn2), xmax = __uint_as_float(*p.xmax_bits);
  const float eps = 1.5f * (p.c_rel * p.qscale * qn * xmax + p.c_abs * (qn + xmax) * (qn + xmax));
  const float dK = len >= p.K ? unord_f32((uint32_t)(list[p.K - 1] >> 32)) : __int_as_float(0x7f800000);
  const float dK_score = dK - (p.is_ip ? 1.0f : qn2);
  const bool unsafe = !(min_full_max - eps > dK_score);
  for (uint32_t i = lane; i < p.K; i += 32) {
    float od = __int_as_float(0x7f800000);
    int32_t ol = -1;
    if (i < len) {
      const uint64_t e = list[i];
      od = unord_f32((uint32_t)(e >> 32));
      ol = __ldg(p.labels + (uint32_t)e);
    }
    p.out_dist[(size_t)qi * p.K + i] = od;
    p.out_label[(size_t)qi * p.K + i] = ol;
  }
  if (lane == 0) {
    atomicAdd(p.n_cand, (unsigned long long)total);
    if (unsafe) p.unsafe_list[atomicAdd(p.n_unsafe, 1u)] = qi;
  }
}

template <int DT, int METRIC, int G, int CH>
static cudaError_t launch_rerank(const RerankParams& p, uint32_t nchunks, cudaStream_t s) {
  const bool exact = nchunks == (uint32_t)(G * CH);
  auto kern = exact ? bf_rerank_kernel<DT, METRIC, G, CH, true> : bf_rerank_kernel<DT, METRIC, G, CH, false>;
  const size_t smem = (size_t)4 * (p.Kcap + 16) * 8;
  kern<<<(p.Q + 3) / 4, 128, smem, s>>>(p);
  return cudaGetLastError();
}
template <int DT, int METRIC>
static cudaError_t rerank_gc(const fnb_index* ix, const RerankParams& p, cudaStream_t s) {
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  if (ix->G == 4) {
    if (ch <= 1) return launch_rerank<DT, METRIC, 4, 1>(p, ix->nchunks, s);
    return launch_rerank<DT, METRIC, 4, 2>(p, ix->nchunks, s);
  }
  if (ix->G == 8) {
    switch (ch) {
      case 1: return launch_rerank<DT, METRIC, 8, 1>(p, ix->nchunks, s);
      case 2: return launch_rerank<DT, METRIC, 8, 2>(p, ix->nchunks, s);
      case 3: return launch_rerank<DT, METRIC, 8, 3>(p, ix->nchunks, s);
      default: return launch_rerank<DT, METRIC, 8, 4>(p, ix->nchunks, s);
    }
  }
  if (ch <= 2) return launch_rerank<DT, METRIC, 32, 2>(p, ix->nchunks, s);
  if (ch <= 4) return launch_rerank<DT, METRIC, 32, 4>(p, ix->nchunks, s);
  if (ch <= 8) return launch_rerank<DT, METRIC, 32, 8>(p, ix->nchunks, s);
  return launch_rerank<DT, METRIC, 32, 16>(p, ix->nchunks, s);
}
static cudaError_t launch_rerank_any(const fnb_index* ix, const RerankParams& p, cudaStream_t s) {
  const bool ip = ix->h.metric == FNB_METRIC_IP;
  switch (ix->h.data_type) {
    case FNB_DTYPE_FLOAT32: return ip ? rerank_gc<DT_F32, M_IP>(ix, p, s) : rerank_gc<DT_F32, M_L2>(ix, p, s);
    case FNB_DTYPE_UINT8: return ip ? rerank_gc<DT_U8, M_IP>(ix, p, s) : rerank_gc<DT_U8, M_L2>(ix, p, s);
    default: return ip ? rerank_gc<DT_I8, M_IP>(ix, p, s) : rerank_gc<DT_I8, M_L2>(ix, p, s);
  }
}

// ---- host ---------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// [rows][Dpad] bf16, K-major; box = 64 elements (128 B) x 128 rows, 128-byte swizzle; rows beyond `rows` read as 0
static bool make_map(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t Dpad) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return false;
  cuuint64_t dims[2] = {Dpad, rows};
  cuuint64_t strides[1] = {Dpad * 2};
  cuuint32_t box[2] = {BK, 128};
  cuuint32_t estr[2] = {1, 1};
  return fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static uint32_t smem_bytes(uint32_t stages, uint32_t nsplit, uint32_t Kp) {
  return 1024u + stages * 2u * nsplit * TILE_BYTES + ACC_BUFS * BN * 4u + Kp * 128u * 8u + 32u * 128u * 4u + (2u * MAX_STAGES + 2u * ACC_BUFS) * 8u + 16u;
}

}  // namespace tc

bool tensor_path_supported(const fnb_index* ix, int64_t Q, int K) {
  if (K > 122 || Q <= 0) return false;
  if (ix->h.cur_nodes < 128 || ix->h.cur_nodes >= (1ull << 31)) return false;
  if (ix->h.dim > 8192) return false;
  return tc::encode_fn() != nullptr;
}

int bruteforce_tensor(fnb_index* ix, Replica& r, const void* d_queries, int64_t Q64, int K, float* d_out_dist,
                      int32_t* d_out_label, BfRun* run) {
  using namespace tc;
  const Header& h = ix->h;
  const uint32_t N = (uint32_t)h.cur_nodes, Q = (uint32_t)Q64, dim = (uint32_t)h.dim;
  const bool is_f32 = h.data_type == FNB_DTYPE_FLOAT32;
  const bool is_ip = h.metric == FNB_METRIC_IP;
  const uint32_t nsplit = is_f32 ? 2u : 1u;
  const uint32_t Dpad = (dim + BK - 1) / BK * BK, kblocks = Dpad / BK;
  const uint32_t n_tiles = (N + BN - 1) / BN, Npad = n_tiles * BN, n_qtiles = (Q + BM - 1) / BM;
  const uint32_t Kp = (uint32_t)K + 6u <= 16u ? 16u : (uint32_t)K + 6u;  // 16: the register-list kernel
  uint32_t stages = MAX_STAGES;
  while (stages > 1 && smem_bytes(stages, nsplit, Kp) > 227u * 1024u) stages--;
  if (smem_bytes(stages, nsplit, Kp) > 227u * 1024u) return fail(FNB_ERR_UNSUPPORTED, "K=%d too large for the tensor path", K);
  // work units (128 queries x one slice of the rows): few and long — admissions into the per-thread lists are
  // ~Kp ln(n / Kp) per slice of n rows, and they are the expensive (divergent) part of the epilogue — but enough
  // to fill the SMs evenly: among the slice counts that give 1..6 units per SM pick the best wave efficiency.
  uint32_t S = 1;
  {
    const uint32_t sms = (uint32_t)r.num_sms, s_max = std::max(1u, n_tiles / 32u);
    double best = -1.0;
    for (uint32_t c = 1; c <= s_max && (uint64_t)c * n_qtiles <= 6ull * sms + n_qtiles; c++) {
      const uint64_t units = (uint64_t)c * n_qtiles;
      const double eff = (double)units / (double)(((units + sms - 1) / sms) * sms);
      if (eff > best + 0.03) {
        best = eff;
        S = c;
      }
    }
  }
  if (const char* e = getenv("FNB_BF_SPLITS")) S = std::max(1, atoi(e));
  S = std::min(S, n_tiles);
  const uint32_t tps = (n_tiles + S - 1) / S;
  S = (n_tiles + tps - 1) / tps;
  const uint32_t n_units = S * n_qtiles;

#define TC_CU(call)                                                                                       \
  do {                                                                                                    \
    cudaError_t e__ = (call);                                                                             \
    if (e__ != cudaSuccess) {                                                                             \
      for (void* ptr__ : allocs) cudaFree(ptr__);                                                         \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    }                                                                                                     \
  } while (0)
  std::vector<void*> allocs;
  auto dalloc = [&](void** p, size_t bytes) {
    cudaError_t e = cudaMalloc(p, bytes ? bytes : 16);
    if (e == cudaSuccess) allocs.push_back(*p);
    return e;
  };
  cudaStream_t s = r.stream;
  __nv_bfloat16 *xh = nullptr, *xl = nullptr, *qh = nullptr, *ql = nullptr;
  float *xnorm2 = nullptr, *xn = nullptr, *qn2 = nullptr, *cand_s = nullptr;
  uint32_t *cand_id = nullptr, *cand_cnt = nullptr, *unsafe_list = nullptr;
  unsigned int* misc = nullptr;  // [0] max|x| bits, [1] n_unsafe, [2..3] n_cand
  TC_CU(dalloc((void**)&xh, (size_t)N * Dpad * 2));
  if (nsplit == 2) TC_CU(dalloc((void**)&xl, (size_t)N * Dpad * 2));
  TC_CU(dalloc((void**)&qh, (size_t)Q * Dpad * 2));
  if (nsplit == 2) TC_CU(dalloc((void**)&ql, (size_t)Q * Dpad * 2));
  TC_CU(dalloc((void**)&xnorm2, (size_t)N * 4));
  TC_CU(dalloc((void**)&xn, (size_t)Npad * 4));
  TC_CU(dalloc((void**)&qn2, (size_t)Q * 4));
  TC_CU(dalloc((void**)&cand_s, (size_t)Q * S * Kp * 4));
  TC_CU(dalloc((void**)&cand_id, (size_t)Q * S * Kp * 4));
  TC_CU(dalloc((void**)&cand_cnt, (size_t)Q * S * 4));
  TC_CU(dalloc((void**)&unsafe_list, (size_t)Q * 4));
  TC_CU(dalloc((void**)&misc, 64));
  TC_CU(cudaMemsetAsync(misc, 0, 64, s));
  TC_CU(cudaMemsetAsync(cand_cnt, 0, (size_t)Q * S * 4, s));

  cudaEvent_t ev[5];
  for (auto& e : ev) TC_CU(cudaEventCreate(&e));
  auto cleanup = [&]() {
    for (void* ptr : allocs) cudaFree(ptr);
    for (auto& e : ev) cudaEventDestroy(e);
  };

  // ---- 1. prep ----
  TC_CU(cudaEventRecord(ev[0], s));
  const float qscale = is_ip ? -1.0f : -2.0f;
  const int pb = 256, pg = r.num_sms * 16;
  const unsigned char* vsrc = reinterpret_cast<const unsigned char*>(r.vec);
  const size_t vrow = (size_t)ix->stride * FNB_CHUNK_BYTES;
  const unsigned char* qsrc = reinterpret_cast<const unsigned char*>(d_queries);
  switch (h.data_type) {
    case FNB_DTYPE_FLOAT32:
      bf_prep_kernel<DT_F32><<<pg, pb, 0, s>>>(vsrc, vrow, N, dim, Dpad, 1.0f, xh, xl, xnorm2, misc);
      bf_prep_kernel<DT_F32><<<pg, pb, 0, s>>>(qsrc, h.data_size, Q, dim, Dpad, qscale, qh, ql, qn2, nullptr);
      break;
    case FNB_DTYPE_UINT8:
      bf_prep_kernel<DT_U8><<<pg, pb, 0, s>>>(vsrc, vrow, N, dim, Dpad, 1.0f, xh, nullptr, xnorm2, misc);
      bf_prep_kernel<DT_U8><<<pg, pb, 0, s>>>(qsrc, h.data_size, Q, dim, Dpad, qscale, qh, nullptr, qn2, nullptr);
      break;
    default:
      bf_prep_kernel<DT_I8><<<pg, pb, 0, s>>>(vsrc, vrow, N, dim, Dpad, 1.0f, xh, nullptr, xnorm2, misc);
      bf_prep_kernel<DT_I8><<<pg, pb, 0, s>>>(qsrc, h.data_size, Q, dim, Dpad, qscale, qh, nullptr, qn2, nullptr);
      break;
  }
  TC_CU(cudaGetLastError());
  bf_xn_kernel<<<(Npad + 255) / 256, 256, 0, s>>>(xnorm2, xn, N, Npad, is_ip ? 1 : 0);
  TC_CU(cudaGetLastError());
  TC_CU(cudaEventRecord(ev[1], s));

  // ---- 2. tensor-core filter ----
  CUtensorMap m_qh, m_ql, m_xh, m_xl;
  bool ok = make_map(&m_qh, qh, Q, Dpad) && make_map(&m_xh, xh, N, Dpad);
  if (nsplit == 2) ok = ok && make_map(&m_ql, ql, Q, Dpad) && make_map(&m_xl, xl, N, Dpad);
  else {
    m_ql = m_qh;
    m_xl = m_xh;
  }
  if (!ok) {
    cleanup();
    return fail(FNB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  }
  FilterParams fp;
  fp.xn = xn;
  fp.cand_s = cand_s;
  fp.cand_id = cand_id;
  fp.cand_cnt = cand_cnt;
  fp.Q = Q;
  fp.n_qtiles = n_qtiles;
  fp.n_tiles = n_tiles;
  fp.S = S;
  fp.tiles_per_split = tps;
  fp.kblocks = kblocks;
  fp.Kp = Kp;
  fp.stages = stages;
  fp.n_units = n_units;
  const uint32_t smem = smem_bytes(stages, nsplit, Kp);
  const uint32_t grid = std::min<uint32_t>(n_units, (uint32_t)r.num_sms);
  const bool reglist = Kp == 16 && !getenv("FNB_BF_SMEM_LIST");
#define TC_LAUNCH(NS, KR)                                                                                       \
  do {                                                                                                          \
    TC_CU(cudaFuncSetAttribute(bf_tc_kernel<NS, KR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    bf_tc_kernel<NS, KR><<<grid, THREADS, smem, s>>>(m_qh, m_ql, m_xh, m_xl, fp);                               \
  } while (0)
  if (nsplit == 2) {
    if (reglist) TC_LAUNCH(2, 16);
    else TC_LAUNCH(2, 0);
  } else {
    if (reglist) TC_LAUNCH(1, 16);
    else TC_LAUNCH(1, 0);
  }
#undef TC_LAUNCH
  TC_CU(cudaGetLastError());
  TC_CU(cudaEventRecord(ev[2], s));

  // ---- 3. exact re-rank ----
  RerankParams rp;
  memset(&rp, 0, sizeof(rp));
  rp.sp.vec = r.vec;
  rp.sp.stride = ix->stride;
  rp.sp.nchunks = ix->nchunks;
  rp.sp.queries = d_queries;
  rp.sp.dim = dim;
  rp.sp.query_vec_ok = ((h.data_size % FNB_CHUNK_BYTES) == 0 && ((uintptr_t)d_queries & 15u) == 0) ? 1u : 0u;
  rp.sp.lines_per_row = (ix->stride * FNB_CHUNK_BYTES + 127u) / 128u;
  rp.labels = r.labels;
  rp.cand_s = cand_s;
  rp.cand_id = cand_id;
  rp.cand_cnt = cand_cnt;
  rp.qn2 = qn2;
  rp.xmax_bits = misc;
  rp.out_dist = d_out_dist;
  rp.out_label = d_out_label;
  rp.n_unsafe = misc + 1;
  rp.unsafe_list = unsafe_list;
  rp.n_cand = reinterpret_cast<unsigned long long*>(misc + 2);
  rp.Q = Q;
  rp.K = (uint32_t)K;
  rp.Kcap = ((uint32_t)K + 31u) & ~31u;
  rp.S = S;
  rp.Kp = Kp;
  rp.is_ip = is_ip ? 1 : 0;
  rp.qscale = is_ip ? 1.0f : 2.0f;
  const double maxabs = h.data_type == FNB_DTYPE_UINT8 ? 255.0 : 128.0;
  if (!is_f32 && (double)dim * maxabs * maxabs * 2.0 < 16777216.0) {
    rp.c_rel = 0.f;  // every product, partial sum and score is an integer below 2^24: the filter is exact
    rp.c_abs = 0.f;
  } else if (!is_f32) {
    rp.c_rel = (float)((3.0 * Dpad / 16.0 + 2.0) * std::ldexp(1.0, -21));
    rp.c_abs = (float)std::ldexp(1.0, -22);
  } else {
    rp.c_rel = (float)(3.0 * std::ldexp(1.0, -18) + (3.0 * Dpad / 16.0) * std::ldexp(1.0, -21));
    rp.c_abs = (float)((double)dim * std::ldexp(1.0, -23));
  }
  TC_CU(launch_rerank_any(ix, rp, s));
  TC_CU(cudaEventRecord(ev[3], s));

  // ---- 4. exact re-scan of the queries whose candidate set could not be proven complete ----
  unsigned int host_misc[4] = {0, 0, 0, 0};
  TC_CU(cudaMemcpyAsync(host_misc, misc, 16, cudaMemcpyDeviceToHost, s));
  TC_CU(cudaStreamSynchronize(s));
  const uint32_t n_unsafe = host_misc[1];
  if (n_unsafe > 0) {
    BfParams bp;
    memset(&bp, 0, sizeof(bp));
    bp.vec = r.vec;
    bp.labels = r.labels;
    bp.queries = d_queries;
    bp.qmap = unsafe_list;
    bp.out_dist = d_out_dist;
    bp.out_label = d_out_label;
    bp.N = N;
    bp.dim = dim;
    bp.nchunks = ix->nchunks;
    bp.stride = ix->stride;
    bp.Q = n_unsafe;
    bp.K = (uint32_t)K;
    bp.Kcap = ((uint32_t)K + 31u) & ~31u;
    bp.query_vec_ok = rp.sp.query_vec_ok;
    TC_CU(launch_exact_scan(ix, bp, s));
  }
  TC_CU(cudaEventRecord(ev[4], s));
  TC_CU(cudaEventSynchronize(ev[4]));
  if (run) {
    cudaEventElapsedTime(&run->prep_ms, ev[0], ev[1]);
    cudaEventElapsedTime(&run->gemm_ms, ev[1], ev[2]);
    cudaEventElapsedTime(&run->rerank_ms, ev[2], ev[3]);
    cudaEventElapsedTime(&run->rescan_ms, ev[3], ev[4]);
    run->n_unsafe = n_unsafe;
    unsigned long long nc;
    memcpy(&nc, host_misc + 2, 8);
    run->n_candidates = (int64_t)nc;
    run->gemm_flops = 2.0 * (double)n_qtiles * BM * (double)n_tiles * BN * (double)Dpad * (nsplit == 2 ? 3.0 : 1.0);
  }
  cleanup();
  return FNB_OK;
#undef TC_CU
}

}  // namespace fnb
