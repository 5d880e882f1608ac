// Measurement tool (not part of the library): the HBM throughput a B200 delivers for RANDOM ROW GATHERS — the access
// pattern of the traversal kernel — as opposed to the streaming copy MEASURED_PEAKS.json reports.
//
//   tools/gather_peak.bin [--rows-mb 4096] [--reps 7]
//
// For each row shape (bytes used / pitch) of the BASELINE configs it times two kernels over uniformly random row ids of
// an array much larger than L2:
//   load      G lanes x 16 B per row (one 128-byte line per instruction and row, like fnb_search_kernel), U x CH loads
//             in flight per lane, results xor-ed into a sink;
//   prefetch  prefetch.global.L2 of every line of the row, nothing returned: no register or scoreboard limit, the
//             memory system's own ceiling for this pattern.
// One JSON line per measurement: {"row_bytes", "pitch", "mode", "gbs" (used bytes / time), "gbs_pitch", "rows_per_s"}.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e = (x);                                                                 \
    if (e != cudaSuccess) {                                                              \
      fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
      exit(1);                                                                           \
    }                                                                                    \
  } while (0)

__device__ __forceinline__ uint4 ldg_stream_if(const uint4* p, bool pred) {
  uint4 r = make_uint4(0, 0, 0, 0);
  asm volatile(
      "{\n\t.reg .pred pp;\n\tsetp.ne.b32 pp, %5, 0;\n\t"
      "@pp ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
      : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w)
      : "l"(p), "r"((int)pred));
  return r;
}

// ids: [n_ids] random row numbers; every warp takes 32-id groups round-robin
template <int G, int CH, int U>
__global__ void __launch_bounds__(128) gather_load(const uint4* __restrict__ vec, uint32_t pitch_chunks, uint32_t nchunks,
                                                   const uint32_t* __restrict__ ids, uint32_t n_groups,
                                                   uint32_t* __restrict__ sink) {
  constexpr int RPI = 32 / G;
  const int lane = threadIdx.x & 31, g = lane / G, pos = lane % G;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  uint32_t acc = 0;
  for (uint32_t grp = warp; grp < n_groups; grp += nwarps) {
    const uint32_t my = ids[(size_t)grp * 32 + lane];
#pragma unroll 1
    for (int r0 = 0; r0 < 32; r0 += RPI * U) {
      uint4 x[U][CH];
#pragma unroll
      for (int u = 0; u < U; u++) {
        const uint32_t rid = __shfl_sync(0xffffffffu, my, (r0 + u * RPI + g) & 31);
        const uint4* row = vec + (size_t)rid * pitch_chunks + pos;
#pragma unroll
        for (int k = 0; k < CH; k++) x[u][k] = ldg_stream_if(row + k * G, (uint32_t)(k * G + pos) < nchunks);
      }
#pragma unroll
      for (int u = 0; u < U; u++)
#pragma unroll
        for (int k = 0; k < CH; k++) acc ^= x[u][k].x ^ x[u][k].y ^ x[u][k].z ^ x[u][k].w;
    }
  }
  if (acc == 0x12345678u) sink[0] = acc;  // keeps the loads alive
}

__global__ void __launch_bounds__(128) gather_prefetch(const uint4* __restrict__ vec, uint32_t pitch_chunks, uint32_t lines_used,
                                                       const uint32_t* __restrict__ ids, uint32_t n_groups) {
  const int lane = threadIdx.x & 31;
  const uint32_t warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t grp = warp; grp < n_groups; grp += nwarps) {
    const uint32_t my = ids[(size_t)grp * 32 + lane];
    // 32 rows x lines_used lines, one line per lane and instruction
    const uint32_t total = 32u * lines_used;
    for (uint32_t t = lane; t < total; t += 32) {
      const uint32_t rid = __shfl_sync(0xffffffffu, my, (t / lines_used) & 31);
      const uint4* p = vec + (size_t)rid * pitch_chunks + (t % lines_used) * 8u;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
  }
}

struct Shape {
  const char* name;
  uint32_t row_bytes, pitch;
};

template <int G, int CH, int U>
static float time_load(const uint4* vec, const Shape& s, const uint32_t* ids, uint32_t n_groups, uint32_t* sink, int blocks,
                       int reps) {
  cudaEvent_t a, b;
  CK(cudaEventCreate(&a));
  CK(cudaEventCreate(&b));
  std::vector<float> ts;
  for (int i = 0; i < reps + 2; i++) {
    CK(cudaEventRecord(a));
    gather_load<G, CH, U><<<blocks, 128>>>(vec, s.pitch / 16, (s.row_bytes + 15) / 16, ids, n_groups, sink);
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    if (i >= 2) ts.push_back(ms);
  }
  std::sort(ts.begin(), ts.end());
  return ts[ts.size() / 2];
}

int main(int argc, char** argv) {
  size_t rows_mb = 4096;
  int reps = 7;
  for (int i = 1; i + 1 < argc; i += 2) {
    if (!strcmp(argv[i], "--rows-mb")) rows_mb = strtoull(argv[i + 1], 0, 10);
    if (!strcmp(argv[i], "--reps")) reps = atoi(argv[i + 1]);
  }
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const size_t bytes = rows_mb << 20;
  uint4* vec;
  CK(cudaMalloc(&vec, bytes));
  CK(cudaMemset(vec, 1, bytes));
  const uint32_t n_ids = 32u * 262144u;  // 8.4 M row fetches per launch
  uint32_t *d_ids, *sink;
  CK(cudaMalloc(&d_ids, (size_t)n_ids * 4));
  CK(cudaMalloc(&sink, 4));
  std::vector<uint32_t> h(n_ids);
  const Shape shapes[] = {{"u8 D=128", 128, 128},        {"f32 D=96", 384, 384},   {"f32 D=100", 400, 512},
                          {"f32 D=128", 512, 512},       {"f32 D=256", 1024, 1024}, {"f32 D=960", 3840, 3840}};
  for (const Shape& s : shapes) {
    const uint64_t n_rows = bytes / s.pitch;
    uint64_t x = 0x9E3779B97F4A7C15ull;
    for (uint32_t i = 0; i < n_ids; i++) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      h[i] = (uint32_t)(x % n_rows);
    }
    CK(cudaMemcpy(d_ids, h.data(), (size_t)n_ids * 4, cudaMemcpyHostToDevice));
    const uint32_t n_groups = n_ids / 32;
    const double used = (double)n_ids * s.row_bytes, pitched = (double)n_ids * s.pitch;
    for (int occ : {4, 8, 12, 16}) {
      const int blocks = prop.multiProcessorCount * occ;
      float ms;
      const uint32_t nch = (s.row_bytes + 15) / 16;
      if (nch <= 8) ms = time_load<8, 1, 8>(vec, s, d_ids, n_groups, sink, blocks, reps);
      else if (nch <= 32) ms = time_load<8, 4, 4>(vec, s, d_ids, n_groups, sink, blocks, reps);
      else if (nch <= 64) ms = time_load<32, 2, 4>(vec, s, d_ids, n_groups, sink, blocks, reps);
      else ms = time_load<32, 8, 2>(vec, s, d_ids, n_groups, sink, blocks, reps);
      printf("{\"shape\": \"%s\", \"row_bytes\": %u, \"pitch\": %u, \"footprint_mb\": %zu, \"mode\": \"load\", \"ctas_per_sm\": %d, "
             "\"ms\": %.4f, \"gbs\": %.1f, \"gbs_pitch\": %.1f, \"rows_per_s\": %.3e}\n",
             s.name, s.row_bytes, s.pitch, rows_mb, occ, ms, used / ms / 1e6, pitched / ms / 1e6, n_ids / (ms * 1e-3));
    }
    {
      cudaEvent_t a, b;
      CK(cudaEventCreate(&a));
      CK(cudaEventCreate(&b));
      std::vector<float> ts;
      const int blocks = prop.multiProcessorCount * 16;
      for (int i = 0; i < reps + 2; i++) {
        CK(cudaEventRecord(a));
        gather_prefetch<<<blocks, 128>>>(vec, s.pitch / 16, (s.row_bytes + 127) / 128, d_ids, n_groups);
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms;
        CK(cudaEventElapsedTime(&ms, a, b));
        if (i >= 2) ts.push_back(ms);
      }
      std::sort(ts.begin(), ts.end());
      const float ms = ts[ts.size() / 2];
      printf("{\"shape\": \"%s\", \"row_bytes\": %u, \"pitch\": %u, \"footprint_mb\": %zu, \"mode\": \"prefetch\", \"ctas_per_sm\": 16, "
             "\"ms\": %.4f, \"gbs\": %.1f, \"gbs_pitch\": %.1f, \"rows_per_s\": %.3e}\n",
             s.name, s.row_bytes, s.pitch, rows_mb, ms, used / ms / 1e6, pitched / ms / 1e6, n_ids / (ms * 1e-3));
    }
    fflush(stdout);
  }
  return 0;
}
