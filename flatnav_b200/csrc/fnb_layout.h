// Layout rules shared by the host loader, the CUDA kernels and (by restatement) the oracle's
// "lanes" summation order.  Host + device.
//
// HBM layout of a loaded index (replaces the reference's AoS node blob
// [vector | M x u32 links | i32 label], include/flatnav/index/Index.h:555-573, which is not 16-byte
// aligned — 644 B per node at D=128 f32, M=32):
//
//   vec    [N][stride]  16-byte chunks; a row is the node's vector, zero-padded to `stride` chunks.
//                       Rows never straddle a 128-byte line more than they must: stride = nchunks rounded up
//                       to a power of two for rows up to 128 B, to a multiple of 8 chunks (128 B) above.  The
//                       padding is never fetched (loads are per 16-byte chunk, DRAM sectors are 32 B); what it
//                       buys is that an 8-lane x 16 B row segment is ONE line request instead of two — measured
//                       on 1.2M x 100 f32 (400-byte rows): +23 % QPS at ef >= 128 for 512-byte vs 416-byte pitch.
//   adj    [N][M]       uint32 node ids (self-loops mark unused slots, Index.h:270)
//   label  [N]          int32
//
// Distance arithmetic order ("lanes"): a row is processed by G lanes; lane p accumulates chunks
// p, p+G, p+2G, ... element by element with fused multiply-add (float32) or dp4a (int8/uint8), and the
// G partial sums are combined by an xor butterfly with offsets G/2 ... 1.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FNB_HD __host__ __device__ __forceinline__
#else
#define FNB_HD inline
#endif

#define FNB_CHUNK_BYTES 16u

FNB_HD uint32_t fnb_nchunks(uint64_t data_size_bytes) { return (uint32_t)((data_size_bytes + FNB_CHUNK_BYTES - 1) / FNB_CHUNK_BYTES); }
FNB_HD uint32_t fnb_stride_chunks(uint32_t nchunks) {
  if (nchunks > 8u) return (nchunks + 7u) & ~7u;
  uint32_t s = 1u;
  while (s < nchunks) s <<= 1;
  return s;
}
// lanes cooperating on one row: 8 lanes x 16 B = one 128-byte line per load instruction and 4 rows per
// warp-wide load; rows longer than 512 B use the whole warp.
// Integer rows of at most 128 B use 4 lanes (two 16-byte chunks per lane at D=128 uint8): their sums are exact, so
// the summation order is free, and the shorter butterfly + 8 rows per load instruction cut the per-row instruction
// count where the kernel is issue-bound rather than HBM-bound.
FNB_HD int fnb_lanes_per_row(uint32_t nchunks, bool integer_data = false) {
  if (integer_data && nchunks <= 8u) return 4;
  return nchunks <= 32u ? 8 : 32;
}
FNB_HD int fnb_chunks_per_lane(uint32_t nchunks, int g) { return (int)((nchunks + (uint32_t)g - 1u) / (uint32_t)g); }
#define FNB_MAX_CHUNKS 512u /* largest row the kernels are instantiated for: 8 KB (D=2048 float32) */
