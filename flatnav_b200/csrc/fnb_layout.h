// Layout rules shared by the host loader, the CUDA kernels and (by restatement) the oracle's
// "lanes" summation order.  Host + device.
//
// HBM layout of a loaded index (replaces the reference's AoS node blob
// [vector | M x u32 links | i32 label], include/flatnav/index/Index.h:555-573, which is not 16-byte
// aligned — 644 B per node at D=128 f32, M=32):
//
//   vec    [N][stride]  16-byte chunks; a row is the node's vector, zero-padded to `stride` chunks,
//                       stride = nchunks rounded up to an even count so every row starts on a 32-byte
//                       sector boundary
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
FNB_HD uint32_t fnb_stride_chunks(uint32_t nchunks) { return (nchunks + 1u) & ~1u; }
// lanes cooperating on one row: 8 lanes x 16 B = one 128-byte line per load instruction and 4 rows per
// warp-wide load; rows longer than 512 B use the whole warp.
FNB_HD int fnb_lanes_per_row(uint32_t nchunks) { return nchunks <= 32u ? 8 : 32; }
FNB_HD int fnb_chunks_per_lane(uint32_t nchunks) {
  int g = fnb_lanes_per_row(nchunks);
  return (int)((nchunks + (uint32_t)g - 1u) / (uint32_t)g);
}
#define FNB_MAX_CHUNKS 512u /* largest row the kernels are instantiated for: 8 KB (D=2048 float32) */
