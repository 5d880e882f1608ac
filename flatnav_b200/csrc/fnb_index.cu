// Host side of libflatnav_b200.so: index loader (reference cereal layout -> HBM arrays), search
// dispatch and the C ABI declared in include/flatnav_b200.h.  No CPU fallback anywhere: every compute
// entry point requires a CUDA device.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/flatnav_b200.h"
#include "fnb_internal.h"

namespace fnb {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

#define CU(call)                                                                                   \
  do {                                                                                             \
    cudaError_t e__ = (call);                                                                      \
    if (e__ != cudaSuccess)                                                                        \
      return fail(FNB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// ---- layout kernels -------------------------------------------------------------------------------
// One warp per node; byte-wise so that any data_size / node_size (including unaligned ones) is handled.
// Reference node layout: [vector (data_size B) | M x u32 links | i32 label]  (Index.h:555-573).
__global__ void deinterleave_kernel(const unsigned char* __restrict__ blob, uint64_t n_nodes, uint64_t node_size,
                                    uint32_t data_size, uint32_t M, uint32_t stride_bytes,
                                    unsigned char* __restrict__ vec, uint32_t* __restrict__ adj,
                                    int32_t* __restrict__ labels) {
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (uint64_t n = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; n < n_nodes; n += warps) {
    const unsigned char* src = blob + n * node_size;
    unsigned char* dst = vec + n * stride_bytes;
    for (uint32_t b = lane; b < stride_bytes; b += 32) dst[b] = b < data_size ? src[b] : (unsigned char)0;
    const unsigned char* ls = src + data_size;
    for (uint32_t j = lane; j <= M; j += 32) {
      uint32_t v = (uint32_t)ls[4 * j] | ((uint32_t)ls[4 * j + 1] << 8) | ((uint32_t)ls[4 * j + 2] << 16) |
                   ((uint32_t)ls[4 * j + 3] << 24);
      if (j < M)
        adj[n * M + j] = v;
      else
        labels[n] = (int32_t)v;
    }
  }
}

__global__ void interleave_kernel(unsigned char* __restrict__ blob, uint64_t n_nodes, uint64_t node_size,
                                  uint32_t data_size, uint32_t M, uint32_t stride_bytes,
                                  const unsigned char* __restrict__ vec, const uint32_t* __restrict__ adj,
                                  const int32_t* __restrict__ labels) {
  const uint64_t warps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  const int lane = threadIdx.x & 31;
  for (uint64_t n = (((uint64_t)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; n < n_nodes; n += warps) {
    unsigned char* dst = blob + n * node_size;
    const unsigned char* src = vec + n * stride_bytes;
    for (uint32_t b = lane; b < data_size; b += 32) dst[b] = src[b];
    unsigned char* ls = dst + data_size;
    for (uint32_t j = lane; j <= M; j += 32) {
      uint32_t v = j < M ? adj[n * M + j] : (uint32_t)labels[n];
      ls[4 * j] = (unsigned char)v;
      ls[4 * j + 1] = (unsigned char)(v >> 8);
      ls[4 * j + 2] = (unsigned char)(v >> 16);
      ls[4 * j + 3] = (unsigned char)(v >> 24);
    }
  }
}

// every link of a live node must point at a live node (Appendix A.6 of SURVEY.md)
__global__ void validate_links_kernel(const uint32_t* __restrict__ adj, uint64_t n_links, uint32_t n_nodes,
                                      unsigned int* __restrict__ bad) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_links; i += stride)
    if (adj[i] >= n_nodes) atomicAdd(bad, 1u);
}

// ---- header ------------------------------------------------------------------------------------------
// cereal BinaryOutputArchive of Index::serialize (Index.h:134-141): int32 data_type, u64 M, u64 data_size,
// u64 node_size, u64 max_node_count, u64 cur_num_nodes, then the distance object's (u64 dimension,
// u64 data_size) (SquaredL2Distance.h:54-57 / InnerProductDistance.h:53-56); little endian, no padding.
static int parse_header(const unsigned char* p, size_t nbytes, int metric, int expect_dtype, Header* h) {
  if (nbytes < FNB_HEADER_BYTES) return fail(FNB_ERR_FORMAT, "index file shorter than the 60-byte header");
  if (metric != FNB_METRIC_L2 && metric != FNB_METRIC_IP) return fail(FNB_ERR_INVALID_ARG, "unknown metric %d", metric);
  int32_t dt;
  uint64_t v[7];
  memcpy(&dt, p, 4);
  memcpy(v, p + 4, 56);
  h->data_type = dt;
  h->metric = metric;
  h->M = v[0];
  h->data_size = v[1];
  h->node_size = v[2];
  h->max_nodes = v[3];
  h->cur_nodes = v[4];
  h->dim = v[5];
  uint64_t es = dt == FNB_DTYPE_FLOAT32 ? 4 : (dt == FNB_DTYPE_UINT8 || dt == FNB_DTYPE_INT8) ? 1 : 0;
  if (es == 0) return fail(FNB_ERR_FORMAT, "unsupported data_type %d in index header", dt);
  if (expect_dtype != FNB_DTYPE_ANY && expect_dtype != dt)
    return fail(FNB_ERR_FORMAT, "index holds data_type %d but %d was requested", dt, expect_dtype);
  // bound every field BEFORE multiplying: a crafted header must not wrap the size arithmetic below (or the kernels'
  // 32-bit copies of M / dim)
  if (h->M == 0 || h->dim == 0) return fail(FNB_ERR_FORMAT, "M and dim must be positive");
  if (h->M > 65536) return fail(FNB_ERR_UNSUPPORTED, "max_edges_per_node %llu exceeds 65536", (unsigned long long)h->M);
  if (h->dim > (uint64_t)FNB_MAX_CHUNKS * FNB_CHUNK_BYTES)
    return fail(FNB_ERR_UNSUPPORTED, "dimension %llu exceeds what the kernels are built for", (unsigned long long)h->dim);
  if (h->max_nodes >= (1ull << 31)) return fail(FNB_ERR_UNSUPPORTED, "more than 2^31 nodes");
  if (v[6] != h->data_size || h->dim * es != h->data_size)
    return fail(FNB_ERR_FORMAT, "header inconsistent: dim=%llu data_size=%llu/%llu", (unsigned long long)h->dim,
                (unsigned long long)h->data_size, (unsigned long long)v[6]);
  if (h->node_size != h->data_size + 4 * h->M + 4)
    return fail(FNB_ERR_FORMAT, "header inconsistent: node_size %llu != data_size + 4M + 4",
                (unsigned long long)h->node_size);
  if (h->cur_nodes > h->max_nodes) return fail(FNB_ERR_FORMAT, "cur_num_nodes > max_node_count");
  // node_size <= 8 KB + 256 KB + 4 and max_nodes < 2^31: the product fits 64 bits
  if (nbytes < FNB_HEADER_BYTES + h->node_size * h->max_nodes)
    return fail(FNB_ERR_FORMAT, "index file truncated: %zu bytes, need %llu", nbytes,
                (unsigned long long)(FNB_HEADER_BYTES + h->node_size * h->max_nodes));
  if (fnb_nchunks(h->data_size) > FNB_MAX_CHUNKS)
    return fail(FNB_ERR_UNSUPPORTED, "vector of %llu bytes exceeds the %u-byte limit of the kernels",
                (unsigned long long)h->data_size, FNB_MAX_CHUNKS * FNB_CHUNK_BYTES);
  return FNB_OK;
}

// Row pitch of the vector array in 16-byte chunks.  FNB_ROW_ALIGN (chunks, power of two) overrides the layout rule
// of fnb_layout.h for experiments.
static uint32_t row_stride_chunks(uint32_t nchunks) {
  if (const char* e = getenv("FNB_ROW_ALIGN")) {
    const uint32_t a = (uint32_t)atoi(e);
    if (a >= 1 && a <= 64 && (a & (a - 1)) == 0) return (nchunks + a - 1) & ~(a - 1);
  }
  return fnb_stride_chunks(nchunks);
}

static int upload_replica(const Header& h, const unsigned char* blob, int device, Replica* r) {
  CU(cudaSetDevice(device));
  r->device = device;
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(FNB_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device,
                prop.major, prop.minor);
  r->num_sms = prop.multiProcessorCount;
  const uint32_t nchunks = fnb_nchunks(h.data_size);
  const uint32_t stride = row_stride_chunks(nchunks);
  const uint64_t n = h.cur_nodes ? h.cur_nodes : 1;
  const uint64_t blob_bytes = h.node_size * h.cur_nodes;
  struct Blob {  // the device copy of the file image: freed on every way out
    unsigned char* p = nullptr;
    ~Blob() { cudaFree(p); }
  } blob_guard;
  unsigned char*& d_blob = blob_guard.p;
  CU(cudaMalloc(&r->vec, n * stride * FNB_CHUNK_BYTES));
  CU(cudaMalloc(&r->adj, n * h.M * 4));
  CU(cudaMalloc(&r->labels, n * 4));
  CU(cudaMalloc(&r->counter, 128));
  r->totals = reinterpret_cast<unsigned long long*>(r->counter + 16);
  CU(cudaStreamCreateWithFlags(&r->stream, cudaStreamNonBlocking));
  r->pool = new LanePool();
  if (const char* e = getenv("FNB_MAX_LANES")) r->pool->max_lanes = std::max(1, atoi(e));
  CU(cudaMalloc(&r->pool->ring, FNB_RING_SLOTS * 64));
  CU(cudaMemset(r->pool->ring, 0, FNB_RING_SLOTS * 64));
  CU(cudaHostAlloc((void**)&r->pool->h_done_seq, 64, cudaHostAllocMapped));
  *r->pool->h_done_seq = 0u;
  CU(cudaHostGetDevicePointer((void**)&r->pool->d_done_seq, (void*)r->pool->h_done_seq, 0));
  r->device_bytes = n * stride * FNB_CHUNK_BYTES + n * h.M * 4 + n * 4;
  r->capacity = n;
  if (h.cur_nodes) {
    CU(cudaMalloc(&d_blob, blob_bytes));
    CU(cudaMemcpyAsync(d_blob, blob, blob_bytes, cudaMemcpyHostToDevice, r->stream));
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((h.cur_nodes * 32 + threads - 1) / threads, (uint64_t)r->num_sms * 32);
    deinterleave_kernel<<<blocks, threads, 0, r->stream>>>(d_blob, h.cur_nodes, h.node_size, (uint32_t)h.data_size, (uint32_t)h.M,
                                             stride * FNB_CHUNK_BYTES, reinterpret_cast<unsigned char*>(r->vec),
                                             r->adj, r->labels);
    CU(cudaGetLastError());
    CU(cudaMemsetAsync(r->counter, 0, 64, r->stream));
    validate_links_kernel<<<r->num_sms * 8, 256, 0, r->stream>>>(r->adj, h.cur_nodes * h.M, (uint32_t)h.cur_nodes, r->counter);
    CU(cudaGetLastError());
    unsigned int bad = 0;
    CU(cudaMemcpyAsync(&bad, r->counter, 4, cudaMemcpyDeviceToHost, r->stream));
    CU(cudaStreamSynchronize(r->stream));
    if (bad) return fail(FNB_ERR_FORMAT, "%u links point outside [0, cur_num_nodes)", bad);
  }
  return FNB_OK;
}

static void free_lane(Lane* l) {
  if (!l) return;
  cudaFree(l->counter);
  if (l->h_totals) cudaFreeHost(l->h_totals);
  cudaFree(l->ws);
  if (l->h_pinned) cudaFreeHost(l->h_pinned);
  if (l->stream) cudaStreamDestroy(l->stream);
  if (l->copy_stream) cudaStreamDestroy(l->copy_stream);
  for (auto& e : l->ev)
    if (e) cudaEventDestroy(e);
  if (l->ev_feed) cudaEventDestroy(l->ev_feed);
  delete l;
}

static void free_replica(Replica* r) {
  if (r->device < 0) return;
  cudaSetDevice(r->device);
  cudaDeviceSynchronize();
  cudaFree(r->vec);
  cudaFree(r->adj);
  cudaFree(r->labels);
  cudaFree(r->counter);
  if (r->stream) cudaStreamDestroy(r->stream);
  if (r->pool) {
    for (Lane* l : r->pool->all) free_lane(l);
    cudaFree(r->pool->ring);
    if (r->pool->h_done_seq) cudaFreeHost((void*)r->pool->h_done_seq);
    delete r->pool;
    r->pool = nullptr;
  }
}

// ---- search lanes ------------------------------------------------------------------------------------
// The calling thread's current device must be r.device.
static int new_lane(Lane** out) {
  Lane* l = new Lane();
  cudaError_t e = cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaMalloc(&l->counter, 128);
  if (e == cudaSuccess) e = cudaMemset(l->counter, 0, 128);
  if (e == cudaSuccess) e = cudaMallocHost(&l->h_totals, 64 + (FNB_FEED_CHUNKS + 1) * 4);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&l->copy_stream, cudaStreamNonBlocking);
  for (int i = 0; i < 4 && e == cudaSuccess; i++) e = cudaEventCreate(&l->ev[i]);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&l->ev_feed, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    free_lane(l);
    return fail(FNB_ERR_CUDA, "cannot create a search lane: %s", cudaGetErrorString(e));
  }
  l->totals = reinterpret_cast<unsigned long long*>(reinterpret_cast<unsigned char*>(l->counter) + FNB_SLOT_TOTALS);
  l->q_ready = l->counter + 16;  // second half of the block
  if (cudaHostGetDevicePointer((void**)&l->h_totals_dev, l->h_totals, 0) != cudaSuccess) {
    free_lane(l);
    return fail(FNB_ERR_CUDA, "cannot map the lane's pinned block");
  }
  l->h_flag = reinterpret_cast<volatile uint32_t*>(l->h_totals + 7);
  *l->h_flag = 0u;
  l->h_flag_dev = reinterpret_cast<unsigned int*>(l->h_totals_dev + 7);
  l->h_marks = reinterpret_cast<uint32_t*>(l->h_totals + 8);
  l->h_marks[FNB_FEED_CHUNKS] = 0xffffffffu;  // "everything is there": releases a launch whose feeding was cut short
  *out = l;
  return FNB_OK;
}

static int acquire_lane(Replica& r, Lane** out) {
  LanePool& P = *r.pool;
  std::unique_lock<std::mutex> lk(P.mu);
  for (;;) {
    if (!P.idle.empty()) {
      *out = P.idle.back();
      P.idle.pop_back();
      return FNB_OK;
    }
    if ((int)P.all.size() < P.max_lanes) {
      Lane* l = nullptr;
      const int rc = new_lane(&l);
      if (rc != FNB_OK) return rc;
      P.all.push_back(l);
      *out = l;
      return FNB_OK;
    }
    P.cv.wait(lk);
  }
}

static void release_lane(Replica& r, Lane* l) {
  LanePool& P = *r.pool;
  {
    std::lock_guard<std::mutex> lk(P.mu);
    P.idle.push_back(l);
  }
  P.cv.notify_one();
}

// A lane held for the length of one fnb_search call.  Whatever way the call ends, nothing it enqueued may still be
// running into the caller's buffers when it returns: the destructor waits for the stream before handing the lane back.
struct LaneHold {
  Replica* r = nullptr;
  Lane* l = nullptr;
  bool feeding = false;  // a host-fed launch has not been fed to the end yet
  bool drained = false;  // the lane's launch was seen to finish (its flag in pinned memory): nothing is left on the stream
  LaneHold() = default;
  LaneHold(const LaneHold&) = delete;
  LaneHold& operator=(const LaneHold&) = delete;
  LaneHold(LaneHold&& o) noexcept : r(o.r), l(o.l), feeding(o.feeding), drained(o.drained) { o.l = nullptr; }
  ~LaneHold() {
    if (!l) return;
    if (!drained) {
      cudaSetDevice(r->device);
      // leaving early (an error after the launch): let the waiting warps through, the call fails anyway
      if (feeding) cudaMemcpyAsync(l->q_ready, l->h_marks + FNB_FEED_CHUNKS, 4, cudaMemcpyHostToDevice, l->copy_stream);
      cudaStreamSynchronize(l->stream);
    }
    release_lane(*r, l);
  }
};

void quiesce(fnb_index* ix) {
  int prev = 0;
  cudaGetDevice(&prev);
  for (Replica& r : ix->replicas) {
    if (r.device < 0) continue;
    cudaSetDevice(r.device);
    cudaDeviceSynchronize();
  }
  cudaSetDevice(prev);
}

static int build_index(const unsigned char* file, size_t nbytes, int metric, int expect_dtype, const int* device_ids,
                       int n_devices, fnb_index** out) {
  if (!out) return fail(FNB_ERR_INVALID_ARG, "out is NULL");
  *out = nullptr;
  Header h;
  int rc = parse_header(file, nbytes, metric, expect_dtype, &h);
  if (rc != FNB_OK) return rc;
  int ndev_avail = 0;
  if (cudaGetDeviceCount(&ndev_avail) != cudaSuccess || ndev_avail == 0)
    return fail(FNB_ERR_CUDA, "no CUDA device available; flatnav_b200 has no CPU path");
  std::vector<int> devs;
  if (device_ids == nullptr || n_devices <= 0) {
    int cur = 0;
    CU(cudaGetDevice(&cur));
    devs.push_back(cur);
  } else {
    if (n_devices > 16) return fail(FNB_ERR_INVALID_ARG, "at most 16 replicas");
    for (int i = 0; i < n_devices; i++) {
      if (device_ids[i] < 0 || device_ids[i] >= ndev_avail)
        return fail(FNB_ERR_INVALID_ARG, "device id %d out of range (have %d)", device_ids[i], ndev_avail);
      devs.push_back(device_ids[i]);
    }
  }
  int prev = 0;
  cudaGetDevice(&prev);
  fnb_index* ix = new fnb_index();
  ix->h = h;
  ix->nchunks = fnb_nchunks(h.data_size);
  ix->stride = row_stride_chunks(ix->nchunks);
  ix->G = fnb_lanes_per_row(ix->nchunks, h.data_type != FNB_DTYPE_FLOAT32 && !getenv("FNB_NO_G4"));
  ix->replicas.resize(devs.size());
  for (size_t i = 0; i < devs.size(); i++) {
    rc = upload_replica(h, file + FNB_HEADER_BYTES, devs[i], &ix->replicas[i]);
    if (rc != FNB_OK) {
      std::string keep = g_last_error;
      for (auto& r : ix->replicas) free_replica(&r);
      delete ix;
      cudaSetDevice(prev);
      g_last_error = keep;
      return rc;
    }
  }
  cudaSetDevice(prev);
  *out = ix;
  return FNB_OK;
}

// ---- search dispatch ---------------------------------------------------------------------------------
// launch_q: queries per kernel launch (the occupancy plan depends on it); <= 0 means Q
int plan_search(const fnb_index* ix, int64_t Q, int K, int ef, int ninit, SearchParams* p, int64_t launch_q,
                uint64_t n_nodes, bool allow_latency_variant) {
  if (Q < 0) return fail(FNB_ERR_INVALID_ARG, "negative query count");
  if (K <= 0) return fail(FNB_ERR_INVALID_ARG, "K must be positive");
  if (ninit <= 0) return fail(FNB_ERR_INVALID_ARG, "num_initializations must be greater than 0.");
  if (Q >= (1ll << 31)) return fail(FNB_ERR_UNSUPPORTED, "more than 2^31 queries in one call");
  const Header& h = ix->h;
  const uint64_t cur_nodes = n_nodes ? n_nodes : h.cur_nodes;
  memset(p, 0, sizeof(*p));
  p->N = (uint32_t)cur_nodes;
  p->M = (uint32_t)h.M;
  p->dim = (uint32_t)h.dim;
  p->nchunks = ix->nchunks;
  p->stride = ix->stride;
  p->Q = (uint32_t)Q;
  p->K = (uint32_t)K;
  p->B = (uint32_t)std::max(ef, K);  // Index.h:392
  p->Bcap = (p->B + 31u) & ~31u;
  // Index.h:851-852: step = cur_num_nodes / num_initializations (integer), at least 1
  uint32_t step = (uint32_t)(cur_nodes / (uint64_t)ninit);
  p->step = step ? step : 1;
  p->nprobe = p->N ? (p->N + p->step - 1) / p->step : 0;
  p->query_vec_ok = (h.data_size % FNB_CHUNK_BYTES) == 0 ? 1u : 0u;

  p->Bpow2 = 1;
  while (p->Bpow2 * 2u <= p->Bcap) p->Bpow2 *= 2u;
  p->lines_per_row = (ix->stride * FNB_CHUNK_BYTES + 127u) / 128u;
  const char* env = getenv("FNB_VS_BUCKETS");  // development / test knob: visited-set buckets per query
  const int cpl = fnb_chunks_per_lane(ix->nchunks, ix->G);
  const int sms = ix->replicas.empty() ? 148 : ix->replicas[0].num_sms;
  if (launch_q <= 0) launch_q = Q;
  p->lat = allow_latency_variant ? choose_latency_variant(launch_q, sms) : 0u;
  // Two-hop prefetch of the CTA latency kernel: extra L2 fills that are free only while the batch leaves the memory
  // system idle (FNB_PF2_MAXQ queries per launch, default 64; 0 disables)
  static const long long pf2_maxq = [] {
    const char* e = getenv("FNB_PF2_MAXQ");
    return e ? atoll(e) : 64ll;
  }();
  static const long long pf2_minb = [] {  // lists shorter than this gain nothing (measured); the tests lower it
    const char* e = getenv("FNB_PF2_MINB");
    return e ? atoll(e) : 64ll;
  }();
  p->pf2 = (p->lat == 2u && launch_q <= pf2_maxq && (long long)p->B >= pf2_minb && h.M % 4 == 0) ? 1u : 0u;
  p->dense = p->lat ? 0u : choose_dense_plan(launch_q, sms, ix->G, cpl, p->B);
  if (p->lat == 2u)  // one query per CTA, up to 4 CTAs per SM: the visited set can have its full size
    size_visited(*p, env ? atoi(env) : 0, 4, 1);
  else
    size_visited(*p, env ? atoi(env) : 0, p->dense ? FNB_CTAS_DENSE : fnb_min_ctas(cpl));
  const uint32_t queries_per_cta = p->lat ? 1u : (uint32_t)FNB_WARPS_PER_CTA;
  if ((uint64_t)p->warp_smem * queries_per_cta + 1024u > 227u * 1024u)
    return fail(FNB_ERR_UNSUPPORTED, "ef_search=%d needs %u bytes of shared memory per query; limit is %u", ef,
                p->warp_smem, 226u * 1024u / queries_per_cta);
  return FNB_OK;
}

cudaError_t dispatch_search(const fnb_index* ix, const SearchParams& p, int num_sms, cudaStream_t s) {
  switch (ix->h.data_type) {
    case FNB_DTYPE_FLOAT32: return dispatch_search_f32(ix, p, num_sms, s);
    case FNB_DTYPE_UINT8: return dispatch_search_u8(ix, p, num_sms, s);
    default: return dispatch_search_i8(ix, p, num_sms, s);
  }
}

static int ensure_workspace(Lane* r, size_t dev_bytes, size_t host_bytes) {
  if (dev_bytes > r->ws_bytes) {
    if (r->ws) CU(cudaFree(r->ws));
    r->ws = nullptr;
    r->ws_bytes = 0;
    CU(cudaMalloc(&r->ws, dev_bytes));
    r->ws_bytes = dev_bytes;
  }
  if (host_bytes > r->h_pinned_bytes) {
    if (r->h_pinned) CU(cudaFreeHost(r->h_pinned));
    r->h_pinned = nullptr;
    r->h_pinned_bytes = 0;
    if (host_bytes < 65536) host_bytes = 65536;
    CU(cudaMallocHost(&r->h_pinned, host_bytes));
    r->h_pinned_bytes = host_bytes;
    CU(cudaHostGetDevicePointer((void**)&r->h_pinned_dev, r->h_pinned, 0));
  }
  return FNB_OK;
}

static inline size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

}  // namespace fnb

using namespace fnb;

// =====================================================================================================
extern "C" {

const char* fnb_last_error(void) { return g_last_error.c_str(); }
const char* fnb_version(void) { return "flatnav_b200 0.1 (sm_100a)"; }

int fnb_index_load(const char* path, int metric, int expect_dtype, const int* device_ids, int n_devices,
                   fnb_index** out) {
  if (!path) return fail(FNB_ERR_INVALID_ARG, "path is NULL");
  FILE* f = fopen(path, "rb");
  if (!f) return fail(FNB_ERR_IO, "Unable to open file for reading: %s", path);
  fseek(f, 0, SEEK_END);
  long long sz = ftell(f);
  fseek(f, 0, SEEK_SET);
  if (sz < 0) {
    fclose(f);
    return fail(FNB_ERR_IO, "cannot stat %s", path);
  }
  unsigned char* buf = (unsigned char*)malloc((size_t)sz ? (size_t)sz : 1);
  if (!buf) {
    fclose(f);
    return fail(FNB_ERR_NOMEM, "cannot allocate %lld bytes to read %s", sz, path);
  }
  size_t got = fread(buf, 1, (size_t)sz, f);
  fclose(f);
  int rc = got == (size_t)sz ? build_index(buf, got, metric, expect_dtype, device_ids, n_devices, out)
                             : fail(FNB_ERR_IO, "short read on %s", path);
  free(buf);
  return rc;
}

int fnb_index_from_memory(const void* file_bytes, size_t nbytes, int metric, int expect_dtype, const int* device_ids,
                          int n_devices, fnb_index** out) {
  if (!file_bytes) return fail(FNB_ERR_INVALID_ARG, "file_bytes is NULL");
  return build_index((const unsigned char*)file_bytes, nbytes, metric, expect_dtype, device_ids, n_devices, out);
}

void fnb_index_free(fnb_index* ix) {
  if (!ix) return;
  int prev = 0;
  cudaGetDevice(&prev);
  if (!ix->replicas.empty() && ix->replicas[0].device >= 0) cudaSetDevice(ix->replicas[0].device);
  fnb_build_scratch_free(ix->build);
  fnb_label_map_free(ix->label_map);
  for (auto& r : ix->replicas) free_replica(&r);
  cudaSetDevice(prev);
  delete ix;
}

int fnb_index_info(const fnb_index* ix, fnb_info* out) {
  if (!ix || !out) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  memset(out, 0, sizeof(*out));
  out->data_type = ix->h.data_type;
  out->metric = ix->h.metric;
  out->max_edges_per_node = ix->h.M;
  out->dim = ix->h.dim;
  out->data_size_bytes = ix->h.data_size;
  out->node_size_bytes = ix->h.node_size;
  out->max_node_count = ix->h.max_nodes;
  out->cur_num_nodes = ix->h.cur_nodes;
  out->n_devices = (int32_t)ix->replicas.size();
  for (size_t i = 0; i < ix->replicas.size() && i < 16; i++) out->device_ids[i] = ix->replicas[i].device;
  out->device_bytes = ix->replicas.empty() ? 0 : ix->replicas[0].device_bytes;
  out->row_stride_bytes = ix->stride * FNB_CHUNK_BYTES;
  out->lanes_per_row = (uint32_t)ix->G;
  return FNB_OK;
}

// scoped device buffer / device selection for the entry points below
namespace {
struct DevBuf {
  void* p = nullptr;
  ~DevBuf() { cudaFree(p); }
};
struct DeviceScope {
  int prev = 0;
  DeviceScope() { cudaGetDevice(&prev); }
  ~DeviceScope() { cudaSetDevice(prev); }
};
struct LastDeviceCall {  // which ring slot the calling thread's last fnb_search_device used
  const fnb_index* ix = nullptr;
  int replica = -1;
  uint32_t slot = 0;
};
thread_local LastDeviceCall t_last_device_call;
}  // namespace

int fnb_index_save(const fnb_index* ix, const char* path) {
  if (!ix || !path) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  SharedLock lock(ix->mu);
  const Header& h = ix->h;
  const Replica& r = ix->replicas[0];
  DeviceScope scope;
  CU(cudaSetDevice(r.device));
  const uint64_t blob_bytes = h.node_size * h.max_nodes;
  std::vector<unsigned char> host(FNB_HEADER_BYTES + blob_bytes, 0);
  int32_t dt = h.data_type;
  uint64_t v[7] = {h.M, h.data_size, h.node_size, h.max_nodes, h.cur_nodes, h.dim, h.data_size};
  memcpy(host.data(), &dt, 4);
  memcpy(host.data() + 4, v, 56);
  if (h.cur_nodes) {
    DevBuf blob;
    cudaStream_t s = nullptr;
    CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    struct StreamGuard {
      cudaStream_t s;
      ~StreamGuard() { cudaStreamDestroy(s); }
    } sg{s};
    CU(cudaMalloc(&blob.p, h.node_size * h.cur_nodes));
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((h.cur_nodes * 32 + threads - 1) / threads, (uint64_t)r.num_sms * 32);
    interleave_kernel<<<blocks, threads, 0, s>>>(static_cast<unsigned char*>(blob.p), h.cur_nodes, h.node_size,
                                                 (uint32_t)h.data_size, (uint32_t)h.M, ix->stride * FNB_CHUNK_BYTES,
                                                 reinterpret_cast<const unsigned char*>(r.vec), r.adj, r.labels);
    CU(cudaGetLastError());
    CU(cudaMemcpyAsync(host.data() + FNB_HEADER_BYTES, blob.p, h.node_size * h.cur_nodes, cudaMemcpyDeviceToHost, s));
    CU(cudaStreamSynchronize(s));
  }
  FILE* f = fopen(path, "wb");
  if (!f) return fail(FNB_ERR_IO, "Unable to open file for writing: %s", path);
  size_t put = fwrite(host.data(), 1, host.size(), f);
  fclose(f);
  if (put != host.size()) return fail(FNB_ERR_IO, "short write on %s", path);
  return FNB_OK;
}

int fnb_search_device(fnb_index* ix, int replica, const void* d_queries, int64_t Q, int K, int ef_search,
                      int num_initializations, float* d_out_dist, int32_t* d_out_label, uint32_t* d_ndist,
                      uint32_t* d_nhops, void* cuda_stream) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  if (replica < 0 || replica >= (int)ix->replicas.size()) return fail(FNB_ERR_INVALID_ARG, "bad replica %d", replica);
  SharedLock lock(ix->mu);  // the plan reads cur_num_nodes: not while a construction batch is changing it
  Replica& r = ix->replicas[replica];
  LanePool& P = *r.pool;
  static const bool no_pdl = getenv("FNB_NO_PDL") != nullptr;
  // Launch state: every call takes the next 64-byte slot of a ring (calls on different streams may overlap) and the
  // kernel leaves the slot clean, so nothing but the kernel goes on the stream.  Consecutive calls on one stream are
  // then adjacent launches, and with programmatic dependent launch the CTAs of the next one fill the SMs the tail of
  // this one leaves idle.  When the previous launch has not finished yet the caller is streaming batches: the tail no
  // longer costs anything, so the plan is chosen as for one long batch (the 28-warp instantiation where it exists).
  const uint32_t seq = P.ring_seq.fetch_add(1u) + 1u;
  const bool lat = choose_latency_variant(Q, r.num_sms) != 0;
  const bool streaming = !no_pdl && !lat && seq > 1u && *P.h_done_seq != seq - 1u;  // (latency variants keep their plan)
  SearchParams p;
  int rc = plan_search(ix, Q, K, ef_search, num_initializations, &p, streaming ? std::max<int64_t>(Q, (int64_t)1 << 22) : 0);
  if (rc != FNB_OK) return rc;
  if (Q == 0) return FNB_OK;
  if (!d_queries || !d_out_dist || !d_out_label) return fail(FNB_ERR_INVALID_ARG, "NULL buffer");
  cudaStream_t s = (cudaStream_t)cuda_stream;
  DeviceScope scope;
  CU(cudaSetDevice(r.device));
  const uint32_t slot = seq % FNB_RING_SLOTS;
  unsigned char* sl = P.ring + (size_t)slot * 64;
  p.vec = r.vec;
  p.adj = r.adj;
  p.labels = r.labels;
  p.queries = d_queries;
  if (((uintptr_t)d_queries & 15u) != 0) p.query_vec_ok = 0;
  p.out_dist = d_out_dist;
  p.out_label = d_out_label;
  p.out_ndist = d_ndist;
  p.out_nhops = d_nhops;
  p.counter = reinterpret_cast<unsigned int*>(sl + FNB_SLOT_COUNTER);
  p.done = reinterpret_cast<unsigned int*>(sl + FNB_SLOT_DONE);
  p.totals = reinterpret_cast<unsigned long long*>(sl + FNB_SLOT_TOTALS);
  p.last_totals = reinterpret_cast<unsigned long long*>(sl + FNB_SLOT_LAST_TOTALS);
  p.done_seq = P.d_done_seq;
  p.seq = seq;
  p.pdl = no_pdl ? 0u : 1u;
  cudaError_t e = dispatch_search(ix, p, r.num_sms, s);  // p.lat: as planned
  if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "search kernel launch failed: %s", cudaGetErrorString(e));
  t_last_device_call.ix = ix;
  t_last_device_call.replica = replica;
  t_last_device_call.slot = slot;
  return FNB_OK;
}

// Name of the traversal-kernel instantiation a search of this shape launches, spelled as ncu prints it.
int fnb_search_kernel_signature(const fnb_index* ix, int64_t Q, int K, int ef_search, char* out, size_t cap) {
  if (!ix || !out || cap == 0) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  SearchParams p;
  int rc = plan_search(ix, Q, K, ef_search, 100, &p);
  if (rc != FNB_OK) return rc;
  const int sms = ix->replicas.empty() ? 148 : ix->replicas[0].num_sms;
  const int lat = (int)p.lat;
  const int ch = fnb_chunks_per_lane(ix->nchunks, ix->G);
  int CH;
  if (ix->G == 4) CH = ch <= 1 ? 1 : 2;
  else if (ix->G == 8) CH = ch <= 3 ? ch : 4;
  else CH = ch <= 2 ? 2 : (ch <= 4 ? 4 : (ch <= 8 ? 8 : 16));
  const int dt = ix->h.data_type == FNB_DTYPE_FLOAT32 ? DT_F32 : (ix->h.data_type == FNB_DTYPE_UINT8 ? DT_U8 : DT_I8);
  const int exact = p.nchunks == (uint32_t)(ix->G * CH) ? 1 : 0;
  const int occ = (!lat && p.dense && CH <= 4 && ix->G <= 8) ? FNB_CTAS_DENSE : 0;
  if (lat == 2)
    snprintf(out, cap, "fnb_search_cta_kernel<%d,%d,%d,%d,%d>", dt, ix->h.metric == FNB_METRIC_IP ? M_IP : M_L2, ix->G, CH, exact);
  else
    snprintf(out, cap, "fnb_search_kernel<%d,%d,%d,%d,%d,%d,%d>", dt, ix->h.metric == FNB_METRIC_IP ? M_IP : M_L2, ix->G,
             CH, exact, lat, occ);
  return FNB_OK;
}

// What a search of this shape is planned with: shared memory per query, visited-set size, resident CTAs per SM.
int fnb_search_plan(const fnb_index* ix, int64_t Q, int K, int ef_search, fnb_plan_info* out) {
  if (!ix || !out) return fail(FNB_ERR_INVALID_ARG, "NULL argument");
  SearchParams p;
  int rc = plan_search(ix, Q, K, ef_search, 100, &p);
  if (rc != FNB_OK) return rc;
  memset(out, 0, sizeof(*out));
  const int sms = ix->replicas.empty() ? 148 : ix->replicas[0].num_sms;
  const int cpl = fnb_chunks_per_lane(ix->nchunks, ix->G);
  out->latency_variant = (int32_t)p.lat;
  out->dense_plan = out->latency_variant ? 0 : (int32_t)p.dense;
  out->list_capacity = (int32_t)p.Bcap;
  out->visited_slots = (int32_t)(p.vs_buckets * (p.vs_wide ? 4u : 8u));
  out->smem_bytes_per_query = (int32_t)p.warp_smem;
  const int by_regs = out->latency_variant == 2 ? 4 : (out->latency_variant ? 8 : (p.dense ? FNB_CTAS_DENSE : fnb_min_ctas(cpl)));
  const int queries_per_cta = out->latency_variant ? 1 : FNB_WARPS_PER_CTA;
  const int by_smem = (int)((227u * 1024u) / ((uint32_t)p.warp_smem * queries_per_cta + 1024u));
  out->ctas_per_sm = std::min(by_regs, by_smem);
  out->warps_per_sm = out->ctas_per_sm * (out->latency_variant == 2 ? FNB_CTA_WARPS : (out->latency_variant ? 1 : FNB_WARPS_PER_CTA));
  out->queries_per_sm = out->ctas_per_sm * queries_per_cta;
  return FNB_OK;
}

// Reads the totals written by the calling thread's last fnb_search_device call on `replica` (after the caller
// synchronised the stream it used).
int fnb_search_device_totals(fnb_index* ix, int replica, int64_t* n_dist, int64_t* n_hops, int64_t* n_short) {
  if (!ix || replica < 0 || replica >= (int)ix->replicas.size()) return fail(FNB_ERR_INVALID_ARG, "bad argument");
  if (t_last_device_call.ix != ix || t_last_device_call.replica != replica)
    return fail(FNB_ERR_INVALID_ARG, "this thread has not called fnb_search_device on replica %d of this index", replica);
  Replica& r = ix->replicas[replica];
  DeviceScope scope;
  CU(cudaSetDevice(r.device));
  unsigned long long t[3];
  CU(cudaMemcpy(t, r.pool->ring + (size_t)t_last_device_call.slot * 64 + FNB_SLOT_LAST_TOTALS, sizeof(t), cudaMemcpyDeviceToHost));
  if (n_dist) *n_dist = (int64_t)t[0];
  if (n_hops) *n_hops = (int64_t)t[1];
  if (n_short) *n_short = (int64_t)t[2];
  return FNB_OK;
}

// ---- latency batches of concurrent callers, combined -------------------------------------------------------------
// A latency batch (a query or a handful, e.g. Index::search from the C++ shim) costs one kernel launch, and the
// launches of one CUDA context are serialised in the driver: T host threads each launching their own kernel stop
// scaling at a few tens of thousands of launches per second.  So requests queue on the replica; whoever finds no
// launch being prepared becomes the collector, takes every waiting request with the same (K, ef_search,
// num_initializations), copies their queries into one lane's pinned block and launches ONE kernel for all of them; the
// next collector starts as soon as that launch is issued (several combined launches are in flight at a time), and the
// collector waits for its kernel (flag in pinned memory), hands every request its results and counters and wakes the
// callers.  Without contention a request is its own collector at once: nothing is added to a lone caller's path.
static int run_combined(fnb_index* ix, Replica& r, std::vector<CombReq*>& batch, bool* launched) {
  const Header& h = ix->h;
  const int K = batch[0]->K;
  int64_t total = 0;
  for (CombReq* c : batch) total += c->nq;
  SearchParams p;
  int rc = plan_search(ix, total, K, batch[0]->ef, batch[0]->ninit, &p);
  if (rc != FNB_OK) return rc;
  CU(cudaSetDevice(r.device));
  LaneHold hold;
  hold.r = &r;
  rc = acquire_lane(r, &hold.l);
  if (rc != FNB_OK) return rc;
  Lane& ln = *hold.l;
  const size_t qb = (size_t)total * h.data_size, ob = (size_t)total * K * 4, cb = (size_t)total * 4;
  const size_t off_dist = align256(qb), off_label = off_dist + align256(ob), off_nd = off_label + align256(ob),
               off_nh = off_nd + align256(cb), off_len = off_nh + align256(cb);
  rc = ensure_workspace(&ln, 0, off_len + align256(cb));
  if (rc != FNB_OK) return rc;
  unsigned char* dp = ln.h_pinned_dev;
  size_t at = 0;
  for (CombReq* c : batch) {
    memcpy(ln.h_pinned + at * h.data_size, c->q, (size_t)c->nq * h.data_size);
    at += (size_t)c->nq;
  }
  p.vec = r.vec;
  p.adj = r.adj;
  p.labels = r.labels;
  p.queries = dp;
  p.out_dist = reinterpret_cast<float*>(dp + off_dist);
  p.out_label = reinterpret_cast<int32_t*>(dp + off_label);
  p.out_ndist = reinterpret_cast<uint32_t*>(dp + off_nd);
  p.out_nhops = reinterpret_cast<uint32_t*>(dp + off_nh);
  p.out_len = reinterpret_cast<uint32_t*>(dp + off_len);
  p.done = ln.counter + FNB_SLOT_DONE / 4;
  p.done_seq = ln.h_flag_dev;
  p.seq = ++ln.seq ? ln.seq : ++ln.seq;
  cudaError_t e = dispatch_search(ix, p, r.num_sms, ln.stream);
  if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "search kernel launch failed: %s", cudaGetErrorString(e));
  // the launch is issued: let the next collector go while this one waits for its kernel
  {
    std::lock_guard<std::mutex> lk(r.pool->cmu);
    r.pool->claunching = false;
    *launched = true;
  }
  r.pool->ccv.notify_all();
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t spins = 0; *ln.h_flag != p.seq; spins++) {
    if (spins < 256u) {
#if defined(__x86_64__) || defined(__i386__)
      __builtin_ia32_pause();
#endif
      continue;
    }
    std::this_thread::yield();
    if ((spins & 1023u) == 0u && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) {
      CU(cudaStreamSynchronize(ln.stream));
      break;
    }
  }
  std::atomic_thread_fence(std::memory_order_acquire);
  hold.drained = true;
  const uint32_t* c_nd = reinterpret_cast<const uint32_t*>(ln.h_pinned + off_nd);
  const uint32_t* c_nh = reinterpret_cast<const uint32_t*>(ln.h_pinned + off_nh);
  const uint32_t* c_len = reinterpret_cast<const uint32_t*>(ln.h_pinned + off_len);
  at = 0;
  for (CombReq* c : batch) {
    memcpy(c->out_dist, ln.h_pinned + off_dist + at * K * 4, (size_t)c->nq * K * 4);
    memcpy(c->out_label, ln.h_pinned + off_label + at * K * 4, (size_t)c->nq * K * 4);
    for (int64_t i = 0; i < c->nq; i++) {
      c->nd += c_nd[at + i];
      c->nh += c_nh[at + i];
      c->ns += c_len[at + i] < (uint32_t)K ? 1 : 0;
    }
    at += (size_t)c->nq;
  }
  return FNB_OK;
}

static int search_latency_combined(fnb_index* ix, Replica& r, CombReq& me) {
  LanePool& P = *r.pool;
  const int64_t cap = 4ll * r.num_sms;  // a combined batch stays a latency batch
  std::unique_lock<std::mutex> lk(P.cmu);
  P.cpending.push_back(&me);
  P.ccv.wait(lk, [&] { return me.taken || !P.claunching; });
  if (me.taken) {  // somebody's launch carries this request
    P.ccv.wait(lk, [&] { return me.done; });
    return me.rc;
  }
  // collector: this request and every compatible one that is waiting
  P.claunching = true;
  std::vector<CombReq*> batch;
  int64_t total = 0;
  for (size_t i = 0; i < P.cpending.size();) {
    CombReq* c = P.cpending[i];
    const bool mine = c == &me;
    if ((mine || (c->K == me.K && c->ef == me.ef && c->ninit == me.ninit)) && (mine || total + c->nq + me.nq <= cap)) {
      c->taken = true;
      batch.push_back(c);
      total += c->nq;
      P.cpending.erase(P.cpending.begin() + (long)i);
    } else {
      i++;
    }
  }
  lk.unlock();
  bool launched = false;
  const int rc = run_combined(ix, r, batch, &launched);
  const std::string err = rc != FNB_OK ? g_last_error : std::string();
  lk.lock();
  if (!launched) P.claunching = false;  // failed before the launch was issued
  for (CombReq* c : batch) {
    c->rc = rc;
    c->err = err;
    c->done = true;
  }
  lk.unlock();
  P.ccv.notify_all();
  return me.rc;
}

// Host-buffer search.  Per replica the call takes a lane (stream + counters + staging) from the replica's pool, so
// concurrent callers proceed side by side.  Three ways for the bytes to travel, chosen per call:
//   * caller buffers are page-locked (cudaHostAlloc / cudaHostRegister, e.g. torch pinned tensors): used in place — the
//     kernel reads each query straight from host memory when a warp picks it up and writes the K results straight
//     back, so the transfers overlap the traversal instead of bracketing it;
//   * pageable buffers (what a numpy caller of the reference binding passes) up to FNB_STAGE_MAX bytes: the host copies
//     the queries chunk by chunk into the lane's pinned block WHILE the kernel runs — the kernel is launched after the
//     first chunk and every warp waits for a watermark (system-scope release / acquire) to pass its query — and the
//     kernel writes the results into the pinned block, from where they are copied out after the launch;
//   * anything larger: asynchronous copies through a device workspace on the lane's stream.
int fnb_search(fnb_index* ix, const void* queries, int64_t Q, int K, int ef_search, int num_initializations,
               float* out_dist, int32_t* out_label, fnb_search_stats* stats) {
  if (!ix) return fail(FNB_ERR_INVALID_ARG, "index is NULL");
  SharedLock lock(ix->mu);
  SearchParams p0;
  const int64_t n_rep = ix->replicas.empty() ? 1 : (int64_t)ix->replicas.size();
  int rc = plan_search(ix, Q, K, ef_search, num_initializations, &p0, (Q + n_rep - 1) / n_rep);
  if (rc != FNB_OK) return rc;
  if (stats) memset(stats, 0, sizeof(*stats));
  if (Q == 0) return FNB_OK;
  if (!queries || !out_dist || !out_label) return fail(FNB_ERR_INVALID_ARG, "NULL buffer");
  const Header& h = ix->h;
  const int R = (int)ix->replicas.size();
  DeviceScope scope;
  static const bool no_combine = getenv("FNB_NO_COMBINE") != nullptr || getenv("FNB_NO_STAGING") != nullptr ||
                                 getenv("FNB_NO_FLAG_WAIT") != nullptr || getenv("FNB_TIME_KERNELS") != nullptr;
  if (p0.lat && R == 1 && !no_combine) {
    CombReq me;
    me.q = static_cast<const unsigned char*>(queries);
    me.nq = Q;
    me.K = K;
    me.ef = ef_search;
    me.ninit = num_initializations;
    me.out_dist = out_dist;
    me.out_label = out_label;
    rc = search_latency_combined(ix, ix->replicas[0], me);
    if (rc != FNB_OK) {
      g_last_error = me.err;
      return rc;
    }
    if (stats) {
      stats->n_queries = Q;
      stats->n_dist = me.nd;
      stats->n_hops = me.nh;
      stats->n_short = me.ns;
      stats->algo_bytes = me.nd * (int64_t)h.data_size + me.nh * (int64_t)h.M * 4 + Q * (int64_t)h.data_size + Q * (int64_t)K * 8;
      stats->kernel_launches = 1;
    }
    if (me.ns > 0) {
      fail(FNB_SHORT_RESULT, "Search did not return the expected number of results for %lld of %lld queries.",
           (long long)me.ns, (long long)Q);
      return FNB_SHORT_RESULT;
    }
    return FNB_OK;
  }
  static const bool no_zero_copy = getenv("FNB_NO_ZEROCOPY") != nullptr;
  static const bool no_staging = getenv("FNB_NO_STAGING") != nullptr;
  static const bool no_feed = getenv("FNB_NO_FEED") != nullptr;
  static const bool no_flag = getenv("FNB_NO_FLAG_WAIT") != nullptr;
  static const bool time_kernels = getenv("FNB_TIME_KERNELS") != nullptr;
  static const size_t stage_max = getenv("FNB_STAGE_MAX") ? (size_t)atoll(getenv("FNB_STAGE_MAX")) : ((size_t)64 << 20);
  auto mapped = [](const void* host) -> unsigned char* {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) {
      cudaGetLastError();
      return nullptr;
    }
    return a.type == cudaMemoryTypeHost ? static_cast<unsigned char*>(a.devicePointer) : nullptr;
  };
  // (a latency batch always goes through the lane's pinned block: no pointer queries — three driver calls under the
  // context lock — on the path that concurrent single-query callers share)
  const bool probe = !no_zero_copy && !p0.lat;
  unsigned char* zq = probe ? mapped(queries) : nullptr;
  unsigned char* zd = probe ? mapped(out_dist) : nullptr;
  unsigned char* zl = probe ? mapped(out_label) : nullptr;
  if (!zd || !zl) zd = zl = nullptr;
  const int64_t per = (Q + R - 1) / R;
  struct Part {
    int64_t q0 = 0, nq = 0;
    bool stage_out = false, counters_in_block = false, by_flag = false, timed = true;
    uint32_t seq = 0;
    size_t off_dist = 0, off_label = 0, off_nd = 0, off_nh = 0, off_len = 0;
  };
  std::vector<Part> parts(R);
  std::vector<LaneHold> held;
  held.reserve(R);
  // enqueue everything on every replica first, then wait: replicas run concurrently from one host thread
  for (int i = 0; i < R; i++) {
    Replica& r = ix->replicas[i];
    Part& pt = parts[i];
    pt.q0 = std::min<int64_t>(Q, per * i);
    pt.nq = std::min<int64_t>(Q, per * (i + 1)) - pt.q0;
    held.emplace_back();
    if (pt.nq <= 0) continue;
    CU(cudaSetDevice(r.device));
    LaneHold& hold = held.back();
    hold.r = &r;
    rc = acquire_lane(r, &hold.l);
    if (rc != FNB_OK) return rc;
    Lane& ln = *hold.l;
    const size_t qb = (size_t)pt.nq * h.data_size, ob = (size_t)pt.nq * K * 4, cb = (size_t)pt.nq * 4;
    const unsigned char* q_src = (const unsigned char*)queries + (size_t)pt.q0 * h.data_size;
    SearchParams p = p0;
    p.Q = (uint32_t)pt.nq;
    p.vec = r.vec;
    p.adj = r.adj;
    p.labels = r.labels;
    // layout of the lane's pinned block: [queries (small batches only) | dist | label | per-query counters (lat only)]
    pt.counters_in_block = p.lat != 0;  // small batches: per-query counters by plain stores, no device totals to clear / fetch
    // queries of a pageable caller.  Small batch (latency variant): copied by the CPU into the pinned block the kernel
    // reads in place — nothing but the kernel on the stream.  Large batch: FED through the device workspace in chunks
    // on the lane's copy stream while the kernel already runs (the throughput variant hands queries out in index
    // order and every warp waits for the watermark to pass its query).
    const bool q_in_block = !zq && p.lat && !no_staging;
    const bool feed = !zq && !q_in_block && !no_feed && qb >= ((size_t)1 << 20);
    pt.off_dist = q_in_block ? align256(qb) : 0;
    pt.off_label = pt.off_dist + (zd ? 0 : align256(ob));
    pt.off_nd = pt.off_label + (zd ? 0 : align256(ob));
    pt.off_nh = pt.off_nd + (pt.counters_in_block ? align256(cb) : 0);
    pt.off_len = pt.off_nh + (pt.counters_in_block ? align256(cb) : 0);
    const size_t block = pt.off_len + (pt.counters_in_block ? align256(cb) : 0);
    pt.stage_out = !zd && !no_staging && block <= stage_max;
    const bool need_block = q_in_block || pt.stage_out || pt.counters_in_block;
    const size_t ws_q = (!zq && !q_in_block) ? align256(qb) : 0;
    rc = ensure_workspace(&ln, ws_q + (!zd && !pt.stage_out ? 2 * align256(ob) : 0) + 256, need_block ? block : 0);
    if (rc != FNB_OK) return rc;
    unsigned char* dp = ln.h_pinned_dev;
    unsigned char *d_q, *d_dist, *d_label;
    if (zq) d_q = zq + (size_t)pt.q0 * h.data_size;
    else if (q_in_block) d_q = dp;
    else d_q = ln.ws;
    if (zd) {
      d_dist = zd + (size_t)pt.q0 * K * 4;
      d_label = zl + (size_t)pt.q0 * K * 4;
    } else if (pt.stage_out) {
      d_dist = dp + pt.off_dist;
      d_label = dp + pt.off_label;
    } else {
      d_dist = ln.ws + ws_q;
      d_label = d_dist + align256(ob);
    }
    p.queries = d_q;
    if (((uintptr_t)d_q & 15u) != 0) p.query_vec_ok = 0;
    p.out_dist = reinterpret_cast<float*>(d_dist);
    p.out_label = reinterpret_cast<int32_t*>(d_label);
    if (pt.counters_in_block) {
      p.out_ndist = reinterpret_cast<uint32_t*>(dp + pt.off_nd);
      p.out_nhops = reinterpret_cast<uint32_t*>(dp + pt.off_nh);
      p.out_len = reinterpret_cast<uint32_t*>(dp + pt.off_len);
      p.counter = nullptr;
      p.totals = nullptr;
    } else {  // the lane's launch-state slot is left clean by the kernel itself: no memset on the stream
      p.counter = ln.counter;
      p.totals = ln.totals;
      p.done = ln.counter + FNB_SLOT_DONE / 4;
      p.last_totals = ln.h_totals_dev;  // straight into pinned host memory: no copy on the stream after the kernel
    }
    if (q_in_block) memcpy(ln.h_pinned, q_src, qb);
    if (feed) p.q_ready = ln.q_ready;
    // Completion: when nothing follows the kernel on the stream (results land in pinned memory by the kernel's own
    // stores) the host waits for a flag the kernel's last warp writes to pinned memory instead of synchronising the
    // stream — no driver call, no lock shared with other calling threads.  Device-side timing (stats->kernel_ms /
    // total_ms: four event records and a query per call) is then off unless FNB_TIME_KERNELS=1.
    pt.by_flag = !no_flag && (zd || pt.stage_out);
    pt.timed = time_kernels || !pt.by_flag;
    if (pt.by_flag) {
      if (!p.done) p.done = ln.counter + FNB_SLOT_DONE / 4;  // latency batches: no counter / totals, but the last-CTA epilogue
      p.done_seq = ln.h_flag_dev;
      p.seq = ++ln.seq ? ln.seq : ++ln.seq;
      pt.seq = p.seq;
    }
    if (pt.timed) CU(cudaEventRecord(ln.ev[0], ln.stream));
    if (!zq && !q_in_block && !feed) CU(cudaMemcpyAsync(d_q, q_src, qb, cudaMemcpyHostToDevice, ln.stream));
    if (feed) {
      CU(cudaMemsetAsync(ln.q_ready, 0, 4, ln.stream));
      CU(cudaEventRecord(ln.ev_feed, ln.stream));  // the first watermark must not land before the clear above
      CU(cudaStreamWaitEvent(ln.copy_stream, ln.ev_feed, 0));
    }
    if (pt.timed) CU(cudaEventRecord(ln.ev[1], ln.stream));
    cudaError_t e = dispatch_search(ix, p, r.num_sms, ln.stream);
    if (e != cudaSuccess) return fail(FNB_ERR_CUDA, "search kernel launch failed: %s", cudaGetErrorString(e));
    if (pt.timed) CU(cudaEventRecord(ln.ev[2], ln.stream));
    if (!zd && !pt.stage_out) {
      CU(cudaMemcpyAsync(out_dist + (size_t)pt.q0 * K, d_dist, ob, cudaMemcpyDeviceToHost, ln.stream));
      CU(cudaMemcpyAsync(out_label + (size_t)pt.q0 * K, d_label, ob, cudaMemcpyDeviceToHost, ln.stream));
    }
    if (pt.timed) CU(cudaEventRecord(ln.ev[3], ln.stream));
    if (feed) {
      // a pageable cudaMemcpyAsync returns once the chunk is staged, so this loop paces itself against the host copy
      // while the DMA of earlier chunks and the traversal of the queries already there proceed
      hold.feeding = true;
      const size_t chunk_q = std::max<size_t>(((size_t)pt.nq + FNB_FEED_CHUNKS - 1) / FNB_FEED_CHUNKS,
                                              std::max<size_t>(64, ((size_t)256 << 10) / h.data_size));
      size_t fed = 0;
      for (int c = 0; fed < (size_t)pt.nq; c++) {
        const size_t n = std::min(chunk_q, (size_t)pt.nq - fed);
        CU(cudaMemcpyAsync(d_q + fed * h.data_size, q_src + fed * h.data_size, n * h.data_size, cudaMemcpyHostToDevice,
                           ln.copy_stream));
        fed += n;
        ln.h_marks[c] = (uint32_t)fed;
        CU(cudaMemcpyAsync(ln.q_ready, ln.h_marks + c, 4, cudaMemcpyHostToDevice, ln.copy_stream));
      }
      hold.feeding = false;
    }
  }
  int64_t nd = 0, nh = 0, ns = 0;
  float kms = 0.f, tms = 0.f;
  int launches = 0;
  for (int i = 0; i < R; i++) {
    const Part& pt = parts[i];
    if (pt.nq <= 0) continue;
    Replica& r = ix->replicas[i];
    Lane& ln = *held[i].l;
    CU(cudaSetDevice(r.device));
    if (pt.by_flag) {
      // spin on the pinned word; after ~10 us offer the core to whoever else wants it on every turn (many callers
      // spinning on all cores of the host must not starve each other's launch paths); after 2 s fall back to the stream
      // (a failed launch never writes the flag: the synchronise reports the error)
      const auto t0 = std::chrono::steady_clock::now();
      for (uint32_t spins = 0; *ln.h_flag != pt.seq; spins++) {
        if (spins < 256u) {
#if defined(__x86_64__) || defined(__i386__)
          __builtin_ia32_pause();
#endif
          continue;
        }
        std::this_thread::yield();
        if ((spins & 1023u) == 0u && std::chrono::steady_clock::now() - t0 > std::chrono::seconds(2)) {
          CU(cudaStreamSynchronize(ln.stream));
          break;
        }
      }
      std::atomic_thread_fence(std::memory_order_acquire);
      held[i].drained = true;
      if (pt.timed) CU(cudaEventSynchronize(ln.ev[3]));  // returns at once: the kernel is done, the record is all that is left
    } else {
      CU(cudaStreamSynchronize(ln.stream));
      held[i].drained = true;
    }
    float a = 0.f, b = 0.f;
    if (pt.timed) {
      CU(cudaEventElapsedTime(&a, ln.ev[1], ln.ev[2]));
      CU(cudaEventElapsedTime(&b, ln.ev[0], ln.ev[3]));
    }
    if (pt.stage_out) {
      const size_t ob = (size_t)pt.nq * K * 4;
      memcpy(out_dist + (size_t)pt.q0 * K, ln.h_pinned + pt.off_dist, ob);
      memcpy(out_label + (size_t)pt.q0 * K, ln.h_pinned + pt.off_label, ob);
    }
    if (pt.counters_in_block) {
      const uint32_t* c_nd = reinterpret_cast<const uint32_t*>(ln.h_pinned + pt.off_nd);
      const uint32_t* c_nh = reinterpret_cast<const uint32_t*>(ln.h_pinned + pt.off_nh);
      const uint32_t* c_len = reinterpret_cast<const uint32_t*>(ln.h_pinned + pt.off_len);
      for (int64_t q = 0; q < pt.nq; q++) {
        nd += c_nd[q];
        nh += c_nh[q];
        ns += c_len[q] < (uint32_t)K ? 1 : 0;
      }
    } else {
      const unsigned long long* t = ln.h_totals;
      nd += (int64_t)t[0];
      nh += (int64_t)t[1];
      ns += (int64_t)t[2];
    }
    kms = std::max(kms, a);
    tms = std::max(tms, b);
    launches++;
  }
  held.clear();  // lanes back to their pools
  if (stats) {
    stats->n_queries = Q;
    stats->n_dist = nd;
    stats->n_hops = nh;
    stats->n_short = ns;
    stats->algo_bytes = nd * (int64_t)h.data_size + nh * (int64_t)h.M * 4 + Q * (int64_t)h.data_size + Q * (int64_t)K * 8;
    stats->kernel_ms = kms;
    stats->total_ms = tms;
    stats->kernel_launches = launches;
  }
  if (ns > 0) {
    fail(FNB_SHORT_RESULT, "Search did not return the expected number of results for %lld of %lld queries.",
         (long long)ns, (long long)Q);
    return FNB_SHORT_RESULT;
  }
  return FNB_OK;
}

}  // extern "C"
