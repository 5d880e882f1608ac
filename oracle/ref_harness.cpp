// TEST INFRASTRUCTURE — not part of the product.
//
// Driver for the UNMODIFIED reference (BlaiseMuhirwa/flatnav).  It is compiled by
// oracle/Makefile against the reference headers where they lie
// (-I$(FLATNAV_REF)/include -I$(FLATNAV_REF)/external/cereal/include); no reference
// source is copied into this repository.  The resulting binaries land in oracle/_ref/
// (git-ignored, but they travel to the GPU box).
//
// It reproduces, with the reference's own classes,
//   * index construction:  Index::addBatch + Index::saveIndex   (Index.h:301, :481)
//   * the batched search loop of the Python binding:
//       executeInParallel(0, Q, T, [&](i){ index->search(q_i, K, ef, ninit); })
//     (python-bindings/src/flatnav/bindings.cpp:196-212, util/Multithreading.h:18-48)
// and dumps raw (distance f32, label i32) arrays so the oracle port and the CUDA path can
// be compared against what the reference itself returns.
//
// Usage:
//   ref_flatnav build  <l2|ip> <f32|u8|i8> <data.bin> <N> <D> <M> <efc> <threads> <out.idx>
//   ref_flatnav search <l2|ip> <f32|u8|i8> <index.idx> <queries.bin> <Q> <K> <ef> <ninit>
//                      <threads> <reps> <out_prefix|->
//   ref_flatnav info   <l2|ip> <f32|u8|i8> <index.idx>
//   ref_flatnav reorder <l2|ip> <f32|u8|i8> <in.idx> <out.idx> <gorder|rcm>[,<gorder|rcm>...]
//                      Index::loadIndex + doGraphReordering (Index.h:412-427) + saveIndex
//   ref_flatnav mtx    <l2|ip> <f32|u8|i8> <data.bin> <N> <D> <M> <graph.mtx> <out.idx>
//                      Index::allocateNode per row (labels 0..N-1, bindings.cpp:308-324) + buildGraphLinks
//                      (Index.h:187-238) + saveIndex

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <mutex>
#include <numeric>
#include <queue>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

// Note: the reference keeps `_collect_stats` private and does not serialise it (Index.h:75), so a
// LOADED index can never count distance computations; per-query distance/hop counts therefore come
// from the oracle port (flatnav_oracle.cpp), not from this binary.
#include <flatnav/index/Index.h>
#include <flatnav/distances/InnerProductDistance.h>
#include <flatnav/distances/SquaredL2Distance.h>
#include <flatnav/util/Multithreading.h>

using flatnav::Index;
using flatnav::distances::InnerProductDistance;
using flatnav::distances::SquaredL2Distance;
using flatnav::util::DataType;

static std::vector<char> read_file(const std::string& path, size_t expect_bytes) {
  std::ifstream f(path, std::ios::binary);
  if (!f.is_open()) {
    std::fprintf(stderr, "cannot open %s\n", path.c_str());
    std::exit(2);
  }
  std::vector<char> buf(expect_bytes);
  f.read(buf.data(), (std::streamsize)expect_bytes);
  if ((size_t)f.gcount() != expect_bytes) {
    std::fprintf(stderr, "%s: expected %zu bytes, got %zu\n", path.c_str(), expect_bytes, (size_t)f.gcount());
    std::exit(2);
  }
  return buf;
}

template <typename dist_t, typename elem_t>
static int do_build(DataType dt, char** a, int nargs) {
  std::string data_path = a[0];
  size_t N = std::strtoull(a[1], nullptr, 10), D = std::strtoull(a[2], nullptr, 10);
  int M = std::atoi(a[3]), efc = std::atoi(a[4]), threads = std::atoi(a[5]);
  std::string out = a[6];
  auto data = read_file(data_path, N * D * sizeof(elem_t));
  auto dist = std::make_unique<dist_t>(D);
  auto index = std::make_unique<Index<dist_t, int>>(std::move(dist), (int)N, M, /*collect_stats=*/false, dt);
  index->setNumThreads((uint32_t)threads);
  std::vector<int> labels(N);
  std::iota(labels.begin(), labels.end(), nargs > 7 ? std::atoi(a[7]) : 0);  // optional first label (dataset shards)
  auto t0 = std::chrono::steady_clock::now();
  index->template addBatch<elem_t>((void*)data.data(), labels, efc);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  index->saveIndex(out);
  std::printf("{\"op\":\"build\",\"N\":%zu,\"D\":%zu,\"M\":%d,\"efc\":%d,\"threads\":%d,\"seconds\":%.3f}\n", N, D, M,
              efc, threads, secs);
  return 0;
}

template <typename dist_t, typename elem_t>
static int do_search(char** a, int nargs) {
  std::string idx = a[0], qpath = a[1];
  size_t Q = std::strtoull(a[2], nullptr, 10);
  int K = std::atoi(a[3]), ef = std::atoi(a[4]), ninit = std::atoi(a[5]);
  int threads = std::atoi(a[6]), reps = std::atoi(a[7]);
  std::string out_prefix = a[8];
  (void)nargs;

  auto index = Index<dist_t, int>::loadIndex(idx);
  size_t D = index->dataDimension();
  auto qbuf = read_file(qpath, Q * D * sizeof(elem_t));
  const elem_t* queries = reinterpret_cast<const elem_t*>(qbuf.data());

  std::vector<float> dists(Q * (size_t)K, std::numeric_limits<float>::infinity());
  std::vector<int> labels(Q * (size_t)K, -1);
  std::vector<int> counts(Q, 0);

  auto one = [&](uint32_t i) {
    auto r = index->search((const void*)(queries + (size_t)i * D), K, ef, ninit);
    counts[i] = (int)r.size();
    for (size_t j = 0; j < r.size() && j < (size_t)K; j++) {
      dists[(size_t)i * K + j] = r[j].first;
      labels[(size_t)i * K + j] = r[j].second;
    }
  };
  auto pass = [&]() {
    auto t0 = std::chrono::steady_clock::now();
    if (threads == 1) {
      for (uint32_t i = 0; i < Q; i++) one(i);  // bindings.cpp:177-195
    } else {
      flatnav::executeInParallel(0, (uint32_t)Q, (uint32_t)threads, one);  // bindings.cpp:198-211
    }
    return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };

  pass();  // warm-up
  std::vector<double> secs;
  for (int r = 0; r < reps; r++) secs.push_back(pass());
  double best = secs.empty() ? 0.0 : *std::min_element(secs.begin(), secs.end());

  if (out_prefix != "-") {
    std::ofstream fd(out_prefix + ".dist.bin", std::ios::binary), fl(out_prefix + ".label.bin", std::ios::binary);
    fd.write((const char*)dists.data(), (std::streamsize)(dists.size() * sizeof(float)));
    fl.write((const char*)labels.data(), (std::streamsize)(labels.size() * sizeof(int)));
  }
  size_t short_results = 0;
  for (auto c : counts) short_results += (c < K);
  std::ostringstream runs;
  for (size_t i = 0; i < secs.size(); i++) runs << (i ? "," : "") << (double)Q / secs[i];
  std::printf(
      "{\"op\":\"search\",\"Q\":%zu,\"K\":%d,\"ef\":%d,\"ninit\":%d,\"threads\":%d,\"hw_threads\":%u,"
      "\"qps_best\":%.3f,\"qps_runs\":[%s],\"short_results\":%zu}\n",
      Q, K, ef, ninit, threads, std::thread::hardware_concurrency(), best > 0 ? (double)Q / best : 0.0,
      runs.str().c_str(), short_results);
  return 0;
}

// Per-query latency of Index::search on one thread, the loop of experiments/run-benchmark.py:66-84
// (search_single per query, wall clock around each call).  One untimed warm-up pass, then one timed pass.
template <typename dist_t, typename elem_t>
static int do_latency(char** a) {
  std::string idx = a[0], qpath = a[1];
  size_t Q = std::strtoull(a[2], nullptr, 10);
  int K = std::atoi(a[3]), ef = std::atoi(a[4]), ninit = std::atoi(a[5]);
  std::string out = a[6];
  auto index = Index<dist_t, int>::loadIndex(idx);
  size_t D = index->dataDimension();
  auto qbuf = read_file(qpath, Q * D * sizeof(elem_t));
  const elem_t* queries = reinterpret_cast<const elem_t*>(qbuf.data());
  std::vector<double> lat(Q, 0.0);
  size_t sink = 0;
  for (int pass = 0; pass < 2; pass++) {
    for (size_t i = 0; i < Q; i++) {
      auto t0 = std::chrono::steady_clock::now();
      auto r = index->search((const void*)(queries + i * D), K, ef, ninit);
      lat[i] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      sink += r.size();
    }
  }
  std::ofstream f(out, std::ios::binary);
  f.write((const char*)lat.data(), (std::streamsize)(lat.size() * sizeof(double)));
  double total = 0;
  for (double v : lat) total += v;
  std::printf("{\"op\":\"latency\",\"Q\":%zu,\"K\":%d,\"ef\":%d,\"mean_us\":%.3f,\"results\":%zu}\n", Q, K, ef,
              Q ? total / (double)Q * 1e6 : 0.0, sink);
  return 0;
}

template <typename dist_t>
static int do_reorder(char** a) {
  auto index = Index<dist_t, int>::loadIndex(a[0]);
  std::vector<std::string> methods;
  std::stringstream ss(a[2]);
  for (std::string m; std::getline(ss, m, ',');) methods.push_back(m);
  auto t0 = std::chrono::steady_clock::now();
  index->doGraphReordering(methods);
  double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  index->saveIndex(a[1]);
  std::printf("{\"op\":\"reorder\",\"methods\":\"%s\",\"seconds\":%.3f}\n", a[2], secs);
  return 0;
}

template <typename dist_t, typename elem_t>
static int do_mtx(DataType dt, char** a) {
  size_t N = std::strtoull(a[1], nullptr, 10), D = std::strtoull(a[2], nullptr, 10);
  int M = std::atoi(a[3]);
  auto data = read_file(a[0], N * D * sizeof(elem_t));
  auto dist = std::make_unique<dist_t>(D);
  auto index = std::make_unique<Index<dist_t, int>>(std::move(dist), (int)N, M, /*collect_stats=*/false, dt);
  for (size_t i = 0; i < N; i++) {
    int label = (int)i;
    uint32_t id;
    index->allocateNode((void*)(data.data() + i * D * sizeof(elem_t)), label, id);
  }
  index->buildGraphLinks(a[4]);
  index->saveIndex(a[5]);
  std::printf("{\"op\":\"mtx\",\"N\":%zu}\n", N);
  return 0;
}

template <typename dist_t>
static int do_info(char** a) {
  auto index = Index<dist_t, int>::loadIndex(a[0]);
  std::printf(
      "{\"data_type\":%d,\"M\":%zu,\"data_size_bytes\":%zu,\"node_size_bytes\":%zu,\"max_node_count\":%zu,"
      "\"cur_num_nodes\":%zu,\"dim\":%zu}\n",
      (int)index->getDataType(), index->maxEdgesPerNode(), index->dataSizeBytes(), index->nodeSizeBytes(),
      index->maxNodeCount(), index->currentNumNodes(), index->dataDimension());
  return 0;
}

template <typename dist_t, typename elem_t>
static int run(const std::string& op, DataType dt, char** a, int n) {
  if (op == "build" && n >= 7) return do_build<dist_t, elem_t>(dt, a, n);
  if (op == "search" && n >= 9) return do_search<dist_t, elem_t>(a, n);
  if (op == "latency" && n >= 7) return do_latency<dist_t, elem_t>(a);
  if (op == "info" && n >= 1) return do_info<dist_t>(a);
  if (op == "reorder" && n >= 3) return do_reorder<dist_t>(a);
  if (op == "mtx" && n >= 6) return do_mtx<dist_t, elem_t>(dt, a);
  std::fprintf(stderr, "bad arguments for %s\n", op.c_str());
  return 2;
}

int main(int argc, char** argv) {
  if (argc < 4) {
    std::fprintf(stderr, "usage: ref_flatnav build|search|latency|info <l2|ip> <f32|u8|i8> ...\n");
    return 2;
  }
  std::string op = argv[1], metric = argv[2], dtype = argv[3];
  char** a = argv + 4;
  int n = argc - 4;
  try {
    if (metric == "l2" && dtype == "f32") return run<SquaredL2Distance<DataType::float32>, float>(op, DataType::float32, a, n);
    if (metric == "l2" && dtype == "u8") return run<SquaredL2Distance<DataType::uint8>, uint8_t>(op, DataType::uint8, a, n);
    if (metric == "l2" && dtype == "i8") return run<SquaredL2Distance<DataType::int8>, int8_t>(op, DataType::int8, a, n);
    if (metric == "ip" && dtype == "f32") return run<InnerProductDistance<DataType::float32>, float>(op, DataType::float32, a, n);
    if (metric == "ip" && dtype == "u8") return run<InnerProductDistance<DataType::uint8>, uint8_t>(op, DataType::uint8, a, n);
    if (metric == "ip" && dtype == "i8") return run<InnerProductDistance<DataType::int8>, int8_t>(op, DataType::int8, a, n);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "reference threw: %s\n", e.what());
    return 3;
  }
  std::fprintf(stderr, "unknown metric/dtype\n");
  return 2;
}
