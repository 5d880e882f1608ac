// Index::search is re-entrant in the reference (include/flatnav/index/Index.h:387-409; every thread takes its own
// visited set from the pool, util/VisitedSetPool.h:154-172) and its callers fan it out with executeInParallel
// (util/Multithreading.h:18-48, python-bindings/src/flatnav/bindings.cpp:196-212).  Same pattern through the shim:
// T threads pull query ids from an atomic counter and call index->search(query, K, ef) one query at a time.
// usage: concurrent_search <index.idx> <queries.f32> <Q> <K> <ef> <threads>
#define FLATNAV_B200_AS_FLATNAV
#include <flatnav_b200/Index.h>

#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <thread>
#include <vector>

using flatnav::Index;
using flatnav::distances::SquaredL2Distance;
using flatnav::util::DataType;
typedef std::vector<std::pair<float, int>> Result;

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const size_t Q = std::strtoull(argv[3], nullptr, 10);
  const int K = std::atoi(argv[4]), ef = std::atoi(argv[5]), T = std::atoi(argv[6]);
  auto index = Index<SquaredL2Distance<DataType::float32>, int>::loadIndex(argv[1]);
  const size_t D = index->dataDimension();
  std::vector<float> q(Q * D);
  std::ifstream f(argv[2], std::ios::binary);
  f.read((char*)q.data(), (std::streamsize)(q.size() * sizeof(float)));

  std::vector<Result> serial(Q), conc(Q);
  for (size_t i = 0; i < 32 && i < Q; i++) index->search(q.data() + i * D, K, ef);  // warm-up
  auto t0 = std::chrono::steady_clock::now();
  for (size_t i = 0; i < Q; i++) serial[i] = index->search(q.data() + i * D, K, ef);
  const double s_serial = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  std::atomic<size_t> next{0};
  std::atomic<int> failed{0};
  auto worker = [&]() {
    try {
      for (;;) {
        const size_t i = next.fetch_add(1);
        if (i >= Q) break;
        conc[i] = index->search(q.data() + i * D, K, ef);
      }
    } catch (...) {
      failed++;
    }
  };
  t0 = std::chrono::steady_clock::now();
  std::vector<std::thread> th;
  for (int t = 0; t < T; t++) th.emplace_back(worker);
  for (auto& t : th) t.join();
  const double s_conc = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  if (failed) return 10;
  for (size_t i = 0; i < Q; i++)
    if (conc[i] != serial[i] || (int)serial[i].size() != K) return 11;
  std::printf("concurrent == serial for %zu queries; 1 thread: %.0f queries/s, %d threads: %.0f queries/s (x%.2f)\n", Q,
              Q / s_serial, T, Q / s_conc, s_serial / s_conc);
  return 0;
}
