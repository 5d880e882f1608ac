// Exercises the C++ shim (include/flatnav_b200/Index.h) the way the reference's own C++ callers use
// flatnav::Index (tools/query_npy.cpp:43-68, include/flatnav/tests/test_serialization.cpp:50-75).
// usage: shim_test <index.idx> <queries.f32> <Q> <K> <ef> <out_prefix>
#define FLATNAV_B200_AS_FLATNAV
#include <flatnav_b200/Index.h>

#include <cstdio>
#include <cstdlib>
#include <fstream>

using flatnav::Index;
using flatnav::distances::SquaredL2Distance;
using flatnav::util::DataType;

int main(int argc, char** argv) {
  if (argc < 7) return 2;
  const std::string idx = argv[1], qpath = argv[2], out = argv[6];
  const size_t Q = std::strtoull(argv[3], nullptr, 10);
  const int K = std::atoi(argv[4]), ef = std::atoi(argv[5]);
  int checks = 0;

  try {
    Index<SquaredL2Distance<DataType::float32>, int>::loadIndex("/nonexistent/x.idx");
    return 10;
  } catch (const std::runtime_error&) { checks++; }

  auto index = Index<SquaredL2Distance<DataType::float32>, int>::loadIndex(idx);
  const size_t D = index->dataDimension();
  if (index->nodeSizeBytes() != index->dataSizeBytes() + 4 * index->maxEdgesPerNode() + 4) return 11;  // test_serialization.cpp:56
  if (index->getDataType() != DataType::float32) return 12;

  std::vector<float> q(Q * D);
  std::ifstream f(qpath, std::ios::binary);
  f.read((char*)q.data(), (std::streamsize)(q.size() * sizeof(float)));

  try {
    index->search(q.data(), K, ef, 0);
    return 13;
  } catch (const std::invalid_argument&) { checks++; }  // Index.h:847-849
  try {
    index->setNumThreads(0);
    return 14;
  } catch (const std::invalid_argument&) { checks++; }  // Index.h:493-497

  // one query at a time, like tools/query_npy.cpp:51
  std::vector<float> d1(Q * K);
  std::vector<int> l1(Q * K);
  for (size_t i = 0; i < Q; i++) {
    auto r = index->search(q.data() + i * D, K, ef);
    if ((int)r.size() != K) return 15;
    for (int j = 0; j < K; j++) {
      d1[i * K + j] = r[j].first;
      l1[i * K + j] = r[j].second;
    }
  }
  // batched
  std::vector<float> d2(Q * K);
  std::vector<int> l2(Q * K);
  index->searchBatch(q.data(), Q, K, ef, d2.data(), l2.data());
  if (d1 != d2 || l1 != l2) return 16;

  // save -> load -> identical results (test_serialization.cpp:64-75)
  index->saveIndex(out + ".resaved.idx");
  auto again = Index<SquaredL2Distance<DataType::float32>, int>::loadIndex(out + ".resaved.idx");
  std::vector<float> d3(Q * K);
  std::vector<int> l3(Q * K);
  again->searchBatch(q.data(), Q, K, ef, d3.data(), l3.data());
  if (d3 != d2 || l3 != l2) return 17;

  // more results than nodes: search() returns fewer than K, searchBatch throws (bindings.cpp:184-189)
  auto few = index->search(q.data(), (int)index->currentNumNodes() + 3, 8);
  if (few.size() > index->currentNumNodes()) return 18;
  try {
    std::vector<float> dd(Q * (index->currentNumNodes() + 3));
    std::vector<int> ll(dd.size());
    index->searchBatch(q.data(), Q, (int)index->currentNumNodes() + 3, 8, dd.data(), ll.data());
    return 19;
  } catch (const std::runtime_error&) { checks++; }

  // re-ordering (Index.h:412-440): the saved file is compared with the reference's by the caller
  {
    auto r = Index<SquaredL2Distance<DataType::float32>, int>::loadIndex(idx);
    auto table = r->getGraphOutdegreeTable();  // Index.h:240-251
    if (table.size() != r->currentNumNodes()) return 21;
    for (size_t n = 0; n < table.size(); n++)
      for (auto v : table[n])
        if (v == n || v >= table.size()) return 22;
    try {
      r->doGraphReordering({"rcm", "nope"});
      return 23;
    } catch (const std::invalid_argument&) { checks++; }  // Index.h:421-423, after "rcm" has been applied
    r->reorderGOrder();
    r->saveIndex(out + ".rcm_gorder.idx");
  }

  // construction through the shim, the way tools/construct_npy.cpp:37-54,78-80 builds an index: distance object ->
  // constructor -> addBatch -> saveIndex; then every inserted vector must find itself (distance 0, its label)
  {
    using flatnav::distances::InnerProductDistance;
    const int M = 16;
    auto distance = SquaredL2Distance<>::create(D);
    if (distance->dimension() != D || distance->dataSize() != 4 * D) return 30;
    Index<SquaredL2Distance<DataType::float32>, int> built(std::move(distance), (int)Q + 1, M, /*collect_stats=*/true);
    if (built.currentNumNodes() != 0 || built.maxNodeCount() != Q + 1 || built.maxEdgesPerNode() != (size_t)M) return 31;
    std::vector<int> labels(Q);
    for (size_t i = 0; i < Q; i++) labels[i] = 1000 + (int)i;
    try {
      built.addBatch<float>(q.data(), labels, 64, 0);
      return 32;
    } catch (const std::invalid_argument&) { checks++; }  // Index.h:303-305
    built.addBatch<float>(q.data(), labels, 64);
    if (built.currentNumNodes() != Q) return 33;
    int one = 7777;
    uint32_t nid = 0;
    std::vector<float> far(D, 1e3f);
    built.allocateNode(far.data(), one, nid);  // Index.h:262-272: unlinked, unreachable from the graph
    if (nid != Q || built.currentNumNodes() != Q + 1) return 34;
    try {
      built.add(far.data(), one, 64, 100);
      return 35;
    } catch (const std::runtime_error&) { checks++; }  // Index.h:353-357: index full
    size_t self = 0;
    for (size_t i = 0; i < Q; i++) {
      auto r = built.search(q.data() + i * D, 1, 64);
      if (r.size() == 1 && r[0].second == 1000 + (int)i && r[0].first == 0.0f) self++;
    }
    if (self * 100 < Q * 95) return 36;
    if (built.distanceComputations() == 0 || built.metricHops() == 0) return 37;  // collect_stats (Index.h:529)
    built.resetStats();
    if (built.distanceComputations() != 0) return 38;
    built.getIndexSummary();  // Index.h:538-547
    built.saveIndex(out + ".built.idx");
    auto reloaded = Index<SquaredL2Distance<DataType::float32>, int>::loadIndex(out + ".built.idx");
    if (reloaded->currentNumNodes() != Q + 1 || reloaded->maxEdgesPerNode() != (size_t)M) return 39;
    auto a = built.search(q.data(), 5, 32), b = reloaded->search(q.data(), 5, 32);
    if (a != b) return 40;
    if (reloaded->distanceComputations() != 0) return 41;  // a loaded index does not collect stats
  }

  auto moved = std::move(*index);  // move-only ownership (Index.h:86-132)
  if (moved.currentNumNodes() == 0) return 20;

  std::ofstream fd(out + ".dist.bin", std::ios::binary), fl(out + ".label.bin", std::ios::binary);
  fd.write((const char*)d2.data(), (std::streamsize)(d2.size() * sizeof(float)));
  fl.write((const char*)l2.data(), (std::streamsize)(l2.size() * sizeof(int)));
  std::printf("shim ok, %d exception checks\n", checks);
  return 0;
}
