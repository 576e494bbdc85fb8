// Compiled-host example / test of include/tmc.hpp: reads a CSR (text), runs tfidfScaleSparseMat + hSpecClust through the
// C ABI and prints the make-tree stdout CSV (cell,cluster,path).  Without a CUDA device it reports TMC_ERR_NO_DEVICE.
#include <cstdio>
#include <fstream>
#include <iostream>
#include "../../include/tmc.hpp"

// `host_cpp_example matrix.txt variants`: the non-default calls of the same surface — KMeansGroup with a permutation test on the
// sparse matrix and the --dense overload (HSD.hierarchicalSpectralCluster) — printed as "nodes <n> splits-with-p <k>" per call.
static int run_variants(const tmc::SpMatrix& b2, const std::vector<std::string>& names) {
  std::vector<tmc::CellInfo> items;
  for (int64_t i = 0; i < (int64_t)names.size(); ++i) items.push_back(tmc::CellInfo{names[i], i});
  auto summarize = [](const tmc::ClusteringTree<tmc::CellInfo>& t) {
    long long nodes = 0, with_p = 0;
    std::vector<const tmc::ClusteringTree<tmc::CellInfo>*> st{&t};
    while (!st.empty()) {
      const auto* n = st.back(); st.pop_back();
      ++nodes; if (n->rootLabel._pVal) ++with_p;
      for (const auto& c : n->subForest) st.push_back(&c);
    }
    std::printf("nodes %lld splits-with-p %lld\n", nodes, with_p);
  };
  summarize(tmc::hierarchicalSpectralCluster<tmc::CellInfo>(tmc::EigenGroup::KMeansGroup, false, 1, std::nullopt, std::nullopt, 4, items, b2));
  tmc::DenseMatrix d; d.nrows = b2.nrows; d.ncols = b2.ncols; d.values.assign((size_t)(b2.nrows * b2.ncols), 0.0);
  for (int64_t i = 0; i < b2.nrows; ++i)
    for (int64_t k = b2.rowPtr[i]; k < b2.rowPtr[i + 1]; ++k) d.values[(size_t)(i * b2.ncols + b2.colIdx[k])] = b2.values[k];
  summarize(tmc::hierarchicalSpectralCluster<tmc::CellInfo>(tmc::EigenGroup::SignGroup, false, std::nullopt, std::nullopt, std::nullopt,
                                                           std::nullopt, items, d));
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s matrix.txt [variants]\n", argv[0]); return 2; }
  std::ifstream in(argv[1]);
  tmc::SpMatrix a; int64_t nnz = 0;
  in >> a.nrows >> a.ncols >> nnz;
  a.rowPtr.assign(a.nrows + 1, 0); a.colIdx.resize(nnz); a.values.resize(nnz);
  for (auto& x : a.rowPtr) in >> x;
  for (auto& x : a.colIdx) in >> x;
  for (auto& x : a.values) in >> x;
  std::vector<std::string> names;
  for (int64_t i = 0; i < a.nrows; ++i) names.push_back("CELL" + std::to_string(i));
  try {
    tmc::SpMatrix b2 = tmc::tfidfScaleSparseMat(a);                                               // Preprocess.hs:204-205
    if (argc > 2 && std::string(argv[2]) == "variants") return run_variants(b2, names);
    tmc::ClusterResults r = tmc::hSpecClust(tmc::EigenGroup::SignGroup, std::nullopt, std::nullopt, std::nullopt, b2, names);   // Cluster.hs:135
    std::printf("cell,cluster,path\n");
    for (auto& e : r._clusterList) {
      std::printf("%s,%lld,", e.first._barcode.c_str(), (long long)e.second[0]);
      for (size_t k = 0; k < e.second.size(); ++k) std::printf("%s%lld", k ? "/" : "", (long long)e.second[k]);
      std::printf("\n");
    }
  } catch (const tmc::Error& e) {
    if (e.status == TMC_ERR_NO_DEVICE) { std::printf("no-device\n"); return 0; }
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
  return 0;
}
