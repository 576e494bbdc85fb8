// Compiled-host example / test of include/tmc.hpp: reads a CSR (text), runs tfidfScaleSparseMat + hSpecClust through the
// C ABI and prints the make-tree stdout CSV (cell,cluster,path).  Without a CUDA device it reports TMC_ERR_NO_DEVICE.
#include <cstdio>
#include <fstream>
#include <iostream>
#include "../../include/tmc.hpp"

int main(int argc, char** argv) {
  if (argc < 2) { std::fprintf(stderr, "usage: %s matrix.txt\n", argv[0]); return 2; }
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
