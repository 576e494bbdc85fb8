// C++ host layer over the C ABI (include/tmc.h), mirroring the reference's Haskell call surface for the make-tree hot
// path name by name — the reference is compiled code and its own toolchain (GHC) is not in this image, so this is the
// typed host side a compiled caller links; the Python mirror (too-many-cells_b200/spectral.py) wraps the same ABI.
//
//   tmc::tfidfScaleSparseMat            src/TooManyCells/Matrix/Preprocess.hs:204-205      (MatObsRow . unB2 . b1ToB2 . B1)
//   tmc::b1ToB2 / b2ToB                 Math.Clustering.Spectral.Sparse                    (spectral-clustering)
//   tmc::hierarchicalSpectralCluster    Math.Clustering.Hierarchical.Spectral.Sparse, called at
//                                       src/TooManyCells/MakeTree/Cluster.hs:147-156,169
//   tmc::hSpecClust                     src/TooManyCells/MakeTree/Cluster.hs:135-187       (ClusterResults: clusterList + tree)
//   tmc::ClusteringVertex / Tree        MakeTree/Types.hs:124-130 (ClusteringVertex{_clusteringItems,_ngMod,_pVal})
//
// Errors: the reference calls `error` (a Haskell exception); here every failure is a tmc::Error carrying the tmc_status.
// Header-only; link with -ltmc (libtmc.so).  No CPU fallback: without a CUDA device the calls throw TMC_ERR_NO_DEVICE.
#pragma once
#include <cmath>
#include <cstdint>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tmc.h"

namespace tmc {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error("libtmc error " + std::to_string(s) + ": " + m), status(s) {}
};
inline void check(int st) { if (st != TMC_OK) throw Error(st, tmc_last_error()); }

// cells x features sparse matrix, rows = cells (MatObsRow, Matrix/Types.hs:152-158)
struct SpMatrix {
  int64_t nrows = 0, ncols = 0;
  std::vector<int64_t> rowPtr{0};
  std::vector<int32_t> colIdx;
  std::vector<double> values;
  tmc_csr view() const { return tmc_csr{nrows, ncols, (int64_t)values.size(), rowPtr.data(), colIdx.data(), values.data(), 0, 0}; }
};
struct B1 { SpMatrix unB1; };   // raw
struct B2 { SpMatrix unB2; };   // idf-scaled
struct B  { SpMatrix unB;  };   // row-L2-normalised

enum class EigenGroup { SignGroup, KMeansGroup };          // Options.hs:172
struct Q { double unQ; };                                    // Math.Modularity.Types.Q

// b1ToB2: idf column scaling
inline B2 b1ToB2(const B1& b1) {
  B2 out{b1.unB1};
  tmc_csr v = b1.unB1.view();
  check(tmc_tfidf(&v, out.unB2.values.data(), nullptr));
  return out;
}
// b2ToB: L2-normalise every row
inline B b2ToB(const B2& b2) {
  B out{b2.unB2};
  tmc_csr v = b2.unB2.view();
  check(tmc_row_l2(&v, out.unB.values.data(), nullptr));
  return out;
}
// Preprocess.hs:204-205
inline SpMatrix tfidfScaleSparseMat(const SpMatrix& m) { return b1ToB2(B1{m}).unB2; }

template <class A> struct ClusteringVertex {
  std::vector<A> _clusteringItems;
  Q _ngMod;
  std::optional<double> _pVal;
};
template <class A> struct Tree {                              // Data.Tree: rootLabel + subForest
  ClusteringVertex<A> rootLabel;
  std::vector<Tree<A>> subForest;
};
template <class A> using ClusteringTree = Tree<A>;

// hierarchicalSpectralCluster eigenGroup normFlag numEigen minSize minMod runs items (Left mat)
// (std::nullopt = Haskell Nothing).  Children order [label 0, label 1]; every vertex carries its items and Q.
namespace detail {
inline tmc_opts opts(EigenGroup eigenGroup, bool normFlag, std::optional<int> numEigen, std::optional<int> minSize, std::optional<Q> minMod,
                     std::optional<int> runs) {
  tmc_opts o; tmc_opts_default(&o);
  o.eigen_group = eigenGroup == EigenGroup::SignGroup ? 0 : 1;
  o.normalize = normFlag ? 1 : 0;
  o.num_eigen = numEigen.value_or(1);
  o.min_size = minSize.value_or(1);
  o.min_modularity = minMod ? minMod->unQ : 0.0;
  o.num_runs = runs.value_or(0);
  return o;
}
template <class A> ClusteringTree<A> rebuild(tmc_tree* t, const std::vector<A>& items);
}  // namespace detail

// row-major dense cells x features matrix (the hmatrix `Matrix Double` of the --dense path, Cluster.hs:157-167)
struct DenseMatrix { int64_t nrows = 0, ncols = 0; std::vector<double> values; };

template <class A>
ClusteringTree<A> hierarchicalSpectralCluster(EigenGroup eigenGroup, bool normFlag, std::optional<int> numEigen, std::optional<int> minSize,
                                              std::optional<Q> minMod, std::optional<int> runs, const std::vector<A>& items,
                                              const SpMatrix& mat) {
  if ((int64_t)items.size() != mat.nrows) throw Error(TMC_ERR_INVALID, "items and matrix rows differ");
  tmc_opts o = detail::opts(eigenGroup, normFlag, numEigen, minSize, minMod, runs);
  tmc_csr v = mat.view();
  tmc_tree* t = nullptr;
  check(tmc_make_tree(&v, &o, &t));
  return detail::rebuild<A>(t, items);
}
// HSD.hierarchicalSpectralCluster (Math.Clustering.Hierarchical.Spectral.Dense), same arguments on a dense matrix
template <class A>
ClusteringTree<A> hierarchicalSpectralCluster(EigenGroup eigenGroup, bool normFlag, std::optional<int> numEigen, std::optional<int> minSize,
                                              std::optional<Q> minMod, std::optional<int> runs, const std::vector<A>& items,
                                              const DenseMatrix& mat) {
  if ((int64_t)items.size() != mat.nrows) throw Error(TMC_ERR_INVALID, "items and matrix rows differ");
  tmc_opts o = detail::opts(eigenGroup, normFlag, numEigen, minSize, minMod, runs);
  tmc_tree* t = nullptr;
  check(tmc_make_tree_dense(mat.values.data(), mat.nrows, mat.ncols, &o, &t));
  return detail::rebuild<A>(t, items);
}

template <class A> ClusteringTree<A> detail::rebuild(tmc_tree* t, const std::vector<A>& items) {
  std::unique_ptr<tmc_tree, void (*)(tmc_tree*)> guard(t, tmc_tree_free);
  // rebuild the Data.Tree from the flat pre-order arrays (iteratively: trees can be deep)
  std::vector<Tree<A>> built(t->n_nodes);
  for (int64_t i = t->n_nodes - 1; i >= 0; --i) {
    Tree<A>& n = built[i];
    for (int64_t k = t->item_begin[i]; k < t->item_end[i]; ++k) n.rootLabel._clusteringItems.push_back(items[t->perm[k]]);
    n.rootLabel._ngMod = Q{t->q[i]};
    if (t->p_val[i] == t->p_val[i]) n.rootLabel._pVal = t->p_val[i];      // NaN = Nothing
    if (t->left[i] >= 0) { n.subForest.push_back(std::move(built[t->left[i]])); n.subForest.push_back(std::move(built[t->right[i]])); }
  }
  return std::move(built[0]);
}

// CellInfo{_barcode,_cellRow} (Matrix/Types.hs:322-326) and hSpecClust (Cluster.hs:135-187)
struct CellInfo { std::string _barcode; int64_t _cellRow; };
struct ClusterResults {
  std::vector<std::pair<CellInfo, std::vector<int64_t>>> _clusterList;   // (cell, [leaf, parent, ..., 0]) in DFS-leaf order
  ClusteringTree<CellInfo> _clusterDend;
};
inline ClusterResults hSpecClust(EigenGroup eigenGroup, std::optional<int> numEigen, std::optional<Q> minMod, std::optional<int> runs,
                                 const SpMatrix& matrix, const std::vector<std::string>& rowNames) {
  std::vector<CellInfo> items;
  for (int64_t i = 0; i < (int64_t)rowNames.size(); ++i) items.push_back(CellInfo{rowNames[i], i});     // Cluster.hs:143-146
  ClusterResults r;
  r._clusterDend = hierarchicalSpectralCluster<CellInfo>(eigenGroup, false, numEigen, std::nullopt, minMod, runs, items, matrix);
  // treeToGraph numbers vertices in pre-order from 0; leaves with their parent paths (Cluster.hs:171-179)
  struct Frame { const Tree<CellInfo>* n; std::vector<int64_t> path; };
  std::vector<Frame> st; int64_t next = 0;
  st.push_back(Frame{&r._clusterDend, {}});
  while (!st.empty()) {
    Frame f = std::move(st.back()); st.pop_back();
    std::vector<int64_t> path{next++};
    path.insert(path.end(), f.path.begin(), f.path.end());
    if (f.n->subForest.empty()) {
      for (const CellInfo& c : f.n->rootLabel._clusteringItems) r._clusterList.push_back({c, path});
    } else {
      st.push_back(Frame{&f.n->subForest[1], path});
      st.push_back(Frame{&f.n->subForest[0], path});
    }
  }
  return r;
}

}  // namespace tmc
