#!/usr/bin/env python
"""End-to-end test driver that reproduces the I/O of `too-many-cells make-tree` for the clustering part
(src/TooManyCells/Program/MakeTree.hs:77-400) on top of libtmc — NOT a re-implementation of the reference CLI (no plots,
no R, no --prior): it exists so that the whole path  10x directory -> filter -> tf-idf -> tree -> prune -> output files
can be exercised with the reference's flag names and output files.

  python tools/make_tree.py (--matrix-path DIR ... | --prior PREVIOUS_OUT) --output OUT [--filter-thresholds "(250, 1)"] [--normalization TfIdfNorm|NoneNorm]
         [--min-modularity 0] [--min-size N | --max-step N | --max-proportion X | --min-distance X | --min-distance-search X]
         [--smart-cutoff K | --elbow-cutoff Max|Min] [--custom-cut NODE]... [--root-cut NODE]
         [--eigen-group KMeansGroup --num-eigen E] [--numRuns R] [--dense] [--labels-file labels.csv] [--feature-column 1]  > clusters.csv
"""
import argparse, ast, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import too_many_cells_b200 as T
from too_many_cells_b200 import tree_io


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrix-path", action="append")            # repeatable: the matrices are merged (Matrix/Types.hs:214-254)
    ap.add_argument("--prior")                                   # Options.hs:204: folder of a previous run -> skip clustering
    ap.add_argument("--legacy-format", action="store_true")      # driver only: files without the significance fields (golden-era format)
    ap.add_argument("--output", default="out")
    ap.add_argument("--filter-thresholds")                       # Options.hs:118, LoadMatrix.hs:179-183: "(MINCELL, MINFEATURE)"
    ap.add_argument("--normalization", default="TfIdfNorm")     # Program/Utility.hs:33-35: make-tree default
    ap.add_argument("--min-modularity", type=float, default=0.0)  # Options.hs:178 -> the only stop flag that reaches the engine
    ap.add_argument("--min-size", type=int)                      # Options.hs:175 -> post-prune (birch-beer)
    ap.add_argument("--max-step", type=int)                      # Options.hs:176
    ap.add_argument("--max-proportion", type=float)              # Options.hs:177
    ap.add_argument("--min-distance", type=float)                # Options.hs:179
    ap.add_argument("--min-distance-search", type=float)         # Options.hs:180
    ap.add_argument("--smart-cutoff", type=float)                # Options.hs:181: median + DOUBLE * MAD replaces the cutoff given
    ap.add_argument("--elbow-cutoff", choices=["Max", "Min"])    # Options.hs:182: takes precedence over --smart-cutoff
    ap.add_argument("--custom-cut", type=int, action="append", default=[])   # Options.hs:183
    ap.add_argument("--root-cut", type=int)                      # Options.hs:184
    ap.add_argument("--eigen-group", default="SignGroup", choices=["SignGroup", "KMeansGroup"])   # Options.hs:172
    ap.add_argument("--num-eigen", type=int)                     # Options.hs:173
    ap.add_argument("--numRuns", type=int, dest="num_runs")      # Options.hs:174 -> significance column / _significance
    ap.add_argument("--dense", action="store_true")              # Options.hs:209 -> tmc_make_tree_dense
    ap.add_argument("--matrix-output")                           # Options.hs:186: the filtered matrix as a 10x folder under --output
    ap.add_argument("--labels-file")
    ap.add_argument("--feature-column", type=int, default=1)
    a = ap.parse_args(argv)
    if a.prior:
        # `--prior`: "skips clustering by using the previous clustering files" (Program/MakeTree.hs:213-225 loads cluster_tree.json);
        # cells are addressed by the row numbers stored in the tree (the reference's --no-update-tree-rows behaviour)
        flat, barcodes = tree_io.parse_cluster_tree_json(open(os.path.join(a.prior, "cluster_tree.json")).read())
        return finish(a, flat, barcodes)
    if not a.matrix_path:
        ap.error("--matrix-path is required without --prior")
    mat, barcodes, features = T.load_all(a.matrix_path, a.feature_column)
    if a.filter_thresholds:
        rt, ct = ast.literal_eval(a.filter_thresholds)
        mat, kr, kc = T.filter_num_sparse_mat(mat, rt, ct)
        barcodes = [barcodes[i] for i in kr]; features = [features[j] for j in kc]
    if a.matrix_output:                                          # "filtered and normalized (not including TfIdfNorm)", Program/MakeTree.hs:342
        T.write_matrix_like(os.path.join(a.output, a.matrix_output), mat, barcodes, features)
    if a.normalization not in ("TfIdfNorm", "NoneNorm"):
        raise SystemExit("only TfIdfNorm (default) and NoneNorm are on the accelerated path")
    if a.dense:                                                  # hSpecCommand True: sparseToHMat, then the dense entry point
        import numpy as np
        mat = np.asarray((T.tfidf_scale_sparse_mat(mat) if a.normalization == "TfIdfNorm" else mat).todense())
    tree = T.hierarchical_spectral_cluster(mat, normalize=(a.normalization == "TfIdfNorm" and not a.dense),
                                           min_modularity=a.min_modularity, eigen_group=a.eigen_group, num_eigen=a.num_eigen,
                                           num_runs=a.num_runs)
    # clusteringTreeToTree drops the Q of leaves; pruning is applied to the finished tree, as birch-beer does
    flat = tree_io.FlatTree(tree.left, tree.right, [q if l >= 0 else float("nan") for q, l in zip(tree.q, tree.left)],
                            [tree.items(i) if tree.left[i] < 0 else None for i in range(tree.n_nodes)],
                            tree.p_val if a.num_runs else None)
    return finish(a, flat, barcodes)


def finish(a, flat, barcodes):
    """Stopping rules, label composition and the output files: everything after the clustering (also the whole --prior path)."""
    # stopping rules in the order of Program/MakeTree.hs:268-302; a data-driven cutoff replaces the number given on the flag
    auto = tree_io.elbow_cutoffs(flat, a.elbow_cutoff) if a.elbow_cutoff else (
        tree_io.smart_cutoffs(flat, a.smart_cutoff) if a.smart_cutoff is not None else {})
    if a.root_cut is not None:
        flat = tree_io.root_cut(flat, a.root_cut)
    if a.custom_cut:
        flat = tree_io.prune_custom_cut(flat, a.custom_cut)
    if a.min_size:
        flat = tree_io.prune_min_size(flat, int(round(auto.get("min_size", a.min_size))))
    if a.max_step is not None:
        flat = tree_io.prune_max_step(flat, a.max_step)
    if a.max_proportion is not None:
        flat = tree_io.prune_max_proportion(flat, auto.get("max_proportion", a.max_proportion))
    if a.min_distance is not None:
        flat = tree_io.prune_min_distance(flat, auto.get("min_distance", a.min_distance))
    if a.min_distance_search is not None:
        flat = tree_io.prune_min_distance_search(flat, auto.get("min_distance_search", a.min_distance_search))
    labels = None
    if a.labels_file:
        bc2row = {b: i for i, b in (barcodes.items() if isinstance(barcodes, dict) else enumerate(barcodes))}
        labels = {bc2row[r[0]]: r[1] for r in list(csv.reader(open(a.labels_file)))[1:] if r[0] in bc2row}
    tree_io.write_make_tree_outputs(flat, barcodes, a.output, labels, significance=not a.legacy_format)
    sys.stdout.write(tree_io.clusters_csv(flat, barcodes))
    return flat


if __name__ == "__main__":
    main()
