#!/usr/bin/env python
"""End-to-end test driver that reproduces the I/O of `too-many-cells make-tree` for the clustering part
(src/TooManyCells/Program/MakeTree.hs:77-400) on top of libtmc — NOT a re-implementation of the reference CLI (no plots,
no R, no --prior): it exists so that the whole path  10x directory -> filter -> tf-idf -> tree -> prune -> output files
can be exercised with the reference's flag names and output files.

  python tools/make_tree.py --matrix-path DIR --output OUT [--filter-thresholds "(250, 1)"] [--normalization TfIdfNorm|NoneNorm]
         [--min-modularity 0] [--min-size N] [--labels-file labels.csv] [--feature-column 1]  > clusters.csv
"""
import argparse, ast, csv, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import too_many_cells_b200 as T
from too_many_cells_b200 import tree_io


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--matrix-path", required=True)
    ap.add_argument("--output", default="out")
    ap.add_argument("--filter-thresholds")                       # Options.hs:118, LoadMatrix.hs:179-183: "(MINCELL, MINFEATURE)"
    ap.add_argument("--normalization", default="TfIdfNorm")     # Program/Utility.hs:33-35: make-tree default
    ap.add_argument("--min-modularity", type=float, default=0.0)  # Options.hs:178 -> the only stop flag that reaches the engine
    ap.add_argument("--min-size", type=int)                      # Options.hs:175 -> post-prune (birch-beer)
    ap.add_argument("--labels-file")
    ap.add_argument("--feature-column", type=int, default=1)
    a = ap.parse_args(argv)
    mat, barcodes, features = T.load_cellranger_data(a.matrix_path, a.feature_column)
    if a.filter_thresholds:
        rt, ct = ast.literal_eval(a.filter_thresholds)
        mat, kr, kc = T.filter_num_sparse_mat(mat, rt, ct)
        barcodes = [barcodes[i] for i in kr]; features = [features[j] for j in kc]
    if a.normalization not in ("TfIdfNorm", "NoneNorm"):
        raise SystemExit("only TfIdfNorm (default) and NoneNorm are on the accelerated path")
    tree = T.hierarchical_spectral_cluster(mat, normalize=(a.normalization == "TfIdfNorm"), min_modularity=a.min_modularity)
    # clusteringTreeToTree drops the Q of leaves; pruning is applied to the finished tree, as birch-beer does
    flat = tree_io.FlatTree(tree.left, tree.right, [q if l >= 0 else float("nan") for q, l in zip(tree.q, tree.left)],
                            [tree.items(i) if tree.left[i] < 0 else None for i in range(tree.n_nodes)])
    if a.min_size:
        flat = tree_io.prune_min_size(flat, a.min_size)
    labels = None
    if a.labels_file:
        bc2row = {b: i for i, b in enumerate(barcodes)}
        labels = {bc2row[r[0]]: r[1] for r in list(csv.reader(open(a.labels_file)))[1:] if r[0] in bc2row}
    tree_io.write_make_tree_outputs(flat, barcodes, a.output, labels)
    sys.stdout.write(tree_io.clusters_csv(flat, barcodes))
    return flat


if __name__ == "__main__":
    main()
