#!/usr/bin/env python
"""Known-answer check against the reference's own workshop run — the hook that turns "parity unpinned" into pinned the moment the
two public 10x tarballs are available (there is no network in the build image, and the reference ships only the OUTPUTS of this
run: workshop/out/*, committed under tests/golden/workshop/).

The reference command (workshop/workshop.org:196-202):

    too-many-cells make-tree --matrix-path data/brain/filtered_feature_bc_matrix/ --matrix-path data/heart/filtered_feature_bc_matrix/ \
        --filter-thresholds "(250, 1)" --output out > clusters.csv

with neuron_1k_v3 and heart_1k_v3 (workshop.org:84-131).  This script runs the same steps on libtmc — GPU ingest of both
directories, `mconcat`, --filter-thresholds, tf-idf, the tree — and compares with the golden cluster_tree.json:

  * the cells that survive the filter and their row numbers (`_cellRow`)          -> pins ingest + merge + filter
  * leaf membership and topology (modulo child swap), Q of every matched vertex     -> pins the numeric path

    python tools/workshop_parity.py --brain DIR --heart DIR [--golden tests/golden/workshop] [--json report.json]
"""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np


def compare(tree, barcodes, golden_tree, golden_names):
    """`tree` over rows named `barcodes` vs the parsed golden tree (row -> barcode in golden_names).  Returns a report dict."""
    g_rows = {bc: r for r, bc in golden_names.items()}
    ours = {bc: i for i, bc in enumerate(barcodes)}
    rep = {"cells_ours": len(ours), "cells_golden": len(g_rows), "cells_common": len(set(ours) & set(g_rows)),
           "row_numbers_equal": all(ours.get(bc) == r for bc, r in g_rows.items())}

    def vertex_sets(t, name_of):
        n = t.n_nodes
        sets = [None] * n
        for i in range(n - 1, -1, -1):
            sets[i] = frozenset(name_of(r) for r in t.items(i)) if t.left[i] < 0 else sets[t.left[i]] | sets[t.right[i]]
        return sets
    so = vertex_sets(tree, lambda r: barcodes[int(r)])
    sg = vertex_sets(golden_tree, lambda r: golden_names[int(r)])
    leaves_o = {so[i] for i in range(tree.n_nodes) if tree.left[i] < 0}
    leaves_g = {sg[i] for i in range(golden_tree.n_nodes) if golden_tree.left[i] < 0}
    rep.update(nodes_ours=tree.n_nodes, nodes_golden=golden_tree.n_nodes, leaves_identical=leaves_o == leaves_g,
               leaves_shared=len(leaves_o & leaves_g), vertices_shared=len(set(so) & set(sg)),
               topology_identical=set(so) == set(sg))
    # cells whose leaf differs (the sign-margin cells SURVEY 7.3 predicts for the reference's kappa = 1e-6)
    leaf_of_o = {bc: s for s in leaves_o for bc in s}
    leaf_of_g = {bc: s for s in leaves_g for bc in s}
    rep["cells_in_a_different_leaf"] = sum(1 for bc in leaf_of_g if leaf_of_o.get(bc) != leaf_of_g[bc])
    qo = {so[i]: float(tree.q[i]) for i in range(tree.n_nodes) if tree.left[i] >= 0}
    qg = {sg[i]: float(golden_tree.q[i]) for i in range(golden_tree.n_nodes) if golden_tree.left[i] >= 0}
    rel = [abs(qo[k] - qg[k]) / max(abs(qg[k]), 1e-300) for k in qg if k in qo]
    rep["q_compared"] = len(rel)
    rep["q_max_rel_diff"] = max(rel) if rel else None
    return rep


def main(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--brain", required=True)
    ap.add_argument("--heart", required=True)
    ap.add_argument("--golden", default=os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "workshop"))
    ap.add_argument("--json")
    a = ap.parse_args(argv)
    import too_many_cells_b200 as T
    from too_many_cells_b200 import tree_io
    golden_tree, golden_names = tree_io.parse_cluster_tree_json(open(os.path.join(a.golden, "cluster_tree.json")).read())
    parts = [T.load_cellranger_data(a.brain), T.load_cellranger_data(a.heart)]
    # the golden run's heart barcodes carry the suffix -2 (workshop.org:221-226) while the public tarball has -1
    if any(bc.endswith("-2") for bc in golden_names.values()) and all(bc.endswith("-1") for bc in parts[1][1]):
        parts[1] = (parts[1][0], [bc[:-2] + "-2" for bc in parts[1][1]], parts[1][2])
    mat, barcodes, features = T.merge_single_cells(parts)
    mat, kr, kc = T.filter_num_sparse_mat(mat, 250, 1)
    barcodes = [barcodes[i] for i in kr]
    tree = T.hierarchical_spectral_cluster(mat, normalize=True)
    rep = compare(tree, barcodes, golden_tree, golden_names)
    rep["engine_ms"] = tree.stats["engine_ms"]
    print(json.dumps(rep, indent=1))
    if a.json:
        json.dump(rep, open(a.json, "w"), indent=1)
    return rep


if __name__ == "__main__":
    main()
