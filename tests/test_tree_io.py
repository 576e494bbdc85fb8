"""Output contract: the writers must reproduce, byte for byte, the reference's own
golden make-tree outputs (workshop/out/*, workshop/clusters.csv) from the golden cluster_tree.json."""
import csv
import gzip
import os

import pytest

from too_many_cells_b200 import tree_io as T
from conftest import GOLDEN

W = os.path.join(GOLDEN, "workshop")


@pytest.fixture(scope="module")
def golden_tree():
    txt = open(os.path.join(W, "cluster_tree.json")).read()
    tree, names = T.parse_cluster_tree_json(txt)
    return txt, tree, names


def test_tree_shape(golden_tree):
    _, tree, names = golden_tree
    assert tree.n_nodes == 231 and len(names) == 2312
    assert int((tree.left < 0).sum()) == 116
    for i in range(tree.n_nodes):           # leaf items ascending original row (stable splits)
        if tree.left[i] < 0:
            it = tree.items(i)
            assert list(it) == sorted(it)


def test_cluster_tree_json_roundtrip(golden_tree):
    txt, tree, names = golden_tree
    assert T.cluster_tree_json(tree, names) == txt


def test_clusters_csv(golden_tree):
    _, tree, names = golden_tree
    assert T.clusters_csv(tree, names) == open(os.path.join(W, "clusters.csv"), newline="").read()


def test_cluster_list_json(golden_tree):
    _, tree, names = golden_tree
    assert T.cluster_list_json(tree, names) == gzip.open(os.path.join(W, "cluster_list.json.gz"), "rt").read()


def test_cluster_info_csv(golden_tree):
    _, tree, _ = golden_tree
    assert T.cluster_info_csv(tree) == open(os.path.join(W, "cluster_info.csv"), newline="").read()


def test_node_info_csv(golden_tree):
    _, tree, names = golden_tree
    bc2row = {v: k for k, v in names.items()}
    lab = {bc2row[r[0]]: r[1] for r in list(csv.reader(open(os.path.join(W, "labels.csv"))))[1:] if r[0] in bc2row}
    assert T.node_info_csv(tree, lab, significance=False) == open(os.path.join(W, "node_info.csv"), newline="").read()
    cur = T.node_info_csv(tree, lab, significance=True)          # current header (Print.hs:142)
    assert cur.startswith("node,size,proportion,modularity,significance,composition,subtree\n")


def test_graph_dot(golden_tree):
    _, tree, _ = golden_tree
    assert T.graph_dot(tree) == open(os.path.join(W, "graph.dot")).read()


@pytest.mark.parametrize("x,s", [(5.972810594217861e-2, "5.972810594217861e-2"), (11.363636363636363, "11.363636363636363"),
                                 (2.0, "2.0"), (0.1, "0.1"), (1e7, "1.0e7"), (1e-3, "1.0e-3"), (123456.0, "123456.0"),
                                 (7.34341252699784e-2, "7.34341252699784e-2"), (-0.5, "-0.5"), (0.0, "0.0")])
def test_haskell_show(x, s):
    assert T.haskell_show_double(x) == s


def test_prune_min_size_matches_golden_pruned_outputs(golden_tree):
    """--min-size 30 through --prior (workshop/workshop.org:324-330): every written artefact of the pruned run is reproduced."""
    _, tree, names = golden_tree
    pruned = T.prune_min_size(tree, 30)
    assert pruned.n_nodes == 75 and int((pruned.left < 0).sum()) == 38
    assert T.cluster_tree_json(pruned, names) == open(os.path.join(W, "pruned_cluster_tree.json")).read()
    assert T.clusters_csv(pruned, names) == open(os.path.join(W, "clusters_pruned.csv"), newline="").read()
    assert T.cluster_info_csv(pruned) == open(os.path.join(W, "pruned_cluster_info.csv"), newline="").read()
    bc2row = {v: k for k, v in names.items()}
    lab = {bc2row[r[0]]: r[1] for r in list(csv.reader(open(os.path.join(W, "labels.csv"))))[1:] if r[0] in bc2row}
    assert T.node_info_csv(pruned, lab, significance=False) == open(os.path.join(W, "pruned_node_info.csv"), newline="").read()
    assert T.prune_min_size(tree, 1).n_nodes == tree.n_nodes                      # no-op
    assert T.prune_min_size(tree, 10 ** 9).n_nodes == 1                           # everything collapses into the root
