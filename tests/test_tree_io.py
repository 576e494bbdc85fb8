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


# ---------------------------------------------------------------------------------------------------------------------
# The other birch-beer stopping rules (--max-step, --max-proportion, --min-distance, --min-distance-search, --custom-cut,
# --root-cut, --smart-cutoff).  The reference's golden files only exercise --min-size, and birch-beer is not vendored, so these
# tests check the semantics the flag help texts state (Options.hs:176-184) as properties on the golden tree.
def _leaf_cells_in_order(tree):
    import numpy as np
    out, stack = [], [0]
    while stack:
        u = stack.pop()
        if tree.left[u] < 0:
            out.extend(int(r) for r in tree.items(u))
        else:
            stack.append(int(tree.right[u])); stack.append(int(tree.left[u]))
    return out


def _check_is_pruning_of(pruned, tree):
    """Every pruned tree keeps all cells in the same DFS-leaf order and pre-order ids with children [left = id + 1]."""
    assert _leaf_cells_in_order(pruned) == _leaf_cells_in_order(tree)
    for i in range(pruned.n_nodes):
        if pruned.left[i] >= 0:
            assert pruned.left[i] == i + 1 and pruned.right[i] > pruned.left[i]


def test_prune_max_step(golden_tree):
    _, tree, _ = golden_tree
    depth0 = T.node_depths(tree)
    for s in (0, 1, 3, 50):
        p = T.prune_max_step(tree, s)
        _check_is_pruning_of(p, tree)
        d = T.node_depths(p)
        assert d.max() == min(s, depth0.max())
        assert sum(1 for i in range(p.n_nodes) if p.left[i] >= 0 and d[i] >= s) == 0
    assert T.prune_max_step(tree, 0).n_nodes == 1 and T.prune_max_step(tree, 50).n_nodes == tree.n_nodes


def test_prune_min_distance_and_search(golden_tree):
    _, tree, _ = golden_tree
    import numpy as np
    qs = sorted(float(tree.q[i]) for i in range(tree.n_nodes) if tree.left[i] >= 0)
    t = qs[len(qs) // 2]
    p = T.prune_min_distance(tree, t)
    _check_is_pruning_of(p, tree)
    par = T.parents_of(p)
    for i in range(p.n_nodes):                      # an inner vertex survives only below a split of distance >= t
        if p.left[i] >= 0 and par[i] >= 0:
            assert p.q[par[i]] >= t
    assert 1 < p.n_nodes < tree.n_nodes
    assert T.prune_min_distance(tree, -1.0).n_nodes == tree.n_nodes          # nothing is below the cutoff
    s = T.prune_min_distance_search(tree, t)
    _check_is_pruning_of(s, tree)
    for i in range(s.n_nodes - 1, -1, -1):          # every surviving split leads to (or is) a split of distance >= t
        if s.left[i] >= 0:
            stack, ok = [i], False
            while stack and not ok:
                u = stack.pop()
                if s.left[u] >= 0:
                    ok = s.q[u] >= t
                    stack += [int(s.left[u]), int(s.right[u])]
            assert ok
    assert s.n_nodes >= 1 and T.prune_min_distance_search(s, t).n_nodes == s.n_nodes      # idempotent
    assert T.prune_min_distance_search(tree, max(qs) + 1.0).n_nodes == 1


def test_prune_max_proportion(golden_tree):
    _, tree, _ = golden_tree
    import numpy as np
    p = T.prune_max_proportion(tree, 0.5)
    _check_is_pruning_of(p, tree)
    prop, par = T.split_proportions(p), T.parents_of(p)
    for i in range(p.n_nodes):                      # below a split more lopsided than 1:2 nothing is split further
        if p.left[i] >= 0 and par[i] >= 0:
            assert prop[par[i]] <= 1.0
    assert T.prune_max_proportion(tree, 1e-9).n_nodes == tree.n_nodes


def test_custom_cut_and_root_cut(golden_tree):
    _, tree, _ = golden_tree
    inner = [i for i in range(1, tree.n_nodes) if tree.left[i] >= 0][:3]
    p = T.prune_custom_cut(tree, inner[:1])
    _check_is_pruning_of(p, tree)
    sizes = T.subtree_sizes(tree)
    assert p.n_nodes < tree.n_nodes and any(p.left[i] < 0 and len(p.items(i)) == sizes[inner[0]] for i in range(p.n_nodes))
    r = T.root_cut(tree, inner[1])
    assert T.subtree_sizes(r)[0] == sizes[inner[1]] and r.q[0] == tree.q[inner[1]]
    assert T.root_cut(tree, 0).n_nodes == tree.n_nodes


def test_smart_cutoff(golden_tree):
    _, tree, _ = golden_tree
    assert T.smart_cutoff([1.0, 2.0, 3.0, 4.0, 100.0], 2.0) == 3.0 + 2.0 * 1.0
    c = T.smart_cutoffs(tree, 1.0)
    assert set(c) == {"min_size", "max_proportion", "min_distance", "min_distance_search"}
    assert 1.0 <= c["min_size"] <= 2312 and c["max_proportion"] >= 1.0 and c["min_distance"] > 0
    p = T.prune_min_size(tree, int(round(c["min_size"])))
    _check_is_pruning_of(p, tree)


def test_elbow_cutoff(golden_tree):
    _, tree, _ = golden_tree
    # a convex curve (long flat start, steep end) dips below its chord: the Min elbow sits where it takes off
    conv = [0.0, 0.01, 0.02, 0.03, 0.04, 0.05, 0.5, 1.0]
    assert T.elbow_cutoff(conv, "Min") == 0.05
    # a concave curve (steep start, long plateau) bulges above the chord: the Max elbow is the top-left hump
    conc = [0.0, 0.5, 0.95, 0.96, 0.97, 0.98, 0.99, 1.0]
    assert T.elbow_cutoff(conc, "Max") == 0.95
    assert T.elbow_cutoff(list(reversed(conc)), "Max") == 0.95                    # order of the input does not matter
    assert T.elbow_cutoff([10 * v for v in conc], "Max") == 9.5                   # nor do the units
    assert T.elbow_cutoff([3.0, 3.0, 3.0], "Max") == 3.0 and T.elbow_cutoff([1.0, 2.0], "Min") == 1.0
    with pytest.raises(ValueError):
        T.elbow_cutoff(conc, "Median")
    c = T.elbow_cutoffs(tree, "Max")
    assert set(c) == {"min_size", "max_proportion", "min_distance", "min_distance_search"}
    inner_q = [float(tree.q[i]) for i in range(tree.n_nodes) if tree.left[i] >= 0]
    assert min(inner_q) <= c["min_distance"] <= max(inner_q)
    _check_is_pruning_of(T.prune_min_distance(tree, c["min_distance"]), tree)


def _oracle_as_clustering_tree(ot):
    """OracleTree -> the flat ClusteringTree the engine returns (pre-order ids are the same; perm = leaves in DFS order)."""
    import numpy as np
    from too_many_cells_b200.spectral import ClusteringTree
    n = ot.n_nodes
    ib, ie, perm = np.zeros(n, np.int64), np.zeros(n, np.int64), []
    for i in range(n):
        ib[i] = len(perm)
        if ot.left[i] < 0:
            perm.extend(int(x) for x in ot.items[i])
    for i in range(n):
        ie[i] = ib[i] + len(ot.items[i])
    par = np.full(n, -1, np.int64)
    for i in range(n):
        if ot.left[i] >= 0:
            par[ot.left[i]] = par[ot.right[i]] = i
    pv = np.array([np.nan if p is None else p for p in ot.p_val])
    return ClusteringTree(left=np.array(ot.left), right=np.array(ot.right), parent=par, q=np.array(ot.q), theta=np.array(ot.theta),
                          iters=np.zeros(n, np.int32), item_begin=ib, item_end=ie, perm=np.array(perm, np.int64), stats={}, p_val=pv)


def test_make_tree_driver_flag_plumbing(tmp_path, monkeypatch, capsys, golden_tiny):
    """tools/make_tree.py with the engine call replaced by the oracle (this is the CPU suite: the real call is exercised by
    tests/test_gpu_ingest.py): the variant flags reach the engine call, p-values reach the files, stopping rules apply in order."""
    import importlib.util, json
    import numpy as np
    import tmc_oracle as O
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_tree_driver", os.path.join(ROOT, "tools", "make_tree.py"))
    drv = importlib.util.module_from_spec(spec); spec.loader.exec_module(drv)
    a, _ = golden_tiny
    a = a[:150]
    names = [f"C{i:03d}-1" for i in range(a.shape[0])]
    seen = {}

    def fake_cluster(mat, normalize=False, min_modularity=None, eigen_group="SignGroup", num_eigen=None, num_runs=None, **kw):
        seen.update(normalize=normalize, eigen_group=eigen_group, num_eigen=num_eigen, num_runs=num_runs, dense=isinstance(mat, np.ndarray))
        import scipy.sparse as sp
        return _oracle_as_clustering_tree(O.hierarchical_spectral_cluster(
            sp.csr_matrix(mat), normalize=normalize, min_modularity=min_modularity or 0.0, eigen_group=eigen_group,
            num_eigen=num_eigen or 1, num_runs=num_runs or 0))

    monkeypatch.setattr(drv.T, "load_all", lambda paths, col: (a, names, [f"G{j}" for j in range(a.shape[1])]))
    monkeypatch.setattr(drv.T, "hierarchical_spectral_cluster", fake_cluster)
    monkeypatch.setattr(drv.T, "tfidf_scale_sparse_mat", O.b1_to_b2)
    out = tmp_path / "o1"
    flat = drv.main(["--matrix-path", "x", "--output", str(out), "--numRuns", "6", "--eigen-group", "KMeansGroup", "--num-eigen", "2"])
    capsys.readouterr()
    assert seen == dict(normalize=True, eigen_group="KMeansGroup", num_eigen=2, num_runs=6, dense=False)
    rows = list(csv.reader(open(out / "node_info.csv", newline="")))
    assert rows[0][4] == "significance"
    inner = [r for r in rows[1:] if r[3] != ""]
    assert inner and all(r[4] != "" for r in inner) and all(r[4] == "" for r in rows[1:] if r[3] == "")
    nested = json.loads((out / "cluster_tree.json").read_text())
    assert nested[0]["_significance"] is not None and flat.p_val is not None
    # --dense hands the engine the tf-idf matrix as a dense array and must not ask for idf twice
    flat_d = drv.main(["--matrix-path", "x", "--output", str(tmp_path / "o2"), "--dense"])
    capsys.readouterr()
    assert seen["dense"] and seen["normalize"] is False and seen["num_runs"] is None
    flat_s = drv.main(["--matrix-path", "x", "--output", str(tmp_path / "o3")])
    capsys.readouterr()
    assert np.array_equal(flat_d.left, flat_s.left)
    # stopping rules: a data-driven cutoff replaces the number on the flag; custom cuts and the root cut come first
    full = flat_s
    f1 = drv.main(["--matrix-path", "x", "--output", str(tmp_path / "o4"), "--min-distance", "0", "--elbow-cutoff", "Max"])
    capsys.readouterr()
    cut = T.elbow_cutoffs(full, "Max")["min_distance"]
    assert np.array_equal(f1.left, T.prune_min_distance(full, cut).left)
    f2 = drv.main(["--matrix-path", "x", "--output", str(tmp_path / "o5"), "--min-size", "1", "--smart-cutoff", "1.0"])
    capsys.readouterr()
    assert np.array_equal(f2.left, T.prune_min_size(full, int(round(T.smart_cutoffs(full, 1.0)["min_size"]))).left)
    first_inner_child = int(full.left[0]) if full.left[int(full.left[0])] >= 0 else int(full.right[0])
    f3 = drv.main(["--matrix-path", "x", "--output", str(tmp_path / "o6"), "--root-cut", str(first_inner_child), "--max-step", "1"])
    capsys.readouterr()
    assert f3.n_nodes == 3


def test_merge_single_cells_like_the_semigroup_instance():
    """Several --matrix-path (Matrix/Types.hs:214-254): equal feature lists stack as they are, different ones go to the ascending
    union of the names."""
    import numpy as np
    import scipy.sparse as sp
    from too_many_cells_b200 import merge_single_cells
    a = sp.csr_matrix(np.array([[1.0, 0, 2], [0, 3, 0]]))
    b = sp.csr_matrix(np.array([[0, 4.0, 0]]))
    m, bc, ft = merge_single_cells([(a, ["a1", "a2"], ["g1", "g2", "g3"]), (b, ["b1"], ["g1", "g2", "g3"])])
    assert bc == ["a1", "a2", "b1"] and ft == ["g1", "g2", "g3"]
    assert np.array_equal(m.toarray(), np.array([[1.0, 0, 2], [0, 3, 0], [0, 4, 0]]))
    c = sp.csr_matrix(np.array([[5.0, 6.0]]))
    m, bc, ft = merge_single_cells([(a, ["a1", "a2"], ["g3", "g1", "g2"]), (c, ["c1"], ["g0", "g2"])])
    assert ft == ["g0", "g1", "g2", "g3"] and bc == ["a1", "a2", "c1"]
    #                  g0  g1  g2  g3      a: g3=col0, g1=col1, g2=col2
    assert np.array_equal(m.toarray(), np.array([[0, 0, 2.0, 1.0], [0, 3.0, 0, 0], [5.0, 0, 6.0, 0]]))
    assert m.has_sorted_indices
    one = merge_single_cells([(a, ["a1", "a2"], ["g1", "g2", "g3"])])
    assert np.array_equal(one[0].toarray(), a.toarray())
    with pytest.raises(ValueError):
        merge_single_cells([])


def test_workshop_parity_comparator_on_the_golden_tree(golden_tree):
    """tools/workshop_parity.py's comparison, fed the golden tree against itself and against a perturbed copy (the tool itself
    needs the two public 10x tarballs of the workshop, which are not in this image: tests/test_gpu_ingest.py runs it when
    TMC_WORKSHOP_DATA points at them)."""
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("workshop_parity", os.path.join(ROOT, "tools", "workshop_parity.py"))
    wp = importlib.util.module_from_spec(spec); spec.loader.exec_module(wp)
    _, tree, names = golden_tree
    barcodes = [names[r] for r in range(len(names))]
    rep = wp.compare(tree, barcodes, tree, names)
    assert rep["cells_common"] == 2312 and rep["row_numbers_equal"] and rep["leaves_identical"] and rep["topology_identical"]
    assert rep["cells_in_a_different_leaf"] == 0 and rep["q_max_rel_diff"] == 0.0 and rep["q_compared"] == 115
    pruned = T.prune_min_size(tree, 30)
    rep = wp.compare(pruned, barcodes, tree, names)
    assert not rep["leaves_identical"] and not rep["topology_identical"] and rep["vertices_shared"] >= pruned.n_nodes - len(
        [i for i in range(pruned.n_nodes) if pruned.left[i] < 0])


def test_make_tree_driver_prior_reproduces_the_reference_workshop_files(tmp_path, capsys):
    """The reference's own `--prior` runs (workshop/workshop.org:253-330): from the golden `out/cluster_tree.json` the driver must
    write, byte for byte, the files the reference wrote — the unpruned set and the `--min-size 30` set (golden-era format: the
    shipped files predate the significance fields)."""
    import importlib.util
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("make_tree_driver", os.path.join(ROOT, "tools", "make_tree.py"))
    drv = importlib.util.module_from_spec(spec); spec.loader.exec_module(drv)
    rd = lambda p: open(p, newline="").read()
    labels = os.path.join(W, "labels.csv")
    out = tmp_path / "plain"
    drv.main(["--prior", W, "--output", str(out), "--labels-file", labels, "--legacy-format"])
    assert capsys.readouterr().out == rd(os.path.join(W, "clusters.csv"))
    for ours, golden in (("cluster_tree.json", "cluster_tree.json"), ("cluster_info.csv", "cluster_info.csv"),
                         ("node_info.csv", "node_info.csv"), ("graph.dot", "graph.dot"), ("clusters.csv", "clusters.csv")):
        assert rd(out / ours) == rd(os.path.join(W, golden)), ours
    assert rd(out / "cluster_list.json") == gzip.open(os.path.join(W, "cluster_list.json.gz"), "rt").read()
    out = tmp_path / "pruned"
    drv.main(["--prior", W, "--output", str(out), "--labels-file", labels, "--legacy-format", "--min-size", "30"])
    assert capsys.readouterr().out == rd(os.path.join(W, "clusters_pruned.csv"))
    for ours, golden in (("cluster_tree.json", "pruned_cluster_tree.json"), ("cluster_info.csv", "pruned_cluster_info.csv"),
                         ("node_info.csv", "pruned_node_info.csv")):
        assert rd(out / ours) == rd(os.path.join(W, golden)), ours
    # current format: the same run carries the significance fields (all Nothing here: no permutation test was stored)
    drv.main(["--prior", W, "--output", str(tmp_path / "cur")])
    capsys.readouterr()
    assert rd(tmp_path / "cur" / "node_info.csv").startswith("node,size,proportion,modularity,significance,composition,subtree")
    assert '"_significance":null' in rd(tmp_path / "cur" / "cluster_tree.json")


def test_write_matrix_like_round_trips_as_a_10x_directory(tmp_path):
    """--matrix-output FOLDER (Matrix/Utility.hs:220-250): features are rows in matrix.mtx.gz, features.tsv.gz has the reference's
    three columns; an independent reader (scipy) gets the same matrix back."""
    import numpy as np
    import scipy.io, scipy.sparse as sp
    from too_many_cells_b200 import write_matrix_like
    rng = np.random.default_rng(5)
    d = ((rng.random((7, 5)) < 0.5) * rng.integers(1, 9, (7, 5))).astype(np.float64)
    d[2, 3] = 0.1 + 0.2                                                          # a value that needs all its digits
    a = sp.csr_matrix(d)
    bcs, feats = [f"BC{i}-1" for i in range(7)], [f"Gene{j}" for j in range(5)]
    write_matrix_like(tmp_path / "mo", a, bcs, feats)
    back = sp.csr_matrix(scipy.io.mmread(str(tmp_path / "mo" / "matrix.mtx.gz"), spmatrix=True))
    assert back.shape == (5, 7) and np.array_equal(back.T.toarray(), a.toarray())
    assert gzip.open(tmp_path / "mo" / "features.tsv.gz", "rt").read().splitlines() == [f"{g}\t{g}\tNA" for g in feats]
    assert gzip.open(tmp_path / "mo" / "barcodes.tsv.gz", "rt").read().splitlines() == bcs
    with pytest.raises(ValueError):
        write_matrix_like(tmp_path / "bad", a, bcs[:-1], feats)
