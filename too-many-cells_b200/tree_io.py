"""make-tree output contract on flat pre-order trees.

Byte-exact writers for the artefacts `too-many-cells make-tree` produces from the
ClusteringTree, pinned in tests/test_tree_io.py against the reference's own golden
outputs (workshop/out/*, workshop/clusters.csv):

  cluster_tree.json   aeson Data.Tree encoding, src/TooManyCells/Program/MakeTree.hs:321-330
  cluster_list.json   src/TooManyCells/MakeTree/Types.hs:82-85
  clusters.csv        stdout of make-tree, src/TooManyCells/Program/MakeTree.hs:383-396
  cluster_info.csv    src/TooManyCells/MakeTree/Print.hs:61-88
  node_info.csv       src/TooManyCells/MakeTree/Print.hs:109-154
  graph.dot           fgl/graphviz dump of treeToGraph (pre-order ids)

CSV bodies use cassava's CRLF record terminator after an LF-terminated header, as the
reference does.  Doubles use Haskell `show` (shortest round-trip digits).
"""
from __future__ import annotations

import json
from collections import Counter
from decimal import Decimal

import numpy as np


def haskell_show_double(x: float) -> str:
    """Haskell `show :: Double -> String` (also what aeson emits for the golden files)."""
    x = float(x)
    if x != x:
        return "NaN"
    if x in (float("inf"), float("-inf")):
        return "Infinity" if x > 0 else "-Infinity"
    if x == 0:
        return "-0.0" if str(x).startswith("-") else "0.0"
    sign = "-" if x < 0 else ""
    t = Decimal(repr(abs(x))).as_tuple()
    digits = "".join(map(str, t.digits)).rstrip("0") or "0"
    # value = 0.d1d2... * 10^e
    e = len(t.digits) + t.exponent
    if 0.1 <= abs(x) < 1e7:
        if e <= 0:
            s = "0." + "0" * (-e) + digits
        elif e >= len(digits):
            s = digits + "0" * (e - len(digits)) + ".0"
        else:
            s = digits[:e] + "." + digits[e:]
        return sign + s
    mant = digits[0] + "." + (digits[1:] if len(digits) > 1 else "0")
    return f"{sign}{mant}e{e - 1}"


# ----------------------------------------------------------------------------- flat tree <-> nested
class FlatTree:
    """Minimal flat pre-order tree: what the writers need (left/right/q/items)."""

    keeps_leaf_distance = True    # a leaf made by pruning keeps the Q of the node it collapsed (golden out_pruned tree)

    def __init__(self, left, right, q, items, p_val=None):
        self.left = np.asarray(left, dtype=np.int64)
        self.right = np.asarray(right, dtype=np.int64)
        self.q = np.asarray(q, dtype=np.float64)
        self.p_val = None if p_val is None else np.asarray(p_val, dtype=np.float64)   # _significance of the kept splits
        self._items = items           # leaf id -> sequence of row ids (None for inner nodes is fine)

    @property
    def n_nodes(self):
        return int(self.left.size)

    def items(self, i):
        return self._items[i]


def _as_flat(tree):
    if hasattr(tree, "items") and hasattr(tree, "left"):
        return tree
    raise TypeError("expected a flat tree (ClusteringTree / FlatTree)")


def parents_of(tree):
    par = np.full(tree.n_nodes, -1, dtype=np.int64)
    for i in range(tree.n_nodes):
        if tree.left[i] >= 0:
            par[tree.left[i]] = i
            par[tree.right[i]] = i
    return par


def leaf_paths(tree):
    """DFS-leaf order list of (leaf, [leaf, parent, ..., 0])  (getGraphLeavesWithParents gr 0)."""
    par = parents_of(tree)
    out = []
    for i in range(tree.n_nodes):       # pre-order ids => ascending id over leaves is DFS order
        if tree.left[i] < 0:
            path, u = [], i
            while u >= 0:
                path.append(int(u))
                u = par[u]
            out.append((i, path))
    return out


def subtree_sizes(tree):
    size = np.zeros(tree.n_nodes, dtype=np.int64)
    for i in range(tree.n_nodes - 1, -1, -1):
        if tree.left[i] < 0:
            size[i] = len(tree.items(i))
        else:
            size[i] = size[tree.left[i]] + size[tree.right[i]]
    return size


def _cell_info(barcode, row):
    return {"_barcode": {"unCell": barcode}, "_cellRow": {"unRow": int(row)}}


class _HsFloat(float):
    pass


def _dumps(obj) -> str:
    """Compact JSON with Haskell-show doubles (aeson's encoding of Double in the golden files)."""
    if obj is None:
        return "null"
    if isinstance(obj, bool):
        return "true" if obj else "false"
    if isinstance(obj, (int, np.integer)):
        return str(int(obj))
    if isinstance(obj, float):
        return haskell_show_double(obj)
    if isinstance(obj, str):
        return json.dumps(obj, ensure_ascii=False)
    if isinstance(obj, (list, tuple)):
        return "[" + ",".join(_dumps(x) for x in obj) + "]"
    if isinstance(obj, dict):
        return "{" + ",".join(json.dumps(k) + ":" + _dumps(v) for k, v in obj.items()) + "}"
    raise TypeError(type(obj))


def _p_val(tree, i):
    """`_significance` of vertex i: the split's permutation p-value (`_pVal`), None = Nothing (leaf / no test run)."""
    p = getattr(tree, "p_val", None)
    if p is None or tree.left[i] < 0 or not (p[i] == p[i]):
        return None
    return float(p[i])


def cluster_tree_json(tree, row_names, significance=False) -> str:
    """`[label,[children]]`; inner: _item null, _distance Q; leaf: _item cells, _distance null.
    `significance=True` adds the `_significance` key current birch-beer emits (absent in the golden file)."""
    tree = _as_flat(tree)
    # iterative post-order string assembly (trees can be ~1e5 nodes deep-ish)
    strs = [None] * tree.n_nodes
    for i in range(tree.n_nodes - 1, -1, -1):
        if tree.left[i] < 0:
            dist = None
            if getattr(tree, "keeps_leaf_distance", False) and tree.q[i] == tree.q[i]:
                dist = float(tree.q[i])
            label = {"_item": [_cell_info(row_names[r], r) for r in tree.items(i)], "_distance": dist}
            kids = "[]"
        else:
            label = {"_item": None, "_distance": float(tree.q[i])}
            kids = "[" + strs[tree.left[i]] + "," + strs[tree.right[i]] + "]"
            strs[tree.left[i]] = strs[tree.right[i]] = None
        if significance:
            label["_significance"] = _p_val(tree, i)
        strs[i] = "[" + _dumps(label) + "," + kids + "]"
    return strs[0]


def parse_cluster_tree_json(text: str):
    """Inverse of cluster_tree_json: returns (FlatTree, row_names dict row->barcode)."""
    nested = json.loads(text)
    left, right, q, items, names = [], [], [], [], {}
    stack = [(nested, -1, 0)]
    # explicit pre-order walk, children [0] then [1]
    def walk(node):
        order = []
        st = [node]
        while st:
            n = st.pop()
            order.append(n)
            for ch in reversed(n[1]):
                st.append(ch)
        return order
    order = walk(nested)
    ids = {id(n): i for i, n in enumerate(order)}
    pv, any_p = [], False
    for n in order:
        label, kids = n
        sig = label.get("_significance")          # current birch-beer; absent in the shipped golden file (--prior round trip)
        pv.append(float("nan") if sig is None else float(sig)); any_p = any_p or sig is not None
        if kids:
            assert len(kids) == 2
            left.append(ids[id(kids[0])]); right.append(ids[id(kids[1])])
            q.append(label["_distance"]); items.append(None)
        else:
            left.append(-1); right.append(-1)
            q.append(float("nan") if label["_distance"] is None else label["_distance"])
            rows = []
            for ci in label["_item"]:
                rows.append(ci["_cellRow"]["unRow"]); names[ci["_cellRow"]["unRow"]] = ci["_barcode"]["unCell"]
            items.append(rows)
    return FlatTree(left, right, q, items, pv if any_p else None), names


def cluster_list(tree, row_names):
    """[(barcode, row, [leaf,...,0])] in DFS-leaf order, item order inside the leaf (treeToClusterList)."""
    tree = _as_flat(tree)
    out = []
    for leaf, path in leaf_paths(tree):
        for r in tree.items(leaf):
            out.append((row_names[r], int(r), path))
    return out


def cluster_list_json(tree, row_names) -> str:
    parts = []
    for bc, r, path in cluster_list(tree, row_names):
        parts.append("[" + _dumps(_cell_info(bc, r)) + "," + _dumps([{"unCluster": c} for c in path]) + "]")
    return "[" + ",".join(parts) + "]"


def _csv_field(s: str) -> str:
    if any(ch in s for ch in ',"\r\n'):
        return '"' + s.replace('"', '""') + '"'
    return s


def _csv(header: str, rows) -> str:
    return header + "\n" + "".join(",".join(_csv_field(str(f)) for f in r) + "\r\n" for r in rows)


def clusters_csv(tree, row_names) -> str:
    """stdout of make-tree: header `cell,cluster,path` (Program/MakeTree.hs:383-396)."""
    rows = ((bc, path[0], "/".join(map(str, path))) for bc, _, path in cluster_list(tree, row_names))
    return _csv("cell,cluster,path", rows) + "\n"       # B.putStrLn appends the final LF


def cluster_info_csv(tree) -> str:
    """cluster,modularity,size — modularity path starts at the leaf's parent (Print.hs:61-88)."""
    tree = _as_flat(tree)
    size = subtree_sizes(tree)
    rows = []
    for leaf, path in leaf_paths(tree):
        qs = [haskell_show_double(tree.q[u]) for u in path if tree.left[u] >= 0]
        rows.append((leaf, "/".join(qs), "/".join(str(int(size[u])) for u in path)))
    return _csv("cluster,modularity,size", rows)


def node_info_csv(tree, label_of_row=None, significance=True) -> str:
    """node,size,proportion,modularity,[significance,]composition,subtree (Print.hs:109-154).
    `significance=False` reproduces the older header of the shipped golden file."""
    tree = _as_flat(tree)
    n = tree.n_nodes
    size = subtree_sizes(tree)
    # subtree node sets: pre-order => node i's subtree is the id range [i, i + count_i)
    count = np.ones(n, dtype=np.int64)
    for i in range(n - 1, -1, -1):
        if tree.left[i] >= 0:
            count[i] += count[tree.left[i]] + count[tree.right[i]]
    comp = [None] * n
    if label_of_row is not None:
        for i in range(n - 1, -1, -1):
            if tree.left[i] < 0:
                comp[i] = Counter(label_of_row[r] for r in tree.items(i) if r in label_of_row)
            else:
                comp[i] = comp[tree.left[i]] + comp[tree.right[i]]
    rows = []
    for i in range(n):
        inner = tree.left[i] >= 0
        prop = haskell_show_double(size[tree.left[i]] / size[tree.right[i]]) if inner else ""
        mod = haskell_show_double(tree.q[i]) if inner else ""
        if comp[i] is not None:
            tot = sum(comp[i].values())
            cs = "/".join(f"{k}: {haskell_show_double(v / tot)} ({v})" for k, v in sorted(comp[i].items()))
        else:
            cs = ""
        sub = "/".join(str(j) for j in range(i, i + int(count[i])))
        sig = _p_val(tree, i)
        row = [i, int(size[i]), prop, mod] + (["" if sig is None else haskell_show_double(sig)] if significance else []) + [cs, sub]
        rows.append(row)
    hdr = "node,size,proportion,modularity," + ("significance," if significance else "") + "composition,subtree"
    return _csv(hdr, rows)


def _collapse(tree, becomes_leaf) -> FlatTree:
    """Rebuild the tree top-down; an inner vertex u with becomes_leaf(u) True turns into a leaf holding all cells of its subtree
    in DFS-leaf order.  Its `_distance` becomes that of its LEFT child (the Q of the left child if that was an inner vertex, null
    if it was a leaf) -- a quirk of birch-beer's merge that the reference's golden pruned tree shows."""
    tree = _as_flat(tree)
    n = tree.n_nodes
    count = np.ones(n, dtype=np.int64)            # pre-order ids => subtree(i) = ids [i, i + count_i)
    for i in range(n - 1, -1, -1):
        if tree.left[i] >= 0:
            count[i] += count[tree.left[i]] + count[tree.right[i]]
    left, right, q, items, pv = [], [], [], [], []
    src_p = getattr(tree, "p_val", None)
    stack = [(0, -1)]
    while stack:
        u, parent = stack.pop()
        i = len(left)
        left.append(-1); right.append(-1); q.append(float("nan")); items.append(None); pv.append(float("nan"))
        if parent >= 0:
            if left[parent] < 0:
                left[parent] = i
            else:
                right[parent] = i
        inner = tree.left[u] >= 0
        if inner and not becomes_leaf(u):
            q[i] = float(tree.q[u])
            if src_p is not None:
                pv[i] = float(src_p[u])
            stack.append((int(tree.right[u]), i)); stack.append((int(tree.left[u]), i))
        else:
            merged = []
            for v in range(u, u + int(count[u])):
                if tree.left[v] < 0:
                    merged.extend(int(r) for r in tree.items(v))
            items[i] = merged
            if inner and tree.left[tree.left[u]] >= 0:
                q[i] = float(tree.q[tree.left[u]])
    return FlatTree(left, right, q, items, pv if src_p is not None else None)


def prune_min_size(tree, min_size: int) -> FlatTree:
    """`--min-size N` (birch-beer, applied after clustering: Program/MakeTree.hs:98,271; Options.hs:175): a vertex whose
    either child holds fewer than N cells becomes a leaf.  Pinned against the reference's golden `workshop/out_pruned/*`
    (made with --min-size 30 from `workshop/out` through --prior)."""
    tree = _as_flat(tree)
    size = subtree_sizes(tree)
    return _collapse(tree, lambda u: size[tree.left[u]] < min_size or size[tree.right[u]] < min_size)


# The remaining birch-beer stopping rules (Program/MakeTree.hs:268-276).  birch-beer itself is an un-vendored dependency: these
# follow the flag help texts (Options.hs:176-184), which are the only in-tree specification, and are NOT pinned to reference
# outputs (the golden files only exercise --min-size).  Same merge as above.
def node_depths(tree):
    tree = _as_flat(tree)
    depth = np.zeros(tree.n_nodes, dtype=np.int64)
    for i in range(tree.n_nodes):
        if tree.left[i] >= 0:
            depth[tree.left[i]] = depth[tree.right[i]] = depth[i] + 1
    return depth


def prune_max_step(tree, max_step: int) -> FlatTree:
    """`--max-step S`: "Only keep clusters that are INT steps from the root" -- vertices S steps below the root become leaves."""
    depth = node_depths(tree)
    return _collapse(tree, lambda u: depth[u] >= max_step)


def split_proportions(tree):
    """|log2(|L| / |R|)| of every inner vertex (nan for leaves): the quantity --max-proportion and its smart cutoff use."""
    tree = _as_flat(tree)
    size = subtree_sizes(tree)
    out = np.full(tree.n_nodes, np.nan)
    for i in range(tree.n_nodes):
        if tree.left[i] >= 0:
            out[i] = abs(np.log2(size[tree.left[i]] / size[tree.right[i]]))
    return out


def prune_max_proportion(tree, max_proportion: float) -> FlatTree:
    """`--max-proportion X`: "stop at the node immediate after a node with DOUBLE proportion split ... at 0.5 if |L| / |R| < 0.5
    or > 2 (absolute log2 transformed) ... Includes L and R in the final result": the children of such a vertex become leaves."""
    tree = _as_flat(tree)
    prop, par = split_proportions(tree), parents_of(tree)
    cut = abs(np.log2(max_proportion))
    return _collapse(tree, lambda u: par[u] >= 0 and prop[par[u]] > cut)


def prune_min_distance(tree, min_distance: float) -> FlatTree:
    """`--min-distance t`: "a node N with L and R children will stop with this criteria [if] the distance at N to L and R is
    < DOUBLE.  Includes L and R in the final result": the children of a vertex with Q < t become leaves."""
    tree = _as_flat(tree)
    par = parents_of(tree)
    return _collapse(tree, lambda u: par[u] >= 0 and tree.q[par[u]] < min_distance)


def prune_min_distance_search(tree, min_distance: float) -> FlatTree:
    """`--min-distance-search t`: "searches from the leaves to the root -- if a path from a subtree contains a distance of at
    least DOUBLE, keep that path, otherwise prune it": a vertex with no split of distance >= t in its subtree becomes a leaf."""
    tree = _as_flat(tree)
    best = np.full(tree.n_nodes, -np.inf)
    for i in range(tree.n_nodes - 1, -1, -1):
        if tree.left[i] >= 0:
            best[i] = max(float(tree.q[i]), best[tree.left[i]], best[tree.right[i]])
    return _collapse(tree, lambda u: best[u] < min_distance)


def prune_custom_cut(tree, nodes) -> FlatTree:
    """`--custom-cut NODE ...`: "List of nodes to prune (make these nodes leaves)"."""
    cut = set(int(x) for x in nodes)
    return _collapse(tree, lambda u: u in cut)


def root_cut(tree, node: int) -> FlatTree:
    """`--root-cut NODE`: "Assign a new root to the tree, removing all nodes outside of the subtree" (ids are renumbered in
    pre-order from the new root)."""
    tree = _as_flat(tree)
    n = tree.n_nodes
    count = np.ones(n, dtype=np.int64)
    for i in range(n - 1, -1, -1):
        if tree.left[i] >= 0:
            count[i] += count[tree.left[i]] + count[tree.right[i]]
    lo, hi = int(node), int(node) + int(count[node])
    rel = lambda v: int(v) - lo if v >= 0 else -1
    return FlatTree([rel(tree.left[v]) for v in range(lo, hi)], [rel(tree.right[v]) for v in range(lo, hi)],
                    [float(tree.q[v]) for v in range(lo, hi)],
                    [list(int(r) for r in tree.items(v)) if tree.left[v] < 0 else None for v in range(lo, hi)],
                    None if getattr(tree, "p_val", None) is None else [float(tree.p_val[v]) for v in range(lo, hi)])


def smart_cutoff(values, n_mads: float) -> float:
    """`--smart-cutoff k`: median + k * MAD of the distribution over all nodes (MAD = median absolute deviation, unscaled)."""
    v = np.asarray([x for x in values if np.isfinite(x)], dtype=np.float64)
    med = float(np.median(v))
    return med + n_mads * float(np.median(np.abs(v - med)))


def smart_cutoffs(tree, n_mads: float) -> dict:
    """The four data-driven cutoffs of `--smart-cutoff` ("--max-proportion and --min-size distributions are log2 transformed"):
    pass the one you want to the corresponding prune_* function."""
    tree = _as_flat(tree)
    size = subtree_sizes(tree)
    inner_q = [float(tree.q[i]) for i in range(tree.n_nodes) if tree.left[i] >= 0]
    return {"min_size": float(2.0 ** smart_cutoff(np.log2(size), n_mads)),
            "max_proportion": float(2.0 ** smart_cutoff(split_proportions(tree), n_mads)),
            "min_distance": smart_cutoff(inner_q, n_mads),
            "min_distance_search": smart_cutoff(inner_q, n_mads)}


def elbow_cutoff(values, kind: str = "Max") -> float:
    """`--elbow-cutoff Max | Min` (Options.hs:184): the elbow point of the distribution of `values`.  birch-beer delegates to the
    un-vendored `elbow` package (Math.Elbow), so the construction is restated from the help text — "for a distribution in positive
    x and y on a graph, the top left hump would be Max and the bottom right dip would be Min": the values are sorted ascending and
    plotted against their rank, the curve is rotated so that the chord from its first to its last point becomes the x axis, and the
    elbow is the point farthest above (Max) or below (Min) the chord; the cutoff is that point's value.  Both axes are scaled to
    [0, 1] first so that the answer does not depend on units.  Unpinned (no golden output exercises it)."""
    if kind not in ("Max", "Min"):
        raise ValueError("--elbow-cutoff takes Max or Min")
    y = np.sort(np.asarray(values, dtype=np.float64))
    if y.size == 0:
        raise ValueError("elbow of an empty distribution")
    if y.size < 3 or y[-1] == y[0]:
        return float(y[0] if kind == "Min" else y[-1])
    x = np.linspace(0.0, 1.0, y.size)
    yn = (y - y[0]) / (y[-1] - y[0])
    # after rotating the chord (0,0)-(1,1) onto the x axis, height above the chord = (yn - x) / sqrt(2)
    h = yn - x
    k = int(np.argmax(h)) if kind == "Max" else int(np.argmin(h))
    return float(y[k])


def elbow_cutoffs(tree, kind: str = "Max") -> dict:
    """The four cutoffs of `--elbow-cutoff` (same distributions as `--smart-cutoff`; it "takes precedent" over it, Options.hs:184)."""
    tree = _as_flat(tree)
    size = subtree_sizes(tree)
    inner_q = [float(tree.q[i]) for i in range(tree.n_nodes) if tree.left[i] >= 0]
    return {"min_size": float(2.0 ** elbow_cutoff(np.log2(size), kind)),
            "max_proportion": float(2.0 ** elbow_cutoff(split_proportions(tree), kind)),
            "min_distance": elbow_cutoff(inner_q, kind),
            "min_distance_search": elbow_cutoff(inner_q, kind)}


def graph_dot(tree) -> str:
    tree = _as_flat(tree)
    lines = ["digraph {"]
    lines += [f"    {i};" for i in range(tree.n_nodes)]
    for i in range(tree.n_nodes):
        if tree.left[i] >= 0:
            lines.append(f"    {i} -> {int(tree.left[i])};")
            lines.append(f"    {i} -> {int(tree.right[i])};")
    return "\n".join(lines) + "\n}"


def write_make_tree_outputs(tree, row_names, out_dir, label_of_row=None, significance=True):
    """Write the files `make-tree --output DIR` writes from the clustering (no plots).  `significance=False` writes the format of
    the program version that made the reference's golden workshop files (no `_significance` key / `significance` column)."""
    import os
    os.makedirs(out_dir, exist_ok=True)
    files = {
        "cluster_tree.json": cluster_tree_json(tree, row_names, significance=significance),
        "cluster_list.json": cluster_list_json(tree, row_names),
        "cluster_info.csv": cluster_info_csv(tree),
        "node_info.csv": node_info_csv(tree, label_of_row, significance=significance),
        "graph.dot": graph_dot(tree),
        "clusters.csv": clusters_csv(tree, row_names),
    }
    for name, text in files.items():
        with open(os.path.join(out_dir, name), "w", newline="") as f:
            f.write(text)
    return sorted(files)
