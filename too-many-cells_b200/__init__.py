"""B200-native spectral-bipartition engine for `too-many-cells make-tree` (libtmc + host mirror)."""
from . import _lib  # noqa: F401
from .spectral import (DeviceB, ClusteringTree, TmcError, b1_to_b2, b2_to_b, b_to_d, second_left, spectral_cluster,  # noqa: F401
                       get_b_modularity, partition_rows, spmv_pair, bench_spmv_pair, tfidf_scale_sparse_mat,
                       hierarchical_spectral_cluster, h_spec_clust, make_opts, lsa_sparse_mat, permutation_test, tree_signature, tree_signature_strict, tree_vertex_names,
                       svd_sparse_mat, pca_sparse_mat, center_scale_sparse_cell)
from . import tree_io  # noqa: F401
from . import ingest  # noqa: F401,E402
from .ingest import load_cellranger_data, filter_num_sparse_mat, parse_matrix_market, merge_single_cells, load_all, write_matrix_like  # noqa: F401,E402
