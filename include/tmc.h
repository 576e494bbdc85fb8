/* libtmc — C ABI of the B200-native spectral-bipartition engine for
 * `too-many-cells make-tree`.
 *
 * The reference has no FFI of its own for this path; the seam is the Haskell
 * call in src/TooManyCells/MakeTree/Cluster.hs:147-156,169
 *     hierarchicalSpectralCluster eigenGroup False numEigen Nothing minMod runs items (Left mat)
 * and the one-liner src/TooManyCells/Matrix/Preprocess.hs:204-205
 *     tfidfScaleSparseMat = MatObsRow . unB2 . b1ToB2 . B1 . unMatObsRow
 * Every entry point below names the reference / upstream function it replaces.
 * INTEGRATION.md shows the `foreign import ccall` stubs a maintainer would add.
 *
 * Conventions: plain pointers and sizes only.  Inputs are borrowed HOST memory
 * for the duration of the call unless the name says `_dev`.  Outputs are
 * caller-allocated unless documented as library-owned.  Every function returns
 * 0 on success or a negative tmc_status; tmc_last_error() gives a thread-local
 * message.  Nothing aborts, nothing hangs: NaN / zero-norm rows are detected up
 * front (the reference hangs in SVDLIBC on those,
 * src/TooManyCells/Program/LoadMatrix.hs:189-193).
 * There is NO CPU fallback: without a CUDA device every compute entry point
 * returns TMC_ERR_NO_DEVICE.
 */
#ifndef TMC_H
#define TMC_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  TMC_OK = 0,
  TMC_ERR_INVALID = -1,     /* bad argument (null pointer, unsorted / out-of-range column, ...) */
  TMC_ERR_NO_DEVICE = -2,   /* no CUDA device / driver */
  TMC_ERR_CUDA = -3,        /* CUDA runtime error (message in tmc_last_error) */
  TMC_ERR_ZERO_ROW = -4,    /* a row has zero L2 norm (reference: NaN -> svdlibc hang) */
  TMC_ERR_NAN = -5,         /* NaN / Inf in the values */
  TMC_ERR_NEGATIVE = -6,    /* (no longer raised: signed matrices are accepted -- the degrees use |B| as the reference's bToD does,
                               everything else B itself; kept so that the numbering of the codes does not move) */
  TMC_ERR_NOMEM = -7,
  TMC_ERR_COMM = -8,        /* NCCL error */
  TMC_ERR_UNSUPPORTED = -9  /* e.g. the non-default variants (KMeansGroup, num_runs > 0) under a multi-GPU communicator */
} tmc_status;

/* cells x features CSR, rows = cells (MatObsRow, src/TooManyCells/Matrix/Types.hs:152-158).
 * Columns need not be sorted within a row; duplicates are not allowed. */
typedef enum { TMC_VAL_F64 = 0, TMC_VAL_F32 = 1, TMC_VAL_U16 = 2, TMC_VAL_U8 = 3 } tmc_val_type;
typedef enum { TMC_IDX_I32 = 0, TMC_IDX_U16 = 1 } tmc_idx_type;
typedef struct {
  int64_t n_rows, n_cols, nnz;
  const int64_t* row_ptr;   /* n_rows + 1 */
  const int32_t* col_idx;   /* nnz column indices; element type per idx_type (cast the pointer) */
  const double*  val;       /* nnz values; element type per val_type (cast the pointer) */
  /* multi-GPU row blocks: leave both 0 for a whole matrix.  With a communicator, a whole
   * matrix passed on every rank is sliced into contiguous row blocks; a rank may instead pass only its
   * own block: n_rows = rows in the block, row_offset = global id of its first row. */
  int64_t row_offset, n_rows_global;
  /* Compact host element types (0, 0 = the defaults double / int32).  The reference's values are `Double`s, but the usual
   * make-tree input is a matrix of small UMI counts: a shim that marshals them as uint16 (and gene indices as uint16 when
   * n_cols <= 65536) moves 4-6 instead of 12 bytes per non-zero over PCIe; the values are widened to fp64 on the device, so
   * every result is bit-identical to the fp64 call.  Accepted by tmc_matrix_upload / tmc_make_tree; the entry points that
   * return host values (tmc_tfidf, tmc_row_l2, tmc_left_singular, tmc_filter_thresholds) want the defaults. */
  int32_t val_type;         /* tmc_val_type */
  int32_t idx_type;         /* tmc_idx_type */
} tmc_csr;

/* Flag contract of the make-tree engine call (src/TooManyCells/Program/Options.hs:172-178,
 * src/TooManyCells/MakeTree/Cluster.hs:147-156). Zero-initialise, then tmc_opts_default(). */
typedef struct {
  int32_t eigen_group;      /* 0 = SignGroup (default), 1 = KMeansGroup (2-means on eigenvectors 2..num_eigen+1; with num_eigen = 1 inside
                               the batched sweeps, else a per-node host recursion over the GPU exports)   Options.hs:172 */
  int32_t num_eigen;        /* default 1; only read by KMeansGroup                      Options.hs:173 */
  double  min_modularity;   /* default 0.0; split iff Q > min_modularity                Options.hs:178 */
  int32_t min_size;         /* make-tree always passes 1 (Nothing)                      Cluster.hs:152 */
  int32_t normalize;        /* 0 from make-tree (tf-idf done by caller), 1 = idf inside Cluster.hs:150 */
  int32_t num_runs;         /* permutation test of every split, 0 = off (fills tmc_tree.p_val; batched sweeps)  Options.hs:174 */
  uint64_t seed;            /* start-vector seed; also seeds the label shuffles of the permutation test */
  /* engine knobs (no reference counterpart) */
  double  tol;              /* Lanczos residual tolerance ||M u - theta u|| <= tol*theta; default 1e-10 */
  int32_t max_lanczos;      /* Krylov dimension before an explicit restart; default 320 */
  int32_t max_restarts;     /* default 8 */
  int32_t max_active;       /* concurrently iterating tree nodes (y-vector slots); default 2048 */
  int32_t device;           /* CUDA device ordinal, -1 = current */
  int32_t profile;          /* 1 = time the SpMV kernels with CUDA events (fills tmc_tree.spmv_ms) */
  int32_t scatter_mode;     /* y = B^T x: 0 = column-sorted transposed copy per node (default), 1 = CSR scatter with fp64 atomics */
  int32_t sign_stop;        /* 0 = default: a node's Lanczos iteration stops as soon as every label (the sign of every entry of
                               u2) is certified final -- min_i |x_i| > 4 r / gap for the Ritz vector x (Davis-Kahan) --
                               instead of running on to `tol`; -1 = off (always iterate to `tol`).  The tree is the same either way. */
  int32_t warm_start;       /* 0 = default: a child's Lanczos run starts from the parent's Krylov space (the Ritz vector of the
                               parent's NEXT singular direction, restricted to the child's cells) instead of a pseudo-random
                               vector; -1 = off.  Fewer steps, the same converged vectors, the same tree. */
  int32_t small_rows;       /* nodes of at most this many cells that live on one rank are solved directly (Gram matrix + dense
                               eigen-solve of the whole subtree in one CTA) instead of iterating in the sweeps: 0 = default (64),
                               -1 = off, else the threshold (at most 64).  Same arithmetic, same tree. */
  int32_t reserved[2];
} tmc_opts;

/* Result tree: flat, pre-order (node 0 = root, children [label 0, label 1]) exactly as
 * birch-beer's treeToGraph numbers it (src/TooManyCells/MakeTree/Cluster.hs:171-181), so
 * node ids equal the `cluster` ids of clusters.csv.  Library-owned until tmc_tree_free. */
typedef struct {
  int64_t n_nodes, n_rows;
  const int64_t* left;        /* n_nodes; -1 for a leaf */
  const int64_t* right;
  const int64_t* parent;      /* -1 for the root */
  const double*  q;           /* Newman-Girvan Q of the node's bipartition (also computed for leaves;
                                 the host maps leaf -> Nothing as clusteringTreeToTree does) */
  const double*  theta;       /* sigma_2(C)^2 at the node */
  const int32_t* iters;       /* Lanczos steps spent on the node */
  const int64_t* item_begin;  /* n_nodes; [item_begin,item_end) into perm: the node's cells */
  const int64_t* item_end;
  const int64_t* perm;        /* n_rows original row ids; ascending inside every LEAF (stable splits) */
  /* engine statistics for the roofline report */
  int64_t n_sweeps;           /* kernel sweeps over the active node set */
  int64_t spmv_pairs_nnz;     /* sum over Lanczos steps of nnz touched (one pair = 2 passes) */
  int64_t kernel_launches;
  double  engine_ms;          /* device time of the tree build (CUDA events) */
  double  spmv_ms;            /* device time inside the scatter+gather kernels (only if opts.profile) */
  double  spmv_alg_bytes;     /* algorithmic bytes of those SpMV pairs (BASELINE.md section 4 formula, b_v as stored) */
  int32_t value_bytes;        /* b_v: 8 = B kept as fp64; 2 = raw counts kept as uint16 + diagonal factors; 0 = all-ones matrix */
  int32_t pad_;
  const double*  p_val;       /* n_nodes; _pVal of the split (MakeTree/Types.hs:124-130): NaN = Nothing (leaf, or num_runs = 0) */
} tmc_tree;

/* Device-resident matrix handle: B = row-L2-normalised (tf-idf) CSR kept in HBM. */
typedef struct tmc_matrix tmc_matrix;

void        tmc_opts_default(tmc_opts* o);
const char* tmc_last_error(void);
const char* tmc_version(void);
int         tmc_device_count(void);

/* == b1ToB2 (spectral-clustering), drop-in for Preprocess.hs:204-205.
 * val_out[k] = val[k] * ln(n_rows / df[col[k]]), df_j = #{entries > 0 in column j}; idf_out optional. */
int tmc_tfidf(const tmc_csr* b1, double* val_out, double* idf_out);

/* == b2ToB: val_out = row-L2-normalised values; norm_out (n_rows, optional) = the row norms. */
int tmc_row_l2(const tmc_csr* b2, double* val_out, double* norm_out);

/* Upload a B2 matrix (host CSR) and keep B = b2ToB(B2) on the device (normalize=1 applies idf first). */
int  tmc_matrix_upload(const tmc_csr* b2, int normalize, int device, tmc_matrix** out);
/* Adopt DEVICE arrays (row_ptr int64, col int32, val double; not copied, caller keeps them alive).
 * val_dev is normalised in place (idf if normalize, then row L2) only when the values are neither small integers nor
 * an exactly recognisable B2 = counts * diag(column weight) (a tf-idf matrix, normalize = 0); in those two cases the array
 * is left untouched and a compact copy of the counts is kept instead.  row_offset / n_rows_global describe this
 * rank's row block under a communicator (0, 0 = the whole matrix). */
int  tmc_matrix_adopt_dev(int64_t n_rows, int64_t n_cols, int64_t nnz, const int64_t* row_ptr_dev,
                          const int32_t* col_dev, double* val_dev, int64_t row_offset, int64_t n_rows_global,
                          int normalize, int device, tmc_matrix** out);
void tmc_matrix_free(tmc_matrix*);
/* How the device copy is stored (diagnostics, tests, bench accounting): info[0] value storage (0 fp64 B, 1 uint16 counts +
 * diagonal factors, 2 all ones), [1] columns in the dense panel, [2] panel bytes per row, [3] entries left in the residual
 * CSR, [4] entries of the caller's CSR, [5] local rows, [6] bits per panel cell (8, 4, or 0 without a panel).
 * Writes min(n, 7) values. */
int  tmc_matrix_info(const tmc_matrix* b, int64_t* info, int n);
/* tmc_matrix_free keeps one retired workspace (device arenas, pinned buffers) for the next matrix; this releases it. */
void tmc_cache_clear(void);

/* == hierarchicalSpectralCluster ... (Left mat)  (Cluster.hs:147-156): the whole tree. */
int  tmc_make_tree(const tmc_csr* b2, const tmc_opts* opts, tmc_tree** out);
int  tmc_make_tree_dev(tmc_matrix* b, const tmc_opts* opts, tmc_tree** out);
/* == HSD.hierarchicalSpectralCluster ... (Left (sparseToHMat mat))  (--dense, Cluster.hs:157-167): `mat` is the dense
 * cells x features matrix, row-major.  Zeros carry no weight in any product of the path, so the tree is the one the sparse
 * call builds from the non-zero entries (signed entries, the usual case after PCA / batch correction, included). */
int  tmc_make_tree_dense(const double* mat, int64_t n_rows, int64_t n_cols, const tmc_opts* opts, tmc_tree** out);
void tmc_tree_free(tmc_tree*);
/* == the `runs` argument of hierarchicalSpectralCluster (--numRuns, Options.hs:174) on a finished tree: for every split the
 * labels are shuffled num_runs times and Q recomputed on the GPU; p_out[node] = #{Q_shuffled >= Q_observed} / num_runs,
 * NaN for leaves.  tmc_make_tree* with opts.num_runs > 0 call this and store the result in tmc_tree.p_val. */
int  tmc_permutation_test(tmc_matrix* b, const tmc_tree* tree, int num_runs, uint64_t seed, double* p_out);

/* Fine-grained exports so every row of SURVEY.md 8(a) is individually parity-testable.
 * `rows` (n_sel original row ids, ascending) selects the node S; NULL = all rows. */
/* == bToD: d = |B_S| (|B_S|^T 1) */
int tmc_degree(tmc_matrix* b, const int64_t* rows, int64_t n_sel, double* d_out);
/* == secondLeft 2 1 (bdToC B D) via sparseSvd: u2 (n_sel), theta = sigma_2^2, Lanczos steps used. */
int tmc_second_left(tmc_matrix* b, const int64_t* rows, int64_t n_sel, const tmc_opts* opts,
                    double* u2_out, double* theta_out, int32_t* iters_out);
/* == getBModularity labels B_S */
int tmc_modularity(tmc_matrix* b, const int64_t* rows, int64_t n_sel, const uint8_t* labels, double* q_out);
/* == stable row split (extractRow/fromRowsL over the sorted index lists): rows_out = [label0 | label1] */
int tmc_partition(const int64_t* rows, int64_t n_sel, const uint8_t* labels, int64_t* rows_out, int64_t* n_left);
/* one fused pair x' = C_S (C_S^T x) (the las2 `opb`), for the roofline harness and parity tests */
int tmc_spmv_pair(tmc_matrix* b, const int64_t* rows, int64_t n_sel, const double* x, double* x_out);

/* Kernel-only timing of the SpMV pair on the full matrix: `reps` launches, average ms per
 * scatter (y = C^T x) and gather (x' = C y) kernel, measured with CUDA events on the engine stream. */
int tmc_bench_spmv_pair(tmc_matrix* b, int reps, double* scatter_ms, double* gather_ms);

/* ---- the steps right before the hot path, on the GPU.  Results are library-owned HOST arrays. */
typedef struct {
  int64_t n_rows, n_cols, nnz;
  const int64_t* row_ptr; const int32_t* col_idx; const double* val;
  const int64_t* kept_rows; const int64_t* kept_cols;      /* tmc_filter_thresholds: original ids of the kept rows / columns */
  int64_t n_kept_rows, n_kept_cols;
} tmc_csr_out;
/* == readMatrix + matToSpMat + transposeSM + sparsifySM (src/TooManyCells/Matrix/Load.hs:85-152): the text of a 10x
 * `matrix.mtx` (already gunzipped; features x cells, coordinate real|integer|pattern) -> CSR.  transpose=1 gives
 * cells x features (what make-tree wants); explicit zeros are dropped; entries sorted by (row, column). */
int  tmc_mtx_parse(const char* text, int64_t n_bytes, int transpose, tmc_csr_out** out);
/* == filterNumSparseMat (src/TooManyCells/Matrix/Preprocess.hs:342-386), --filter-thresholds "(MINCELL, MINFEATURE)":
 * keep rows with sum >= row_thresh and columns with sum >= col_thresh, both sums over the ORIGINAL matrix. */
int  tmc_filter_thresholds(const tmc_csr* a, double row_thresh, double col_thresh, tmc_csr_out** out);
void tmc_csr_out_free(tmc_csr_out*);

/* ---- == secondLeft 1 N (b1ToB2 mat) (lsaSparseMat, src/TooManyCells/Matrix/Preprocess.hs:560-571): the first
 * n_vec left singular vectors of the (optionally idf-scaled) matrix, by Lanczos over the same SpMV pair as the tree.
 * u_out: n_rows x n_vec row-major (one n_vec-vector per cell, as S.fromRowsL builds it; each column unit-norm, sign arbitrary);
 * sigma_out: n_vec singular values, descending; converged_out: 1 if every pair met opts->tol. */
int tmc_left_singular(const tmc_csr* a, int normalize, int n_vec, const tmc_opts* opts, double* u_out, double* sigma_out,
                      int* converged_out);

/* Multi-GPU (one process per GPU): NCCL communicator owned by the library.
 * Rank 0 calls tmc_comm_unique_id and ships the 128 bytes to the other ranks (torch.distributed),
 * then every rank calls tmc_comm_init. */
int  tmc_comm_unique_id(uint8_t id_out[128]);
int  tmc_comm_init(const uint8_t id[128], int rank, int world, int device);
void tmc_comm_destroy(void);
/* Collective self-test of the one-shot all-reduce over NVLink peer memory (csrc/tmc_p2p.cuh, enabled for the tree with
 * TMC_P2P=1) against ncclAllReduce on n_elems doubles: largest absolute difference and average device microseconds per call of
 * either path (p2p_us = -1 if the peer windows could not be set up).  tools/p2p_check.py drives it under torchrun. */
int  tmc_comm_selftest(int64_t n_elems, int reps, double* max_abs_diff, double* p2p_us, double* nccl_us);

#ifdef __cplusplus
}
#endif
#endif
