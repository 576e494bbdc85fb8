{-# LANGUAGE ForeignFunctionInterface #-}
{- |
Module      : TooManyCells.MakeTree.TmcFFI
Description : FFI binding of libtmc (include/tmc.h), the B200 engine behind @make-tree@.

The file a too-many-cells maintainer drops in as @src/TooManyCells/MakeTree/TmcFFI.hs@ (INTEGRATION.md).  It replaces the
two calls of the hot path and nothing else:

  * @hierarchicalSpectralCluster eigenGroup False numEigen Nothing minMod runs items (Left mat)@
    (src/TooManyCells/MakeTree/Cluster.hs:147-156) and its @--dense@ twin (Cluster.hs:157-167)
  * @tfidfScaleSparseMat@ (src/TooManyCells/Matrix/Preprocess.hs:204-205)

No GHC in the build image of this repository: the module is written against the documented APIs of base, vector,
sparse-linear-algebra and hmatrix and has not been compiled here; what IS machine-checked is the struct layout it relies
on (the @off_*@ / @size_*@ constants below are compared with the C layout by
tests/test_host.py::test_haskell_binding_offsets_match_the_header).
-}
module TooManyCells.MakeTree.TmcFFI
  ( tmcHierarchicalSpectralCluster
  , tmcHierarchicalSpectralClusterDense
  , tmcTfIdfScaleSparseMat
  ) where

import Control.Exception (bracket)
import Control.Monad (when)
import Data.Int (Int32, Int64)
import Data.List (sortOn)
import Data.Maybe (fromMaybe)
import Data.Tree (Tree (..))
import Foreign
import Foreign.C.String (peekCString)
import Foreign.C.Types
import qualified Data.Sparse.Common as S
import qualified Data.Vector as V
import qualified Data.Vector.Storable as VS
import qualified Numeric.LinearAlgebra as H

import Math.Clustering.Hierarchical.Spectral.Types
       (ClusteringTree, ClusteringVertex (..), EigenGroup (..))
import Math.Modularity.Types (Q (..))

-- ---------------------------------------------------------------------------------------------------------------------
-- struct layouts of include/tmc.h (x86-64 / aarch64 LP64)
size_csr, size_opts, size_tree :: Int
size_csr = 72
size_opts = 96
size_tree = 152

off_csr_n_rows, off_csr_n_cols, off_csr_nnz, off_csr_row_ptr, off_csr_col_idx, off_csr_val, off_csr_row_offset, off_csr_n_rows_global :: Int
off_csr_n_rows = 0
off_csr_n_cols = 8
off_csr_nnz = 16
off_csr_row_ptr = 24
off_csr_col_idx = 32
off_csr_val = 40
off_csr_row_offset = 48
off_csr_n_rows_global = 56
off_csr_val_type, off_csr_idx_type :: Int
off_csr_val_type = 64
off_csr_idx_type = 68

off_opts_eigen_group, off_opts_num_eigen, off_opts_min_modularity, off_opts_min_size, off_opts_normalize, off_opts_num_runs, off_opts_seed :: Int
off_opts_eigen_group = 0
off_opts_num_eigen = 4
off_opts_min_modularity = 8
off_opts_min_size = 16
off_opts_normalize = 20
off_opts_num_runs = 24
off_opts_seed = 32

off_tree_n_nodes, off_tree_n_rows, off_tree_left, off_tree_right, off_tree_q, off_tree_item_begin, off_tree_item_end, off_tree_perm, off_tree_p_val :: Int
off_tree_n_nodes = 0
off_tree_n_rows = 8
off_tree_left = 16
off_tree_right = 24
off_tree_q = 40
off_tree_item_begin = 64
off_tree_item_end = 72
off_tree_perm = 80
off_tree_p_val = 144

data TmcCsr
data TmcOpts
data TmcTree

foreign import ccall safe "tmc_opts_default"    c_opts_default    :: Ptr TmcOpts -> IO ()
foreign import ccall safe "tmc_make_tree"       c_make_tree       :: Ptr TmcCsr -> Ptr TmcOpts -> Ptr (Ptr TmcTree) -> IO CInt
foreign import ccall safe "tmc_make_tree_dense" c_make_tree_dense :: Ptr CDouble -> Int64 -> Int64 -> Ptr TmcOpts -> Ptr (Ptr TmcTree) -> IO CInt
foreign import ccall safe "tmc_tree_free"       c_tree_free       :: Ptr TmcTree -> IO ()
foreign import ccall safe "tmc_tfidf"           c_tfidf           :: Ptr TmcCsr -> Ptr CDouble -> Ptr CDouble -> IO CInt
foreign import ccall safe "tmc_last_error"      c_last_error      :: IO (Ptr CChar)

-- | The reference reports failures with 'error'; so does the binding (message from the library, thread-local).
orDie :: CInt -> IO ()
orDie rc = when (rc /= 0) $ do
  msg <- c_last_error >>= peekCString
  error ("libtmc (" ++ show rc ++ "): " ++ msg)

-- ---------------------------------------------------------------------------------------------------------------------
-- marshalling

-- | @S.toListSM@ -> CSR.  Entries are grouped by row (the library does not need sorted columns); explicit zeros are dropped
-- as @sparsifySM@ would (src/TooManyCells/Program/LoadMatrix.hs:229).
toCsr :: S.SpMatrix Double -> (VS.Vector Int64, VS.Vector Int32, VS.Vector CDouble)
toCsr mat = (rowPtr, cols, vals)
  where
    es     = sortOn (\(i, j, _) -> (i, j)) . filter (\(_, _, x) -> x /= 0) . S.toListSM $ mat
    counts = VS.accum (+) (VS.replicate (S.nrows mat) 0) [ (i, 1) | (i, _, _) <- es ]
    rowPtr = VS.scanl' (+) 0 counts                                  -- length nrows + 1
    cols   = VS.fromList [ fromIntegral j | (_, j, _) <- es ]
    vals   = VS.fromList [ realToFrac x   | (_, _, x) <- es ]

withCsr :: S.SpMatrix Double -> (Ptr TmcCsr -> Int -> IO b) -> IO b
withCsr mat act =
  VS.unsafeWith rowPtr $ \pr -> VS.unsafeWith cols $ \pc -> VS.unsafeWith vals $ \pv ->
    allocaBytes size_csr $ \csr -> do
      pokeByteOff csr off_csr_n_rows        (fromIntegral (S.nrows mat)   :: Int64)
      pokeByteOff csr off_csr_n_cols        (fromIntegral (S.ncols mat)   :: Int64)
      pokeByteOff csr off_csr_nnz           (fromIntegral (VS.length vals) :: Int64)
      pokeByteOff csr off_csr_row_ptr       pr
      pokeByteOff csr off_csr_col_idx       pc
      pokeByteOff csr off_csr_val           pv
      pokeByteOff csr off_csr_row_offset    (0 :: Int64)              -- whole matrix, single process
      pokeByteOff csr off_csr_n_rows_global (0 :: Int64)
      pokeByteOff csr off_csr_val_type      (0 :: Int32)              -- CDouble values / CInt indices (see INTEGRATION.md for Word16 counts)
      pokeByteOff csr off_csr_idx_type      (0 :: Int32)
      act csr (VS.length vals)
  where
    (rowPtr, cols, vals) = toCsr mat

-- | The same for a matrix of raw UMI counts (the make-tree input before tf-idf): counts travel as Word16 and gene indices as
-- Word16 when they fit (tmc_csr.val_type = TMC_VAL_U16 = 2, idx_type = TMC_IDX_U16 = 1), 4 instead of 12 bytes per non-zero
-- over PCIe; libtmc widens them on the device, so the tree is bit-identical to the CDouble call.  Returns Nothing when an entry
-- is not an integer in [0, 65535] or the matrix has more than 65536 features (the caller then falls back to 'withCsr').
withCsrCounts :: S.SpMatrix Double -> (Ptr TmcCsr -> Int -> IO b) -> Maybe (IO b)
withCsrCounts mat act
  | S.ncols mat > 65536 || any bad es = Nothing
  | otherwise = Just $
      VS.unsafeWith rowPtr $ \pr -> VS.unsafeWith cols16 $ \pc -> VS.unsafeWith vals16 $ \pv ->
        allocaBytes size_csr $ \csr -> do
          pokeByteOff csr off_csr_n_rows        (fromIntegral (S.nrows mat)     :: Int64)
          pokeByteOff csr off_csr_n_cols        (fromIntegral (S.ncols mat)     :: Int64)
          pokeByteOff csr off_csr_nnz           (fromIntegral (VS.length vals16) :: Int64)
          pokeByteOff csr off_csr_row_ptr       pr
          pokeByteOff csr off_csr_col_idx       (castPtr pc :: Ptr Int32)   -- element type is given by idx_type
          pokeByteOff csr off_csr_val           (castPtr pv :: Ptr CDouble) -- element type is given by val_type
          pokeByteOff csr off_csr_row_offset    (0 :: Int64)
          pokeByteOff csr off_csr_n_rows_global (0 :: Int64)
          pokeByteOff csr off_csr_val_type      (2 :: Int32)                -- TMC_VAL_U16
          pokeByteOff csr off_csr_idx_type      (1 :: Int32)                -- TMC_IDX_U16
          act csr (VS.length vals16)
  where
    es      = sortOn (\(i, j, _) -> (i, j)) . filter (\(_, _, x) -> x /= 0) . S.toListSM $ mat
    bad (_, _, x) = x < 0 || x > 65535 || x /= fromIntegral (round x :: Int)
    counts  = VS.accum (+) (VS.replicate (S.nrows mat) 0) [ (i, 1) | (i, _, _) <- es ]
    rowPtr  = VS.scanl' (+) 0 counts :: VS.Vector Int64
    cols16  = VS.fromList [ fromIntegral j | (_, j, _) <- es ] :: VS.Vector Word16
    vals16  = VS.fromList [ round x        | (_, _, x) <- es ] :: VS.Vector Word16

-- | Flag contract of the engine call (src/TooManyCells/Program/Options.hs:172-178; Cluster.hs:147-156).
withOpts :: EigenGroup -> Maybe Int -> Maybe Q -> Maybe Int -> (Ptr TmcOpts -> IO b) -> IO b
withOpts eigenGroup numEigen minMod runs act =
  allocaBytes size_opts $ \o -> do
    c_opts_default o
    pokeByteOff o off_opts_eigen_group    (case eigenGroup of { SignGroup -> 0; KMeansGroup -> 1 } :: Int32)
    pokeByteOff o off_opts_num_eigen      (fromIntegral (fromMaybe 1 numEigen) :: Int32)
    pokeByteOff o off_opts_min_modularity (maybe 0 unQ minMod :: Double)        -- split iff Q > minMod
    pokeByteOff o off_opts_min_size       (1 :: Int32)                          -- Cluster.hs:152 passes Nothing
    pokeByteOff o off_opts_normalize      (0 :: Int32)                          -- Cluster.hs:150: normFlag False
    pokeByteOff o off_opts_num_runs       (fromIntegral (fromMaybe 0 runs) :: Int32)
    act o

peekVec :: Storable a => Ptr TmcTree -> Int -> Int -> IO (VS.Vector a)
peekVec t off n = do
  p <- peekByteOff t off
  VS.generateM n (peekElemOff p)

-- | Rebuild @Tree (ClusteringVertex a)@ from the flat pre-order arrays (children order [label 0, label 1], the order
-- birch-beer's @treeToGraph@ numbers them in, Cluster.hs:171-181) and release the library-owned tree.
takeTree :: V.Vector a -> Ptr (Ptr TmcTree) -> IO (ClusteringTree a)
takeTree items out = bracket (peek out) c_tree_free $ \t -> do
  n      <- fromIntegral <$> (peekByteOff t off_tree_n_nodes :: IO Int64)
  m      <- fromIntegral <$> (peekByteOff t off_tree_n_rows  :: IO Int64)
  left   <- peekVec t off_tree_left       n :: IO (VS.Vector Int64)
  right  <- peekVec t off_tree_right      n :: IO (VS.Vector Int64)
  qs     <- peekVec t off_tree_q          n :: IO (VS.Vector Double)
  ps     <- peekVec t off_tree_p_val      n :: IO (VS.Vector Double)
  begins <- peekVec t off_tree_item_begin n :: IO (VS.Vector Int64)
  ends   <- peekVec t off_tree_item_end   n :: IO (VS.Vector Int64)
  perm   <- peekVec t off_tree_perm       m :: IO (VS.Vector Int64)
  let vertex i = ClusteringVertex
        { _clusteringItems = V.generate (fromIntegral (ends VS.! i - begins VS.! i))
                                        (\k -> items V.! fromIntegral (perm VS.! (fromIntegral (begins VS.! i) + k)))
        , _ngMod           = Q (qs VS.! i)
        , _pVal            = let p = ps VS.! i in if isNaN p then Nothing else Just p
        }
      build i
        | left VS.! i < 0 = Node (vertex i) []
        | otherwise       = Node (vertex i) [build (fromIntegral (left VS.! i)), build (fromIntegral (right VS.! i))]
      tree = build 0
  -- force the spine and the item vectors before the bracket frees the C arrays (the vectors above are copies already;
  -- this only avoids holding them through thunks)
  length (foldr (:) [] (fmap (V.length . _clusteringItems) tree)) `seq` pure tree

-- ---------------------------------------------------------------------------------------------------------------------
-- the drop-in calls

-- | Drop-in for @hierarchicalSpectralCluster eigenGroup False numEigen Nothing minMod runs items (Left mat)@
-- (Cluster.hs:147-156): @hSpecCommand False = tmcHierarchicalSpectralCluster eigenGroup (fmap unNumEigen numEigen) minModMay
-- (fmap unNumRuns runsMay) items@.
tmcHierarchicalSpectralCluster
  :: EigenGroup -> Maybe Int -> Maybe Q -> Maybe Int -> V.Vector a -> S.SpMatrix Double -> IO (ClusteringTree a)
tmcHierarchicalSpectralCluster eigenGroup numEigen minMod runs items mat =
  withCsr mat $ \csr _ -> withOpts eigenGroup numEigen minMod runs $ \opts -> alloca $ \out -> do
    c_make_tree csr opts out >>= orDie
    takeTree items out

-- | Drop-in for the @--dense@ arm, @HSD.hierarchicalSpectralCluster ... (Left (sparseToHMat mat))@ (Cluster.hs:157-167).
-- hmatrix matrices are row-major, which is what @tmc_make_tree_dense@ reads.
tmcHierarchicalSpectralClusterDense
  :: EigenGroup -> Maybe Int -> Maybe Q -> Maybe Int -> V.Vector a -> H.Matrix Double -> IO (ClusteringTree a)
tmcHierarchicalSpectralClusterDense eigenGroup numEigen minMod runs items mat =
  VS.unsafeWith (H.flatten mat) $ \p -> withOpts eigenGroup numEigen minMod runs $ \opts -> alloca $ \out -> do
    c_make_tree_dense (castPtr p) (fromIntegral (H.rows mat)) (fromIntegral (H.cols mat)) opts out >>= orDie
    takeTree items out

-- | Drop-in for @unB2 . b1ToB2 . B1@ inside @tfidfScaleSparseMat@ (Preprocess.hs:204-205): same sparsity pattern,
-- values scaled by @ln (nrows / df_j)@.
tmcTfIdfScaleSparseMat :: S.SpMatrix Double -> IO (S.SpMatrix Double)
tmcTfIdfScaleSparseMat mat =
  withCsr mat $ \csr nnz -> allocaArray nnz $ \pOut -> do
    c_tfidf csr pOut nullPtr >>= orDie
    scaled <- peekArray nnz pOut
    let es = sortOn (\(i, j, _) -> (i, j)) . filter (\(_, _, x) -> x /= 0) . S.toListSM $ mat     -- the order toCsr used
    pure $ S.fromListSM (S.nrows mat, S.ncols mat) [ (i, j, realToFrac v) | ((i, j, _), v) <- zip es (scaled :: [CDouble]) ]
