{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE NoImplicitPrelude #-}
-- | Binding of libfwgpu.so (include/fwgpu.h) for jinilover/floydWarshall.
--
-- NEW module for the reference's library stanza (src/lib/FwGpu.hs).  It provides
-- 'floydWarshallGpu', a drop-in for the body of @floydWarshall@
-- (reference src/lib/Algorithms.hs:19-20): same argument, same result type, results
-- bit-identical to @runAlgo@ (Algorithms.hs:42-61) including the exact @_path@ lists.
-- @buildMatrix@ and @runAlgo@ both run on the device ('fw_solve_edges'); only the
-- rate map (as index triples) goes up and the dense matrices come down.
--
-- STATUS: GHC is not installed in the build image of this repository, so this file has
-- NOT been compiled.  The C ABI it binds is exercised end to end by the C++ mirror
-- (floydwarshall_b200/host/algorithms.hpp + tests/cpp) and the ctypes layer
-- (floydwarshall_b200/algorithms.py + tests/).  See INTEGRATION.md for the cabal changes.
module FwGpu
  ( floydWarshallGpu
  , floydWarshallMultiGpu
  , fwCtxLastError
  ) where

import Protolude
import Foreign
import Foreign.C.String (peekCString)
import Foreign.C.Types
import System.IO.Unsafe (unsafePerformIO)

import qualified Data.List as L
import qualified Data.Map as M
import qualified Data.Vector as V
import qualified Data.Vector.Storable as S
import qualified Data.Vector.Storable.Mutable as SM

import Types

-- int fw_solve_edges(fw_ctx*, int32_t n, const int32_t* ccy, int32_t m, const int32_t* src,
--                    const int32_t* dst, const double* val, double* rate, int32_t* next,
--                    int32_t* init_next, int32_t* mid, int32_t* csT, int32_t* rs);
-- `safe`: the call runs for seconds on large graphs and the executable is -threaded
-- (floydWarshall.cabal:49).  A NULL context selects the library's default context on device 0.
foreign import ccall safe "fw_solve_edges"
  c_fw_solve_edges :: Ptr () -> Int32 -> Ptr Int32 -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr CDouble
                   -> Ptr CDouble -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32
                   -> IO CInt

-- const char* fw_ctx_last_error(fw_ctx*);  NULL = the default context.
-- The text is kept PER CONTEXT, not per OS thread: a `safe` call and the query after it may run on
-- different OS threads for an unbound Haskell thread, so the thread-local fw_last_error() is not used here.
foreign import ccall unsafe "fw_ctx_last_error"
  c_fw_ctx_last_error :: Ptr () -> IO (Ptr CChar)

fwCtxLastError :: Ptr () -> IO Text
fwCtxLastError ctx = toS <$> (c_fw_ctx_last_error ctx >>= peekCString)

-- The multi-GPU object (include/fwgpu.h, "multi-GPU solve"): one process drives every GPU of the box.
--   int fw_multi_create(int32_t ndev, const int32_t* devices, fw_multi** out);
--   int fw_multi_solve_edges(fw_multi*, int32_t n, const int32_t* ccy, int32_t m, const int32_t* src,
--                            const int32_t* dst, const double* val, double* rate, int32_t* next,
--                            int32_t* init_next, int32_t* mid, int32_t* csT, int32_t* rs);
--   const char* fw_multi_last_error(fw_multi*);
-- For the REPL flow bind fw_multi_sync (on OutSync) + fw_multi_optimum (per request) instead: the optimised
-- matrix then stays sharded in HBM and only the requested entry and its path cross PCIe.
foreign import ccall safe "fw_multi_create"
  c_fw_multi_create :: Int32 -> Ptr Int32 -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "fw_multi_solve_edges"
  c_fw_multi_solve_edges :: Ptr () -> Int32 -> Ptr Int32 -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr CDouble
                         -> Ptr CDouble -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32
                         -> IO CInt
foreign import ccall unsafe "fw_multi_last_error"
  c_fw_multi_last_error :: Ptr () -> IO (Ptr CChar)
foreign import ccall unsafe "fw_device_count"
  c_fw_device_count :: IO CInt

-- | One fw_multi object over all GPUs of the box, created on first use and kept for the process.
multiHandle :: Ptr ()
multiHandle = unsafePerformIO $ do
  ndev <- c_fw_device_count
  alloca $ \out -> do
    rc <- c_fw_multi_create (fromIntegral ndev) nullPtr out
    when (rc /= 0) $ throwIO (userError ("fw_multi_create failed (" <> show rc <> ")"))
    peek out
{-# NOINLINE multiHandle #-}

-- | The five dense n x n tables of a solve (row-major).
data Solved = Solved
  { sRate     :: !(S.Vector CDouble)  -- ^ _bestRate
  , sInitNext :: !(S.Vector Int32)    -- ^ next-hops of buildMatrix (edge (a,b) exists iff >= 0)
  , sMid      :: !(S.Vector Int32)    -- ^ k of the last replacement of (i,j), -1 if never replaced
  , sCsT      :: !(S.Vector Int32)    -- ^ mid of (i,k) as of step k   (stored at [i*n+k])
  , sRs       :: !(S.Vector Int32)    -- ^ mid of (k,j) as of step k   (stored at [k*n+j])
  }

-- | Same contract as the reference's @floydWarshall@ (Algorithms.hs:19-20).
floydWarshallGpu :: M.Map (Vertex, Vertex) Double -> Matrix RateEntry
floydWarshallGpu exRates
  | V.null vertices = V.empty                                -- AlgorithmsTest.hs:62-64
  | otherwise = unsafePerformIO $ do                         -- pure caller: ProcessRequests.hs:82-84
      solved <- solveEdges (c_fw_solve_edges nullPtr) (fwCtxLastError nullPtr) n ccyIds edges
      pure $ V.generate n $ \i -> V.generate n $ \j ->
        RateEntry { _bestRate = realToFrac (sRate solved S.! (i * n + j))
                  , _start    = vertices V.! i
                  , _path     = map (vertices V.!) (pathOf n solved i j) }   -- lazy per entry
  where
    -- vertex order of the reference: sort . nub of every key component (Algorithms.hs:29)
    vertices = V.fromList . L.sort . L.nub $ M.keys exRates >>= \(a, b) -> [a, b]
    n        = V.length vertices
    index    = M.fromList (zip (V.toList vertices) [0 ..]) :: M.Map Vertex Int
    ix v     = fromIntegral (index M.! v) :: Int32
    -- any integer id per currency: the library only tests ids for equality (same-currency rule,
    -- Algorithms.hs:35)
    ccyTable = M.fromList (zip (L.nub (map _ccy (V.toList vertices))) [0 ..]) :: M.Map Text Int32
    ccyIds   = S.fromList [ccyTable M.! _ccy v | v <- V.toList vertices]
    edges    = [(ix a, ix b, r) | ((a, b), r) <- M.toList exRates]
{-# NOINLINE floydWarshallGpu #-}

-- | The same function on every GPU of the box (rows sharded over the GPUs, pivot-row panels over NVLink).
floydWarshallMultiGpu :: M.Map (Vertex, Vertex) Double -> Matrix RateEntry
floydWarshallMultiGpu exRates
  | V.null vertices = V.empty
  | otherwise = unsafePerformIO $ do
      solved <- solveEdges (c_fw_multi_solve_edges multiHandle)
                           (toS <$> (c_fw_multi_last_error multiHandle >>= peekCString)) n ccyIds edges
      pure $ V.generate n $ \i -> V.generate n $ \j ->
        RateEntry { _bestRate = realToFrac (sRate solved S.! (i * n + j))
                  , _start    = vertices V.! i
                  , _path     = map (vertices V.!) (pathOf n solved i j) }
  where
    vertices = V.fromList . L.sort . L.nub $ M.keys exRates >>= \(a, b) -> [a, b]
    n        = V.length vertices
    index    = M.fromList (zip (V.toList vertices) [0 ..]) :: M.Map Vertex Int
    ix v     = fromIntegral (index M.! v) :: Int32
    ccyTable = M.fromList (zip (L.nub (map _ccy (V.toList vertices))) [0 ..]) :: M.Map Text Int32
    ccyIds   = S.fromList [ccyTable M.! _ccy v | v <- V.toList vertices]
    edges    = [(ix a, ix b, r) | ((a, b), r) <- M.toList exRates]
{-# NOINLINE floydWarshallMultiGpu #-}

type SolveEdgesCall = Int32 -> Ptr Int32 -> Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr CDouble
                    -> Ptr CDouble -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> Ptr Int32 -> IO CInt

solveEdges :: SolveEdgesCall -> IO Text -> Int -> S.Vector Int32 -> [(Int32, Int32, Double)] -> IO Solved
solveEdges call lastError n ccyIds edges = do
  let m    = length edges
      srcV = S.fromList [a | (a, _, _) <- edges]
      dstV = S.fromList [b | (_, b, _) <- edges]
      valV = S.fromList [realToFrac r | (_, _, r) <- edges] :: S.Vector CDouble
  rate <- SM.new (n * n)
  next <- SM.new (n * n)
  ini  <- SM.new (n * n)
  mid  <- SM.new (n * n)
  csT  <- SM.new (n * n)
  rs   <- SM.new (n * n)
  rc <- S.unsafeWith ccyIds $ \pc -> S.unsafeWith srcV $ \ps -> S.unsafeWith dstV $ \pd ->
        S.unsafeWith valV $ \pv -> SM.unsafeWith rate $ \pr -> SM.unsafeWith next $ \pn ->
        SM.unsafeWith ini $ \pi' -> SM.unsafeWith mid $ \pm -> SM.unsafeWith csT $ \pcs ->
        SM.unsafeWith rs $ \prs ->
          call (fromIntegral n) pc (fromIntegral m) ps pd pv pr pn pi' pm pcs prs
  when (rc /= 0) $ do
    msg <- lastError
    -- no CPU fallback: the REPL keeps its previous state on an exception (src/app/Main.hs:30-33)
    throwIO (userError ("fwgpu (" <> show rc <> "): " <> toS msg))
  Solved <$> S.unsafeFreeze rate <*> S.unsafeFreeze ini <*> S.unsafeFreeze mid
         <*> S.unsafeFreeze csT <*> S.unsafeFreeze rs

-- | The reference's @_path@ of entry (i,j): @ikPath ++ kjPath@ with both halves AS OF the step that
-- made the replacement (Algorithms.hs:55).  mid/csT/rs record exactly which step that was for the
-- entry itself, for column snapshots and for row snapshots; an entry that was never replaced
-- contributes its buildMatrix path (@[j]@ if the edge exists).  Same recursion as
-- floydwarshall_b200/csrc/fw_paths.cuh and oracle/fw_oracle.py:reconstruct_path.
pathOf :: Int -> Solved -> Int -> Int -> [Int]
pathOf n s i j = go (sMid s) i j []
  where
    at tbl a b = fromIntegral (tbl S.! (a * n + b)) :: Int
    -- `go tbl a b rest` prepends the path of (a,b), read through table `tbl`, to `rest`
    go tbl a b rest =
      case at tbl a b of
        k | k < 0     -> if at (sInitNext s) a b >= 0 then b : rest else rest
          | otherwise -> go (sCsT s) a k (go (sRs s) k b rest)
