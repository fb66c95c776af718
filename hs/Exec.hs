{-# LANGUAGE NamedFieldPuns #-}
-- | Exec backend for mplan2vdl: runs the Vlite graph on a B200 through libvdl_cuda.
-- SOURCE ONLY (no GHC in the build image).  Hook: MainFuns.compile, right after the cleanup passes produce
-- `vexps'` (MainFuns.hs:184-186), besides `Vdl.vdlFromVexps` (187):
--
-- >   case exec_mode of
-- >     ExecText  -> show <$> runReader (Vdl.vdlFromVexps vexps') config      -- today's behaviour
-- >     ExecCuda  -> Exec.runOnGpu config vexps'                              -- this module
--
-- Two ways to drive the library, both shown:
--   * `runProgramText`: hand the printed program to vdl_plan_load -- the library hash-conses it, runs its own
--     select->map->fold fusion and executes; nothing else to do on the Haskell side.
--   * `evalVexp`: walk the `Vexp` DAG (shared through memoized_hash, Vlite.hs:143-157) op by op, one FFI call per
--     `Vx` constructor, using the `ColInfo` of each node for the explicit lengths the C ABI wants
--     (Scatter: App. G2).
module Exec (runProgramText, evalVexp) where

import qualified Data.ByteString.Char8 as C
import qualified Data.HashMap.Strict as Map
import Control.Monad.State
import Foreign
import Foreign.C.String
import Foreign.C.Types

import Config (Config, ColInfo (..))
import qualified Vlite as V
import VdlCuda

check :: Ptr VdlCtx -> CInt -> IO ()
check ctx rc
  | rc == vdlOk = return ()
  | otherwise = do msg <- c_vdl_last_error ctx >>= peekCString
                   ioError (userError ("libvdl_cuda error " ++ show rc ++ ": " ++ msg))

-- | Run the program text (exactly what `show (Vdl ...)` prints) and return (output name, values) pairs in
-- MaterializeCompact order -- the shape resolve.py decodes.
runProgramText :: Ptr VdlCtx -> String -> IO [(String, [Int64])]
runProgramText ctx text =
  withCString text $ \ctext -> alloca $ \pplan -> do
    c_vdl_plan_load ctx ctext 1 pplan >>= check ctx          -- 1 = VDL_PLAN_FUSE
    plan <- peek pplan
    c_vdl_plan_run plan >>= check ctx
    n <- c_vdl_plan_num_outputs plan
    outs <- forM [0 .. n - 1] $ \i -> alloca $ \pname -> alloca $ \pdata -> alloca $ \plen -> do
      c_vdl_plan_output plan i pname pdata plen >>= check ctx
      name <- peek pname >>= peekCString
      len <- peek plen
      vals <- peek pdata >>= peekArray (fromIntegral len)
      return (name, vals)
    _ <- c_vdl_plan_destroy plan
    return outs

type Memo = Map.HashMap V.Vexp VdlVec

binopCode :: V.BinaryOp -> Maybe CInt     -- order of Voodop (Vdl.hs:110-123); Lt/Leq/Geq/Neq/Min/Max are lowered first
binopCode op = lookup op [ (V.LogAnd, 0), (V.LogOr, 1), (V.BitAnd, 2), (V.BitOr, 3), (V.BitShift, 4), (V.Eq, 5)
                         , (V.Add, 6), (V.Sub, 7), (V.Gt, 8), (V.Mul, 9), (V.Div, 10), (V.Mod, 11) ]

-- | Evaluate one node, memoised on the Vexp's structural hash.
evalVexp :: Ptr VdlCtx -> V.Vexp -> StateT Memo IO VdlVec
evalVexp ctx vexp@V.Vexp { V.vx, V.info = ColInfo { count } } = do
  memo <- get
  case Map.lookup vexp memo of
    Just h -> return h
    Nothing -> do
      h <- go vx
      modify (Map.insert vexp h)
      return h
  where
    out f = liftIO $ alloca $ \p -> f p >>= check ctx >> peek p
    lenOf h = liftIO $ alloca $ \p -> c_vdl_vec_len ctx h p >>= check ctx >> peek p
    go (V.Load n) = out $ \p -> withCString (show n) $ \s -> c_vdl_column_lookup ctx s p
    go V.RangeV { V.rmin, V.rstep, V.rref } = do
      r <- evalVexp ctx rref
      l <- lenOf r
      out (c_vdl_op_range ctx (fromInteger rmin) (fromInteger rstep) l)
    go V.RangeC { V.rmin, V.rstep, V.rcount } = out (c_vdl_op_range ctx (fromInteger rmin) (fromInteger rstep) (fromInteger rcount))
    go V.Binop { V.binop, V.left, V.right } = case binopCode binop of
      Just code -> do { l <- evalVexp ctx left; r <- evalVexp ctx right; out (c_vdl_op_binary ctx code l r) }
      Nothing -> error "Exec: run Vlite.loweringPass first (Vlite.hs:1335-1337)"
    go V.Shuffle { V.shop = V.Gather, V.shsource, V.shpos } = do
      s <- evalVexp ctx shsource; p <- evalVexp ctx shpos; out (c_vdl_op_gather ctx s p)
    go V.Shuffle { V.shop = V.Scatter, V.shsource, V.shpos } = do
      s <- evalVexp ctx shsource; p <- evalVexp ctx shpos
      -- explicit output length (App. G2): the index space of the positions, which the library tracks (FoldSelect / pos_ /
      -- Partition positions index their input's rows -- for a join scatter that is the dimension's row count, the
      -- reference's `dimref`, Vlite.hs:782 --, `p % k` indexes k slots).  The node's metadata count (Vlite.hs:316-320) is
      -- off by one and data-independent, so it is not used.
      n <- liftIO $ alloca $ \q -> c_vdl_vec_index_space ctx p q >>= check ctx >> peek q
      if n < 0 then error "Exec: Scatter positions without a known index space" else out (c_vdl_op_scatter ctx s p n)
    go V.Fold { V.foldop = V.FSel, V.fdata } = do { d <- evalVexp ctx fdata; out (c_vdl_op_fold_select ctx d) }
    go V.Fold { V.foldop, V.fgroups, V.fdata } = do
      g <- evalVexp ctx fgroups; d <- evalVexp ctx fdata
      let code = case foldop of { V.FSum -> 0; V.FMin -> 1; V.FMax -> 2; V.FChoose -> 3; V.FSel -> error "handled above" }
      out (c_vdl_op_fold ctx code g d)
    go V.Partition { V.pdata, V.pivots = V.Vexp { V.vx = V.RangeC { V.rmin, V.rstep, V.rcount } } } = do
      d <- evalVexp ctx pdata
      out (c_vdl_op_partition ctx d (fromInteger rmin) (fromInteger rstep) (fromInteger rcount))
    go V.Like { V.ldata, V.lpattern, V.lcol } = do           -- the dictionary is the column's string heap (Vdl.hs:244-247)
      d <- evalVexp ctx ldata
      h <- out $ \p -> withCString (show lcol ++ ".heap") $ \s -> c_vdl_column_lookup ctx s p
      out (\p -> C.useAsCString lpattern $ \pat -> c_vdl_op_like ctx d h pat p)
    go V.VShuffle { V.varg } = evalVexp ctx varg
    go V.CrossProduct { V.left, V.right, V.variant } = do      -- joins under --use_cross_product (Vlite.hs:89-93, 672-680)
      l <- evalVexp ctx left; r <- evalVexp ctx right
      out (c_vdl_op_cross_product ctx l r (case variant of { V.COuter -> 0; V.CInner -> 1 }))
    go other = error ("Exec: op outside the supported vocabulary: " ++ show other)
