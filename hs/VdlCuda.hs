{-# LANGUAGE ForeignFunctionInterface #-}
-- | Raw FFI bindings to libvdl_cuda (include/vdl_cuda.h), the B200 executor for the Voodoo programs mplan2vdl
-- emits.  SOURCE ONLY: this image has no GHC, so this module has never been compiled; it mirrors, call for call,
-- the ctypes binding in mplan2vdl_b200/lib.py that IS exercised by the test-suite.
--
-- Every import is `safe`: the calls block (kernel launches + stream synchronisation) and must not stall the
-- Haskell runtime's capability.  Handles: VdlCtx / VdlPlan / VdlFused are opaque pointers; vectors are Int32
-- handles owned by the context.
module VdlCuda where

import Data.Int (Int32, Int64)
import Data.Word (Word64)
import Foreign.C.String (CString)
import Foreign.C.Types (CInt (..))
import Foreign.Ptr (Ptr)

data VdlCtx
data VdlPlan
data VdlFused
data VdlProbe
data VdlMapDesc   -- vdl_map_desc: marshalled with Foreign.Storable by the caller (layout in vdl_cuda.h)
type VdlVec = Int32

-- status codes (vdl_cuda.h)
vdlOk, vdlEInval, vdlECuda, vdlENotFound, vdlEUnsupported, vdlERange, vdlENoMem :: CInt
vdlOk = 0; vdlEInval = 1; vdlECuda = 2; vdlENotFound = 3; vdlEUnsupported = 4; vdlERange = 5; vdlENoMem = 6

-- storage types: Types.hs sizeOf SInt32 = 4, SInt64 / SDecimal = 8
vdlI32, vdlI64 :: CInt
vdlI32 = 4; vdlI64 = 8

foreign import ccall safe "vdl_abi_version" c_vdl_abi_version :: IO CInt
foreign import ccall safe "vdl_ctx_create" c_vdl_ctx_create :: CInt -> Ptr (Ptr VdlCtx) -> IO CInt
foreign import ccall safe "vdl_ctx_destroy" c_vdl_ctx_destroy :: Ptr VdlCtx -> IO CInt
foreign import ccall safe "vdl_last_error" c_vdl_last_error :: Ptr VdlCtx -> IO CString

-- columns: what `Load n` (Vlite.hs Vx) binds
foreign import ccall safe "vdl_column_alloc" c_vdl_column_alloc :: Ptr VdlCtx -> CString -> CInt -> Int64 -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_column_upload" c_vdl_column_upload :: Ptr VdlCtx -> VdlVec -> Ptr () -> Int64 -> IO CInt
foreign import ccall safe "vdl_column_fill_synthetic" c_vdl_column_fill_synthetic
  :: Ptr VdlCtx -> VdlVec -> Word64 -> Word64 -> CInt -> Int64 -> Int64 -> Int64 -> Int64 -> Int64 -> IO CInt
foreign import ccall safe "vdl_column_lookup" c_vdl_column_lookup :: Ptr VdlCtx -> CString -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_vec_len" c_vdl_vec_len :: Ptr VdlCtx -> VdlVec -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_vec_download" c_vdl_vec_download :: Ptr VdlCtx -> VdlVec -> Ptr Int64 -> Int64 -> IO CInt
foreign import ccall safe "vdl_vec_free" c_vdl_vec_free :: Ptr VdlCtx -> VdlVec -> IO CInt

-- one entry point per Vx constructor (Vlite.hs:102-116)
foreign import ccall safe "vdl_op_range" c_vdl_op_range :: Ptr VdlCtx -> Int64 -> Int64 -> Int64 -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_op_binary" c_vdl_op_binary :: Ptr VdlCtx -> CInt -> VdlVec -> VdlVec -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_op_fold_select" c_vdl_op_fold_select :: Ptr VdlCtx -> VdlVec -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_op_map" c_vdl_op_map :: Ptr VdlCtx -> Ptr VdlMapDesc -> Ptr VdlVec -> Ptr VdlVec -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_jit_selftest" c_vdl_jit_selftest :: CString -> CInt -> IO CInt
foreign import ccall unsafe "vdl_abi_sizeof_map_desc" c_vdl_abi_sizeof_map_desc :: IO CInt
foreign import ccall safe "vdl_op_gather" c_vdl_op_gather :: Ptr VdlCtx -> VdlVec -> VdlVec -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_op_scatter" c_vdl_op_scatter :: Ptr VdlCtx -> VdlVec -> VdlVec -> Int64 -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_op_partition" c_vdl_op_partition :: Ptr VdlCtx -> VdlVec -> Int64 -> Int64 -> Int64 -> Ptr VdlVec -> IO CInt
foreign import ccall safe "vdl_op_fold" c_vdl_op_fold :: Ptr VdlCtx -> CInt -> VdlVec -> VdlVec -> Ptr VdlVec -> IO CInt

-- whole programs: the text Vdl.vdlFromVexps prints
foreign import ccall safe "vdl_plan_load" c_vdl_plan_load :: Ptr VdlCtx -> CString -> CInt -> Ptr (Ptr VdlPlan) -> IO CInt
foreign import ccall safe "vdl_plan_set_row_base" c_vdl_plan_set_row_base :: Ptr VdlPlan -> Int64 -> IO CInt
foreign import ccall safe "vdl_plan_run" c_vdl_plan_run :: Ptr VdlPlan -> IO CInt
foreign import ccall safe "vdl_plan_run_local" c_vdl_plan_run_local :: Ptr VdlPlan -> IO CInt
foreign import ccall safe "vdl_plan_finish" c_vdl_plan_finish :: Ptr VdlPlan -> Ptr (Ptr ()) -> CInt -> IO CInt
foreign import ccall safe "vdl_plan_num_outputs" c_vdl_plan_num_outputs :: Ptr VdlPlan -> IO CInt
foreign import ccall safe "vdl_plan_output" c_vdl_plan_output :: Ptr VdlPlan -> CInt -> Ptr CString -> Ptr (Ptr Int64) -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_plan_destroy" c_vdl_plan_destroy :: Ptr VdlPlan -> IO CInt

-- multi-GPU combine over peer memory (vdl_cuda.h "multi-GPU combine over peer memory"): after vdl_plan_set_peers,
-- vdl_plan_run returns the GLOBAL result on every rank with one kernel launch per GPU and no collective library
foreign import ccall safe "vdl_plan_exchange_bytes" c_vdl_plan_exchange_bytes :: Ptr VdlPlan -> CInt -> CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_plan_set_peers" c_vdl_plan_set_peers :: Ptr VdlPlan -> CInt -> CInt -> CInt -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "vdl_ipc_alloc" c_vdl_ipc_alloc :: Ptr VdlCtx -> Int64 -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "vdl_ipc_export" c_vdl_ipc_export :: Ptr VdlCtx -> Ptr () -> CString -> IO CInt      -- 64-byte handle
foreign import ccall safe "vdl_ipc_open" c_vdl_ipc_open :: Ptr VdlCtx -> CString -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "vdl_ipc_close" c_vdl_ipc_close :: Ptr VdlCtx -> Ptr () -> IO CInt
foreign import ccall safe "vdl_ipc_free" c_vdl_ipc_free :: Ptr VdlCtx -> Ptr () -> IO CInt

-- what the fusion passes did with a loaded program (Folds on the fused scan; FK-join Folds / vectors on the probe kernel)
foreign import ccall safe "vdl_plan_stats" c_vdl_plan_stats :: Ptr VdlPlan -> Ptr CInt -> Ptr CInt -> Ptr CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_probe_exchange_bytes" c_vdl_probe_exchange_bytes :: Ptr VdlProbe -> CInt -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_probe_set_peers" c_vdl_probe_set_peers :: Ptr VdlProbe -> CInt -> CInt -> Ptr (Ptr ()) -> IO CInt
foreign import ccall safe "vdl_plan_map_stats" c_vdl_plan_map_stats :: Ptr VdlPlan -> Ptr CInt -> Ptr CInt -> IO CInt
foreign import ccall safe "vdl_plan_probe_stats" c_vdl_plan_probe_stats :: Ptr VdlPlan -> Ptr CInt -> Ptr CInt -> Ptr CInt -> IO CInt

-- sharded runs of FK-join plans: partial tables of every fused scan / probe fold group, and the vectors probe passes emit
foreign import ccall safe "vdl_plan_num_partials" c_vdl_plan_num_partials :: Ptr VdlPlan -> IO CInt
foreign import ccall safe "vdl_plan_partials" c_vdl_plan_partials :: Ptr VdlPlan -> CInt -> Ptr (Ptr ()) -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_plan_num_emits" c_vdl_plan_num_emits :: Ptr VdlPlan -> IO CInt
foreign import ccall safe "vdl_plan_emit" c_vdl_plan_emit :: Ptr VdlPlan -> CInt -> Ptr (Ptr ()) -> Ptr Int64 -> IO CInt
foreign import ccall safe "vdl_plan_emit_replace" c_vdl_plan_emit_replace :: Ptr VdlPlan -> CInt -> Ptr () -> Int64 -> IO CInt
