{-# LANGUAGE ForeignFunctionInterface #-}
{-# LANGUAGE EmptyDataDecls           #-}

--------------------------------------------------------------------------------
-- |
-- Module      :  Data.Tensor.Device.FFI
--
-- Raw @foreign import ccall@ bindings of @include/t5b200.h@ (libt5b200.so), one
-- per entry point, exactly the boundary a device-backed sibling of
-- "Data.Tensor.Vector" binds.  NOT TYPE-CHECKED in this repository: the image has
-- no GHC (SURVEY.md §0 F2).  The same entry points are exercised from Python
-- (ctypes) and C++ in @tests/@, so the ABI itself is fully covered.
--
-- Every call is @safe@: kernels run for micro- to milliseconds and an @unsafe@
-- call would stall the calling capability and the GC (SURVEY.md §8b, Threading).
--------------------------------------------------------------------------------

module Data.Tensor.Device.FFI where

import           Data.Word
import           Foreign.C.String
import           Foreign.C.Types
import           Foreign.Ptr

data Ctx   -- ^ opaque @t5_ctx@
data Buf   -- ^ opaque @t5_buf@

-- element types / layouts / status codes of the header
t5F64, t5F32, t5I64, t5U8, t5AOS, t5SOA :: CInt
t5F64 = 0; t5F32 = 1; t5I64 = 2; t5U8 = 3; t5AOS = 0; t5SOA = 1

t5Ok, t5ErrShape, t5ErrUnsupported, t5ErrCuda, t5ErrNccl, t5ErrOom, t5ErrInvalid :: CInt
t5Ok = 0; t5ErrShape = 1; t5ErrUnsupported = 2; t5ErrCuda = 3; t5ErrNccl = 4; t5ErrOom = 5; t5ErrInvalid = 6

foreign import ccall safe "t5_strerror"   c_strerror   :: CInt -> IO CString
foreign import ccall safe "t5_last_error" c_last_error :: Ptr Ctx -> IO CString
foreign import ccall safe "t5_version"    c_version    :: IO CString

foreign import ccall safe "t5_init"     c_init     :: Ptr CInt -> CInt -> Ptr (Ptr Ctx) -> IO CInt
foreign import ccall safe "t5_shutdown" c_shutdown :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_live_buffers" c_live_buffers :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_ndev"     c_ndev     :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_sync"     c_sync     :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_partition"
  c_partition :: CSize -> CInt -> CInt -> Ptr CSize -> Ptr CSize -> IO CInt

foreign import ccall safe "t5_alloc"
  c_alloc :: Ptr Ctx -> CInt -> CSize -> CSize -> CInt -> Ptr (Ptr Buf) -> IO CInt
foreign import ccall safe "t5_free"     c_free     :: Ptr Buf -> IO CInt
-- | finalizer for 'Foreign.ForeignPtr.newForeignPtr'; t5_free is thread-safe and
-- sets its own device, as a finalizer running on an arbitrary OS thread requires.
foreign import ccall safe "&t5_free"    p_free     :: FunPtr (Ptr Buf -> IO ())
foreign import ccall safe "t5_upload"   c_upload   :: Ptr Buf -> Ptr () -> CSize -> IO CInt
foreign import ccall safe "t5_download" c_download :: Ptr Buf -> Ptr () -> CSize -> IO CInt
foreign import ccall safe "t5_relayout" c_relayout :: Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_pinned_alloc" c_pinned_alloc :: CSize -> IO (Ptr ())
foreign import ccall safe "t5_pinned_free"  c_pinned_free  :: Ptr () -> IO ()
foreign import ccall safe "t5_buf_batch"  c_buf_batch  :: Ptr Buf -> IO CSize
foreign import ccall safe "t5_buf_elems"  c_buf_elems  :: Ptr Buf -> IO CSize
foreign import ccall safe "t5_buf_dtype"  c_buf_dtype  :: Ptr Buf -> IO CInt
foreign import ccall safe "t5_buf_layout" c_buf_layout :: Ptr Buf -> IO CInt

foreign import ccall safe "t5_matmul"
  c_matmul :: Ptr Ctx -> CInt -> CInt -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_dot"
  c_dot :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_dot_sum"
  c_dot_sum :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr () -> IO CInt
foreign import ccall safe "t5_dot_sum_used_nccl" c_dot_sum_used_nccl :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_matmul_is_exact" c_matmul_is_exact :: CInt -> CInt -> CInt -> CInt -> IO CInt
foreign import ccall safe "t5_prod"
  c_prod :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> CInt -> Ptr Word64 -> CInt
         -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_transpose"
  c_transpose :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_det"
  c_det :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_inverse"
  c_inverse :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_solve"
  c_solve :: Ptr Ctx -> CInt -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_zip"
  c_zip :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_scale"
  c_scale :: Ptr Ctx -> CInt -> Ptr () -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_replicate"
  c_replicate :: Ptr Ctx -> CInt -> Ptr () -> Ptr Buf -> IO CInt

-- the rest of the old SquareMatrix class (CHANGELOG.md:28-29; SURVEY.md §8f-4)
foreign import ccall safe "t5_trace"
  c_trace :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_poly_eval"
  c_poly_eval :: Ptr Ctx -> CInt -> CInt -> Ptr () -> CInt -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_char_poly"
  c_char_poly :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_min_poly"
  c_min_poly :: Ptr Ctx -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_row_echelon"
  c_row_echelon :: Ptr Ctx -> CInt -> CInt -> CInt -> CInt -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt

-- structural operations (src/Data/Tensor/Vector.hs:260-354, :400-463) as device gathers
foreign import ccall safe "t5_append"
  c_append :: Ptr Ctx -> CInt -> CInt -> CInt -> Ptr Word64 -> Ptr Word64 -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_split"
  c_split :: Ptr Ctx -> CInt -> CInt -> CInt -> Ptr Word64 -> Word64 -> Ptr Buf -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_at"
  c_at :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> Ptr Word64 -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_ta"
  c_ta :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> Word64 -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_slice"
  c_slice :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> Ptr Word64 -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_permute_axes"
  c_permute_axes :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> Ptr CInt -> Ptr Buf -> Ptr Buf -> IO CInt
foreign import ccall safe "t5_structural_table"
  c_structural_table :: CInt -> CInt -> Ptr Word64 -> Ptr Word64 -> CInt -> Ptr Word32 -> CSize -> Ptr CSize -> IO CInt

-- host-buffer (end-to-end) entries: operands and results are ordinary (ideally pinned) host arrays
foreign import ccall safe "t5_matmul_host"
  c_matmul_host :: Ptr Ctx -> CInt -> CInt -> CInt -> CInt -> CSize -> Ptr () -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "t5_dot_host"
  c_dot_host :: Ptr Ctx -> CInt -> CInt -> CSize -> Ptr () -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "t5_prod_host"
  c_prod_host :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> CInt -> Ptr Word64 -> CInt -> CSize
              -> Ptr () -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "t5_transpose_host"
  c_transpose_host :: Ptr Ctx -> CInt -> CInt -> Ptr Word64 -> CSize -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "t5_det_host"
  c_det_host :: Ptr Ctx -> CInt -> CInt -> CSize -> Ptr () -> Ptr () -> IO CInt
foreign import ccall safe "t5_inverse_host"
  c_inverse_host :: Ptr Ctx -> CInt -> CInt -> CSize -> Ptr () -> Ptr () -> Ptr Word8 -> IO CInt
foreign import ccall safe "t5_solve_host"
  c_solve_host :: Ptr Ctx -> CInt -> CInt -> CInt -> CSize -> Ptr () -> Ptr () -> Ptr () -> Ptr Word8 -> IO CInt

foreign import ccall safe "t5_timer_begin" c_timer_begin :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_timer_end"   c_timer_end   :: Ptr Ctx -> Ptr CFloat -> IO CInt
foreign import ccall safe "t5_timer_per_device"
  c_timer_per_device :: Ptr Ctx -> Ptr CFloat -> CInt -> IO CInt
foreign import ccall safe "t5_flush_l2"    c_flush_l2    :: Ptr Ctx -> IO CInt
foreign import ccall safe "t5_fill_synthetic" c_fill_synthetic :: Ptr Buf -> Word64 -> Word64 -> IO CInt
foreign import ccall safe "t5_plant_specials" c_plant_specials :: Ptr Buf -> CInt -> Word64 -> Word64 -> IO CInt
foreign import ccall safe "t5_launch_count"   c_launch_count   :: Ptr Ctx -> IO Word64
