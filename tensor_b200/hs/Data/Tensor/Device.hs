{-# LANGUAGE DataKinds           #-}
{-# LANGUAGE FlexibleContexts    #-}
{-# LANGUAGE KindSignatures      #-}
{-# LANGUAGE ScopedTypeVariables #-}
{-# LANGUAGE TypeFamilies        #-}

--------------------------------------------------------------------------------
-- |
-- Module      :  Data.Tensor.Device
--
-- Device-resident batches of fixed-shape tensors: the storage the B200 backend
-- substitutes for @Data.Tensor.Vector.Tensor {form, content}@
-- (reference: src/Data/Tensor/Vector.hs:121-124).  The per-item shape stays at
-- the type level (@is :: ['PI']@, src/Data/MultiIndex.hs:46-61) exactly as in the
-- reference; the batch count is a run-time value because unary type-level
-- naturals cannot express 2^24 (SURVEY.md §2 N6, §7 H1).
--
-- Marshalling follows the reference's 'IsList' instance
-- (src/Data/Tensor/Vector.hs:360-367): row-major, last index fastest; a list of
-- the wrong length throws 'WrongListLength' (src/Data/Tensor.hs:96-100).  Boxed,
-- lazy elements (N2) are forced ('rnf', Vector.hs:372-373) and copied through
-- 'Storable' before upload.
--
-- NOT TYPE-CHECKED here (no GHC in the image, SURVEY.md §0 F2).
--------------------------------------------------------------------------------

module Data.Tensor.Device
    ( Context(..), withContext, contextDevices, liveBuffers
    , DeviceElt(..)
    , Batch, batchCount, batchShape
    , fromTensors, toTensors, fromListB, toListB
    , TensorDeviceException(..)
    -- * internals used by Data.Tensor.LinearAlgebra.Device
    , unsafeWithBatch, newBatch, newStatusBatch, itemShape, checkRC, batchCtx
    ) where

import           Control.DeepSeq            (NFData, force)
import           Control.Exception
import           Control.Monad              (when)
import           Data.Int                   (Int64)
import           Data.Proxy
import           Data.Singletons
import           Data.Typeable              (Typeable)
import           Data.Word                  (Word8)
import qualified Data.Vector.Storable       as S
import           Foreign.C.String           (peekCString)
import           Foreign.C.Types
import           Foreign.ForeignPtr
import           Foreign.Marshal.Alloc      (alloca)
import           Foreign.Marshal.Array      (withArrayLen)
import           Foreign.Ptr
import           Foreign.Storable
import           GHC.Exts                   (IsList (toList))

import           Data.MultiIndex            (PI, Shape, fromShape)
import           Data.Tensor                (TensorException (..))
import qualified Data.Tensor.Vector         as V
import           Data.Tensor.Device.FFI

-- | Errors of the device layer that are not 'WrongListLength'.
data TensorDeviceException
    = Unsupported String   -- ^ T5_ERR_UNSUPPORTED: no kernel; there is no CPU fallback
    | DeviceError CInt String
      deriving (Show, Typeable)
instance Exception TensorDeviceException

-- | A set of GPUs with one stream each; the batch axis of every 'Batch' created
-- in it is sharded over them (t5_partition).
newtype Context = Context (Ptr Ctx)

-- | Run an action with a device context.  @t5_shutdown@ runs when the action ends - typically BEFORE the garbage
-- collector has run the @t5_free@ finalizers of the batches the action created.  That order is legal by contract
-- (include/t5b200.h, t5_shutdown): the library releases the device memory of the still-live buffers at shutdown
-- and turns them into orphans, every later operation on one of them throws 'DeviceError' (T5_ERR_INVALID), and
-- the finalizer of an orphan only releases its handle.  No 'touchForeignPtr' choreography is needed.
withContext :: [Int] -> (Context -> IO a) -> IO a
withContext devs act =
    withArrayLen (map fromIntegral devs) $ \n p ->
    alloca $ \out -> do
      rc <- c_init (if null devs then nullPtr else p) (fromIntegral n) out
      when (rc /= t5Ok) $ throwIO (DeviceError rc "t5_init")
      h <- peek out
      act (Context h) `finally` c_shutdown h

contextDevices :: Context -> IO Int
contextDevices (Context h) = fromIntegral <$> c_ndev h

-- | Batches allocated in this context whose finalizer has not run yet.
liveBuffers :: Context -> IO Int
liveBuffers (Context h) = fromIntegral <$> c_live_buffers h

-- | Element types that have device kernels.
class (Storable e, NFData e) => DeviceElt e where
    dtypeCode :: Proxy e -> CInt
instance DeviceElt Double where dtypeCode _ = t5F64
instance DeviceElt Float  where dtypeCode _ = t5F32
instance DeviceElt Int64  where dtypeCode _ = t5I64
instance DeviceElt Int    where dtypeCode _ = t5I64   -- GHC Int is 64-bit two's complement

-- | @count@ tensors of shape @is@ and element type @e@ in HBM.
data Batch (is :: [PI]) e = Batch
    { batchCtx   :: Context
    , batchCount :: Int
    , batchBuf   :: ForeignPtr Buf   -- finalizer: t5_free
    }

itemShape :: forall is proxy. SingI is => proxy is -> [Int]
itemShape _ = fromShape (sing :: Shape is)

batchShape :: forall is e. SingI is => Batch is e -> [Int]
batchShape _ = itemShape (Proxy :: Proxy is)

checkRC :: Context -> String -> CInt -> IO ()
checkRC (Context h) what rc
    | rc == t5Ok             = return ()
    | rc == t5ErrShape       = throwIO WrongListLength
    | rc == t5ErrUnsupported = throwIO (Unsupported what)
    | otherwise              = do msg <- c_last_error h >>= peekCString
                                  throwIO (DeviceError rc (what ++ ": " ++ msg))

newBatch :: forall is e. (SingI is, DeviceElt e) => Context -> Int -> IO (Batch is e)
newBatch ctx@(Context h) n =
    alloca $ \out -> do
      let d = product (itemShape (Proxy :: Proxy is))
      c_alloc h (dtypeCode (Proxy :: Proxy e)) (fromIntegral d) (fromIntegral n) t5AOS out
          >>= checkRC ctx "t5_alloc"
      Batch ctx n <$> (peek out >>= newForeignPtr p_free)

-- | One status byte per item (the 'Maybe' of inverse / solve): 0 = Just, 1 = Nothing.
newStatusBatch :: Context -> Int -> IO (Batch '[] Word8)
newStatusBatch ctx@(Context h) n =
    alloca $ \out -> do
      c_alloc h t5U8 1 (fromIntegral n) t5AOS out >>= checkRC ctx "t5_alloc"
      Batch ctx n <$> (peek out >>= newForeignPtr p_free)

unsafeWithBatch :: Batch is e -> (Ptr Buf -> IO a) -> IO a
unsafeWithBatch = withForeignPtr . batchBuf

-- | Upload a flat row-major list holding @n@ tensors (the batched 'fromList').
fromListB :: forall is e. (SingI is, DeviceElt e) => Context -> Int -> [e] -> IO (Batch is e)
fromListB ctx n xs = do
    let d  = product (itemShape (Proxy :: Proxy is))
        sv = S.fromList (force xs)
    when (S.length sv /= n * d) $ throwIO WrongListLength   -- Vector.hs:364-366
    b <- newBatch ctx n
    S.unsafeWith sv $ \p -> unsafeWithBatch b $ \hb ->
        c_upload hb (castPtr p) (fromIntegral (S.length sv * sizeOf (undefined :: e)))
            >>= checkRC ctx "t5_upload"
    return b

toListB :: forall is e. (SingI is, DeviceElt e) => Batch is e -> IO [e]
toListB b = do
    let len = batchCount b * product (itemShape (Proxy :: Proxy is))
    fp <- mallocForeignPtrArray len
    withForeignPtr fp $ \p -> unsafeWithBatch b $ \hb ->
        c_download hb (castPtr p) (fromIntegral (len * sizeOf (undefined :: e)))
            >>= checkRC (batchCtx b) "t5_download"
    return (S.toList (S.unsafeFromForeignPtr0 fp len))

-- | Marshal ordinary "Data.Tensor.Vector" tensors (their 'toList' order IS the wire order).
fromTensors :: (SingI is, DeviceElt e) => Context -> [V.Tensor is e] -> IO (Batch is e)
fromTensors ctx ts = fromListB ctx (length ts) (concatMap toList ts)

toTensors :: forall is e. (SingI is, DeviceElt e) => Batch is e -> IO [V.Tensor is e]
toTensors b = do
    xs <- toListB b
    let d = product (itemShape (Proxy :: Proxy is))
        chunks [] = []
        chunks ys = let (h, t) = splitAt d ys in V.fromList h : chunks t
    return (chunks xs)
