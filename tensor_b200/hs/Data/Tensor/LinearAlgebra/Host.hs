{-# LANGUAGE DataKinds           #-}
{-# LANGUAGE FlexibleContexts    #-}
{-# LANGUAGE KindSignatures      #-}
{-# LANGUAGE ScopedTypeVariables #-}

--------------------------------------------------------------------------------
-- |
-- Module      :  Data.Tensor.LinearAlgebra.Host
--
-- The end-to-end form of the hot path: operands and results are ORDINARY host
-- batches (a storable vector holding @count@ row-major items, the wire order of
-- src/Data/Tensor/Vector.hs:360-367), and each function is ONE @ccall@ of a
-- @t5_*_host@ entry, which streams chunks through the GPU (H2D, kernel, D2H
-- overlapped on three slots) and returns with the results in host memory.  This
-- is what a caller that keeps its data in Haskell values uses; callers that chain
-- operations keep 'Data.Tensor.Device.Batch'es on the device instead and pay the
-- PCIe transfer once.
--
-- NOT TYPE-CHECKED here (no GHC in the image, SURVEY.md §0 F2).
--------------------------------------------------------------------------------

module Data.Tensor.LinearAlgebra.Host
    ( HostBatch(..), hostShape
    , matmulH, dotH, prodH, transposeH, detH, inverseH, solveH
    ) where

import           Control.Exception            (throwIO)
import           Control.Monad                (when)
import           Data.Proxy
import           Data.Singletons
import qualified Data.Vector.Storable         as S
import qualified Data.Vector.Storable.Mutable as SM
import           Data.Word
import           Foreign.Marshal.Array        (withArrayLen)
import           Foreign.Ptr

import           Data.MultiIndex              (PI)
import           Data.Tensor                  (TensorException (..))
import           Data.Tensor.Device
import           Data.Tensor.Device.FFI

-- | @hostCount@ tensors of shape @is@, concatenated row-major.
data HostBatch (is :: [PI]) e = HostBatch { hostCount :: Int, hostData :: S.Vector e }

hostShape :: forall is e. SingI is => HostBatch is e -> [Int]
hostShape _ = itemShape (Proxy :: Proxy is)

w64 :: [Int] -> [Word64]
w64 = map fromIntegral

checkLen :: forall is e. (SingI is, S.Storable e) => HostBatch is e -> IO ()
checkLen b = when (S.length (hostData b) /= hostCount b * product (hostShape b)) $ throwIO WrongListLength

result :: forall is e. (SingI is, S.Storable e) => Int -> (Ptr () -> IO ()) -> IO (HostBatch is e)
result n fill = do
    mv <- SM.new (n * product (itemShape (Proxy :: Proxy is)))
    SM.unsafeWith mv (fill . castPtr)
    HostBatch n <$> S.unsafeFreeze mv

-- | Batched @.*.@ of @count@ pairs; shapes @[m,k]@ and @[k,n]@.
matmulH :: forall i j k e. (SingI i, SingI j, SingI k, DeviceElt e)
        => Context -> HostBatch '[i, j] e -> HostBatch '[j, k] e -> IO (HostBatch '[i, k] e)
matmulH ctx@(Context h) a b = do
    checkLen a; checkLen b
    let [m, kk] = hostShape a; [_, n] = hostShape b
    S.unsafeWith (hostData a) $ \pa -> S.unsafeWith (hostData b) $ \pb -> result (hostCount a) $ \pc ->
        c_matmul_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral m) (fromIntegral kk) (fromIntegral n)
                      (fromIntegral (hostCount a)) (castPtr pa) (castPtr pb) pc >>= checkRC ctx "t5_matmul_host"

dotH :: forall i e. (SingI i, DeviceElt e) => Context -> HostBatch '[i] e -> HostBatch '[i] e -> IO (HostBatch '[] e)
dotH ctx@(Context h) x y = do
    checkLen x; checkLen y
    let [n] = hostShape x
    S.unsafeWith (hostData x) $ \px -> S.unsafeWith (hostData y) $ \py -> result (hostCount x) $ \po ->
        c_dot_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral n) (fromIntegral (hostCount x)) (castPtr px) (castPtr py) po
            >>= checkRC ctx "t5_dot_host"

-- | @prod n@: contract the last @n@ indices of @a@ with the first @n@ of @b@ (@n = 0@: tensor product).
prodH :: forall is js ks e. (SingI is, SingI js, SingI ks, DeviceElt e)
      => Context -> Int -> HostBatch is e -> HostBatch js e -> IO (HostBatch ks e)
prodH ctx@(Context h) nc a b = do
    checkLen a; checkLen b
    withArrayLen (w64 (hostShape a)) $ \ra pda -> withArrayLen (w64 (hostShape b)) $ \rb pdb ->
      S.unsafeWith (hostData a) $ \pa -> S.unsafeWith (hostData b) $ \pb -> result (hostCount a) $ \pc ->
        c_prod_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral ra) pda (fromIntegral rb) pdb (fromIntegral nc)
                    (fromIntegral (hostCount a)) (castPtr pa) (castPtr pb) pc >>= checkRC ctx "t5_prod_host"

transposeH :: forall is js e. (SingI is, SingI js, DeviceElt e) => Context -> HostBatch is e -> IO (HostBatch js e)
transposeH ctx@(Context h) a = do
    checkLen a
    withArrayLen (w64 (hostShape a)) $ \r pd -> S.unsafeWith (hostData a) $ \pa -> result (hostCount a) $ \po ->
        c_transpose_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral r) pd (fromIntegral (hostCount a)) (castPtr pa) po
            >>= checkRC ctx "t5_transpose_host"

detH :: forall i e. (SingI i, DeviceElt e, Fractional e) => Context -> HostBatch '[i, i] e -> IO (HostBatch '[] e)
detH ctx@(Context h) a = do
    checkLen a
    let [n, _] = hostShape a
    S.unsafeWith (hostData a) $ \pa -> result (hostCount a) $ \po ->
        c_det_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral n) (fromIntegral (hostCount a)) (castPtr pa) po
            >>= checkRC ctx "t5_det_host"

-- | The reference's @Maybe@ per item comes back as a status vector (0 = Just, 1 = Nothing).
inverseH :: forall i e. (SingI i, DeviceElt e, Fractional e)
         => Context -> HostBatch '[i, i] e -> IO (HostBatch '[i, i] e, S.Vector Word8)
inverseH ctx@(Context h) a = do
    checkLen a
    let [n, _] = hostShape a
    st <- SM.new (hostCount a)
    x <- S.unsafeWith (hostData a) $ \pa -> SM.unsafeWith st $ \ps -> result (hostCount a) $ \po ->
        c_inverse_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral n) (fromIntegral (hostCount a)) (castPtr pa) po ps
            >>= checkRC ctx "t5_inverse_host"
    (,) x <$> S.unsafeFreeze st

solveH :: forall i e. (SingI i, DeviceElt e, Fractional e)
       => Context -> HostBatch '[i, i] e -> HostBatch '[i] e -> IO (HostBatch '[i] e, S.Vector Word8)
solveH ctx@(Context h) a b = do
    checkLen a; checkLen b
    let [n, _] = hostShape a
    st <- SM.new (hostCount a)
    x <- S.unsafeWith (hostData a) $ \pa -> S.unsafeWith (hostData b) $ \pb -> SM.unsafeWith st $ \ps ->
         result (hostCount a) $ \po ->
        c_solve_host h (dtypeCode (Proxy :: Proxy e)) (fromIntegral n) 1 (fromIntegral (hostCount a))
                     (castPtr pa) (castPtr pb) po ps >>= checkRC ctx "t5_solve_host"
    (,) x <$> S.unsafeFreeze st
