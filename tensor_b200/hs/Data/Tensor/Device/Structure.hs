{-# LANGUAGE DataKinds           #-}
{-# LANGUAGE FlexibleContexts    #-}
{-# LANGUAGE KindSignatures      #-}
{-# LANGUAGE ScopedTypeVariables #-}

--------------------------------------------------------------------------------
-- |
-- Module      :  Data.Tensor.Device.Structure
--
-- The reshuffles of the reference's @IsTensor@ / @MultiIndexable@ / @Sliceable@
-- instances for "Data.Tensor.Vector" (src/Data/Tensor/Vector.hs:260-354,
-- :400-463) on device batches: each is ONE gather kernel behind one @ccall@.
-- Indices and axes are 1-based like the reference's MultiIndex
-- (src/Data/Tensor/Vector.hs:150-168).  The result shapes are the ones the
-- reference's type families compute (@Append@, @SplitAt@, @Slice@ ...); they are
-- taken here from the result type's 'SingI' dictionary and checked by the
-- library against the operand shapes (T5_ERR_SHAPE -> 'WrongListLength').
--
-- NOT TYPE-CHECKED here (no GHC in the image, SURVEY.md §0 F2).
--------------------------------------------------------------------------------

module Data.Tensor.Device.Structure
    ( appendB, splitB, atB, taB, sliceB, permuteAxesB
    ) where

import           Data.Proxy
import           Data.Singletons
import           Data.Word
import           Foreign.Marshal.Array  (withArray, withArrayLen)

import           Data.Tensor.Device
import           Data.Tensor.Device.FFI

w64 :: [Int] -> [Word64]
w64 = map fromIntegral

-- | @append n x y@ (Vector.hs:322-335): concatenate along axis @n@ (1-based).
appendB :: forall is js ks e. (SingI is, SingI js, SingI ks, DeviceElt e)
        => Int -> Batch is e -> Batch js e -> IO (Batch ks e)
appendB axis x y = do
    let Context h = batchCtx x
    z <- newBatch (batchCtx x) (batchCount x)
    withArrayLen (w64 (batchShape x)) $ \r pdx -> withArray (w64 (batchShape y)) $ \pdy ->
      unsafeWithBatch x $ \px -> unsafeWithBatch y $ \py -> unsafeWithBatch z $ \pz ->
        c_append h (dtypeCode (Proxy :: Proxy e)) (fromIntegral axis) (fromIntegral r) pdx pdy px py pz
            >>= checkRC (batchCtx x) "t5_append"
    return z

-- | @split n t@ (Vector.hs:336-354): the inverse of 'appendB'; the first result keeps the
-- first @k@ entries of axis @n@.
splitB :: forall is js ks e. (SingI is, SingI js, SingI ks, DeviceElt e)
       => Int -> Int -> Batch is e -> IO (Batch js e, Batch ks e)
splitB axis k t = do
    let Context h = batchCtx t
    a <- newBatch (batchCtx t) (batchCount t)
    b <- newBatch (batchCtx t) (batchCount t)
    withArrayLen (w64 (batchShape t)) $ \r pd ->
      unsafeWithBatch t $ \pt -> unsafeWithBatch a $ \pa -> unsafeWithBatch b $ \pb ->
        c_split h (dtypeCode (Proxy :: Proxy e)) (fromIntegral axis) (fromIntegral r) pd (fromIntegral k) pt pa pb
            >>= checkRC (batchCtx t) "t5_split"
    return (a, b)

-- | @t `at` ix@ (Vector.hs:277-283): fix every index but the first.
atB :: forall is i e. (SingI is, SingI i, DeviceElt e) => Batch is e -> [Int] -> IO (Batch '[i] e)
atB t ix = do
    let Context h = batchCtx t
    o <- newBatch (batchCtx t) (batchCount t)
    withArrayLen (w64 (batchShape t)) $ \r pd -> withArray (w64 ix) $ \pix ->
      unsafeWithBatch t $ \pt -> unsafeWithBatch o $ \po ->
        c_at h (dtypeCode (Proxy :: Proxy e)) (fromIntegral r) pd pix pt po >>= checkRC (batchCtx t) "t5_at"
    return o

-- | @[i] `ta` t@: fix the first index (GADT semantics, src/Data/Tensor.hs:245-247; SURVEY N1).
taB :: forall is js e. (SingI is, SingI js, DeviceElt e) => Int -> Batch is e -> IO (Batch js e)
taB i t = do
    let Context h = batchCtx t
    o <- newBatch (batchCtx t) (batchCount t)
    withArrayLen (w64 (batchShape t)) $ \r pd ->
      unsafeWithBatch t $ \pt -> unsafeWithBatch o $ \po ->
        c_ta h (dtypeCode (Proxy :: Proxy e)) (fromIntegral r) pd (fromIntegral i) pt po >>= checkRC (batchCtx t) "t5_ta"
    return o

-- | @slice t sl@ (Vector.hs:400-429, :461-463): @Nothing@ keeps an axis, @Just k@ fixes it.
sliceB :: forall is js e. (SingI is, SingI js, DeviceElt e) => Batch is e -> [Maybe Int] -> IO (Batch js e)
sliceB t sl = do
    let Context h = batchCtx t
    o <- newBatch (batchCtx t) (batchCount t)
    withArrayLen (w64 (batchShape t)) $ \r pd -> withArray (map (maybe 0 fromIntegral) sl) $ \ps ->
      unsafeWithBatch t $ \pt -> unsafeWithBatch o $ \po ->
        c_slice h (dtypeCode (Proxy :: Proxy e)) (fromIntegral r) pd ps pt po >>= checkRC (batchCtx t) "t5_slice"
    return o

-- | Output axis @a@ is input axis @perm !! a@ (0-based): @concat@ / @unConcat@ / @rev@ /
-- @unRev@ of nested tensors (Vector.hs:260-301) are axis permutations of the flat layout.
permuteAxesB :: forall is js e. (SingI is, SingI js, DeviceElt e) => [Int] -> Batch is e -> IO (Batch js e)
permuteAxesB perm t = do
    let Context h = batchCtx t
    o <- newBatch (batchCtx t) (batchCount t)
    withArrayLen (w64 (batchShape t)) $ \r pd -> withArray (map fromIntegral perm) $ \pp ->
      unsafeWithBatch t $ \pt -> unsafeWithBatch o $ \po ->
        c_permute_axes h (dtypeCode (Proxy :: Proxy e)) (fromIntegral r) pd pp pt po
            >>= checkRC (batchCtx t) "t5_permute_axes"
    return o
