{-# LANGUAGE DataKinds              #-}
{-# LANGUAGE FlexibleContexts       #-}
{-# LANGUAGE KindSignatures         #-}
{-# LANGUAGE ScopedTypeVariables    #-}
{-# LANGUAGE TypeOperators          #-}

--------------------------------------------------------------------------------
-- |
-- Module      :  Data.Tensor.LinearAlgebra.Device
--
-- The IO layer under "Data.Tensor.LinearAlgebra": one function per operation,
-- each ONE @ccall@ into libt5b200.so on device batches ("Data.Tensor.Device").
-- Shapes come from the type-level index list exactly as Data.Tensor.Vector takes
-- them (@fromShape sing@, src/Data/Tensor/Vector.hs:189, :231, :363).  Results
-- are in 'IO' because they allocate device buffers; the pure class instances a
-- user of the reference sees are in "Data.Tensor.LinearAlgebra".
--
-- NOT TYPE-CHECKED here (no GHC in the image, SURVEY.md §0 F2).
--------------------------------------------------------------------------------

module Data.Tensor.LinearAlgebra.Device
    ( matmulB, matvecB, dotB, dotSum, tensorProductB, transposeB, detB, inverseB, solveB
    , prodN, tr, polyEval, charPoly, minPoly, rowEchelonForm, echelonWith
    , addB, subB, hadamardB, scaleB, replicateB, pureB, unitB
    ) where

import           Control.Exception     (throwIO)
import           Data.Proxy
import           Data.Singletons
import           Data.Word
import           Foreign.C.Types       (CInt)
import           Foreign.Marshal.Alloc (alloca)
import           Foreign.Marshal.Utils (with)
import           Foreign.Marshal.Array (withArrayLen)
import qualified Data.Vector.Storable  as S
import           Foreign.Ptr
import           Foreign.Storable

import           Data.MultiIndex       (PI (..), Shape, fromShape)
import           Data.Tensor           (TensorException (WrongListLength))
import           Data.Tensor.Device
import           Data.Tensor.Device.FFI

dims :: [Int] -> [Word64]
dims = map fromIntegral

code :: forall e. DeviceElt e => Proxy e -> CInt
code = dtypeCode

-- | matrix . matrix: @c[i,k] = fold_{j ascending} (acc + a[i,j] * b[j,k])@, @acc0 = 0@.
matmulB :: forall i j k e. (SingI i, SingI j, SingI k, DeviceElt e)
        => Batch '[i, j] e -> Batch '[j, k] e -> IO (Batch '[i, k] e)
matmulB a b = do
    let [m, kk] = batchShape a; [_, n] = batchShape b
        Context h = batchCtx a
    c <- newBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch b $ \pb -> unsafeWithBatch c $ \pc ->
        c_matmul h (code (Proxy :: Proxy e)) (fromIntegral m) (fromIntegral kk) (fromIntegral n) pa pb pc
            >>= checkRC (batchCtx a) "t5_matmul"
    return c

-- | matrix . vector
matvecB :: forall i j e. (SingI i, SingI j, DeviceElt e) => Batch '[i, j] e -> Batch '[j] e -> IO (Batch '[i] e)
matvecB a x = do
    let [m, kk] = batchShape a
        Context h = batchCtx a
    y <- newBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch x $ \px -> unsafeWithBatch y $ \py ->
        c_matmul h (code (Proxy :: Proxy e)) (fromIntegral m) (fromIntegral kk) 1 pa px py
            >>= checkRC (batchCtx a) "t5_matmul"
    return y

-- | One scalar per pair: @fold_{i ascending} (acc + x_i * y_i)@, @acc0 = 0@.
dotB :: forall i e. (SingI i, DeviceElt e) => Batch '[i] e -> Batch '[i] e -> IO (Batch '[] e)
dotB x y = do
    let [n] = batchShape x
        Context h = batchCtx x
    s <- newBatch (batchCtx x) (batchCount x)
    unsafeWithBatch x $ \px -> unsafeWithBatch y $ \py -> unsafeWithBatch s $ \ps ->
        c_dot h (code (Proxy :: Proxy e)) (fromIntegral n) px py ps >>= checkRC (batchCtx x) "t5_dot"
    return s

-- | Sum over the batch of 'dot' — the only operation with a cross-GPU exchange
-- (1-element ncclAllReduce, SURVEY.md §8e).
dotSum :: forall i e. (SingI i, DeviceElt e) => Batch '[i] e -> Batch '[i] e -> IO e
dotSum x y = alloca $ \out -> do
    let [n] = batchShape x
        Context h = batchCtx x
    unsafeWithBatch x $ \px -> unsafeWithBatch y $ \py ->
        c_dot_sum h (code (Proxy :: Proxy e)) (fromIntegral n) px py (castPtr out)
            >>= checkRC (batchCtx x) "t5_dot_sum"
    peek out

-- | Outer product, one multiply per component; @ks@ must be @is ++ js@ (the pure class instance in
-- "Data.Tensor.LinearAlgebra" computes it; the library checks the element count).
tensorProductB :: forall is js ks e. (SingI is, SingI js, SingI ks, DeviceElt e)
               => Batch is e -> Batch js e -> IO (Batch ks e)
tensorProductB a b = do
    let da = batchShape a; db = batchShape b
        Context h = batchCtx a
    c <- newBatch (batchCtx a) (batchCount a)
    withArrayLen (dims da) $ \ra pda -> withArrayLen (dims db) $ \rb pdb ->
      unsafeWithBatch a $ \pa -> unsafeWithBatch b $ \pb -> unsafeWithBatch c $ \pc ->
        c_prod h (code (Proxy :: Proxy e)) (fromIntegral ra) pda (fromIntegral rb) pdb 0 pa pb pc
            >>= checkRC (batchCtx a) "t5_prod"
    return c

-- | Full index reversal - @transpose = unRev . t0@ (src/Data/Indexable.hs:101-102); @js@ must be
-- @Reverse is@.
transposeB :: forall is js e. (SingI is, SingI js, DeviceElt e) => Batch is e -> IO (Batch js e)
transposeB a = do
    let da = batchShape a
        Context h = batchCtx a
    t <- newBatch (batchCtx a) (batchCount a)
    withArrayLen (dims da) $ \r pd -> unsafeWithBatch a $ \pa -> unsafeWithBatch t $ \pt ->
        c_transpose h (code (Proxy :: Proxy e)) (fromIntegral r) pd pa pt >>= checkRC (batchCtx a) "t5_transpose"
    return t

detB :: forall i e. (SingI i, DeviceElt e, Fractional e) => Batch '[i, i] e -> IO (Batch '[] e)
detB a = do
    let [n, _] = batchShape a
        Context h = batchCtx a
    d <- newBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch d $ \pd ->
        c_det h (code (Proxy :: Proxy e)) (fromIntegral n) pa pd >>= checkRC (batchCtx a) "t5_det"
    return d
-- | The per-item 'Maybe' comes back as a status batch (0 = Just, 1 = Nothing).
inverseB :: forall i e. (SingI i, DeviceElt e, Fractional e) => Batch '[i, i] e -> IO (Batch '[i, i] e, Batch '[] Word8)
inverseB a = do
    let [n, _] = batchShape a
        Context h = batchCtx a
    x  <- newBatch (batchCtx a) (batchCount a)
    st <- newStatusBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch x $ \px -> unsafeWithBatch st $ \ps ->
        c_inverse h (code (Proxy :: Proxy e)) (fromIntegral n) pa px ps >>= checkRC (batchCtx a) "t5_inverse"
    return (x, st)

-- | Unique-solution square systems (first-non-zero pivoting, SURVEY §8a).
solveB :: forall i e. (SingI i, DeviceElt e, Fractional e)
       => Batch '[i, i] e -> Batch '[i] e -> IO (Batch '[i] e, Batch '[] Word8)
solveB a b = do
    let [n, _] = batchShape a
        Context h = batchCtx a
    x  <- newBatch (batchCtx a) (batchCount a)
    st <- newStatusBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch b $ \pb -> unsafeWithBatch x $ \px -> unsafeWithBatch st $ \ps ->
        c_solve h (code (Proxy :: Proxy e)) (fromIntegral n) 1 pa pb px ps >>= checkRC (batchCtx a) "t5_solve"
    return (x, st)

-- | @prodN n a b@ (@prod n@): contract the last @n@ indices of @a@ with the first @n@ of @b@.
prodN :: forall is js ks e. (SingI is, SingI js, SingI ks, DeviceElt e)
      => Int -> Batch is e -> Batch js e -> IO (Batch ks e)
prodN n a b = do
    let Context h = batchCtx a
    c <- newBatch (batchCtx a) (batchCount a)
    withArrayLen (dims (batchShape a)) $ \ra pda -> withArrayLen (dims (batchShape b)) $ \rb pdb ->
      unsafeWithBatch a $ \pa -> unsafeWithBatch b $ \pb -> unsafeWithBatch c $ \pc ->
        c_prod h (code (Proxy :: Proxy e)) (fromIntegral ra) pda (fromIntegral rb) pdb (fromIntegral n) pa pb pc
            >>= checkRC (batchCtx a) "t5_prod"
    return c

-- The rest of the old SquareMatrix class (CHANGELOG.md:28-29; SURVEY.md §8f-4) ------------

-- | Trace: @fold_{i ascending} (acc + a_ii)@ from 0.
tr :: forall i e. (SingI i, DeviceElt e) => Batch '[i, i] e -> IO (Batch '[] e)
tr a = do
    let [n, _] = batchShape a; Context h = batchCtx a
    o <- newBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch o $ \po ->
        c_trace h (code (Proxy :: Proxy e)) (fromIntegral n) pa po >>= checkRC (batchCtx a) "t5_trace"
    return o

-- | @polyEval cs a = c0 I + c1 a + .. + cd a^d@ (Horner), coefficients low order first.
polyEval :: forall i e. (SingI i, DeviceElt e) => [e] -> Batch '[i, i] e -> IO (Batch '[i, i] e)
polyEval cs a = do
    let [n, _] = batchShape a; Context h = batchCtx a
    o <- newBatch (batchCtx a) (batchCount a)
    S.unsafeWith (S.fromList cs) $ \pc -> unsafeWithBatch a $ \pa -> unsafeWithBatch o $ \po ->
        c_poly_eval h (code (Proxy :: Proxy e)) (fromIntegral n) (castPtr pc) (fromIntegral (length cs)) pa po
            >>= checkRC (batchCtx a) "t5_poly_eval"
    return o

-- | Coefficients of @det (xI - a)@, low order first (@n + 1@ per item; the result shape
-- @'[j]@ must be @i + 1@).
charPoly :: forall i j e. (SingI i, SingI j, DeviceElt e, Fractional e) => Batch '[i, i] e -> IO (Batch '[j] e)
charPoly a = do
    let [n, _] = batchShape a; Context h = batchCtx a
    o <- newBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch o $ \po ->
        c_char_poly h (code (Proxy :: Proxy e)) (fromIntegral n) pa po >>= checkRC (batchCtx a) "t5_char_poly"
    return o

-- | Minimal polynomial (zero-padded above its degree) and the degree per item.
minPoly :: forall i j e. (SingI i, SingI j, DeviceElt e, Fractional e)
        => Batch '[i, i] e -> IO (Batch '[j] e, Batch '[] Word8)
minPoly a = do
    let [n, _] = batchShape a; Context h = batchCtx a
    o <- newBatch (batchCtx a) (batchCount a)
    d <- newStatusBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch o $ \po -> unsafeWithBatch d $ \pd ->
        c_min_poly h (code (Proxy :: Proxy e)) (fromIntegral n) pa po pd >>= checkRC (batchCtx a) "t5_min_poly"
    return (o, d)

-- | Reduced row echelon form of every item and its rank (the @rowEchelonForm@ of the old
-- @Data.Tensor.LinearAlgebra@, CHANGELOG.md:24-29): Gauss-Jordan with the first-non-zero pivot rule.
rowEchelonForm :: forall i j e. (SingI i, SingI j, DeviceElt e, Fractional e)
               => Batch '[i, j] e -> IO (Batch '[i, j] e, Batch '[] Word8)
rowEchelonForm a = let [_, c] = batchShape a in echelonWith c a

-- | @triangularSolve@ on an augmented batch @[a | b]@ whose first @k@ columns are @a@: row operations
-- until @a@ is in reduced row echelon form, @b@ carried along; pivots are only looked for in @a@.
-- (Build the augmented batch with 'append' and cut the result with 'split'.)
echelonWith :: forall i j e. (SingI i, SingI j, DeviceElt e, Fractional e)
            => Int -> Batch '[i, j] e -> IO (Batch '[i, j] e, Batch '[] Word8)
echelonWith k a = do
    let [r, c] = batchShape a; Context h = batchCtx a
    o <- newBatch (batchCtx a) (batchCount a)
    d <- newStatusBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch o $ \po -> unsafeWithBatch d $ \pd ->
        c_row_echelon h (code (Proxy :: Proxy e)) (fromIntegral r) (fromIntegral c) (fromIntegral k) pa po pd
            >>= checkRC (batchCtx a) "t5_row_echelon"
    return (o, d)

-- Vector-space operations: what @liftA2 (+)@, @liftA2 (-)@, @liftA2 (*)@ and @fmap (k *)@ of the
-- Functor / Applicative instances (src/Data/Tensor/Vector.hs:180-190, :250-251) are on batches.
zipB :: forall is e. (SingI is, DeviceElt e) => CInt -> Batch is e -> Batch is e -> IO (Batch is e)
zipB op a b = do
    let Context h = batchCtx a
    c <- newBatch (batchCtx a) (batchCount a)
    unsafeWithBatch a $ \pa -> unsafeWithBatch b $ \pb -> unsafeWithBatch c $ \pc ->
        c_zip h (code (Proxy :: Proxy e)) op pa pb pc >>= checkRC (batchCtx a) "t5_zip"
    return c

addB, subB, hadamardB :: (SingI is, DeviceElt e) => Batch is e -> Batch is e -> IO (Batch is e)
addB = zipB 0
subB = zipB 1
hadamardB = zipB 2

-- | Every item a copy of one tensor given as its row-major element list (length checked against the
-- shape like 'fromList', src/Data/Tensor/Vector.hs:362-366).
replicateB :: forall is e. (SingI is, DeviceElt e) => Context -> Int -> [e] -> IO (Batch is e)
replicateB ctx@(Context h) n xs = do
    o <- newBatch ctx n
    let d = fromIntegral (product (batchShape o)) :: Int
    if length xs /= d then throwIO WrongListLength else
        S.unsafeWith (S.fromList xs) $ \px -> unsafeWithBatch o $ \po ->
            c_replicate h (code (Proxy :: Proxy e)) (castPtr px) po >>= checkRC ctx "t5_replicate"
    return o

-- | @pure a@ of the Applicative instance (src/Data/Tensor/Vector.hs:185-188) on a batch.
pureB :: forall is e. (SingI is, DeviceElt e) => Context -> Int -> e -> IO (Batch is e)
pureB ctx n a = replicateB ctx n (replicate (fromIntegral (product (fromShape (sing :: Shape is)))) a)

-- | @unit@ of the old @SquareMatrix@ class: a batch of identity matrices.
unitB :: forall i e. (SingI i, DeviceElt e, Num e) => Context -> Int -> IO (Batch '[i, i] e)
unitB ctx n = let k = fromIntegral (head (fromShape (sing :: Shape '[i, i]))) :: Int
              in replicateB ctx n [if r == c then 1 else 0 | r <- [1 .. k], c <- [1 .. k]]

scaleB :: forall is e. (SingI is, DeviceElt e) => e -> Batch is e -> IO (Batch is e)
scaleB k a = do
    let Context h = batchCtx a
    c <- newBatch (batchCtx a) (batchCount a)
    with k $ \pk -> unsafeWithBatch a $ \pa -> unsafeWithBatch c $ \pc ->
        c_scale h (code (Proxy :: Proxy e)) (castPtr pk) pa pc >>= checkRC (batchCtx a) "t5_scale"
    return c
