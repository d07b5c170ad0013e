{-# LANGUAGE DataKinds             #-}
{-# LANGUAGE FlexibleContexts      #-}
{-# LANGUAGE FlexibleInstances     #-}
{-# LANGUAGE KindSignatures        #-}
{-# LANGUAGE MultiParamTypeClasses #-}
{-# LANGUAGE PolyKinds             #-}
{-# LANGUAGE ScopedTypeVariables   #-}
{-# LANGUAGE TypeFamilies          #-}
{-# LANGUAGE TypeOperators         #-}
{-# LANGUAGE UndecidableInstances  #-}

--------------------------------------------------------------------------------
-- |
-- Module      :  Data.Tensor.LinearAlgebra
--
-- The linear-algebra classes the reference dropped in 0.2.0 (CHANGELOG.md:24-29:
-- @MatrixProduct@, @TensorProduct@; @SquareMatrix@ named at :28), re-introduced
-- with the semantics fixed in SURVEY.md §8a, and their instances for device
-- batches ("Data.Tensor.Device").
--
-- The class methods are PURE, like every operation of the reference
-- (@transpose@, src/Data/Indexable.hs:101-102; the Functor / Applicative
-- instances, src/Data/Tensor/Vector.hs:180-190): a device batch is an immutable
-- value, every operation returns a fresh buffer and never writes its operands
-- (include/t5b200.h, "Inputs are never written"), so wrapping the one @ccall@ of
-- "Data.Tensor.LinearAlgebra.Device" in 'unsafePerformIO' is referentially
-- transparent.  Each wrapper is NOINLINE (one allocation per evaluation, no
-- duplication by the simplifier) and uses 'unsafePerformIO', not
-- 'unsafeDupablePerformIO': a duplicated evaluation would launch the kernel twice.
-- Laziness is safe with respect to context lifetime: a thunk forced after
-- 'withContext' has returned finds a closed context, the library answers
-- T5_ERR_INVALID (never a crash - buffers outliving t5_shutdown are orphans,
-- t5b200.h), and the wrapper throws 'DeviceError'.
--
-- Result SHAPES are computed at the type level - @is ++ js@ for the tensor
-- product, the reference's own 'ReverseI' relation (src/Data/Indexable.hs:82-97)
-- for 'transpose' - through associated type families, so that a shape error is a
-- type error exactly as with "Data.Tensor.Vector".
--
-- NOT TYPE-CHECKED here (no GHC in the image, SURVEY.md §0 F2).
--------------------------------------------------------------------------------

module Data.Tensor.LinearAlgebra
    ( -- * Classes (pure)
      MatrixProduct(..), DotProduct(..), TensorProduct(..), Product(..)
    , SquareMatrix(..), LinearSystem(..)
    , transposeB'
      -- * Type-level shape arithmetic
    , type (++), Drop, Take
      -- * Pure forms of the remaining operations
    , dotSum', tr', polyEval', charPoly', minPoly', rowEchelonForm', triangularSolve'
    , productIsExact
    , (.+.), (.-.), (.*), hadamard
      -- * The IO layer, re-exported for callers that sequence device work themselves
    , module Data.Tensor.LinearAlgebra.Device
    ) where

import           Data.Proxy
import           Data.Singletons
import           Data.Word                        (Word8)
import           System.IO.Unsafe                 (unsafePerformIO)

import           Data.Indexable                   (ReverseI)
import           Data.MultiIndex                  (PI (..), fromPI)
import           Data.Tensor.Device
import           Data.Tensor.Device.FFI            (c_matmul_is_exact)
import           Data.Tensor.LinearAlgebra.Device

infixl 7 .*.
infixl 7 ⊗
infixl 6 .+., .-.
infixl 7 .*

-- | List concatenation on index lists: the shape of a tensor product.
type family (is :: [PI]) ++ (js :: [PI]) :: [PI] where
    '[]       ++ js = js
    (i ': is) ++ js = i ': (is ++ js)

-- | @Take n is@ / @Drop n is@ with a unary count: the shape arithmetic of @prod n@.
type family Drop (n :: PI) (is :: [PI]) :: [PI] where
    Drop 'One      (i ': is) = is
    Drop ('S n)    (i ': is) = Drop n is
type family Take (n :: PI) (is :: [PI]) :: [PI] where
    Take 'One      (i ': is) = '[i]
    Take ('S n)    (i ': is) = i ': Take n is
-- | All but the last @n@ indices.
type family DropLast (n :: PI) (is :: [PI]) :: [PI] where
    DropLast n is = RevList (Drop n (RevList is))
type family RevList (is :: [PI]) :: [PI] where
    RevList is = RevOnto is '[]
type family RevOnto (is :: [PI]) (acc :: [PI]) :: [PI] where
    RevOnto '[]       acc = acc
    RevOnto (i ': is) acc = RevOnto is (i ': acc)

-- | @a .*. b@ contracts the last index of @a@ with the first of @b@.
class MatrixProduct a b where
    type MatrixProductSpace a b
    (.*.) :: a -> b -> MatrixProductSpace a b

-- | One scalar per pair: @fold_{i ascending} (acc + x_i * y_i)@, @acc0 = 0@.
class DotProduct a where
    type DotSpace a
    dot :: a -> a -> DotSpace a

-- | @a ⊗ b@: the rank-(r+s) outer product, one multiply per component.
class TensorProduct a b where
    type a :⊗: b
    (⊗) :: a -> b -> a :⊗: b

-- | @prod n a b@: contract the last @n@ indices of @a@ with the first @n@ of @b@ (the count is a
-- type-level 'PI' passed by proxy, so the result shape is computed).
class Product (n :: PI) a b where
    type ProdSpace n a b
    prod :: proxy n -> a -> b -> ProdSpace n a b

-- | @det@ and @inverse@; the per-item 'Maybe' of @inverse@ is the status batch (0 = Just, 1 = Nothing).
class SquareMatrix m where
    type Scalars m
    type Statuses m
    det     :: m -> Scalars m
    inverse :: m -> (m, Statuses m)

-- | @solve a b@: unique-solution square systems (first-non-zero pivoting, SURVEY §8a).
class LinearSystem m v where
    type SolStatuses m v
    solve :: m -> v -> (v, SolStatuses m v)

-- Instances for device batches ---------------------------------------------------------------

instance (SingI i, SingI j, SingI k, DeviceElt e) => MatrixProduct (Batch '[i, j] e) (Batch '[j, k] e) where
    type MatrixProductSpace (Batch '[i, j] e) (Batch '[j, k] e) = Batch '[i, k] e
    a .*. b = pure2 matmulB a b

instance (SingI i, SingI j, DeviceElt e) => MatrixProduct (Batch '[i, j] e) (Batch '[j] e) where
    type MatrixProductSpace (Batch '[i, j] e) (Batch '[j] e) = Batch '[i] e
    a .*. x = pure2 matvecB a x

instance (SingI i, DeviceElt e) => DotProduct (Batch '[i] e) where
    type DotSpace (Batch '[i] e) = Batch '[] e
    dot x y = pure2 dotB x y

instance (SingI is, SingI js, SingI (is ++ js), DeviceElt e) => TensorProduct (Batch is e) (Batch js e) where
    type Batch is e :⊗: Batch js e = Batch (is ++ js) e
    a ⊗ b = pure2 tensorProductB a b

instance (SingI is, SingI js, SingI (DropLast n is ++ Drop n js), SingI n, DeviceElt e)
      => Product n (Batch is e) (Batch js e) where
    type ProdSpace n (Batch is e) (Batch js e) = Batch (DropLast n is ++ Drop n js) e
    prod _ a b = pure2 (prodN (piToInt (sing :: Sing n))) a b

instance (SingI i, DeviceElt e, Fractional e) => SquareMatrix (Batch '[i, i] e) where
    type Scalars  (Batch '[i, i] e) = Batch '[] e
    type Statuses (Batch '[i, i] e) = Batch '[] Word8
    det a     = pure1 detB a
    inverse a = pure1 inverseB a

instance (SingI i, DeviceElt e, Fractional e) => LinearSystem (Batch '[i, i] e) (Batch '[i] e) where
    type SolStatuses (Batch '[i, i] e) (Batch '[i] e) = Batch '[] Word8
    solve a b = pure2 solveB a b

-- | 'Data.Indexable.transpose' for device batches, with the reference's own constraint
-- (@ReverseI is '[] js@, src/Data/Indexable.hs:101) computing the result shape.  It is a separate
-- function rather than a 'MultiIndexable' instance because a batch has no per-element 'Indexable'
-- access from Haskell (its elements live in HBM); on a batch it is one gather kernel.
transposeB' :: forall is js e. (ReverseI is '[] js, SingI is, SingI js, DeviceElt e) => Batch is e -> Batch js e
transposeB' a = pure1 transposeB a

-- Pure forms of the remaining operations --------------------------------------------------------

-- | Sum over the batch of 'dot' - the only operation with a cross-GPU exchange (SURVEY.md §8e).
dotSum' :: (SingI i, DeviceElt e) => Batch '[i] e -> Batch '[i] e -> e
dotSum' x y = pure2 dotSum x y

-- | Is @(.*.)@ of an @m x k@ by a @k x n@ matrix of this element type the bit-exact sequential fold ('True'), or
-- does the library run it on the tensor cores within BASELINE.json's tolerance (Double 1e-12, Float 1e-5 of
-- @sum_j |a_ij||b_jc|@; include/t5b200.h, t5_matmul)?  Depends on element type and shape only.  The C entry is a
-- pure function of its arguments (no context, no device), hence the plain 'unsafePerformIO'.
productIsExact :: forall e. DeviceElt e => Proxy e -> Int -> Int -> Int -> Bool
productIsExact pe m k n = unsafePerformIO $ do
    r <- c_matmul_is_exact (dtypeCode pe) (fromIntegral m) (fromIntegral k) (fromIntegral n)
    return (r /= 0)        -- (negative = rejected arguments: (.*.) itself throws for those)
{-# NOINLINE productIsExact #-}

tr' :: (SingI i, DeviceElt e) => Batch '[i, i] e -> Batch '[] e
tr' a = pure1 tr a

polyEval' :: (SingI i, DeviceElt e) => [e] -> Batch '[i, i] e -> Batch '[i, i] e
polyEval' cs a = pure1 (polyEval cs) a

charPoly' :: (SingI i, DeviceElt e, Fractional e) => Batch '[i, i] e -> Batch '[ 'S i ] e
charPoly' a = pure1 charPoly a

minPoly' :: (SingI i, DeviceElt e, Fractional e) => Batch '[i, i] e -> (Batch '[ 'S i ] e, Batch '[] Word8)
minPoly' a = pure1 minPoly a

rowEchelonForm' :: (SingI i, SingI j, DeviceElt e, Fractional e) => Batch '[i, j] e -> (Batch '[i, j] e, Batch '[] Word8)
rowEchelonForm' a = pure1 rowEchelonForm a

-- | @triangularSolve' k ab@: reduce the first @k@ columns of the augmented batch @[a | b]@.
triangularSolve' :: (SingI i, SingI j, DeviceElt e, Fractional e)
                 => Int -> Batch '[i, j] e -> (Batch '[i, j] e, Batch '[] Word8)
triangularSolve' k a = pure1 (echelonWith k) a

-- | The vector-space operations (@liftA2 (+)@, @liftA2 (-)@, @fmap (k *)@ of
-- src/Data/Tensor/Vector.hs:180-190, :250-251) on batches.
(.+.), (.-.), hadamard :: (SingI is, DeviceElt e) => Batch is e -> Batch is e -> Batch is e
a .+. b = pure2 addB a b
a .-. b = pure2 subB a b
hadamard a b = pure2 hadamardB a b

(.*) :: (SingI is, DeviceElt e) => e -> Batch is e -> Batch is e
k .* a = pure1 (scaleB k) a

-- Purification ----------------------------------------------------------------------------------

pure1 :: (a -> IO r) -> a -> r
pure1 f a = unsafePerformIO (f a)
{-# NOINLINE pure1 #-}

pure2 :: (a -> b -> IO r) -> a -> b -> r
pure2 f a b = unsafePerformIO (f a b)
{-# NOINLINE pure2 #-}

-- | The run-time value of a type-level positive integer ('fromPI', src/Data/MultiIndex.hs:64-66).
piToInt :: Sing (n :: PI) -> Int
piToInt = fromPI
