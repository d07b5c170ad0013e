"""tensor_b200 — B200-native (sm_100a) batched fixed-dimension linear algebra behind
the tensor5/tensor class API.  The compute lives in hand-written CUDA kernels
(tensor_b200/csrc) reached through the C ABI of include/t5b200.h; this package is
the Python mirror of the Haskell-side modules described in INTEGRATION.md and is
what the tests and bench drive.  There is no CPU implementation in this package.
"""
from ._lib import AOS, SOA, F64, F32, I64, U8, LibraryMissing, partition  # noqa: F401
from .device import (Batch, Context, DeviceError, PinnedArray, TensorException,  # noqa: F401
                     Unsupported, WrongListLength)
from .linear_algebra import (char_poly, det, dot, dot_sum, inverse, matmul, matmul_host, matmul_is_exact, min_poly,  # noqa: F401
                             poly_eval, prod, pure, replicate, row_echelon_form, scale, solve, tensor_product, tr, transpose,
                             triangular_solve, unit, zip_with)
from . import host, structure, synthetic  # noqa: F401
from .structure import Nested  # noqa: F401
