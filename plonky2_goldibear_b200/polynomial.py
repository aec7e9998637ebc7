"""Host mirror of the batched polynomial transforms the prover runs around the commit
(field/src/polynomial/mod.rs, field/src/fft.rs), on column-major arrays [num_polys][n], through p2g_transform."""
import numpy as np

from . import _lib
from .fields import Field, REPR_CANONICAL

_IFFT, _COSET_LDE, _COSET_IFFT = 0, 1, 2


def _run(field: Field, kind: int, cols, rate_bits: int, ctx, repr: int):
    ctx = ctx or _lib.default_context()
    cols = np.ascontiguousarray(np.asarray(cols, dtype=field.dtype))
    if cols.ndim != 2 or cols.shape[0] == 0:
        raise ValueError("expected a non-empty [num_polys][n] array")
    n = cols.shape[1]
    if n <= 0 or n & (n - 1):
        raise ValueError("Not a power of two: %d" % n)  # log2_strict
    out = np.empty((cols.shape[0], n << rate_bits), dtype=field.dtype)
    ctx.check(ctx.lib.p2g_transform(ctx.handle, field.field_id, repr, kind, n.bit_length() - 1, cols.shape[0], rate_bits, 0,
                                    cols.ctypes.data, out.ctypes.data))
    return out


def ifft(field: Field, values, ctx=None, repr: int = REPR_CANONICAL):
    """PolynomialValues::ifft per row (polynomial/mod.rs:55-57, fft.rs:66-94)."""
    return _run(field, _IFFT, values, 0, ctx, repr)


def coset_ifft(field: Field, values, ctx=None, repr: int = REPR_CANONICAL):
    """PolynomialValues::coset_ifft(F::generator()) per row (polynomial/mod.rs:62-72): what compute_quotient_polys ends
    with (plonk/prover.rs:921-925)."""
    return _run(field, _COSET_IFFT, values, 0, ctx, repr)


def coset_lde(field: Field, coeffs, rate_bits: int, ctx=None, repr: int = REPR_CANONICAL):
    """PolynomialCoeffs::lde(rate_bits).coset_fft(F::generator()) per row, natural order (oracle.rs:125-150)."""
    return _run(field, _COSET_LDE, coeffs, rate_bits, ctx, repr)
