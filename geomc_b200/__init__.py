"""geomc_b200 -- B200-native batched small-matrix factorization / solve / inverse.

Drop-in for ONE hot path of trbabb/geomc (geomc/linalg: decomp_plu, decomp_lup, backsolve_lu,
backsolve_plu, linear_solve, cholesky, inv), executed for millions of independent systems per
call by hand-written sm_100a CUDA kernels behind a C ABI (include/geomc_b200.h).
There is no CPU fallback: importing works anywhere, computing needs the built library and a GPU.
"""
from . import _lib
from ._lib import LAYOUT_AOS, LAYOUT_SOA, GmbError, lib, set_option
from .linalg import (BatchSoA, apply_permutation, backsolve_lu, backsolve_plu, checksum, cholesky, cholesky_solve,
                     decomp_lup, decomp_plu, host_cholesky_solve, host_decomp_plu, host_inv, host_linear_solve, inv,
                     linear_solve, linear_solve_vec, mul, nullspace, orthogonal, simplex_contains, simplex_intersect,
                     solve_residual, trace_simplex, xf_apply, xf_compose, xf_from_matrix, xf_scale,
                     xf_translation)

__all__ = [
    "BatchSoA", "GmbError", "LAYOUT_AOS", "LAYOUT_SOA", "lib",
    "decomp_plu", "decomp_lup", "backsolve_lu", "backsolve_plu", "apply_permutation", "linear_solve",
    "linear_solve_vec", "cholesky", "cholesky_solve", "inv", "orthogonal", "nullspace", "solve_residual", "checksum",
    "simplex_contains", "trace_simplex", "simplex_intersect",
    "mul", "xf_from_matrix", "xf_translation", "xf_scale", "xf_compose", "xf_apply",
    "host_linear_solve", "host_inv", "host_cholesky_solve", "host_decomp_plu",
]
