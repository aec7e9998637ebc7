"""Shared helpers for the tests: the deterministic synthetic input of SURVEY.md 8(d)."""
import numpy as np

GL_P = 0xFFFFFFFF00000001
BB_P = 2013265921


def splitmix64(x: np.ndarray) -> np.ndarray:
    x = (x + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = x
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def synthetic_columns(order: int, dtype, num_cols: int, n: int, seed: int = 42, col_begin: int = 0) -> np.ndarray:
    """value[col][row] = SplitMix64(seed ^ (col << 32 | row)) mod p, for columns col_begin .. col_begin + num_cols - 1"""
    with np.errstate(over="ignore"):
        col = np.arange(col_begin, col_begin + num_cols, dtype=np.uint64)[:, None] << np.uint64(32)
        row = np.arange(n, dtype=np.uint64)[None, :]
        v = splitmix64(np.uint64(seed) ^ (col | row))
    return (v % np.uint64(order)).astype(dtype)
