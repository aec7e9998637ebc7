"""Executable model of the NEXT NTT design: a multi-pass coset LDE with decimation-in-time
butterflies whose twiddles absorb the coset factor, checked against the oracle's leaf order.

Why: today's LDE pass 1 multiplies every element by its coset factor s^j on load and by an inter-pass factor on store
(2 of its 7 multiplications per element) and stores through a bit-reversed row index.  With t = b*tw first and then
(a + t, a - t) on NATURAL-order input, level l of the size-n transform splits block i (modulus X^(2h) - rho) into
X^h -+ sqrt(rho) with  sqrt(rho) = s^h * w_{2^(l+1)}^{bitrev_l(i)},  h = n / 2^(l+1):  one twiddle per (level, block),
the coset factor inside it, the result at position q = P(s * w_n^{bitrev(q)}) -- the reference's leaf order
(fri/oracle.rs:108-109 + util reverse_index_bits) -- and nothing to do on load or store.
A pass over levels [L0, L0 + lp) works inside chunks (= level-L0 blocks, index q, natural position order); its twiddle
for local level m and local block i' factors as
    B[m][i'] * U_m(q),   B[m][i'] = s^(C / 2^(m+1)) * w_{2^(m+1)}^{bitrev_m(i')}   (C = chunk length; per pass and coset)
                         U_m(q)   = w_{2^(L0+m+1)}^{bitrev_L0(q)}                    (per chunk: lp values, by squaring)
so the first pass (one chunk, U = 1) uses B as it is and a later pass builds its 2^lp - 1 twiddles per CTA with one
multiplication each (shared by the W columns of the tile).
This file is the index arithmetic of that kernel in plain Python (small sizes); tests/test_oracle.py runs it."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def bitrev(x, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (x & 1)
        x >>= 1
    return r


def pass_base_table(p, s, w_of, chunk_log, lp):
    """B[(1 << m) + i'] for m < lp, i' < 2^m  (heap order, entry 0 unused); w_of(k) = primitive 2^k-th root."""
    table = [0] * (1 << lp)
    for m in range(lp):
        sm = pow(s, 1 << (chunk_log - m - 1), p)
        w = w_of(m + 1)
        for i in range(1 << m):
            table[(1 << m) + i] = sm * pow(w, bitrev(i, m), p) % p
    return table


def chunk_factors(p, w_of, l0, lp, q):
    """U_m(q), m < lp."""
    g = pow(w_of(l0 + lp), bitrev(q, l0), p)  # U_{lp-1}
    u = [0] * lp
    for m in range(lp - 1, -1, -1):
        u[m] = g
        g = g * g % p
    return u


def coset_lde_ct(p, s, w_of, coeffs, lps):
    """Size-n coset transform of `coeffs` (natural order) over passes of lps[i] levels each; returns the list whose entry q
    is P(s * w_n^{bitrev_n(q)}).  Data stays in place between passes (no permutation anywhere)."""
    n = len(coeffs)
    log_n = n.bit_length() - 1
    assert sum(lps) == log_n
    x = list(coeffs)
    l0 = 0
    for lp in lps:
        chunk_log = log_n - l0
        base = pass_base_table(p, s, w_of, chunk_log, lp)
        for q in range(1 << l0):  # one chunk = one CTA's (strided or contiguous) tile family
            u = chunk_factors(p, w_of, l0, lp, q) if l0 else [1] * lp
            tw = [0] * (1 << lp)
            for m in range(lp):
                for i in range(1 << m):
                    tw[(1 << m) + i] = base[(1 << m) + i] * u[m] % p
            lo = q << chunk_log
            for m in range(lp):
                half = 1 << (chunk_log - m - 1)
                for i in range(1 << m):
                    t_w = tw[(1 << m) + i]
                    b0 = lo + i * 2 * half
                    for j in range(half):
                        a, b = x[b0 + j], x[b0 + j + half]
                        t = b * t_w % p
                        x[b0 + j], x[b0 + j + half] = (a + t) % p, (a - t) % p
        l0 += lp
    return x


def check(field_id, log_n, lps, rate_bits=2, seed=1):
    """all 2^rate_bits coset blocks of the LDE against the oracle's from_coeffs leaves"""
    import numpy as np
    from oracle import p2o
    lib = p2o.load()
    p = int(lib.p2o_order(field_id))
    g = int(lib.p2o_generator(field_id))
    w_of = lambda k: int(lib.p2o_two_adic_generator(field_id, k)) if k else 1
    n = 1 << log_n
    rng = np.random.default_rng(seed)
    coeffs = rng.integers(0, p, size=(1, n), dtype=np.uint64).astype(p2o.DTYPE[field_id])
    ref = p2o.commit(field_id, coeffs, p2o.COEFFS, rate_bits, 0)["leaves"][:, 0]
    w_big = w_of(log_n + rate_bits)
    for b in range(1 << rate_bits):
        s = g * pow(w_big, bitrev(b, rate_bits), p) % p  # block b of the leaves is the coset with this shift
        got = coset_lde_ct(p, s, w_of, [int(v) for v in coeffs[0]], lps)
        assert got == [int(v) for v in ref[b * n:(b + 1) * n]], (field_id, log_n, lps, b)
    return True


if __name__ == "__main__":
    for fid in (0, 1):
        for log_n, lps in ((6, [6]), (6, [3, 3]), (7, [4, 3]), (8, [3, 3, 2]), (5, [1, 2, 2])):
            check(fid, log_n, lps)
            print("field", fid, "2^%d" % log_n, lps, "ok")
