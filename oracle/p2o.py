"""ctypes binding of the CPU oracle (oracle/libp2o.so, see oracle/p2o.h).  TEST INFRASTRUCTURE ONLY."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
GL, BB = 0, 1
VALUES, COEFFS = 0, 1
DTYPE = {GL: np.uint64, BB: np.uint32}
DIGEST_ELEMS = {GL: 4, BB: 8}
ORDER = {GL: 0xFFFFFFFF00000001, BB: 2013265921}


class CommitDesc(ctypes.Structure):
    _fields_ = [
        ("field", ctypes.c_int), ("input_kind", ctypes.c_int), ("log_n", ctypes.c_uint), ("num_cols", ctypes.c_uint),
        ("rate_bits", ctypes.c_uint), ("cap_height", ctypes.c_uint), ("num_salt_cols", ctypes.c_uint),
        ("threads", ctypes.c_int), ("input", ctypes.c_void_p), ("salts", ctypes.c_void_p),
        ("coeffs_out", ctypes.c_void_p), ("leaves_out", ctypes.c_void_p), ("digests_out", ctypes.c_void_p),
        ("cap_out", ctypes.c_void_p), ("stage_seconds", ctypes.c_double * 4),
    ]


def build(native: bool = False) -> str:
    """Compile the oracle with the system g++ (portable x86-64-v3 build, or -march=native)."""
    target = "native" if native else "all"
    subprocess.run(["make", "-C", _HERE, target], check=True, stdout=subprocess.DEVNULL)
    return os.path.join(_HERE, "libp2o_native.so" if native else "libp2o.so")


_LIBS = {}


def load(native: bool = False) -> ctypes.CDLL:
    if native in _LIBS:
        return _LIBS[native]
    path = os.path.join(_HERE, "libp2o_native.so" if native else "libp2o.so")
    if not os.path.exists(path):
        path = build(native)
    lib = ctypes.CDLL(path)
    vp, u64, sz, ui, ci = ctypes.c_void_p, ctypes.c_uint64, ctypes.c_size_t, ctypes.c_uint, ctypes.c_int
    for name in ("p2o_order", "p2o_generator"):
        getattr(lib, name).argtypes = [ci]
        getattr(lib, name).restype = u64
    lib.p2o_two_adic_generator.argtypes = [ci, ui]
    lib.p2o_two_adic_generator.restype = u64
    lib.p2o_mul.argtypes = [ci, u64, u64]
    lib.p2o_mul.restype = u64
    lib.p2o_pow.argtypes = [ci, u64, u64]
    lib.p2o_pow.restype = u64
    lib.p2o_inv.argtypes = [ci, u64]
    lib.p2o_inv.restype = u64
    for name in ("p2o_poseidon_gl", "p2o_poseidon_gl_naive", "p2o_poseidon2_bb"):
        getattr(lib, name).argtypes = [vp]
        getattr(lib, name).restype = None
    lib.p2o_poseidon2_bb_gate_wires.argtypes = [vp, ui, vp]
    lib.p2o_poseidon2_bb_gate_wires.restype = None
    lib.p2o_fft.argtypes = [ci, vp, ui, ui]
    lib.p2o_ifft.argtypes = [ci, vp, ui]
    lib.p2o_coset_fft.argtypes = [ci, vp, ui, u64, ui]
    lib.p2o_coset_ifft.argtypes = [ci, vp, ui, u64]
    lib.p2o_reverse_index_bits.argtypes = [vp, sz, ui]
    lib.p2o_reverse_bits.argtypes = [sz, ui]
    lib.p2o_reverse_bits.restype = sz
    lib.p2o_hash_or_noop.argtypes = [ci, vp, sz, vp]
    lib.p2o_hash_no_pad.argtypes = [ci, vp, sz, vp]
    lib.p2o_two_to_one.argtypes = [ci, vp, vp, vp]
    lib.p2o_merkle_tree.argtypes = [ci, vp, sz, sz, ui, vp, vp, ci]
    lib.p2o_merkle_tree.restype = ci
    lib.p2o_merkle_prove.argtypes = [vp, sz, ui, sz, vp]
    lib.p2o_verify_merkle_proof_to_cap.argtypes = [ci, vp, sz, sz, vp, ui, vp, ui]
    lib.p2o_verify_merkle_proof_to_cap.restype = ci
    lib.p2o_commit.argtypes = [ctypes.POINTER(CommitDesc)]
    lib.p2o_commit.restype = ci
    lib.p2o_max_threads.restype = ci
    lib.p2o_ext_degree.argtypes = [ci]
    lib.p2o_ext_degree.restype = ui
    lib.p2o_ext_w.argtypes = [ci]
    lib.p2o_ext_w.restype = u64
    lib.p2o_ext_mul.argtypes = [ci, vp, vp, vp]
    lib.p2o_fri_pow.argtypes = [ci, vp, vp, ui, ui, ctypes.POINTER(u64)]
    lib.p2o_fri_pow.restype = ci
    lib.p2o_fri_layer_leaves.argtypes = [ci, vp, ui, vp]
    lib.p2o_fri_fold_layer.argtypes = [ci, vp, ui, ui, vp, u64, vp, vp, ctypes.POINTER(u64)]
    _LIBS[native] = lib
    return lib


def _p(a):
    return a.ctypes.data if a is not None and a.size else None


def permute(field: int, state: np.ndarray, naive: bool = False) -> np.ndarray:
    lib = load()
    s = np.ascontiguousarray(state, dtype=DTYPE[field]).copy()
    if field == GL:
        (lib.p2o_poseidon_gl_naive if naive else lib.p2o_poseidon_gl)(s.ctypes.data)
    else:
        lib.p2o_poseidon2_bb(s.ctypes.data)
    return s


def poseidon2_bb_gate_wires(state: np.ndarray, swap: int = 0) -> np.ndarray:
    """The 166 wires of one Poseidon2BabyBearGate operation on `state` (gates/poseidon2_babybear.rs:79-147, :573-680)."""
    lib = load()
    s = np.ascontiguousarray(state, dtype=np.uint32)
    wires = np.zeros(166, dtype=np.uint32)
    lib.p2o_poseidon2_bb_gate_wires(s.ctypes.data, int(swap), wires.ctypes.data)
    return wires


def fft(field, col, zero_factor=0):
    lib = load()
    a = np.ascontiguousarray(col, dtype=DTYPE[field]).copy()
    lib.p2o_fft(field, a.ctypes.data, int(np.log2(a.size)), zero_factor)
    return a


def ifft(field, col):
    lib = load()
    a = np.ascontiguousarray(col, dtype=DTYPE[field]).copy()
    lib.p2o_ifft(field, a.ctypes.data, int(np.log2(a.size)))
    return a


def coset_fft(field, col, shift, zero_factor=0):
    lib = load()
    a = np.ascontiguousarray(col, dtype=DTYPE[field]).copy()
    lib.p2o_coset_fft(field, a.ctypes.data, int(np.log2(a.size)), shift, zero_factor)
    return a


def coset_ifft(field, col, shift):
    a = np.ascontiguousarray(col, dtype=DTYPE[field]).copy()
    load().p2o_coset_ifft(field, a.ctypes.data, a.size.bit_length() - 1, int(shift))
    return a


def reverse_index_bits(a):
    lib = load()
    a = np.ascontiguousarray(a).copy()
    lib.p2o_reverse_index_bits(a.ctypes.data, a.dtype.itemsize, int(np.log2(a.size)))
    return a


def hash_or_noop(field, elems):
    lib = load()
    e = np.ascontiguousarray(elems, dtype=DTYPE[field])
    out = np.zeros(DIGEST_ELEMS[field], dtype=DTYPE[field])
    lib.p2o_hash_or_noop(field, _p(e), e.size, out.ctypes.data)
    return out


def two_to_one(field, l, r):
    lib = load()
    l = np.ascontiguousarray(l, dtype=DTYPE[field])
    r = np.ascontiguousarray(r, dtype=DTYPE[field])
    out = np.zeros(DIGEST_ELEMS[field], dtype=DTYPE[field])
    lib.p2o_two_to_one(field, l.ctypes.data, r.ctypes.data, out.ctypes.data)
    return out


def merkle_tree(field, leaves, cap_height, threads=0):
    """returns (digests [2(N-2^h)][D], cap [2^h][D]); raises ValueError where the reference panics."""
    lib = load()
    leaves = np.ascontiguousarray(leaves, dtype=DTYPE[field])
    n, ln = leaves.shape
    d = DIGEST_ELEMS[field]
    if n == 0 or n & (n - 1) or cap_height > n.bit_length() - 1:
        raise ValueError("reference panics: bad leaf count / cap_height")
    digests = np.zeros((2 * (n - (1 << cap_height)), d), dtype=DTYPE[field])
    cap = np.zeros((1 << cap_height, d), dtype=DTYPE[field])
    rc = lib.p2o_merkle_tree(field, _p(leaves), n, ln, cap_height, _p(digests), cap.ctypes.data, threads)
    if rc != 0:
        raise ValueError("p2o_merkle_tree failed")
    return digests, cap


def merkle_prove(field, digests, num_leaves, cap_height, leaf_index):
    lib = load()
    ns = (num_leaves.bit_length() - 1) - cap_height
    out = np.zeros((ns, DIGEST_ELEMS[field]), dtype=DTYPE[field])
    digests = np.ascontiguousarray(digests)
    lib.p2o_merkle_prove(_p(digests), num_leaves, cap_height, leaf_index, _p(out))
    return out


def verify_merkle_proof_to_cap(field, leaf, leaf_index, cap, siblings) -> bool:
    lib = load()
    leaf = np.ascontiguousarray(leaf, dtype=DTYPE[field])
    cap = np.ascontiguousarray(cap, dtype=DTYPE[field])
    sib = np.ascontiguousarray(siblings, dtype=DTYPE[field])
    cap_height = (cap.shape[0]).bit_length() - 1
    return bool(lib.p2o_verify_merkle_proof_to_cap(field, _p(leaf), leaf.size, leaf_index, cap.ctypes.data, cap_height,
                                                   _p(sib), sib.shape[0] if sib.size else 0))


def commit(field, columns, input_kind, rate_bits, cap_height, salts=None, threads=0, want_leaves=True, native=False):
    """PolynomialBatch::from_values / from_coeffs on the CPU.  `columns`: [c][n] array.  Returns a dict."""
    lib = load(native)
    cols = np.ascontiguousarray(columns, dtype=DTYPE[field])
    c, n = cols.shape
    log_n = n.bit_length() - 1
    N = n << rate_bits
    d = DIGEST_ELEMS[field]
    salt_arr = None if salts is None else np.ascontiguousarray(salts, dtype=DTYPE[field])
    ns = 0 if salt_arr is None else salt_arr.shape[0]
    coeffs = np.zeros((c, n), dtype=DTYPE[field])
    leaves = np.zeros((N, c + ns), dtype=DTYPE[field]) if want_leaves else None
    digests = np.zeros((2 * (N - (1 << cap_height)), d), dtype=DTYPE[field])
    cap = np.zeros((1 << cap_height, d), dtype=DTYPE[field])
    desc = CommitDesc()
    desc.field, desc.input_kind, desc.log_n, desc.num_cols = field, input_kind, log_n, c
    desc.rate_bits, desc.cap_height, desc.num_salt_cols, desc.threads = rate_bits, cap_height, ns, threads
    desc.input, desc.salts = _p(cols), _p(salt_arr)
    desc.coeffs_out, desc.leaves_out = _p(coeffs), _p(leaves)
    desc.digests_out, desc.cap_out = _p(digests), cap.ctypes.data
    rc = lib.p2o_commit(ctypes.byref(desc))
    if rc != 0:
        raise ValueError("reference panics on these arguments")
    return dict(coeffs=coeffs, leaves=leaves, digests=digests, cap=cap, stage_seconds=list(desc.stage_seconds))


def max_threads() -> int:
    return load().p2o_max_threads()


# ---- FRI commit phase (plonky2/src/fri/prover.rs:83-133) -------------------------------------------------------------
def ext_degree(field) -> int:
    return int(load().p2o_ext_degree(field))


def ext_mul(field, a, b) -> np.ndarray:
    a = np.ascontiguousarray(a, dtype=DTYPE[field])
    b = np.ascontiguousarray(b, dtype=DTYPE[field])
    out = np.empty_like(a)
    load().p2o_ext_mul(field, a.ctypes.data, b.ctypes.data, out.ctypes.data)
    return out


def fri_layer_leaves(field, values, arity_bits) -> np.ndarray:
    """values [n][D] -> leaf matrix [n >> arity_bits][D << arity_bits] (:103-109)"""
    values = np.ascontiguousarray(values, dtype=DTYPE[field])
    n, d = values.shape
    out = np.empty_like(values)
    load().p2o_fri_layer_leaves(field, values.ctypes.data, n.bit_length() - 1, out.ctypes.data)
    return out.reshape(n >> arity_bits, d << arity_bits)


def fri_fold_layer(field, coeffs, arity_bits, beta, shift):
    """(:114-125) -> (folded coeffs [n >> a][D], their coset evaluations [n >> a][D] in natural order, shift^arity)"""
    coeffs = np.ascontiguousarray(coeffs, dtype=DTYPE[field])
    beta = np.ascontiguousarray(beta, dtype=DTYPE[field])
    n, d = coeffs.shape
    co = np.empty((n >> arity_bits, d), dtype=DTYPE[field])
    vo = np.empty((n >> arity_bits, d), dtype=DTYPE[field])
    so = ctypes.c_uint64()
    load().p2o_fri_fold_layer(field, coeffs.ctypes.data, n.bit_length() - 1, arity_bits, beta.ctypes.data, int(shift),
                              co.ctypes.data, vo.ctypes.data, ctypes.byref(so))
    return co, vo, int(so.value)


def fri_committed_trees(field, coeffs, values, challenger, reduction_arity_bits, rate_bits, cap_height):
    """fri_committed_trees (:83-133) composed from the pieces above; `challenger` is any object with observe_cap /
    get_extension_challenge / observe_extension_elements (tests pass the package's Challenger with this module's
    permutation injected).  Returns ([(leaves, digests, cap)], final_coeffs)."""
    trees = []
    shift = int(load().p2o_generator(field))
    coeffs = np.ascontiguousarray(coeffs, dtype=DTYPE[field])
    values = np.ascontiguousarray(values, dtype=DTYPE[field])
    for arity_bits in reduction_arity_bits:
        leaves = fri_layer_leaves(field, values, arity_bits)
        digests, cap = merkle_tree(field, leaves, cap_height)
        trees.append((leaves, digests, cap))
        challenger.observe_cap(cap)
        beta = challenger.get_extension_challenge()
        coeffs, values, shift = fri_fold_layer(field, coeffs, arity_bits, beta, shift)
    final = coeffs[: coeffs.shape[0] >> rate_bits].copy()
    challenger.observe_extension_elements(final)
    return trees, final


def fri_pow(field, sponge_state, input_buffer, min_leading_zeros) -> int:
    """fri_proof_of_work's search (:147-179): the smallest valid witness."""
    st = np.ascontiguousarray(sponge_state, dtype=DTYPE[field])
    inp = np.ascontiguousarray(input_buffer, dtype=DTYPE[field])
    w = ctypes.c_uint64()
    rc = load().p2o_fri_pow(field, st.ctypes.data, _p(inp), inp.size, min_leading_zeros, ctypes.byref(w))
    if rc != 0:
        raise ValueError("Proof of work failed. This is highly unlikely!")
    return int(w.value)
