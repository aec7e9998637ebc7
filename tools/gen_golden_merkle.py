#!/usr/bin/env python3
"""Extract Merkle-path golden vectors from the reference's frozen Goldilocks proof.

Run once in the build container (needs /root/reference and oracle/libp2o.so); output is committed as
tests/golden/gl_proof_merkle_paths.json so tests and the GPU box never read /root/reference.

Source bytes: plonky2/src/recursion/regression_test_data.rs (:5 common data, :62 verifier data, :93 proof),
the fixture of `test_recursive_verifier_gl_regression` (recursion/recursive_verifier.rs:280-314).
Wire format followed: plonky2/src/util/serialization/mod.rs (read_common_circuit_data :774-, read_proof :1023-,
read_fri_proof :575-, read_fri_query_rounds :524-, read_fri_initial_proof :434-, read_merkle_proof :406-).

Every FRI query round opens one leaf of each of the four PolynomialBatch oracles (constants+sigmas, wires,
Zs+partial products, quotient) plus one leaf per FRI commit-phase tree, each with its Merkle path to a cap
that the reference's own `from_values`/`MerkleTree::new` produced.  The leaf index is not stored in the proof
(the verifier derives it by Fiat-Shamir); it is recovered here by searching for the index whose path reaches
the cap -- all four initial oracles must agree on it, which is itself a strong check of the hashing semantics.
"""
import ctypes
import json
import os
import re
import struct
import sys

REF = os.environ.get("P2G_REFERENCE", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_arrays():
    src = open(os.path.join(REF, "plonky2/src/recursion/regression_test_data.rs")).read()
    out = {}
    for m in re.finditer(r"static (\w+): \[u8; (\d+)\] = \[(.*?)\];", src, re.S):
        data = bytes(int(x) for x in re.findall(r"\d+", m.group(3)))
        assert len(data) == int(m.group(2))
        out[m.group(1)] = data
    return out


class Rd:
    def __init__(self, b):
        self.b, self.o = b, 0

    def u8(self):
        v = self.b[self.o]
        self.o += 1
        return v

    def u32(self):
        v = struct.unpack_from("<I", self.b, self.o)[0]
        self.o += 4
        return v

    def usize(self):
        v = struct.unpack_from("<Q", self.b, self.o)[0]
        self.o += 8
        return v

    def usize_vec(self):
        return [self.usize() for _ in range(self.usize())]

    def field_vec(self, n):
        v = list(struct.unpack_from("<%dQ" % n, self.b, self.o))
        self.o += 8 * n
        return v

    def hash(self):
        return self.field_vec(4)

    def cap(self, h):
        return [self.hash() for _ in range(1 << h)]

    def merkle_proof(self):
        return [self.hash() for _ in range(self.u8())]


def read_fri_config(r):
    c = dict(rate_bits=r.usize(), cap_height=r.usize(), num_query_rounds=r.usize(), pow_bits=r.u32())
    variant = r.u8()
    if variant == 0:
        c["strategy"] = ("fixed", r.usize_vec())
    elif variant == 1:
        c["strategy"] = ("const_arity", r.usize(), r.usize())
    else:
        c["strategy"] = ("min_size", r.usize() if r.u8() else None)
    return c


def read_common(b):
    r = Rd(b)
    cfg = dict(num_wires=r.usize(), num_routed_wires=r.usize(), num_constants=r.usize(), security_bits=r.usize(),
               num_challenges=r.usize(), max_quotient_degree_factor=r.usize(), use_base_arithmetic_gate=r.u8(),
               zero_knowledge=r.u8())
    cfg["fri_config"] = read_fri_config(r)
    fri_params = dict(config=read_fri_config(r), reduction_arity_bits=r.usize_vec(), degree_bits=r.usize(), hiding=r.u8())
    sel = dict(selector_indices=r.usize_vec(), groups=[(r.usize(), r.usize()) for _ in range(r.usize())])
    cd = dict(config=cfg, fri_params=fri_params, selectors=sel, quotient_degree_factor=r.usize(),
              num_gate_constraints=r.usize(), num_constants=r.usize(), num_public_inputs=r.usize())
    cd["k_is"] = r.field_vec(r.usize())
    cd["num_partial_products"] = r.usize()
    cd["num_lookup_polys"] = r.usize()
    cd["num_lookup_selectors"] = r.usize()
    cd["num_luts"] = r.usize()
    return cd


def main():
    arrays = load_arrays()
    cd = read_common(arrays["RECURSIVE_VERIFIER_GL_COMMON_DATA"])
    cfg, fp = cd["config"], cd["fri_params"]
    assert cd["num_luts"] == 0 and cd["num_lookup_polys"] == 0
    D = 2
    salt = 4 if fp["hiding"] else 0
    cap_h = cfg["fri_config"]["cap_height"]
    print("config:", {k: v for k, v in cfg.items() if k != "fri_config"}, cfg["fri_config"], file=sys.stderr)
    print("fri_params:", fp, "k_is[:5]:", cd["k_is"][:5], file=sys.stderr)
    # k_is are powers of F::generator() (field/src/cosets.rs:8-20): pins the GL generator 7
    assert cd["k_is"][:4] == [1, 7, 49, 343]

    vr = Rd(arrays["RECURSIVE_VERIFIER_GL_VERIFIER_DATA"])
    cs_cap = vr.cap(vr.usize())
    circuit_digest = vr.hash()

    r = Rd(arrays["RECURSIVE_VERIFIER_GL_PROOF"])
    wires_cap, zs_cap, quot_cap = r.cap(cap_h), r.cap(cap_h), r.cap(cap_h)
    n_open = (cd["num_constants"] + cfg["num_routed_wires"] + cfg["num_wires"] + 2 * cfg["num_challenges"]
              + cd["num_partial_products"] * cfg["num_challenges"] + cd["quotient_degree_factor"] * cfg["num_challenges"])
    r.field_vec(n_open * D)  # OpeningSet, not needed here
    arity_bits = fp["reduction_arity_bits"]
    fri_caps = [r.cap(cap_h) for _ in arity_bits]
    rounds = []
    lens = [cd["num_constants"] + cfg["num_routed_wires"], cfg["num_wires"] + salt,
            cfg["num_challenges"] * (1 + cd["num_partial_products"] + cd["num_lookup_polys"]) + salt,
            cfg["num_challenges"] * cd["quotient_degree_factor"] + salt]
    for _ in range(cfg["fri_config"]["num_query_rounds"]):
        initial = [(r.field_vec(L), r.merkle_proof()) for L in lens]
        steps = [(r.field_vec((1 << ab) * D), r.merkle_proof()) for ab in arity_bits]
        rounds.append((initial, steps))
    final_poly_len = 1 << (fp["degree_bits"] - sum(arity_bits))
    r.field_vec(final_poly_len * D)
    r.field_vec(1)  # pow witness
    pis = r.field_vec(r.usize())
    assert r.o == len(r.b), (r.o, len(r.b))
    print("proof parsed ok; public inputs:", len(pis), file=sys.stderr)

    lib = ctypes.CDLL(os.path.join(ROOT, "oracle/libp2o.so"))
    lib.p2o_verify_merkle_proof_to_cap.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                                                   ctypes.c_void_p, ctypes.c_uint, ctypes.c_void_p, ctypes.c_uint]

    def arr(vals):
        return (ctypes.c_uint64 * len(vals))(*vals)

    def flat(hs):
        return [x for h in hs for x in h]

    lib.p2o_hash_or_noop.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.p2o_two_to_one.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]

    def find_index(leaf, proof, cap, candidates):
        l, s, c = arr(leaf), arr(flat(proof)), arr(flat(cap))
        return [idx for idx in candidates
                if lib.p2o_verify_merkle_proof_to_cap(0, l, len(leaf), idx, c, cap_h, s, len(proof))]

    def search_index(leaf, proof, cap):
        """All leaf indices whose path reaches the cap: walk the 2^len(proof) left/right choices level by level."""
        d = arr([0] * 4)
        lib.p2o_hash_or_noop(0, arr(leaf), len(leaf), d)
        frontier = [(0, list(d))]  # (low index bits, digest)
        for lvl, sib in enumerate(proof):
            nxt = []
            for bits, dig in frontier:
                for bit in (0, 1):
                    o = arr([0] * 4)
                    a, b = (arr(sib), arr(dig)) if bit else (arr(dig), arr(sib))
                    lib.p2o_two_to_one(0, a, b, o)
                    nxt.append((bits | (bit << lvl), list(o)))
            frontier = nxt
        return [bits | (ci << len(proof)) for bits, dig in frontier for ci, ch in enumerate(cap) if ch == dig]

    caps = [cs_cap, wires_cap, zs_cap, quot_cap]
    lde_bits = fp["degree_bits"] + cfg["fri_config"]["rate_bits"]
    out_paths = []
    for qi, (initial, steps) in enumerate(rounds):
        assert all(len(p) == lde_bits - cap_h for _, p in initial)
        hits = search_index(initial[1][0], initial[1][1], caps[1])
        assert len(hits) == 1, (qi, hits)
        x = hits[0]
        for t, (leaf, proof) in enumerate(initial):
            assert find_index(leaf, proof, caps[t], [x]) == [x], (qi, t)
            out_paths.append(dict(tree=["constants_sigmas", "wires", "zs_partial_products", "quotient"][t], query=qi,
                                  leaf_index=x, leaf=leaf, siblings=proof, cap_id=t))
        # FRI commit-phase trees: leaf = arity*D base elements, index = x >> arity_bits (fri/verifier.rs)
        xi = x
        for s, ((evals, proof), ab) in enumerate(zip(steps, arity_bits)):
            xi >>= ab
            assert find_index(evals, proof, fri_caps[s], [xi]) == [xi], (qi, s)
            out_paths.append(dict(tree="fri_layer_%d" % s, query=qi, leaf_index=xi, leaf=evals, siblings=proof,
                                  cap_id=4 + s))
    keep_queries = {0, 1, 2, 13, 27}
    doc = dict(
        source="plonky2/src/recursion/regression_test_data.rs:5,62,93 (golden proof of test_recursive_verifier_gl)",
        note="all %d query rounds verified against the oracle at generation time; queries %s kept" % (len(rounds), sorted(keep_queries)),
        field="goldilocks", cap_height=cap_h, degree_bits=fp["degree_bits"], rate_bits=cfg["fri_config"]["rate_bits"],
        hiding=bool(fp["hiding"]), k_is_prefix=cd["k_is"][:6],
        caps=[[[hex(x) for x in h] for h in c] for c in caps + fri_caps],
        paths=[dict(p, leaf=[hex(v) for v in p["leaf"]], siblings=[[hex(v) for v in h] for h in p["siblings"]])
               for p in out_paths if p["query"] in keep_queries],
    )
    with open(os.path.join(ROOT, "tests/golden/gl_proof_merkle_paths.json"), "w") as f:
        json.dump(doc, f)
    print("wrote %d paths (%d verified)" % (len(doc["paths"]), len(out_paths)))


if __name__ == "__main__":
    main()
