#!/usr/bin/env python3
"""Extract the FRI-verification fixture from the reference's frozen Goldilocks proof and verify it with the oracle.

Run once in the build container (needs /root/reference and oracle/libp2o.so); writes tests/golden/gl_proof_fri.json.
Source bytes and wire format: see tools/gen_golden_merkle.py (same proof, plonky2/src/recursion/regression_test_data.rs)."""
import ctypes
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import gen_golden_merkle as gm  # noqa: E402
import gl_fri_check  # noqa: E402


def main():
    arrays = gm.load_arrays()
    cd = gm.read_common(arrays["RECURSIVE_VERIFIER_GL_COMMON_DATA"])
    cfg, fp = cd["config"], cd["fri_params"]
    D = 2
    salt = 4 if fp["hiding"] else 0
    cap_h = cfg["fri_config"]["cap_height"]
    vr = gm.Rd(arrays["RECURSIVE_VERIFIER_GL_VERIFIER_DATA"])
    cs_cap = vr.cap(vr.usize())
    circuit_digest = vr.hash()
    r = gm.Rd(arrays["RECURSIVE_VERIFIER_GL_PROOF"])
    wires_cap, zs_cap, quot_cap = r.cap(cap_h), r.cap(cap_h), r.cap(cap_h)
    nch = cfg["num_challenges"]

    def ext_vec(n):
        v = r.field_vec(n * D)
        return [v[2 * i:2 * i + 2] for i in range(n)]

    # read_opening_set, serialization/mod.rs:343-375 (no lookups in this circuit)
    openings = dict(constants=ext_vec(cd["num_constants"]), plonk_sigmas=ext_vec(cfg["num_routed_wires"]),
                    wires=ext_vec(cfg["num_wires"]), plonk_zs=ext_vec(nch), plonk_zs_next=ext_vec(nch),
                    partial_products=ext_vec(cd["num_partial_products"] * nch),
                    quotient_polys=ext_vec(cd["quotient_degree_factor"] * nch))
    arity_bits = fp["reduction_arity_bits"]
    fri_caps = [r.cap(cap_h) for _ in arity_bits]
    lens = [cd["num_constants"] + cfg["num_routed_wires"], cfg["num_wires"] + salt,
            nch * (1 + cd["num_partial_products"]) + salt, nch * cd["quotient_degree_factor"] + salt]
    rounds = []
    for _ in range(cfg["fri_config"]["num_query_rounds"]):
        initial = [(r.field_vec(L), r.merkle_proof()) for L in lens]
        steps = [(ext_vec(1 << ab), r.merkle_proof()) for ab in arity_bits]
        rounds.append((initial, steps))
    final_poly = ext_vec(1 << (fp["degree_bits"] - sum(arity_bits)))
    pow_witness = r.field_vec(1)[0]
    pis = r.field_vec(r.usize())
    assert r.o == len(r.b)

    # leaf indices recovered from the Merkle paths (independent of the transcript)
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle/libp2o.so"))
    lib.p2o_hash_or_noop.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]
    lib.p2o_two_to_one.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p]
    lib.p2o_two_adic_generator.restype = ctypes.c_uint64
    lib.p2o_two_adic_generator.argtypes = [ctypes.c_int, ctypes.c_uint]
    lib.p2o_generator.restype = ctypes.c_uint64

    def arr(v):
        return (ctypes.c_uint64 * len(v))(*v)

    def search_index(leaf, proof, cap):
        d = arr([0] * 4)
        lib.p2o_hash_or_noop(0, arr(leaf), len(leaf), d)
        frontier = [(0, list(d))]
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

    indices = []
    for initial, _ in rounds:
        hits = search_index(initial[1][0], initial[1][1], wires_cap)
        assert len(hits) == 1
        indices.append(hits[0])

    hx = lambda v: [hex(x) for x in v]
    keep = [0, 1, 5, 13, 20, 27]
    fx = dict(
        source="plonky2/src/recursion/regression_test_data.rs:5,62,93 (golden proof of test_recursive_verifier_gl)",
        note="transcript inputs of the whole proof + %d of the %d FRI query rounds (all %d verified at generation time)" % (
            len(keep), len(rounds), len(rounds)),
        config=dict(num_challenges=nch, degree_bits=fp["degree_bits"], rate_bits=cfg["fri_config"]["rate_bits"], cap_height=cap_h,
                    reduction_arity_bits=arity_bits, hiding=bool(fp["hiding"]), num_constants=cd["num_constants"],
                    num_routed_wires=cfg["num_routed_wires"], num_wires=cfg["num_wires"],
                    num_partial_products=cd["num_partial_products"], quotient_degree_factor=cd["quotient_degree_factor"],
                    proof_of_work_bits=cfg["fri_config"]["pow_bits"], num_query_rounds=cfg["fri_config"]["num_query_rounds"]),
        circuit_digest=hx(circuit_digest), num_public_inputs=len(pis),
        wires_cap=[hx(h) for h in wires_cap], zs_cap=[hx(h) for h in zs_cap], quotient_cap=[hx(h) for h in quot_cap],
        openings={k: [hx(e) for e in v] for k, v in openings.items()},
        fri_caps=[[hx(h) for h in c] for c in fri_caps], final_poly=[hx(e) for e in final_poly], pow_witness=hex(pow_witness),
        query_indices_from_merkle_paths=indices,
        queries=[dict(x_index=indices[i], initial_evals=[hx(ev) for ev, _ in rounds[i][0]],
                      steps=[[hx(e) for e in ev] for ev, _ in rounds[i][1]]) for i in range(len(rounds))],
    )

    def permute(state):
        s = arr(list(state))
        lib.p2o_poseidon_gl(s)
        return list(s)

    res = gl_fri_check.verify(fx, permute, int(lib.p2o_generator(0)), lambda b: int(lib.p2o_two_adic_generator(0, b)))
    print("verified %d query rounds; zeta = %s" % (res["queries_checked"], res["zeta"]))
    fx["queries"] = [fx["queries"][i] for i in keep]
    with open(os.path.join(ROOT, "tests/golden/gl_proof_fri.json"), "w") as f:
        json.dump(fx, f)
    print("wrote tests/golden/gl_proof_fri.json")


if __name__ == "__main__":
    main()
