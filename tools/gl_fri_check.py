"""Goldilocks FRI verifier pieces needed to check the reference's frozen proof against the ORACLE's field constants and
permutation: the Fiat-Shamir transcript (plonk/get_challenges.rs:28-93, iop/challenger.rs:32-160, fri/challenges.rs:12-62)
and the per-query consistency checks (fri/verifier.rs:19-46,131-255).  Pure Python big-int arithmetic; the only outside
ingredients are passed in: `permute` (Poseidon width 12), `generator`, `two_adic_generator(bits)`.

What a passing run pins, with bytes the reference itself produced:
  * the transcript reproduces the 28 query indices that the Merkle paths were found to open  -> sponge/duplex semantics;
  * row k of each oracle is consumed as the evaluation at generator * w_N^bitrev(k) (verifier.rs:198-199) and the
    folding checks only close if that is the point the prover's from_values put there
    -> leaf order, coset shift g = 7 and the two-adic root chain (N = 2^17 and the arity-16 sub-roots) are the reference's.
"""
P = 0xFFFFFFFF00000001
W = 7  # Goldilocks quadratic extension: x^2 = 7 (p3 BinomialExtensionField<Goldilocks, 2>)


def e_add(a, b):
    return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)


def e_sub(a, b):
    return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)


def e_mul(a, b):
    return ((a[0] * b[0] + W * a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)


def e_inv(a):
    norm = (a[0] * a[0] - W * a[1] * a[1]) % P
    ni = pow(norm, P - 2, P)
    return (a[0] * ni % P, (-a[1]) * ni % P)


def e_pow(a, e):
    r = (1, 0)
    while e:
        if e & 1:
            r = e_mul(r, a)
        a = e_mul(a, a)
        e >>= 1
    return r


def base(x):
    return (x % P, 0)


def reverse_bits(x, n):
    r = 0
    for i in range(n):
        r |= ((x >> i) & 1) << (n - 1 - i)
    return r


class Challenger:
    """iop/challenger.rs: overwrite-mode duplex sponge, rate 8, challenges popped from the END of the output buffer."""

    def __init__(self, permute):
        self.permute = permute
        self.state = [0] * 12
        self.inp = []
        self.out = []

    def observe(self, x):
        self.out = []
        self.inp.append(x % P)
        if len(self.inp) == 8:
            self.duplex()

    def observe_many(self, xs):
        for x in xs:
            self.observe(x)

    def observe_ext(self, e):
        self.observe(e[0])
        self.observe(e[1])

    def duplex(self):
        for i, x in enumerate(self.inp):
            self.state[i] = x
        self.inp = []
        self.state = list(self.permute(self.state))
        self.out = list(self.state[:8])

    def challenge(self):
        if self.inp or not self.out:
            self.duplex()
        return self.out.pop()

    def ext_challenge(self):
        a = self.challenge()
        b = self.challenge()
        return (a, b)


def reduce_with(alpha, values):
    """ReducingFactor::reduce: sum_i values[i] * alpha^i; returns (result, count)"""
    acc = (0, 0)
    for v in reversed(values):
        acc = e_add(e_mul(acc, alpha), v)
    return acc, len(values)


def fri_fold_chain(x_index, old_eval, step_evals, arity_bits, fri_betas, final_poly, lde_bits, generator, two_adic_generator):
    """The folding half of fri_verifier_query_round (verifier.rs:209-255): starting from the value `old_eval` the first
    layer must hold at leaf position x_index, check every layer's opened coset against the previous fold
    (compute_evaluation, verifier.rs:19-46) and the last fold against the final polynomial."""
    subgroup_x = generator * pow(two_adic_generator(lde_bits), reverse_bits(x_index, lde_bits), P) % P
    for i, ab in enumerate(arity_bits):
        arity = 1 << ab
        evals = step_evals[i]
        coset_index, within = x_index >> ab, x_index & (arity - 1)
        assert evals[within] == old_eval, "FRI layer %d: committed value differs from the folded one" % i
        # compute_evaluation (verifier.rs:19-46)
        g = two_adic_generator(ab)
        ordered = [evals[reverse_bits(j, ab)] for j in range(arity)]
        start = subgroup_x * pow(g, arity - reverse_bits(within, ab), P) % P
        xs = [base(start * pow(g, j, P) % P) for j in range(arity)]
        beta = fri_betas[i]
        lx = (1, 0)
        for xj in xs:
            lx = e_mul(lx, e_sub(beta, xj))
        acc = (0, 0)
        for a in range(arity):
            wa = (1, 0)
            for b in range(arity):
                if a != b:
                    wa = e_mul(wa, e_sub(xs[a], xs[b]))
            term = e_mul(e_inv(e_mul(wa, e_sub(beta, xs[a]))), ordered[a])
            acc = e_add(acc, term)
        old_eval = e_mul(lx, acc)
        subgroup_x = pow(subgroup_x, arity, P)
        x_index = coset_index
    fin = (0, 0)
    for cf in reversed(final_poly):
        fin = e_add(e_mul(fin, base(subgroup_x)), cf)
    assert fin == old_eval, "final polynomial evaluation is invalid"


def verify(fx, permute, generator, two_adic_generator):
    """fx: the fixture dict (tests/golden/gl_proof_fri.json).  Raises AssertionError on any failed check; returns a summary."""
    H = lambda xs: [int(x, 16) for x in xs]
    cfg = fx["config"]
    num_ch, degree_bits, rate_bits, cap_h = cfg["num_challenges"], cfg["degree_bits"], cfg["rate_bits"], cfg["cap_height"]
    arity_bits = cfg["reduction_arity_bits"]
    lde_bits = degree_bits + rate_bits
    # ---- transcript (get_challenges.rs:28-93)
    ch = Challenger(permute)
    ch.observe_many(H(fx["circuit_digest"]))
    ch.observe_many([0, 0, 0, 0])  # hash_no_pad of zero public inputs: the untouched zero state (hashing.rs:100-123)
    assert fx["num_public_inputs"] == 0
    for h in fx["wires_cap"]:
        ch.observe_many(H(h))
    betas = [ch.challenge() for _ in range(num_ch)]
    gammas = [ch.challenge() for _ in range(num_ch)]
    for h in fx["zs_cap"]:
        ch.observe_many(H(h))
    alphas = [ch.challenge() for _ in range(num_ch)]
    for h in fx["quotient_cap"]:
        ch.observe_many(H(h))
    zeta = ch.ext_challenge()
    op = {k: [tuple(H(e)) for e in v] for k, v in fx["openings"].items()}
    zeta_batch = op["constants"] + op["plonk_sigmas"] + op["wires"] + op["plonk_zs"] + op["partial_products"] + op["quotient_polys"]
    next_batch = op["plonk_zs_next"]
    for e in zeta_batch + next_batch:
        ch.observe_ext(e)
    fri_alpha = ch.ext_challenge()
    fri_betas = []
    for cap in fx["fri_caps"]:
        for h in cap:
            ch.observe_many(H(h))
        fri_betas.append(ch.ext_challenge())
    final_poly = [tuple(H(e)) for e in fx["final_poly"]]
    for e in final_poly:
        ch.observe_ext(e)
    pre_pow = (list(ch.state), list(ch.inp))  # what fri_proof_of_work starts from (prover.rs:166-168)
    ch.observe(int(fx["pow_witness"], 16))
    pow_response = ch.challenge()
    assert pow_response < (1 << (64 - cfg["proof_of_work_bits"])), "proof of work (verifier.rs:48-60)"
    indices = [ch.challenge() % (1 << lde_bits) for _ in range(cfg["num_query_rounds"])]
    assert indices == fx["query_indices_from_merkle_paths"], "Fiat-Shamir query indices differ from the opened Merkle paths"

    # ---- per query (fri_verifier_query_round, verifier.rs:161-255)
    w_lde = two_adic_generator(lde_bits)
    g_deg = two_adic_generator(degree_bits)
    points = [zeta, e_mul(base(g_deg), zeta)]
    reduced_openings = [reduce_with(fri_alpha, zeta_batch)[0], reduce_with(fri_alpha, next_batch)[0]]
    salt = 4 if cfg["hiding"] else 0
    n_cs = cfg["num_constants"] + cfg["num_routed_wires"]
    n_zs_pp = num_ch * (1 + cfg["num_partial_products"])
    n_q = num_ch * cfg["quotient_degree_factor"]
    checked = 0
    for q in fx["queries"]:
        x_index = q["x_index"]
        ev = [H(e) for e in q["initial_evals"]]
        unsalted = [ev[0], ev[1][:len(ev[1]) - salt], ev[2][:len(ev[2]) - salt], ev[3][:len(ev[3]) - salt]]
        assert len(unsalted[0]) == n_cs and len(unsalted[1]) == cfg["num_wires"] and len(unsalted[2]) == n_zs_pp and len(unsalted[3]) == n_q
        subgroup_x = generator * pow(w_lde, reverse_bits(x_index, lde_bits), P) % P
        # fri_combine_initial (verifier.rs:106-159)
        batch_evals = [[base(v) for v in unsalted[0] + unsalted[1] + unsalted[2] + unsalted[3]],
                       [base(v) for v in unsalted[2][:num_ch]]]
        total, count = (0, 0), 0
        for evals, red_open, point in zip(batch_evals, reduced_openings, points):
            red, cnt = reduce_with(fri_alpha, evals)
            numerator = e_sub(red, red_open)
            denominator = e_sub(base(subgroup_x), point)
            total = e_mul(total, e_pow(fri_alpha, cnt))  # alpha.shift(sum): the count of the reduce just done
            total = e_add(total, e_mul(numerator, e_inv(denominator)))
        fri_fold_chain(x_index, total, [[tuple(H(e)) for e in step] for step in q["steps"]], arity_bits, fri_betas, final_poly,
                       lde_bits, generator, two_adic_generator)
        checked += 1
    return {"queries_checked": checked, "query_indices": indices, "zeta": zeta, "pre_pow_state": pre_pow[0], "pre_pow_input": pre_pow[1],
            "pow_witness": int(fx["pow_witness"], 16), "proof_of_work_bits": cfg["proof_of_work_bits"]}
