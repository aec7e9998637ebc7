"""Poseidon2-BabyBear pinned by the reference's SECOND in-tree definition: the gate's constraint system.

The reference computes the permutation with the external p3_poseidon2 crate (hash/poseidon2_babybear.rs:69-84,150-159);
no literal known-answer vector for it exists in the tree.  What the tree does hold twice is the algorithm:
  (1) the witness generator Poseidon2BabyBearGenerator::run_once  gates/poseidon2_babybear.rs:573-680 with the plain
      linear layers :787-832, :903-917 -- this is what oracle/p2o.cpp (namespace p2bb) follows;
  (2) the constraint system Poseidon2BabyBearGate::eval_unfiltered :204-309, and the recursive-circuit form
      eval_unfiltered_circuit :403-506 with its own linear layers sbox_circuit :707-717, add_rc_circuit :719-734,
      permute_internal_mut_circuit :742-785, permute_external_mut_circuit :834-862, apply_mat4_circuit :864-901, whose
      comments also spell the 4x4 block out as a matrix (:893-900).
This file restates (2) in plain Python integers (nothing shared with the oracle or the CUDA code) and checks that all
150 constraints vanish on the oracle's round-by-round states, for swap = 0 and 1, that they do NOT vanish when a wire,
a shift, the Montgomery factor or a round constant is perturbed, and that the constraint system run forward (every
`state - wire` constraint read as the wire's definition) reproduces the oracle's permutation output.  The reference's
own tests tie (1) to p3: gates/poseidon2_babybear.rs:957-1004 (generated witness == p3 permutation) -- so (1) == (2)
on our side closes the chain up to the p3 arithmetic itself.
"""
import os
import re

import numpy as np
import pytest

P = 2013265921
W = 16
CAP = 8
HALF_FULL, FULL, PARTIAL = 4, 8, 13
MONTY_INV = 943718400  # gates/poseidon2_babybear.rs:776
DIAG_SHIFTS = [0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15]  # :41-42

REF = "/root/reference/plonky2/src/hash/poseidon2_babybear.rs"


def _constants():
    """Round constants: parsed from the reference when it is present (this container), else from the header the
    oracle is built from (generated from the same literals by tools/gen_constants.py)."""
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "oracle", "p2o_constants.h")).read()
    ext = [int(x) for x in re.findall(r"\d+", hdr.split("P2O_BB_P2_EXTERNAL_RC")[1].split("=", 1)[1].split(";")[0])]
    internal = [int(x) for x in re.findall(r"\d+", hdr.split("P2O_BB_P2_INTERNAL_RC")[1].split("=", 1)[1].split(";")[0])]
    ext = [ext[16 * r:16 * r + 16] for r in range(8)]
    assert len(internal) == 13 and all(len(r) == 16 for r in ext)
    if os.path.exists(REF):
        src = open(REF).read()
        e2 = [int(x) for x in re.findall(r"\d+", src.split("EXTERNAL_CONSTANTS:")[1].split("=", 1)[1].split("];")[0])]
        i2 = [int(x) for x in re.findall(r"\d+", src.split("INTERNAL_CONSTANTS:")[1].split("=", 1)[1].split("];")[0])]
        assert [e2[16 * r:16 * r + 16] for r in range(8)] == ext and i2 == internal
    return ext, internal


# ---- the circuit-flavoured layers (the *_circuit functions), on integers mod p ----------------------------------------
def sbox_circuit(x):  # :707-717
    x2 = x * x % P
    x3 = x * x2 % P
    x4 = x2 * x2 % P
    return x3 * x4 % P


def apply_mat4_circuit(x):  # :864-901 (the non-gate branch)
    t01 = (x[0] + x[1]) % P
    t23 = (x[2] + x[3]) % P
    t0123 = (t01 + t23) % P
    t01123 = (t0123 + x[1]) % P
    t01233 = (t0123 + x[3]) % P
    x3 = (2 * x[0] + t01233) % P  # mul_const_add_extension(F::two(), x[0], t01233)
    x1 = (2 * x[2] + t01123) % P
    x0 = (t01123 + t01) % P
    x2 = (t01233 + t23) % P
    return [x0, x1, x2, x3]


def permute_external_circuit(state):  # :834-862
    s = list(state)
    for i in range(0, W, 4):
        s[i:i + 4] = apply_mat4_circuit(s[i:i + 4])
    sums = [sum(s[j + k] for j in range(0, W, 4)) % P for k in range(4)]
    return [(s[i] + sums[i % 4]) % P for i in range(W)]


def permute_internal_circuit(state, shifts=DIAG_SHIFTS, monty_inv=MONTY_INV):  # :772-784
    s = [monty_inv * x % P for x in state]
    part_sum = sum(s[1:]) % P
    full_sum = (part_sum + s[0]) % P
    out = [(part_sum - s[0]) % P]
    for i in range(W - 1):
        out.append(((1 << shifts[i]) * s[i + 1] + full_sum) % P)  # mul_const_add_extension(shift, state[i+1], full_sum)
    return out


# ---- wire indices, num_ops = 1 (:79-147) ------------------------------------------------------------------------------
ROUTED = 2 * W + 1
wire_input = lambda i: i
wire_output = lambda i: W + i
WIRE_SWAP = 2 * W
START_DELTA = ROUTED
START_FULL_0 = START_DELTA + CAP
wire_full_sbox_0 = lambda r, i: START_FULL_0 + W * (r - 1) + i
START_PARTIAL = START_FULL_0 + W * (HALF_FULL - 1)
START_FULL_1 = START_PARTIAL + PARTIAL
wire_full_sbox_1 = lambda r, i: START_FULL_1 + W * r + i
NUM_WIRES = START_FULL_1 + W * HALF_FULL
NUM_CONSTRAINTS = W * (FULL - 1) + PARTIAL + W + 1 + CAP  # :540-547


def eval_unfiltered(wires, ext_rc, int_rc, define_wires=False, **layer_kw):
    """Constraint values of one operation (:204-309).  With define_wires the `state - wire` constraints are read the other
    way round -- the wire takes the computed value -- which turns the constraint system into a forward evaluation."""
    w = [int(x) for x in wires]
    cons = []
    swap = w[WIRE_SWAP]
    cons.append(swap * (swap - 1) % P)
    for i in range(CAP):
        lhs, rhs = w[wire_input(i)], w[wire_input(i + CAP)]
        if define_wires:
            w[START_DELTA + i] = swap * (rhs - lhs) % P
        cons.append((swap * (rhs - lhs) - w[START_DELTA + i]) % P)
    state = [1] * W
    for i in range(CAP):
        d = w[START_DELTA + i]
        state[i] = (w[wire_input(i)] + d) % P
        state[i + CAP] = (w[wire_input(i + CAP)] - d) % P
    state = permute_external_circuit(state)

    def take(idx, value):
        if define_wires:
            w[idx] = value
        cons.append((value - w[idx]) % P)
        return w[idx]

    for r in range(HALF_FULL):
        state = [(state[i] + ext_rc[r][i]) % P for i in range(W)]
        if r > 0:
            state = [take(wire_full_sbox_0(r, i), state[i]) for i in range(W)]
        state = permute_external_circuit([sbox_circuit(x) for x in state])
    for r in range(PARTIAL):
        state[0] = (state[0] + int_rc[r]) % P
        state[0] = sbox_circuit(take(START_PARTIAL + r, state[0]))
        state = permute_internal_circuit(state, **layer_kw)
    for r in range(HALF_FULL, FULL):
        state = [(state[i] + ext_rc[r][i]) % P for i in range(W)]
        state = [take(wire_full_sbox_1(r - HALF_FULL, i), state[i]) for i in range(W)]
        state = permute_external_circuit([sbox_circuit(x) for x in state])
    for i in range(W):
        take(wire_output(i), state[i])
    assert len(cons) == NUM_CONSTRAINTS
    return cons, w


def _inputs():
    rng = np.random.default_rng(2024)
    yield np.arange(16, dtype=np.uint32)  # the input of the reference's generated_output test, :960-1004
    yield np.zeros(16, dtype=np.uint32)
    yield np.full(16, P - 1, dtype=np.uint32)
    for _ in range(40):
        yield rng.integers(0, P, 16, dtype=np.uint64).astype(np.uint32)


def test_wire_layout_matches_the_gate():
    assert NUM_WIRES == 166 and NUM_CONSTRAINTS == 150
    assert (WIRE_SWAP, START_DELTA, START_FULL_0, START_PARTIAL, START_FULL_1) == (32, 33, 41, 89, 102)


def test_mat4_is_the_matrix_the_reference_comments_spell_out():
    """:893-900: rows (2,3,1,1), (1,2,3,1), (1,1,2,3), (3,1,1,2); SURVEY 8(a) a14 quotes the same block."""
    m = [[2, 3, 1, 1], [1, 2, 3, 1], [1, 1, 2, 3], [3, 1, 1, 2]]
    rng = np.random.default_rng(1)
    for _ in range(20):
        x = [int(v) for v in rng.integers(0, P, 4)]
        assert apply_mat4_circuit(x) == [sum(m[r][c] * x[c] for c in range(4)) % P for r in range(4)]


def test_gate_constraints_vanish_on_the_oracle_witness(oracle):
    ext_rc, int_rc = _constants()
    for state in _inputs():
        for swap in (0, 1):
            wires = oracle.poseidon2_bb_gate_wires(state, swap)
            assert wires.shape == (NUM_WIRES,)
            cons, _ = eval_unfiltered(wires, ext_rc, int_rc)
            assert not any(cons), "constraint %d does not vanish" % next(i for i, c in enumerate(cons) if c)
            # the routed output wires are the permutation of the (possibly swapped) input
            src = np.concatenate([state[8:], state[:8]]) if swap else state
            assert (wires[16:32] == oracle.permute(oracle.BB, src)).all()


def test_gate_constraints_run_forward_reproduce_the_oracle_permutation(oracle):
    ext_rc, int_rc = _constants()
    for state in _inputs():
        wires = np.zeros(NUM_WIRES, dtype=np.uint64)
        wires[:16] = state
        _, w = eval_unfiltered(wires, ext_rc, int_rc, define_wires=True)
        assert w[16:32] == [int(x) for x in oracle.permute(oracle.BB, state)]


def test_gate_constraints_reject_perturbed_definitions(oracle):
    """Negative controls: the constraint system is not vacuous."""
    ext_rc, int_rc = _constants()
    state = np.arange(16, dtype=np.uint32)
    wires = oracle.poseidon2_bb_gate_wires(state, 0)
    for idx in (16, 41, 89, 101, 102, 165):
        bad = wires.copy()
        bad[idx] = (int(bad[idx]) + 1) % P
        assert any(eval_unfiltered(bad, ext_rc, int_rc)[0])
    shifted = list(DIAG_SHIFTS)
    shifted[-1] = 14  # the real table skips 14
    assert any(eval_unfiltered(wires, ext_rc, int_rc, shifts=shifted)[0])
    assert any(eval_unfiltered(wires, ext_rc, int_rc, monty_inv=1)[0])
    rc2 = [list(r) for r in ext_rc]
    rc2[5][3] = (rc2[5][3] + 1) % P
    assert any(eval_unfiltered(wires, rc2, int_rc)[0])
    int2 = list(int_rc)
    int2[12] = (int2[12] + 1) % P
    assert any(eval_unfiltered(wires, ext_rc, int2)[0])


@pytest.mark.gpu
def test_gpu_permutation_satisfies_the_gate_constraints(oracle, ctx):
    """The CUDA permutation's outputs as the gate's output wires, interior wires from the constraint system run forward:
    every constraint vanishes, i.e. the kernel computes what the reference's gate proves."""
    import plonky2_goldibear_b200 as p2g
    ext_rc, int_rc = _constants()
    states = np.ascontiguousarray(np.stack(list(_inputs())), dtype=np.uint32)
    out = states.copy()
    ctx.check(ctx.lib.p2g_permute(ctx.handle, p2g.BABYBEAR.hash_id, 0, out.ctypes.data, len(out)))
    for s_in, s_out in zip(states, out):
        wires = np.zeros(NUM_WIRES, dtype=np.uint64)
        wires[:16] = s_in
        _, w = eval_unfiltered(wires, ext_rc, int_rc, define_wires=True)
        w[16:32] = [int(x) for x in s_out]
        assert not any(eval_unfiltered(w, ext_rc, int_rc)[0])
