// Poseidon2 over BabyBear, width 16, x^7, 4 + 13 + 4 rounds -- one permutation per thread, state in registers,
// Montgomery form throughout.
//
// Replaces the external p3_poseidon2::Poseidon2::permute_mut the reference calls from
// plonky2/src/hash/poseidon2_babybear.rs:69-84,150-159; the permutation is the one restated in-tree by
// plonky2/src/gates/poseidon2_babybear.rs:609-671 (external layer :804-832 with apply_mat4 :903-917, internal
// layer :787-802: multiply by 2^-32, then s0 <- sum_{i>=1} s_i - s0, s_i <- sum_all + 2^shift_i * s_i).
// In Montgomery form the 2^-32 factor is one Montgomery reduction of the 64-bit sum, as in p3-baby-bear.
#pragma once
#include "field_bb.cuh"
#include "hash_tables.h"

namespace p2g {

// Plain C++ (P2G_BB_FN, field_bb.cuh): tests/cpp/poseidon2_bb_emul_test.cpp runs this very code on the host against the
// oracle with every range assumption asserted.  The including translation unit defines the table c_p2bb
// (hash.cu: __constant__ Poseidon2BbTables c_p2bb).

// (s + c)^7 for a lane s < p + 2^15 and a round constant handed over as c - p (hash_tables.h): the sum is a signed word
// of magnitude <= p + 2^15 -- one add, no reduction -- and the four products of the chain are signed Montgomery products
// that need no correction between them (field_bb.cuh).  Result signed, in (-p, p).
P2G_BB_FN int p2bb_sbox_rc_signed(u32 s, u32 c_minus_p) {
    P2G_BB_CHECK(s < P2G_BB_P + 32768);
    const int x = (int)(s + c_minus_p);
    const int x2 = bb_mul_signed(x, x);
    const int x3 = bb_mul_signed(x2, x);
    const int x4 = bb_mul_signed(x2, x2);
    return bb_mul_signed(x3, x4);
}

P2G_BB_FN void p2bb_mat4(u32 &x0, u32 &x1, u32 &x2, u32 &x3) {
    u32 t01 = bb_add(x0, x1);
    u32 t23 = bb_add(x2, x3);
    u32 t0123 = bb_add(t01, t23);
    u32 t01123 = bb_add(t0123, x1);
    u32 t01233 = bb_add(t0123, x3);
    u32 n3 = bb_add(t01233, bb_add(x0, x0));
    u32 n1 = bb_add(t01123, bb_add(x2, x2));
    u32 n0 = bb_add(t01123, t01);
    u32 n2 = bb_add(t01233, t23);
    x0 = n0;
    x1 = n1;
    x2 = n2;
    x3 = n3;
}

P2G_BB_FN void p2bb_external(u32 s[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) p2bb_mat4(s[i], s[i + 1], s[i + 2], s[i + 3]);
    u32 sums[4];
#pragma unroll
    for (int k = 0; k < 4; k++) sums[k] = bb_add(bb_add(s[k], s[4 + k]), bb_add(s[8 + k], s[12 + k]));
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = bb_add(s[i], sums[i % 4]);
}

// The 13 internal rounds keep their lanes LAZY: a lane leaves a round as the Montgomery reduction of a sum below 2^47
// WITHOUT the conditional subtraction, i.e. in [0, p + 2^15) -- a range the next round's 64-bit sums and shifts take as
// it is (the sums stay below 2^47), so the loop holds no conditional subtraction: per lane one IMAD.WIDE (sum +
// lane * 2^shift), one IMAD and one IMAD.HI, all on the multiplier pipe.  Lane 0 goes through the signed S-box and comes
// back as x^7 + p in (0, 2p).
P2G_BB_FN void p2bb_internal(u32 s[16]) {
    const int shifts[15] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15};
    P2G_BB_CHECK(s[0] < 2 * P2G_BB_P);
    u64 part = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) {
        P2G_BB_CHECK(s[i] < P2G_BB_P + 32768);
        part += s[i];
    }
    const u64 full = part + s[0];
    const u32 n0 = bb_monty_reduce_lazy(part + (u64)(2 * P2G_BB_P - s[0]));
#pragma unroll
    for (int i = 1; i < 16; i++) s[i] = bb_monty_reduce_lazy(full + ((u64)s[i] << shifts[i - 1]));
    s[0] = n0;
}

P2G_BB_FN void poseidon2_bb_permute(u32 s[16]) {
    p2bb_external(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = bb_from_signed(p2bb_sbox_rc_signed(s[i], c_p2bb.ext_rc[r][i]));
        p2bb_external(s);
    }
#pragma unroll 1
    for (int r = 0; r < 13; r++) {
        s[0] = (u32)p2bb_sbox_rc_signed(s[0], c_p2bb.int_rc[r]) + P2G_BB_P;
        p2bb_internal(s);
    }
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = bb_min_u32(s[i], s[i] - P2G_BB_P);  // lazy lanes back into [0, p)
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = bb_from_signed(p2bb_sbox_rc_signed(s[i], c_p2bb.ext_rc[r][i]));
        p2bb_external(s);
    }
}

}  // namespace p2g
