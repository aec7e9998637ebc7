// Constant tables of the two permutations (built on the host by poseidon_tables.cpp, uploaded by hash.cu).
#pragma once
#include <stdint.h>

#define P2G_PGL_LIMB_ROUNDS 22  // all partial rounds, 4..25
#define P2G_PGL_LIMB_BIAS (1u << 18)

namespace p2g {

struct PoseidonGlTables {
    unsigned long long rc0[12];     // ALL_ROUND_CONSTANTS[0..12] (poseidon_goldilocks.rs:114-212), added before round 0
    // constants of rounds 1..29 (and zeros for "round 30") split into the MDS layer's three limbs, bits 0-21 / 22-43 / 44-63:
    // rcl[r][k][i] = limb k of ALL_ROUND_CONSTANTS[12 r + i]; the MDS of round r-1 adds rcl[r]
    unsigned int rcl[31][3][12];
    // limb-domain partial rounds 4..25 (poseidon_gl.cuh, pgl_partial_rounds_limbs): rclb[k] = lane 0's constant after
    // the MDS of round 4 + k (round constants of lanes 1..11 pushed through the linear layers, bias 2^18 taken out
    // mod p), split into limbs with the bias added back on limb 0; rc_limb_exit[i] = what lanes 1..11 still owe when
    // they leave the loop (canonical; entry 0 unused)
    unsigned int rclb[P2G_PGL_LIMB_ROUNDS][4];
    unsigned long long rc_limb_exit[12];
};

struct Poseidon2BbTables {
    // Montgomery form MINUS p as a wrapped 32-bit word: lane + entry is then a signed word in [-p, p + 2^15), which is what
    // the S-box's signed multiplication chain takes (poseidon2_bb.cuh, p2bb_sbox_rc_signed)
    uint32_t ext_rc[8][16];  // hash/poseidon2_babybear.rs:26-61
    uint32_t int_rc[13];     // hash/poseidon2_babybear.rs:63-67
};

}  // namespace p2g
