// Constant tables of the two permutations (built on the host by poseidon_tables.cpp, uploaded by hash.cu).
#pragma once
#include <stdint.h>

namespace p2g {

struct PoseidonGlTables {
    unsigned long long rc0[12];     // ALL_ROUND_CONSTANTS[0..12] (poseidon_goldilocks.rs:114-212), added before round 0
    // constants of rounds 1..29 (and zeros for "round 30") split into the MDS layer's three limbs, bits 0-21 / 22-43 / 44-63:
    // rcl[r][k][i] = limb k of ALL_ROUND_CONSTANTS[12 r + i]; the MDS of round r-1 adds rcl[r]
    unsigned int rcl[31][3][12];
};

struct Poseidon2BbTables {
    uint32_t ext_rc[8][16];  // Montgomery form; hash/poseidon2_babybear.rs:26-61
    uint32_t int_rc[13];     // Montgomery form; hash/poseidon2_babybear.rs:63-67
};

}  // namespace p2g
