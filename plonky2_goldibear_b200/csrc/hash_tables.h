// Constant tables of the two permutations (built on the host by poseidon_tables.cpp, uploaded by hash.cu).
#pragma once
#include <stdint.h>

namespace p2g {

struct PoseidonGlTables {
    unsigned long long pre_add[2][12];    // round constants of the first full round of each half (rounds 0 and 26)
    unsigned long long mds_add[8][12];    // constants folded into the MDS of full round i: those of the NEXT layer
                                          // (rounds 1,2,3, the partial rounds' first constants, rounds 27,28,29, zero);
                                          // all from ALL_ROUND_CONSTANTS, poseidon_goldilocks.rs:114-212
    unsigned long long k[22];             // added to s0 after its S-box in partial round r
    unsigned long long row[22][11];       // d = 25*s0 + sum_i row[r][i] * s[i+1]
    unsigned long long col[22][11];       // s[i+1] += s0 * col[r][i]
    unsigned long long init[11][11];      // dense block applied to s[1..] before the first partial round
};

struct Poseidon2BbTables {
    uint32_t ext_rc[8][16];  // Montgomery form; hash/poseidon2_babybear.rs:26-61
    uint32_t int_rc[13];     // Montgomery form; hash/poseidon2_babybear.rs:63-67
};

}  // namespace p2g
