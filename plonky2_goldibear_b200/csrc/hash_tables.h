// Constant tables of the two permutations (built on the host by poseidon_tables.cpp, uploaded by hash.cu).
#pragma once
#include <stdint.h>

namespace p2g {

struct PoseidonGlTables {
    unsigned long long rc[31][12];  // rc[r] = ALL_ROUND_CONSTANTS[12r..12r+12] (poseidon_goldilocks.rs:114-212), rc[30] = 0
};

struct Poseidon2BbTables {
    uint32_t ext_rc[8][16];  // Montgomery form; hash/poseidon2_babybear.rs:26-61
    uint32_t int_rc[13];     // Montgomery form; hash/poseidon2_babybear.rs:63-67
};

}  // namespace p2g
