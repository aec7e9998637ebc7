// Host run of the CUDA kernels' Poseidon2-BabyBear code (csrc/field_bb.cuh, csrc/poseidon2_bb.cuh) against the oracle.
// The permutation keeps its lanes in lazy / signed ranges between corrections (signed Montgomery products in the S-box,
// uncorrected reductions in the internal rounds); here every range assumption the code states is asserted
// (P2G_BB_CHECK) on random and extreme states, and the outputs are compared word for word with p2o_poseidon2_bb.
// Test infrastructure: links the oracle; nothing here is product code.
#include <cstdio>
#include <cstdlib>
#include <cstring>

#define P2G_BB_CHECK(cond)                                                              \
    do {                                                                                \
        if (!(cond)) {                                                                  \
            std::fprintf(stderr, "range check failed: %s (%s:%d)\n", #cond, __FILE__, __LINE__); \
            std::abort();                                                               \
        }                                                                               \
    } while (0)

#include "../../plonky2_goldibear_b200/csrc/field_bb.cuh"
#include "../../plonky2_goldibear_b200/csrc/hash_tables.h"
namespace p2g {
Poseidon2BbTables c_p2bb;
}
#include "../../plonky2_goldibear_b200/csrc/poseidon2_bb.cuh"
#include "../../plonky2_goldibear_b200/csrc/poseidon_tables.cpp"
#include "p2o.h"

using namespace p2g;

static unsigned long long rng_state = 0x9e3779b97f4a7c15ULL;
static unsigned long long rnd() {
    rng_state ^= rng_state << 13;
    rng_state ^= rng_state >> 7;
    rng_state ^= rng_state << 17;
    return rng_state;
}

static int check_state(const uint32_t canon[16], const char *what) {
    uint32_t want[16], s[16];
    std::memcpy(want, canon, sizeof want);
    p2o_poseidon2_bb(want);
    for (int i = 0; i < 16; i++) s[i] = BbHost::to_internal(canon[i]);
    poseidon2_bb_permute(s);
    for (int i = 0; i < 16; i++) {
        if (s[i] >= P2G_BB_P) {
            std::printf("FAIL %s: lane %d not canonical\n", what, i);
            return 1;
        }
        const uint32_t got = BbHost::mul(s[i], P2G_BB_RINV);
        if (got != want[i]) {
            std::printf("FAIL %s: lane %d got %u want %u\n", what, i, got, want[i]);
            return 1;
        }
    }
    return 0;
}

int main() {
    build_poseidon2_bb_tables(&c_p2bb);
    int bad = 0;
    uint32_t st[16];
    // extreme states: constants, one-hot, alternating extremes
    const uint32_t ext[] = {0u, 1u, 2u, P2G_BB_P - 1, P2G_BB_P - 2, (P2G_BB_P - 1) / 2, (P2G_BB_P + 1) / 2, 0x7fffffffu % P2G_BB_P};
    for (uint32_t v : ext) {
        for (int i = 0; i < 16; i++) st[i] = v;
        bad += check_state(st, "constant state");
        for (int hot = 0; hot < 16; hot++) {
            for (int i = 0; i < 16; i++) st[i] = i == hot ? v : 0;
            bad += check_state(st, "one-hot state");
            for (int i = 0; i < 16; i++) st[i] = i == hot ? v : P2G_BB_P - 1;
            bad += check_state(st, "one-cold state");
        }
    }
    for (int i = 0; i < 16; i++) st[i] = (i & 1) ? P2G_BB_P - 1 : 0;
    bad += check_state(st, "alternating");
    // random states, and a chain (each output is the next input)
    for (int t = 0; t < 20000 && !bad; t++) {
        for (int i = 0; i < 16; i++) st[i] = (uint32_t)(rnd() % P2G_BB_P);
        bad += check_state(st, "random state");
    }
    for (int i = 0; i < 16; i++) st[i] = i;
    for (int t = 0; t < 2000 && !bad; t++) {
        bad += check_state(st, "chained state");
        p2o_poseidon2_bb(st);
    }
    // the signed product on the edge of its stated input range
    const int edge[] = {0, 1, -1, (int)P2G_BB_P, -(int)P2G_BB_P, (int)P2G_BB_P + 65536, -(int)P2G_BB_P - 65536, 0x40000000, -0x40000000};
    for (int a : edge)
        for (int b : edge) {
            const int r = bb_mul_signed(a, b);
            const long long lhs = (((long long)r % (long long)P2G_BB_P) + P2G_BB_P) % P2G_BB_P;
            const long long ab = ((((long long)a * b) % (long long)P2G_BB_P) + P2G_BB_P) % P2G_BB_P;
            if ((unsigned long long)lhs != (unsigned long long)BbHost::mul((uint32_t)ab, P2G_BB_RINV)) {
                std::printf("FAIL signed product %d * %d\n", a, b);
                bad++;
            }
        }
    std::printf(bad ? "FAILED\n" : "all ok\n");
    return bad ? 1 : 0;
}
