// BabyBear field (p = 15*2^27 + 1 = 2013265921) for sm_100a integer pipes.
//
// Internal representation: Montgomery form with R = 2^32, value in [0, p) -- the same form the external
// p3-monty-31 crate keeps in memory (Cargo.toml:22; 943718400 = 2^-32 mod p in the reference's
// gates/poseidon2_babybear.rs:776,790 is that factor).  One multiplication = IMAD.WIDE + IMAD + IMAD.WIDE
// and a conditional subtract.
#pragma once
#include <stdint.h>

namespace p2g {

typedef unsigned long long u64;
typedef unsigned int u32;

#define P2G_BB_P 2013265921u
#define P2G_BB_MU 0x77ffffffu      // -p^-1 mod 2^32
#define P2G_BB_PINV 0x88000001u    // p^-1 mod 2^32
#define P2G_BB_R 268435454u        // 2^32 mod p  (Montgomery form of 1)
#define P2G_BB_R2 1172168163u      // 2^64 mod p
#define P2G_BB_RINV 943718400u     // 2^-32 mod p

struct BbHost {  // host-side exact arithmetic on canonical values
    typedef u32 E;
    static constexpr u64 P = P2G_BB_P;
    static inline u32 mul(u32 a, u32 b) { return (u32)(((u64)a * b) % P); }
    static inline u32 add(u32 a, u32 b) { return (u32)(((u64)a + b) % P); }
    static inline u32 sub(u32 a, u32 b) { return (u32)(((u64)a + P - b) % P); }
    static inline u32 pow(u32 a, u64 e) {
        u32 r = 1;
        while (e) {
            if (e & 1) r = mul(r, a);
            a = mul(a, a);
            e >>= 1;
        }
        return r;
    }
    static inline u32 inv(u32 a) { return pow(a, P - 2); }
    static inline u32 to_internal(u32 canonical) { return mul(canonical, P2G_BB_R); }  // -> Montgomery form
};

// The arithmetic below is plain C++ apart from three PTX lines, so that tests/cpp/poseidon2_bb_emul_test.cpp can run
// the kernels' permutation on the host with every range assumption checked (P2G_BB_CHECK).
#if defined(__CUDACC__)
#define P2G_BB_FN __device__ __forceinline__
#else
#define P2G_BB_FN static inline
#endif
#ifndef P2G_BB_CHECK
#define P2G_BB_CHECK(cond) ((void)0)
#endif
P2G_BB_FN u32 bb_min_u32(u32 a, u32 b) { return a < b ? a : b; }

// x < 2^32 * p  ->  x * 2^-32 mod p, in [0, p)
P2G_BB_FN u32 bb_monty_reduce(u64 x) {
    P2G_BB_CHECK(x < ((u64)P2G_BB_P << 32));
    u32 m = (u32)x * P2G_BB_MU;
    u64 t = x + (u64)m * P2G_BB_P;  // low word becomes 0; no overflow: x < 2^32 p, m p < 2^32 p, 2^33 p < 2^64
    u32 r = (u32)(t >> 32);         // < 2p
    u32 r2 = r - P2G_BB_P;
    return bb_min_u32(r, r2);       // r - p wraps above r exactly when r < p
}
// the same without the conditional subtraction: x < 2^32 * B  ->  [0, p + B)
P2G_BB_FN u32 bb_monty_reduce_lazy(u64 x) {
    P2G_BB_CHECK(x < ((u64)1 << 62));
    u32 m = (u32)x * P2G_BB_MU;
    return (u32)((x + (u64)m * P2G_BB_P) >> 32);
}
P2G_BB_FN u32 bb_mul(u32 a, u32 b) { return bb_monty_reduce((u64)a * b); }
// Signed Montgomery product for multiplication CHAINS: |a|, |b| <= p + 2^16  ->  a b 2^-32 mod p as a signed word with
// |r| < |a b| / 2^32 + p / 2 < 0.97 p, so the result feeds the next product as it is.  m = (low word of a b) * p^-1 taken
// as a SIGNED word; m p has the same low word as a b, so the difference of the high words is exact: IMAD.WIDE, IMAD,
// IMAD.HI and one subtraction that either integer pipe can take -- where the unsigned form ends in a conditional
// subtraction (VIADDMNMX) that only the half-rate ALU pipe executes.  Measured on the permutation (tools/microbench):
// 3.51 -> 3.93 G permutations/s; pinning those subtractions to the multiplier pipe (an IMAD by a run-time -1) loses
// 2-13 % of that again, and so does taking m off the multiplier pipe (low word * p^-1 as two shifts, an add and an
// exclusive-or on the ALU pipe: 3.54-3.80 G/s, more registers), so the instruction choice is left as it is.  (A fused form -- the high word of a b - m p in one
// IMAD.HI with a 64-bit addend, as ptxas builds for the unsigned reduction -- is not reachable for signed operands:
// mad.wide.s32 with an addend is split into a product and a two-instruction 64-bit add.)
P2G_BB_FN int bb_mul_signed(int a, int b) {
    P2G_BB_CHECK((a < 0 ? -(long long)a : a) <= (long long)P2G_BB_P + 65536 && (b < 0 ? -(long long)b : b) <= (long long)P2G_BB_P + 65536);
#if defined(__CUDA_ARCH__)
    long long t;
    int m, h;
    asm("mul.wide.s32 %0, %1, %2;" : "=l"(t) : "r"(a), "r"(b));
    asm("mul.lo.s32 %0, %1, %2;" : "=r"(m) : "r"((int)t), "r"((int)P2G_BB_PINV));
    asm("mul.hi.s32 %0, %1, %2;" : "=r"(h) : "r"(m), "r"((int)P2G_BB_P));
    return (int)(t >> 32) - h;
#else
    const long long t = (long long)a * b;
    const int m = (int)((u32)t * P2G_BB_PINV);
    const int h = (int)(((long long)m * (long long)P2G_BB_P) >> 32);
    const int r = (int)(t >> 32) - h;
    P2G_BB_CHECK((r < 0 ? -(long long)r : r) < (long long)P2G_BB_P);
    return r;
#endif
}
// signed word in (-p, p) -> [0, p)
P2G_BB_FN u32 bb_from_signed(int r) {
    const u32 y = (u32)r;
    return bb_min_u32(y, y + P2G_BB_P);  // r < 0: y + p wraps to r + p < y;  r >= 0: y + p > y
}
P2G_BB_FN u32 bb_add(u32 a, u32 b) {
    P2G_BB_CHECK(a < P2G_BB_P && b < P2G_BB_P);
    u32 s = a + b;  // < 2p < 2^32
    return bb_min_u32(s, s - P2G_BB_P);
}
P2G_BB_FN u32 bb_sub(u32 a, u32 b) {
    u32 d = a - b;
    return bb_min_u32(d, d + P2G_BB_P);  // if a < b, d wrapped (large) and d + p is the answer
}

#ifdef __CUDACC__

struct BbDev {
    typedef u32 E;        // Montgomery form, [0, p)
    typedef u32 Storage;  // canonical or Montgomery u32, chosen per call (p2g `repr`)
    static __device__ __forceinline__ E mul(E a, E b) { return bb_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return bb_add(a, b); }
    static __device__ __forceinline__ E sub(E a, E b) { return bb_sub(a, b); }
    static __device__ __forceinline__ E zero() { return 0; }
    // canonical (any u32) -> Montgomery:  x * R^2 * R^-1
    static __device__ __forceinline__ E from_canonical(u32 x) { return bb_mul(x, P2G_BB_R2); }
    static __device__ __forceinline__ u32 to_canonical(E x) { return bb_monty_reduce((u64)x); }
    // Montgomery words from outside are accepted up to 2^32-1 and brought into [0, p)
    static __device__ __forceinline__ E from_monty(u32 x) {
        x = bb_min_u32(x, x - P2G_BB_P);
        return bb_min_u32(x, x - P2G_BB_P);
    }
};

#endif  // __CUDACC__

}  // namespace p2g
