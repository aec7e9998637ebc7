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

#ifdef __CUDACC__

// x < 2^32 * p  ->  x * 2^-32 mod p, in [0, p)
__device__ __forceinline__ u32 bb_monty_reduce(u64 x) {
    u32 m = (u32)x * P2G_BB_MU;
    u64 t = x + (u64)m * P2G_BB_P;  // low word becomes 0; no overflow: x < 2^32 p, m p < 2^32 p, 2^33 p < 2^64
    u32 r = (u32)(t >> 32);         // < 2p
    u32 r2 = r - P2G_BB_P;
    return min(r, r2);              // r - p wraps above r exactly when r < p
}
__device__ __forceinline__ u32 bb_mul(u32 a, u32 b) { return bb_monty_reduce((u64)a * b); }
__device__ __forceinline__ u32 bb_add(u32 a, u32 b) {
    u32 s = a + b;  // < 2p < 2^32
    return min(s, s - P2G_BB_P);
}
__device__ __forceinline__ u32 bb_sub(u32 a, u32 b) {
    u32 d = a - b;
    return min(d, d + P2G_BB_P);  // if a < b, d wrapped (large) and d + p is the answer
}

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
        x = min(x, x - P2G_BB_P);
        return min(x, x - P2G_BB_P);
    }
};

#endif  // __CUDACC__

}  // namespace p2g
