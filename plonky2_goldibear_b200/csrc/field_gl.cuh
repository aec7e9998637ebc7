// Goldilocks field (p = 2^64 - 2^32 + 1) for sm_100a integer pipes.
//
// Representation: a u64 holding ANY representative of the residue (0 <= x < 2^64, not necessarily < p).
// Reduction is Montgomery-free: 2^64 = 2^32 - 1 (=: EPS) and 2^96 = -1 (mod p), so a 128-bit product
// (h1:h0:l1:l0 in 32-bit limbs) reduces to  lo + h0*EPS - h1  plus at most one +-EPS correction.
// Values are made canonical (< p) only when they are stored as results.
//
// The reference gets this arithmetic from the external p3-goldilocks crate (Cargo.toml:19); the identities
// are the ones the reference itself uses in plonky2/src/hash/poseidon_goldilocks.rs:254-268.
#pragma once
#include <stdint.h>

namespace p2g {

typedef unsigned long long u64;
typedef unsigned int u32;

#define P2G_GL_P 0xFFFFFFFF00000001ULL
#define P2G_GL_EPS 0xFFFFFFFFULL

struct GlHost {  // host-side exact arithmetic used to build tables
    typedef u64 E;
    static constexpr u64 P = P2G_GL_P;
    static inline u64 reduce128(unsigned __int128 x) { return (u64)(x % P); }
    static inline u64 mul(u64 a, u64 b) { return (u64)(((unsigned __int128)a * b) % P); }
    static inline u64 add(u64 a, u64 b) { return (u64)(((unsigned __int128)a + b) % P); }
    static inline u64 sub(u64 a, u64 b) { return (u64)(((unsigned __int128)a + P - b) % P); }
    static inline u64 pow(u64 a, u64 e) {
        u64 r = 1;
        while (e) {
            if (e & 1) r = mul(r, a);
            a = mul(a, a);
            e >>= 1;
        }
        return r;
    }
    static inline u64 inv(u64 a) { return pow(a, P - 2); }
    static inline u64 to_internal(u64 canonical) { return canonical; }
};

#ifdef __CUDACC__

__device__ __forceinline__ u64 gl_pack(u32 lo, u32 hi) { return ((u64)hi << 32) | lo; }

// (hi:lo) -> any-representative u64.   v = lo + h0*EPS - h1 lies in (-2^32, 2^65); the signed carry word c
// (-1, 0, +1) says which multiple of 2^64 to fold back in as c*EPS; the folded value always fits in 64 bits.
__device__ __forceinline__ u64 gl_reduce128(u64 lo, u64 hi) {
    u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
    u32 u0, u1, c;
    asm("{\n\t"
        "mad.lo.cc.u32 %0, %5, 0xffffffff, %3;\n\t"
        "madc.hi.cc.u32 %1, %5, 0xffffffff, %4;\n\t"
        "addc.u32 %2, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, %6;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 %2, %2, 0;\n\t"
        "}"
        : "=&r"(u0), "=&r"(u1), "=&r"(c)
        : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
    u32 nc = 0u - c;                 // low word of c*EPS
    u32 sc = (u32)((int)c >> 31);    // high word of c*EPS (sign extension)
    asm("{\n\t"
        "add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, %3;\n\t"
        "}"
        : "+r"(u0), "+r"(u1)
        : "r"(nc), "r"(sc));
    return gl_pack(u0, u1);
}

// full 64x64 -> 128 product; four IMAD.WIDE plus carry glue
__device__ __forceinline__ void gl_mul128(u64 a, u64 b, u64 &lo, u64 &hi) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    u32 r0, r1, r2, r3;
    asm("{\n\t"
        "mul.lo.u32 %0,%4,%6;\n\t"
        "mul.hi.u32 %1,%4,%6;\n\t"
        "mad.lo.cc.u32 %1,%4,%7,%1;\n\t"
        "madc.hi.u32 %2,%4,%7,0;\n\t"
        "mad.lo.cc.u32 %1,%5,%6,%1;\n\t"
        "madc.hi.cc.u32 %2,%5,%6,%2;\n\t"
        "addc.u32 %3,0,0;\n\t"
        "mad.lo.cc.u32 %2,%5,%7,%2;\n\t"
        "madc.hi.u32 %3,%5,%7,%3;\n\t"
        "}"
        : "=&r"(r0), "=&r"(r1), "=&r"(r2), "=&r"(r3)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    lo = gl_pack(r0, r1);
    hi = gl_pack(r2, r3);
}

__device__ __forceinline__ u64 gl_mul(u64 a, u64 b) {
    u64 lo, hi;
    gl_mul128(a, b, lo, hi);
    return gl_reduce128(lo, hi);
}

// a*b + c  (c any u64): the addend rides in the 128-bit product (hi <= 2^64 - 2^33 so the carry cannot overflow)
__device__ __forceinline__ u64 gl_mul_add(u64 a, u64 b, u64 c) {
    u64 lo, hi;
    gl_mul128(a, b, lo, hi);
    u64 lo2 = lo + c;
    hi += (lo2 < lo);
    return gl_reduce128(lo2, hi);
}

__device__ __forceinline__ u64 gl_canon(u64 a) { return a >= P2G_GL_P ? a - P2G_GL_P : a; }

// a + b for arbitrary representatives (two possible wrap corrections)
__device__ __forceinline__ u64 gl_add(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a) {
        u64 t = s + P2G_GL_EPS;
        if (t < s) t += P2G_GL_EPS;
        s = t;
    }
    return s;
}
// a - b for arbitrary representatives
__device__ __forceinline__ u64 gl_sub(u64 a, u64 b) {
    u64 d = a - b;
    if (a < b) {
        u64 t = d - P2G_GL_EPS;
        if (t > d) t -= P2G_GL_EPS;
        d = t;
    }
    return d;
}
// a + b where b is canonical (< p): a single correction is exact (a + b < 2^64 + p)
__device__ __forceinline__ u64 gl_add_c(u64 a, u64 b) {
    u64 s = a + b;
    if (s < a) s += P2G_GL_EPS;
    return s;
}
// a - b where b is canonical (< p): a single correction is exact
__device__ __forceinline__ u64 gl_sub_c(u64 a, u64 b) {
    u64 d = a - b;
    if (a < b) d -= P2G_GL_EPS;
    return d;
}

struct GlDev {
    typedef u64 E;       // internal representation: any representative
    typedef u64 Storage;  // memory representation: canonical u64 (any u64 accepted on load)
    static __device__ __forceinline__ E load(Storage x) { return x; }
    static __device__ __forceinline__ Storage store(E x) { return gl_canon(x); }
    static __device__ __forceinline__ E mul(E a, E b) { return gl_mul(a, b); }
    static __device__ __forceinline__ E add(E a, E b) { return gl_add(a, b); }
    static __device__ __forceinline__ E sub(E a, E b) { return gl_sub(a, b); }
    static __device__ __forceinline__ E zero() { return 0; }
};

#endif  // __CUDACC__

}  // namespace p2g
