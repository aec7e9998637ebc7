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

// Issue costs measured on B200 (tools/microbench/imad_probe.cu): IMAD.WIDE / IMAD.HI ~3.6 cycles per warp
// instruction, 32-bit IMAD ~2, IADD3 / LOP3 / SEL / SHF ~1, with no overlap between them.  So the arithmetic below is
// written to use exactly four IMAD.WIDE per 64x64 product and carry-chain adds for everything else.

// (hi:lo) -> any-representative u64.   v = lo + h0*2^32 - (h0 + h1) lies in (-2^32, 2^65); the signed carry word t
// (-1, 0, +1) says which multiple of 2^64 to fold back in as t*EPS; the folded value always fits in 64 bits.
__device__ __forceinline__ u64 gl_reduce128(u64 lo, u64 hi) {
    u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
    u32 v0, v1, t;
    asm("{\n\t"
        ".reg .u32 s0, cs;\n\t"
        "add.cc.u32 s0, %5, %6;\n\t"     // s = h0 + h1 (33 bits)
        "addc.u32 cs, 0, 0;\n\t"
        "sub.cc.u32 %0, %3, s0;\n\t"     // lo - s
        "subc.cc.u32 %1, %4, cs;\n\t"
        "subc.u32 %2, 0, 0;\n\t"         // t = -borrow
        "add.cc.u32 %1, %1, %5;\n\t"     // + h0 * 2^32
        "addc.u32 %2, %2, 0;\n\t"        // t in {-1, 0, 1}
        "}"
        : "=&r"(v0), "=&r"(v1), "=&r"(t)
        : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
    u32 nt = 0u - t;                // low word of t*EPS
    u32 st = (u32)((int)t >> 31);   // high word of t*EPS (sign extension)
    asm("{\n\t"
        "add.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, %3;\n\t"
        "}"
        : "+r"(v0), "+r"(v1)
        : "r"(nt), "r"(st));
    return gl_pack(v0, v1);
}

// full 64x64 -> 128 product: four IMAD.WIDE (a0b0, a0b1, a1b0 accumulated onto a0b1 with carry-out, a1b1) joined by
// 64-bit carry adds.  Written with 64-bit PTX values on purpose: with 32-bit mul.lo/mul.hi pairs ptxas splits a0*b0 into
// IMAD.HI + IMAD (5.6 issue cycles instead of 3.6) to fold an addend.
__device__ __forceinline__ void gl_mul128(u64 a, u64 b, u64 &lo, u64 &hi) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32);
    asm("{\n\t"
        ".reg .u64 p0, p3, m, t;\n\t"
        ".reg .u32 m0, m1, mc, z;\n\t"
        "mul.wide.u32 p0, %2, %4;\n\t"
        "mul.wide.u32 m, %2, %5;\n\t"
        "mov.b64 {m0, m1}, m;\n\t"
        "mad.lo.cc.u32 m0, %3, %4, m0;\n\t"
        "madc.hi.cc.u32 m1, %3, %4, m1;\n\t"
        "addc.u32 mc, 0, 0;\n\t"
        "mul.wide.u32 p3, %3, %5;\n\t"
        "mov.u32 z, 0;\n\t"
        "mov.b64 t, {z, m0};\n\t"
        "add.cc.u64 %0, p0, t;\n\t"
        "mov.b64 t, {m1, mc};\n\t"
        "addc.u64 %1, p3, t;\n\t"
        "}"
        : "=l"(lo), "=l"(hi)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}

// a^2 with three products (a0^2, a0 a1, a1^2) and the cross term doubled by funnel shifts: one IMAD.WIDE and the
// register-pair copies of the second cross product less, two shifts more.  Costed as a loss in round 1 (the ALU pipe is
// the scarce one), measured as a gain in round 2: 27.32k -> 26.47k cycles per warp-permutation with the two squarings of
// every S-box (tools/microbench/poseidon_bench.cu).
__device__ __forceinline__ void gl_sqr128(u64 a, u64 &lo, u64 &hi) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32);
    asm("{\n\t"
        ".reg .u64 p0, p3, m, t;\n\t"
        ".reg .u32 m0, m1, d0, d1, d2, z;\n\t"
        "mul.wide.u32 p0, %2, %2;\n\t"
        "mul.wide.u32 m, %2, %3;\n\t"
        "mul.wide.u32 p3, %3, %3;\n\t"
        "mov.b64 {m0, m1}, m;\n\t"
        "shl.b32 d0, m0, 1;\n\t"
        "shf.l.wrap.b32 d1, m0, m1, 1;\n\t"
        "shr.b32 d2, m1, 31;\n\t"
        "mov.u32 z, 0;\n\t"
        "mov.b64 t, {z, d0};\n\t"
        "add.cc.u64 %0, p0, t;\n\t"
        "mov.b64 t, {d1, d2};\n\t"
        "addc.u64 %1, p3, t;\n\t"
        "}"
        : "=l"(lo), "=l"(hi)
        : "r"(a0), "r"(a1));
}

__device__ __forceinline__ u64 gl_mul(u64 a, u64 b) {
    u64 lo, hi;
    gl_mul128(a, b, lo, hi);
    return gl_reduce128(lo, hi);
}

__device__ __forceinline__ u64 gl_sqr(u64 a) {
    u64 lo, hi;
    gl_sqr128(a, lo, hi);
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

// a + b for arbitrary representatives: 2^64 = EPS, so a carry folds back in as +EPS; that can carry once more
// (only when both inputs were >= p), never a third time.
__device__ __forceinline__ u64 gl_add(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), s0, s1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "add.cc.u32 %0, %2, %4;\n\t"
        "addc.cc.u32 %1, %3, %5;\n\t"
        "addc.u32 m, 0, 0;\n\t"           // carry
        "sub.u32 m, 0, m;\n\t"            // -carry = low word of carry*EPS
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.cc.u32 %1, %1, 0;\n\t"
        "addc.u32 m, 0, 0;\n\t"
        "sub.u32 m, 0, m;\n\t"
        "add.cc.u32 %0, %0, m;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(s0), "=&r"(s1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return gl_pack(s0, s1);
}
// a - b for arbitrary representatives: a borrow folds back in as -EPS (= subtracting the mask 0xffffffff from the low
// word with borrow propagation), at most twice.
__device__ __forceinline__ u64 gl_sub(u64 a, u64 b) {
    u32 a0 = (u32)a, a1 = (u32)(a >> 32), b0 = (u32)b, b1 = (u32)(b >> 32), d0, d1;
    asm("{\n\t"
        ".reg .u32 m;\n\t"
        "sub.cc.u32 %0, %2, %4;\n\t"
        "subc.cc.u32 %1, %3, %5;\n\t"
        "subc.u32 m, 0, 0;\n\t"          // m = -borrow
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.cc.u32 %1, %1, 0;\n\t"
        "subc.u32 m, 0, 0;\n\t"
        "sub.cc.u32 %0, %0, m;\n\t"
        "subc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(d0), "=&r"(d1)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
    return gl_pack(d0, d1);
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
