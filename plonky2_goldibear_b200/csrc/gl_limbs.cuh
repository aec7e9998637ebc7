// Goldilocks butterflies without multiplications: radix-2^K DFT kernels on 24-bit limbs.
//
// In Goldilocks 2 has order 192 (2^96 = -1), and the two-adic roots the reference's FFT uses up to order 64 are powers
// of two: two_adic_generator(4) = 2^156 = -2^60, two_adic_generator(3) = 2^120 = -2^24, two_adic_generator(2) = 2^48
// (field/src/fft.rs:16 takes them from the external p3-goldilocks crate; values checked in tests/test_oracle.py).
// So a 16-point DFT needs no field multiplication at all -- only multiplications by 2^(12 j).
//
// Representation: X = 2^24, X^4 = 2^96 = -1.  An element is c0 + c1 X + c2 X^2 + c3 X^3 with SIGNED 32-bit
// coefficients ("limbs") that are allowed to grow: a fresh split has limbs in [0, 2^24), every butterfly level adds one
// bit, and 16 points take four levels, so limbs stay below 2^28 in magnitude.  In this ring
//   * add / sub are four independent 32-bit adds with no carries: they issue on either integer pipe (IADD3 on the
//     ALU pipe or IMAD.IADD on the FMA pipe) -- the carry-chain add/sub of the u64 representation (10 + 8 ALU-pipe
//     instructions per butterfly) were what bound the round-1 NTT;
//   * multiplication by X^q is a negacyclic rotation of the limbs: free (register renaming, signs folded into the
//     following add / sub);
//   * multiplication by 2^12 X^q, the odd powers of the 16th root, is a rotation plus a carry-less renormalising shift
//     (3 instructions per limb).  Only 4 of the 32 butterflies of a 16-point DFT need it; radix 8, 4 and 2 need none.
// Leaving the ring costs one "join": sum_i (c_i + bias_i) * W_i with W_i = w * 2^(24 i) mod p as 8 IMAD.WIDE into two
// 64-bit accumulators (no carries: 4 * 2^29.1 * 2^32 < 2^64), followed by a 96-bit -> 64-bit fold.  The round's twiddle
// factor w rides in the W_i for free: the twiddle table holds the four W_i of every root (32 bytes per entry).
//
// Everything here is plain C++ with a PTX twin for three device-side multiply-adds, so the host-side emulation test
// (tests/cpp/ntt_tile_emul_test.cpp) runs the very same algorithm against a naive DFT.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define P2G_HD __host__ __device__ __forceinline__
#else
#define P2G_HD inline
#endif

namespace p2g {
namespace gll {

typedef unsigned long long u64;
typedef unsigned int u32;

// bias_i >= 2^28 with sum_i bias_i 2^(24 i) = 0 (mod p): makes the limbs non-negative for the unsigned IMAD.WIDE
// without changing the value.  (2^28 (1 + X + X^2 + X^3) + 0xffefeffef0001011 = 0 mod p; tests/test_oracle.py)
#define P2G_GLL_BIAS0 0x10001011u
#define P2G_GLL_BIAS1 0x10effef0u
#define P2G_GLL_BIAS2 0x1000ffefu
#define P2G_GLL_BIAS3 0x10000000u

P2G_HD void split(u64 x, int (&c)[4]) {
    const u32 lo = (u32)x, hi = (u32)(x >> 32);
    c[0] = (int)(lo & 0xffffffu);
    c[1] = (int)(((lo >> 24) | (hi << 8)) & 0xffffffu);
    c[2] = (int)(hi >> 16);
    c[3] = 0;
}

// d *= 2^sh, sh a multiple of 12 in [0, 192).  `sh` must be a compile-time constant after unrolling.
P2G_HD void mul_pow2(int (&d)[4], int sh) {
    const bool neg = sh >= 96;
    if (neg) sh -= 96;
    const int q = sh / 24, t = sh % 24;
    int r[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int k = i + q;
        int val = d[i];
        if (k >= 4) {
            k -= 4;
            val = -val;
        }
        if (neg) val = -val;
        r[k] = val;
    }
    if (t) {
        // times 2^12: limb i = h_i 2^12 + l_i  ->  l_i 2^12 stays in limb i, h_i moves up one limb (h_3 wraps negated)
        int h[4], l[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            h[i] = r[i] >> 12;  // arithmetic
            l[i] = r[i] & 0xfff;
        }
        d[0] = (l[0] << 12) - h[3];
#pragma unroll
        for (int i = 1; i < 4; i++) d[i] = (l[i] << 12) + h[i - 1];
    } else {
#pragma unroll
        for (int i = 0; i < 4; i++) d[i] = r[i];
    }
}

// E16: the 16th root of the transform direction is 2^(12 * E16): 13 forward (2^156), 3 inverse (2^36)
template <int DIR>
struct Dir {
    static const int E16 = DIR ? 3 : 13;
};

P2G_HD u32 bitrev_small(u32 x, int bits) {
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
    return r;
}

// Decimation-in-frequency DFT of R = 2^K elements in DEPTH-FIRST order: frequency digit k = bitrev_K(s) of the inputs
// v[base .. base + R) is handed to emit(s, limbs) as soon as it is final, v[bitrev_K(k)] = sum_j v_in[j] w_R^(j k), w_R the
// R-th root of the direction.  One butterfly level over the block, then the two halves one after the other: the joins
// of a finished half (multiplier pipe) sit next to the additions of the next one (either pipe) in program order, and
// the limbs of finished outputs die early.  Fully unrolled; every twiddle is a compile-time power of two.
template <int K, int DIR, int N, class Emit>
P2G_HD void dft_emit(int (&v)[N][4], int base, Emit &emit) {
    if constexpr (K == 0) {
        emit(base, v[base]);
    } else {
        constexpr int hj = 1 << (K - 1);
#pragma unroll
        for (int j = 0; j < hj; j++) {
            // twiddle w_{2^K}^j = 2^(12 * E16 * (16 / 2^K) * j)
            const int e = (j * (16 >> K) * Dir<DIR>::E16) & 15;
            int d[4];
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const int a = v[base + j][i], b = v[base + j + hj][i];
                v[base + j][i] = a + b;
                d[i] = a - b;
            }
            mul_pow2(d, 12 * e);
#pragma unroll
            for (int i = 0; i < 4; i++) v[base + j + hj][i] = d[i];
        }
        dft_emit<K - 1, DIR, N, Emit>(v, base, emit);
        dft_emit<K - 1, DIR, N, Emit>(v, base + hj, emit);
    }
}

struct EmitNothing {
    P2G_HD void operator()(int, const int (&)[4]) const {}
};
// the same, in place, outputs left in v
template <int K, int DIR>
P2G_HD void dft(int (&v)[1 << K][4]) {
    EmitNothing e;
    dft_emit<K, DIR, 1 << K, EmitNothing>(v, 0, e);
}

// L + H * 2^32 (mod p) for L, H < 2^63.1: the 96-bit value w0 + w1 2^32 + w2 2^64 folds with 2^64 = 2^32 - 1 into
// (w1:w0) + w2 * (2^32 - 1); that sum can pass 2^64 once, which shows in the high word alone (the addend is below
// 2^63.2, so a wrapped sum is more than 2^32 below the first operand), and the wrap folds back in as + (2^32 - 1).
// `eps` = 2^32 - 1, handed in as a run-time value on purpose: with the literal, ptxas rewrites each multiply-add by
// 0xffffffff into a four-instruction negate / high-product / carry sequence instead of one IMAD.WIDE.
#ifndef P2G_FOLD96_VARIANT
#define P2G_FOLD96_VARIANT 2
#endif
P2G_HD u64 fold96(u64 L, u64 H, u32 eps) {
#if defined(__CUDA_ARCH__)
    u64 r;
#if P2G_FOLD96_VARIANT == 2
    // all on the ALU pipe: w2 (2^32 - 1) = (w2 << 32) - w2 is a subtraction from the low word and an addition to the high
    // one; the multiplier pipe is the busier one in the NTT kernels (8 IMAD.WIDE per join before the fold)
    (void)eps;
    asm("{\n\t"
        ".reg .u32 l0, l1, h0, h1, w1, w2, r0, r1, t;\n\t"
        ".reg .pred q;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "mov.b64 {h0, h1}, %2;\n\t"
        "add.cc.u32 w1, l1, h0;\n\t"
        "addc.u32 w2, h1, 0;\n\t"
        "sub.cc.u32 r0, l0, w2;\n\t"
        "subc.u32 t, w1, 0;\n\t"
        "add.u32 r1, t, w2;\n\t"
        "setp.lt.u32 q, r1, w1;\n\t"
        "@q add.cc.u32 r0, r0, 0xffffffff;\n\t"
        "@q addc.u32 r1, r1, 0;\n\t"
        "mov.b64 %0, {r0, r1};\n\t"
        "}"
        : "=&l"(r)
        : "l"(L), "l"(H));
#else
    asm("{\n\t"
        ".reg .u32 l0, l1, h0, h1, w1, w2, r0, r1, e;\n\t"
        ".reg .u64 a;\n\t"
        ".reg .pred q;\n\t"
        "mov.b64 {l0, l1}, %1;\n\t"
        "mov.b64 {h0, h1}, %2;\n\t"
        "add.cc.u32 w1, l1, h0;\n\t"
        "addc.u32 w2, h1, 0;\n\t"
        "mov.b64 a, {l0, w1};\n\t"
        "mad.wide.u32 %0, w2, %3, a;\n\t"
        "mov.b64 {r0, r1}, %0;\n\t"
        "setp.lt.u32 q, r1, w1;\n\t"
#if P2G_FOLD96_VARIANT == 0
        "selp.u32 e, 1, 0, q;\n\t"
        "mad.wide.u32 %0, e, %3, %0;\n\t"
#else  // the wrap folded back in by two predicated adds on the ALU pipe instead of a third IMAD.WIDE
        "@q add.cc.u32 r0, r0, 0xffffffff;\n\t"
        "@q addc.u32 r1, r1, 0;\n\t"
        "mov.b64 %0, {r0, r1};\n\t"
#endif
        "}"
        : "=&l"(r)
        : "l"(L), "l"(H), "r"(eps));
#endif
    return r;
#else
    const u32 l0 = (u32)L, l1 = (u32)(L >> 32), h0 = (u32)H, h1 = (u32)(H >> 32);
    const u64 mid = (u64)l1 + h0;
    const u32 w1 = (u32)mid;
    const u32 w2 = h1 + (u32)(mid >> 32);
    const u64 a = ((u64)w1 << 32) | l0;
    u64 r = a + (u64)w2 * eps;
    if ((u32)(r >> 32) < w1) r += eps;
    return r;
#endif
}

// a * b + c on the multiplier pipe (IMAD.WIDE.U32); written as PTX on the device because the C expression leaves
// adds of a zero high word behind
P2G_HD u64 madw(u32 a, u32 b, u64 c) {
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(r) : "r"(a), "r"(b), "l"(c));
    return r;
#else
    return (u64)a * b + c;
#endif
}
P2G_HD u64 mulw(u32 a, u32 b) {
#if defined(__CUDA_ARCH__)
    u64 r;
    asm("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b));
    return r;
#else
    return (u64)a * b;
#endif
}

// (c0 + c1 X + c2 X^2 + c3 X^3) * w  (mod p), any representative, from NON-NEGATIVE limbs below 2^29 + 2^24 that carry
// the bias (see add_bias) and W_i = w * 2^(24 i) mod p (any u64 values congruent to that).
P2G_HD u64 join(const int (&c)[4], u64 W0, u64 W1, u64 W2, u64 W3, u32 eps) {
    const u32 u0 = (u32)c[0], u1 = (u32)c[1], u2 = (u32)c[2], u3 = (u32)c[3];
    u64 L = mulw(u0, (u32)W0);
    L = madw(u1, (u32)W1, L);
    L = madw(u2, (u32)W2, L);
    L = madw(u3, (u32)W3, L);
    u64 H = mulw(u0, (u32)(W0 >> 32));
    H = madw(u1, (u32)(W1 >> 32), H);
    H = madw(u2, (u32)(W2 >> 32), H);
    H = madw(u3, (u32)(W3 >> 32), H);
    return fold96(L, H, eps);
}

// the same for w = 1: W = (1, 2^24, 2^48, 2^72 = 2^40 - 2^8), five products instead of eight
P2G_HD u64 join_one(const int (&c)[4], u32 eps) {
    const u32 u0 = (u32)c[0], u1 = (u32)c[1], u2 = (u32)c[2], u3 = (u32)c[3];
    u64 L = madw(u1, 0x1000000u, (u64)u0);
    L = madw(u3, 0xffffff00u, L);
    u64 H = mulw(u2, 0x10000u);
    H = madw(u3, 0xffu, H);
    return fold96(L, H, eps);
}

// The joins multiply unsigned: the limbs must not be negative.  A DFT output is sum_j x_j w^(jk) and x_0 enters every
// output with the factor 1, so adding the bias to the limbs of x_0 BEFORE the transform adds it to every output:
// 4 adds per unit instead of 4 per element.  (x_0 never meets a shifting twiddle: those sit on odd j.)
P2G_HD void add_bias(int (&c)[4]) {
    c[0] += (int)P2G_GLL_BIAS0;
    c[1] += (int)P2G_GLL_BIAS1;
    c[2] += (int)P2G_GLL_BIAS2;
    c[3] += (int)P2G_GLL_BIAS3;
}

}  // namespace gll
}  // namespace p2g
