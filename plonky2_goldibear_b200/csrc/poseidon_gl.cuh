// Poseidon over Goldilocks, width 12, x^7, 4 + 22 + 4 rounds -- one permutation per thread, state in registers.
//
// Replaces the reference's CPU permutation plonky2/src/hash/poseidon_goldilocks.rs:912-922 (full rounds :889-897,
// fast partial rounds :899-909 with mds_partial_layer_init :657-683 and mds_partial_layer_fast :718-744).
// Differences that do not change a single output bit:
//   * the MDS layer multiplies 32-bit halves of the state by the small circulant constants with IMAD.WIDE and
//     folds the NEXT round's constants into the accumulators (the reference adds them at the top of a round);
//   * state words are kept as arbitrary u64 representatives and made canonical only when a digest is stored;
//   * the sparse-factor tables are re-derived on the host at context creation (poseidon_tables.cpp), not pasted.
#pragma once
#include "field_gl.cuh"
#include "hash_tables.h"

namespace p2g {

#ifdef __CUDACC__

// the including translation unit (hash.cu) defines:  __constant__ PoseidonGlTables c_pgl;

__device__ __forceinline__ u64 pgl_sbox(u64 x) {
    u64 x2 = gl_mul(x, x);
    u64 x4 = gl_mul(x2, x2);
    u64 x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

// acc_lo + acc_hi * 2^32 (both < 2^42: sums of twelve 32-bit halves times constants <= 41, plus a constant half)
// -> representative.  With acc_hi = ah1*2^32 + ah0:  v = acc_lo + ah1*EPS + ah0*2^32  (2^64 = EPS); the first two
// terms cannot overflow, the last add can carry once, and that carry folds back in as +EPS without a second carry.
__device__ __forceinline__ u64 pgl_fold(u64 acc_lo, u64 acc_hi) {
    u32 al0 = (u32)acc_lo, al1 = (u32)(acc_lo >> 32), ah0 = (u32)acc_hi, ah1 = (u32)(acc_hi >> 32);
    u32 w0, w1;
    asm("{\n\t"
        ".reg .u32 c, nc;\n\t"
        "sub.cc.u32 %0, %2, %5;\n\t"      // acc_lo - ah1 ...
        "subc.u32 %1, %3, 0;\n\t"
        "add.u32 %1, %1, %5;\n\t"         // ... + ah1 * 2^32   (acc_lo + ah1*EPS, < 2^43)
        "add.cc.u32 %1, %1, %4;\n\t"      // + ah0 * 2^32
        "addc.u32 c, 0, 0;\n\t"
        "sub.u32 nc, 0, c;\n\t"           // carry * EPS = (0 : -carry)
        "add.cc.u32 %0, %0, nc;\n\t"
        "addc.u32 %1, %1, 0;\n\t"
        "}"
        : "=&r"(w0), "=&r"(w1)
        : "r"(al0), "r"(al1), "r"(ah0), "r"(ah1));
    return gl_pack(w0, w1);
}

// out = MDS * s (+ next round constants).  MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...)
// (poseidon_goldilocks.rs:301-302; row r of the product is sum_i circ[i] * s[(i+r)%12], :547-557).
template <bool ADD_CONST>
__device__ __forceinline__ void pgl_mds(u64 s[12], const u64 *next_rc) {
    const u32 circ[12] = {17, 15, 41, 16, 2, 28, 13, 13, 39, 18, 34, 20};
    u32 lo[12], hi[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        lo[i] = (u32)s[i];
        hi[i] = (u32)(s[i] >> 32);
    }
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = 0, ah = 0;
        if (ADD_CONST) {
            u64 c = next_rc[r];
            al = (u32)c;
            ah = c >> 32;
        }
#pragma unroll
        for (int i = 0; i < 12; i++) {
            al += (u64)lo[(i + r) % 12] * circ[i];
            ah += (u64)hi[(i + r) % 12] * circ[i];
        }
        if (r == 0) {
            al += (u64)lo[0] * 8u;
            ah += (u64)hi[0] * 8u;
        }
        s[r] = pgl_fold(al, ah);
    }
}

// Accumulator for sums of up to 2^32 128-bit products, kept per 32-bit column so that every partial product is one
// IMAD.WIDE with carry-out plus one carry add (no per-product 128-bit assembly):
//   sum = c0 + c1 * 2^32 + c2 * 2^64,   c0 = sum x0*y0, c1 = sum (x0*y1 + x1*y0), c2 = sum x1*y1, each 96 bits.
struct Acc192 {
    u32 c0l, c0h, c0x, c1l, c1h, c1x, c2l, c2h, c2x;
    __device__ __forceinline__ void init() { c0l = c0h = c0x = c1l = c1h = c1x = c2l = c2h = c2x = 0; }
    __device__ __forceinline__ void mac(u64 x, u64 y) {
        u32 x0 = (u32)x, x1 = (u32)(x >> 32), y0 = (u32)y, y1 = (u32)(y >> 32);
        asm("{\n\t"
            "mad.lo.cc.u32 %0, %9, %11, %0;\n\t"
            "madc.hi.cc.u32 %1, %9, %11, %1;\n\t"
            "addc.u32 %2, %2, 0;\n\t"
            "mad.lo.cc.u32 %3, %9, %12, %3;\n\t"
            "madc.hi.cc.u32 %4, %9, %12, %4;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "mad.lo.cc.u32 %3, %10, %11, %3;\n\t"
            "madc.hi.cc.u32 %4, %10, %11, %4;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "mad.lo.cc.u32 %6, %10, %12, %6;\n\t"
            "madc.hi.cc.u32 %7, %10, %12, %7;\n\t"
            "addc.u32 %8, %8, 0;\n\t"
            "}"
            : "+r"(c0l), "+r"(c0h), "+r"(c0x), "+r"(c1l), "+r"(c1h), "+r"(c1x), "+r"(c2l), "+r"(c2h), "+r"(c2x)
            : "r"(x0), "r"(x1), "r"(y0), "r"(y1));
    }
    // x * small (y < 2^32): only the two low columns
    __device__ __forceinline__ void mac_small(u64 x, u32 y) {
        u32 x0 = (u32)x, x1 = (u32)(x >> 32);
        asm("{\n\t"
            "mad.lo.cc.u32 %0, %6, %8, %0;\n\t"
            "madc.hi.cc.u32 %1, %6, %8, %1;\n\t"
            "addc.u32 %2, %2, 0;\n\t"
            "mad.lo.cc.u32 %3, %7, %8, %3;\n\t"
            "madc.hi.cc.u32 %4, %7, %8, %4;\n\t"
            "addc.u32 %5, %5, 0;\n\t"
            "}"
            : "+r"(c0l), "+r"(c0h), "+r"(c0x), "+r"(c1l), "+r"(c1h), "+r"(c1x)
            : "r"(x0), "r"(x1), "r"(y));
    }
    __device__ __forceinline__ u64 reduce() const {
        // limbs t0..t4 of the 160-bit sum
        u32 t0, t1, t2, t3, t4;
        asm("{\n\t"
            "add.cc.u32 %1, %6, %8;\n\t"       // t1 = c0h + c1l
            "addc.cc.u32 %2, %7, %9;\n\t"      // t2 = c0x + c1h
            "addc.cc.u32 %3, %10, 0;\n\t"      // t3 = c1x
            "addc.u32 %4, 0, 0;\n\t"
            "add.cc.u32 %2, %2, %11;\n\t"      // + c2l
            "addc.cc.u32 %3, %3, %12;\n\t"     // + c2h
            "addc.u32 %4, %4, %13;\n\t"        // + c2x
            "mov.u32 %0, %5;\n\t"
            "}"
            : "=&r"(t0), "=&r"(t1), "=&r"(t2), "=&r"(t3), "=&r"(t4)
            : "r"(c0l), "r"(c0h), "r"(c0x), "r"(c1l), "r"(c1h), "r"(c1x), "r"(c2l), "r"(c2h), "r"(c2x));
        u64 mid = gl_reduce128(gl_pack(t2, t3), (u64)t4);  // (t4:t3:t2) * 2^64 part
        return gl_reduce128(gl_pack(t0, t1), mid);
    }
};

// Code-size discipline: the whole permutation is ~28k dynamic instructions, and a warp that strays from its
// neighbours misses the instruction cache on every line (first ncu capture: 13 of 18 stalled warps per issue were
// `no_instruction`).  So there is ONE full-round body (used by both halves), ONE partial-round body and ONE
// pre-matrix row body, each a real loop, and P2G_ROUND_SYNC keeps the warps of a CTA on the same lines.
#ifndef P2G_ROUND_SYNC
#define P2G_ROUND_SYNC() __syncthreads()
#endif

__device__ __forceinline__ void poseidon_gl_permute(u64 s[12]) {
#pragma unroll 1
    for (int half = 0; half < 2; half++) {
        // ---- 4 full rounds; the constants of the following layer are folded into each MDS (c_pgl.mds_add)
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_pgl.pre_add[half][i]);
#pragma unroll 1
        for (int r = 0; r < 4; r++) {
            P2G_ROUND_SYNC();
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = pgl_sbox(s[i]);
            pgl_mds<true>(s, c_pgl.mds_add[half * 4 + r]);
        }
        if (half == 0) {
            // ---- dense pre-matrix of the partial rounds (mds_partial_layer_init), one output row per iteration
            u64 t[11];
#pragma unroll 1
            for (int i = 0; i < 11; i++) {
                Acc192 acc;
                acc.init();
#pragma unroll
                for (int j = 0; j < 11; j++) acc.mac(s[j + 1], c_pgl.init[i][j]);
                t[i] = acc.reduce();
            }
#pragma unroll
            for (int i = 0; i < 11; i++) s[i + 1] = t[i];
            // ---- 22 partial rounds in sparse-factor form
#pragma unroll 1
            for (int r = 0; r < 22; r++) {
                P2G_ROUND_SYNC();
                u64 s0 = gl_add_c(pgl_sbox(s[0]), c_pgl.k[r]);
                Acc192 acc;
                acc.init();
                acc.mac_small(s0, 25u);  // M[0][0] = 17 + 8
#pragma unroll
                for (int i = 0; i < 11; i++) acc.mac(s[i + 1], c_pgl.row[r][i]);
#pragma unroll
                for (int i = 0; i < 11; i++) s[i + 1] = gl_mul_add(s0, c_pgl.col[r][i], s[i + 1]);
                s[0] = acc.reduce();
            }
        }
    }
}

#endif  // __CUDACC__
}  // namespace p2g
