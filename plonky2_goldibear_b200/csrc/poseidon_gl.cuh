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

// acc_lo + acc_hi * 2^32 (both < 2^63) -> representative
__device__ __forceinline__ u64 pgl_fold(u64 acc_lo, u64 acc_hi) {
    u64 x = acc_lo + (acc_hi << 32);
    u32 top = (u32)(acc_hi >> 32) + (u32)(x < acc_lo);  // multiples of 2^64
    u64 t = (u64)top * 0xffffffffu;
    u64 r = x + t;
    if (r < t) r += P2G_GL_EPS;
    return r;
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

// 192-bit accumulator for sums of 128-bit products
struct Acc192 {
    u64 a0, a1, a2;
    __device__ __forceinline__ void init() { a0 = a1 = a2 = 0; }
    __device__ __forceinline__ void mac(u64 x, u64 y) {
        u64 lo, hi;
        gl_mul128(x, y, lo, hi);
        u32 l0 = (u32)lo, l1 = (u32)(lo >> 32), h0 = (u32)hi, h1 = (u32)(hi >> 32);
        u32 r0 = (u32)a0, r1 = (u32)(a0 >> 32), r2 = (u32)a1, r3 = (u32)(a1 >> 32), r4 = (u32)a2;
        asm("{\n\t"
            "add.cc.u32 %0,%0,%5;\n\t"
            "addc.cc.u32 %1,%1,%6;\n\t"
            "addc.cc.u32 %2,%2,%7;\n\t"
            "addc.cc.u32 %3,%3,%8;\n\t"
            "addc.u32 %4,%4,0;\n\t"
            "}"
            : "+r"(r0), "+r"(r1), "+r"(r2), "+r"(r3), "+r"(r4)
            : "r"(l0), "r"(l1), "r"(h0), "r"(h1));
        a0 = gl_pack(r0, r1);
        a1 = gl_pack(r2, r3);
        a2 = r4;
    }
    __device__ __forceinline__ u64 reduce() const {
        u64 mid = gl_reduce128(a1, a2);  // (a2:a1) * 2^64 part
        return gl_reduce128(a0, mid);
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
                acc.mac(s0, 25);  // M[0][0] = 17 + 8
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
