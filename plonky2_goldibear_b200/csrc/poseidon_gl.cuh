// Poseidon over Goldilocks, width 12, x^7, 4 + 22 + 4 rounds -- one permutation per thread, state in registers.
//
// Replaces the reference's CPU permutation plonky2/src/hash/poseidon_goldilocks.rs:912-922 (same function as its
// plain form :927-948: constant layer :519-535, S-box :746-760, MDS :547-557).
// Differences that do not change a single output bit:
//   * the MDS layer works on three 22-bit limbs of the state with shifts and adds only (pgl_circ_limb) and folds the
//     NEXT round's constants into its sums (the reference adds them at the top of a round);
//   * the partial rounds use the plain form (S-box on lane 0, full MDS) rather than the sparse-factor form;
//   * state words are kept as arbitrary u64 representatives and made canonical only when a digest is stored.
#pragma once
#include "field_gl.cuh"
#include "hash_tables.h"

namespace p2g {

#ifdef __CUDACC__

// the including translation unit (hash.cu) defines:  __constant__ PoseidonGlTables c_pgl;

__device__ __forceinline__ u64 pgl_sbox(u64 x) {
#ifdef P2G_GL_SBOX_NO_SQUARE  // tools/microbench: the general product for x^2 and x^4 (27.32k cycles per warp-permutation)
    u64 x2 = gl_mul(x, x);
    u64 x4 = gl_mul(x2, x2);
#else  // dedicated squaring (field_gl.cuh, gl_sqr128): 26.47k
    u64 x2 = gl_sqr(x);
    u64 x4 = gl_sqr(x2);
#endif
    u64 x3 = gl_mul(x, x2);
    return gl_mul(x3, x4);
}

// acc_lo + acc_hi * 2^32 (acc_lo < 2^53, acc_hi < 2^43: the reassembled limbs of one MDS output)
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

// Circulant part of the MDS layer on one 22-bit LIMB of the state: y[r] = sum_i circ[i] * x[(i + r) % 12] (+ add[r]) with
// circ = (17,15,41,16,2,28,13,13,39,18,34,20) (poseidon_goldilocks.rs:301), computed WITHOUT multiplications.
// The direct product needs 144 IMAD.WIDE per 32-bit half; instead the length-12 cyclic correlation is split as
// 12 = 3 x 4: a 4-point DFT along t of x[j + 3t] (integers and Gaussian integers, adds only), three 3-point "twisted"
// cyclic products modulo z^3 - lambda (lambda = 1, i, -1; the fourth is the conjugate of the second because the input
// is real) against the transformed circulant, and the inverse 4-point DFT.  The transformed circulant, with the 1/4
// (and 1/2 for the complex pair) of the inverse folded in, is all small:
//   f=0: (16, 32, 16)   f=2: (-1, -8, 2)   f=1: (2+i, -4-i, 16-i)
// so every product is a shift, an add or a negation.  The reference reaches the same product through
// mds_multiply_freq (:956-1088); this derivation and its constants were re-done from the circulant and checked against
// the direct product (tests: Poseidon known answers).
// Arithmetic is plain 32-bit wrap-around: inputs < 2^22, outputs < 264 * 2^22 < 2^31, intermediates may wrap (every
// operation is a ring operation mod 2^32 and the true result fits).  There are no carries, so each add / shift-add can
// issue on either integer pipe (IADD3 / LEA on the half-rate ALU pipe, IMAD.IADD / IMAD on the FMA pipe) and ptxas
// balances them; an earlier formulation on 32-bit halves with 64-bit accumulators kept its carry chains on the ALU pipe
// (ncu: ALU 82% busy, FMA 30%) and was 17% slower per layer (profiles/r1_leaf_sponge_cost_model.json).
__device__ __forceinline__ void pgl_circ_limb(const u32 x[12], u32 y[12], const u32 add[12]) {
    u32 f0[3], f2[3], re[3], im[3];
#pragma unroll
    for (int j = 0; j < 3; j++) {
        u32 x0 = x[j], x1 = x[j + 3], x2 = x[j + 6], x3 = x[j + 9];
        u32 p02 = x0 + x2, p13 = x1 + x3;
        f0[j] = p02 + p13;
        f2[j] = p02 - p13;
        re[j] = x0 - x2;
        im[j] = x1 - x3;
    }
    u32 s = f0[0] + f0[1] + f0[2];
    u32 y0[3] = {(s + f0[2]) << 4, (s + f0[0]) << 4, (s + f0[1]) << 4};
    u32 y2[3] = {(f2[2] << 3) - f2[0] - (f2[1] << 1), 0u - (f2[0] << 3) - f2[1] - (f2[2] << 1), (f2[0] << 1) - (f2[1] << 3) - f2[2]};
    u32 r0 = re[0], i0 = im[0], r1 = re[1], i1 = im[1], r2 = re[2], i2 = im[2];
    u32 t_re = (r1 << 4) + i1 - (r2 << 2) + i2;
    u32 t_im = (i1 << 4) - r1 - r2 - (i2 << 2);
    u32 y1r[3], y1i[3];
    y1r[0] = (r0 << 1) - i0 - t_im;
    y1i[0] = r0 + (i0 << 1) + t_re;
    y1r[1] = i0 - (r0 << 2) + (r1 << 1) - i1 + r2 - (i2 << 4);
    y1i[1] = r1 - r0 - (i0 << 2) + (i1 << 1) + (r2 << 4) + i2;
    y1r[2] = (r0 << 4) + i0 - (r1 << 2) + i1 + (r2 << 1) - i2;
    y1i[2] = (i0 << 4) - r0 - r1 - (i1 << 2) + r2 + (i2 << 1);
#pragma unroll
    for (int j = 0; j < 3; j++) {
        u32 p = y0[j] + y2[j], m = y0[j] - y2[j];
        y[j] = p + y1r[j] + add[j];
        y[j + 3] = m + y1i[j] + add[j + 3];
        y[j + 6] = p - y1r[j] + add[j + 6];
        y[j + 9] = m - y1i[j] + add[j + 9];
    }
}

// out = MDS * s + next round constants.  MDS = circ(17,15,41,16,2,28,13,13,39,18,34,20) + diag(8,0,...)
// (poseidon_goldilocks.rs:301-302; row r of the product is sum_i circ[i] * s[(i+r)%12], :547-557), applied to three
// limbs (22 + 22 + 20 bits) of every lane; out = l0 + l1 * 2^22 + l2 * 2^44 is reassembled with two IMAD.WIDE into the
// (acc_lo, acc_hi) pair pgl_fold takes:  acc_lo = l1 * 2^22 + l0 (< 2^53),  acc_hi = l2 * 2^12 (< 2^43).
// rcl = the next round's constants pre-split into the same limbs (hash_tables.h); they and the diagonal term ride as
// third addends of the last adds.
__device__ __forceinline__ void pgl_mds(u64 s[12], const u32 (*rcl)[12]) {
    u32 l0[12], l1[12], l2[12], c0[12], c1[12], c2[12];
#pragma unroll
    for (int i = 0; i < 12; i++) {
        u32 lo = (u32)s[i], hi = (u32)(s[i] >> 32);
        l0[i] = lo & 0x3fffffu;
        l1[i] = __funnelshift_r(lo, hi, 22) & 0x3fffffu;
        l2[i] = hi >> 12;
        c0[i] = rcl[0][i];
        c1[i] = rcl[1][i];
        c2[i] = rcl[2][i];
    }
    c0[0] += l0[0] << 3;  // the diagonal entry 8
    c1[0] += l1[0] << 3;
    c2[0] += l2[0] << 3;
    u32 y0[12], y1[12], y2[12];
    pgl_circ_limb(l0, y0, c0);
    pgl_circ_limb(l1, y1, c1);
    pgl_circ_limb(l2, y2, c2);
#pragma unroll
    for (int r = 0; r < 12; r++) {
        u64 al = (u64)y1[r] * 0x400000u + y0[r];
        u64 ah = (u64)y2[r] * 0x1000u;
        s[r] = pgl_fold(al, ah);
    }
}


// Round structure (poseidon_goldilocks.rs:927-948, the reference's `poseidon_naive`): 30 rounds of
// [add constants, S-box, MDS]; rounds 0-3 and 26-29 apply the S-box to all 12 lanes, rounds 4-25 to lane 0 only.
// The reference's production form (:912-922) replaces the 22 partial rounds by a dense pre-matrix plus sparse
// factors with 64-bit constants (23 full multiplications a round); with the multiplication-free MDS above the plain
// form is cheaper on this machine (measured per warp: S-box 122 + MDS 606 cycles a partial round against ~810, and no pre-matrix) and needs no derived
// tables.  Both forms compute the same permutation (the oracle checks naive == fast, tests check this == oracle).
//
// Code-size discipline: the permutation is ~25k dynamic instructions, and a warp that strays from its neighbours
// misses the instruction cache on every line (first ncu capture: 13 of 18 stalled warps per issue were
// `no_instruction`).  So there is ONE round body for all 30 rounds (the S-box of lanes 1..11 sits behind a uniform
// branch), a real loop, and P2G_ROUND_SYNC keeps the warps of a CTA on the same lines.
#ifndef P2G_ROUND_SYNC
#define P2G_ROUND_SYNC() __syncthreads()
#endif

// ---- partial rounds in the limb domain ------------------------------------------------------------------------------
// In rounds 4-25 only lane 0 meets the S-box; lanes 1..11 go from one MDS layer straight into the next.  pgl_mds pays
// for that hand-over twice per lane and round: 2 IMAD.WIDE + an 8-instruction carry-chain fold to rebuild the u64, then
// 5 instructions to cut it into limbs again.  Here lanes 1..11 STAY in limbs for rounds 4..25: the MDS sums (< 2^31)
// are brought back to limb size by a carry-less pass
//     c0 = y0 >> 22, t1 = y1 + c0, c1 = t1 >> 22, t2 = y2 + c1, top = t2 >> 20   (top < 2^9: the bits at and above 2^64)
//     l0 = (y0 & M22) - top,  l1 = (t1 & M22) + top * 2^10,  l2 = t2 & M20        (2^64 = 2^32 - 1 mod p)
// (10 plain 32-bit instructions, either integer pipe), leaving l0 in (-2^9, 2^22), l1 < 2^22 + 2^19, l2 < 2^20: the next
// layer's sums stay below 264 * (2^22 + 2^19) + 2^23 < 2^31.  A slightly negative l0 could make a limb-0 sum negative,
// so every limb-0 sum carries a bias of 2^18 (> 264 * 2^9).
// Round constants: lanes 1..11 see only linear maps in these rounds, so their constants are pushed forward through
// the MDS layers on the host (poseidon_tables.cpp; the reference's fast form does the same thing backwards,
// poseidon_goldilocks.rs:304-492): with x = x' + delta, delta_0 = 0,  M S(x) + c = M S(x') + (M delta + c), so the
// loop adds a constant to lane 0 only (rclb, with the bias taken out mod p and put back on limb 0) plus the bare bias
// on lanes 1..11, and the accumulated delta (minus the biases) is added once on the way out (rc_limb_exit).
// Lane 0 is rebuilt and cut every round as before (it feeds the S-box).  After round 25 the limbs are joined once
// ( + 2^10 before the join so it never sees a negative limb, taken out again inside rc_limb_exit) for the full rounds
// 26..29.  Same permutation bit for bit (tests: known answers, oracle parity).
__device__ __forceinline__ void pgl_split3(u64 v, u32 &a, u32 &b, u32 &c) {
    u32 lo = (u32)v, hi = (u32)(v >> 32);
    a = lo & 0x3fffffu;
    b = __funnelshift_r(lo, hi, 22) & 0x3fffffu;
    c = hi >> 12;
}
__device__ __forceinline__ u64 pgl_join3(u32 y0, u32 y1, u32 y2) {
    return pgl_fold((u64)y1 * 0x400000u + y0, (u64)y2 * 0x1000u);
}

__device__ __forceinline__ void pgl_partial_rounds_limbs(u64 s[12]) {
    u32 l0[12], l1[12], l2[12];
#pragma unroll
    for (int i = 1; i < 12; i++) pgl_split3(s[i], l0[i], l1[i], l2[i]);
    u64 s0 = s[0];
#pragma unroll 1
    for (int r = 0; r < P2G_PGL_LIMB_ROUNDS; r++) {
        P2G_ROUND_SYNC();
        s0 = pgl_sbox(s0);
        pgl_split3(s0, l0[0], l1[0], l2[0]);
        // constants: lane 0 gets this round's (propagated, biased) constant + the diagonal entry 8; lanes 1..11 only the
        // bias on limb 0 (an immediate) -- their constants were pushed through the linear layers on the host
        u32 c0[12], c1[12], c2[12];
        c0[0] = c_pgl.rclb[r][0] + (l0[0] << 3);
        c1[0] = c_pgl.rclb[r][1] + (l1[0] << 3);
        c2[0] = c_pgl.rclb[r][2] + (l2[0] << 3);
#pragma unroll
        for (int i = 1; i < 12; i++) {
            c0[i] = P2G_PGL_LIMB_BIAS;
            c1[i] = 0;
            c2[i] = 0;
        }
        u32 y0[12], y1[12], y2[12];
        pgl_circ_limb(l0, y0, c0);
        pgl_circ_limb(l1, y1, c1);
        pgl_circ_limb(l2, y2, c2);
        s0 = pgl_join3(y0[0], y1[0], y2[0]);
#pragma unroll
        for (int i = 1; i < 12; i++) {
            u32 t1 = y1[i] + (y0[i] >> 22);
            u32 t2 = y2[i] + (t1 >> 22);
            u32 top = t2 >> 20;
            l0[i] = (y0[i] & 0x3fffffu) - top;
            l1[i] = (t1 & 0x3fffffu) + (top << 10);
            l2[i] = t2 & 0xfffffu;
        }
    }
    s[0] = s0;
#pragma unroll
    for (int i = 1; i < 12; i++) s[i] = gl_add_c(pgl_join3(l0[i] + 1024u, l1[i], l2[i]), c_pgl.rc_limb_exit[i]);
}

__device__ __forceinline__ void poseidon_gl_permute(u64 s[12]) {
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = gl_add_c(s[i], c_pgl.rc0[i]);
#ifdef P2G_PGL_PLAIN_PARTIAL
#pragma unroll 1
    for (int r = 0; r < 30; r++) {
        P2G_ROUND_SYNC();
        s[0] = pgl_sbox(s[0]);
        if (r < 4 || r >= 26) {
#pragma unroll
            for (int i = 1; i < 12; i++) s[i] = pgl_sbox(s[i]);
        }
        pgl_mds(s, c_pgl.rcl[r + 1]);  // the next round's constants ride along; rcl[30] = 0
    }
#else
    // the full rounds 0..3 and 26..29 share one body (two trips through the same loop); the 22 partial rounds run in
    // the limb domain in between
#pragma unroll 1
    for (int phase = 0; phase < 2; phase++) {
        const int r_end = phase ? 30 : 4;
#pragma unroll 1
        for (int r = phase ? 26 : 0; r < r_end; r++) {
            P2G_ROUND_SYNC();
#pragma unroll
            for (int i = 0; i < 12; i++) s[i] = pgl_sbox(s[i]);
            pgl_mds(s, c_pgl.rcl[r + 1]);  // the next round's constants ride along; rcl[30] = 0
        }
        if (phase == 0) pgl_partial_rounds_limbs(s);
    }
#endif
}

#endif  // __CUDACC__
}  // namespace p2g
