// Poseidon2 over BabyBear, width 16, x^7, 4 + 13 + 4 rounds -- one permutation per thread, state in registers,
// Montgomery form throughout.
//
// Replaces the external p3_poseidon2::Poseidon2::permute_mut the reference calls from
// plonky2/src/hash/poseidon2_babybear.rs:69-84,150-159; the permutation is the one restated in-tree by
// plonky2/src/gates/poseidon2_babybear.rs:609-671 (external layer :804-832 with apply_mat4 :903-917, internal
// layer :787-802: multiply by 2^-32, then s0 <- sum_{i>=1} s_i - s0, s_i <- sum_all + 2^shift_i * s_i).
// In Montgomery form the 2^-32 factor is one Montgomery reduction of the 64-bit sum, as in p3-baby-bear.
#pragma once
#include "field_bb.cuh"
#include "hash_tables.h"

namespace p2g {

#ifdef __CUDACC__

// the including translation unit (hash.cu) defines:  __constant__ Poseidon2BbTables c_p2bb;

__device__ __forceinline__ u32 p2bb_sbox(u32 x) {
    u32 x2 = bb_mul(x, x);
    u32 x3 = bb_mul(x2, x);
    u32 x4 = bb_mul(x2, x2);
    return bb_mul(x3, x4);
}

__device__ __forceinline__ void p2bb_mat4(u32 &x0, u32 &x1, u32 &x2, u32 &x3) {
    u32 t01 = bb_add(x0, x1);
    u32 t23 = bb_add(x2, x3);
    u32 t0123 = bb_add(t01, t23);
    u32 t01123 = bb_add(t0123, x1);
    u32 t01233 = bb_add(t0123, x3);
    u32 n3 = bb_add(t01233, bb_add(x0, x0));
    u32 n1 = bb_add(t01123, bb_add(x2, x2));
    u32 n0 = bb_add(t01123, t01);
    u32 n2 = bb_add(t01233, t23);
    x0 = n0;
    x1 = n1;
    x2 = n2;
    x3 = n3;
}

__device__ __forceinline__ void p2bb_external(u32 s[16]) {
#pragma unroll
    for (int i = 0; i < 16; i += 4) p2bb_mat4(s[i], s[i + 1], s[i + 2], s[i + 3]);
    u32 sums[4];
#pragma unroll
    for (int k = 0; k < 4; k++) sums[k] = bb_add(bb_add(s[k], s[4 + k]), bb_add(s[8 + k], s[12 + k]));
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = bb_add(s[i], sums[i % 4]);
}

__device__ __forceinline__ void p2bb_internal(u32 s[16]) {
    const int shifts[15] = {0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 10, 11, 12, 13, 15};
    u64 part = 0;
#pragma unroll
    for (int i = 1; i < 16; i++) part += s[i];
    u64 full = part + s[0];
    u32 n0 = bb_monty_reduce(part + (u64)(P2G_BB_P - s[0]));
#pragma unroll
    for (int i = 1; i < 16; i++) s[i] = bb_monty_reduce(full + ((u64)s[i] << shifts[i - 1]));
    s[0] = n0;
}

__device__ __forceinline__ void poseidon2_bb_permute(u32 s[16]) {
    p2bb_external(s);
#pragma unroll 1
    for (int r = 0; r < 4; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = p2bb_sbox(bb_add(s[i], c_p2bb.ext_rc[r][i]));
        p2bb_external(s);
    }
#pragma unroll 1
    for (int r = 0; r < 13; r++) {
        s[0] = p2bb_sbox(bb_add(s[0], c_p2bb.int_rc[r]));
        p2bb_internal(s);
    }
#pragma unroll 1
    for (int r = 4; r < 8; r++) {
#pragma unroll
        for (int i = 0; i < 16; i++) s[i] = p2bb_sbox(bb_add(s[i], c_p2bb.ext_rc[r][i]));
        p2bb_external(s);
    }
}

#endif  // __CUDACC__
}  // namespace p2g
