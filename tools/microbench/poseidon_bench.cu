// Stand-alone throughput harness for the Goldilocks Poseidon permutation kernel variants.
// Builds against the product headers (csrc/) so what is timed is the shipped code.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -I../../plonky2_goldibear_b200/csrc \
//        [-DVARIANT_FLAGS...] poseidon_bench.cu ../../plonky2_goldibear_b200/csrc/poseidon_tables.cpp -o poseidon_bench
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "p2g_internal.h"
namespace p2g {
__constant__ PoseidonGlTables c_pgl;
__constant__ Poseidon2BbTables c_p2bb;
}
#include "poseidon2_bb.cuh"
#include "poseidon_gl.cuh"
using namespace p2g;

#ifndef BLOCK
#define BLOCK 128
#endif
#ifndef MINB
#define MINB 1
#endif

__global__ void __launch_bounds__(BLOCK, MINB) gl_chain(const u64 *in, u64 *out, int chain) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = in[t * 12 + i];
#pragma unroll 1
    for (int k = 0; k < chain; k++) poseidon_gl_permute(s);
#pragma unroll
    for (int i = 0; i < 12; i++) out[t * 12 + i] = gl_canon(s[i]);
}
// component probes: 30 x (12 S-boxes) and 30 x MDS layers, same loop shape as the permutation
__global__ void __launch_bounds__(BLOCK, MINB) gl_sbox_chain(const u64 *in, u64 *out, int chain) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = in[t * 12 + i];
#pragma unroll 1
    for (int k = 0; k < chain * 30; k++) {
#pragma unroll
        for (int i = 0; i < 12; i++) s[i] = pgl_sbox(s[i]);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) out[t * 12 + i] = gl_canon(s[i]);
}
__global__ void __launch_bounds__(BLOCK, MINB) gl_mds_chain(const u64 *in, u64 *out, int chain) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u64 s[12];
#pragma unroll
    for (int i = 0; i < 12; i++) s[i] = in[t * 12 + i];
#pragma unroll 1
    for (int k = 0; k < chain; k++) {
#pragma unroll 1
        for (int r = 0; r < 30; r++) pgl_mds(s, c_pgl.rcl[r + 1]);
    }
#pragma unroll
    for (int i = 0; i < 12; i++) out[t * 12 + i] = gl_canon(s[i]);
}

__global__ void __launch_bounds__(BLOCK, MINB) bb_chain(const u32 *in, u32 *out, int chain) {
    size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    u32 s[16];
#pragma unroll
    for (int i = 0; i < 16; i++) s[i] = in[t * 16 + i];
#pragma unroll 1
    for (int k = 0; k < chain; k++) poseidon2_bb_permute(s);
#pragma unroll
    for (int i = 0; i < 16; i++) out[t * 16 + i] = s[i];
}

int main(int argc, char **argv) {
    int chain = argc > 1 ? atoi(argv[1]) : 17;
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    static PoseidonGlTables gl;
    static Poseidon2BbTables bb;
    build_poseidon_gl_tables(&gl);
    build_poseidon2_bb_tables(&bb);
    cudaMemcpyToSymbol(c_pgl, &gl, sizeof gl);
    cudaMemcpyToSymbol(c_p2bb, &bb, sizeof bb);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, gl_chain, BLOCK, 0);
    size_t threads = (size_t)sms * occ * BLOCK * 8;  // 8 full waves
    u64 *d_in, *d_out;
    cudaMalloc(&d_in, threads * 96);
    cudaMalloc(&d_out, threads * 96);
    u64 *h = (u64 *)malloc(threads * 96);
    for (size_t i = 0; i < threads * 12; i++) h[i] = (i % 12);  // KAT input 0..11 in every thread
    cudaMemcpy(d_in, h, threads * 96, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    gl_chain<<<threads / BLOCK, BLOCK>>>(d_in, d_out, 1);
    cudaMemcpy(h, d_out, 96, cudaMemcpyDeviceToHost);
    bool ok = h[0] == 0xd64e1e3efc5b8e9eULL && h[11] == 0x5c0a27fcb0e1459bULL;  // poseidon_goldilocks.rs:1175-1178
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        gl_chain<<<threads / BLOCK, BLOCK>>>(d_in, d_out, chain);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    cudaError_t err = cudaGetLastError();
    printf("GL  block %d occ %d blocks/SM KAT %s: %.3f ms for %zu perms -> %.1f M perm/s (%s)\n", BLOCK, occ, ok ? "ok" : "FAIL",
           best, threads * chain, threads * (double)chain / best * 1e-3, cudaGetErrorString(err));
    {
        const double warp_slots = (double)sms * 4 * 1.9625e9;  // SMSP cycles per second at the usual boost clock
        for (int which = 0; which < 2; which++) {
            float b2 = 1e30f;
            for (int rep = 0; rep < 3; rep++) {
                cudaEventRecord(e0);
                if (which == 0) gl_sbox_chain<<<threads / BLOCK, BLOCK>>>(d_in, d_out, chain);
                else gl_mds_chain<<<threads / BLOCK, BLOCK>>>(d_in, d_out, chain);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (ms < b2) b2 = ms;
            }
            double calls = (double)threads / 32 * chain * 30 * (which == 0 ? 12 : 1);
            printf("GL  %s: %.3f ms -> %.1f SMSP cycles per warp-%s\n", which == 0 ? "sbox chain" : "mds chain", b2,
                   b2 * 1e-3 * warp_slots / calls, which == 0 ? "sbox" : "mds");
        }
        printf("GL  permutation: %.0f SMSP cycles per warp-permutation\n", best * 1e-3 * warp_slots / ((double)threads / 32 * chain));
    }
    // BabyBear
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, bb_chain, BLOCK, 0);
    size_t bthreads = (size_t)sms * occ * BLOCK * 8;
    u32 *b_in, *b_out;
    cudaMalloc(&b_in, bthreads * 64);
    cudaMalloc(&b_out, bthreads * 64);
    cudaMemset(b_in, 1, bthreads * 64);
    bb_chain<<<bthreads / BLOCK, BLOCK>>>(b_in, b_out, 1);
    best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        bb_chain<<<bthreads / BLOCK, BLOCK>>>(b_in, b_out, chain);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    u32 hb[16];
    bb_chain<<<bthreads / BLOCK, BLOCK>>>(b_in, b_out, 3);
    cudaMemcpy(hb, b_out, 64, cudaMemcpyDeviceToHost);  // compare across variants: three chained permutations of 0x01010101 words
    printf("BB  block %d occ %d blocks/SM: %.3f ms for %zu perms -> %.1f M perm/s   out %08x %08x %08x %08x (%s)\n", BLOCK, occ, best,
           bthreads * chain, bthreads * (double)chain / best * 1e-3, hb[0], hb[1], hb[14], hb[15], cudaGetErrorString(cudaGetLastError()));
    return ok ? 0 : 1;
}
