// Second integer-pipe probe: costs of IMAD.WIDE.U32 (immediate and register multiplier), IMAD (32-bit), IMAD.HI,
// IADD3 carry pairs and their mixes, written so that ptxas keeps the accumulate forms (checked with cuobjdump).
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITER 512

template <int OP>
__global__ void __launch_bounds__(256) probe(u32 *out, u32 seed, u32 m0, u32 m1, u32 m2) {
    u64 acc[8];
    u32 x[8], y[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        x[i] = seed * (threadIdx.x + 1) + i * 77;
        y[i] = x[i] ^ 0x5555;
        acc[i] = x[i];
    }
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int kk = 0; kk < 32; kk++) {
            const int k = kk & 7;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                if (OP == 0) acc[j] += (u64)x[(j + k) & 7] * (u32)(17 + 2 * k);          // IMAD.WIDE imm
                if (OP == 1) acc[j] += (u64)x[(j + k) & 7] * (k & 1 ? m0 : m1);          // IMAD.WIDE reg
                if (OP == 2) y[j] += x[(j + k) & 7] * (u32)(17 + 2 * k);                  // IMAD 32
                if (OP == 3) y[j] += __umulhi(x[(j + k) & 7], (k & 1 ? m0 : m1));         // IMAD.HI
                if (OP == 4) { acc[j] += (u64)x[(j + k) & 7] * (u32)(17 + 2 * k); y[j] = (y[j] + x[(j + k + 1) & 7]) ^ m0; }  // WIDE + IADD3 + LOP3
                if (OP == 5) { acc[j] += (u64)x[(j + k) & 7] * (u32)(17 + 2 * k); y[j] += x[(j + k + 1) & 7]; }             // WIDE + IADD3
                if (OP == 6) { y[j] += x[(j + k + 1) & 7]; }                                                                  // IADD3 only
                if (OP == 7) { acc[j] += ((u64)x[(j + k) & 7] << 32) | y[(j + k) & 7]; }                                      // 64-bit add (IADD3 + IADD3.X)
                if (OP == 9) { acc[j] += acc[(j + k + 1) & 7] ^ 0; }                                                           // 64-bit reg add (ptxas picks IADD3.X or IMAD.X)
                if (OP == 10) {  // same with a third, opaque-zero operand on the high word: must be IADD3.X
                    u32 al = (u32)acc[j], ah = (u32)(acc[j] >> 32), bl = (u32)acc[(j + k + 1) & 7], bh = (u32)(acc[(j + k + 1) & 7] >> 32);
                    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %3;" : "+r"(al), "+r"(ah) : "r"(bl), "r"(bh + m2));
                    acc[j] = ((u64)ah << 32) | al;
                }
                if (OP == 11) { y[j] = x[(j + k + 1) & 7] - y[(j+k)&7]; }                                                      // 2-input sub (IADD3 or IMAD.IADD)
                if (OP == 12) { y[j] = x[(j + k + 1) & 7] - y[(j+k)&7] + m2; }                                                 // 3-input (IADD3)
                if (OP == 8) { acc[j] += (u64)x[(j + k) & 7] * (u32)(17 + 2 * k); acc[(j + 1) & 7] += ((u64)y[j] << 32) | x[j]; }  // WIDE + 64-bit add
            }
        }
#pragma unroll
        for (int j = 0; j < 8; j++) x[j] ^= (u32)(acc[j] >> 13) + y[j];
    }
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= x[i] ^ y[i] ^ (u32)acc[i] ^ (u32)(acc[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char *name, double ops_per_inner) {
    int sms;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    u32 *out;
    const int threads = 256, blocks = sms * 8;
    cudaMalloc(&out, (size_t)threads * blocks * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<OP><<<blocks, threads>>>(out, 12345, 0x9e3779b9u, 0x7f4a7c15u, 0u);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, 12345, 0x9e3779b9u, 0x7f4a7c15u, 0u);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double slots = (double)blocks * (threads / 32) * ITER * 256.0;  // inner-statement executions per warp
    double cyc_per_slot = (ms * 1e-3 * 1.9625e9) / (slots / sms / 4);  // SMSP cycles per inner statement
    printf("%-28s %8.3f ms  %.2f SMSP-cycles per inner statement (%.0f counted ops)\n", name, ms, cyc_per_slot, ops_per_inner);
    cudaFree(out);
}

int main() {
    run<0>("IMAD.WIDE imm", 1);
    run<1>("IMAD.WIDE reg", 1);
    run<2>("IMAD32 imm", 1);
    run<3>("IMAD.HI reg", 1);
    run<6>("IADD3", 1);
    run<7>("add64 (IADD3+IADD3.X)", 2);
    run<5>("IMAD.WIDE + IADD3", 2);
    run<4>("IMAD.WIDE + IADD3 + LOP3", 3);
    run<8>("IMAD.WIDE + add64", 3);
    run<9>("add64 reg+reg (ptxas mix)", 2);
    run<10>("add64 with opaque zero", 2);
    run<11>("sub32 2-input", 1);
    run<12>("sub32 3-input", 1);
    return 0;
}
