// Integer-pipe throughput probe for sm_100a: how many warp-instructions per cycle per SM for the opcodes the
// Goldilocks / BabyBear arithmetic is made of.  Each kernel runs ITER iterations of 8 independent chains of one op.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_pipes int_pipes.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITER 4096

template <int OP>
__global__ void probe(u32 *out, u32 seed) {
    u32 a[8], b[8];
    u64 w[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        a[i] = seed + threadIdx.x * 8 + i;
        b[i] = seed * 3 + i;
        w[i] = ((u64)a[i] << 32) | b[i];
    }
    u32 c = seed | 1;
#pragma unroll 1
    for (int it = 0; it < ITER; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) {  // IMAD.WIDE.U32 with 64-bit accumulate
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 1) & 7]), "r"(c));
            } else if (OP == 1) {  // IMAD (32-bit lo)
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b[i]));
            } else if (OP == 2) {  // IMAD.HI
                asm volatile("mad.hi.u32 %0, %0, %1, %0;" : "+r"(a[i]) : "r"(c));
            } else if (OP == 3) {  // IADD3
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 4) {  // LOP3
                asm volatile("xor.b32 %0, %0, %1;" : "+r"(a[i]) : "r"(b[i]));
            } else if (OP == 5) {  // 64-bit add = IADD3 + IADD3.X
                asm volatile("add.u64 %0, %0, %1;" : "+l"(w[i]) : "l"((u64)b[i] << 7 | c));
            } else if (OP == 6) {  // mix: IMAD.WIDE + IADD3 (1:1)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 1) & 7]), "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(c));
            } else if (OP == 7) {  // mix: IMAD.WIDE + 2 IADD3
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(w[i]) : "r"((u32)w[(i + 1) & 7]), "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(c));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(a[i]) : "r"(c));
            } else if (OP == 8) {  // add with carry chain (IADD3 + IADD3.X pairs via cc)
                asm volatile("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, %1, %2;" : "+r"(a[i]), "+r"(b[i]) : "r"(c));
            } else if (OP == 9) {  // SEL/ISETP pair
                asm volatile("{.reg .pred p; setp.ge.u32 p, %0, %1; selp.u32 %0, %1, %2, p;}" : "+r"(a[i]) : "r"(b[i]), "r"(c));
            } else if (OP == 10) {  // mix IMAD(lo) + IADD3
                asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(c), "r"(b[i]));
                asm volatile("add.u32 %0, %0, %1;" : "+r"(b[i]) : "r"(c));
            } else if (OP == 11) {  // FFMA for reference
                float f = __uint_as_float(a[i]);
                asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f) : "f"(1.0001f), "f"(0.5f));
                a[i] = __float_as_uint(f);
            }
        }
    }
    u32 r = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i] ^ (u32)w[i] ^ (u32)(w[i] >> 32);
    out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

template <int OP>
void run(const char *name, int instr_per_slot) {
    int dev_sms;
    cudaDeviceGetAttribute(&dev_sms, cudaDevAttrMultiProcessorCount, 0);
    int clock_khz;
    cudaDeviceGetAttribute(&clock_khz, cudaDevAttrClockRate, 0);
    u32 *out;
    const int threads = 256, blocks = dev_sms * 8;
    cudaMalloc(&out, (size_t)threads * blocks * 4);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    probe<OP><<<blocks, threads>>>(out, 12345);
    cudaDeviceSynchronize();
    cudaEventRecord(e0);
    probe<OP><<<blocks, threads>>>(out, 12345);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    double warp_instr = (double)blocks * (threads / 32) * ITER * 8.0 * instr_per_slot;
    double per_sm_per_s = warp_instr / dev_sms / (ms * 1e-3);
    printf("%-34s %8.3f ms  %7.2f G warp-instr/s/SM  = %5.2f warp-instr/clk/SM @%.0f MHz nominal\n", name, ms,
           per_sm_per_s * 1e-9, per_sm_per_s / (clock_khz * 1e3), clock_khz * 1e-3);
    cudaFree(out);
}

int main() {
    run<11>("FFMA", 1);
    run<0>("IMAD.WIDE.U32 (64b acc)", 1);
    run<1>("IMAD lo", 1);
    run<2>("IMAD.HI", 1);
    run<3>("IADD3", 1);
    run<4>("LOP3", 1);
    run<5>("add.u64 (IADD3+IADD3.X)", 2);
    run<8>("add.cc/addc pair", 2);
    run<9>("ISETP+SEL", 2);
    run<6>("mix IMAD.WIDE + IADD3", 2);
    run<7>("mix IMAD.WIDE + 2 IADD3", 3);
    run<10>("mix IMAD lo + IADD3", 2);
    return 0;
}
