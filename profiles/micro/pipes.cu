// micro-benchmark: issue rates of IMAD, IMAD.WIDE, ALU ops and their mixes on sm_100a (lanes per clock per SM).
// Dependent rings (each op reads a neighbour chain) keep ptxas from strength-reducing the sequences.
// nvcc -gencode arch=compute_100a,code=sm_100a -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
typedef unsigned int u32;
#define ITERS 2048
#define IMAD(D, A, B) asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(D) : "r"(A), "r"(B))
#define WIDE(W, A, B) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(W) : "r"(A), "r"(B))
#define ADD(D, A) asm("add.u32 %0, %0, %1;" : "+r"(D) : "r"(A))
#define XOR(D, A) asm("xor.b32 %0, %0, %1;" : "+r"(D) : "r"(A))
template <int MODE>
__global__ void __launch_bounds__(256) kern(u32 *out, u32 m) {
    u32 x0 = threadIdx.x, x1 = x0 * 3 + 1, x2 = x0 * 5 + 2, x3 = x0 * 7 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    u64 w0 = x0, w1 = x1, w2 = x2, w3 = x3, w4 = x4, w5 = x5, w6 = x6, w7 = x7;
#pragma unroll 1
    for (int i = 0; i < ITERS; i++) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            if (MODE == 0) {          // 8 IMAD, ring of 8
                IMAD(x0, x1, x2); IMAD(x1, x2, x3); IMAD(x2, x3, x4); IMAD(x3, x4, x5); IMAD(x4, x5, x6); IMAD(x5, x6, x7); IMAD(x6, x7, x0); IMAD(x7, x0, x1);
            } else if (MODE == 1) {   // 8 IMAD.WIDE, ring of 8
                WIDE(w0, (u32)w1, (u32)w2); WIDE(w1, (u32)w2, (u32)w3); WIDE(w2, (u32)w3, (u32)w4); WIDE(w3, (u32)w4, (u32)w5);
                WIDE(w4, (u32)w5, (u32)w6); WIDE(w5, (u32)w6, (u32)w7); WIDE(w6, (u32)w7, (u32)w0); WIDE(w7, (u32)w0, (u32)w1);
            } else if (MODE == 2) {   // 8 ALU (add / xor alternating, ring)
                ADD(x0, x1); XOR(x1, x2); ADD(x2, x3); XOR(x3, x4); ADD(x4, x5); XOR(x5, x6); ADD(x6, x7); XOR(x7, x0);
            } else if (MODE == 3) {   // 4 IMAD + 4 ALU
                IMAD(x0, x1, x2); XOR(x1, x2); IMAD(x2, x3, x4); ADD(x3, x4); IMAD(x4, x5, x6); XOR(x5, x6); IMAD(x6, x7, x0); ADD(x7, x0);
            } else if (MODE == 4) {   // 4 IMAD.WIDE + 4 ALU
                WIDE(w0, (u32)w1, x1); XOR(x1, x2); WIDE(w1, (u32)w2, x3); ADD(x3, x4); WIDE(w2, (u32)w3, x5); XOR(x5, x6); WIDE(w3, (u32)w0, x7); ADD(x7, x0);
            } else if (MODE == 5) {   // 6 IMAD.WIDE + 2 ALU
                WIDE(w0, (u32)w1, x1); WIDE(w1, (u32)w2, x3); WIDE(w2, (u32)w3, x5); XOR(x1, x2); WIDE(w3, (u32)w4, x7); WIDE(w4, (u32)w5, x1); WIDE(w5, (u32)w0, x3); ADD(x3, x4);
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 ^ x1 ^ x2 ^ x3 ^ x4 ^ x5 ^ x6 ^ x7 ^ (u32)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7) ^ (u32)((w0 ^ w3 ^ w5) >> 32);
}
template <int MODE>
void run(const char *name, u32 *d, int sms, float clk_ghz) {
    const int blocks = sms * 8;
    kern<MODE><<<blocks, 256>>>(d, 3);
    cudaDeviceSynchronize();
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    kern<MODE><<<blocks, 256>>>(d, 3);
    cudaEventRecord(b);
    cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double instr = (double)blocks * 256 * ITERS * 32;
    printf("{\"micro\": \"%s\", \"ms\": %.3f, \"lane_instr_per_clk_per_sm_at_nominal_clock\": %.1f}\n", name, ms, instr / (ms * 1e-3) / (clk_ghz * 1e9) / sms);
}
int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const float ghz = khz / 1e6f;
    u32 *d; cudaMalloc(&d, (size_t)p.multiProcessorCount * 8 * 256 * 4);
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_ghz_nominal\": %.3f}\n", p.name, p.multiProcessorCount, ghz);
    run<0>("8xIMAD", d, p.multiProcessorCount, ghz);
    run<1>("8xIMAD.WIDE", d, p.multiProcessorCount, ghz);
    run<2>("8xALU", d, p.multiProcessorCount, ghz);
    run<3>("4xIMAD+4xALU", d, p.multiProcessorCount, ghz);
    run<4>("4xIMAD.WIDE+4xALU", d, p.multiProcessorCount, ghz);
    run<5>("6xIMAD.WIDE+2xALU", d, p.multiProcessorCount, ghz);
    return 0;
}
