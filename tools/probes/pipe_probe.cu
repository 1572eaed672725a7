// Issue-rate probe for the integer instructions the AECM kernel leans on (sm_100a): warp-instructions per cycle per SM
// sub-partition for pure streams and for 1:1 mixes, 8 independent chains per thread, 8 warps per sub-partition.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITERS 2048

#define OP_SHF(x, k) asm volatile("shr.s32 %0, %0, %1;" : "+r"(x) : "r"(k))
#define OP_IMAD(x, k) asm volatile("mad.lo.s32 %0, %0, %1, %1;" : "+r"(x) : "r"(k))
#define OP_IMADHI(x, k) asm volatile("mul.hi.s32 %0, %0, %1;" : "+r"(x) : "r"(k))
#define OP_IADD(x, k) asm volatile("add.s32 %0, %0, %1;" : "+r"(x) : "r"(k))
#define OP_LOP(x, k) asm volatile("xor.b32 %0, %0, %1;" : "+r"(x) : "r"(k))
#define OP_MNMX(x, k) asm volatile("max.s32 %0, %0, %1;" : "+r"(x) : "r"(k))
#define OP_PRMT(x, k) asm volatile("prmt.b32 %0, %0, %1, 0x7632;" : "+r"(x) : "r"(k))
#define OP_SHL(x, k) asm volatile("shl.b32 %0, %0, %1;" : "+r"(x) : "r"(k))
#define OP_SETP(x, k) asm volatile("{.reg .pred p; setp.gt.s32 p, %0, %1; selp.s32 %0, %0, %1, p;}" : "+r"(x) : "r"(k))

template <int MODE>
__global__ void probe(int *out, int k, long long *cycles) {
    int v[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; c++) v[c] = threadIdx.x * 7 + c + k;
    const long long t0 = clock64();
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int c = 0; c < CHAINS; c++) {
            const int y = v[(c + 1) % CHAINS];  // the neighbour chain: nothing for ptxas to fold
            if (MODE == 0) OP_SHF(v[c], y);
            if (MODE == 1) OP_IMAD(v[c], y);
            if (MODE == 2) OP_IMADHI(v[c], y);
            if (MODE == 3) OP_IADD(v[c], y);
            if (MODE == 4) OP_LOP(v[c], y);
            if (MODE == 5) OP_MNMX(v[c], y);
            if (MODE == 6) OP_PRMT(v[c], y);
            if (MODE == 7) { if (c & 1) OP_SHF(v[c], y); else OP_IMAD(v[c], y); }
            if (MODE == 8) { if (c & 1) OP_SHF(v[c], y); else OP_IMADHI(v[c], y); }
            if (MODE == 9) { if (c & 1) OP_IADD(v[c], y); else OP_IMAD(v[c], y); }
            if (MODE == 10) { if (c & 1) OP_SHF(v[c], y); else OP_IADD(v[c], y); }
            if (MODE == 11) OP_SHL(v[c], y);
            if (MODE == 12) OP_SETP(v[c], y);
            if (MODE == 13) { if (c & 1) OP_IMADHI(v[c], y); else OP_IMAD(v[c], y); }
        }
    }
    const long long t1 = clock64();
    int acc = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; c++) acc ^= v[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
    if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE>
void run(const char *name, int ops_per_slot) {
    const int blocks = 148, threads = 1024;  // 32 warps per SM = 8 per sub-partition
    int *out;
    long long *cyc, h[148];
    cudaMalloc(&out, blocks * threads * sizeof(int));
    cudaMalloc(&cyc, blocks * sizeof(long long));
    probe<MODE><<<blocks, threads>>>(out, 3, cyc);
    probe<MODE><<<blocks, threads>>>(out, 3, cyc);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    double mean = 0;
    for (int i = 0; i < blocks; i++) mean += (double)h[i] / blocks;
    const double warp_instr = (double)ITERS * CHAINS * ops_per_slot * (threads / 32) / 4.0;  // per sub-partition
    printf("%-28s %.3f warp-instr/cycle/sub-partition (%.0f cycles)\n", name, warp_instr / mean, mean);
    cudaFree(out);
    cudaFree(cyc);
}

int main() {
    run<0>("SHF (shr.s32)", 1);
    run<11>("SHL (shl.b32)", 1);
    run<1>("IMAD (mad.lo)", 1);
    run<2>("IMAD.HI (mul.hi)", 1);
    run<3>("IADD (add)", 1);
    run<4>("LOP3 (xor)", 1);
    run<5>("MNMX (max)", 1);
    run<6>("PRMT", 1);
    run<12>("ISETP+SEL", 2);
    run<7>("SHF : IMAD 1:1", 1);
    run<8>("SHF : IMAD.HI 1:1", 1);
    run<9>("IADD : IMAD 1:1", 1);
    run<10>("SHF : IADD 1:1", 1);
    run<13>("IMAD.HI : IMAD 1:1", 1);
    return 0;
}
