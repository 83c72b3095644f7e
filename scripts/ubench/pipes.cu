// Microbenchmark: latency and issue rate of the integer instructions the BoxBlur segment kernels are built from, on one SM.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench_pipes scripts/ubench/pipes.cu ; run on the GPU box.
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

enum Op { IDP_LO, IDP_HI, IMAD, IADD3, IADD2, LEAHI, SHF, PRMT, LOP, MIX_IDP_IMAD, MIX_IDP_IADD3, MIX_IDP_SHF, MIX_IMAD_IADD3, MIX_IDP_PRMT, NOPS };
const char* names[] = {"IDP.2A lo", "IDP.2A hi", "IMAD", "IADD3(3-in)", "IADD(2-in)", "LEA.HI(add>>16)", "SHF", "PRMT", "LOP3",
                       "IDP+IMAD alt", "IDP+IADD3 alt", "IDP+SHF alt", "IMAD+IADD3 alt", "IDP+PRMT alt"};

template <int OP>
__device__ __forceinline__ uint32_t op(uint32_t x, uint32_t a, uint32_t b, int i) {
    uint32_t d;
    switch (OP) {
        case IDP_LO: asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0x0001u), "r"(x)); return d;
        case IDP_HI: asm volatile("dp2a.lo.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0xff00u), "r"(x)); return d;
        case IMAD: asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(a), "r"(b)); return d;
        case IADD3: asm volatile("{ .reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3; }" : "=r"(d) : "r"(x), "r"(a), "r"(b)); return d;
        case IADD2: asm volatile("add.u32 %0, %1, %2;" : "=r"(d) : "r"(x), "r"(a)); return d;
        case LEAHI: asm volatile("{ .reg .u32 t; shr.u32 t, %2, 16; add.u32 %0, %1, t; }" : "=r"(d) : "r"(x), "r"(a)); return d;
        case SHF: asm volatile("shf.r.wrap.b32 %0, %1, %2, 16;" : "=r"(d) : "r"(x), "r"(a)); return d;
        case PRMT: asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(d) : "r"(x), "r"(a)); return d;
        case LOP: asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(x), "r"(a), "r"(b)); return d;
        case MIX_IDP_IMAD: return (i & 1) ? op<IMAD>(x, a, b, i) : op<IDP_LO>(x, a, b, i);
        case MIX_IDP_IADD3: return (i & 1) ? op<IADD3>(x, a, b, i) : op<IDP_LO>(x, a, b, i);
        case MIX_IDP_SHF: return (i & 1) ? op<SHF>(x, a, b, i) : op<IDP_LO>(x, a, b, i);
        case MIX_IMAD_IADD3: return (i & 1) ? op<IADD3>(x, a, b, i) : op<IMAD>(x, a, b, i);
        case MIX_IDP_PRMT: return (i & 1) ? op<PRMT>(x, a, b, i) : op<IDP_LO>(x, a, b, i);
    }
    return x;
}

// CHAINS independent dependency chains per thread, ITER*UNR ops per chain
template <int OP, int CHAINS>
__global__ void bench(uint32_t* out, long long* cyc, uint32_t a, uint32_t b, int iters) {
    uint32_t x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x + c;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 16; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = op<OP>(x[c], a + c, b, u);
    }
    const long long t1 = clock64();
    uint32_t s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int OP, int CHAINS>
double run(int threads) {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 4096 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    bench<OP, CHAINS><<<1, threads>>>(out, cyc, 0x12345678u, 3u, iters);
    bench<OP, CHAINS><<<1, threads>>>(out, cyc, 0x12345678u, 3u, iters);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    cudaFree(out); cudaFree(cyc);
    return (double)c / ((double)iters * 16 * CHAINS);  // cycles per warp-instruction of one warp
}

template <int OP>
void report() {
    // latency: 1 warp, 1 chain.  issue rate per SMSP: 4 warps/SMSP (512 threads), 8 chains -> cycles per instruction per SMSP
    const double lat = run<OP, 1>(32);
    const double t1 = run<OP, 8>(128);    // 1 warp per SMSP
    const double t4 = run<OP, 8>(512);    // 4 warps per SMSP
    const double t8 = run<OP, 4>(1024);   // 8 warps per SMSP
    printf("%-18s latency %5.2f cyc | cycles per warp-instr per SMSP: 1 warp x8 chains %5.2f, 4 warps %5.2f, 8 warps(x4) %5.2f\n", names[OP], lat, t1,
           t4 / 4, t8 / 8);
}

int main() {
    report<IDP_LO>(); report<IDP_HI>(); report<IMAD>(); report<IADD3>(); report<IADD2>(); report<LEAHI>(); report<SHF>(); report<PRMT>(); report<LOP>();
    report<MIX_IDP_IMAD>(); report<MIX_IDP_IADD3>(); report<MIX_IDP_SHF>(); report<MIX_IMAD_IADD3>(); report<MIX_IDP_PRMT>();
    return 0;
}
