// Dependent-chain latencies on one warp (sm_100a): cycles per op, for the ops of the exact symbol step.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/mb/latency.cu -o /tmp/latency && /tmp/latency
#include <cstdio>
#include <cuda_runtime.h>
#define N 4096
#define CHAIN(name, init, body, fin) \
__global__ void k_##name(long long *out, float seed) { \
    init; \
    long long t0 = clock64(); \
    _Pragma("unroll 64") for (int i = 0; i < N; i++) { body; } \
    long long t1 = clock64(); \
    fin; \
    if (threadIdx.x == 0) out[0] = t1 - t0; \
}
CHAIN(fadd, float x = seed, x = __fadd_rn(x, 1.0001f), if (x == 12345.f) out[1] = 1)
CHAIN(fmul, float x = seed, x = __fmul_rn(x, 1.0000001f), if (x == 12345.f) out[1] = 1)
CHAIN(dadd, double x = seed, x = __dadd_rn(x, 1.0001), if (x == 12345.) out[1] = 1)
CHAIN(dmul, double x = seed, x = __dmul_rn(x, 1.0000001), if (x == 12345.) out[1] = 1)
CHAIN(dfma, double x = seed, x = __fma_rn(x, 1.0000001, 0.5), if (x == 12345.) out[1] = 1)
CHAIN(f2d2f, float x = seed, x = __double2float_rn(__dadd_rn((double)x, 1.0001)), if (x == 12345.f) out[1] = 1)
CHAIN(f2d_only, float x = seed, { double d = (double)x; x = __double2float_rn(d) + 1.0f; }, if (x == 12345.f) out[1] = 1)
CHAIN(d2i, double x = seed, { int v = __double2int_rz(x); x = (double)v + 0.5; }, if (x == 12345.) out[1] = 1)
CHAIN(f2i_i2f, float x = seed, { int v = __float2int_rz(x); x = (float)v + 0.5f; }, if (x == 12345.f) out[1] = 1)
CHAIN(rsqrt, float x = seed + 2.f, { float y; asm volatile("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); x = y + 1.5f; }, if (x == 12345.f) out[1] = 1)
CHAIN(imad, int x = (int)seed, x = x * 3 + 7, if (x == 12345) out[1] = 1)
CHAIN(prmt, unsigned x = (unsigned)seed, x = __byte_perm(x, 0x4B000000u, 0x7610) + 1u, if (x == 12345u) out[1] = 1)
CHAIN(shfl, float x = seed, x = __shfl_sync(0xffffffffu, x, (threadIdx.x + 1) & 31) + 1.0f, if (x == 12345.f) out[1] = 1)
CHAIN(vote, unsigned x = (unsigned)seed, x = __ballot_sync(0xffffffffu, x & 1) + threadIdx.x, if (x == 12345u) out[1] = 1)
__global__ void k_lds(long long *out, float seed) {
    __shared__ int next[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) next[i] = (i * 37 + 11) & 1023;
    __syncwarp();
    int p = (int)seed & 1023;
    long long t0 = clock64();
#pragma unroll 64
    for (int i = 0; i < N; i++) p = next[p];
    long long t1 = clock64();
    if (p == 12345) out[1] = 1;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
// mailbox ping-pong between two warps on different SM sub-partitions: cycles per round trip
__global__ void k_pingpong(long long *out, float seed) {
    __shared__ volatile int a[32], b[32];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x < 32) { a[lane] = 0; b[lane] = 0; }
    __syncthreads();
    long long t0 = clock64();
    if (warp == 0) {
        for (int i = 1; i <= 2048; i++) { a[lane] = i; while (__any_sync(0xffffffffu, b[lane] != i)) { } }
    } else if (warp == 1) {
        for (int i = 1; i <= 2048; i++) { while (__any_sync(0xffffffffu, a[lane] != i)) { } b[lane] = i; }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out[0] = (t1 - t0) * 2;   // scaled so that /N(4096) prints cycles per ROUND TRIP
}
// taken-branch cost for a lone warp: a chain of data-dependent taken branches
__global__ void k_branch(long long *out, float seed) {
    int x = (int)seed;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < N; i++) { x = x * 3 + 1; }
    long long t1 = clock64();
    if (x == 12345) out[1] = 1;
    if (threadIdx.x == 0) out[0] = t1 - t0;
}
int main() {
    long long *d, h[2];
    cudaMalloc(&d, 16);
#define RUN(name, threads) do { k_##name<<<1, threads>>>(d, 1.0f); cudaDeviceSynchronize(); k_##name<<<1, threads>>>(d, 1.0f); cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost); \
    printf("%-10s %7.2f cycles/op%s\n", #name, (double)h[0] / N, cudaGetLastError() ? " (error)" : ""); } while (0)
    RUN(fadd, 32); RUN(fmul, 32); RUN(dadd, 32); RUN(dmul, 32); RUN(dfma, 32); RUN(f2d2f, 32); RUN(f2d_only, 32); RUN(d2i, 32);
    RUN(f2i_i2f, 32); RUN(rsqrt, 32); RUN(imad, 32); RUN(prmt, 32); RUN(shfl, 32); RUN(vote, 32); RUN(lds, 32); RUN(branch, 32);
    RUN(pingpong, 64);
    printf("notes: f2d2f = F2F.F64.F32 + DADD + F2F.F32.F64; f2d_only = F2F.F64.F32 + F2F.F32.F64 + FADD; d2i = F2I.F64 + I2F.F64 + DADD;\n"
           "f2i_i2f = F2I + I2F + FADD; rsqrt = MUFU.RSQ + FADD; prmt = PRMT + IADD; shfl = SHFL + FADD; vote = VOTE + IADD; branch = loop of IMAD + taken BRA;\n"
           "pingpong = one mailbox round trip between two warps (volatile shared store -> polled load, both ways)\n");
    return 0;
}
