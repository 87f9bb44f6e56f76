// Micro-benchmark: issue cost of Blackwell's packed FFMA2 next to scalar FFMA / FMUL+FADD, alone and mixed with ALU work.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false tools/ffma2_bench.cu -o ab/ffma2_bench
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
struct K { float2 one, nz, a, b; };
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c) { u64 r; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r; }
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }

// MODE 0: 8 chains of scalar FFMA        (1 flop-pair per lane per instruction)
// MODE 1: 8 chains of scalar FMUL + FADD (the render kernels' unfused mix)
// MODE 2: 8 chains of packed FFMA2       (two FMAs per lane per instruction)
// MODE 3: 8 chains of FFMA2 + 8 integer LOP3/IADD ops per step (does the ALU pipe issue beside FFMA2?)
// MODE 4: 8 chains of scalar FFMA + the same 8 integer ops
// MODE 5: 8 chains of FFMA2 emulating unfused mul then add (2 FFMA2 per mul-add pair) = the packed form of MODE 1
template <int MODE>
__global__ void __launch_bounds__(256) chain(const __grid_constant__ K k, float *out, int iters) {
    float x[8]; u64 X[8]; unsigned n[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { x[i] = (threadIdx.x + i) * 1e-3f; X[i] = pk(x[i], x[i] + 1.0f); n[i] = threadIdx.x * 7 + i; }
    const u64 A = pk(k.a.x, k.a.y), B = pk(k.b.x, k.b.y), ONE = pk(k.one.x, k.one.y), NZ = pk(k.nz.x, k.nz.y);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (MODE == 0 || MODE == 4) x[i] = __fmaf_rn(x[i], k.a.x, k.b.x);
            if (MODE == 1) x[i] = x[i] * k.a.x + k.b.x;
            if (MODE == 2 || MODE == 3) X[i] = fma2(X[i], A, B);
            if (MODE == 5) X[i] = fma2(fma2(X[i], A, NZ), ONE, B);
            if (MODE == 3 || MODE == 4) n[i] = (n[i] ^ (n[i] >> 3)) + 0x9e3779b9u;
        }
    }
    float s = 0.0f; unsigned m = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(X[i])); s += x[i] + lo + hi; m ^= n[i]; }
    if (s == 123.456f || m == 0x12345u) out[0] = s;
}
template <int MODE>
static void run(const char *name, double flops_per_step, double inst_per_step) {
    int dev = 0, sm = 0, khz = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    float *d; cudaMalloc(&d, 4);
    K k{{1.0f, 1.0f}, {-0.0f, -0.0f}, {0.999f, 0.998f}, {1e-3f, 2e-3f}};
    const int iters = 1 << 14, blocks = sm * 8, threads = 256;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0); chain<MODE><<<blocks, threads>>>(k, d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1); if (r && ms < best) best = ms;
    }
    const double steps = 8.0 * iters * (double)blocks * threads;  // chain steps (thread level)
    const double warp_inst = steps * inst_per_step / 32.0;
    const double cyc = best * 1e-3 * khz * 1e3;                   // at the nominal max clock
    printf("%-46s %8.3f ms  %7.2f TFLOP/s  %6.3f warp-inst/clk/SMSP (at %d MHz nominal)\n", name, best,
           steps * flops_per_step / (best * 1e-3) / 1e12, warp_inst / cyc / (sm * 4.0), khz / 1000);
    cudaFree(d);
}
int main() {
    run<0>("scalar FFMA", 2, 1);
    run<1>("scalar FMUL+FADD (unfused)", 2, 2);
    run<2>("packed FFMA2", 4, 1);
    run<5>("packed FFMA2 x2 = unfused mul,add on 2 floats", 4, 2);
    run<3>("packed FFMA2 + 3 ALU ops (LOP3/SHF/IADD)", 4, 4);
    run<4>("scalar FFMA + 3 ALU ops", 2, 4);
    return cudaDeviceSynchronize() != cudaSuccess;
}
