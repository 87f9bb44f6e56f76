// rr_util.cu — row-band un-interleave kernel and the FP32 pipe calibration micro-benchmark.
#include "rr_kernels.h"

namespace rr {

// One block per image row: copy the row from the shard that rendered it (packed band order) to its
// place in the row-major frame. 16-byte vectors when everything is aligned, bytes otherwise.
__global__ void bands_unpack_kernel(const uint8_t *__restrict__ packed, size_t shard_stride, uint8_t *__restrict__ frame,
                                    int W, int H, int band_rows, int band_count) {
    const int iy = blockIdx.x;
    if (iy >= H) return;
    const int b = iy / band_rows;
    const int shard = b % band_count;
    const int local_row = (b / band_count) * band_rows + (iy - b * band_rows);
    const size_t row_bytes = (size_t)W * 3;
    const uint8_t *src = packed + (size_t)shard * shard_stride + (size_t)local_row * row_bytes;
    uint8_t *dst = frame + (size_t)iy * row_bytes;
    if (((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst) | row_bytes) & 15) == 0) {
        const uint4 *s4 = reinterpret_cast<const uint4 *>(src);
        uint4 *d4 = reinterpret_cast<uint4 *>(dst);
        for (size_t i = threadIdx.x; i < row_bytes / 16; i += blockDim.x) d4[i] = s4[i];
    } else {
        for (size_t i = threadIdx.x; i < row_bytes; i += blockDim.x) dst[i] = src[i];
    }
}

cudaError_t launch_bands_unpack(const FrameParams &P, const void *d_packed, size_t shard_stride, void *d_frame,
                                cudaStream_t stream) {
    if (P.yres <= 0 || P.xres <= 0) return cudaSuccess;
    const int br = P.band_rows <= 0 ? 1 : P.band_rows;
    const int bc = P.band_count <= 1 ? 1 : P.band_count;
    bands_unpack_kernel<<<P.yres, 256, 0, stream>>>(reinterpret_cast<const uint8_t *>(d_packed), shard_stride,
                                                   reinterpret_cast<uint8_t *>(d_frame), P.xres, P.yres, br, bc);
    return cudaGetLastError();
}

// ---- primary-ray tables of the trace kernel (primary_dir_tab, rr_device.cuh) ---------------------
// xres column entries then yres row entries; the same f32 products the per-pixel code would form (this file is compiled
// with -fmad=false like the render kernels).
__global__ void prim_table_kernel(const __grid_constant__ FrameParams P, float4 *__restrict__ tab) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < P.xres) tab[i] = prim_col_entry(P, i);
    else if (i < P.xres + P.yres) tab[i] = prim_row_entry(P, i - P.xres);
}

cudaError_t launch_prim_table(const FrameParams &P, float4 *d_tab, cudaStream_t stream) {
    const int n = P.xres + P.yres;
    if (n <= 0) return cudaSuccess;
    prim_table_kernel<<<(n + 255) / 256, 256, 0, stream>>>(P, d_tab);
    return cudaGetLastError();
}

// ---- completion signal / wait (multi-GPU placed frames) ----------------------------------------
__global__ void signal_kernel(const Signal sig) {
    // everything earlier on this stream has completed and is visible; publish with system scope
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(sig.flag), "r"(sig.epoch) : "memory");
}

cudaError_t launch_signal(const Signal &sig, cudaStream_t stream) {
    if (!sig.flag) return cudaSuccess;
    signal_kernel<<<1, 1, 0, stream>>>(sig);
    return cudaGetLastError();
}

// One thread per flag word. Epochs only grow, so "reached" is a signed distance test (wrap-safe).
__global__ void fence_wait_kernel(const unsigned *flags, int count, unsigned epoch, unsigned long long timeout_ns, unsigned *status) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    unsigned long long t0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
    for (;;) {
        unsigned v;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + i) : "memory");
        if ((int)(v - epoch) >= 0) break;
        unsigned long long t1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
        if (t1 - t0 > timeout_ns) {
            if (status) atomicExch(status, 1u);
            break;
        }
        __nanosleep(100);
    }
    __threadfence_system();
}

cudaError_t launch_fence_wait(const unsigned *d_flags, int count, unsigned epoch, unsigned timeout_ms, unsigned *d_status,
                              cudaStream_t stream) {
    if (count <= 0) return cudaSuccess;
    const int threads = count < 128 ? count : 128;
    fence_wait_kernel<<<(count + threads - 1) / threads, threads, 0, stream>>>(d_flags, count, epoch,
                                                                             (unsigned long long)timeout_ms * 1000000ull, d_status);
    return cudaGetLastError();
}

// With CUDA's lazy module loading a kernel is loaded at its first launch, which cannot complete while another kernel
// (a spinning fence_wait_kernel) occupies the device. Load the two signalling kernels up front.
cudaError_t preload_signal_kernels() {
    cudaFuncAttributes a;
    cudaError_t e = cudaFuncGetAttributes(&a, signal_kernel);
    if (e != cudaSuccess) return e;
    return cudaFuncGetAttributes(&a, fence_wait_kernel);
}

// ---- self-test of normalized() (rr_device.cuh): shared-reciprocal path against three plain IEEE divisions -----------
// Operands: a counter-based hash gives three floats with independent signs, mantissas and exponents; every 8th vector gets
// exponents from the whole f32 range (zero, subnormal, huge, infinite and NaN components included) so that the guard and the
// plain-division branch are exercised too, the others stay within 2^-44 .. 2^44 around the guard's limits.
__device__ __forceinline__ unsigned st_hash(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
    return (unsigned)x;
}
__global__ void normalize_selftest_kernel(unsigned long long n, unsigned long long seed, unsigned long long *mismatches) {
    unsigned long long bad = 0;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) {
        float c[3];
        const bool wild = (i & 7ull) == 7ull;
        for (int k = 0; k < 3; ++k) {
            const unsigned h = st_hash(seed + 3ull * i + k), h2 = st_hash(~seed + 5ull * i + k);
            const unsigned ex = wild ? (h2 % 256u) : (127u - 44u + h2 % 89u);
            c[k] = __uint_as_float((h & 0x807fffffu) | (ex << 23));
            if (!wild && (h2 >> 28) == 0u) c[k] = c[(k + 1) % 3] * 0.0f + c[k] * 1e-3f;  // some strongly unequal magnitudes
        }
        const V3 v = mk(c[0], c[1], c[2]);
        const V3 a = normalized(v), b = normalized_plain(v);
        if (__float_as_uint(a.x) != __float_as_uint(b.x) || __float_as_uint(a.y) != __float_as_uint(b.y) || __float_as_uint(a.z) != __float_as_uint(b.z)) ++bad;
        // the same three numbers as (numerator, numerator, divisor) of div_pair(): quotients far from 1
        float u, w;
        div_pair(c[0], c[1], c[2], u, w);
        if (__float_as_uint(u) != __float_as_uint(c[0] / c[2]) || __float_as_uint(w) != __float_as_uint(c[1] / c[2])) ++bad;
    }
    if (bad) atomicAdd(mismatches, bad);
}

cudaError_t normalize_selftest(int device, unsigned long long n, unsigned long long seed, unsigned long long *mismatches) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    unsigned long long *d = nullptr;
    e = cudaMalloc(&d, sizeof *d);
    if (e != cudaSuccess) return e;
    cudaMemset(d, 0, sizeof *d);
    normalize_selftest_kernel<<<1184, 256>>>(n, seed, d);
    e = cudaMemcpy(mismatches, d, sizeof *d, cudaMemcpyDeviceToHost);
    cudaFree(d);
    return e != cudaSuccess ? e : cudaGetLastError();
}

// ---- FP32 pipe calibration -------------------------------------------------------------------
// 8 independent dependency chains per thread so the 4-cycle FMA-pipe latency is covered at
// 16 warps/SMSP. FUSED=false compiles (under -fmad=false) to FMUL+FADD pairs: the instruction mix
// bit-exact parity forces on the render kernels; FUSED=true uses explicit fmaf -> FFMA.
template <bool FUSED>
__global__ void __launch_bounds__(256) fp32_chain_kernel(float *out, int iters, float a, float b) {
    float x[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) x[k] = (float)(threadIdx.x + k) * 1e-3f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 8; ++k) x[k] = FUSED ? __fmaf_rn(x[k], a, b) : (x[k] * a + b);
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) s += x[k];
    if (s == 123.456f) out[0] = s;  // keep the chains alive
}

template <bool FUSED>
static cudaError_t time_chain(int sm_count, float *tflops) {
    float *d = nullptr;
    cudaError_t e = cudaMalloc(&d, 4);
    if (e != cudaSuccess) return e;
    const int iters = 1 << 15, blocks = sm_count * 8, threads = 256;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        fp32_chain_kernel<FUSED><<<blocks, threads>>>(d, iters, 0.999f, 1e-3f);
        cudaEventRecord(e1);
        e = cudaEventSynchronize(e1);
        if (e != cudaSuccess) break;
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d);
    if (e != cudaSuccess) return e;
    const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * (double)threads;
    *tflops = (float)(flops / (best * 1e-3) / 1e12);
    return cudaGetLastError();
}

cudaError_t fp32_peak(int device, float *unfused_tflops, float *ffma_tflops) {
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return e;
    int sm = 0;
    e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device);
    if (e != cudaSuccess) return e;
    e = time_chain<false>(sm, unfused_tflops);
    if (e != cudaSuccess) return e;
    return time_chain<true>(sm, ffma_tflops);
}

}  // namespace rr
