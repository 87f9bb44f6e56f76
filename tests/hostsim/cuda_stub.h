// Minimal host stand-ins for the CUDA built-ins used by ray-rust_b200/csrc/rr_device.cuh, rr_trace.cuh and
// rr_march.cuh, so the per-pixel device logic can be compiled with g++ and run on the CPU (tests only).
// Scalar f32 with -ffp-contract=off gives the same +,-,*,/,sqrt bits as the GPU build (-fmad=false).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <algorithm>

#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__
#define __restrict__
#define __grid_constant__
#define RR_HOSTSIM 1

struct float2 { float x, y; };
struct float4 { float x, y, z, w; };
struct int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
static inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
static inline int4 make_int4(int x, int y, int z, int w) { return int4{x, y, z, w}; }
static inline float2 make_float2(float x, float y) { return float2{x, y}; }
struct dim3s { unsigned x = 0, y = 0, z = 0; };
static dim3s threadIdx, blockIdx, blockDim, gridDim;

template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
static inline int __float_as_int(float f) { int i; std::memcpy(&i, &f, 4); return i; }
static inline int __float2int_rz(float x) {
    if (x != x) return 0;
    if (x >= 2147483648.0f) return INT32_MAX;
    if (x <= -2147483648.0f) return INT32_MIN;
    return (int)x;
}
static inline unsigned __float2uint_rz(float x) {
    if (x != x) return 0;
    if (x >= 4294967296.0f) return UINT32_MAX;
    if (x <= 0.0f) return 0;
    return (unsigned)x;
}
static inline float __fmaf_rn(float a, float b, float c) { return std::fmaf(a, b, c); }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline unsigned __shfl_sync(unsigned, unsigned v, int) { return v; }
static inline unsigned __activemask() { return 1u; }
static inline bool __all_sync(unsigned, bool p) { return p; }  // a one-lane warp
static inline unsigned long long __shfl_down_sync(unsigned, unsigned long long v, int) { return v; }
static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { unsigned long long o = *p; *p += v; return o; }
using std::min;
using std::max;
typedef int cudaError_t;
typedef void *cudaStream_t;
