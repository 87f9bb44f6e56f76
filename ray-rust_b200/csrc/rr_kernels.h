// rr_kernels.h — host-callable launchers of the sm_100a kernels (implemented in rr_trace.cu,
// rr_march.cu, rr_util.cu). Internal to libray_rust_b200.so; the public surface is include/rr_ffi.h.
#pragma once
#include <cuda_runtime.h>

#include "rr_device.cuh"

namespace rr {

constexpr int RR_MAX_STACK_HOST = 32;  // == RR_MAX_STACK / RR_MARCH_MAX_STACK in the kernels

struct LaunchInfo {
    int sm_count;
    size_t smem_optin;  // max opt-in dynamic shared memory per block
};

// Tile-row order of a launch (a permutation of its (local_rows + 3) / 4 tile rows of 8x4 warp tiles, or nullptr) and where
// the kernel records the longest tile of every tile row (zeroed by the caller, or nullptr): see MarchProfile in rr_ffi.cu.
// Ray-march kernel only: the same order on the BVH instances of the ray-trace kernel measured neutral on one GPU (1.790 vs
// 1.786 ms) and 14 % SLOWER on band-sharded frames (4 GPUs: 0.597 -> 0.679 ms, profiles/r3l_bench_n4_rowprofile_bvh.json).
struct RowProfile {
    const int *order;
    unsigned *cost;
};
// Ray-trace mode. d_out: RGB8 (row_stride bytes per row) or, when f32_out, packed float rgb.
// d_cnt != nullptr selects the instrumented instantiation.
cudaError_t launch_trace(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride, bool f32_out,
                         Counters *d_cnt, cudaStream_t stream, const LaunchInfo &li, bool allow_bvh, const Signal &sig = Signal{nullptr, nullptr, nullptr, 0u});
// Ray-march mode (same contract). sig.work / sig.done: this launch's tile-queue word and block counter (both zero at
// launch; the last block resets them), sig.flag: optional completion word, as for the trace kernel.
cudaError_t launch_march(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride, bool f32_out,
                         Counters *d_cnt, const Signal &sig, cudaStream_t stream, const LaunchInfo &li, bool allow_bvh = true,
                         const RowProfile &prof = RowProfile{nullptr, nullptr});
// Fills the primary-ray tables of a frame (FrameParams::ptab layout: xres column entries, then yres row entries).
cudaError_t launch_prim_table(const FrameParams &P, float4 *d_tab, cudaStream_t stream);
// Row-band un-interleave (multi-GPU gather epilogue).
cudaError_t launch_bands_unpack(const FrameParams &P, const void *d_packed, size_t shard_stride, void *d_frame,
                                cudaStream_t stream);
// Completion signalling between GPUs without a collective (see Signal in rr_device.cuh):
//   launch_signal: stand-alone publisher for kernels that do not carry the signal themselves (march mode);
//   launch_fence_wait: the frame owner's stream waits until `count` flag words have reached `epoch`
//   (bounded spin; *d_status = 1 on timeout when d_status is given).
cudaError_t launch_signal(const Signal &sig, cudaStream_t stream);
cudaError_t preload_signal_kernels();
cudaError_t launch_fence_wait(const unsigned *d_flags, int count, unsigned epoch, unsigned timeout_ms, unsigned *d_status,
                              cudaStream_t stream);
// normalized() self-test: number of vectors (of n hashed ones) on which the shared-reciprocal path and three plain IEEE
// divisions differ in any bit.
cudaError_t normalize_selftest(int device, unsigned long long n, unsigned long long seed, unsigned long long *mismatches);
// FP32 pipe calibration (roofline denominator): achieved TFLOP/s of unfused FMUL+FADD and of FFMA.
cudaError_t fp32_peak(int device, float *unfused_tflops, float *ffma_tflops);

}  // namespace rr
