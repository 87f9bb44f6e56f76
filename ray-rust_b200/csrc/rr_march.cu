// rr_march.cu — ray-march kernel for sm_100a (render.rs:806-827 + :1226-1411 + quantiser).
//
// Work distribution: the iteration count per marched ray is heavy-tailed (1.3 % of
// rays run the 10 001-iteration cap and hold 74 % of all iterations, concentrated in the horizon
// rows), so tiles are handed out dynamically: each warp of a persistent grid pulls the next 8x4
// pixel tile from a global atomic counter when it finishes one.
// The ray-march kernel keeps plain IEEE divisions (SharedRcp, rr_device.cuh): its frame time is bound by the longest
// dependent march chains, not by instruction issue, and the shared-reciprocal build measured 1 % slower here
// (6.21 -> 6.29 ms at 4K, profiles/r2y_ab.txt) where the trace kernels gain 3-5 %.
#define RR_SHARED_RCP 0
#include "rr_kernels.h"
#include "rr_march.cuh"

namespace rr {

constexpr int MARCH_THREADS = 128;

template <typename T>
__device__ __forceinline__ void copy_tail(T *dst, const T *src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// Only the list tails (objects beyond the SceneHead) are staged; small scenes stage nothing.
__device__ __forceinline__ MarchView stage_march(const DevScene &G, float4 *smem, bool stage) {
    MarchView S;
    S.n_spheres = G.n_spheres;
    S.n_floors = G.n_floors;
    S.sph = G.sph_m; S.sph_glow = G.sph_glow; S.sph_oi = G.sph_oi; S.flo_o = G.flo_o; S.flo_n = G.flo_n; S.flo_oi = G.flo_oi;
    // MBVH instances read the BVH and the leaf-ordered spheres from global memory through L1 (40 KB for 1 024 spheres)
    S.bvh_a = G.bvh_a; S.bvh_b = G.bvh_b; S.bsph = G.bsph_m; S.bsph_oi = G.bsph_oi; S.n_bvh_nodes = G.n_bvh_nodes;
    S.scene_abs = 0.0f;
    for (int k = 0; k < 3; ++k) S.scene_abs = fmaxf(S.scene_abs, fmaxf(fabsf(G.scene_lo[k]), fabsf(G.scene_hi[k])));
    if (!stage) return S;
    const int ts = max(G.n_spheres - RR_HEAD_SPHERES, 0), tf = max(G.n_floors - RR_HEAD_FLOORS, 0);
    if (ts + tf == 0) return S;
    float4 *p = smem;
    float4 *sph = p; p += ts;
    float4 *flo_o = p; p += tf;
    float4 *flo_n = p; p += tf;
    float *sph_glow = reinterpret_cast<float *>(p);
    int *sph_oi = reinterpret_cast<int *>(sph_glow + ts);
    int *flo_oi = sph_oi + ts;
    copy_tail(sph, G.sph_m + RR_HEAD_SPHERES, ts);
    copy_tail(sph_glow, G.sph_glow + RR_HEAD_SPHERES, ts);
    copy_tail(sph_oi, G.sph_oi + RR_HEAD_SPHERES, ts);
    copy_tail(flo_o, G.flo_o + RR_HEAD_FLOORS, tf);
    copy_tail(flo_n, G.flo_n + RR_HEAD_FLOORS, tf);
    copy_tail(flo_oi, G.flo_oi + RR_HEAD_FLOORS, tf);
    __syncthreads();
    S.sph = sph - RR_HEAD_SPHERES; S.sph_glow = sph_glow - RR_HEAD_SPHERES; S.sph_oi = sph_oi - RR_HEAD_SPHERES;
    S.flo_o = flo_o - RR_HEAD_FLOORS; S.flo_n = flo_n - RR_HEAD_FLOORS; S.flo_oi = flo_oi - RR_HEAD_FLOORS;
    return S;
}

// (no minimum block count on purpose: __launch_bounds__(128, 7) yields the same 72 registers, scheduled differently, and a
// 7-10 % slower kernel; (128, 8) = 64 registers is 6 % slower too: profiles/r2k_ab_march.txt)
template <bool COUNT, bool F32OUT, bool STAGE, int GLOW, bool MBVH>
__global__ void __launch_bounds__(MARCH_THREADS)
march_kernel(const __grid_constant__ DevScene G, const __grid_constant__ SceneHead H, const __grid_constant__ FrameParams P,
             void *__restrict__ out, size_t row_stride, Counters *gcnt, int fast_store, const Signal sig, const RowProfile prof) {
    extern __shared__ float4 rr_smem[];
    const MarchView S = stage_march(G, rr_smem, STAGE);

    const int W = P.xres, rows = P.local_rows;
    const int tiles_x = (W + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const int ntiles = tiles_x * tiles_y;
    const int lane = threadIdx.x & 31;
    const int col = lane & 7, row = lane >> 3;
    Counters cnt = {};
    const int rot = P.march_tile_rot < tiles_y ? P.march_tile_rot : 0;

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(sig.work, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int tq = tile / tiles_x, tx = tile - tq * tiles_x;
        // tile ROW order: the handle's profile of an earlier launch of this frame geometry (longest rows first), else from
        // the horizon rows on (FrameParams::march_tile_rot)
        const int ty = prof.order ? __ldg(&prof.order[tq]) : (tq + rot < tiles_y ? tq + rot : tq + rot - tiles_y);
        const long long t_begin = prof.cost ? clock64() : 0ll;
        const int x0 = tx << 3, ly0 = ty << 2;
        const int ix = x0 + col, ly = ly0 + row;
        const bool valid = ix < W && ly < rows;
        V3 c = mk(0.0f, 0.0f, 0.0f);
        if (valid) c = march_pixel<COUNT, GLOW, MBVH>(G, H, S, P, ix, local_to_image_row(P, ly), cnt);
        if (F32OUT) {
            if (valid) {
                float *o = reinterpret_cast<float *>(out) + ((size_t)(P.placed ? local_to_image_row(P, ly) : ly) * W + ix) * 3;
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            }
        } else {
            const unsigned rgb = quantize(c.x) | (quantize(c.y) << 8) | (quantize(c.z) << 16);
            store_tile_rgb8(reinterpret_cast<uint8_t *>(out), row_stride, x0, ly0, W, rows, rgb, fast_store != 0,
                            P.placed ? local_to_image_row(P, ly) : ly);
        }
        if (prof.cost && lane == 0) {  // the row's longest tile, in units of 64 clocks
            const long long d = (clock64() - t_begin) >> 6;
            atomicMax(&prof.cost[ty], (unsigned)(d < 0 ? 0 : (d > 0xffffffffll ? 0xffffffffll : d)));
        }
    }
    if (COUNT) flush_counters(cnt, gcnt);
    // Like the trace kernel: the block that finishes last resets the launch's queue word and block counter (so the slot
    // is clean for its next use without a memset on the stream) and, for placed multi-GPU frames, publishes the
    // completion word in the frame owner's memory.
    finish_launch(sig);
}

static size_t march_smem_bytes(const DevScene &G) {
    const size_t ts = G.n_spheres > RR_HEAD_SPHERES ? G.n_spheres - RR_HEAD_SPHERES : 0;
    const size_t tf = G.n_floors > RR_HEAD_FLOORS ? G.n_floors - RR_HEAD_FLOORS : 0;
    return ts * (sizeof(float4) + sizeof(float) + sizeof(int)) + tf * (2 * sizeof(float4) + sizeof(int)) + 16;
}

template <bool COUNT, bool F32OUT, bool STAGE, int GLOW, bool MBVH = false>
static cudaError_t launch_one(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                              Counters *d_cnt, const Signal &sig, cudaStream_t stream, const LaunchInfo &li, size_t smem, const RowProfile &prof) {
    auto kern = march_kernel<COUNT, F32OUT, STAGE, GLOW, MBVH>;
    cudaError_t e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MARCH_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const long long tiles = (long long)((P.xres + 7) / 8) * ((P.local_rows + 3) / 4);
    const long long need = (tiles + (MARCH_THREADS / 32) - 1) / (MARCH_THREADS / 32);
    long long grid = (long long)li.sm_count * per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    const int fast = (!F32OUT && (P.xres % 8 == 0) && (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3) == 0)) ? 1 : 0;
    kern<<<(unsigned)grid, MARCH_THREADS, smem, stream>>>(G, H, P, d_out, row_stride, d_cnt, fast, sig, prof);
    return cudaGetLastError();
}

template <bool COUNT, bool F32OUT>
static cudaError_t launch_two(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                              Counters *d_cnt, const Signal &sig, cudaStream_t stream, const LaunchInfo &li, bool allow_bvh, const RowProfile &prof) {
    size_t smem = march_smem_bytes(G);
    const bool stage = smem <= li.smem_optin / 2;
    if (!stage) smem = 0;
    // glow tracking: 0 = nothing reads it, 1 = separate pass over the <= RR_HEAD_GLOW glowing objects,
    // 2 = inline in the scan (many glowing objects)
    const int glow = !(P.glow_enabled && G.n_glow > 0) ? 0 : (H.n_glow_head >= 0 ? 1 : 2);
    // Large scenes: the sphere scan goes through the BVH (rr_march.cuh, MBVH); nothing is staged (floor tails and BVH are
    // read through L1). Inline glow keeps the linear scan.
    if (allow_bvh && G.n_bvh_nodes > 0 && glow != 2) {
        if (glow == 0) return launch_one<COUNT, F32OUT, false, 0, true>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, 0, prof);
        return launch_one<COUNT, F32OUT, false, 1, true>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, 0, prof);
    }
    if (stage) {
        if (glow == 0) return launch_one<COUNT, F32OUT, true, 0>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, smem, prof);
        if (glow == 1) return launch_one<COUNT, F32OUT, true, 1>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, smem, prof);
        return launch_one<COUNT, F32OUT, true, 2>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, smem, prof);
    }
    if (glow == 0) return launch_one<COUNT, F32OUT, false, 0>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, smem, prof);
    if (glow == 1) return launch_one<COUNT, F32OUT, false, 1>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, smem, prof);
    return launch_one<COUNT, F32OUT, false, 2>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, smem, prof);
}

cudaError_t launch_march(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                         bool f32_out, Counters *d_cnt, const Signal &sig, cudaStream_t stream, const LaunchInfo &li, bool allow_bvh, const RowProfile &prof) {
    if (P.xres <= 0 || P.local_rows <= 0) return cudaSuccess;
    if (d_cnt) return f32_out ? launch_two<true, true>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, allow_bvh, prof)
                              : launch_two<true, false>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, allow_bvh, prof);
    return f32_out ? launch_two<false, true>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, allow_bvh, prof)
                   : launch_two<false, false>(G, H, P, d_out, row_stride, d_cnt, sig, stream, li, allow_bvh, prof);
}

}  // namespace rr
