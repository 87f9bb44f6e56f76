// rr_march.cu — ray-march kernel for sm_100a (render.rs:806-827 + :1226-1411 + quantiser).
//
// Work distribution: the iteration count per marched ray is heavy-tailed (1.3 % of
// rays run the 10 001-iteration cap and hold 74 % of all iterations, concentrated in the horizon
// rows), so tiles are handed out dynamically: each warp of a persistent grid pulls the next 8x4
// pixel tile from a global atomic counter when it finishes one.
#include "rr_kernels.h"
#include "rr_march.cuh"

namespace rr {

constexpr int MARCH_THREADS = 128;

__device__ __forceinline__ MarchView stage_march(const DevScene &G, float4 *smem, bool stage) {
    MarchView S;
    S.n_spheres = G.n_spheres;
    S.n_floors = G.n_floors;
    if (!stage) {
        S.sph = G.sph_m; S.sph_glow = G.sph_glow; S.sph_oi = G.sph_oi; S.flo_o = G.flo_o; S.flo_n = G.flo_n; S.flo_oi = G.flo_oi;
        return S;
    }
    float4 *sph = smem;
    float4 *flo_o = sph + G.n_spheres;
    float4 *flo_n = flo_o + G.n_floors;
    float *sph_glow = reinterpret_cast<float *>(flo_n + G.n_floors);
    int *sph_oi = reinterpret_cast<int *>(sph_glow + G.n_spheres);
    int *flo_oi = sph_oi + G.n_spheres;
    for (int i = threadIdx.x; i < G.n_spheres; i += blockDim.x) {
        sph[i] = G.sph_m[i];
        sph_glow[i] = G.sph_glow[i];
        sph_oi[i] = G.sph_oi[i];
    }
    for (int i = threadIdx.x; i < G.n_floors; i += blockDim.x) {
        flo_o[i] = G.flo_o[i];
        flo_n[i] = G.flo_n[i];
        flo_oi[i] = G.flo_oi[i];
    }
    __syncthreads();
    S.sph = sph; S.sph_glow = sph_glow; S.sph_oi = sph_oi; S.flo_o = flo_o; S.flo_n = flo_n; S.flo_oi = flo_oi;
    return S;
}

template <bool COUNT, bool F32OUT, bool STAGE, bool GLOW>
__global__ void __launch_bounds__(MARCH_THREADS)
march_kernel(const DevScene G, const FrameParams P, void *__restrict__ out, size_t row_stride, Counters *gcnt,
             unsigned *work, int fast_store) {
    extern __shared__ float4 rr_smem[];
    const MarchView S = stage_march(G, rr_smem, STAGE);

    const int W = P.xres, rows = P.local_rows;
    const int tiles_x = (W + 7) >> 3, tiles_y = (rows + 3) >> 2;
    const int ntiles = tiles_x * tiles_y;
    const int lane = threadIdx.x & 31;
    const int col = lane & 7, row = lane >> 3;
    Counters cnt = {};

    for (;;) {
        int tile = 0;
        if (lane == 0) tile = (int)atomicAdd(work, 1u);
        tile = __shfl_sync(0xffffffffu, tile, 0);
        if (tile >= ntiles) break;
        const int ty = tile / tiles_x, tx = tile - ty * tiles_x;
        const int x0 = tx << 3, ly0 = ty << 2;
        const int ix = x0 + col, ly = ly0 + row;
        const bool valid = ix < W && ly < rows;
        V3 c = mk(0.0f, 0.0f, 0.0f);
        if (valid) c = march_pixel<COUNT, GLOW>(G, S, P, ix, local_to_image_row(P, ly), cnt);
        if (F32OUT) {
            if (valid) {
                float *o = reinterpret_cast<float *>(out) + ((size_t)ly * W + ix) * 3;
                o[0] = c.x; o[1] = c.y; o[2] = c.z;
            }
        } else {
            const unsigned rgb = quantize(c.x) | (quantize(c.y) << 8) | (quantize(c.z) << 16);
            store_tile_rgb8(reinterpret_cast<uint8_t *>(out), row_stride, x0, ly0, W, rows, rgb, fast_store != 0);
        }
    }
    if (COUNT) flush_counters(cnt, gcnt);
}

static size_t march_smem_bytes(const DevScene &G) {
    return (size_t)G.n_spheres * (sizeof(float4) + sizeof(float) + sizeof(int)) +
           (size_t)G.n_floors * (2 * sizeof(float4) + sizeof(int)) + 16;
}

template <bool COUNT, bool F32OUT, bool STAGE, bool GLOW>
static cudaError_t launch_one(const DevScene &G, const FrameParams &P, void *d_out, size_t row_stride, Counters *d_cnt,
                              unsigned *d_work, cudaStream_t stream, const LaunchInfo &li, size_t smem) {
    auto kern = march_kernel<COUNT, F32OUT, STAGE, GLOW>;
    cudaError_t e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, MARCH_THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int tiles = ((P.xres + 7) / 8) * ((P.local_rows + 3) / 4);
    const int need = (tiles + (MARCH_THREADS / 32) - 1) / (MARCH_THREADS / 32);
    int grid = li.sm_count * per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    e = cudaMemsetAsync(d_work, 0, sizeof(unsigned), stream);
    if (e != cudaSuccess) return e;
    const int fast = (!F32OUT && (P.xres % 8 == 0) && (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3) == 0)) ? 1 : 0;
    kern<<<grid, MARCH_THREADS, smem, stream>>>(G, P, d_out, row_stride, d_cnt, d_work, fast);
    return cudaGetLastError();
}

template <bool COUNT, bool F32OUT>
static cudaError_t launch_two(const DevScene &G, const FrameParams &P, void *d_out, size_t row_stride, Counters *d_cnt,
                              unsigned *d_work, cudaStream_t stream, const LaunchInfo &li) {
    size_t smem = march_smem_bytes(G);
    const bool stage = smem <= li.smem_optin / 2;
    if (!stage) smem = 0;
    const bool glow = P.glow_enabled && G.n_glow > 0;
    if (stage) return glow ? launch_one<COUNT, F32OUT, true, true>(G, P, d_out, row_stride, d_cnt, d_work, stream, li, smem)
                           : launch_one<COUNT, F32OUT, true, false>(G, P, d_out, row_stride, d_cnt, d_work, stream, li, smem);
    return glow ? launch_one<COUNT, F32OUT, false, true>(G, P, d_out, row_stride, d_cnt, d_work, stream, li, smem)
                : launch_one<COUNT, F32OUT, false, false>(G, P, d_out, row_stride, d_cnt, d_work, stream, li, smem);
}

cudaError_t launch_march(const DevScene &G, const FrameParams &P, void *d_out, size_t row_stride, bool f32_out,
                         Counters *d_cnt, unsigned *d_work, cudaStream_t stream, const LaunchInfo &li) {
    if (P.xres <= 0 || P.local_rows <= 0) return cudaSuccess;
    if (d_cnt) return f32_out ? launch_two<true, true>(G, P, d_out, row_stride, d_cnt, d_work, stream, li)
                              : launch_two<true, false>(G, P, d_out, row_stride, d_cnt, d_work, stream, li);
    return f32_out ? launch_two<false, true>(G, P, d_out, row_stride, d_cnt, d_work, stream, li)
                   : launch_two<false, false>(G, P, d_out, row_stride, d_cnt, d_work, stream, li);
}

}  // namespace rr
