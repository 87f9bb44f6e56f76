// rr_trace.cu — ray-trace kernel for sm_100a (render.rs:806-827 + :993-1224, fused with the
// putpoint quantiser of main.rs:148-152).
//
// Mapping: a warp owns an 8x4 pixel tile (coherent primary rays, 24-byte row runs for the RGB8
// store); warps walk the tile list with a static grid stride from a persistent grid of
// (SM count x resident blocks), so the scene is staged into shared memory once per block.
// The kernel is FP32-ALU / divergence bound: compiled with -fmad=false for bit parity, no tensor
// cores, DRAM traffic = the framebuffer store only.
#include "rr_kernels.h"
#include "rr_trace.cuh"

namespace rr {

#ifndef RR_TRACE_THREADS
#define RR_TRACE_THREADS 256
#endif
constexpr int TRACE_THREADS = RR_TRACE_THREADS;
#ifndef RR_BYTE_STAGE
#define RR_BYTE_STAGE 1  // 0: A/B build that assembles the 24 words of a sub-run with shuffles (profiles/r3c_ab.txt: 0.163 -> 0.156 ms at 4K with 1)
#endif
#ifndef RR_TRACE_MIN_BLOCKS
#define RR_TRACE_MIN_BLOCKS 4
#endif
// BVH instances run ONE block of 1 024 threads per SM: the staged tree (37 KB for 1 024 spheres) and the per-thread
// traversal stacks (8 bytes x RR_BVH_SMEM_STACK per thread) then exist once per SM, which leaves room for all 32 resident
// warps the 64-register budget allows. With 256-thread blocks the same shared memory is needed per block and only three
// fit (ncu, profiles/r2e_trace_synthetic1024_4k.md: 37 % warps active instead of 50 %).
__host__ __device__ constexpr int trace_threads(bool bvh) { return bvh ? RR_TRACE_THREADS_BVH : TRACE_THREADS; }
__host__ __device__ constexpr int trace_min_blocks(bool bvh) { return bvh ? (RR_TRACE_MIN_BLOCKS * TRACE_THREADS) / RR_TRACE_THREADS_BVH : RR_TRACE_MIN_BLOCKS; }

template <typename T>
__device__ __forceinline__ void copy_list(T *dst, const T *src, int n) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) dst[i] = src[i];
}

// Stage what the scan loops read into shared memory: the list tails (objects beyond the SceneHead)
// for the brute-force scan, or the BVH nodes + leaf-ordered spheres. Small scenes stage nothing.
template <bool BVH>
__device__ __forceinline__ SceneView stage_scene(const DevScene &G, float4 *smem, bool stage) {
    SceneView S;
    S.n_spheres = G.n_spheres;
    S.n_floors = G.n_floors;
    S.sph = G.sph; S.sph_oi = G.sph_oi; S.flo_o = G.flo_o; S.flo_n = G.flo_n; S.flo_oi = G.flo_oi;
    S.bvh_a = G.bvh_a; S.bvh_b = G.bvh_b; S.bvh_w = G.bvh_w; S.bsph = G.bsph; S.bsph_oi = G.bsph_oi; S.n_bvh_nodes = G.n_bvh_nodes;
    // BVH instances: the first float4s of the dynamic shared memory are the per-thread traversal stacks (trace_smem_bytes)
    S.stk = reinterpret_cast<uint2 *>(smem);
    S.stk_stride = (int)blockDim.x;
    S.stk_s = (unsigned)__cvta_generic_to_shared(smem);
    S.bvh_w_s = 0u; S.bsph_s = 0u;
    if (BVH) smem += (RR_BVH_SMEM_STACK * trace_threads(true) * sizeof(uint2)) / sizeof(float4);
    if (!stage) return S;
    const int tf = max(G.n_floors - RR_HEAD_FLOORS, 0);
    const int ts = BVH ? 0 : max(G.n_spheres - RR_HEAD_SPHERES, 0);
    // ordered traversal: 4 float4 per inner node in ONE array (bvh_a region); depth-first arrays otherwise
    const int nb = !BVH ? 0 : (RR_BVH_ORDERED ? 2 * G.n_bvh_inner : G.n_bvh_nodes), bs = BVH ? G.n_spheres : 0;
    if (ts + tf + nb == 0) return S;
    float4 *p = smem;
    float4 *sph = p; p += ts;
    float4 *flo_o = p; p += tf;
    float4 *flo_n = p; p += tf;
    float4 *bvh_a = p; p += nb;
    float4 *bvh_b = p; p += nb;
    float4 *bsph = p; p += bs;
    int *q = reinterpret_cast<int *>(p);
    int *sph_oi = q; q += ts;
    int *flo_oi = q; q += tf;
    int *bsph_oi = q;
    copy_list(sph, G.sph + RR_HEAD_SPHERES, ts);
    copy_list(sph_oi, G.sph_oi + RR_HEAD_SPHERES, ts);
    copy_list(flo_o, G.flo_o + RR_HEAD_FLOORS, tf);
    copy_list(flo_n, G.flo_n + RR_HEAD_FLOORS, tf);
    copy_list(flo_oi, G.flo_oi + RR_HEAD_FLOORS, tf);
    if (RR_BVH_ORDERED) {
        copy_list(bvh_a, G.bvh_w, 2 * nb);  // bvh_a and bvh_b regions are adjacent: 4 * n_bvh_inner float4
    } else {
        copy_list(bvh_a, G.bvh_a, nb);
        copy_list(bvh_b, G.bvh_b, nb);
    }
    copy_list(bsph, G.bsph, bs);
    copy_list(bsph_oi, G.bsph_oi, bs);
    __syncthreads();
    // tail views are indexed with the global list index
    S.sph = sph - RR_HEAD_SPHERES; S.sph_oi = sph_oi - RR_HEAD_SPHERES;
    S.flo_o = flo_o - RR_HEAD_FLOORS; S.flo_n = flo_n - RR_HEAD_FLOORS; S.flo_oi = flo_oi - RR_HEAD_FLOORS;
    if (BVH) {
        S.bvh_a = bvh_a; S.bvh_b = bvh_b; S.bvh_w = bvh_a; S.bsph = bsph; S.bsph_oi = bsph_oi;
        S.bvh_w_s = (unsigned)__cvta_generic_to_shared(bvh_a);
        S.bsph_s = (unsigned)__cvta_generic_to_shared(bsph);
    }
    return S;
}

template <bool COUNT, bool F32OUT, bool STAGE, bool BVH, int TW, bool HEADONLY>
__global__ void __launch_bounds__(trace_threads(BVH), trace_min_blocks(BVH))
trace_kernel(const __grid_constant__ DevScene G, const __grid_constant__ SceneHead H, const __grid_constant__ FrameParams P,
             void *__restrict__ out, size_t row_stride, Counters *gcnt, int fast_store, float inv_tiles_x, const Signal sig) {
    extern __shared__ float4 rr_smem[];
    const SceneView S = stage_scene<BVH>(G, rr_smem, STAGE);

    // warp tile: TW x TH pixels. 8x4 (best ray coherence); 32x1 for placed output (one 96-byte run per store); 128x1 =
    // four 32x1 sub-tiles rendered one after the other and stored together as ONE 384-byte run (three full 128-byte
    // lines, 16 bytes per lane) for frames that live in a peer GPU's memory: the owner's NVLink ingress is what bounds
    // the 8-GPU frame, and full-line writes use it best.
    constexpr int SUB = TW == 128 ? 4 : 1;  // sub-tiles per tile
    constexpr int TWS = TW / SUB;           // sub-tile width
    constexpr int TH = 32 / TWS;
    __shared__ uint4 rr_wbuf[TW == 128 ? TRACE_THREADS / 32 : 1][24];
    const int W = P.xres, rows = P.local_rows;
    const int tiles_x = (W + TW - 1) / TW, tiles_y = (rows + TH - 1) / TH;
    const int ntiles = tiles_x * tiles_y;
    const int lane = threadIdx.x & 31;
    const int col = lane % TWS, row = lane / TWS;
    const int warps_per_block = blockDim.x >> 5;
    const int gw = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int nw = gridDim.x * warps_per_block;
    Counters cnt = {};

    // Tile schedule: static interleaved stride for the first part of the tile list (each warp samples the whole image, so
    // the correlated cost of neighbouring tiles — sky, floor, glass — is spread over all SMs), then a global queue hands
    // out the rest one tile at a time (same-address atomics: ~2 G/s on this part, a 4K frame has 259 k tiles). Tile
    // cost varies by 10x and more; with the static stride alone the busiest SM sub-partition ended ~15 % after the average
    // one (ncu: 41.7 % warps active of a possible 50 %). Measured static share, kernel ms at 4K default / 8K default /
    // 4K 1 024 spheres (profiles/r1_s2_tile_schedule.md): 16/16 0.249 / 0.955 / 2.78, 13/16 0.239 / 0.910 / 2.71,
    // 8/16 0.225 / 0.862 / 2.35, 6/16 0.225 / 0.864 / 2.06, 0/16 0.234 / 0.892 / 2.05. Contiguous chunks per warp
    // (guided self-scheduling) lose badly (0.35 / 1.28 / 4.0): neighbouring tiles cost alike.
    // Tried and dropped in round 2 (profiles/r2i_shard_time.txt, r2h_bench_n8_tail*.json): a smaller static share for small
    // launches and handing the last macro tiles out as their four 32x1 sub-tiles. Neither changes the time of a 1/8-frame
    // launch (0.105 ms against 0.082 ideal at 8K: that gap is not the tile schedule), and the 96-byte stores of the sub-tiles
    // cost the 8-GPU step 5 % of its NVLink store rate. Issuing the queue grab for the NEXT tile before the current one is rendered
    // (to hide the atomic's round trip: 1.2 % of warp time in ncu) is slower, 0.168 -> 0.177 ms at 4K (profiles/r2w_ab_trace.txt).
    constexpr int STATIC_16THS = BVH ? 6 : 8;
    const int n_static = sig.work ? (int)(((long long)ntiles * STATIC_16THS) >> 4) : ntiles;
    int nt = gw;
    for (;;) {
        if (sig.work && nt >= n_static) {  // warp-uniform: the static share is done (queue tiles are >= n_static, so it stays done)
            nt = 0;
            if (lane == 0) nt = n_static + (int)atomicAdd(sig.work, 1u);
            nt = __shfl_sync(0xffffffffu, nt, 0);
        }
        if (nt >= ntiles) break;
        const int tile = nt;
        nt = tile + nw;
        {
            // ty = tile / tiles_x without an integer division: float estimate + one-step correction
            // (exact for tile < 2^24; the launcher falls back to inv_tiles_x = 0 -> integer division above that)
            int ty;
            if (inv_tiles_x > 0.0f) {
                ty = (int)(((float)tile + 0.5f) * inv_tiles_x);
                const int r = tile - ty * tiles_x;
                ty += (r >= tiles_x) ? 1 : ((r < 0) ? -1 : 0);
            } else {
                ty = tile / tiles_x;
            }
            const int tx = tile - ty * tiles_x;
            const int x0 = tx * TW, ly0 = ty * TH;
            if constexpr (TW == 128) {
                // launcher guarantees: RGB8, W % 128 == 0, 16-byte aligned rows (so every lane's uint4 is aligned)
                const int ly = ly0;
                const int irow = local_to_image_row(P, ly);
                const int orow = P.placed ? irow : ly;
                const float4 prow = __ldg(&P.ptab[P.xres + irow]);  // one row: the same for the four sub-tiles
                unsigned *wb = reinterpret_cast<unsigned *>(rr_wbuf[threadIdx.x >> 5]);
#if RR_BYTE_STAGE
                const unsigned wb_s = (unsigned)__cvta_generic_to_shared(wb);
#endif
#pragma unroll 1
                for (int sub = 0; sub < SUB; ++sub) {
                    const V3 c = trace_pixel<COUNT, BVH, HEADONLY, BVH && STAGE && RR_BVH_ORDERED>(G, H, S, P, primary_dir_tab(P, __ldg(&P.ptab[x0 + sub * 32 + lane]), prow), cnt);
#if RR_BYTE_STAGE
                    // every lane puts its own three bytes into the warp's staging run (3 byte stores; the 4 lanes that share a
                    // 32-bit word serialise, which costs less than assembling words with two shuffles and a 64-bit shift)
                    const unsigned qr = quantize(c.x), qg = quantize(c.y), qb = quantize(c.z);
                    const unsigned a = wb_s + (unsigned)(sub * 96 + 3 * lane);
                    asm volatile("st.shared.u8 [%0], %1;\n\tst.shared.u8 [%0+1], %2;\n\tst.shared.u8 [%0+2], %3;" ::"r"(a), "r"(qr), "r"(qg), "r"(qb) : "memory");
#else
                    const unsigned rgb = quantize(c.x) | (quantize(c.y) << 8) | (quantize(c.z) << 16);
                    // word w of the 96-byte sub-run holds bytes 4w..4w+3 = pixels pa (and pa+1); lanes 0..23 own one word
                    const int pa = (4 * lane) / 3;
                    const unsigned va = __shfl_sync(0xffffffffu, rgb, min(pa, 31));
                    const unsigned vb = __shfl_sync(0xffffffffu, rgb, min(pa + 1, 31));
                    const unsigned long long both = (unsigned long long)va | ((unsigned long long)vb << 24);
                    if (lane < 24) wb[sub * 24 + lane] = (unsigned)(both >> (8 * ((4 * lane - 3 * pa) & 3)));
#endif
                }
                __syncwarp();
                if (lane < 24)
                    *reinterpret_cast<uint4 *>(reinterpret_cast<uint8_t *>(out) + (size_t)orow * row_stride + (size_t)x0 * 3 + 16 * lane) =
                        rr_wbuf[threadIdx.x >> 5][lane];
                __syncwarp();
            } else {
            const int ix = x0 + col, ly = ly0 + row;
            const bool valid = ix < W && ly < rows;
            V3 c = mk(0.0f, 0.0f, 0.0f);
            if (valid) c = trace_pixel<COUNT, BVH, HEADONLY, BVH && STAGE && RR_BVH_ORDERED>(G, H, S, P, ix, local_to_image_row(P, ly), cnt);
            if (F32OUT) {
                if (valid) {
                    float *o = reinterpret_cast<float *>(out) + ((size_t)(P.placed ? local_to_image_row(P, ly) : ly) * W + ix) * 3;
                    o[0] = c.x; o[1] = c.y; o[2] = c.z;
                }
            } else {
                const unsigned rgb = quantize(c.x) | (quantize(c.y) << 8) | (quantize(c.z) << 16);
                store_tile_rgb8<TWS>(reinterpret_cast<uint8_t *>(out), row_stride, x0, ly0, W, rows, rgb, fast_store != 0,
                                     P.placed ? local_to_image_row(P, ly) : ly);
            }
            }
        }
    }
    if (COUNT) flush_counters(cnt, gcnt);
    finish_launch(sig);
}

static size_t trace_stack_bytes(bool bvh) { return bvh ? (size_t)RR_BVH_SMEM_STACK * trace_threads(true) * sizeof(uint2) : 0; }
static size_t trace_smem_bytes(const DevScene &G, bool bvh) {
    const size_t tf = G.n_floors > RR_HEAD_FLOORS ? G.n_floors - RR_HEAD_FLOORS : 0;
    size_t b = tf * (2 * sizeof(float4) + sizeof(int)) + 16;
    const size_t node_f4 = RR_BVH_ORDERED ? (size_t)G.n_bvh_inner * 4 : (size_t)G.n_bvh_nodes * 2;
    if (bvh) return b + node_f4 * sizeof(float4) + (size_t)G.n_spheres * (sizeof(float4) + sizeof(int));
    const size_t ts = G.n_spheres > RR_HEAD_SPHERES ? G.n_spheres - RR_HEAD_SPHERES : 0;
    return b + ts * (sizeof(float4) + sizeof(int));
}

template <bool COUNT, bool F32OUT, bool STAGE, bool BVH, int TW, bool HEADONLY>
static cudaError_t launch_hd(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                             Counters *d_cnt, cudaStream_t stream, const LaunchInfo &li, size_t smem, const Signal &sig) {
    auto kern = trace_kernel<COUNT, F32OUT, STAGE, BVH, TW, HEADONLY>;
    constexpr int TH = TW == 128 ? 1 : 32 / TW;
    constexpr int THREADS = trace_threads(BVH);
    cudaError_t e;
    if (smem > 48 * 1024) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, THREADS, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    const int tiles_x = (P.xres + TW - 1) / TW;
    const long long tiles = (long long)tiles_x * ((P.local_rows + TH - 1) / TH);
    const long long need = (tiles + (THREADS / 32) - 1) / (THREADS / 32);
    long long grid = (long long)li.sm_count * per_sm;
    if (grid > need) grid = need;
    if (grid < 1) grid = 1;
    const int fast = (!F32OUT && (P.xres % TW == 0) && (row_stride % 4 == 0) && ((reinterpret_cast<uintptr_t>(d_out) & 3) == 0)) ? 1 : 0;
    const float inv_tx = tiles < (1 << 24) ? 1.0f / (float)tiles_x : 0.0f;
    kern<<<(unsigned)grid, THREADS, smem, stream>>>(G, H, P, d_out, row_stride, d_cnt, fast, inv_tx, sig);
    return cudaGetLastError();
}

// Head-only instance: the whole scene sits in the constant-bank SceneHead (the built-in scene: 1 floor + 4 spheres), the
// scan has no count checks and no tail loops, nothing is staged.
template <bool COUNT, bool F32OUT, bool STAGE, bool BVH, int TW>
static cudaError_t launch_tw(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                             Counters *d_cnt, cudaStream_t stream, const LaunchInfo &li, size_t smem, const Signal &sig) {
    if constexpr (!BVH && !COUNT && !F32OUT) {
        if (G.n_floors <= RR_HEAD_FLOORS && G.n_spheres <= RR_HEAD_SPHERES)
            return launch_hd<COUNT, F32OUT, false, BVH, TW, true>(G, H, P, d_out, row_stride, d_cnt, stream, li, 0, sig);
    }
    return launch_hd<COUNT, F32OUT, STAGE, BVH, TW, false>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig);
}

// Tile shape (measured, profiles/r1_s2_tile_schedule.md): 128x1 macro tiles (four 32x1 sub-tiles, one 384-byte store)
// whenever the frame allows it — they win locally too (4K 0.226 -> 0.217 ms, 8K 0.862 -> 0.807 ms: a quarter of the tile
// bookkeeping and queue traffic, full-line stores) — except for the BVH instances, whose incoherent secondary rays want
// the 8x4 footprint; 32x1 row tiles for other placed frames; 8x4 otherwise.
template <bool COUNT, bool F32OUT, bool STAGE, bool BVH>
static cudaError_t launch_one(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                              Counters *d_cnt, cudaStream_t stream, const LaunchInfo &li, size_t smem, const Signal &sig) {
    if constexpr (!F32OUT && !COUNT) {
        static const bool wide = [] { const char *e = getenv("RR_WIDE_TILES"); return e ? atoi(e) != 0 : true; }();
        // ... and only when there are enough of them to keep every resident warp busy (>= 4 per warp): the small chunk
        // launches of the host pipeline (rr_render_rgb8) need the finer 8x4 granularity
        const long long macro_tiles = (long long)(P.xres / 128) * P.local_rows;
        const bool enough = macro_tiles >= 4ll * li.sm_count * (RR_TRACE_MIN_BLOCKS * TRACE_THREADS / 32);
        // BVH instances always keep the 8x4 footprint, placed or not: their incoherent secondary rays lose far more to a
        // 128x1 / 32x1 footprint than the peer stores win (measured on 2 GPUs, 1 024 spheres at 4K: the placed step took
        // 1.78 ms with 128x1 tiles against a 1.23 ms kernel with 8x4 tiles; at ~2 ms per 25 MB frame the NVLink store rate is
        // irrelevant). This was the unexplained 0.53 efficiency of that scene on 8 GPUs in round 1.
        if (wide && enough && !BVH && P.xres % 128 == 0 && row_stride % 16 == 0 && (reinterpret_cast<uintptr_t>(d_out) & 15) == 0)
            return launch_tw<COUNT, F32OUT, STAGE, BVH, 128>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig);
        if (P.placed && !BVH && P.xres % 32 == 0)
            return launch_tw<COUNT, F32OUT, STAGE, BVH, 32>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig);
    }
    return launch_tw<COUNT, F32OUT, STAGE, BVH, 8>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig);
}

template <bool COUNT, bool F32OUT>
static cudaError_t launch_two(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                              Counters *d_cnt, cudaStream_t stream, const LaunchInfo &li, bool allow_bvh, const Signal &sig) {
    const bool bvh = allow_bvh && G.n_bvh_nodes > 0;
    size_t smem = trace_smem_bytes(G, bvh);
#ifndef RR_BVH_STAGE
#define RR_BVH_STAGE 1
#endif
    // keep >= 2 blocks per SM resident (a BVH instance with 1 024-thread blocks fills the SM with one)
    const size_t stage_limit = bvh && trace_min_blocks(true) == 1 ? li.smem_optin - 1024 : li.smem_optin / 2;
    const bool stage = (RR_BVH_STAGE || !bvh) && smem + trace_stack_bytes(bvh) <= stage_limit;
    if (!stage) smem = 0;
    smem += trace_stack_bytes(bvh);  // the traversal stacks are always there
    if (bvh) return stage ? launch_one<COUNT, F32OUT, true, true>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig)
                          : launch_one<COUNT, F32OUT, false, true>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig);
    return stage ? launch_one<COUNT, F32OUT, true, false>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig)
                 : launch_one<COUNT, F32OUT, false, false>(G, H, P, d_out, row_stride, d_cnt, stream, li, smem, sig);
}

cudaError_t launch_trace(const DevScene &G, const SceneHead &H, const FrameParams &P, void *d_out, size_t row_stride,
                         bool f32_out, Counters *d_cnt, cudaStream_t stream, const LaunchInfo &li, bool allow_bvh, const Signal &sig) {
    if (P.xres <= 0 || P.local_rows <= 0) return cudaSuccess;
    if (d_cnt) return f32_out ? launch_two<true, true>(G, H, P, d_out, row_stride, d_cnt, stream, li, allow_bvh, sig)
                              : launch_two<true, false>(G, H, P, d_out, row_stride, d_cnt, stream, li, allow_bvh, sig);
    return f32_out ? launch_two<false, true>(G, H, P, d_out, row_stride, d_cnt, stream, li, allow_bvh, sig)
                   : launch_two<false, false>(G, H, P, d_out, row_stride, d_cnt, stream, li, allow_bvh, sig);
}

}  // namespace rr
