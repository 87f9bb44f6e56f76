// simt_model.cpp — lane-utilisation model of the BVH trace instance, built from the KERNEL SOURCE on the CPU (tests/hostsim
// stand-ins). For every pixel it records the traversal steps (inner nodes visited, leaf spheres tested) of every ray, then
// replays warps over those records under two schedules:
//   A  one pixel per lane per tile (the round-1 kernel): a lane that has finished its pixel idles until the tile is done;
//   B  a pool of pixels per warp, a lane that finishes takes the next pixel of the pool (dynamic refill).
// Cost of one ray slot of a warp = max over busy lanes of (inner * CI) + max (leaf * CL) + CS; utilisation = lane work / (32 * cost).
// Build/run: tools/simt_model.py. Analysis tool only — not part of the product or of the tests.
#include <cstdio>
#include <vector>
struct RayRec { unsigned short inner, leaf; };
static std::vector<RayRec> *g_cur;
#define RR_MODEL_RAY() (g_cur->push_back(RayRec{0, 0}))
#define RR_MODEL_INNER() (++g_cur->back().inner)
#define RR_MODEL_LEAF(n) (g_cur->back().leaf += (unsigned short)(n))
#include "../tests/hostsim/hostsim.cpp"

extern "C" int model_run(const rr_scene_desc *desc, const rr_frame_params *params, const int *shapes, int n_shapes, double *out) {
    Flat f;
    flatten(desc, f);
    FrameParams P = to_dev(params, f.H);
    SceneView S{};
    S.sph = f.G.sph; S.sph_oi = f.G.sph_oi; S.flo_o = f.G.flo_o; S.flo_n = f.G.flo_n; S.flo_oi = f.G.flo_oi;
    S.n_spheres = f.G.n_spheres; S.n_floors = f.G.n_floors;
    S.bvh_a = f.G.bvh_a; S.bvh_b = f.G.bvh_b; S.bvh_w = f.G.bvh_w; S.bsph = f.G.bsph; S.bsph_oi = f.G.bsph_oi;
    S.n_bvh_nodes = f.G.n_bvh_nodes;
    if (!S.n_bvh_nodes) return 1;
    const int W = P.xres, Hh = P.yres;
    std::vector<std::vector<RayRec>> px((size_t)W * Hh);
    Counters cnt{};
    for (int iy = 0; iy < Hh; ++iy)
        for (int ix = 0; ix < W; ++ix) {
            g_cur = &px[(size_t)iy * W + ix];
            trace_pixel<false, true>(f.G, f.H, S, P, ix, iy, cnt);
        }
    const double CI = 34, CL = 30, CS = 160;  // instructions per inner node / leaf sphere / rest of a ray (from the SASS listing)
    // shapes: (tile_w, tile_h, refill) triples; tile_w*tile_h is the pool (32 for schedule A)
    for (int s = 0; s < n_shapes; ++s) {
        const int tw = shapes[3 * s], th = shapes[3 * s + 1], refill = shapes[3 * s + 2];
        double work = 0, cost = 0, rays = 0, slots = 0;
        for (int ty = 0; ty < Hh; ty += th)
            for (int tx = 0; tx < W; tx += tw) {
                std::vector<const std::vector<RayRec> *> pool;
                for (int y = ty; y < ty + th && y < Hh; ++y)
                    for (int x = tx; x < tx + tw && x < W; ++x) pool.push_back(&px[(size_t)y * W + x]);
                size_t next = 0;
                const std::vector<RayRec> *lane_px[32] = {};
                size_t lane_k[32] = {};
                auto refill_lane = [&](int l) { lane_px[l] = next < pool.size() ? pool[next++] : nullptr; lane_k[l] = 0; };
                for (;;) {
                    // lanes without a pixel take one (schedule A: only when ALL lanes are idle = next tile of 32)
                    bool any = false;
                    for (int l = 0; l < 32; ++l) any = any || lane_px[l];
                    if (refill || !any)
                        for (int l = 0; l < 32; ++l) if (!lane_px[l]) refill_lane(l);
                    int mi = 0, ml = 0, busy = 0;
                    for (int l = 0; l < 32; ++l) {
                        if (!lane_px[l]) continue;
                        const RayRec &r = (*lane_px[l])[lane_k[l]];
                        mi = r.inner > mi ? r.inner : mi;
                        ml = r.leaf > ml ? r.leaf : ml;
                        work += r.inner * CI + r.leaf * CL + CS;
                        ++busy;
                    }
                    if (!busy) break;
                    rays += busy; slots += 1;
                    cost += 32 * (mi * CI + ml * CL + CS);
                    for (int l = 0; l < 32; ++l)
                        if (lane_px[l] && ++lane_k[l] == lane_px[l]->size()) lane_px[l] = nullptr;
                }
            }
        out[4 * s] = work / cost; out[4 * s + 1] = rays / (32 * slots); out[4 * s + 2] = cost; out[4 * s + 3] = rays;
    }
    return 0;
}
