// march_cost_map.cpp — march steps per pixel of a frame, from the KERNEL SOURCE compiled for the CPU (tests/hostsim).
// Analysis tool (tools/march_cost_map.py): where the ray-march frame's work is, tile by tile. Not part of the product or the tests.
#include "../tests/hostsim/hostsim.cpp"
#include <thread>
#include <vector>
extern "C" int march_cost_map(const rr_scene_desc *desc, const rr_frame_params *params, unsigned *steps, int nthreads) {
    Flat f;
    flatten(desc, f);
    FrameParams P = to_dev(params, f.H);
    MarchView M{};
    M.sph = f.G.sph_m; M.sph_glow = f.G.sph_glow; M.sph_oi = f.G.sph_oi; M.flo_o = f.G.flo_o; M.flo_n = f.G.flo_n; M.flo_oi = f.G.flo_oi;
    M.n_spheres = f.G.n_spheres; M.n_floors = f.G.n_floors;
    std::vector<std::thread> th;
    for (int t = 0; t < nthreads; ++t)
        th.emplace_back([&, t] {
            for (int iy = t; iy < P.yres; iy += nthreads)
                for (int ix = 0; ix < P.xres; ++ix) {
                    Counters cnt{};
                    march_pixel<true, 1>(f.G, f.H, M, P, ix, iy, cnt);
                    steps[(size_t)iy * P.xres + ix] = (unsigned)cnt.march_steps;
                }
        });
    for (auto &x : th) x.join();
    return 0;
}
