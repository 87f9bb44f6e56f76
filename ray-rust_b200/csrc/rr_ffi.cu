// rr_ffi.cu — implementation of the C ABI declared in include/rr_ffi.h.
//
// Host side of the boundary: validates the POD scene, flattens RenderEnv.objects/materials
// (render.rs:658-659) into the SoA device layout of rr_device.cuh, owns the per-handle stream and
// buffers, launches the kernels and moves the frame to the caller. No CPU rendering path exists
// here: every render entry point either runs the CUDA kernels or fails.
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <vector>

#include "../../include/rr_ffi.h"
#include "rr_bvh.h"
#include "rr_kernels.h"

namespace {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg) {
    g_last_error = msg;
    return code;
}
int fail_cuda(cudaError_t e, const char *what) {
    char buf[512];
    snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorString(e), cudaGetErrorName(e));
    cudaGetLastError();  // clear sticky-free errors
    return fail(e == cudaErrorMemoryAllocation ? RR_ERR_OOM : RR_ERR_CUDA, buf);
}
#define CU(call)                                           \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return fail_cuda(e__, #call); \
    } while (0)

}  // namespace

// One in-flight host-facing render of a handle: its own stream pair, device frame and events. A handle owns RR_LANES of
// them, so concurrent calls on one handle (the web server renders from several threads, webserver.rs:268-280) overlap on
// the GPU instead of queueing behind one stream and one buffer, and rr_render_rgb8_async can keep several frames of an
// animation in flight (render.rs:926-989).
constexpr int RR_LANES = 4;
struct Lane {
    cudaStream_t stream = nullptr;
    cudaStream_t copy_stream = nullptr;
    void *d_out = nullptr;
    size_t d_out_cap = 0;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t chunk_ev[32] = {};
    bool busy = false;      // taken by a call
    bool pending = false;   // an async render is in flight on it (owned by its ticket)
    uint32_t gen = 0;       // ticket generation
};

// Primary-ray tables (rr_device.cuh, primary_dir_tab) depend on resolution, fov and camera rotation only. A handle keeps
// the tables of the last RR_PTABS distinct cameras: a launch whose camera is cached binds the table and launches nothing
// extra (a still camera, every chunk of a chunked frame, every band of a multi-GPU frame); a new camera takes the
// oldest slot and queues the 6 000-thread fill kernel in front of the render kernel on the same stream. A slot is
// rewritten RR_PTABS camera changes after it was bound, the same bound on overlapping launches as RR_SLOTS.
constexpr int RR_PTABS = 32;
struct PrimTable {
    float4 *d = nullptr;
    size_t cap = 0;  // entries
    int xres = -1, yres = -1;
    float key[6] = {};  // xfov, yfov, cam_rot[4], compared bitwise
};

// Profile-guided tile-row order of the ray-march kernel. A march frame's time is set by its longest warp tiles and by WHEN
// they start (1.7 % of the tiles hold ~80 % of the work, a single tile runs for up to ~2 ms); which rows
// hold them depends on the scene and the camera. Every march launch records the longest tile of each of its tile rows
// (RowProfile::cost, copied back asynchronously behind the kernel); the next launch of the same frame geometry on this
// handle — the next frame of an animation, the next request of the web front-end, the next step of a benchmark — serves
// the rows longest first (RowProfile::order, a permutation generated on the host from the last COMPLETED profile, uploaded
// on the launch's stream in front of the kernel). Scheduling only: every tile is rendered once, with the same arithmetic.
constexpr int RR_PROF_SLOTS = 8;       // launches of one handle whose profiles may be in flight at once
constexpr int RR_PROF_MAX_ROWS = 4096; // tile rows (16 384 pixel rows); larger launches are not profiled
struct ProfKey {
    int mode, xres, yres, local_rows, row0, band_rows, band_index, band_count, band_span;
    bool operator==(const ProfKey &o) const { return memcmp(this, &o, sizeof *this) == 0; }
};
struct MarchProfile {
    std::mutex mu;
    bool ready = false, failed = false;
    unsigned *d_cost = nullptr, *h_cost = nullptr;  // [RR_PROF_SLOTS][RR_PROF_MAX_ROWS]; h_* page-locked
    int *d_order = nullptr, *h_order = nullptr;
    cudaEvent_t ev[RR_PROF_SLOTS] = {};
    int state[RR_PROF_SLOTS] = {};  // 0 free, 1 reserved by a launch in progress, 2 copy-back queued (ev recorded)
    ProfKey key[RR_PROF_SLOTS] = {};
    int rows[RR_PROF_SLOTS] = {};
    unsigned long long seq[RR_PROF_SLOTS] = {};
    unsigned long long next_seq = 1, order_seq = 0;
    std::vector<int> order;  // the current profile: tile rows, longest first
    ProfKey order_key{};
};

struct rr_scene {
    int device = 0;
    MarchProfile prof;
    rr::DevScene G{};
    rr::SceneHead H{};
    rr::LaunchInfo li{};
    std::vector<void *> allocs;
    std::mutex mu;               // guards lanes' busy flags, last_ms, culling
    std::condition_variable cv;
    Lane lanes[RR_LANES];
    std::mutex cnt_mu;           // rr_render_count: one counter block per handle
    rr::Counters *d_cnt = nullptr;
    unsigned *d_work = nullptr;  // RR_SLOTS x (work, done, -, -): per-launch tile queue word + block counter
    std::atomic<unsigned> launch_seq{0};
    std::mutex ptab_mu;
    PrimTable ptabs[RR_PTABS];
    unsigned ptab_next = 0;
    float last_ms = 0.0f;
    bool timed = false;
    bool culling = true;
    bool profile_rows = true;    // march launches record / use the tile-row profile (RR_ROW_PROFILE=0 turns it off)
};
constexpr unsigned RR_SLOTS = 64;

namespace {

template <typename T>
int upload(rr_scene *s, const std::vector<T> &h, const T **out) {
    void *d = nullptr;
    const size_t bytes = (h.empty() ? 1 : h.size()) * sizeof(T);
    CU(cudaMalloc(&d, bytes));
    s->allocs.push_back(d);
    if (!h.empty()) CU(cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice));
    *out = reinterpret_cast<const T *>(d);
    return RR_OK;
}

int32_t rows_of(const rr_frame_params *p) {
    const int cnt = p->band_count <= 1 ? 1 : p->band_count;
    if (cnt == 1) return p->yres;
    const int br = p->band_rows <= 0 ? 1 : p->band_rows;
    // bands b = band_index, band_index+cnt, ...; each has br rows except a clipped last one
    int64_t rows = 0;
    const int span = p->band_span <= 1 ? 1 : p->band_span;
    for (int64_t b0 = p->band_index; b0 * br < p->yres; b0 += cnt)
        for (int64_t b = b0; b < b0 + span && b * br < p->yres; ++b) {
            int64_t lo = b * br, hi = lo + br;
            if (hi > p->yres) hi = p->yres;
            rows += hi - lo;
        }
    return (int32_t)rows;
}

int check_params(const rr_frame_params *p) {
    if (!p) return fail(RR_ERR_BAD_ARG, "params is null");
    if (p->xres < 0 || p->yres < 0) return fail(RR_ERR_BAD_ARG, "negative resolution");
    if ((int64_t)p->xres * (int64_t)p->yres > (int64_t)1 << 31) return fail(RR_ERR_BAD_ARG, "frame larger than 2^31 pixels");
    if (p->band_count > 1 && (p->band_index < 0 || p->band_index >= p->band_count || p->band_rows <= 0 ||
                              (p->band_span > 1 && p->band_index + p->band_span > p->band_count)))
        return fail(RR_ERR_BAD_ARG, "bad row-band parameters");
    if (p->bg_kind != RR_BG_BGCOLOR && p->bg_kind != RR_BG_BLACK) return fail(RR_ERR_BAD_ARG, "unknown bg_kind");
    if (p->max_refractions > rr::RR_MAX_STACK_HOST)
        return fail(RR_ERR_UNSUPPORTED, "max_refractions above the device recursion stack (32)");
    return RR_OK;
}

rr::FrameParams to_dev(const rr_frame_params *p, const rr_scene *s) {
    rr::FrameParams d{};
    d.xres = p->xres; d.yres = p->yres; d.xfov = p->xfov; d.yfov = p->yfov;
    for (int k = 0; k < 3; ++k) { d.cam_pos[k] = p->cam_position[k]; d.light[k] = p->light[k]; }
    for (int k = 0; k < 4; ++k) d.cam_rot[k] = p->cam_rotation[k];
    d.use_raymarching = p->use_raymarching; d.glow_enabled = p->glow_enabled; d.glow_effect = p->glow_effect;
    d.max_reflections = p->max_reflections; d.max_refractions = p->max_refractions;
    d.bg_kind = p->bg_kind;
    d.band_count = p->band_count <= 1 ? 1 : p->band_count;
    d.band_rows = p->band_rows <= 0 ? 1 : p->band_rows;
    d.band_index = d.band_count == 1 ? 0 : p->band_index;
    d.band_span = d.band_count == 1 || p->band_span <= 1 ? 1 : p->band_span;
    d.local_rows = rows_of(p);
    d.row0 = 0;
    if (s) rr::finish_frame_params(d, s->H);
    return d;
}

// Binds P.ptab to the cached primary-ray table of this camera, filling a slot first (on `st`) when the camera is new.
int bind_prim_table(rr_scene *s, rr::FrameParams &P, cudaStream_t st) {
    P.ptab = nullptr;
    if (P.xres <= 0 || P.yres <= 0) return RR_OK;
    float key[6] = {P.xfov, P.yfov, P.cam_rot[0], P.cam_rot[1], P.cam_rot[2], P.cam_rot[3]};
    std::lock_guard<std::mutex> lk(s->ptab_mu);
    for (PrimTable &t : s->ptabs)
        if (t.d && t.xres == P.xres && t.yres == P.yres && memcmp(t.key, key, sizeof key) == 0) { P.ptab = t.d; return RR_OK; }
    PrimTable &t = s->ptabs[s->ptab_next++ % RR_PTABS];
    const size_t need = (size_t)P.xres + (size_t)P.yres;
    if (t.cap < need) {
        if (t.d) CU(cudaFree(t.d));  // (synchronises the device: nothing still reads the old table)
        t.d = nullptr; t.cap = 0; t.xres = -1;
        CU(cudaMalloc(&t.d, need * sizeof(float4)));
        t.cap = need;
    }
    t.xres = -1;  // not valid until the fill kernel is queued
    cudaError_t e = rr::launch_prim_table(P, t.d, st);
    if (e != cudaSuccess) return fail_cuda(e, "primary-ray table kernel");
    t.xres = P.xres; t.yres = P.yres;
    memcpy(t.key, key, sizeof key);
    P.ptab = t.d;
    return RR_OK;
}

// ---- march profile (see MarchProfile) -----------------------------------------------------------------------------
bool prof_init(MarchProfile &m) {  // under m.mu
    if (m.ready) return true;
    if (m.failed) return false;
    const size_t n = (size_t)RR_PROF_SLOTS * RR_PROF_MAX_ROWS;
    bool ok = cudaMalloc(reinterpret_cast<void **>(&m.d_cost), n * sizeof(unsigned)) == cudaSuccess &&
              cudaMalloc(reinterpret_cast<void **>(&m.d_order), n * sizeof(int)) == cudaSuccess &&
              cudaHostAlloc(reinterpret_cast<void **>(&m.h_cost), n * sizeof(unsigned), cudaHostAllocDefault) == cudaSuccess &&
              cudaHostAlloc(reinterpret_cast<void **>(&m.h_order), n * sizeof(int), cudaHostAllocDefault) == cudaSuccess;
    for (int k = 0; ok && k < RR_PROF_SLOTS; ++k) ok = cudaEventCreateWithFlags(&m.ev[k], cudaEventDisableTiming) == cudaSuccess;
    if (!ok) { cudaGetLastError(); m.failed = true; return false; }  // a hint only: render without it
    m.ready = true;
    return true;
}
void prof_free(MarchProfile &m) {
    if (m.d_cost) cudaFree(m.d_cost);
    if (m.d_order) cudaFree(m.d_order);
    if (m.h_cost) cudaFreeHost(m.h_cost);
    if (m.h_order) cudaFreeHost(m.h_order);
    for (auto &e : m.ev) if (e) cudaEventDestroy(e);
}
void prof_harvest(MarchProfile &m, int k) {  // under m.mu; slot k's copy-back has completed
    m.state[k] = 0;
    if (m.seq[k] < m.order_seq) return;  // a newer launch has already been harvested
    const unsigned *c = m.h_cost + (size_t)k * RR_PROF_MAX_ROWS;
    const int n = m.rows[k];
    m.order.resize((size_t)n);
    for (int i = 0; i < n; ++i) m.order[(size_t)i] = i;
    std::stable_sort(m.order.begin(), m.order.end(), [c](int a, int b) { return c[a] > c[b]; });
    m.order_key = m.key[k];
    m.order_seq = m.seq[k];
}
// Before a march launch: picks a profile slot, queues (on `st`) the upload of the row order when a completed profile of this
// geometry exists and the clearing of the slot's cost words. Returns the slot, or -1 (launch without profile).
int prof_begin(rr_scene *s, const rr::FrameParams &P, cudaStream_t st, rr::RowProfile *out) {
    *out = rr::RowProfile{nullptr, nullptr};
    const int tiles_y = (P.local_rows + 3) / 4;
    if (tiles_y < 8 || tiles_y > RR_PROF_MAX_ROWS) return -1;
    MarchProfile &m = s->prof;
    std::lock_guard<std::mutex> lk(m.mu);
    if (!prof_init(m)) return -1;
    for (int k = 0; k < RR_PROF_SLOTS; ++k)
        if (m.state[k] == 2 && cudaEventQuery(m.ev[k]) == cudaSuccess) prof_harvest(m, k);
    cudaGetLastError();  // (cudaErrorNotReady of the queries)
    int slot = -1;
    for (int k = 0; k < RR_PROF_SLOTS && slot < 0; ++k)
        if (m.state[k] == 0) slot = k;
    if (slot < 0) {
        // RR_PROF_SLOTS launches of this handle are in flight (a caller queueing frames far ahead on its stream): wait for
        // the oldest one, like rr_render_rgb8_async waits for a lane. Bounds the run-ahead to a few frames of GPU work.
        int oldest = -1;
        for (int k = 0; k < RR_PROF_SLOTS; ++k)
            if (m.state[k] == 2 && (oldest < 0 || m.seq[k] < m.seq[oldest])) oldest = k;
        if (oldest < 0 || cudaEventSynchronize(m.ev[oldest]) != cudaSuccess) { cudaGetLastError(); return -1; }
        prof_harvest(m, oldest);
        slot = oldest;
    }
    const ProfKey key{P.use_raymarching ? 1 : 0, P.xres, P.yres, P.local_rows, P.row0, P.band_rows, P.band_index, P.band_count, P.band_span};
    unsigned *d_cost = m.d_cost + (size_t)slot * RR_PROF_MAX_ROWS;
    if (cudaMemsetAsync(d_cost, 0, (size_t)tiles_y * sizeof(unsigned), st) != cudaSuccess) { cudaGetLastError(); return -1; }
    out->cost = d_cost;
    if (m.order_seq && m.order_key == key && (int)m.order.size() == tiles_y) {
        int *h = m.h_order + (size_t)slot * RR_PROF_MAX_ROWS, *d = m.d_order + (size_t)slot * RR_PROF_MAX_ROWS;
        memcpy(h, m.order.data(), (size_t)tiles_y * sizeof(int));
        if (cudaMemcpyAsync(d, h, (size_t)tiles_y * sizeof(int), cudaMemcpyHostToDevice, st) == cudaSuccess) out->order = d;
        else cudaGetLastError();
    }
    m.key[slot] = key;
    m.rows[slot] = tiles_y;
    m.seq[slot] = m.next_seq++;
    m.state[slot] = 1;  // reserved; prof_end records the event (on failure the slot is released there)
    return slot;
}
// After the launch: queue the copy-back of the cost words and the event that marks the slot complete.
void prof_end(rr_scene *s, int slot, cudaStream_t st, bool launched) {
    if (slot < 0) return;
    MarchProfile &m = s->prof;
    std::lock_guard<std::mutex> lk(m.mu);
    bool ok = launched;
    ok = ok && cudaMemcpyAsync(m.h_cost + (size_t)slot * RR_PROF_MAX_ROWS, m.d_cost + (size_t)slot * RR_PROF_MAX_ROWS,
                               (size_t)m.rows[slot] * sizeof(unsigned), cudaMemcpyDeviceToHost, st) == cudaSuccess;
    ok = ok && cudaEventRecord(m.ev[slot], st) == cudaSuccess;
    if (ok) {
        m.state[slot] = 2;
    } else {
        // nothing will complete this slot's event; whatever was queued for it only touches the slot's own buffers, which
        // stay allocated until the handle is destroyed. The slot goes back unharvested.
        cudaGetLastError();
        m.state[slot] = 0;
    }
}

int launch(rr_scene *s, const rr::FrameParams &P_in, void *d_out, size_t row_stride, bool f32, rr::Counters *d_cnt,
           cudaStream_t st, unsigned *d_flag = nullptr, unsigned epoch = 0) {
    rr::FrameParams P = P_in;
    if (!P.use_raymarching) {
        const int rc = bind_prim_table(s, P, st);
        if (rc != RR_OK) return rc;
    }
    // Each launch takes the next of RR_SLOTS (work, done) word pairs: the kernels' tile queue and block counter reset
    // themselves in the block that finishes last, so no memset is queued on the stream, and kernels of one handle running
    // concurrently on different streams (ray-trace AND ray-march mode) never share a queue unless more than RR_SLOTS of
    // them overlap. Both kernels publish the completion signal of a placed multi-GPU render themselves (last block); an
    // empty shard falls back to a one-thread publisher queued on the same stream.
    unsigned *slot = s->d_work + 4 * (s->launch_seq.fetch_add(1u, std::memory_order_relaxed) % RR_SLOTS);
    const rr::Signal sig{slot, slot + 1, d_flag, epoch};
    const bool fused = d_flag && P.xres > 0 && P.local_rows > 0;
    cudaError_t e;
    if (P.use_raymarching) {
        rr::RowProfile prof{nullptr, nullptr};
        const int pslot = (P.xres > 0 && P.local_rows > 0 && s->profile_rows) ? prof_begin(s, P, st, &prof) : -1;
        e = rr::launch_march(s->G, s->H, P, d_out, row_stride, f32, d_cnt, sig, st, s->li, s->culling, prof);
        prof_end(s, pslot, st, e == cudaSuccess);
    } else {
        e = rr::launch_trace(s->G, s->H, P, d_out, row_stride, f32, d_cnt, st, s->li, s->culling, sig);
    }
    if (e == cudaSuccess && d_flag && !fused) e = rr::launch_signal(sig, st);
    if (e != cudaSuccess) return fail_cuda(e, "kernel launch");
    return RR_OK;
}

// ---- lanes --------------------------------------------------------------------------------------------------------
Lane *acquire_lane(rr_scene *s) {
    std::unique_lock<std::mutex> lk(s->mu);
    for (;;) {
        for (Lane &l : s->lanes)
            if (!l.busy) { l.busy = true; return &l; }
        s->cv.wait(lk);
    }
}
void release_lane(rr_scene *s, Lane *l, float ms, bool timed) {
    {
        std::lock_guard<std::mutex> lk(s->mu);
        l->busy = false;
        l->pending = false;
        l->gen += 1;
        if (timed) { s->last_ms = ms; s->timed = true; }
    }
    s->cv.notify_one();
}
// Scope guard of a blocking call: whatever way the call returns, nothing of it is still running on the lane's streams
// (no DMA into the caller's buffer, no kernel reading d_out) when the lane goes back to the pool.
struct LaneHold {
    rr_scene *s; Lane *l; float ms = 0.0f; bool timed = false; bool keep = false;
    LaneHold(rr_scene *s_) : s(s_), l(acquire_lane(s_)) {}
    ~LaneHold() {
        if (keep) return;  // handed over to an async ticket
        if (!timed) { cudaStreamSynchronize(l->copy_stream); cudaStreamSynchronize(l->stream); cudaGetLastError(); }
        release_lane(s, l, ms, timed);
    }
};


using rr::Bvh;
using rr::build_bvh;


// How many row chunks a host-bound frame is cut into (kernel of chunk k+1 overlaps the copy of chunk k).
// Defaults measured on B200 + PCIe Gen5 (profiles/r1e_e2e_chunks.md); RR_E2E_MAX_CHUNKS / RR_E2E_CHUNK_KB override.
int pick_chunks(size_t bytes, const rr::FrameParams &P, const rr_scene *s) {
    static const int max_chunks = [] { const char *e = getenv("RR_E2E_MAX_CHUNKS"); int v = e ? atoi(e) : 16; return v < 1 ? 1 : (v > 32 ? 32 : v); }();
    static const size_t chunk_bytes = [] { const char *e = getenv("RR_E2E_CHUNK_KB"); long v = e ? atol(e) : 1536; return (size_t)(v < 64 ? 64 : v) << 10; }();
    // Chunking pays only while the kernel is shorter than the copy. Long kernels (ray marching with its
    // heavy-tailed rows, scenes with many objects) lose more to per-chunk tails than the overlap wins.
    if (P.use_raymarching || s->G.n_objects > 64) return 1;
    long long n = (long long)(bytes / chunk_bytes);
    if (n < 1) n = 1;
    if (n > max_chunks) n = max_chunks;
    return (int)n;
}

// Row counts of the chunks (whole 4-row tiles): a geometric plan. The first chunk is rows/d, so that the first copy starts
// after ~1/d of the kernel work, every next one g times larger, so that the copy engine sees ~9 copies instead of 16
// (a ~1.5 us gap each) while kernel k+1 (g x the rows at ~150 GB/s-equivalent) still ends before copy k (56 GB/s) does.
// Measured against the uniform 16 x 1.5 MB plan (tools/ab_e2e_geom.py): 4K 0.523-0.531 -> 0.496-0.498 ms, 8K 1.88-1.92 ->
// 1.86 ms with d = 48, g = 1.5 (round 1, 0.22 ms kernel); with the 0.156 ms kernel of round 2 a smaller first chunk and a
// steeper plan win: d = 96, g = 1.7 gives 0.492-0.495 ms / 1.81-1.82 ms (profiles/r3e_e2e_geom.txt; raw copy 0.441 / 1.765).
// RR_E2E_GEOM="d,g" overrides, "0" selects the uniform plan.
int plan_chunks(int rows, int nchunk, int *plan) {
    static const double geom_d = [] { const char *e = getenv("RR_E2E_GEOM"); return e ? atof(e) : 96.0; }();
    static const double geom_g = [] { const char *e = getenv("RR_E2E_GEOM"); const char *c = e ? strchr(e, ',') : nullptr; return c ? atof(c + 1) : 1.7; }();
    int n = 0, done = 0;
    if (geom_d >= 2.0 && nchunk > 1) {
        double want = rows / geom_d;
        while (done < rows && n < 31) {
            int r = ((int)want + 3) & ~3;
            if (r < 4) r = 4;
            if (r > rows - done) r = rows - done;
            plan[n++] = r;
            done += r;
            want *= geom_g;
        }
        if (done < rows) plan[n++] = rows - done;
        return n;
    }
    int chunk_rows = (rows + nchunk - 1) / nchunk;
    chunk_rows = (chunk_rows + 3) & ~3;
    for (; done < rows; done += chunk_rows) plan[n++] = rows - done < chunk_rows ? rows - done : chunk_rows;
    return n;
}

// Long kernels (the same predicate for which chunking does not pay): when the caller's frame is page-locked, mapped host
// memory (rr_host_alloc, rr_host_register, cudaHostAlloc), the kernel stores its RGB8 tiles straight into it over PCIe.
// The copy disappears behind the kernel: at >= 2.7 ms per 24.9 MB frame the stores need < 10 GB/s of the link
// (profiles/r1_s2_zero_copy.md: config 4 e2e 3.23 -> 2.80 ms, config 3 8.14 -> 7.78 ms). Short kernels keep the chunked
// device-buffer + DMA pipeline: 24-byte tile rows reach only 18 GB/s as PCIe writes.
bool long_kernel(const rr::FrameParams &P, const rr_scene *s) { return P.use_raymarching || s->G.n_objects > 64; }

bool mapped_host_alias(const void *host, size_t row_stride, void **dev) {
    static const bool enabled = [] { const char *e = getenv("RR_E2E_ZERO_COPY"); return e ? atoi(e) != 0 : true; }();
    if (!enabled || (row_stride & 3) || (reinterpret_cast<uintptr_t>(host) & 3)) return false;
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return false; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return false;
    *dev = a.devicePointer;
    return true;
}

int ensure_out(Lane *l, size_t bytes) {
    if (bytes <= l->d_out_cap) return RR_OK;
    if (l->d_out) cudaFree(l->d_out);
    l->d_out = nullptr;
    l->d_out_cap = 0;
    CU(cudaMalloc(&l->d_out, bytes));
    l->d_out_cap = bytes;
    return RR_OK;
}

// Timed launch on a lane: events around the launch(es), stream sync, elapsed ms into the hold.
int finish_timed(LaneHold &h) {
    CU(cudaStreamSynchronize(h.l->copy_stream));
    CU(cudaStreamSynchronize(h.l->stream));
    CU(cudaEventElapsedTime(&h.ms, h.l->ev0, h.l->ev1));
    h.timed = true;
    return RR_OK;
}

// Enqueue one host-bound RGB8 frame on a lane (everything rr_render_rgb8 does before it waits): either the kernel stores
// straight into the caller's page-locked frame (long kernels) or the frame is rendered in row chunks on the lane's
// stream with each chunk's device-to-host copy queued on the lane's copy stream behind its kernel.
int enqueue_rgb8(rr_scene *s, Lane *l, const rr::FrameParams &P, uint8_t *out, size_t row_stride);


}  // namespace

extern "C" {

int rr_abi_version(void) { return RR_ABI_VERSION; }

#ifndef RR_SRC_HASH
#define RR_SRC_HASH "unknown"
#endif
// "rr_src_hash=<sha256/32 of the sources and flags this binary was built from>": read back by build.py / ffi.py
const char *rr_build_info(void) { return "rr_src_hash=" RR_SRC_HASH " arch=sm_100a fmad=false"; }

const char *rr_last_error(void) { return g_last_error.c_str(); }

int rr_device_count(int *count) {
    if (!count) return fail(RR_ERR_BAD_ARG, "count is null");
    CU(cudaGetDeviceCount(count));
    return RR_OK;
}

int rr_scene_destroy(rr_scene *s) {
    if (!s) return RR_OK;
    cudaSetDevice(s->device);
    for (Lane &l : s->lanes) {
        if (l.stream) cudaStreamSynchronize(l.stream);
        if (l.copy_stream) cudaStreamSynchronize(l.copy_stream);
    }
    cudaDeviceSynchronize();  // launches on caller streams may still be writing the profile buffers
    for (void *p : s->allocs) cudaFree(p);
    prof_free(s->prof);
    for (PrimTable &t : s->ptabs) if (t.d) cudaFree(t.d);
    for (Lane &l : s->lanes) {
        if (l.d_out) cudaFree(l.d_out);
        if (l.ev0) cudaEventDestroy(l.ev0);
        if (l.ev1) cudaEventDestroy(l.ev1);
        for (auto &e : l.chunk_ev) if (e) cudaEventDestroy(e);
        if (l.stream) cudaStreamDestroy(l.stream);
        if (l.copy_stream) cudaStreamDestroy(l.copy_stream);
    }
    cudaGetLastError();
    delete s;
    return RR_OK;
}

int rr_scene_create(const rr_scene_desc *desc, int device, rr_scene **out) {
    if (!desc || !out) return fail(RR_ERR_BAD_ARG, "desc/out is null");
    *out = nullptr;
    if (desc->n_objects && !desc->objects) return fail(RR_ERR_BAD_ARG, "objects is null");
    if (desc->n_materials && !desc->materials) return fail(RR_ERR_BAD_ARG, "materials is null");
    if (desc->n_textures && !desc->textures) return fail(RR_ERR_BAD_ARG, "textures is null");
    if (desc->n_objects >= (1u << 22)) return fail(RR_ERR_UNSUPPORTED, "more than 2^22 objects (object index field of the device recursion stack)");
    // validate indices and enums before touching the device
    for (uint32_t i = 0; i < desc->n_materials; ++i) {
        const rr_material &m = desc->materials[i];
        if (m.pattern < RR_SOLID || m.pattern > RR_REPEATED_GRADATION) return fail(RR_ERR_BAD_ARG, "unknown pattern");
        if (m.texture_filter != RR_NEAREST && m.texture_filter != RR_BILINEAR) return fail(RR_ERR_BAD_ARG, "unknown texture_filter");
        if (m.texture >= 0 && (uint32_t)m.texture >= desc->n_textures) return fail(RR_ERR_BAD_ARG, "texture index out of range");
    }
    for (uint32_t i = 0; i < desc->n_textures; ++i) {
        const rr_texture &t = desc->textures[i];
        if (!t.rgb8 || t.width == 0 || t.height == 0) return fail(RR_ERR_BAD_ARG, "empty texture");
    }
    for (uint32_t i = 0; i < desc->n_objects; ++i) {
        const rr_object &o = desc->objects[i];
        if (o.kind != RR_SPHERE && o.kind != RR_FLOOR) return fail(RR_ERR_BAD_ARG, "unknown object kind");
        if (o.uvmap < RR_UV_XY || o.uvmap > RR_UV_LL) return fail(RR_ERR_BAD_ARG, "unknown uvmap");
        if (o.material < 0 || (uint32_t)o.material >= desc->n_materials) return fail(RR_ERR_BAD_ARG, "material index out of range");
    }
    int ndev = 0;
    CU(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return fail(RR_ERR_BAD_ARG, "device index out of range");
    CU(cudaSetDevice(device));

    rr_scene *s = new (std::nothrow) rr_scene();
    if (!s) return fail(RR_ERR_OOM, "host allocation failed");
    s->device = device;
    {
        const char *e = getenv("RR_ROW_PROFILE");
        s->profile_rows = !(e && atoi(e) == 0);
    }
    int rc = RR_OK;
    auto bail = [&](int code) { rr_scene_destroy(s); return code; };

    std::vector<float4> sph, sph_m, flo_o, flo_n, obj_a, obj_n;
    std::vector<float> sph_glow;
    std::vector<int> sph_oi, flo_oi;
    std::vector<int4> obj_b;
    int n_glow = 0;
    for (uint32_t i = 0; i < desc->n_objects; ++i) {
        const rr_object &o = desc->objects[i];
        const rr_material &m = desc->materials[o.material];
        if (m.glow_dist != 0.0f) n_glow++;
        obj_a.push_back(make_float4(o.org[0], o.org[1], o.org[2], o.r));
        obj_n.push_back(make_float4(o.face_normal[0], o.face_normal[1], o.face_normal[2], 0.0f));
        obj_b.push_back(make_int4(o.kind, o.uvmap, o.material, 0));
        if (o.kind == RR_SPHERE) {
            const float rr2 = o.r * o.r;  // the single f32 product of render.rs:456
            sph.push_back(make_float4(o.org[0], o.org[1], o.org[2], rr2));
            sph_m.push_back(make_float4(o.org[0], o.org[1], o.org[2], o.r));
            sph_glow.push_back(m.glow_dist);
            sph_oi.push_back((int)i);
        } else {
            flo_o.push_back(make_float4(o.org[0], o.org[1], o.org[2], m.glow_dist));
            flo_n.push_back(make_float4(o.face_normal[0], o.face_normal[1], o.face_normal[2], 0.0f));
            flo_oi.push_back((int)i);
        }
    }
    std::vector<rr::DevTexture> tex;
    for (uint32_t i = 0; i < desc->n_textures; ++i) {
        const rr_texture &t = desc->textures[i];
        std::vector<uint8_t> px(t.rgb8, t.rgb8 + (size_t)t.width * t.height * 3);
        const uint8_t *d = nullptr;
        if ((rc = upload(s, px, &d)) != RR_OK) return bail(rc);
        tex.push_back(rr::DevTexture{d, t.width, t.height});
    }
    std::vector<rr::DevMaterial> mats;
    for (uint32_t i = 0; i < desc->n_materials; ++i) {
        const rr_material &m = desc->materials[i];
        rr::DevMaterial d{};
        for (int k = 0; k < 3; ++k) { d.diffuse[k] = m.diffuse[k]; d.specular[k] = m.specular[k]; }
        d.pn = m.pn; d.t = m.t; d.n = m.n; d.glow_dist = m.glow_dist;
        d.pattern = m.pattern; d.pattern_scale = m.pattern_scale; d.pattern_angle_scale = m.pattern_angle_scale;
        d.texture = m.texture; d.texture_filter = m.texture_filter;
        mats.push_back(d);
    }
    rr::DevScene &G = s->G;
    G.n_spheres = (int)sph.size(); G.n_floors = (int)flo_o.size();
    G.n_objects = (int)desc->n_objects; G.n_materials = (int)desc->n_materials; G.n_glow = n_glow;
    if ((rc = upload(s, sph, &G.sph)) || (rc = upload(s, sph_m, &G.sph_m)) || (rc = upload(s, sph_glow, &G.sph_glow)) ||
        (rc = upload(s, sph_oi, &G.sph_oi)) || (rc = upload(s, flo_o, &G.flo_o)) || (rc = upload(s, flo_n, &G.flo_n)) ||
        (rc = upload(s, flo_oi, &G.flo_oi)) || (rc = upload(s, obj_a, &G.obj_a)) || (rc = upload(s, obj_n, &G.obj_n)) ||
        (rc = upload(s, obj_b, &G.obj_b)) || (rc = upload(s, mats, &G.mat)) || (rc = upload(s, tex, &G.tex)))
        return bail(rc);

    // exact culling structure for large scenes (rr_trace.cuh "BVH")
    {
        Bvh bvh;
        if (build_bvh(sph_m, bvh)) {
            std::vector<float4> bsph, bsph_m;
            std::vector<int> boi;
            std::vector<float> bglow;
            for (int k : bvh.order) {
                bsph.push_back(sph[k]); bsph_m.push_back(sph_m[k]); boi.push_back(sph_oi[k]); bglow.push_back(sph_glow[k]);
            }
            if ((rc = upload(s, bvh.a, &G.bvh_a)) || (rc = upload(s, bvh.b, &G.bvh_b)) || (rc = upload(s, bvh.w, &G.bvh_w)) ||
                (rc = upload(s, bsph, &G.bsph)) ||
                (rc = upload(s, bsph_m, &G.bsph_m)) || (rc = upload(s, boi, &G.bsph_oi)) || (rc = upload(s, bglow, &G.bsph_glow)))
                return bail(rc);
            G.n_bvh_nodes = (int)bvh.a.size();
            G.n_bvh_inner = (int)(bvh.w.size() / 4);
            for (int c = 0; c < 3; ++c) { G.scene_lo[c] = bvh.lo[c]; G.scene_hi[c] = bvh.hi[c]; }
            G.r_min = bvh.r_min;
        }
    }

    for (int k = 0; k < rr::RR_HEAD_SPHERES && k < (int)sph.size(); ++k) { s->H.sph[k] = sph[k]; s->H.sph_oi[k] = sph_oi[k]; s->H.sph_m[k] = sph_m[k]; s->H.sph_glow[k] = sph_glow[k]; }
    {   // glowing objects for the march kernel's separate glow pass
        int ng = 0;
        for (uint32_t i = 0; i < desc->n_objects && ng >= 0; ++i) {
            const rr_object &o = desc->objects[i];
            const rr_material &m = desc->materials[o.material];
            if (m.glow_dist == 0.0f) continue;
            if (ng == rr::RR_HEAD_GLOW) { ng = -1; break; }
            s->H.glow_a[ng] = make_float4(o.org[0], o.org[1], o.org[2], o.kind == RR_SPHERE ? o.r : 0.0f);
            s->H.glow_b[ng] = make_float4(o.face_normal[0], o.face_normal[1], o.face_normal[2], 0.0f);
            s->H.glow_k[ng] = m.glow_dist;
            s->H.glow_kind[ng] = o.kind == RR_SPHERE ? 0 : 1;
            s->H.glow_oi[ng] = (int)i;
            ++ng;
        }
        s->H.n_glow_head = ng;
    }
    rr::fill_march_bounds(s->H, std::min((int)sph.size(), rr::RR_HEAD_SPHERES));
    for (int k = 0; k < rr::RR_HEAD_FLOORS && k < (int)flo_o.size(); ++k) {
        s->H.flo_o[k] = flo_o[k]; s->H.flo_n[k] = flo_n[k]; s->H.flo_oi[k] = flo_oi[k];
    }
    rr::fill_head_pairs(s->H, std::min((int)sph.size(), rr::RR_HEAD_SPHERES), std::min((int)flo_o.size(), rr::RR_HEAD_FLOORS));

    cudaError_t e;
    int sm = 0, optin = 0;
    if ((e = cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, device)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device)) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&s->d_cnt), sizeof(rr::Counters))) != cudaSuccess ||
        (e = cudaMalloc(reinterpret_cast<void **>(&s->d_work), RR_SLOTS * 4 * sizeof(unsigned))) != cudaSuccess)
        return bail(fail_cuda(e, "rr_scene_create"));
    s->allocs.push_back(s->d_cnt);
    s->allocs.push_back(s->d_work);
    if ((e = cudaMemset(s->d_work, 0, RR_SLOTS * 4 * sizeof(unsigned))) != cudaSuccess) return bail(fail_cuda(e, "cudaMemset"));
    if ((e = rr::preload_signal_kernels()) != cudaSuccess) return bail(fail_cuda(e, "preload"));
    for (Lane &l : s->lanes) {
        if ((e = cudaStreamCreateWithFlags(&l.stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaStreamCreateWithFlags(&l.copy_stream, cudaStreamNonBlocking)) != cudaSuccess ||
            (e = cudaEventCreate(&l.ev0)) != cudaSuccess || (e = cudaEventCreate(&l.ev1)) != cudaSuccess)
            return bail(fail_cuda(e, "rr_scene_create"));
        for (auto &ev : l.chunk_ev)
            if ((e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming)) != cudaSuccess) return bail(fail_cuda(e, "cudaEventCreate"));
    }
    s->li.sm_count = sm;
    s->li.smem_optin = (size_t)optin;
    *out = s;
    return RR_OK;
}

int rr_frame_rows(const rr_frame_params *params, int32_t *rows_out) {
    if (!rows_out) return fail(RR_ERR_BAD_ARG, "rows_out is null");
    int rc = check_params(params);
    if (rc) return rc;
    *rows_out = rows_of(params);
    return RR_OK;
}

// Device-resident render on a caller stream: only the launch is enqueued, nothing of the handle is held afterwards, so
// any number of such launches may be in flight on different streams (each has its own queue slot, see launch()).
// cuda_stream == NULL: one of the handle's lanes, timed, synchronous.
static int render_device(rr_scene *s, rr::FrameParams &P, void *d_out, size_t row_stride, bool f32, void *cuda_stream,
                         unsigned *d_flag = nullptr, unsigned epoch = 0) {
    CU(cudaSetDevice(s->device));
    if (cuda_stream) return launch(s, P, d_out, row_stride, f32, nullptr, reinterpret_cast<cudaStream_t>(cuda_stream), d_flag, epoch);
    LaneHold h(s);
    int rc;
    CU(cudaEventRecord(h.l->ev0, h.l->stream));
    if ((rc = launch(s, P, d_out, row_stride, f32, nullptr, h.l->stream, d_flag, epoch))) return rc;
    CU(cudaEventRecord(h.l->ev1, h.l->stream));
    return finish_timed(h);
}

int rr_render_rgb8_device(rr_scene *s, const rr_frame_params *params, void *d_out, size_t row_stride, void *cuda_stream) {
    if (!s || !d_out) return fail(RR_ERR_BAD_ARG, "scene/d_out is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    if (row_stride == 0) row_stride = (size_t)P.xres * 3;
    if (row_stride < (size_t)P.xres * 3) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    return render_device(s, P, d_out, row_stride, false, cuda_stream);
}

int rr_render_f32_device(rr_scene *s, const rr_frame_params *params, void *d_out, void *cuda_stream) {
    if (!s || !d_out) return fail(RR_ERR_BAD_ARG, "scene/d_out is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    return render_device(s, P, d_out, 0, true, cuda_stream);
}

}  // extern "C"

namespace {
int enqueue_rgb8(rr_scene *s, Lane *l, const rr::FrameParams &P, uint8_t *out, size_t row_stride) {
    int rc;
    const size_t packed = (size_t)P.xres * 3;
    const int rows = P.local_rows;
    void *alias = nullptr;
    if (long_kernel(P, s) && P.xres % 8 == 0 && mapped_host_alias(out, row_stride, &alias)) {
        CU(cudaEventRecord(l->ev0, l->stream));
        if ((rc = launch(s, P, alias, row_stride, false, nullptr, l->stream))) return rc;
        CU(cudaEventRecord(l->ev1, l->stream));
        return RR_OK;
    }
    if ((rc = ensure_out(l, packed * rows))) return rc;
    const int nchunk = pick_chunks(packed * rows, P, s);
    int plan[32];
    const int nplan = plan_chunks(rows, nchunk, plan);
    CU(cudaEventRecord(l->ev0, l->stream));
    for (int k = 0, r0 = 0; k < nplan; r0 += plan[k], ++k) {
        rr::FrameParams C = P;
        C.row0 = r0;
        C.local_rows = plan[k];
        uint8_t *d = reinterpret_cast<uint8_t *>(l->d_out) + (size_t)r0 * packed;
        if ((rc = launch(s, C, d, packed, false, nullptr, l->stream))) return rc;
        CU(cudaEventRecord(l->chunk_ev[k], l->stream));
        CU(cudaStreamWaitEvent(l->copy_stream, l->chunk_ev[k], 0));
        if (row_stride == packed)  // contiguous rows: plain 1-D copy
            CU(cudaMemcpyAsync(out + (size_t)r0 * packed, d, packed * (size_t)C.local_rows, cudaMemcpyDeviceToHost, l->copy_stream));
        else
            CU(cudaMemcpy2DAsync(out + (size_t)r0 * row_stride, row_stride, d, packed, packed, (size_t)C.local_rows,
                                 cudaMemcpyDeviceToHost, l->copy_stream));
    }
    CU(cudaEventRecord(l->ev1, l->stream));
    return RR_OK;
}
}  // namespace

extern "C" {

// Frame to host memory. Short kernels: the frame is rendered in up to 32 row chunks (geometric plan, plan_chunks) on a
// lane's stream; each chunk's device-to-host copy is queued on the lane's second stream as soon as its kernel finishes,
// so the PCIe transfer of chunk k overlaps the kernel of chunk k+1 (the copy is the longer leg at 4K/8K).
int rr_render_rgb8(rr_scene *s, const rr_frame_params *params, uint8_t *out, size_t row_stride) {
    if (!s || !out) return fail(RR_ERR_BAD_ARG, "scene/out is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    const size_t packed = (size_t)P.xres * 3;
    if (row_stride == 0) row_stride = packed;
    if (row_stride < packed) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    if (P.local_rows == 0 || P.xres == 0) return RR_OK;
    CU(cudaSetDevice(s->device));
    LaneHold h(s);
    if ((rc = enqueue_rgb8(s, h.l, P, out, row_stride))) return rc;
    return finish_timed(h);
}

// render_frames support (render.rs:926-989: many frames of one scene): the same frame-to-host pipeline, but the call
// returns as soon as everything is queued. Up to RR_LANES frames of one handle can be in flight; a further call waits
// for a free lane. `out` should be page-locked (rr_host_alloc) for the copies to be asynchronous.
int rr_render_rgb8_async(rr_scene *s, const rr_frame_params *params, uint8_t *out, size_t row_stride, int32_t *ticket) {
    if (!s || !out || !ticket) return fail(RR_ERR_BAD_ARG, "scene/out/ticket is null");
    *ticket = -1;
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    const size_t packed = (size_t)P.xres * 3;
    if (row_stride == 0) row_stride = packed;
    if (row_stride < packed) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    CU(cudaSetDevice(s->device));
    LaneHold h(s);
    if (P.local_rows > 0 && P.xres > 0) {
        if ((rc = enqueue_rgb8(s, h.l, P, out, row_stride))) return rc;
    } else {
        CU(cudaEventRecord(h.l->ev0, h.l->stream));
        CU(cudaEventRecord(h.l->ev1, h.l->stream));
    }
    {
        std::lock_guard<std::mutex> lk(s->mu);
        h.l->pending = true;
        *ticket = (int32_t)((h.l - s->lanes) | (int)((h.l->gen & 0xffffffu) << 4));
    }
    h.keep = true;
    return RR_OK;
}

int rr_render_wait(rr_scene *s, int32_t ticket, float *kernel_ms) {
    if (!s || ticket < 0) return fail(RR_ERR_BAD_ARG, "scene is null or bad ticket");
    const int li = ticket & 15;
    if (li >= RR_LANES) return fail(RR_ERR_BAD_ARG, "bad ticket");
    Lane *l = &s->lanes[li];
    {
        std::lock_guard<std::mutex> lk(s->mu);
        if (!l->busy || !l->pending || (int)((l->gen & 0xffffffu) << 4 | li) != ticket) return fail(RR_ERR_BAD_ARG, "stale ticket");
    }
    CU(cudaSetDevice(s->device));
    cudaError_t e1 = cudaStreamSynchronize(l->copy_stream), e2 = cudaStreamSynchronize(l->stream);
    float ms = 0.0f;
    cudaError_t e3 = (e1 == cudaSuccess && e2 == cudaSuccess) ? cudaEventElapsedTime(&ms, l->ev0, l->ev1) : cudaSuccess;
    const bool ok = e1 == cudaSuccess && e2 == cudaSuccess && e3 == cudaSuccess;
    release_lane(s, l, ms, ok);
    if (!ok) return fail_cuda(e1 != cudaSuccess ? e1 : (e2 != cudaSuccess ? e2 : e3), "rr_render_wait");
    if (kernel_ms) *kernel_ms = ms;
    return RR_OK;
}

int rr_render_f32(rr_scene *s, const rr_frame_params *params, float *out_rgb) {
    if (!s || !out_rgb) return fail(RR_ERR_BAD_ARG, "scene/out is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    const size_t bytes = (size_t)P.xres * P.local_rows * 3 * sizeof(float);
    if (bytes == 0) return RR_OK;
    CU(cudaSetDevice(s->device));
    LaneHold h(s);
    if ((rc = ensure_out(h.l, bytes))) return rc;
    CU(cudaEventRecord(h.l->ev0, h.l->stream));
    if ((rc = launch(s, P, h.l->d_out, 0, true, nullptr, h.l->stream))) return rc;
    CU(cudaEventRecord(h.l->ev1, h.l->stream));
    CU(cudaMemcpyAsync(out_rgb, h.l->d_out, bytes, cudaMemcpyDeviceToHost, h.l->stream));
    return finish_timed(h);
}

int rr_render_count(rr_scene *s, const rr_frame_params *params, uint8_t *out, size_t row_stride, rr_ray_counts *counts) {
    if (!s || !counts) return fail(RR_ERR_BAD_ARG, "scene/counts is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    const size_t packed = (size_t)P.xres * 3;
    if (row_stride == 0) row_stride = packed;
    if (row_stride < packed) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    std::memset(counts, 0, sizeof(*counts));
    if (P.local_rows == 0 || P.xres == 0) return RR_OK;
    CU(cudaSetDevice(s->device));
    std::lock_guard<std::mutex> ck(s->cnt_mu);  // one counter block per handle
    LaneHold h(s);
    cudaStream_t st = h.l->stream;
    if ((rc = ensure_out(h.l, packed * P.local_rows))) return rc;
    CU(cudaMemsetAsync(s->d_cnt, 0, sizeof(rr::Counters), st));
    if ((rc = launch(s, P, h.l->d_out, packed, false, s->d_cnt, st))) return rc;
    rr::Counters c{};
    CU(cudaMemcpyAsync(&c, s->d_cnt, sizeof(c), cudaMemcpyDeviceToHost, st));
    if (out)
        CU(cudaMemcpy2DAsync(out, row_stride, h.l->d_out, packed, packed, (size_t)P.local_rows, cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    counts->pixels = c.pixels; counts->primary = c.primary; counts->reflect = c.reflect; counts->refract = c.refract;
    counts->shadow = c.shadow; counts->object_tests = c.object_tests; counts->march_steps = c.march_steps;
    counts->bg_evals = c.bg_evals; counts->sphere_tests = c.sphere_tests; counts->sphere_hits = c.sphere_hits;
    return RR_OK;
}

int rr_bands_unpack_device(const rr_frame_params *params, const void *d_packed, size_t shard_stride_bytes, void *d_frame,
                           void *cuda_stream) {
    if (!d_packed || !d_frame) return fail(RR_ERR_BAD_ARG, "null device pointer");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, nullptr);
    if (P.band_span > 1) return fail(RR_ERR_UNSUPPORTED, "rr_bands_unpack_device: equal band spans only");
    cudaError_t e = rr::launch_bands_unpack(P, d_packed, shard_stride_bytes, d_frame, reinterpret_cast<cudaStream_t>(cuda_stream));
    if (e != cudaSuccess) return fail_cuda(e, "bands_unpack");
    if (!cuda_stream) CU(cudaStreamSynchronize(nullptr));
    return RR_OK;
}

// ---- multi-GPU: placed output into one shared frame (device: peer memory via CUDA IPC; host: registered memory) ----
int rr_render_rgb8_placed_device(rr_scene *s, const rr_frame_params *params, void *d_frame, size_t row_stride, void *cuda_stream) {
    if (!s || !d_frame) return fail(RR_ERR_BAD_ARG, "scene/d_frame is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    P.placed = 1;
    if (row_stride == 0) row_stride = (size_t)P.xres * 3;
    if (row_stride < (size_t)P.xres * 3) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    return render_device(s, P, d_frame, row_stride, false, cuda_stream);
}

// Same render, plus the completion signal carried by the kernel (rr_device.cuh Signal): epoch lands in
// d_flags[band_index] of the frame owner's memory once every row of this shard is in the frame.
int rr_render_rgb8_placed_signal_device(rr_scene *s, const rr_frame_params *params, void *d_frame, size_t row_stride,
                                        uint32_t *d_flags, uint32_t epoch, void *cuda_stream) {
    if (!s || !d_frame || !d_flags) return fail(RR_ERR_BAD_ARG, "scene/d_frame/d_flags is null");
    int rc = check_params(params);
    if (rc) return rc;
    rr::FrameParams P = to_dev(params, s);
    P.placed = 1;
    if (row_stride == 0) row_stride = (size_t)P.xres * 3;
    if (row_stride < (size_t)P.xres * 3) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    CU(cudaSetDevice(s->device));
    return launch(s, P, d_frame, row_stride, false, nullptr, reinterpret_cast<cudaStream_t>(cuda_stream), d_flags + P.band_index, epoch);
}

int rr_fence_wait_device(int device, const uint32_t *d_flags, int32_t count, uint32_t epoch, uint32_t timeout_ms, uint32_t *d_status,
                         void *cuda_stream) {
    if (!d_flags || count < 0) return fail(RR_ERR_BAD_ARG, "d_flags is null or count < 0");
    CU(cudaSetDevice(device));
    CU(rr::launch_fence_wait(d_flags, count, epoch, timeout_ms, d_status, reinterpret_cast<cudaStream_t>(cuda_stream)));
    return RR_OK;
}

int rr_fence_signal_device(int device, uint32_t *d_flag, uint32_t epoch, void *cuda_stream) {
    if (!d_flag) return fail(RR_ERR_BAD_ARG, "d_flag is null");
    CU(cudaSetDevice(device));
    CU(rr::launch_signal(rr::Signal{nullptr, nullptr, d_flag, epoch}, reinterpret_cast<cudaStream_t>(cuda_stream)));
    return RR_OK;
}

int rr_device_memset(int device, void *d_ptr, int value, size_t bytes) {
    if (!d_ptr) return fail(RR_ERR_BAD_ARG, "d_ptr is null");
    CU(cudaSetDevice(device));
    CU(cudaMemset(d_ptr, value, bytes));
    CU(cudaDeviceSynchronize());
    return RR_OK;
}

int rr_device_copy(int device, void *d_dst, const void *d_src, size_t bytes, void *cuda_stream) {
    if (!d_dst || !d_src) return fail(RR_ERR_BAD_ARG, "pointer is null");
    CU(cudaSetDevice(device));
    CU(cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDefault, reinterpret_cast<cudaStream_t>(cuda_stream)));
    return RR_OK;
}

int rr_device_read(int device, const void *d_ptr, void *host, size_t bytes) {
    if (!d_ptr || !host) return fail(RR_ERR_BAD_ARG, "d_ptr/host is null");
    CU(cudaSetDevice(device));
    CU(cudaMemcpy(host, d_ptr, bytes, cudaMemcpyDeviceToHost));
    return RR_OK;
}

// This shard's bands rendered packed on the device, then copied straight to their rows of a full
// frame in host memory (each GPU over its own PCIe link when several processes share the frame).
int rr_render_rgb8_placed(rr_scene *s, const rr_frame_params *params, uint8_t *host_frame, size_t row_stride) {
    if (!s || !host_frame) return fail(RR_ERR_BAD_ARG, "scene/host_frame is null");
    int rc = check_params(params);
    if (rc) return rc;
    CU(cudaSetDevice(s->device));
    rr::FrameParams P = to_dev(params, s);
    if (P.band_span > 1) return fail(RR_ERR_UNSUPPORTED, "rr_render_rgb8_placed: unequal band spans are a device-frame feature");
    const size_t packed = (size_t)P.xres * 3;
    if (row_stride == 0) row_stride = packed;
    if (row_stride < packed) return fail(RR_ERR_BAD_ARG, "row_stride smaller than a row");
    const int rows = P.local_rows;
    if (rows == 0 || P.xres == 0) return RR_OK;
    LaneHold h(s);
    Lane *l = h.l;
    void *alias = nullptr;
    if (long_kernel(P, s) && P.xres % 8 == 0 && mapped_host_alias(host_frame, row_stride, &alias)) {
        rr::FrameParams Z = P;
        Z.placed = 1;  // rows go to their image position in the (shared) host frame
        CU(cudaEventRecord(l->ev0, l->stream));
        if ((rc = launch(s, Z, alias, row_stride, false, nullptr, l->stream))) return rc;
        CU(cudaEventRecord(l->ev1, l->stream));
        return finish_timed(h);
    }
    if ((rc = ensure_out(l, packed * rows))) return rc;
    const int B = P.band_count <= 1 ? 4 : P.band_rows, n = P.band_count, k = P.band_index;
    // Same pipeline as rr_render_rgb8: up to 16 chunks of whole bands, each chunk's device-to-host copy
    // queued on the copy stream behind its kernel so PCIe overlaps the next chunk's rendering.
    const int nchunk = pick_chunks(packed * rows, P, s);
    int chunk_rows = (rows + nchunk - 1) / nchunk;
    const int align = (B % 4 == 0) ? B : B * 4;  // whole bands and whole 4-row tiles
    chunk_rows = ((chunk_rows + align - 1) / align) * align;
    CU(cudaEventRecord(l->ev0, l->stream));
    int ci = 0;
    for (int r0 = 0; r0 < rows; r0 += chunk_rows, ++ci) {
        rr::FrameParams Cp = P;
        Cp.row0 = r0;
        Cp.local_rows = rows - r0 < chunk_rows ? rows - r0 : chunk_rows;
        uint8_t *d = reinterpret_cast<uint8_t *>(l->d_out) + (size_t)r0 * packed;
        if ((rc = launch(s, Cp, d, packed, false, nullptr, l->stream))) return rc;
        CU(cudaEventRecord(l->chunk_ev[ci], l->stream));
        CU(cudaStreamWaitEvent(l->copy_stream, l->chunk_ev[ci], 0));
        const int nr = Cp.local_rows;
        if (n <= 1) {
            CU(cudaMemcpy2DAsync(host_frame + (size_t)r0 * row_stride, row_stride, d, packed, packed, (size_t)nr,
                                 cudaMemcpyDeviceToHost, l->copy_stream));
        } else if (row_stride == packed) {
            // a band is B contiguous rows; this shard's bands are n*B rows apart in the frame: one strided copy
            const int full = nr / B, tail = nr - full * B;
            const size_t band_bytes = (size_t)B * packed;
            const size_t first_band = (size_t)(r0 / B);
            if (full > 0)
                CU(cudaMemcpy2DAsync(host_frame + (first_band * n + k) * band_bytes, (size_t)n * band_bytes, d, band_bytes, band_bytes,
                                     (size_t)full, cudaMemcpyDeviceToHost, l->copy_stream));
            if (tail > 0)
                CU(cudaMemcpyAsync(host_frame + ((first_band + full) * n + k) * band_bytes, d + (size_t)full * band_bytes,
                                   (size_t)tail * packed, cudaMemcpyDeviceToHost, l->copy_stream));
        } else {
            for (int b0 = 0; b0 < nr; b0 += B) {  // padded rows: one 2D copy per band
                const int br = nr - b0 < B ? nr - b0 : B;
                const size_t iy = ((size_t)((r0 + b0) / B) * n + k) * B;
                CU(cudaMemcpy2DAsync(host_frame + iy * row_stride, row_stride, d + (size_t)b0 * packed, packed, packed, (size_t)br,
                                     cudaMemcpyDeviceToHost, l->copy_stream));
            }
        }
    }
    CU(cudaEventRecord(l->ev1, l->stream));
    return finish_timed(h);
}

int rr_device_alloc(int device, size_t bytes, void **d_ptr) {
    if (!d_ptr) return fail(RR_ERR_BAD_ARG, "d_ptr is null");
    *d_ptr = nullptr;
    CU(cudaSetDevice(device));
    CU(cudaMalloc(d_ptr, bytes ? bytes : 1));
    return RR_OK;
}
int rr_device_free(int device, void *d_ptr) {
    if (!d_ptr) return RR_OK;
    CU(cudaSetDevice(device));
    CU(cudaFree(d_ptr));
    return RR_OK;
}
int rr_ipc_export(void *d_ptr, uint8_t handle[64]) {
    if (!d_ptr || !handle) return fail(RR_ERR_BAD_ARG, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
    cudaIpcMemHandle_t h;
    CU(cudaIpcGetMemHandle(&h, d_ptr));
    std::memcpy(handle, &h, 64);
    return RR_OK;
}
int rr_ipc_open(int device, const uint8_t handle[64], void **d_ptr) {
    if (!d_ptr || !handle) return fail(RR_ERR_BAD_ARG, "null argument");
    *d_ptr = nullptr;
    CU(cudaSetDevice(device));
    cudaIpcMemHandle_t h;
    std::memcpy(&h, handle, 64);
    CU(cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return RR_OK;
}
int rr_ipc_close(int device, void *d_ptr) {
    if (!d_ptr) return RR_OK;
    CU(cudaSetDevice(device));
    CU(cudaIpcCloseMemHandle(d_ptr));
    return RR_OK;
}
int rr_host_register(void *ptr, size_t bytes) {
    if (!ptr) return fail(RR_ERR_BAD_ARG, "ptr is null");
    CU(cudaHostRegister(ptr, bytes, cudaHostRegisterPortable));
    return RR_OK;
}
int rr_host_unregister(void *ptr) {
    if (!ptr) return RR_OK;
    CU(cudaHostUnregister(ptr));
    return RR_OK;
}

int rr_scene_set_culling(rr_scene *s, int enabled) {
    if (!s) return fail(RR_ERR_BAD_ARG, "scene is null");
    std::lock_guard<std::mutex> lk(s->mu);
    s->culling = enabled != 0;
    return RR_OK;
}

int rr_host_alloc(size_t bytes, void **out) {
    if (!out) return fail(RR_ERR_BAD_ARG, "out is null");
    *out = nullptr;
    CU(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
    return RR_OK;
}

int rr_host_free(void *ptr) {
    if (!ptr) return RR_OK;
    CU(cudaFreeHost(ptr));
    return RR_OK;
}

int rr_last_kernel_ms(rr_scene *s, float *ms) {
    if (!s || !ms) return fail(RR_ERR_BAD_ARG, "scene/ms is null");
    std::lock_guard<std::mutex> lk(s->mu);
    if (!s->timed) return fail(RR_ERR_BAD_ARG, "no timed render on this handle yet");
    *ms = s->last_ms;
    return RR_OK;
}

int rr_fp32_peak_tflops(int device, float *unfused_tflops, float *ffma_tflops) {
    if (!unfused_tflops || !ffma_tflops) return fail(RR_ERR_BAD_ARG, "null output");
    cudaError_t e = rr::fp32_peak(device, unfused_tflops, ffma_tflops);
    if (e != cudaSuccess) return fail_cuda(e, "fp32_peak");
    return RR_OK;
}

int rr_selftest_normalize(int device, uint64_t n, uint64_t seed, uint64_t *mismatches) {
    if (!mismatches) return fail(RR_ERR_BAD_ARG, "null output");
    unsigned long long bad = 0;
    cudaError_t e = rr::normalize_selftest(device, (unsigned long long)n, (unsigned long long)seed, &bad);
    if (e != cudaSuccess) return fail_cuda(e, "normalize_selftest");
    *mismatches = (uint64_t)bad;
    return RR_OK;
}

}  // extern "C"
