/*
 * rr_ffi.h — C ABI of the B200 per-pixel tracing path (libray_rust_b200.so).
 *
 * This is the drop-in boundary for ray-rust's `render()`:
 *
 *   pub fn render(ren: &RenderEnv, pointproc: &mut impl FnMut(i32, i32, &RenderColor),
 *                 thread_count: i32) -> anyhow::Result<()>         (reference src/render.rs:801-805)
 *
 * The reference has no FFI of its own; `render()` *is* the seam (callers: src/main.rs:338,
 * src/render.rs:979, src/webserver.rs:48).  A Rust `-sys` crate binds exactly the entry points
 * below (see INTEGRATION.md for the stub); the C++ host layer in ray-rust_b200/host/ and the
 * Python ctypes binding in ray-rust_b200/__init__.py are the two bindings built and tested here.
 *
 * Conventions
 *   - POD only. No callbacks cross the ABI. No exceptions / panics cross the ABI.
 *   - Every function returns int: RR_OK (0) or a negative rr_status. The message for the last
 *     failure on the calling thread is returned by rr_last_error().
 *   - All floats are IEEE-754 binary32, all colour channels are linear f32 as in RenderColor
 *     (src/render.rs:23-28); RGB8 output is `(c*255).min(255) as u8` (src/main.rs:148-152).
 *   - A scene handle may be used from several host threads concurrently (the web server calls
 *     render() from several tokio workers, src/webserver.rs:268-280): up to 4 host-facing renders of
 *     one handle run concurrently on the device (each on its own stream pair and device frame), a
 *     fifth waits for a free one; *_device launches on caller streams never wait.
 *   - The device path is the only path: if no CUDA device is usable every entry point fails with
 *     RR_ERR_CUDA. There is no CPU fallback inside this library.
 */
#ifndef RR_FFI_H
#define RR_FFI_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RR_ABI_VERSION 2

/* ------------------------------------------------------------------------------------------ */
/* status codes                                                                                */
/* ------------------------------------------------------------------------------------------ */
typedef enum rr_status {
    RR_OK = 0,
    RR_ERR_BAD_ARG = -1,     /* null pointer, negative size, out-of-range index ...              */
    RR_ERR_CUDA = -2,        /* any CUDA runtime failure (message carries cudaGetErrorString)    */
    RR_ERR_OOM = -3,         /* host or device allocation failed                                 */
    RR_ERR_UNSUPPORTED = -4  /* a scene the device path cannot represent (see rr_scene_create)   */
} rr_status;

/* ------------------------------------------------------------------------------------------ */
/* scene description (host side, read once by rr_scene_create)                                 */
/* ------------------------------------------------------------------------------------------ */

/* RenderObject enum, src/render.rs:585-589 */
typedef enum rr_object_kind { RR_SPHERE = 0, RR_FLOOR = 1 } rr_object_kind;
/* UVMap, src/render.rs:51-57 */
typedef enum rr_uvmap { RR_UV_XY = 0, RR_UV_YZ = 1, RR_UV_ZX = 2, RR_UV_LL = 3 } rr_uvmap;
/* RenderPattern, src/render.rs:44-49 */
typedef enum rr_pattern { RR_SOLID = 0, RR_CHECKERBOARD = 1, RR_REPEATED_GRADATION = 2 } rr_pattern;
/* TextureFilter, src/render.rs:59-63 */
typedef enum rr_texture_filter { RR_NEAREST = 0, RR_BILINEAR = 1 } rr_texture_filter;
/* RenderEnv.bgproc is a host fn pointer (src/render.rs:661) and cannot cross to the device.
 * The only implementation in the reference is `bgcolor` (src/main.rs:231-260). */
typedef enum rr_bg_kind { RR_BG_BGCOLOR = 0, RR_BG_BLACK = 1 } rr_bg_kind;

/* RenderMaterial, src/render.rs:82-97 (name/texture_name stay on the host). */
typedef struct rr_material {
    float diffuse[3];           /* r,g,b */
    float specular[3];
    int32_t pn;                 /* Phong exponent for powi */
    float t;                    /* transparency == refraction blend weight (src/render.rs:1095) */
    float n;                    /* "refraction constant" (src/render.rs:1098) */
    float glow_dist;
    float frac[3];              /* carried for round-trips only; never read by the path */
    int32_t pattern;            /* rr_pattern */
    float pattern_scale;
    float pattern_angle_scale;
    int32_t texture;            /* index into rr_scene_desc.textures, or -1 for None */
    int32_t texture_filter;     /* rr_texture_filter */
} rr_material;

/* RenderSphere (src/render.rs:378-384) / RenderFloor (src/render.rs:487-493). */
typedef struct rr_object {
    int32_t kind;               /* rr_object_kind */
    int32_t material;           /* index into rr_scene_desc.materials */
    int32_t uvmap;              /* rr_uvmap */
    float r;                    /* sphere radius; ignored for floors */
    float org[3];               /* centre */
    float face_normal[3];       /* floors only; used as given, never normalised */
} rr_object;

/* DynamicImage::ImageRgb8 — the only texture format the path honours (src/render.rs:251). */
typedef struct rr_texture {
    uint32_t width, height;
    const uint8_t *rgb8;        /* width*height*3, row-major, no padding */
} rr_texture;

typedef struct rr_scene_desc {
    uint32_t n_objects;
    const rr_object *objects;   /* order is significant: index 0 ends the bounce loop (render.rs:1187) */
    uint32_t n_materials;
    const rr_material *materials;
    uint32_t n_textures;
    const rr_texture *textures;
} rr_scene_desc;

/* ------------------------------------------------------------------------------------------ */
/* per-frame parameters: the RenderEnv fields the path reads per ray (src/render.rs:646-666)   */
/* ------------------------------------------------------------------------------------------ */
typedef struct rr_frame_params {
    int32_t xres, yres;
    float xfov, yfov;
    float cam_position[3];      /* Camera.position */
    float cam_rotation[4];      /* Camera.rotation as x,y,z,w (Quat, src/quat.rs:6-11) */
    float light[3];             /* already normalised by RenderEnv::light (render.rs:720-723) */
    int32_t use_raymarching;    /* 0: raytrace (render.rs:1142), 1: raymarch (render.rs:1299) */
    int32_t glow_enabled;       /* glow_effect.is_some() */
    float glow_effect;
    int32_t max_reflections;    /* honoured by raytrace only (render.rs:1195 vs :1368) */
    int32_t max_refractions;
    int32_t bg_kind;            /* rr_bg_kind */
    /* Row-band sharding (multi-GPU). Rows are grouped in bands of `band_rows`; band b sits in slot
     * (b % band_count) of its period; this call renders only the bands whose slot is in
     * [band_index, band_index + band_span), packed contiguously in band order. band_count<=1 renders
     * the whole frame; band_span <= 1 is one slot per shard (equal shares). Unequal spans give a rank a
     * larger share of the rows: the owner of a multi-GPU frame receives every other rank's rows over
     * its NVLink ports, and rows it renders itself do not cross them. */
    int32_t band_rows;
    int32_t band_index;
    int32_t band_count;
    int32_t band_span;
} rr_frame_params;

/* per-class ray counters written by rr_render_count (definition of "ray": SURVEY.md 8d) */
typedef struct rr_ray_counts {
    uint64_t pixels;
    uint64_t primary;           /* first raycast()/raymarch_single() of a top-level trace */
    uint64_t reflect;           /* further iterations of a bounce loop */
    uint64_t refract;           /* first iteration of a refraction child */
    uint64_t shadow;            /* shadow raycast()/raymarch_single() inside shading() */
    uint64_t object_tests;      /* trace: per-object raycast calls; march: per-object distance calls */
    uint64_t march_steps;       /* march mode: iterations of raymarch_single's loop */
    uint64_t bg_evals;          /* bgproc invocations (reference-equivalent count) */
    uint64_t sphere_tests;      /* the part of object_tests that went to spheres */
    uint64_t sphere_hits;       /* shading() calls whose object is a sphere (normal = 3 div + sqrt) */
} rr_ray_counts;

typedef struct rr_scene rr_scene; /* opaque: device-resident flattened scene + per-handle stream */

/* ------------------------------------------------------------------------------------------ */
/* entry points                                                                                */
/* ------------------------------------------------------------------------------------------ */

/* ABI version of the loaded library (== RR_ABI_VERSION). */
int rr_abi_version(void);

/* "rr_src_hash=<hex> arch=sm_100a fmad=false": the content hash of the sources and compiler flags this binary was built
 * from. Bindings compare it with the sources they sit next to and refuse (or rebuild) a stale binary. */
const char *rr_build_info(void);

/* Thread-local message of the last failure on this thread ("" if none). Never NULL. */
const char *rr_last_error(void);

/* Number of CUDA devices visible to the process. */
int rr_device_count(int *count);

/* Flatten `desc` (RenderEnv.objects + materials + textures, src/render.rs:658-659) to SoA arrays
 * on `device` and return a handle. Replaces the immutable borrow `ren: &RenderEnv` that
 * render() shares with its worker threads (src/render.rs:845-868).
 * RR_ERR_BAD_ARG on dangling material/texture indices or unknown enum values. */
int rr_scene_create(const rr_scene_desc *desc, int device, rr_scene **out);
int rr_scene_destroy(rr_scene *scene);

/* Large scenes (>= 24 spheres) are rendered through an exact culling structure (a BVH whose leaf test
 * is the reference's sphere test; results are bit-identical to the brute-force scan). enabled=0 forces
 * the brute-force scan of render.rs:993-1018 (used by tests to compare the two on the device). */
int rr_scene_set_culling(rr_scene *scene, int enabled);

/* Number of rows a call with `params` produces (all rows, or this shard's bands). */
int rr_frame_rows(const rr_frame_params *params, int32_t *rows_out);

/* render(): RGB8, result in HOST memory. Replaces `render(&ren, &mut putpoint, threads)` +
 * the putpoint quantiser at all three call sites (src/main.rs:148-152,338; src/render.rs:973-979;
 * src/webserver.rs:42-48). out[(x + y*xres)*3 + c]; `row_stride` in bytes (0 => xres*3).
 * Blocking. The timed region of bench.py's `e2e` is exactly one call of this function. */
int rr_render_rgb8(rr_scene *scene, const rr_frame_params *params, uint8_t *out, size_t row_stride);

/* render_frames() support (src/render.rs:926-989 renders many frames of one scene, one render() + image::save_buffer
 * per frame). rr_render_rgb8_async enqueues exactly what rr_render_rgb8 does and returns a ticket; rr_render_wait blocks
 * until that frame is complete in `out` (and reports the device time of its kernels, kernel_ms may be NULL). A handle
 * keeps up to 4 frames in flight (a further call blocks until one is waited for), so the host can overlap the PNG
 * encode of frame k with the kernel and copy of frame k+1, and deal frames to one handle per GPU. `out` should be
 * page-locked (rr_host_alloc); tickets must be waited for exactly once, on any thread. */
int rr_render_rgb8_async(rr_scene *scene, const rr_frame_params *params, uint8_t *out, size_t row_stride,
                         int32_t *ticket);
int rr_render_wait(rr_scene *scene, int32_t ticket, float *kernel_ms);

/* render(): unquantised RenderColor stream (r,g,b f32 per pixel, row-major) in HOST memory, for
 * callers whose pointproc is not the stock quantiser (generic `pointproc(x, y, &RenderColor)`). */
int rr_render_f32(rr_scene *scene, const rr_frame_params *params, float *out_rgb);

/* Same kernels, result left in DEVICE memory on `device` of the handle. `cuda_stream` is a
 * cudaStream_t (NULL = the handle's own stream); the call is asynchronous w.r.t. that stream when
 * a stream is given, and synchronises when NULL. Used for device-resident timing and to hand the
 * rows to a collective (NCCL gather of row bands).
 * Mind that CUDA's DEFAULT stream has the handle 0: passing it means NULL here, i.e. a blocking launch on the handle's
 * own stream that is NOT ordered with work queued on the default stream. Callers that time or pipeline launches create a
 * stream (cudaStreamCreate, torch.cuda.Stream()) and pass that. */
int rr_render_rgb8_device(rr_scene *scene, const rr_frame_params *params, void *d_out,
                          size_t row_stride, void *cuda_stream);
int rr_render_f32_device(rr_scene *scene, const rr_frame_params *params, void *d_out_rgb,
                         void *cuda_stream);

/* Instrumented render: same arithmetic, additionally counts rays per class on the device.
 * Slower; never used for timing. `out` may be NULL (counts only). */
int rr_render_count(rr_scene *scene, const rr_frame_params *params, uint8_t *out, size_t row_stride,
                    rr_ray_counts *counts);

/* Un-interleave row bands gathered from `band_count` shards (each packed as rr_render_*_device
 * wrote them, shard s at d_packed + s*shard_stride_bytes) into a row-major frame, on the device. */
int rr_bands_unpack_device(const rr_frame_params *params, const void *d_packed,
                           size_t shard_stride_bytes, void *d_frame, void *cuda_stream);

/* ---- multi-GPU, one process per GPU: shards write into ONE shared frame ----------
 * The reference gathers finished rows to the caller thread over an mpsc channel (render.rs:846-886).
 * Here every rank renders its interleaved row bands (params->band_*) and the rows land directly at
 * their image position in a frame owned by rank 0:
 *   device frame: rank 0 allocates it with rr_device_alloc and publishes rr_ipc_export's 64-byte handle;
 *     the other ranks map it with rr_ipc_open (NVLink peer memory) and pass the mapped pointer to
 *     rr_render_rgb8_placed_device — the render kernel's own stores cross NVLink, there is no gather
 *     or un-interleave pass. Completion is ordered by the caller (stream sync + a barrier).
 *   host frame: a buffer shared between the processes (e.g. POSIX shared memory), page-locked in each
 *     with rr_host_register; rr_render_rgb8_placed copies this rank's bands over its own PCIe link. */
int rr_render_rgb8_placed_device(rr_scene *scene, const rr_frame_params *params, void *d_frame,
                                 size_t row_stride, void *cuda_stream);
int rr_render_rgb8_placed(rr_scene *scene, const rr_frame_params *params, uint8_t *host_frame,
                          size_t row_stride);
/* Fused completion for the device frame: no collective after the kernel. The render kernel itself publishes
 * `epoch` into d_flags[params->band_index] (a uint32 array of band_count words in the FRAME OWNER's memory; with
 * unequal band spans pass d_flags - band_index + <this shard's word>,
 * allocated with rr_device_alloc + rr_device_memset(0) and mapped by the other ranks like the frame) once all of
 * this shard's rows are in the frame: system-scope fence per block, last block does the release store
 * (both modes; an empty shard queues a one-thread publisher instead). The owner calls rr_fence_wait_device to make
 * its stream wait until all `count` words have reached `epoch` (acquire loads, bounded by timeout_ms; on timeout
 * *d_status is set to 1, d_status may be NULL). Epochs must grow from frame to frame. Both calls only enqueue work
 * (cuda_stream NULL = the default stream); signalled renders of one shard must be ordered on one stream (epochs grow).
 * This replaces the channel receive loop of render.rs:871-886 across GPUs. */
int rr_render_rgb8_placed_signal_device(rr_scene *scene, const rr_frame_params *params, void *d_frame,
                                        size_t row_stride, uint32_t *d_flags, uint32_t epoch,
                                        void *cuda_stream);
int rr_fence_wait_device(int device, const uint32_t *d_flags, int32_t count, uint32_t epoch,
                         uint32_t timeout_ms, uint32_t *d_status, void *cuda_stream);
/* Stand-alone publisher: once everything queued earlier on `cuda_stream` has completed, store `epoch` into *d_flag
 * (device or peer memory) with system-scope release semantics. With rr_fence_wait_device this builds device-side
 * barriers / doorbells between the GPUs of one box (bench.py aligns the ranks' frame starts with it). */
int rr_fence_signal_device(int device, uint32_t *d_flag, uint32_t epoch, void *cuda_stream);
int rr_device_memset(int device, void *d_ptr, int value, size_t bytes);
int rr_device_read(int device, const void *d_ptr, void *host, size_t bytes);
/* Asynchronous device-to-device copy on `cuda_stream` (NULL = default stream); dst/src may be local or peer (rr_ipc_open)
 * memory. bench.py measures the raw NVLink ingress rate of the frame owner with it, as the roofline of the multi-GPU step. */
int rr_device_copy(int device, void *d_dst, const void *d_src, size_t bytes, void *cuda_stream);
int rr_device_alloc(int device, size_t bytes, void **d_ptr);
int rr_device_free(int device, void *d_ptr);
int rr_ipc_export(void *d_ptr, uint8_t handle[64]);
int rr_ipc_open(int device, const uint8_t handle[64], void **d_ptr);
int rr_ipc_close(int device, void *d_ptr);
int rr_host_register(void *ptr, size_t bytes);
int rr_host_unregister(void *ptr);

/* Pinned host memory for frame buffers (full-speed D2H). Plain malloc'd buffers also work. */
int rr_host_alloc(size_t bytes, void **out);
int rr_host_free(void *ptr);

/* Average device time in milliseconds of the kernel(s) of the last rr_render_* call on this
 * handle that was issued on the handle's own stream (CUDA events around the launches). */
int rr_last_kernel_ms(rr_scene *scene, float *ms);

/* FP32 pipe calibration used for the roofline denominator: runs an unfused FMUL+FADD chain and an
 * FFMA chain on every SM and reports achieved TFLOP/s. */
int rr_fp32_peak_tflops(int device, float *unfused_tflops, float *ffma_tflops);

/* Device self-test of Vec3::normalized (vec3.rs:36-39) as the kernels compute it: the three IEEE divisions share one
 * refined reciprocal of the length (csrc/rr_device.cuh). Runs `n` hashed vectors (special values included) through that
 * path and through three plain divisions and reports the number that differ in any bit; 0 is the only passing value. */
int rr_selftest_normalize(int device, uint64_t n, uint64_t seed, uint64_t *mismatches);

#ifdef __cplusplus
}
#endif
#endif /* RR_FFI_H */
