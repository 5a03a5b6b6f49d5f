/*
 * rtw.h — C ABI of the B200-native path-tracing backend (librtw.so).
 *
 * This is the drop-in boundary for the render hot path of
 * themeshpotato/rust-ray-tracing-in-a-weekend.  The reference has no FFI of its
 * own (ray_color and the worker loop are private fns of the binary,
 * src/main.rs:19, :507), so the boundary replaces exactly src/main.rs:474-589
 * (thread fan-out + framebuffer reduce): the host keeps building a World and a
 * Camera with the reference's constructors, a flatten step walks
 * World.hittables / World.materials and calls the functions below, and
 * write_color (src/main.rs:591-596) consumes the per-pixel radiance sums that
 * rtw_render returns.  Each entry point cites the reference item it mirrors.
 *
 * Conventions
 *  - plain pointers and sizes only; every input is COPIED before the call
 *    returns; outputs are caller-allocated.
 *  - constructors return an id (>= 0; materials: the reference's 1-based
 *    MaterialHandle) or a negative rtw_status.  Other calls return rtw_status.
 *  - no call unwinds or aborts across the boundary (the reference panics:
 *    src/texture.rs:14-18, src/main.rs:461-463, :26); rtw_last_error() returns
 *    a thread-local message.
 *  - all scene inputs are f64 like the reference; the device path computes in
 *    f32 with f64 islands (see DESIGN.md "Precision").
 *  - every constructor rejects non-finite numbers and the degenerate values the
 *    reference would turn into NaN pixels (radius 0, density 0, time0 == time1,
 *    ir 0, look_from == look_at) with RTW_ERR_INVALID_ARG.
 *  - there is NO CPU fallback: without a CUDA device rtw_scene_commit fails
 *    with RTW_ERR_NO_DEVICE.
 */
#ifndef RTW_H
#define RTW_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum rtw_status {
    RTW_OK = 0,
    RTW_ERR_INVALID_ARG = -1,
    RTW_ERR_UNSUPPORTED_NESTING = -2, /* composition outside SURVEY §8a-H */
    RTW_ERR_CUDA = -3,
    RTW_ERR_OOM = -4,
    RTW_ERR_NO_DEVICE = -5,
    RTW_ERR_NOT_COMMITTED = -6
} rtw_status;

typedef struct rtw_scene rtw_scene; /* opaque; = reference `World` (src/main.rs:40-43) + device replicas */

/* The 10 public fields of the reference Camera (src/camera.rs:4-15). */
typedef struct rtw_camera {
    double origin[3];
    double lower_left_corner[3];
    double horizontal[3];
    double vertical[3];
    double u[3], v[3], w[3];
    double lens_radius;
    double time0, time1;
} rtw_camera;

/* Everything the reference hard-codes in main() (src/main.rs:309-312, :25, :316-459). */
typedef struct rtw_render_params {
    int32_t width, height;      /* explicit; aspect = width/height (not src/main.rs:467) */
    int32_t spp;                /* samples per pixel (all of them; no spp/thread_count truncation); <= 1048576 */
    int32_t max_depth;          /* 50 in the reference (src/main.rs:310); <= 63 */
    double background[3];       /* constant miss colour (src/main.rs:37) */
    double t_min;               /* 0.001 (src/main.rs:25), ray-parameter units */
    uint64_t seed;              /* Philox key; counter = (draw block, bounce, pixel, sample) */
    int32_t n_gpus;             /* 0 = every device the scene was committed to */
    int32_t samples_per_unit;   /* 0 = auto; work unit = 32-pixel tile x this many samples */
    int32_t flags;              /* RTW_FLAG_* */
    int32_t reserved;
} rtw_render_params;

#define RTW_FLAG_DEVICE_OUT 1   /* out_rgb_sum is a device pointer on the first device (no D2H) */
#define RTW_FLAG_KERNEL_MEGA 2  /* force the one-path-per-lane megakernel */
#define RTW_FLAG_NO_TILE_CULL 8 /* diagnostic: skip the per-tile candidate lists, primary rays traverse the BVH (same image) */
#define RTW_FLAG_KERNEL_POOL 4  /* force the warp-pool (shared-memory wavefront) kernel; default: chosen by measurement */
#define RTW_FLAG_KERNEL_WAVEFRONT 16 /* force the global-memory wavefront pipeline (path pool in HBM, dynamic ray fetch) */

typedef struct rtw_stats {
    double ms_render;           /* CUDA-event time of the render kernels, max over devices */
    double ms_total;            /* host wall time of the rtw_render call */
    double ms_commit;           /* host wall time of the last rtw_scene_commit */
    uint64_t paths;             /* width*height*spp */
    uint64_t rays;              /* ray segments traced (all devices) */
    uint64_t units_per_device[8];
    int32_t n_devices;
    int32_t kernel_launches;    /* kernels launched by this call */
    uint64_t h2d_bytes, d2h_bytes;
    int32_t n_prims, n_nodes, n_materials, n_media;
} rtw_stats;

const char* rtw_last_error(void);
int rtw_device_count(void);             /* CUDA devices visible; 0 = none */
const char* rtw_version(void);

/* ---- scene = reference World (src/main.rs:40-50) -------------------------------------------- */
rtw_scene* rtw_scene_new(void);
void rtw_scene_free(rtw_scene*);

/* Texture (src/texture.rs:4-9) -> texture id */
int rtw_tex_solid(rtw_scene*, const double rgb[3]);                                   /* SolidColor */
int rtw_tex_checker(rtw_scene*, const double even[3], const double odd[3]);           /* Checker(even, odd) */
int rtw_tex_noise(rtw_scene*, const double* ranvec /*256*3*/, const int32_t* perm_x,  /* Noise(Perlin, scale); */
                  const int32_t* perm_y, const int32_t* perm_z, double scale);        /* tables: src/perlin.rs:5-10 */
int rtw_tex_image(rtw_scene*, int32_t width, int32_t height, int32_t bytes_per_scanline,
                  const uint8_t* rgb8);                                               /* Image(w,h,bps,data) */

/* Material (src/material.rs:6-12) -> 1-based handle like World::register_material (src/main.rs:46-49) */
int rtw_mat_lambertian(rtw_scene*, int tex);
int rtw_mat_metal(rtw_scene*, const double albedo[3], double fuzz);
int rtw_mat_dielectric(rtw_scene*, double ir);
int rtw_mat_diffuse_light(rtw_scene*, int tex);
int rtw_mat_isotropic(rtw_scene*, int tex);

/* Hittable (src/hittable.rs:29-41) -> hittable id.  Ids may be reused (= Rust .clone()). */
int rtw_sphere(rtw_scene*, int mat, const double center[3], double radius);
/* n static spheres pushed straight into the world (= n x { rtw_sphere; rtw_world_push }); for the 1M-16M sweep */
int rtw_sphere_batch(rtw_scene*, int32_t n, const int32_t* mats, const double* centers /*3n*/, const double* radii);
int rtw_moving_sphere(rtw_scene*, int mat, const double center0[3], const double center1[3],
                      double time0, double time1, double radius);
int rtw_xy_rect(rtw_scene*, int mat, double x0, double x1, double y0, double y1, double k);
int rtw_xz_rect(rtw_scene*, int mat, double x0, double x1, double z0, double z1, double k);
int rtw_yz_rect(rtw_scene*, int mat, double y0, double y1, double z0, double z1, double k);
int rtw_box(rtw_scene*, const double min[3], const double max[3], int mat);           /* new_box :132-145 */
int rtw_translate(rtw_scene*, int child, const double offset[3]);                     /* Translate :38 */
int rtw_rotate_y(rtw_scene*, double angle_deg, int child);                            /* new_rotate_y :147-199 */
/* RotateY as the reference STORES it (sin_theta, cos_theta; src/hittable.rs:39): for hosts that walk an existing
 * Hittable tree — no round trip of the angle through atan2 */
int rtw_rotate_y_sincos(rtw_scene*, double sin_theta, double cos_theta, int child);
int rtw_constant_medium(rtw_scene*, int child, double density, int phase_mat);        /* :201-207 */
int rtw_bvh_node(rtw_scene*, const int32_t* children, int32_t n, double time0, double time1); /* :77-130 */
int rtw_world_push(rtw_scene*, int hittable);                                         /* world.hittables.push */

/* Camera::new (src/camera.rs:18-56) — host-side, pure arithmetic. */
int rtw_camera_new(const double look_from[3], const double look_at[3], const double vup[3],
                   double vfov_deg, double aspect_ratio, double aperture, double focus_dist,
                   double time0, double time1, rtw_camera* out);

/* Flatten (SoA + linearised BVH), upload one replica per device.  n_gpus <= rtw_device_count();
 * first_device lets one-process-per-GPU launchers pin a rank to its LOCAL_RANK.
 * Scenes of >= 256 Ki primitives are built on the device: the spheres travel through 64 MiB of pinned staging
 * buffers that the library allocates once per process, and the builder's scratch comes from a memory pool that is
 * kept across commits (a re-commit of 16 M spheres: 0.05 s) and returned to the driver by rtw_scene_free. */
int rtw_scene_commit(rtw_scene*, int32_t n_gpus, int32_t first_device);

/* The hot path: replaces src/main.rs:497-589.  out_rgb_sum = per-pixel SUM of radiance over spp,
 * row-major H x W x 3 float, row 0 = TOP (= y = H-1 of src/main.rs:591). */
int rtw_render(rtw_scene*, const rtw_camera*, const rtw_render_params*, float* out_rgb_sum, rtw_stats* stats);

/* Progressive accumulation / resume with a progress callback: replaces the busy-poll progress thread
 * (src/main.rs:557-582) and the all-or-nothing output (src/main.rs:591-596).
 * Renders samples [first_sample, params->spp) in passes of `samples_per_pass` (0 = ten passes).  When first_sample > 0,
 * inout_rgb_sum must hold the sums of samples [0, first_sample) on entry (the buffer of an earlier, interrupted call);
 * it holds the sums of everything rendered so far whenever `progress` runs and on return.  The sample index is a
 * coordinate of the Philox counter, so ANY split into passes or resumes yields the image of one rtw_render call
 * (up to f32 summation order).  `progress` (may be NULL) returns non-zero to stop after the current pass: RTW_OK is
 * returned and stats->paths says how far the render got.  RTW_FLAG_DEVICE_OUT is not supported here. */
typedef int (*rtw_progress_fn)(int32_t samples_done, int32_t samples_total, const float* rgb_sum, void* user);
int rtw_render_progressive(rtw_scene*, const rtw_camera*, const rtw_render_params*, int32_t first_sample,
                           int32_t samples_per_pass, float* inout_rgb_sum, rtw_progress_fn progress, void* user,
                           rtw_stats* stats);

/* Page-locked host memory for out_rgb_sum: the D2H of the framebuffer then runs at PCIe speed (optional). */
void* rtw_host_alloc(uint64_t bytes);
void rtw_host_free(void* p);

/* write_color (src/math.rs:119-132) on the device: sums -> 8-bit rgb, same row order. */
int rtw_write_color(const float* rgb_sum, int32_t n_pixels, int32_t spp, uint8_t* out_rgb8);

/* Output files (host only).  rtw_write_ppm: the P3 text the reference prints on stdout — header "P3\n{W} {H}\n255\n\n"
 * (src/main.rs:472), one "r g b" line per pixel (src/math.rs:127-131), rows top to bottom (src/main.rs:591-596).
 * rtw_write_png: the same pixels as an 8-bit RGB PNG (stored deflate blocks). */
int rtw_write_ppm(const char* path, const uint8_t* rgb8, int32_t width, int32_t height);
int rtw_write_png(const char* path, const uint8_t* rgb8, int32_t width, int32_t height);

/* ---- multi-process (one rank per GPU) plumbing: CUDA IPC on the shared framebuffer + tile counter ---- */
#define RTW_IPC_HANDLE_BYTES 160
/* rank 0: allocate the shared framebuffer (W*H*3 float) + tile counter on its device, export handle */
int rtw_shared_create(rtw_scene*, int32_t width, int32_t height, uint8_t handle[RTW_IPC_HANDLE_BYTES]);
/* other ranks: map rank 0's allocation */
int rtw_shared_open(rtw_scene*, int32_t width, int32_t height, const uint8_t handle[RTW_IPC_HANDLE_BYTES]);
/* rank 0, before each step (all ranks must be between steps): zero framebuffer + counter */
int rtw_shared_reset(rtw_scene*);
/* every rank: pull tiles from the shared counter until exhausted, accumulate into the shared framebuffer */
int rtw_render_shared(rtw_scene*, const rtw_camera*, const rtw_render_params*, rtw_stats* stats);
/* The same without reset call and pre-launch barrier: epoch e renders into half (e & 1) of the shared allocation while
 * rank 0 zeroes the other half for epoch e + 1.  Contract: every rank has returned from epoch e - 1 before any rank calls
 * epoch e, and rank 0's device is synchronised in between (one barrier per step).  Epochs count up from 0 after
 * rtw_shared_create / rtw_shared_open. */
int rtw_render_shared_epoch(rtw_scene*, const rtw_camera*, const rtw_render_params*, int32_t epoch, rtw_stats* stats);
int rtw_shared_read_epoch(rtw_scene*, int32_t epoch, float* out_rgb_sum);   /* rank 0: the image of that epoch */
/* rank 0: copy the shared framebuffer to host */
int rtw_shared_read(rtw_scene*, float* out_rgb_sum);
int rtw_shared_close(rtw_scene*);

/* ---- kernel-level parity hooks (each has a same-signature orc_* twin in oracle/) ----------------
 * All arrays are f64 at the boundary; the device converts to f32.  `xi` is an explicit stream of
 * U[0,1) draws, `stride` per item, consumed in the reference's draw order. */
int rtw_test_get_ray(const rtw_camera*, int32_t n, const double* s, const double* t,
                     const double* xi, int32_t stride,
                     double* out_origin, double* out_dir, double* out_time, int32_t* out_ndraw);
/* target < 0: the committed world (BVH path); else hittable id (flattened on the fly). */
int rtw_test_hit(rtw_scene*, int32_t target, int32_t n, const double* origin, const double* dir,
                 const double* time, double t_min, double t_max, const double* xi, int32_t stride,
                 int32_t* out_hit, double* out_t, double* out_p, double* out_normal,
                 int32_t* out_front, double* out_u, double* out_v, int32_t* out_mat, int32_t* out_ndraw);
int rtw_test_aabb(int32_t n, const double* box_min, const double* box_max, const double* origin,
                  const double* dir, double t_min, double t_max, int32_t* out_hit);
int rtw_test_scatter(rtw_scene*, int32_t mat, int32_t n, const double* ray_origin, const double* ray_dir,
                     const double* ray_time, const double* p, const double* normal, const int32_t* front,
                     const double* u, const double* v, const double* xi, int32_t stride,
                     int32_t* out_scattered, double* out_origin, double* out_dir, double* out_time,
                     double* out_attenuation, double* out_emitted, int32_t* out_ndraw);
int rtw_test_texture(rtw_scene*, int32_t tex, int32_t n, const double* u, const double* v,
                     const double* p, double* out_rgb);
/* radiance of individual (x, y(bottom-up), sample) paths with the render's Philox keys */
int rtw_trace_paths(rtw_scene*, const rtw_camera*, const rtw_render_params*, int32_t n,
                    const int32_t* px, const int32_t* py, const int32_t* sample,
                    double* out_rgb, int32_t* out_segments);
/* Philox4x32-10 block (counter[4], key[2]) -> out[4]; pins the RNG to Random123's known answers */
int rtw_test_philox(int32_t n, const uint32_t* counter, const uint32_t* key, uint32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* RTW_H */
