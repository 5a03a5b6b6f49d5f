// NOT COMPILED IN THIS REPO: rustc/cargo are absent from the build image.
// Reviewed source of the Rust side of the drop-in; the same call sequence is exercised by host/rtw.hpp (C++) and api.py.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_double, c_int, c_void};

#[repr(C)] pub struct rtw_scene { _private: [u8; 0] }

#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct rtw_camera {                       // the 10 pub fields of Camera, src/camera.rs:4-15
    pub origin: [c_double; 3], pub lower_left_corner: [c_double; 3],
    pub horizontal: [c_double; 3], pub vertical: [c_double; 3],
    pub u: [c_double; 3], pub v: [c_double; 3], pub w: [c_double; 3],
    pub lens_radius: c_double, pub time0: c_double, pub time1: c_double,
}
#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct rtw_render_params {
    pub width: i32, pub height: i32, pub spp: i32, pub max_depth: i32,
    pub background: [c_double; 3], pub t_min: c_double, pub seed: u64,
    pub n_gpus: i32, pub samples_per_unit: i32, pub flags: i32, pub reserved: i32,
}
#[repr(C)] #[derive(Default, Clone, Copy)]
pub struct rtw_stats {
    pub ms_render: c_double, pub ms_total: c_double, pub ms_commit: c_double,
    pub paths: u64, pub rays: u64, pub units_per_device: [u64; 8],
    pub n_devices: i32, pub kernel_launches: i32, pub h2d_bytes: u64, pub d2h_bytes: u64,
    pub n_prims: i32, pub n_nodes: i32, pub n_materials: i32, pub n_media: i32,
}

extern "C" {
    pub fn rtw_last_error() -> *const c_char;
    pub fn rtw_device_count() -> c_int;
    pub fn rtw_scene_new() -> *mut rtw_scene;
    pub fn rtw_scene_free(s: *mut rtw_scene);
    pub fn rtw_tex_solid(s: *mut rtw_scene, rgb: *const c_double) -> c_int;
    pub fn rtw_tex_checker(s: *mut rtw_scene, even: *const c_double, odd: *const c_double) -> c_int;
    pub fn rtw_tex_noise(s: *mut rtw_scene, ranvec: *const c_double, px: *const i32, py: *const i32, pz: *const i32, scale: c_double) -> c_int;
    pub fn rtw_tex_image(s: *mut rtw_scene, w: i32, h: i32, bytes_per_scanline: i32, rgb8: *const u8) -> c_int;
    pub fn rtw_mat_lambertian(s: *mut rtw_scene, tex: c_int) -> c_int;
    pub fn rtw_mat_metal(s: *mut rtw_scene, albedo: *const c_double, fuzz: c_double) -> c_int;
    pub fn rtw_mat_dielectric(s: *mut rtw_scene, ir: c_double) -> c_int;
    pub fn rtw_mat_diffuse_light(s: *mut rtw_scene, tex: c_int) -> c_int;
    pub fn rtw_mat_isotropic(s: *mut rtw_scene, tex: c_int) -> c_int;
    pub fn rtw_sphere(s: *mut rtw_scene, mat: c_int, c: *const c_double, r: c_double) -> c_int;
    pub fn rtw_moving_sphere(s: *mut rtw_scene, mat: c_int, c0: *const c_double, c1: *const c_double, t0: c_double, t1: c_double, r: c_double) -> c_int;
    pub fn rtw_xy_rect(s: *mut rtw_scene, mat: c_int, x0: c_double, x1: c_double, y0: c_double, y1: c_double, k: c_double) -> c_int;
    pub fn rtw_xz_rect(s: *mut rtw_scene, mat: c_int, x0: c_double, x1: c_double, z0: c_double, z1: c_double, k: c_double) -> c_int;
    pub fn rtw_yz_rect(s: *mut rtw_scene, mat: c_int, y0: c_double, y1: c_double, z0: c_double, z1: c_double, k: c_double) -> c_int;
    pub fn rtw_box(s: *mut rtw_scene, min: *const c_double, max: *const c_double, mat: c_int) -> c_int;
    pub fn rtw_translate(s: *mut rtw_scene, child: c_int, offset: *const c_double) -> c_int;
    pub fn rtw_rotate_y(s: *mut rtw_scene, angle_deg: c_double, child: c_int) -> c_int;
    pub fn rtw_rotate_y_sincos(s: *mut rtw_scene, sin_theta: c_double, cos_theta: c_double, child: c_int) -> c_int;
    pub fn rtw_constant_medium(s: *mut rtw_scene, child: c_int, density: c_double, phase_mat: c_int) -> c_int;
    pub fn rtw_bvh_node(s: *mut rtw_scene, children: *const i32, n: i32, t0: c_double, t1: c_double) -> c_int;
    pub fn rtw_world_push(s: *mut rtw_scene, hittable: c_int) -> c_int;
    pub fn rtw_scene_commit(s: *mut rtw_scene, n_gpus: i32, first_device: i32) -> c_int;
    pub fn rtw_render(s: *mut rtw_scene, cam: *const rtw_camera, p: *const rtw_render_params, out_rgb_sum: *mut f32, stats: *mut rtw_stats) -> c_int;
    pub fn rtw_write_color(rgb_sum: *const f32, n_pixels: i32, spp: i32, out_rgb8: *mut u8) -> c_int;
    // progressive passes / resume + progress callback (replaces the progress thread, src/main.rs:557-582)
    pub fn rtw_render_progressive(s: *mut rtw_scene, cam: *const rtw_camera, p: *const rtw_render_params, first_sample: i32, samples_per_pass: i32,
                                  inout_rgb_sum: *mut f32, progress: Option<extern "C" fn(i32, i32, *const f32, *mut c_void) -> c_int>,
                                  user: *mut c_void, stats: *mut rtw_stats) -> c_int;
    // output files: the P3 text of src/main.rs:472 + :591-596, or a PNG of the same pixels
    pub fn rtw_write_ppm(path: *const c_char, rgb8: *const u8, width: i32, height: i32) -> c_int;
    pub fn rtw_write_png(path: *const c_char, rgb8: *const u8, width: i32, height: i32) -> c_int;
}
