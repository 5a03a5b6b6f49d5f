// NOT COMPILED IN THIS REPO: rustc/cargo are absent from the build image.
// Reviewed source of the Rust side of the drop-in; the same call sequence is exercised by host/rtw.hpp (C++) and api.py.
use crate::{ffi::*, hittable::Hittable, material::Material, texture::Texture, camera::Camera, World};

fn v(p: &crate::math::Vector3) -> [f64; 3] { [p.x, p.y, p.z] }
fn ok(rc: i32) -> i32 { if rc < 0 { panic!("rtw: {}", unsafe { std::ffi::CStr::from_ptr(rtw_last_error()).to_string_lossy() }) } rc }

unsafe fn tex(s: *mut rtw_scene, t: &Texture) -> i32 {
    match t {                                                        // src/texture.rs:4-9
        Texture::SolidColor(c) => ok(rtw_tex_solid(s, v(c).as_ptr())),
        Texture::Checker(even, odd) => ok(rtw_tex_checker(s, v(even).as_ptr(), v(odd).as_ptr())),
        Texture::Noise(p, scale) => {                                // src/perlin.rs:5-10
            let rv: Vec<f64> = p.ranvec.iter().flat_map(|q| [q.x, q.y, q.z]).collect();
            ok(rtw_tex_noise(s, rv.as_ptr(), p.perm_x.as_ptr(), p.perm_y.as_ptr(), p.perm_z.as_ptr(), *scale))
        }
        Texture::Image(w, h, bps, data) => ok(rtw_tex_image(s, *w as i32, *h as i32, *bps as i32, data.as_ptr())),
    }
}

unsafe fn hit(s: *mut rtw_scene, h: &Hittable) -> i32 {
    match h {                                                        // src/hittable.rs:29-41
        Hittable::Sphere { mat_handle, center, radius } => ok(rtw_sphere(s, mat_handle.0 as i32, v(center).as_ptr(), *radius)),
        Hittable::MovingSphere { mat_handle, center_0, center_1, time_0, time_1, radius } =>
            ok(rtw_moving_sphere(s, mat_handle.0 as i32, v(center_0).as_ptr(), v(center_1).as_ptr(), *time_0, *time_1, *radius)),
        Hittable::XYRect { mat_handle, x0, x1, y0, y1, k } => ok(rtw_xy_rect(s, mat_handle.0 as i32, *x0, *x1, *y0, *y1, *k)),
        Hittable::XZRect { mat_handle, x0, x1, z0, z1, k } => ok(rtw_xz_rect(s, mat_handle.0 as i32, *x0, *x1, *z0, *z1, *k)),
        Hittable::YZRect { mat_handle, y0, y1, z0, z1, k } => ok(rtw_yz_rect(s, mat_handle.0 as i32, *y0, *y1, *z0, *z1, *k)),
        Hittable::Box { mat_handle, min, max, .. } => ok(rtw_box(s, v(min).as_ptr(), v(max).as_ptr(), mat_handle.0 as i32)),
        Hittable::Translate { offset, ptr } => { let c = hit(s, ptr); ok(rtw_translate(s, c, v(offset).as_ptr())) }
        Hittable::RotateY { sin_theta, cos_theta, ptr, .. } => {     // new_rotate_y stores sin/cos (:147-152): passed on unchanged
            let c = hit(s, ptr); ok(rtw_rotate_y_sincos(s, *sin_theta, *cos_theta, c))
        }
        Hittable::ConstantMedium { phase_function, boundary, neg_inv_density } => {
            let c = hit(s, boundary); ok(rtw_constant_medium(s, c, -1.0 / *neg_inv_density, phase_function.0 as i32))
        }
        Hittable::BvhNode { .. } => {                                // membership only: the backend builds its own BVH
            let mut leaves = Vec::new(); collect(s, h, &mut leaves);
            ok(rtw_bvh_node(s, leaves.as_ptr(), leaves.len() as i32, 0.0, 1.0))
        }
    }
}
// new_bvh_node puts a single-object span into BOTH children as two separate Box::new(clone) (:96-98), so pointer
// identity never detects the copy.  The walk registers every leaf it meets; rtw_bvh_node drops members that are
// field-by-field equal to an earlier member (what #[derive(Clone)] produces), so the backend sees each object once
// (tests/test_host_abi.py::test_bvh_node_members_cloned_by_the_reference_builder_are_emitted_once walks such a tree).
unsafe fn collect(s: *mut rtw_scene, h: &Hittable, out: &mut Vec<i32>) {
    if let Hittable::BvhNode { left, right, .. } = h {
        collect(s, left, out);
        collect(s, right, out);
    } else { out.push(hit(s, h)); }
}

pub fn render(world: &World, cam: &Camera, w: usize, h: usize, spp: usize, depth: i32, bg: &crate::math::Color, n_gpus: i32) -> Vec<f32> {
    unsafe {
        let s = rtw_scene_new();
        for m in &world.materials {                                  // handles stay 1-based (src/main.rs:46-49)
            match m {
                Material::Lambertian { albedo } => { let t = tex(s, albedo); ok(rtw_mat_lambertian(s, t)); }
                Material::Metal { albedo, fuzz } => { ok(rtw_mat_metal(s, v(albedo).as_ptr(), *fuzz)); }
                Material::Dielectric { ir } => { ok(rtw_mat_dielectric(s, *ir)); }
                Material::DiffuseLight { emit } => { let t = tex(s, emit); ok(rtw_mat_diffuse_light(s, t)); }
                Material::Isotropic { albedo } => { let t = tex(s, albedo); ok(rtw_mat_isotropic(s, t)); }
            }
        }
        for hh in &world.hittables { let id = hit(s, hh); ok(rtw_world_push(s, id)); }
        ok(rtw_scene_commit(s, n_gpus, 0));
        let c = rtw_camera { origin: v(&cam.origin), lower_left_corner: v(&cam.lower_left_corner), horizontal: v(&cam.horizontal),
                             vertical: v(&cam.vertical), u: v(&cam.u), v: v(&cam.v), w: v(&cam.w),
                             lens_radius: cam.lense_radius, time0: cam.time_0, time1: cam.time_1 };
        let p = rtw_render_params { width: w as i32, height: h as i32, spp: spp as i32, max_depth: depth, background: v(bg),
                                    t_min: 0.001, seed: 1, n_gpus: 0, ..Default::default() };
        let mut out = vec![0f32; w * h * 3];
        let mut st = rtw_stats::default();
        ok(rtw_render(s, &c, &p, out.as_mut_ptr(), &mut st));
        rtw_scene_free(s);
        out                                                          // per-pixel SUMS, row 0 = top
    }
}
