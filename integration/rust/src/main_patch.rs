// NOT COMPILED IN THIS REPO: rustc/cargo are absent from the build image.
// Reviewed source of the Rust side of the drop-in; the same call sequence is exercised by host/rtw.hpp (C++) and api.py.
let sums = gpu::render(&scene.world, &camera, image_width, image_height, scene.samples_per_pixel, max_depth, &scene.background, 0);
println!("P3\n{} {}\n255\n", image_width, image_height);              // src/main.rs:472
for px in sums.chunks(3) {                                            // already top row first (src/main.rs:591)
    Color::new(px[0] as f64, px[1] as f64, px[2] as f64).write_color(scene.samples_per_pixel as i32);
}

// Variant with progress and partial output (replaces the busy-poll progress thread, src/main.rs:557-582): ten passes,
// a line on stderr after each, and the sums so far are a complete image of `done` samples at every callback.
//
// extern "C" fn on_pass(done: i32, total: i32, _sums: *const f32, _user: *mut std::os::raw::c_void) -> i32 {
//     eprint!("\rProgress: {}/{} samples", done, total);
//     0                                                        // non-zero: stop; resume later with first_sample = done
// }
// let mut sums = vec![0f32; image_width * image_height * 3];
// gpu::render_progressive(&scene.world, &camera, image_width, image_height, scene.samples_per_pixel, max_depth,
//                         &scene.background, 0 /* first_sample */, 0 /* ten passes */, &mut sums, Some(on_pass));
//
// gpu::render_progressive is gpu::render with the last call replaced by
//     ok(rtw_render_progressive(s, &c, &p, first_sample, samples_per_pass, sums.as_mut_ptr(), cb, std::ptr::null_mut(), &mut st));
