// NOT COMPILED IN THIS REPO: rustc/cargo are absent from the build image.
// Reviewed source of the Rust side of the drop-in; the same call sequence is exercised by host/rtw.hpp (C++) and api.py.
let sums = gpu::render(&scene.world, &camera, image_width, image_height, scene.samples_per_pixel, max_depth, &scene.background, 0);
println!("P3\n{} {}\n255\n", image_width, image_height);              // src/main.rs:472
for px in sums.chunks(3) {                                            // already top row first (src/main.rs:591)
    Color::new(px[0] as f64, px[1] as f64, px[2] as f64).write_color(scene.samples_per_pixel as i32);
}
