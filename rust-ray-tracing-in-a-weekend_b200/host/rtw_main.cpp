// rtw_main.cpp — replacement for the reference's main() (src/main.rs:307-599) on top of the C++ mirror: pick a scene,
// build World + Camera with the reference's constructors, flatten, render on the GPU(s), write_color, print the P3 image
// byte-compatibly with src/main.rs:472 and :591-596.
//   rtw_main --scene 0 --width 120 --height 80 --spp 16 [--gpus N] [--seed 1] [--earth earth.rgb] [--passes K] [--png out.png] [--dry-run] > out.ppm
//   --passes K: render in K passes and report "\rProgress: done/total samples" on stderr after each (the reference's progress
//   thread, src/main.rs:557-582, busy-polls per-thread pixel counters instead)
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>

#include "scenes.hpp"

extern "C" int rtw_debug_flatten(rtw_scene*, int32_t*, double*);

int main(int argc, char** argv) {
    int scene_id = 0, width = 0, height = 0, spp = 0, gpus = 1, dry = 0, passes = 0, max_depth = 50;     // max_depth: src/main.rs:310
    unsigned long long seed = 1;
    std::string earth, png;
    for (int i = 1; i < argc; ++i) {
        auto is = [&](const char* k) { return !strcmp(argv[i], k) && i + 1 < argc; };
        if (is("--scene")) scene_id = atoi(argv[++i]);
        else if (is("--width")) width = atoi(argv[++i]);
        else if (is("--height")) height = atoi(argv[++i]);
        else if (is("--spp")) spp = atoi(argv[++i]);
        else if (is("--gpus")) gpus = atoi(argv[++i]);
        else if (is("--seed")) seed = strtoull(argv[++i], nullptr, 10);
        else if (is("--depth")) max_depth = atoi(argv[++i]);
        else if (is("--earth")) earth = argv[++i];
        else if (is("--passes")) passes = atoi(argv[++i]);
        else if (is("--png")) png = argv[++i];
        else if (!strcmp(argv[i], "--dry-run")) dry = 1;
        else { fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    try {
        using namespace rtw_host;
        std::vector<uint8_t> texels;
        if (!earth.empty()) {
            std::ifstream f(earth, std::ios::binary);
            texels.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
        }
        SceneDesc sd = make_scene(scene_id, seed, texels.empty() ? nullptr : &texels);
        if (width > 0) sd.image_width = width;
        if (height > 0) sd.image_height = height;
        if (spp > 0) sd.samples_per_pixel = spp;
        const double aspect = (double)sd.image_width / (double)sd.image_height;
        Camera cam = Camera::create(sd.look_from, sd.look_at, Vector3(0.0, 1.0, 0.0), sd.vfov, aspect, 0.1, 10.0, 0.0, 1.0);   // src/main.rs:469
        Backend be;
        be.flatten(sd.world);
        if (dry) {                        // host-only: report what the flatten produced (no GPU needed)
            int32_t c[8]; double sah = 0;
            if (rtw_debug_flatten(be.scene(), c, &sah) != RTW_OK) { fprintf(stderr, "flatten failed: %s\n", rtw_last_error()); return 1; }
            printf("prims %d bvh_prims %d nodes %d xforms %d media %d mats %d texs %d depth %d\n", c[0], c[1], c[2], c[3], c[4], c[5], c[6], c[7]);
            return 0;
        }
        fprintf(stderr, "Rendering %dx%d (%d pixels) image with %d samples per pixel and a max depth of %d, using %d GPU(s)\n",
                sd.image_width, sd.image_height, sd.image_width * sd.image_height, sd.samples_per_pixel, max_depth, gpus);
        auto t0 = std::chrono::steady_clock::now();
        be.commit(gpus, 0);
        rtw_stats st{};
        std::vector<float> sums;
        if (passes > 1) {
            const int per_pass = (sd.samples_per_pixel + passes - 1) / passes;
            be.render_progressive(cam, sd.image_width, sd.image_height, sd.samples_per_pixel, max_depth, sd.background, seed, 0, per_pass, sums,
                                  [](int done, int total, const float*) { fprintf(stderr, "\rProgress: %d/%d samples", done, total); return false; }, &st);
            fprintf(stderr, "\n");
        } else {
            sums = be.render(cam, sd.image_width, sd.image_height, sd.samples_per_pixel, max_depth, sd.background, seed, &st);
        }
        const int n = sd.image_width * sd.image_height;
        std::vector<uint8_t> rgb((size_t)n * 3);
        if (rtw_write_color(sums.data(), n, sd.samples_per_pixel, rgb.data()) != RTW_OK) throw std::runtime_error(rtw_last_error());
        if (!png.empty() && rtw_write_png(png.c_str(), rgb.data(), sd.image_width, sd.image_height) != RTW_OK) throw std::runtime_error("cannot write " + png);
        printf("P3\n%d %d\n255\n\n", sd.image_width, sd.image_height);              // println!("P3\n{} {}\n255\n") src/main.rs:472
        for (int i = 0; i < n; ++i) printf("%d %d %d\n", rgb[3 * i], rgb[3 * i + 1], rgb[3 * i + 2]);   // rows already top first (:591)
        double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        fprintf(stderr, "Rendering finished in %.3f seconds (kernel %.3f ms, %.1f Mpaths/s)\n", secs, st.ms_render,
                (double)st.paths / st.ms_render / 1e3);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
