// rtw.hpp — C++ mirror of the reference's scene-construction surface, on top of the C ABI (include/rtw.h).
//
// The Rust toolchain is absent from this image, so this header plays the role of the Rust shim of INTEGRATION.md:
// the SAME type and constructor names as the reference — Vector3/Point3/Color (src/math.rs:12-20), Camera::new
// (src/camera.rs:18-56), Texture (src/texture.rs:4-9), Perlin (src/perlin.rs:5-30), Material (src/material.rs:6-12),
// MaterialHandle (:97-98), Hittable + new_box / new_rotate_y / new_constant_medium / new_bvh_node
// (src/hittable.rs:29-41, :77-207), World::register_material (src/main.rs:40-50) — and a flatten() that walks
// World.hittables / World.materials exactly like the Rust flatten would, calling rtw_* for every node.
// Host-side only; nothing here computes a pixel.
#ifndef RTW_HOST_HPP
#define RTW_HOST_HPP

#include <cmath>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/rtw.h"

namespace rtw_host {

struct Vector3 {                                                      // src/math.rs:12-20
    double x = 0, y = 0, z = 0;
    Vector3() {}
    Vector3(double a, double b, double c) : x(a), y(b), z(c) {}
    static Vector3 make(double a, double b, double c) { return Vector3(a, b, c); }   // Vector3::new
    double length() const { return std::sqrt(x * x + y * y + z * z); }
};
using Point3 = Vector3;
using Color = Vector3;
inline Vector3 operator+(Vector3 a, Vector3 b) { return Vector3(a.x + b.x, a.y + b.y, a.z + b.z); }
inline Vector3 operator-(Vector3 a, Vector3 b) { return Vector3(a.x - b.x, a.y - b.y, a.z - b.z); }
inline Vector3 operator*(double s, Vector3 a) { return Vector3(a.x * s, a.y * s, a.z * s); }

// Host RNG standing in for rand::thread_rng on the scene-building side (SplitMix64; same stream as scenes.py).
struct HostRng {
    uint64_t s;
    explicit HostRng(uint64_t seed = 1) : s(seed) {}
    double random_double() {                                          // src/math.rs:268-271
        s += 0x9E3779B97F4A7C15ull;
        uint64_t z = s;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
    double random_double_range(double a, double b) { return a + (b - a) * random_double(); }    // :273-276
    int random_int_range(int a, int b) { return (int)random_double_range((double)a, (double)(b + 1)); }   // :278-280
    Vector3 random() { double a = random_double(), b = random_double(), c = random_double(); return Vector3(a, b, c); }   // :35-41
    Vector3 random_range(double lo, double hi) {                      // :43-49
        double a = random_double_range(lo, hi), b = random_double_range(lo, hi), c = random_double_range(lo, hi);
        return Vector3(a, b, c);
    }
};

struct Perlin {                                                       // src/perlin.rs:5-10
    std::vector<Vector3> ranvec;
    std::vector<int32_t> perm_x, perm_y, perm_z;
    static Perlin create(HostRng& g) {                                // Perlin::new :13-30
        Perlin p;
        p.ranvec.resize(256);
        for (int i = 0; i < 256; ++i) { Vector3 v = g.random_range(-1.0, 1.0); p.ranvec[i] = (1.0 / v.length()) * v; }
        p.perm_x = generate_perm(g); p.perm_y = generate_perm(g); p.perm_z = generate_perm(g);
        return p;
    }
    static std::vector<int32_t> generate_perm(HostRng& g) {           // :110-129, including `p[i] = target`
        std::vector<int32_t> p(256);
        for (int i = 0; i < 256; ++i) p[i] = i;
        for (int i = 255; i >= 0; --i) {
            int target = g.random_int_range(0, i);
            if (target > 255) target = 255;
            int tmp = p[i];
            p[i] = target;
            p[target] = tmp;
        }
        return p;
    }
};

struct Texture {                                                      // src/texture.rs:4-9
    enum Kind { SolidColorK, CheckerK, NoiseK, ImageK } kind = SolidColorK;
    Color a, b;
    std::shared_ptr<Perlin> perlin; double scale = 1;
    size_t w = 0, h = 0, bytes_per_scanline = 0; std::shared_ptr<std::vector<uint8_t>> data;
    static Texture SolidColor(Color c) { Texture t; t.kind = SolidColorK; t.a = c; return t; }
    static Texture Checker(Color even, Color odd) { Texture t; t.kind = CheckerK; t.a = even; t.b = odd; return t; }
    static Texture Noise(Perlin p, double scale) { Texture t; t.kind = NoiseK; t.perlin = std::make_shared<Perlin>(std::move(p)); t.scale = scale; return t; }
    static Texture Image(size_t w, size_t h, size_t bps, std::vector<uint8_t> d) {
        Texture t; t.kind = ImageK; t.w = w; t.h = h; t.bytes_per_scanline = bps; t.data = std::make_shared<std::vector<uint8_t>>(std::move(d)); return t;
    }
};

struct MaterialHandle { size_t v = 0; };                              // src/material.rs:97-98 (1-based)

struct Material {                                                     // src/material.rs:6-12
    enum Kind { LambertianK, MetalK, DielectricK, DiffuseLightK, IsotropicK } kind = LambertianK;
    Texture tex; Color albedo; double fuzz = 0, ir = 1;
    static Material Lambertian(Texture albedo) { Material m; m.kind = LambertianK; m.tex = std::move(albedo); return m; }
    static Material Metal(Color albedo, double fuzz) { Material m; m.kind = MetalK; m.albedo = albedo; m.fuzz = fuzz; return m; }
    static Material Dielectric(double ir) { Material m; m.kind = DielectricK; m.ir = ir; return m; }
    static Material DiffuseLight(Texture emit) { Material m; m.kind = DiffuseLightK; m.tex = std::move(emit); return m; }
    static Material Isotropic(Texture albedo) { Material m; m.kind = IsotropicK; m.tex = std::move(albedo); return m; }
};

struct Hittable {                                                     // src/hittable.rs:29-41
    enum Kind { SphereK, MovingSphereK, BvhNodeK, XYRectK, XZRectK, YZRectK, BoxK, TranslateK, RotateYK, ConstantMediumK } kind = SphereK;
    MaterialHandle mat_handle;
    Point3 center, center_1; double radius = 0, time_0 = 0, time_1 = 1;
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0;
    Point3 min, max;
    Vector3 offset; double sin_theta = 0, cos_theta = 1, density = 0;   // RotateY keeps sin / cos, not the angle (:39)
    std::vector<Hittable> children;                                   // Box<Hittable> ptr; BvhNode: { left, right }

    static Hittable Sphere(MaterialHandle m, Point3 c, double r) { Hittable h; h.kind = SphereK; h.mat_handle = m; h.center = c; h.radius = r; return h; }
    static Hittable MovingSphere(MaterialHandle m, Point3 c0, Point3 c1, double t0, double t1, double r) {
        Hittable h; h.kind = MovingSphereK; h.mat_handle = m; h.center = c0; h.center_1 = c1; h.time_0 = t0; h.time_1 = t1; h.radius = r; return h;
    }
    static Hittable XYRect(MaterialHandle m, double x0, double x1, double y0, double y1, double k) { return rect(XYRectK, m, x0, x1, y0, y1, k); }
    static Hittable XZRect(MaterialHandle m, double x0, double x1, double z0, double z1, double k) { return rect(XZRectK, m, x0, x1, z0, z1, k); }
    static Hittable YZRect(MaterialHandle m, double y0, double y1, double z0, double z1, double k) { return rect(YZRectK, m, y0, y1, z0, z1, k); }
    static Hittable new_box(Point3 mn, Point3 mx, MaterialHandle m) { Hittable h; h.kind = BoxK; h.mat_handle = m; h.min = mn; h.max = mx; return h; }   // :132-145
    static Hittable Translate(Vector3 offset, Hittable ptr) { Hittable h; h.kind = TranslateK; h.offset = offset; h.children.push_back(std::move(ptr)); return h; }
    static Hittable new_rotate_y(double angle, Hittable ptr) {                                                                                                      // :147-199
        Hittable h; h.kind = RotateYK;
        const double radians = angle * 3.1415926535897932385 / 180.0;           // degrees_to_radians, src/math.rs:8-10
        h.sin_theta = std::sin(radians); h.cos_theta = std::cos(radians);
        h.children.push_back(std::move(ptr)); return h;
    }
    static Hittable new_constant_medium(Hittable boundary, double d, MaterialHandle m) {                                                                            // :201-207
        Hittable h; h.kind = ConstantMediumK; h.density = d; h.mat_handle = m; h.children.push_back(std::move(boundary)); return h;
    }
    // The SHAPE the reference builds (:77-130): a binary tree of { left, right }; a single-object span is cloned into BOTH
    // children (:96-98).  The per-node sort (random axis, :82-94, :108-110) only permutes members and is left out: the
    // backend takes membership from this tree and builds its own BVH (Backend::collect walks it like the Rust shim).
    static Hittable new_bvh_node(const std::vector<Hittable>& list, size_t start, size_t end, double t0, double t1) {
        Hittable h; h.kind = BvhNodeK; h.time_0 = t0; h.time_1 = t1;
        const size_t span = end - start;
        if (span == 1) { h.children.push_back(list[start]); h.children.push_back(list[start]); }
        else if (span == 2) { h.children.push_back(list[start]); h.children.push_back(list[start + 1]); }
        else {
            const size_t mid = start + span / 2;
            h.children.push_back(new_bvh_node(list, start, mid, t0, t1));
            h.children.push_back(new_bvh_node(list, mid, end, t0, t1));
        }
        return h;
    }
private:
    static Hittable rect(Kind kd, MaterialHandle m, double a0, double a1, double b0, double b1, double k) {
        Hittable h; h.kind = kd; h.mat_handle = m; h.a0 = a0; h.a1 = a1; h.b0 = b0; h.b1 = b1; h.k = k; return h;
    }
};

struct World {                                                        // src/main.rs:40-50
    std::vector<Material> materials;
    std::vector<Hittable> hittables;
    MaterialHandle register_material(Material m) { materials.push_back(std::move(m)); MaterialHandle h; h.v = materials.size(); return h; }
};

struct Camera {                                                       // src/camera.rs:4-15
    rtw_camera c;
    static Camera create(Point3 look_from, Point3 look_at, Vector3 vup, double vfov, double aspect_ratio, double aperture,
                         double focus_dist, double time_0, double time_1) {                                          // Camera::new
        Camera cam;
        double lf[3] = {look_from.x, look_from.y, look_from.z}, la[3] = {look_at.x, look_at.y, look_at.z}, vu[3] = {vup.x, vup.y, vup.z};
        if (rtw_camera_new(lf, la, vu, vfov, aspect_ratio, aperture, focus_dist, time_0, time_1, &cam.c) != RTW_OK) throw std::runtime_error(rtw_last_error());
        return cam;
    }
};

// ---- flatten: World -> rtw_scene (the C++ twin of `fn flatten(&World)` in INTEGRATION.md) -----------------------
class Backend {
public:
    Backend() : s_(rtw_scene_new()) { if (!s_) throw std::runtime_error("rtw_scene_new failed"); }
    ~Backend() { rtw_scene_free(s_); }
    Backend(const Backend&) = delete;
    Backend& operator=(const Backend&) = delete;
    rtw_scene* scene() { return s_; }

    void flatten(const World& w) {
        for (const Material& m : w.materials) material(m);
        for (const Hittable& h : w.hittables) ok(rtw_world_push(s_, hittable(h)));
    }
    void commit(int n_gpus = 1, int first_device = 0) { ok(rtw_scene_commit(s_, n_gpus, first_device)); }
    // per-pixel radiance SUMS, row 0 = top (src/main.rs:591)
    std::vector<float> render(const Camera& cam, int width, int height, int spp, int max_depth, Color background, uint64_t seed = 1,
                              rtw_stats* stats = nullptr) {
        rtw_render_params p{};
        p.width = width; p.height = height; p.spp = spp; p.max_depth = max_depth;
        p.background[0] = background.x; p.background[1] = background.y; p.background[2] = background.z;
        p.t_min = 0.001; p.seed = seed;
        std::vector<float> out((size_t)width * height * 3);
        rtw_stats st{};
        ok(rtw_render(s_, &cam.c, &p, out.data(), &st));
        if (stats) *stats = st;
        return out;
    }
    // The same render in passes with a progress callback (what the reference's progress thread reports, src/main.rs:557-582);
    // `sums` may carry an interrupted render (then first_sample > 0).  progress(done, total, sums) -> true to stop.
    template <class Progress>
    void render_progressive(const Camera& cam, int width, int height, int spp, int max_depth, Color background, uint64_t seed,
                            int first_sample, int samples_per_pass, std::vector<float>& sums, Progress progress, rtw_stats* stats = nullptr) {
        rtw_render_params p{};
        p.width = width; p.height = height; p.spp = spp; p.max_depth = max_depth;
        p.background[0] = background.x; p.background[1] = background.y; p.background[2] = background.z;
        p.t_min = 0.001; p.seed = seed;
        sums.resize((size_t)width * height * 3);
        rtw_stats st{};
        auto tramp = [](int32_t done, int32_t total, const float* rgb, void* user) -> int { return (*static_cast<Progress*>(user))(done, total, rgb) ? 1 : 0; };
        ok(rtw_render_progressive(s_, &cam.c, &p, first_sample, samples_per_pass, sums.data(), tramp, &progress, &st));
        if (stats) *stats = st;
    }

private:
    rtw_scene* s_;
    static int ok(int rc) { if (rc < 0) throw std::runtime_error(std::string("rtw: ") + rtw_last_error()); return rc; }
    static void v3(const Vector3& v, double o[3]) { o[0] = v.x; o[1] = v.y; o[2] = v.z; }

    int texture(const Texture& t) {
        double a[3], b[3]; v3(t.a, a); v3(t.b, b);
        switch (t.kind) {
        case Texture::SolidColorK: return ok(rtw_tex_solid(s_, a));
        case Texture::CheckerK: return ok(rtw_tex_checker(s_, a, b));
        case Texture::NoiseK: {
            std::vector<double> rv; rv.reserve(768);
            for (const Vector3& q : t.perlin->ranvec) { rv.push_back(q.x); rv.push_back(q.y); rv.push_back(q.z); }
            return ok(rtw_tex_noise(s_, rv.data(), t.perlin->perm_x.data(), t.perlin->perm_y.data(), t.perlin->perm_z.data(), t.scale));
        }
        default: return ok(rtw_tex_image(s_, (int)t.w, (int)t.h, (int)t.bytes_per_scanline, t.data->data()));
        }
    }
    void material(const Material& m) {
        double a[3]; v3(m.albedo, a);
        switch (m.kind) {
        case Material::LambertianK: ok(rtw_mat_lambertian(s_, texture(m.tex))); break;
        case Material::MetalK: ok(rtw_mat_metal(s_, a, m.fuzz)); break;
        case Material::DielectricK: ok(rtw_mat_dielectric(s_, m.ir)); break;
        case Material::DiffuseLightK: ok(rtw_mat_diffuse_light(s_, texture(m.tex))); break;
        default: ok(rtw_mat_isotropic(s_, texture(m.tex))); break;
        }
    }
    int hittable(const Hittable& h) {
        double a[3], b[3];
        const int mat = (int)h.mat_handle.v;
        switch (h.kind) {
        case Hittable::SphereK: v3(h.center, a); return ok(rtw_sphere(s_, mat, a, h.radius));
        case Hittable::MovingSphereK: v3(h.center, a); v3(h.center_1, b); return ok(rtw_moving_sphere(s_, mat, a, b, h.time_0, h.time_1, h.radius));
        case Hittable::XYRectK: return ok(rtw_xy_rect(s_, mat, h.a0, h.a1, h.b0, h.b1, h.k));
        case Hittable::XZRectK: return ok(rtw_xz_rect(s_, mat, h.a0, h.a1, h.b0, h.b1, h.k));
        case Hittable::YZRectK: return ok(rtw_yz_rect(s_, mat, h.a0, h.a1, h.b0, h.b1, h.k));
        case Hittable::BoxK: v3(h.min, a); v3(h.max, b); return ok(rtw_box(s_, a, b, mat));
        case Hittable::TranslateK: { int c = hittable(h.children[0]); v3(h.offset, a); return ok(rtw_translate(s_, c, a)); }
        case Hittable::RotateYK: { int c = hittable(h.children[0]); return ok(rtw_rotate_y_sincos(s_, h.sin_theta, h.cos_theta, c)); }
        case Hittable::ConstantMediumK: { int c = hittable(h.children[0]); return ok(rtw_constant_medium(s_, c, h.density, mat)); }
        default: {      // BvhNode: every leaf of the { left, right } tree, clones included — rtw_bvh_node drops the clones
            std::vector<int32_t> ids;
            collect(h, ids);
            return ok(rtw_bvh_node(s_, ids.data(), (int32_t)ids.size(), h.time_0, h.time_1));
        }
        }
    }
    void collect(const Hittable& h, std::vector<int32_t>& out) {      // = `collect` of integration/rust/src/gpu.rs
        if (h.kind == Hittable::BvhNodeK) { collect(h.children[0], out); collect(h.children[1], out); }
        else out.push_back(hittable(h));
    }
};

}  // namespace rtw_host
#endif
