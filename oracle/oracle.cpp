// oracle.cpp — CPU ORACLE.  TEST INFRASTRUCTURE ONLY.
//
// A scalar f64 restatement of the render hot path of
// themeshpotato/rust-ray-tracing-in-a-weekend, one section per reference module, every
// function citing the reference file:line it follows (paths relative to /root/reference).
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may
// load this library, and only as the checker / the timed CPU baseline — never on the product
// path (librtw.so has no CPU fallback and never links this file).
//
// PARITY PINNING: the reference has no tests and cannot be compiled here (no rustc/cargo, `rand`
// and `stb_image` not vendored).  The restatement is pinned to the known answers the reference
// holds — the sphere_uv table in src/math.rs:292-294 and the one output image that matches a HEAD
// scene, generated_images/earth.ppm (block means in tests/golden/ref_earth_400x225_blocks5.npy:
// correlation 0.999, |delta| 1.3/255 on the globe) — plus closed-form values of the reference
// formulas (tests/test_oracle_kat.py).  RNG (rand 0.8 thread_rng, OS-seeded) is third-party and
// unpinned ("parity unpinned": only the distributions carry over); JPEG decode (stb_image) is
// pinned only through that image.  See DESIGN.md section 2.
//
// Deliberate departures (none changes a distribution):
//  * random_double() is replaced by an injectable stream: either an explicit array of U[0,1)
//    draws, or Philox4x32-10 keyed by seed with counter (draw block, bounce, pixel, sample);
//    a draw is (word >> 8) * 2^-24 so that the f32 device path consumes the identical value.
//    The two rejection loops (unit sphere, unit disk) take their attempts from blocks of their own,
//    indexed by the attempt number (Rng::sphere_attempt / disk_attempt) — same distribution, same
//    acceptance test; explicit streams stay sequential like the reference.
//  * random_double_range(a,b) (inclusive in rand 0.8, src/math.rs:273-276) is a + (b-a)*xi.
//  * optional `media_deferred` world scan (all non-medium items first, then the media in list
//    order) — the order the device uses; the default is the literal list order.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <vector>

#include "../include/rtw.h"

namespace {

// ---------------------------------------------------------------------------------------------
// src/math.rs
// ---------------------------------------------------------------------------------------------
const double PI = 3.1415926535897932385;                       // math.rs:5
const double INF = std::numeric_limits<double>::infinity();    // math.rs:6
inline double degrees_to_radians(double d) { return d * PI / 180.0; }  // math.rs:8-10

struct V3 {                                                    // math.rs:12-20
    double x = 0, y = 0, z = 0;
    V3() {}
    V3(double a, double b, double c) : x(a), y(b), z(c) {}
    double operator[](int i) const { return i == 0 ? x : (i == 1 ? y : z); }  // as_array :31-33
};
inline V3 operator+(V3 a, V3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }   // :147-157
inline V3 operator-(V3 a, V3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }   // :175-185
inline V3 operator-(V3 a) { return V3(-a.x, -a.y, -a.z); }                          // :187-209
inline V3 operator*(V3 a, V3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }   // :211-221
inline V3 operator*(V3 a, double s) { return V3(a.x * s, a.y * s, a.z * s); }      // :223-233
inline V3 operator*(double s, V3 a) { return V3(a.x * s, a.y * s, a.z * s); }      // :235-257
inline V3 operator/(V3 a, double s) { return (1.0 / s) * a; }                       // :260-266 (reciprocal!)
inline double dot(V3 u, V3 v) { return u.x * v.x + u.y * v.y + u.z * v.z; }        // :82-84
inline double length_squared(V3 v) { return v.x * v.x + v.y * v.y + v.z * v.z; }   // :86-88
inline double length(V3 v) { return std::sqrt(length_squared(v)); }                 // :90-92
inline V3 cross(V3 u, V3 v) {                                                       // :94-100
    return V3(u.y * v.z - u.z * v.y, u.z * v.x - u.x * v.z, u.x * v.y - u.y * v.x);
}
inline V3 normalize(V3 v) { return v / length(v); }                                 // :102-104
inline V3 reflect(V3 v, V3 n) { return v - 2.0 * dot(v, n) * n; }                  // :106-108
inline V3 refract(V3 uv, V3 n, double etai_over_etat) {                             // :110-117
    double cos_theta = std::fmin(dot(-uv, n), 1.0);
    V3 r_out_perp = etai_over_etat * (uv + cos_theta * n);
    double r_out_perp_length = length_squared(r_out_perp);
    V3 r_out_parallel = -std::sqrt(std::fabs(1.0 - r_out_perp_length)) * n;
    return r_out_perp + r_out_parallel;
}
inline bool near_zero(V3 v) {                                                        // :134-137
    const double S = 1e-8;
    return std::fabs(v.x) < S && std::fabs(v.y) < S && std::fabs(v.z) < S;
}
inline double clampd(double x, double lo, double hi) { return x < lo ? lo : (x > hi ? hi : x); }  // :282-286
inline void sphere_uv(V3 p, double& u, double& v) {                                  // :288-300
    double theta = std::acos(-p.y);
    double phi = std::atan2(-p.z, p.x) + PI;
    u = phi / (2.0 * PI);
    v = theta / PI;
}
// write_color (math.rs:119-132): gamma 2, clamp [0,0.999], *256 truncated; NaN -> 0 (Rust `as i32`).
inline int to_byte(double sum, int spp) {
    double scale = 1.0 / (double)spp;
    double r = std::sqrt(sum * scale);
    double c = 256.0 * clampd(r, 0.0, 0.999);
    if (c != c) return 0;
    return (int)c;
}

// ---------------------------------------------------------------------------------------------
// RNG — replaces rand::thread_rng (math.rs:268-280); see header comment.
// ---------------------------------------------------------------------------------------------
inline void philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    // Salmon et al., "Parallel random numbers: as easy as 1, 2, 3" (SC'11), Philox-4x32 with 10 rounds.
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct Rng {
    // mode 0: explicit stream
    const double* stream = nullptr;
    int stream_len = 0;
    // mode 1: philox
    bool philox = false;
    uint32_t key[2] = {0, 0};
    uint32_t pixel = 0, sample = 0, bounce = 0;
    uint32_t cached_block = 0xffffffffu;
    uint32_t words[4];
    // common
    int draw = 0;        // draws consumed in the current stream / bounce
    long total = 0;
    bool exhausted = false;
    void set_bounce(uint32_t b) { bounce = b; draw = 0; cached_block = 0xffffffffu; }
    double next() {                                       // = random_double() math.rs:268-271
        ++total;
        if (!philox) {
            if (draw >= stream_len) { exhausted = true; ++draw; return 0.5; }
            return stream[draw++];
        }
        uint32_t blk = (uint32_t)draw >> 2;
        if (blk != cached_block) {
            uint32_t ctr[4] = {blk, bounce, pixel, sample};
            philox4x32_10(ctr, key, words);
            cached_block = blk;
        }
        uint32_t w = words[draw & 3];
        ++draw;
        return (double)(w >> 8) * (1.0 / 16777216.0);
    }
    double range(double a, double b) { return a + (b - a) * next(); }  // random_double_range :273-276
    // The rejection loops (math.rs:51-58, :69-76) draw from their OWN Philox blocks, indexed by the attempt: attempt a of the
    // unit-sphere loop of this bounce is block 0x40000000 + a (words 0..2 = x, y, z), attempt a of the unit-disk loop is
    // words 2(a & 1), 2(a & 1) + 1 of block 0x20000000 + (a >> 1).  They do not advance the scalar stream.  (The device can
    // then evaluate ANY attempt of ANY lane: finished lanes help unlucky ones, csrc/rtw_device.cuh coop_unit_sphere.)
    // With an explicit stream (parity hooks) the draws stay sequential, as in the reference.
    static double u01(uint32_t w) { return (double)(w >> 8) * (1.0 / 16777216.0); }
    void sphere_attempt(uint32_t a, double& x, double& y, double& z) {
        if (!philox) { x = next(); y = next(); z = next(); return; }
        total += 3;
        uint32_t ctr[4] = {0x40000000u + a, bounce, pixel, sample}, w[4];
        philox4x32_10(ctr, key, w);
        x = u01(w[0]); y = u01(w[1]); z = u01(w[2]);
    }
    void disk_attempt(uint32_t a, double& x, double& y) {
        if (!philox) { x = next(); y = next(); return; }
        total += 2;
        uint32_t ctr[4] = {0x20000000u + (a >> 1), bounce, pixel, sample}, w[4];
        philox4x32_10(ctr, key, w);
        x = u01(w[2 * (a & 1)]); y = u01(w[2 * (a & 1) + 1]);
    }
};

inline V3 random_range_v3(Rng& g, double a, double b) {          // Vector3::random_range :43-49 (x, y, z order)
    double x = g.range(a, b); double y = g.range(a, b); double z = g.range(a, b);
    return V3(x, y, z);
}
inline V3 random_in_unit_sphere(Rng& g) {                         // :51-58; Vector3::random_range(-1, 1) :43-49 per attempt
    for (uint32_t a = 0;; ++a) {
        double x, y, z;
        g.sphere_attempt(a, x, y, z);
        V3 p(-1.0 + 2.0 * x, -1.0 + 2.0 * y, -1.0 + 2.0 * z);      // a + (b - a) * random_double() :273-276
        if (length_squared(p) < 1.0) return p;
    }
}
inline V3 random_in_unit_disk(Rng& g) {                           // :69-76
    for (uint32_t a = 0;; ++a) {
        double x, y;
        g.disk_attempt(a, x, y);
        V3 p(-1.0 + 2.0 * x, -1.0 + 2.0 * y, 0.0);
        if (length_squared(p) < 1.0) return p;
    }
}
inline V3 random_unit_vector(Rng& g) { return normalize(random_in_unit_sphere(g)); }  // :78-80

// ---------------------------------------------------------------------------------------------
// src/ray.rs
// ---------------------------------------------------------------------------------------------
struct Ray {                                                       // ray.rs:3-7
    V3 origin, direction; double time = 0;
    V3 at(double t) const { return origin + t * direction; }      // ray.rs:19-21
};

// ---------------------------------------------------------------------------------------------
// src/aabb.rs
// ---------------------------------------------------------------------------------------------
struct AABB { V3 minimum, maximum; };                              // aabb.rs:6-9
inline AABB surrounding_box(const AABB& a, const AABB& b) {        // aabb.rs:19-33
    AABB r;
    r.minimum = V3(std::fmin(a.minimum.x, b.minimum.x), std::fmin(a.minimum.y, b.minimum.y), std::fmin(a.minimum.z, b.minimum.z));
    r.maximum = V3(std::fmax(a.maximum.x, b.maximum.x), std::fmax(a.maximum.y, b.maximum.y), std::fmax(a.maximum.z, b.maximum.z));
    return r;
}
inline bool aabb_hit(const AABB& b, const Ray& ray, double t_min, double t_max) {   // aabb.rs:77-103
    double mn = t_min, mx = t_max;
    for (int a = 0; a < 3; ++a) {
        double inv_d = 1.0 / ray.direction[a];
        double t0 = (b.minimum[a] - ray.origin[a]) * inv_d;
        double t1 = (b.maximum[a] - ray.origin[a]) * inv_d;
        if (inv_d < 0.0) std::swap(t0, t1);
        mn = t0 > mn ? t0 : mn;
        mx = t1 < mx ? t1 : mx;
        if (mx <= mn) return false;
    }
    return true;
}

// ---------------------------------------------------------------------------------------------
// src/perlin.rs (device-side part: noise / perlin_interp / turb; tables come through the ABI)
// ---------------------------------------------------------------------------------------------
struct Perlin {                                                    // perlin.rs:5-10
    V3 ranvec[256];
    int perm_x[256], perm_y[256], perm_z[256];
};
inline double perlin_interp(const V3 c[2][2][2], double u, double v, double w) {   // perlin.rs:70-94
    double uu = u * u * (3.0 - 2.0 * u);
    double vv = v * v * (3.0 - 2.0 * v);
    double ww = w * w * (3.0 - 2.0 * w);
    double accum = 0.0;
    for (int i = 0; i < 2; ++i)
        for (int j = 0; j < 2; ++j)
            for (int k = 0; k < 2; ++k) {
                V3 val = c[i][j][k];
                double fi = i, fj = j, fk = k;
                V3 weight_v(u - fi, v - fj, w - fk);
                accum += (fi * uu + (1.0 - fi) * (1.0 - uu)) *
                         (fj * vv + (1.0 - fj) * (1.0 - vv)) *
                         (fk * ww + (1.0 - fk) * (1.0 - ww)) * dot(val, weight_v);
            }
    return accum;
}
inline int f2i_sat(double x) {   // Rust `as i32`: saturating, NaN -> 0
    if (x != x) return 0;
    if (x >= 2147483647.0) return 2147483647;
    if (x <= -2147483648.0) return (int)(-2147483647 - 1);
    return (int)x;
}
inline double perlin_noise(const Perlin& pl, V3 p) {               // perlin.rs:32-68
    double x = std::floor(p.x), y = std::floor(p.y), z = std::floor(p.z);
    double u = p.x - x, v = p.y - y, w = p.z - z;
    u = u * u * (3.0 - 2.0 * u);                                   // first smoothstep (:41-43)
    v = v * v * (3.0 - 2.0 * v);
    w = w * w * (3.0 - 2.0 * w);
    int i = f2i_sat(x), j = f2i_sat(y), k = f2i_sat(z);
    V3 c[2][2][2];
    for (int di = 0; di < 2; ++di)
        for (int dj = 0; dj < 2; ++dj)
            for (int dk = 0; dk < 2; ++dk) {
                unsigned xi = (unsigned)((int)((unsigned)i + (unsigned)di) & 255);
                unsigned yi = (unsigned)((int)((unsigned)j + (unsigned)dj) & 255);
                unsigned zi = (unsigned)((int)((unsigned)k + (unsigned)dk) & 255);
                c[di][dj][dk] = pl.ranvec[(unsigned)(pl.perm_x[xi] ^ pl.perm_y[yi] ^ pl.perm_z[zi]) & 255u];
            }
    return perlin_interp(c, u, v, w);                              // second smoothstep inside (:71-73)
}
inline double perlin_turb(const Perlin& pl, V3 p, int depth) {     // perlin.rs:96-108
    double accum = 0.0; V3 temp_p = p; double weight = 1.0;
    for (int i = 0; i < depth; ++i) {
        accum += weight * perlin_noise(pl, temp_p);
        weight *= 0.5;
        temp_p = temp_p * 2.0;
    }
    return std::fabs(accum);
}

// ---------------------------------------------------------------------------------------------
// src/texture.rs
// ---------------------------------------------------------------------------------------------
enum TexKind { TEX_SOLID = 0, TEX_CHECKER = 1, TEX_NOISE = 2, TEX_IMAGE = 3 };
struct Texture {                                                   // texture.rs:4-9
    int kind = TEX_SOLID;
    V3 c0, c1;                 // Solid: c0; Checker: even=c0, odd=c1
    double scale = 1.0;
    Perlin* perlin = nullptr;
    int w = 0, h = 0, bps = 0;
    std::vector<uint8_t> data;
};

struct Counters {   // event counts for the flop model
    uint64_t paths = 0, rays = 0, aabb = 0, sphere = 0, sphere_accept = 0, moving = 0, rect = 0, rect_accept = 0,
             translate = 0, rotate = 0, medium = 0, scatter[5] = {0, 0, 0, 0, 0}, tex[4] = {0, 0, 0, 0},
             accum = 0, draws = 0;
    void add(const Counters& o) {
        paths += o.paths; rays += o.rays; aabb += o.aabb; sphere += o.sphere; sphere_accept += o.sphere_accept;
        moving += o.moving; rect += o.rect; rect_accept += o.rect_accept; translate += o.translate;
        rotate += o.rotate; medium += o.medium; accum += o.accum; draws += o.draws;
        for (int i = 0; i < 5; ++i) scatter[i] += o.scatter[i];
        for (int i = 0; i < 4; ++i) tex[i] += o.tex[i];
    }
};

struct Ctx { Rng* rng; Counters* cnt; };

inline V3 texture_value(const Texture& t, double u, double v, V3 p, Counters* cnt) {   // texture.rs:30-75
    if (cnt) cnt->tex[t.kind]++;
    switch (t.kind) {
    case TEX_SOLID: return t.c0;                                   // :32-34
    case TEX_CHECKER: {                                            // :35-42
        double sines = std::sin(10.0 * p.x) * std::sin(10.0 * p.y) * std::sin(10.0 * p.z);
        return sines < 0.0 ? t.c1 : t.c0;
    }
    case TEX_NOISE:                                                // :43-45
        return V3(1.0, 1.0, 1.0) * 0.5 * (1.0 + std::sin(t.scale * p.z + 10.0 * perlin_turb(*t.perlin, p, 7)));
    default: {                                                     // :46-73
        double uu = clampd(u, 0.0, 1.0);
        double vv = 1.0 - clampd(v, 0.0, 1.0);
        double fi = uu * (double)t.w, fj = vv * (double)t.h;
        size_t i = fi != fi ? 0 : (size_t)fi;                      // `as usize`: NaN -> 0
        size_t j = fj != fj ? 0 : (size_t)fj;
        if (i >= (size_t)t.w) i = t.w - 1;
        if (j >= (size_t)t.h) j = t.h - 1;
        const double color_scale = 1.0 / 255.0;
        const uint8_t* ptr = t.data.data() + (j * (size_t)t.bps + i * 3);
        return V3(color_scale * ptr[0], color_scale * ptr[1], color_scale * ptr[2]);
    }
    }
}

// ---------------------------------------------------------------------------------------------
// src/material.rs
// ---------------------------------------------------------------------------------------------
enum MatKind { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2, MAT_DIFFUSE_LIGHT = 3, MAT_ISOTROPIC = 4 };
struct Material {                                                  // material.rs:6-12
    int kind = MAT_LAMBERTIAN;
    int tex = -1;            // Lambertian/DiffuseLight/Isotropic
    V3 albedo;               // Metal
    double fuzz = 0, ir = 1;
};

struct HitRecord {                                                 // hittable.rs:6-15
    V3 point, normal; double t = 0; bool front_face = false; int mat_handle = 0; double u = 0, v = 0;
    void set_face_normal(const Ray& ray, V3 outward_normal) {     // hittable.rs:23-26
        front_face = dot(ray.direction, outward_normal) < 0.0;
        normal = front_face ? outward_normal : -outward_normal;
    }
};

inline double reflectance(double cosine, double ref_idx) {         // material.rs:89-94
    double r0 = (1.0 - ref_idx) / (1.0 + ref_idx);
    r0 = r0 * r0;
    return r0 + (1.0 - r0) * std::pow(1.0 - cosine, 5.0);
}

struct Scene;
V3 material_emitted(const Scene& sc, const Material& m, double u, double v, V3 p, Counters* cnt);
bool material_scatter(const Scene& sc, const Material& m, const Ray& ray, const HitRecord& rec, Ctx& cx,
                      Ray& scattered, V3& attenuation);

// ---------------------------------------------------------------------------------------------
// src/hittable.rs
// ---------------------------------------------------------------------------------------------
enum HKind { H_SPHERE, H_MOVING_SPHERE, H_BVH_NODE, H_XY, H_XZ, H_YZ, H_BOX, H_TRANSLATE, H_ROTATE_Y, H_MEDIUM };
struct Hittable {                                                  // hittable.rs:29-41
    int kind;
    int mat = 0;                         // mat_handle / phase_function (1-based)
    V3 c0, c1; double radius = 0, time0 = 0, time1 = 0;           // spheres
    double a0 = 0, a1 = 0, b0 = 0, b1 = 0, k = 0;                 // rects
    V3 bmin, bmax; std::vector<int> sides;                        // Box
    V3 offset; int child = -1;                                     // Translate / RotateY / ConstantMedium
    double sin_theta = 0, cos_theta = 1; bool has_box = false; AABB bbox;
    double neg_inv_density = 0;
    int left = -1, right = -1; AABB aabb_box;                     // BvhNode
};

struct Scene {
    std::vector<Texture> textures;
    std::vector<Perlin*> perlins;
    std::vector<Material> materials;     // handle h -> materials[h-1]   (main.rs:26, :46-49)
    std::vector<Hittable> nodes;
    std::vector<int> world;              // world.hittables (main.rs:42)
    bool media_deferred = false;
    uint64_t build_rng = 0x9E3779B97F4A7C15ull;
    ~Scene() { for (auto* p : perlins) delete p; }
    double build_random() {              // host RNG for new_bvh_node's axis pick (hittable.rs:82)
        build_rng += 0x9E3779B97F4A7C15ull;
        uint64_t z = build_rng;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        z ^= z >> 31;
        return (double)(z >> 11) * (1.0 / 9007199254740992.0);
    }
};

inline V3 get_center_at_time(V3 c0, V3 c1, double time0, double time1, double time) {   // hittable.rs:556-558
    return c0 + ((time - time0) / (time1 - time0)) * (c1 - c0);
}

bool hittable_hit(const Scene& sc, int id, const Ray& ray, double t_min, double t_max, Ctx& cx, HitRecord& rec);

bool hit_list(const Scene& sc, const std::vector<int>& list, const Ray& ray, double t_min, double t_max,
              Ctx& cx, HitRecord& rec) {                                                // hittable.rs:43-55
    double closest_so_far = t_max;
    bool any = false;
    HitRecord tmp;
    for (int id : list) {
        if (hittable_hit(sc, id, ray, t_min, closest_so_far, cx, tmp)) {
            closest_so_far = tmp.t;
            rec = tmp;
            any = true;
        }
    }
    return any;
}

// World-level scan.  media_deferred=false is the literal hit_hittables over world.hittables;
// true visits non-medium items first, then the media (same relative order) — identical in
// distribution (a medium's free-flight draw vs the nearest surface), and the order the device uses.
bool world_hit(const Scene& sc, const Ray& ray, double t_min, double t_max, Ctx& cx, HitRecord& rec) {
    if (cx.cnt) cx.cnt->rays++;
    if (!sc.media_deferred) return hit_list(sc, sc.world, ray, t_min, t_max, cx, rec);
    double closest_so_far = t_max;
    bool any = false;
    HitRecord tmp;
    for (int pass = 0; pass < 2; ++pass)
        for (int id : sc.world) {
            bool is_medium = sc.nodes[id].kind == H_MEDIUM;
            if ((pass == 1) != is_medium) continue;
            if (hittable_hit(sc, id, ray, t_min, closest_so_far, cx, tmp)) {
                closest_so_far = tmp.t; rec = tmp; any = true;
            }
        }
    return any;
}

bool sphere_hit(V3 center, double radius, const Ray& ray, double t_min, double t_max, int mat, Ctx& cx,
                HitRecord& rec) {                                                       // hittable.rs:254-288
    if (cx.cnt) cx.cnt->sphere++;
    V3 oc = ray.origin - center;
    double a = length_squared(ray.direction);
    double half_b = dot(oc, ray.direction);
    double c = length_squared(oc) - radius * radius;
    double discriminant = half_b * half_b - a * c;
    if (discriminant < 0.0) return false;
    double sqrtd = std::sqrt(discriminant);
    double root = (-half_b - sqrtd) / a;
    if (root < t_min || t_max < root) {
        root = (-half_b + sqrtd) / a;
        if (root < t_min || t_max < root) return false;
    }
    if (cx.cnt) cx.cnt->sphere_accept++;
    rec = HitRecord();
    rec.mat_handle = mat;
    rec.t = root;
    rec.point = ray.at(rec.t);
    V3 outward_normal = (rec.point - center) / radius;
    rec.set_face_normal(ray, outward_normal);
    sphere_uv(outward_normal, rec.u, rec.v);
    return true;
}

// axis: 0 = XY (normal +z), 1 = XZ (normal +y), 2 = YZ (normal +x)          hittable.rs:308-384
bool rect_hit(int axis, double a0, double a1, double b0, double b1, double k, const Ray& ray, double t_min,
              double t_max, int mat, Ctx& cx, HitRecord& rec) {
    if (cx.cnt) cx.cnt->rect++;
    double ok, dk, oa, da, ob, db;
    V3 outward;
    if (axis == 0) { ok = ray.origin.z; dk = ray.direction.z; oa = ray.origin.x; da = ray.direction.x; ob = ray.origin.y; db = ray.direction.y; outward = V3(0, 0, 1); }
    else if (axis == 1) { ok = ray.origin.y; dk = ray.direction.y; oa = ray.origin.x; da = ray.direction.x; ob = ray.origin.z; db = ray.direction.z; outward = V3(0, 1, 0); }
    else { ok = ray.origin.x; dk = ray.direction.x; oa = ray.origin.y; da = ray.direction.y; ob = ray.origin.z; db = ray.direction.z; outward = V3(1, 0, 0); }
    double t = (k - ok) / dk;
    if (t < t_min || t > t_max) return false;
    double a = oa + t * da;
    double b = ob + t * db;
    if (a < a0 || a > a1 || b < b0 || b > b1) return false;
    if (cx.cnt) cx.cnt->rect_accept++;
    rec = HitRecord();
    rec.u = (a - a0) / (a1 - a0);
    rec.v = (b - b0) / (b1 - b0);
    rec.t = t;
    rec.set_face_normal(ray, outward);
    rec.mat_handle = mat;
    rec.point = ray.at(t);
    return true;
}

bool bvh_node_hit(const Scene& sc, const Hittable& h, const Ray& ray, double t_min, double t_max, Ctx& cx,
                  HitRecord& rec) {                                                     // hittable.rs:290-306
    if (cx.cnt) cx.cnt->aabb++;
    if (!aabb_hit(h.aabb_box, ray, t_min, t_max)) return false;
    HitRecord hl;
    if (hittable_hit(sc, h.left, ray, t_min, t_max, cx, hl)) {
        HitRecord hr;
        if (hittable_hit(sc, h.right, ray, t_min, hl.t, cx, hr)) rec = hr; else rec = hl;
        return true;
    }
    return hittable_hit(sc, h.right, ray, t_min, t_max, cx, rec);
}

bool hit_rotate_y(const Scene& sc, const Hittable& h, const Ray& ray, double t_min, double t_max, Ctx& cx,
                  HitRecord& rec) {                                                     // hittable.rs:386-415
    if (cx.cnt) cx.cnt->rotate++;
    double s = h.sin_theta, c = h.cos_theta;
    V3 origin = ray.origin, direction = ray.direction;
    origin.x = c * ray.origin.x - s * ray.origin.z;
    origin.z = s * ray.origin.x + c * ray.origin.z;
    direction.x = c * ray.direction.x - s * ray.direction.z;
    direction.z = s * ray.direction.x + c * ray.direction.z;
    Ray rotated{origin, direction, ray.time};
    if (!hittable_hit(sc, h.child, rotated, t_min, t_max, cx, rec)) return false;
    V3 p = rec.point, normal = rec.normal;
    p.x = c * rec.point.x + s * rec.point.z;
    p.z = -s * rec.point.x + c * rec.point.z;
    normal.x = c * rec.normal.x + s * rec.normal.z;
    normal.z = -s * rec.normal.x + c * rec.normal.z;
    rec.point = p;
    rec.set_face_normal(rotated, normal);     // quirk: object-space ray vs world-space normal (:409)
    return true;
}

bool hit_constant_medium(const Scene& sc, const Hittable& h, const Ray& ray, double t_min, double t_max, Ctx& cx,
                         HitRecord& rec) {                                              // hittable.rs:417-473
    if (cx.cnt) cx.cnt->medium++;
    HitRecord rec1, rec2;
    if (!hittable_hit(sc, h.child, ray, -INF, INF, cx, rec1)) return false;             // :422
    if (!hittable_hit(sc, h.child, ray, rec1.t + 0.0001, INF, cx, rec2)) return false;  // :423
    if (rec1.t < t_min) rec1.t = t_min;
    if (rec2.t > t_max) rec2.t = t_max;
    if (rec1.t >= rec2.t) return false;
    if (rec1.t < 0.0) rec1.t = 0.0;
    double ray_length = length(ray.direction);
    double distance_inside_boundary = (rec2.t - rec1.t) * ray_length;
    double hit_distance = h.neg_inv_density * std::log(cx.rng->next());                // :446 (the draw)
    if (hit_distance > distance_inside_boundary) return false;
    rec = HitRecord();
    rec.t = rec1.t + hit_distance / ray_length;
    rec.point = ray.at(rec.t);
    rec.normal = V3(1.0, 0.0, 0.0);
    rec.front_face = true;
    rec.mat_handle = h.mat;
    return true;
}

bool hittable_hit(const Scene& sc, int id, const Ray& ray, double t_min, double t_max, Ctx& cx,
                  HitRecord& rec) {                                                     // hittable.rs:209-252
    const Hittable& h = sc.nodes[id];
    switch (h.kind) {
    case H_SPHERE: return sphere_hit(h.c0, h.radius, ray, t_min, t_max, h.mat, cx, rec);
    case H_MOVING_SPHERE:
        if (cx.cnt) cx.cnt->moving++;
        return sphere_hit(get_center_at_time(h.c0, h.c1, h.time0, h.time1, ray.time), h.radius, ray, t_min, t_max,
                          h.mat, cx, rec);
    case H_BVH_NODE: return bvh_node_hit(sc, h, ray, t_min, t_max, cx, rec);
    case H_XY: return rect_hit(0, h.a0, h.a1, h.b0, h.b1, h.k, ray, t_min, t_max, h.mat, cx, rec);
    case H_XZ: return rect_hit(1, h.a0, h.a1, h.b0, h.b1, h.k, ray, t_min, t_max, h.mat, cx, rec);
    case H_YZ: return rect_hit(2, h.a0, h.a1, h.b0, h.b1, h.k, ray, t_min, t_max, h.mat, cx, rec);
    case H_BOX: return hit_list(sc, h.sides, ray, t_min, t_max, cx, rec);               // :229-231
    case H_TRANSLATE: {                                                                 // :232-244
        if (cx.cnt) cx.cnt->translate++;
        Ray moved{ray.origin - h.offset, ray.direction, ray.time};
        if (!hittable_hit(sc, h.child, moved, t_min, t_max, cx, rec)) return false;
        rec.point = rec.point + h.offset;
        V3 normal = rec.normal;
        rec.set_face_normal(moved, normal);
        return true;
    }
    case H_ROTATE_Y: return hit_rotate_y(sc, h, ray, t_min, t_max, cx, rec);
    default: return hit_constant_medium(sc, h, ray, t_min, t_max, cx, rec);
    }
}

bool bounding_box(const Scene& sc, int id, double time0, double time1, AABB& out) {   // hittable.rs:475-525
    const Hittable& h = sc.nodes[id];
    switch (h.kind) {
    case H_SPHERE:                                                                      // :527-534
        out.minimum = h.c0 - V3(h.radius, h.radius, h.radius);
        out.maximum = h.c0 + V3(h.radius, h.radius, h.radius);
        return true;
    case H_MOVING_SPHERE: {       // :480-482 — the arm's own time_0/time_1 shadow the arguments
        V3 r(h.radius, h.radius, h.radius);
        V3 ca = get_center_at_time(h.c0, h.c1, h.time0, h.time1, h.time0);
        V3 cb = get_center_at_time(h.c0, h.c1, h.time0, h.time1, h.time1);
        AABB b0{ca - r, ca + r}, b1{cb - r, cb + r};
        out = surrounding_box(b0, b1);
        return true;
    }
    case H_BVH_NODE: out = h.aabb_box; return true;
    case H_XY: out.minimum = V3(h.a0, h.b0, h.k - 0.0001); out.maximum = V3(h.a1, h.b1, h.k + 0.0001); return true;
    case H_XZ: out.minimum = V3(h.a0, h.k - 0.0001, h.b0); out.maximum = V3(h.a1, h.k + 0.0001, h.b1); return true;
    case H_YZ: out.minimum = V3(h.k - 0.0001, h.a0, h.b0); out.maximum = V3(h.k + 0.0001, h.a1, h.b1); return true;
    case H_BOX: out.minimum = h.bmin; out.maximum = h.bmax; return true;
    case H_TRANSLATE: {
        AABB b;
        if (!bounding_box(sc, h.child, time0, time1, b)) return false;
        out.minimum = b.minimum + h.offset; out.maximum = b.maximum + h.offset;
        return true;
    }
    case H_ROTATE_Y: if (h.has_box) { out = h.bbox; return true; } return false;
    default: return bounding_box(sc, h.child, time0, time1, out);
    }
}

// new_bvh_node (hittable.rs:77-130): random axis, sort on bounding_box(0,0).minimum[axis], median split,
// single-object leaves duplicated into both children.  `objs` holds hittable ids (the reference deep-clones
// the list per node, :78 — same tree, without the O(N^2) copies).
int build_bvh(Scene& sc, std::vector<int>& objs, size_t start, size_t end, double time0, double time1) {
    int axis = (int)(0.0 + (3.0 - 0.0) * sc.build_random());   // random_int_range(0, 2) = range(0, 3) as i32
    if (axis > 2) axis = 2;
    auto less = [&](int a, int b) {                              // aabb.rs:35-48
        AABB ba, bb;
        bounding_box(sc, a, 0.0, 0.0, ba); bounding_box(sc, b, 0.0, 0.0, bb);
        return ba.minimum[axis] < bb.minimum[axis];
    };
    int left, right;
    size_t span = end - start;
    if (span == 1) { left = right = objs[start]; }
    else if (span == 2) {
        if (less(objs[start], objs[start + 1])) { left = objs[start]; right = objs[start + 1]; }
        else { left = objs[start + 1]; right = objs[start]; }
    } else {
        std::stable_sort(objs.begin() + start, objs.begin() + end, less);
        size_t mid = start + span / 2;
        left = build_bvh(sc, objs, start, mid, time0, time1);
        right = build_bvh(sc, objs, mid, end, time0, time1);
    }
    Hittable h; h.kind = H_BVH_NODE; h.left = left; h.right = right;
    AABB bl, br;
    if (bounding_box(sc, left, time0, time1, bl) && bounding_box(sc, right, time0, time1, br)) h.aabb_box = surrounding_box(bl, br);
    else h.aabb_box = AABB{V3(0, 0, 0), V3(0, 0, 0)};
    sc.nodes.push_back(h);
    return (int)sc.nodes.size() - 1;
}

// ---------------------------------------------------------------------------------------------
// src/material.rs (scatter / emitted)
// ---------------------------------------------------------------------------------------------
V3 material_emitted(const Scene& sc, const Material& m, double u, double v, V3 p, Counters* cnt) {  // :25-34
    if (m.kind == MAT_DIFFUSE_LIGHT) return texture_value(sc.textures[m.tex], u, v, p, cnt);
    return V3(0, 0, 0);
}
bool material_scatter(const Scene& sc, const Material& m, const Ray& ray, const HitRecord& rec, Ctx& cx,
                      Ray& scattered, V3& attenuation) {                                            // :15-23
    if (cx.cnt) cx.cnt->scatter[m.kind]++;
    Rng& g = *cx.rng;
    switch (m.kind) {
    case MAT_LAMBERTIAN: {                                                                          // :36-48
        V3 scatter_direction = rec.normal + random_unit_vector(g);
        if (near_zero(scatter_direction)) scatter_direction = rec.normal;
        scattered = Ray{rec.point, scatter_direction, ray.time};
        attenuation = texture_value(sc.textures[m.tex], rec.u, rec.v, rec.point, cx.cnt);
        return true;
    }
    case MAT_METAL: {                                                                               // :50-60
        V3 reflected = reflect(normalize(ray.direction), rec.normal);
        V3 with_fuzz = reflected + m.fuzz * random_in_unit_sphere(g);
        scattered = Ray{rec.point, with_fuzz, ray.time};
        if (dot(scattered.direction, rec.normal) > 0.0) { attenuation = m.albedo; return true; }
        return false;
    }
    case MAT_DIELECTRIC: {                                                                          // :62-82
        attenuation = V3(1.0, 1.0, 1.0);
        double refraction_ratio = rec.front_face ? 1.0 / m.ir : m.ir;
        V3 unit_direction = normalize(ray.direction);
        double cos_theta = std::fmin(dot(-unit_direction, rec.normal), 1.0);
        double sin_theta = std::sqrt(1.0 - cos_theta * cos_theta);
        bool cannot_refract = refraction_ratio * sin_theta > 1.0;
        V3 direction;
        if (cannot_refract || reflectance(cos_theta, refraction_ratio) > g.next())   // short-circuit :72
            direction = reflect(unit_direction, rec.normal);
        else
            direction = refract(unit_direction, rec.normal, refraction_ratio);
        scattered = Ray{rec.point, direction, ray.time};
        return true;
    }
    case MAT_DIFFUSE_LIGHT: return false;                                                           // :20
    default: {                                                                                      // :84-87
        scattered = Ray{rec.point, random_in_unit_sphere(g), ray.time};
        attenuation = texture_value(sc.textures[m.tex], rec.u, rec.v, rec.point, cx.cnt);
        return true;
    }
    }
}

// ---------------------------------------------------------------------------------------------
// src/camera.rs
// ---------------------------------------------------------------------------------------------
inline V3 v3(const double a[3]) { return V3(a[0], a[1], a[2]); }
inline void put(double a[3], V3 v) { a[0] = v.x; a[1] = v.y; a[2] = v.z; }

void camera_new(V3 look_from, V3 look_at, V3 vup, double vfov, double aspect_ratio, double aperture,
                double focus_dist, double time0, double time1, rtw_camera* out) {                  // camera.rs:18-56
    double theta = degrees_to_radians(vfov);
    double h = std::tan(theta / 2.0);
    double viewport_height = 2.0 * h;
    double viewport_width = aspect_ratio * viewport_height;
    V3 w = normalize(look_from - look_at);
    V3 u = normalize(cross(vup, w));
    V3 v = cross(w, u);
    V3 origin = look_from;
    V3 horizontal = focus_dist * viewport_width * u;
    V3 vertical = focus_dist * viewport_height * v;
    V3 lower_left_corner = origin - horizontal * 0.5 - vertical * 0.5 - focus_dist * w;
    put(out->origin, origin); put(out->lower_left_corner, lower_left_corner);
    put(out->horizontal, horizontal); put(out->vertical, vertical);
    put(out->u, u); put(out->v, v); put(out->w, w);
    out->lens_radius = aperture * 0.5;
    out->time0 = time0; out->time1 = time1;
}
Ray camera_get_ray(const rtw_camera& cam, double s, double t, Rng& g) {                             // camera.rs:58-66
    V3 rd = cam.lens_radius * random_in_unit_disk(g);
    V3 offset = v3(cam.u) * rd.x + v3(cam.v) * rd.y;
    V3 o = v3(cam.origin) + offset;
    V3 d = v3(cam.lower_left_corner) + s * v3(cam.horizontal) + t * v3(cam.vertical) - v3(cam.origin) - offset;
    double time = g.range(cam.time0, cam.time1);
    return Ray{o, d, time};
}

// ---------------------------------------------------------------------------------------------
// src/main.rs: ray_color (:19-38) and the pixel-sample loop (:512-523)
// ---------------------------------------------------------------------------------------------
V3 ray_color(const Scene& sc, const Ray& ray, V3 background, int depth, int max_depth, double t_min, Ctx& cx,
             int* segments) {
    if (depth <= 0) return V3(0, 0, 0);                                                // :21-23
    if (cx.rng->philox) cx.rng->set_bounce((uint32_t)(max_depth - depth + 1));
    if (segments) ++*segments;
    HitRecord rec;
    if (world_hit(sc, ray, t_min, INF, cx, rec)) {                                     // :25
        const Material& material = sc.materials[rec.mat_handle - 1];                   // :26
        V3 emitted = material_emitted(sc, material, rec.u, rec.v, rec.point, cx.cnt);  // :28
        Ray scattered; V3 attenuation;
        if (cx.cnt) cx.cnt->accum++;
        if (material_scatter(sc, material, ray, rec, cx, scattered, attenuation))      // :30-31
            return emitted + attenuation * ray_color(sc, scattered, background, depth - 1, max_depth, t_min, cx, segments);
        return emitted;                                                                // :32-33
    }
    return background;                                                                 // :37
}

V3 trace_pixel_sample(const Scene& sc, const rtw_camera& cam, const rtw_render_params& prm, int x, int y, int s,
                      Counters* cnt, int* segments) {
    Rng g; g.philox = true;
    g.key[0] = (uint32_t)prm.seed; g.key[1] = (uint32_t)(prm.seed >> 32);
    g.pixel = (uint32_t)(y * prm.width + x); g.sample = (uint32_t)s;
    g.set_bounce(0);
    Ctx cx{&g, cnt};
    double u = ((double)x + g.next()) / ((double)prm.width - 1.0);                     // :517
    double v = ((double)y + g.next()) / ((double)prm.height - 1.0);                    // :518
    Ray r = camera_get_ray(cam, u, v, g);                                              // :520
    if (cnt) cnt->paths++;
    V3 c = ray_color(sc, r, v3(prm.background), prm.max_depth, prm.max_depth, prm.t_min, cx, segments);  // :522
    if (cnt) cnt->draws += (uint64_t)g.total;
    return c;
}

thread_local std::string g_err;
int fail(int code, const char* msg) { g_err = msg; return code; }

}  // namespace

// =============================================================================================
// C ABI — same shape as include/rtw.h with the orc_ prefix
// =============================================================================================
extern "C" {

typedef struct orc_scene orc_scene;
#define SC(p) (reinterpret_cast<Scene*>(p))

typedef struct orc_counters {
    uint64_t paths, rays, aabb, sphere, sphere_accept, moving, rect, rect_accept, translate, rotate, medium;
    uint64_t scatter[5];
    uint64_t tex[4];
    uint64_t accum, draws;
} orc_counters;

const char* orc_last_error(void) { return g_err.c_str(); }

orc_scene* orc_scene_new(void) { return reinterpret_cast<orc_scene*>(new Scene()); }
void orc_scene_free(orc_scene* s) { delete SC(s); }
int orc_scene_set_media_deferred(orc_scene* s, int on) { SC(s)->media_deferred = on != 0; return 0; }
int orc_scene_set_build_seed(orc_scene* s, uint64_t seed) { SC(s)->build_rng = seed; return 0; }

int orc_tex_solid(orc_scene* s, const double rgb[3]) {
    Texture t; t.kind = TEX_SOLID; t.c0 = v3(rgb); SC(s)->textures.push_back(t); return (int)SC(s)->textures.size() - 1;
}
int orc_tex_checker(orc_scene* s, const double even[3], const double odd[3]) {
    Texture t; t.kind = TEX_CHECKER; t.c0 = v3(even); t.c1 = v3(odd); SC(s)->textures.push_back(t);
    return (int)SC(s)->textures.size() - 1;
}
int orc_tex_noise(orc_scene* s, const double* ranvec, const int32_t* px, const int32_t* py, const int32_t* pz, double scale) {
    Perlin* p = new Perlin();
    for (int i = 0; i < 256; ++i) {
        p->ranvec[i] = V3(ranvec[3 * i], ranvec[3 * i + 1], ranvec[3 * i + 2]);
        p->perm_x[i] = px[i]; p->perm_y[i] = py[i]; p->perm_z[i] = pz[i];
    }
    SC(s)->perlins.push_back(p);
    Texture t; t.kind = TEX_NOISE; t.perlin = p; t.scale = scale; SC(s)->textures.push_back(t);
    return (int)SC(s)->textures.size() - 1;
}
int orc_tex_image(orc_scene* s, int32_t w, int32_t h, int32_t bps, const uint8_t* data) {
    if (w <= 0 || h <= 0 || bps < 3 * w || !data) return fail(RTW_ERR_INVALID_ARG, "bad image");
    Texture t; t.kind = TEX_IMAGE; t.w = w; t.h = h; t.bps = bps; t.data.assign(data, data + (size_t)bps * h);
    SC(s)->textures.push_back(std::move(t));
    return (int)SC(s)->textures.size() - 1;
}
static int push_mat(orc_scene* s, const Material& m) { SC(s)->materials.push_back(m); return (int)SC(s)->materials.size(); }
static bool tex_ok(orc_scene* s, int tex) { return tex >= 0 && tex < (int)SC(s)->textures.size(); }
int orc_mat_lambertian(orc_scene* s, int tex) { if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad tex"); Material m; m.kind = MAT_LAMBERTIAN; m.tex = tex; return push_mat(s, m); }
int orc_mat_metal(orc_scene* s, const double albedo[3], double fuzz) { Material m; m.kind = MAT_METAL; m.albedo = v3(albedo); m.fuzz = fuzz; return push_mat(s, m); }
int orc_mat_dielectric(orc_scene* s, double ir) { Material m; m.kind = MAT_DIELECTRIC; m.ir = ir; return push_mat(s, m); }
int orc_mat_diffuse_light(orc_scene* s, int tex) { if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad tex"); Material m; m.kind = MAT_DIFFUSE_LIGHT; m.tex = tex; return push_mat(s, m); }
int orc_mat_isotropic(orc_scene* s, int tex) { if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad tex"); Material m; m.kind = MAT_ISOTROPIC; m.tex = tex; return push_mat(s, m); }

static int push_node(orc_scene* s, const Hittable& h) { SC(s)->nodes.push_back(h); return (int)SC(s)->nodes.size() - 1; }
static bool node_ok(orc_scene* s, int id) { return id >= 0 && id < (int)SC(s)->nodes.size(); }
static bool mat_ok(orc_scene* s, int m) { return m >= 1 && m <= (int)SC(s)->materials.size(); }

int orc_sphere(orc_scene* s, int mat, const double c[3], double r) {
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    Hittable h; h.kind = H_SPHERE; h.mat = mat; h.c0 = v3(c); h.radius = r; return push_node(s, h);
}
int orc_sphere_batch(orc_scene* s, int32_t n, const int32_t* mats, const double* centers, const double* radii) {
    for (int i = 0; i < n; ++i) {
        int id = orc_sphere(s, mats[i], centers + 3 * i, radii[i]);
        if (id < 0) return id;
        SC(s)->world.push_back(id);
    }
    return 0;
}
int orc_moving_sphere(orc_scene* s, int mat, const double c0[3], const double c1[3], double t0, double t1, double r) {
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    Hittable h; h.kind = H_MOVING_SPHERE; h.mat = mat; h.c0 = v3(c0); h.c1 = v3(c1); h.time0 = t0; h.time1 = t1; h.radius = r;
    return push_node(s, h);
}
static int push_rect(orc_scene* s, int kind, int mat, double a0, double a1, double b0, double b1, double k) {
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    Hittable h; h.kind = kind; h.mat = mat; h.a0 = a0; h.a1 = a1; h.b0 = b0; h.b1 = b1; h.k = k; return push_node(s, h);
}
int orc_xy_rect(orc_scene* s, int mat, double x0, double x1, double y0, double y1, double k) { return push_rect(s, H_XY, mat, x0, x1, y0, y1, k); }
int orc_xz_rect(orc_scene* s, int mat, double x0, double x1, double z0, double z1, double k) { return push_rect(s, H_XZ, mat, x0, x1, z0, z1, k); }
int orc_yz_rect(orc_scene* s, int mat, double y0, double y1, double z0, double z1, double k) { return push_rect(s, H_YZ, mat, y0, y1, z0, z1, k); }
int orc_box(orc_scene* s, const double mn[3], const double mx[3], int mat) {                       // new_box hittable.rs:132-145
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    Hittable h; h.kind = H_BOX; h.mat = mat; h.bmin = v3(mn); h.bmax = v3(mx);
    h.sides.push_back(orc_xy_rect(s, mat, mn[0], mx[0], mn[1], mx[1], mx[2]));
    h.sides.push_back(orc_xy_rect(s, mat, mn[0], mx[0], mn[1], mx[1], mn[2]));
    h.sides.push_back(orc_xz_rect(s, mat, mn[0], mx[0], mn[2], mx[2], mx[1]));
    h.sides.push_back(orc_xz_rect(s, mat, mn[0], mx[0], mn[2], mx[2], mn[1]));
    h.sides.push_back(orc_yz_rect(s, mat, mn[1], mx[1], mn[2], mx[2], mx[0]));
    h.sides.push_back(orc_yz_rect(s, mat, mn[1], mx[1], mn[2], mx[2], mn[0]));
    return push_node(s, h);
}
int orc_translate(orc_scene* s, int child, const double offset[3]) {
    if (!node_ok(s, child)) return fail(RTW_ERR_INVALID_ARG, "bad child");
    Hittable h; h.kind = H_TRANSLATE; h.child = child; h.offset = v3(offset); return push_node(s, h);
}
int orc_rotate_y(orc_scene* s, double angle, int child) {                                          // new_rotate_y :147-199
    if (!node_ok(s, child)) return fail(RTW_ERR_INVALID_ARG, "bad child");
    Hittable h; h.kind = H_ROTATE_Y; h.child = child;
    double radians = degrees_to_radians(angle);
    h.sin_theta = std::sin(radians); h.cos_theta = std::cos(radians);
    AABB bb;
    h.has_box = bounding_box(*SC(s), child, 0.0, 1.0, bb);
    if (!h.has_box) bb = AABB{V3(0, 0, 0), V3(0, 0, 0)};
    double mn[3] = {INF, INF, INF}, mx[3] = {-INF, -INF, -INF};
    for (int i = 0; i < 2; ++i) for (int j = 0; j < 2; ++j) for (int k = 0; k < 2; ++k) {
        double x = i * bb.maximum.x + (1.0 - i) * bb.minimum.x;
        double y = j * bb.maximum.y + (1.0 - j) * bb.minimum.y;
        double z = k * bb.maximum.z + (1.0 - k) * bb.minimum.z;
        double newx = h.cos_theta * x + h.sin_theta * z;
        double newz = -h.sin_theta * x + h.cos_theta * z;
        double tester[3] = {newx, y, newz};
        for (int c = 0; c < 3; ++c) { mn[c] = std::fmin(mn[c], tester[c]); mx[c] = std::fmax(mx[c], tester[c]); }
    }
    h.bbox = AABB{V3(mn[0], mn[1], mn[2]), V3(mx[0], mx[1], mx[2])};
    return push_node(s, h);
}
int orc_constant_medium(orc_scene* s, int child, double density, int phase_mat) {                  // :201-207
    if (!node_ok(s, child)) return fail(RTW_ERR_INVALID_ARG, "bad child");
    if (!mat_ok(s, phase_mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    Hittable h; h.kind = H_MEDIUM; h.child = child; h.mat = phase_mat; h.neg_inv_density = -1.0 / density;
    return push_node(s, h);
}
int orc_bvh_node(orc_scene* s, const int32_t* children, int32_t n, double t0, double t1) {
    if (n <= 0 || !children) return fail(RTW_ERR_INVALID_ARG, "empty bvh");
    std::vector<int> objs(children, children + n);
    for (int id : objs) if (!node_ok(s, id)) return fail(RTW_ERR_INVALID_ARG, "bad child");
    return build_bvh(*SC(s), objs, 0, objs.size(), t0, t1);
}
int orc_world_push(orc_scene* s, int id) {
    if (!node_ok(s, id)) return fail(RTW_ERR_INVALID_ARG, "bad hittable");
    SC(s)->world.push_back(id); return 0;
}
int orc_world_clear(orc_scene* s) { SC(s)->world.clear(); return 0; }

int orc_camera_new(const double look_from[3], const double look_at[3], const double vup[3], double vfov,
                   double aspect, double aperture, double focus_dist, double t0, double t1, rtw_camera* out) {
    camera_new(v3(look_from), v3(look_at), v3(vup), vfov, aspect, aperture, focus_dist, t0, t1, out);
    return 0;
}

int orc_bounding_box(orc_scene* s, int id, double t0, double t1, double out_min[3], double out_max[3]) {
    AABB b;
    if (!node_ok(s, id)) return fail(RTW_ERR_INVALID_ARG, "bad hittable");
    if (!bounding_box(*SC(s), id, t0, t1, b)) return 1;
    put(out_min, b.minimum); put(out_max, b.maximum);
    return 0;
}

int orc_sphere_uv(int32_t n, const double* p, double* out_u, double* out_v) {
    for (int i = 0; i < n; ++i) sphere_uv(V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), out_u[i], out_v[i]);
    return 0;
}
int orc_reflectance(int32_t n, const double* cosine, const double* ref_idx, double* out) {
    for (int i = 0; i < n; ++i) out[i] = reflectance(cosine[i], ref_idx[i]);
    return 0;
}
int orc_refract(const double uv[3], const double nrm[3], double eta, double out[3]) { put(out, refract(v3(uv), v3(nrm), eta)); return 0; }
int orc_reflect(const double v[3], const double nrm[3], double out[3]) { put(out, reflect(v3(v), v3(nrm))); return 0; }
int orc_perlin_noise(orc_scene* s, int tex, int32_t n, const double* p, double* out_noise, double* out_turb7) {
    if (!tex_ok(s, tex) || SC(s)->textures[tex].kind != TEX_NOISE) return fail(RTW_ERR_INVALID_ARG, "not a noise texture");
    const Perlin& pl = *SC(s)->textures[tex].perlin;
    for (int i = 0; i < n; ++i) {
        V3 q(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        out_noise[i] = perlin_noise(pl, q); out_turb7[i] = perlin_turb(pl, q, 7);
    }
    return 0;
}

int orc_test_philox(int32_t n, const uint32_t* counter, const uint32_t* key, uint32_t* out) {
    for (int i = 0; i < n; ++i) philox4x32_10(counter + 4 * i, key + 2 * i, out + 4 * i);
    return 0;
}

int orc_test_get_ray(const rtw_camera* cam, int32_t n, const double* s, const double* t, const double* xi,
                     int32_t stride, double* out_origin, double* out_dir, double* out_time, int32_t* out_ndraw) {
    for (int i = 0; i < n; ++i) {
        Rng g; g.stream = xi + (size_t)i * stride; g.stream_len = stride;
        Ray r = camera_get_ray(*cam, s[i], t[i], g);
        put(out_origin + 3 * i, r.origin); put(out_dir + 3 * i, r.direction); out_time[i] = r.time;
        out_ndraw[i] = g.exhausted ? -1 : g.draw;
    }
    return 0;
}

int orc_test_hit(orc_scene* s, int32_t target, int32_t n, const double* origin, const double* dir,
                 const double* time, double t_min, double t_max, const double* xi, int32_t stride,
                 int32_t* out_hit, double* out_t, double* out_p, double* out_normal, int32_t* out_front,
                 double* out_u, double* out_v, int32_t* out_mat, int32_t* out_ndraw) {
    if (target >= 0 && !node_ok(s, target)) return fail(RTW_ERR_INVALID_ARG, "bad target");
    for (int i = 0; i < n; ++i) {
        Rng g; g.stream = xi ? xi + (size_t)i * stride : nullptr; g.stream_len = xi ? stride : 0;
        Ctx cx{&g, nullptr};
        Ray r{V3(origin[3 * i], origin[3 * i + 1], origin[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), time ? time[i] : 0.0};
        HitRecord rec;
        bool hit = target < 0 ? world_hit(*SC(s), r, t_min, t_max, cx, rec) : hittable_hit(*SC(s), target, r, t_min, t_max, cx, rec);
        out_hit[i] = hit ? 1 : 0;
        if (!hit) rec = HitRecord();
        out_t[i] = rec.t; put(out_p + 3 * i, rec.point); put(out_normal + 3 * i, rec.normal);
        out_front[i] = rec.front_face ? 1 : 0; out_u[i] = rec.u; out_v[i] = rec.v; out_mat[i] = rec.mat_handle;
        out_ndraw[i] = g.exhausted ? -1 : g.draw;
    }
    return 0;
}

int orc_test_aabb(int32_t n, const double* bmin, const double* bmax, const double* origin, const double* dir,
                  double t_min, double t_max, int32_t* out_hit) {
    for (int i = 0; i < n; ++i) {
        AABB b{V3(bmin[3 * i], bmin[3 * i + 1], bmin[3 * i + 2]), V3(bmax[3 * i], bmax[3 * i + 1], bmax[3 * i + 2])};
        Ray r{V3(origin[3 * i], origin[3 * i + 1], origin[3 * i + 2]), V3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]), 0.0};
        out_hit[i] = aabb_hit(b, r, t_min, t_max) ? 1 : 0;
    }
    return 0;
}

int orc_test_scatter(orc_scene* s, int32_t mat, int32_t n, const double* ro, const double* rd, const double* rt,
                     const double* p, const double* normal, const int32_t* front, const double* u, const double* v,
                     const double* xi, int32_t stride, int32_t* out_scattered, double* out_origin, double* out_dir,
                     double* out_time, double* out_att, double* out_emitted, int32_t* out_ndraw) {
    if (!mat_ok(s, mat)) return fail(RTW_ERR_INVALID_ARG, "bad material handle");
    const Material& m = SC(s)->materials[mat - 1];
    for (int i = 0; i < n; ++i) {
        Rng g; g.stream = xi + (size_t)i * stride; g.stream_len = stride;
        Ctx cx{&g, nullptr};
        Ray r{V3(ro[3 * i], ro[3 * i + 1], ro[3 * i + 2]), V3(rd[3 * i], rd[3 * i + 1], rd[3 * i + 2]), rt ? rt[i] : 0.0};
        HitRecord rec; rec.point = V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]);
        rec.normal = V3(normal[3 * i], normal[3 * i + 1], normal[3 * i + 2]);
        rec.front_face = front[i] != 0; rec.u = u ? u[i] : 0.0; rec.v = v ? v[i] : 0.0; rec.mat_handle = mat;
        V3 em = material_emitted(*SC(s), m, rec.u, rec.v, rec.point, nullptr);
        Ray sc_ray; V3 att(0, 0, 0);
        bool ok = material_scatter(*SC(s), m, r, rec, cx, sc_ray, att);
        out_scattered[i] = ok ? 1 : 0;
        if (!ok) { sc_ray = Ray(); att = V3(0, 0, 0); }
        put(out_origin + 3 * i, sc_ray.origin); put(out_dir + 3 * i, sc_ray.direction); out_time[i] = sc_ray.time;
        put(out_att + 3 * i, att); put(out_emitted + 3 * i, em);
        out_ndraw[i] = g.exhausted ? -1 : g.draw;
    }
    return 0;
}

int orc_test_texture(orc_scene* s, int32_t tex, int32_t n, const double* u, const double* v, const double* p, double* out_rgb) {
    if (!tex_ok(s, tex)) return fail(RTW_ERR_INVALID_ARG, "bad tex");
    for (int i = 0; i < n; ++i)
        put(out_rgb + 3 * i, texture_value(SC(s)->textures[tex], u ? u[i] : 0.0, v ? v[i] : 0.0, V3(p[3 * i], p[3 * i + 1], p[3 * i + 2]), nullptr));
    return 0;
}

int orc_trace_paths(orc_scene* s, const rtw_camera* cam, const rtw_render_params* prm, int32_t n, const int32_t* px,
                    const int32_t* py, const int32_t* sample, double* out_rgb, int32_t* out_segments) {
    for (int i = 0; i < n; ++i) {
        int seg = 0;
        V3 c = trace_pixel_sample(*SC(s), *cam, *prm, px[i], py[i], sample[i], nullptr, &seg);
        put(out_rgb + 3 * i, c);
        if (out_segments) out_segments[i] = seg;
    }
    return 0;
}

// The pixel-sample loop of src/main.rs:507-540, tile-split over n_threads std::threads (rows from an atomic
// counter) instead of the reference's sample-split + mutex reduce (:516, :542-547).  out_sum: H x W x 3 f64,
// row 0 = top; out_sumsq (optional): per-pixel-channel sum of squares for the 3-sigma test.
int orc_render(orc_scene* s, const rtw_camera* cam, const rtw_render_params* prm, int32_t n_threads,
               double* out_sum, double* out_sumsq, orc_counters* out_counters, double* out_seconds) {
    const Scene& sc = *SC(s);
    const int W = prm->width, H = prm->height, spp = prm->spp;
    if (W <= 1 || H <= 1 || spp <= 0) return fail(RTW_ERR_INVALID_ARG, "bad render params");
    if (n_threads <= 0) n_threads = (int)std::thread::hardware_concurrency();
    if (n_threads <= 0) n_threads = 1;
    std::atomic<int> next_row(0);
    std::vector<Counters> cnts(n_threads);
    auto t0 = std::chrono::steady_clock::now();
    auto worker = [&](int tid) {
        Counters* cnt = out_counters ? &cnts[tid] : nullptr;
        for (;;) {
            int y = next_row.fetch_add(1);
            if (y >= H) break;
            for (int x = 0; x < W; ++x) {
                V3 pixel_color(0, 0, 0), sq(0, 0, 0);
                for (int sidx = 0; sidx < spp; ++sidx) {
                    V3 c = trace_pixel_sample(sc, *cam, *prm, x, y, sidx, cnt, nullptr);
                    pixel_color = pixel_color + c;
                    sq = sq + c * c;
                }
                size_t o = ((size_t)(H - 1 - y) * W + x) * 3;
                put(out_sum + o, pixel_color);
                if (out_sumsq) put(out_sumsq + o, sq);
            }
        }
    };
    std::vector<std::thread> th;
    for (int i = 1; i < n_threads; ++i) th.emplace_back(worker, i);
    worker(0);
    for (auto& t : th) t.join();
    auto t1 = std::chrono::steady_clock::now();
    if (out_seconds) *out_seconds = std::chrono::duration<double>(t1 - t0).count();
    if (out_counters) {
        Counters tot; for (auto& c : cnts) tot.add(c);
        out_counters->paths = tot.paths; out_counters->rays = tot.rays; out_counters->aabb = tot.aabb;
        out_counters->sphere = tot.sphere; out_counters->sphere_accept = tot.sphere_accept; out_counters->moving = tot.moving;
        out_counters->rect = tot.rect; out_counters->rect_accept = tot.rect_accept; out_counters->translate = tot.translate;
        out_counters->rotate = tot.rotate; out_counters->medium = tot.medium;
        for (int i = 0; i < 5; ++i) out_counters->scatter[i] = tot.scatter[i];
        for (int i = 0; i < 4; ++i) out_counters->tex[i] = tot.tex[i];
        out_counters->accum = tot.accum; out_counters->draws = tot.draws;
    }
    return 0;
}

int orc_write_color(const double* rgb_sum, int32_t n_pixels, int32_t spp, uint8_t* out_rgb8) {      // math.rs:119-132
    for (int i = 0; i < 3 * n_pixels; ++i) out_rgb8[i] = (uint8_t)to_byte(rgb_sum[i], spp);
    return 0;
}

int orc_hardware_threads(void) { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
