// scenes.hpp — the reference's scene functions (src/main.rs:52-289) and scene table (src/main.rs:316-459) written
// against the mirror types of rtw.hpp, with the seeded HostRng consumed in the reference's draw order.  These are
// INPUT generators for the example driver and the tests (same streams as scenes.py), not part of the hot path.
#ifndef RTW_HOST_SCENES_HPP
#define RTW_HOST_SCENES_HPP

#include <string>
#include <vector>

#include "rtw.hpp"

namespace rtw_host {

struct SceneDesc {               // the `Scene` struct of src/main.rs:296-305 (+ explicit height)
    World world;
    Point3 look_from, look_at;
    double vfov = 20;
    Color background;
    int image_width = 400, image_height = 225, samples_per_pixel = 100;
};

inline Texture earth_texture(const std::vector<uint8_t>* texels) {
    if (texels && texels->size() == 1024u * 512u * 3u) return Texture::Image(1024, 512, 3 * 1024, *texels);
    std::vector<uint8_t> d(1024u * 512u * 3u);          // procedural stand-in (same as scenes.py's fallback)
    for (unsigned y = 0; y < 512; ++y)
        for (unsigned x = 0; x < 1024; ++x) {
            uint8_t* p = &d[(y * 1024u + x) * 3u];
            p[0] = (uint8_t)(x * 255u / 1023u); p[1] = (uint8_t)(y * 255u / 511u); p[2] = (uint8_t)((x ^ y) & 255u);
        }
    return Texture::Image(1024, 512, 3 * 1024, std::move(d));
}

inline World random_scene(HostRng& g) {                                               // src/main.rs:245-289
    World world;
    auto ground = world.register_material(Material::Lambertian(Texture::Checker(Color(0.2, 0.5, 0.5), Color(0.9, 0.9, 0.9))));
    world.hittables.push_back(Hittable::Sphere(ground, Point3(0.0, -1000.0, 0.0), 1000.0));
    for (int a = -11; a < 11; ++a)
        for (int b = -11; b < 11; ++b) {
            double choose_mat = g.random_double();
            double cx = a + 0.9 * g.random_double();
            double cz = b + 0.9 * g.random_double();
            Point3 center(cx, 0.2, cz);
            if ((center - Point3(4.0, 0.2, 0.0)).length() > 0.9) {
                if (choose_mat < 0.8) {
                    Color albedo = g.random();
                    auto m = world.register_material(Material::Lambertian(Texture::SolidColor(albedo)));
                    Point3 center2 = center + Vector3(0.0, g.random_double_range(0.0, 0.5), 0.0);
                    world.hittables.push_back(Hittable::MovingSphere(m, center, center2, 0.0, 1.0, 0.2));
                } else if (choose_mat < 0.95) {
                    Color albedo = g.random_range(0.5, 1.0);
                    double fuzz = g.random_double_range(0.0, 0.5);
                    auto m = world.register_material(Material::Metal(albedo, fuzz));
                    world.hittables.push_back(Hittable::Sphere(m, center, 0.2));
                } else {
                    auto m = world.register_material(Material::Dielectric(1.5));
                    world.hittables.push_back(Hittable::Sphere(m, center, 0.2));
                }
            }
        }
    world.hittables.push_back(Hittable::Sphere(world.register_material(Material::Dielectric(1.5)), Point3(0.0, 1.0, 0.0), 1.0));
    world.hittables.push_back(Hittable::Sphere(world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.4, 0.2, 0.1)))), Point3(-4.0, 1.0, 0.0), 1.0));
    world.hittables.push_back(Hittable::Sphere(world.register_material(Material::Metal(Color(0.7, 0.6, 0.5), 0.0)), Point3(4.0, 1.0, 0.0), 1.0));
    return world;
}

inline World two_spheres_scene() {                                                    // src/main.rs:52-63
    World world;
    auto m = world.register_material(Material::Lambertian(Texture::Checker(Color(0.2, 0.3, 0.1), Color(0.9, 0.9, 0.9))));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, -10.0, 0.0), 10.0));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, 10.0, 0.0), 10.0));
    return world;
}

inline World two_perlin_spheres_scene(HostRng& g) {                                   // src/main.rs:65-76
    World world;
    auto m = world.register_material(Material::Lambertian(Texture::Noise(Perlin::create(g), 4.0)));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, -1000.0, 0.0), 1000.0));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, 2.0, 0.0), 2.0));
    return world;
}

inline World earth_scene(const std::vector<uint8_t>* texels) {                        // src/main.rs:78-89
    World world;
    auto m = world.register_material(Material::Lambertian(earth_texture(texels)));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, 0.0, 0.0), 2.0));
    return world;
}

inline World simple_light_scene(HostRng& g) {                                         // src/main.rs:91-105
    World world;
    auto m = world.register_material(Material::Lambertian(Texture::Noise(Perlin::create(g), 4.0)));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, -1000.0, 0.0), 1000.0));
    world.hittables.push_back(Hittable::Sphere(m, Point3(0.0, 2.0, 0.0), 2.0));
    auto light = world.register_material(Material::DiffuseLight(Texture::SolidColor(Color(4.0, 4.0, 4.0))));
    world.hittables.push_back(Hittable::XYRect(light, 3.0, 5.0, 1.0, 3.0, -2.0));
    return world;
}

inline MaterialHandle cornell_walls(World& world, Color light_rgb, double lx0, double lx1, double lz0, double lz1) {
    auto red = world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.65, 0.05, 0.05))));
    auto white = world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.73, 0.73, 0.73))));
    auto green = world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.12, 0.45, 0.15))));
    auto light = world.register_material(Material::DiffuseLight(Texture::SolidColor(light_rgb)));
    world.hittables.push_back(Hittable::YZRect(green, 0.0, 555.0, 0.0, 555.0, 555.0));
    world.hittables.push_back(Hittable::YZRect(red, 0.0, 555.0, 0.0, 555.0, 0.0));
    world.hittables.push_back(Hittable::XZRect(light, lx0, lx1, lz0, lz1, 554.0));
    world.hittables.push_back(Hittable::XZRect(white, 0.0, 555.0, 0.0, 555.0, 0.0));
    world.hittables.push_back(Hittable::XZRect(white, 0.0, 555.0, 0.0, 555.0, 555.0));
    world.hittables.push_back(Hittable::XYRect(white, 0.0, 555.0, 0.0, 555.0, 555.0));
    return white;
}

inline World cornell_box_scene() {                                                    // src/main.rs:107-136
    World world;
    auto white = cornell_walls(world, Color(15.0, 15.0, 15.0), 213.0, 343.0, 227.0, 332.0);
    world.hittables.push_back(Hittable::Translate(Vector3(265.0, 0.0, 295.0),
        Hittable::new_rotate_y(15.0, Hittable::new_box(Point3(0.0, 0.0, 0.0), Point3(165.0, 330.0, 165.0), white))));
    world.hittables.push_back(Hittable::Translate(Vector3(130.0, 0.0, 65.0),
        Hittable::new_rotate_y(-18.0, Hittable::new_box(Point3(0.0, 0.0, 0.0), Point3(165.0, 165.0, 165.0), white))));
    return world;
}

inline World cornell_box_smoke_scene() {                                              // src/main.rs:138-171
    World world;
    auto white = cornell_walls(world, Color(7.0, 7.0, 7.0), 113.0, 443.0, 127.0, 432.0);
    auto p1 = world.register_material(Material::Isotropic(Texture::SolidColor(Color(0.0, 0.0, 0.0))));
    world.hittables.push_back(Hittable::new_constant_medium(Hittable::Translate(Vector3(265.0, 0.0, 295.0),
        Hittable::new_rotate_y(15.0, Hittable::new_box(Point3(0.0, 0.0, 0.0), Point3(165.0, 330.0, 165.0), white))), 0.01, p1));
    auto p2 = world.register_material(Material::Isotropic(Texture::SolidColor(Color(1.0, 1.0, 1.0))));
    world.hittables.push_back(Hittable::new_constant_medium(Hittable::Translate(Vector3(130.0, 0.0, 65.0),
        Hittable::new_rotate_y(-18.0, Hittable::new_box(Point3(0.0, 0.0, 0.0), Point3(165.0, 165.0, 165.0), white))), 0.01, p2));
    return world;
}

inline World final_scene(HostRng& g, const std::vector<uint8_t>* texels) {            // src/main.rs:173-243
    World world;
    std::vector<Hittable> boxes1;
    auto ground = world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.48, 0.83, 0.53))));
    for (int i = 0; i < 20; ++i)
        for (int j = 0; j < 20; ++j) {
            double w = 100.0, x0 = -1000.0 + i * w, z0 = -1000.0 + j * w, y1 = g.random_double_range(1.0, 101.0);
            boxes1.push_back(Hittable::new_box(Point3(x0, 0.0, z0), Point3(x0 + w, y1, z0 + w), ground));
        }
    world.hittables.push_back(Hittable::new_bvh_node(boxes1, 0, boxes1.size(), 0.0, 1.0));
    auto light = world.register_material(Material::DiffuseLight(Texture::SolidColor(Color(7.0, 7.0, 7.0))));
    world.hittables.push_back(Hittable::XZRect(light, 123.0, 423.0, 147.0, 412.0, 554.0));
    Point3 c1(400.0, 400.0, 200.0), c2(430.0, 400.0, 200.0);
    world.hittables.push_back(Hittable::MovingSphere(world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.7, 0.3, 0.1)))), c1, c2, 0.0, 1.0, 50.0));
    auto dielectric = world.register_material(Material::Dielectric(1.5));
    world.hittables.push_back(Hittable::Sphere(dielectric, Point3(260.0, 150.0, 45.0), 50.0));
    world.hittables.push_back(Hittable::Sphere(world.register_material(Material::Metal(Color(0.8, 0.8, 0.9), 1.0)), Point3(0.0, 150.0, 145.0), 50.0));
    Hittable boundary = Hittable::Sphere(dielectric, Point3(360.0, 150.0, 145.0), 70.0);
    world.hittables.push_back(boundary);
    world.hittables.push_back(Hittable::new_constant_medium(boundary, 0.2, world.register_material(Material::Isotropic(Texture::SolidColor(Color(0.2, 0.4, 0.9))))));
    boundary = Hittable::Sphere(dielectric, Point3(0.0, 0.0, 0.0), 5000.0);
    world.hittables.push_back(Hittable::new_constant_medium(boundary, 0.0001, world.register_material(Material::Isotropic(Texture::SolidColor(Color(1.0, 1.0, 1.0))))));
    world.hittables.push_back(Hittable::Sphere(world.register_material(Material::Lambertian(earth_texture(texels))), Point3(400.0, 200.0, 400.0), 100.0));
    world.hittables.push_back(Hittable::Sphere(world.register_material(Material::Lambertian(Texture::Noise(Perlin::create(g), 0.1))), Point3(220.0, 280.0, 300.0), 80.0));
    std::vector<Hittable> boxes2;
    auto white = world.register_material(Material::Lambertian(Texture::SolidColor(Color(0.73, 0.73, 0.73))));
    for (int j = 0; j < 1000; ++j) boxes2.push_back(Hittable::Sphere(white, g.random_range(0.0, 165.0), 10.0));
    world.hittables.push_back(Hittable::Translate(Vector3(-100.0, 270.0, 395.0),
        Hittable::new_rotate_y(15.0, Hittable::new_bvh_node(boxes2, 0, boxes2.size(), 0.0, 1.0))));
    return world;
}

// `match` of src/main.rs:314-464: scene id -> world + camera row.  Heights are explicit (HEAD's width*aspect is a bug).
inline SceneDesc make_scene(int id, uint64_t seed, const std::vector<uint8_t>* earth_texels) {
    HostRng g(seed);
    SceneDesc s;
    s.look_from = Point3(13.0, 2.0, 3.0); s.look_at = Point3(0.0, 0.0, 0.0); s.vfov = 20.0; s.background = Color(0.7, 0.8, 1.0);
    switch (id) {
    case 0: s.world = random_scene(g); s.image_width = 1200; s.image_height = 800; s.samples_per_pixel = 500; break;
    case 1: s.world = two_spheres_scene(); s.image_width = 800; s.image_height = 450; s.samples_per_pixel = 200; break;
    case 2: s.world = two_perlin_spheres_scene(g); s.image_width = 800; s.image_height = 450; s.samples_per_pixel = 200; break;
    case 3: s.world = earth_scene(earth_texels); s.image_width = 800; s.image_height = 450; s.samples_per_pixel = 200; break;
    case 4: s.world = simple_light_scene(g); s.look_from = Point3(26.0, 3.0, 6.0); s.look_at = Point3(0.0, 2.0, 0.0);
            s.background = Color(0, 0, 0); s.image_width = 600; s.image_height = 600; s.samples_per_pixel = 1000; break;
    case 5: case 6:
        s.world = id == 5 ? cornell_box_scene() : cornell_box_smoke_scene();
        s.look_from = Point3(278.0, 278.0, -800.0); s.look_at = Point3(278.0, 278.0, 0.0); s.vfov = 40.0;
        s.background = Color(0, 0, 0); s.image_width = 600; s.image_height = 600; s.samples_per_pixel = 1000; break;
    case 7:
        s.world = final_scene(g, earth_texels);
        s.look_from = Point3(478.0, 278.0, -600.0); s.look_at = Point3(278.0, 278.0, 0.0); s.vfov = 40.0;
        s.background = Color(0, 0, 0); s.image_width = 800; s.image_height = 800; s.samples_per_pixel = 10000; break;
    default: throw std::runtime_error("Unsupported scene selected");              // src/main.rs:461-463 (panic -> exception)
    }
    return s;
}

}  // namespace rtw_host
#endif
