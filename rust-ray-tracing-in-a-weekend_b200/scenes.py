"""Scene library — the bench/test INPUTS.

Restates the reference's scene functions (src/main.rs:52-289) and scene table (src/main.rs:316-459)
on top of the constructor API in api.py, with a seeded host RNG (SplitMix64) consumed in the
reference's draw order, so the oracle and the CUDA library are fed byte-identical scenes.
Not on the hot path: this is host-side scene description only.
"""
import math
import os
from dataclasses import dataclass, field

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EARTHMAP = os.path.join(_ROOT, "tests", "golden", "earthmap.jpg")

_M64 = (1 << 64) - 1


class HostRng:
    """SplitMix64 -> U[0,1) doubles; stands in for rand::thread_rng on the scene-building side."""

    def __init__(self, seed=1):
        self.s = seed & _M64

    def random_double(self):                                   # src/math.rs:268-271
        self.s = (self.s + 0x9E3779B97F4A7C15) & _M64
        z = self.s
        z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & _M64
        z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & _M64
        z ^= z >> 31
        return (z >> 11) * (1.0 / 9007199254740992.0)

    def range(self, a, b):                                     # src/math.rs:273-276
        return a + (b - a) * self.random_double()

    def int_range(self, a, b):                                 # src/math.rs:278-280
        return int(self.range(float(a), float(b + 1)))

    def color(self):                                           # Vector3::random src/math.rs:35-41
        return (self.random_double(), self.random_double(), self.random_double())

    def color_range(self, a, b):                               # Vector3::random_range src/math.rs:43-49
        return (self.range(a, b), self.range(a, b), self.range(a, b))


def perlin_tables(rng):
    """Perlin::new (src/perlin.rs:13-30) incl. the non-permutation `permute` (src/perlin.rs:122-129)."""
    ranvec = np.zeros((256, 3))
    for i in range(256):
        v = np.array(rng.color_range(-1.0, 1.0))
        ranvec[i] = (1.0 / math.sqrt(float(v @ v))) * v          # normalize = (1/len)*v, src/math.rs:102-104,260-266

    def gen_perm():
        p = list(range(256))
        for i in range(255, -1, -1):
            target = min(rng.int_range(0, i), 255)
            tmp = p[i]
            p[i] = target                                        # sic: index, not p[target] (src/perlin.rs:126)
            p[target] = tmp
        return np.array(p, np.int32)

    return ranvec, gen_perm(), gen_perm(), gen_perm()


_earth_cache = {}


def earth_texels():
    """RGB8 texels of textures/earthmap.jpg (copied to tests/golden/).  stb_image is replaced by Pillow;
    both libraries get the same decoded bytes (SURVEY §9 decoder note).  A missing file or decoder is an error:
    a bench / test input must never silently change the workload."""
    if "e" not in _earth_cache:
        try:
            from PIL import Image
        except ImportError as e:
            raise RuntimeError("scenes.earth_texels: Pillow is needed to decode tests/golden/earthmap.jpg") from e
        if not os.path.exists(EARTHMAP):
            raise FileNotFoundError(f"scenes.earth_texels: {EARTHMAP} is missing (the earth / final_scene image texture)")
        _earth_cache["e"] = np.asarray(Image.open(EARTHMAP).convert("RGB"), dtype=np.uint8).copy()
    return _earth_cache["e"]


@dataclass
class SceneSpec:
    """One row of the reference scene table (src/main.rs:316-459) + the config's W/H/spp."""
    name: str
    look_from: tuple
    look_at: tuple
    vfov: float
    background: tuple
    width: int
    height: int
    spp: int
    aperture: float = 0.1            # src/main.rs:469
    focus_dist: float = 10.0         # src/main.rs:312
    vup: tuple = (0.0, 1.0, 0.0)     # src/main.rs:311
    max_depth: int = 50              # src/main.rs:310
    info: dict = field(default_factory=dict)

    def camera(self, lib, width=None, height=None):
        w = width or self.width
        h = height or self.height
        return lib.camera_new(self.look_from, self.look_at, self.vup, self.vfov, w / h, self.aperture,
                              self.focus_dist, 0.0, 1.0)


class _World:
    """Collects world.hittables; optionally wraps them in one reference-style BvhNode (the oracle's
    'BVH model' used for the op-count roofline model)."""

    def __init__(self, sc, wrap_bvh=False):
        self.sc, self.wrap, self.items = sc, wrap_bvh, []

    def push(self, h):
        if self.wrap:
            self.items.append(h)
        else:
            self.sc.push(h)

    def done(self):
        if self.wrap and self.items:
            self.sc.push(self.sc.bvh_node(self.items, 0.0, 1.0))


def random_scene(sc, seed=1, wrap_bvh=False):                                       # src/main.rs:245-289
    rng = HostRng(seed)
    w = _World(sc, wrap_bvh)
    ground = sc.lambertian(sc.tex_checker((0.2, 0.5, 0.5), (0.9, 0.9, 0.9)))
    w.push(sc.sphere(ground, (0.0, -1000.0, 0.0), 1000.0))
    counts = dict(lambertian=0, metal=0, dielectric=0)
    for a in range(-11, 11):
        for b in range(-11, 11):
            choose_mat = rng.random_double()
            cx = a + 0.9 * rng.random_double()
            cz = b + 0.9 * rng.random_double()
            center = (cx, 0.2, cz)
            if math.sqrt((cx - 4.0) ** 2 + (0.2 - 0.2) ** 2 + (cz - 0.0) ** 2) > 0.9:
                if choose_mat < 0.8:
                    albedo = rng.color()
                    m = sc.lambertian(sc.tex_solid(albedo))
                    center2 = (cx, 0.2 + rng.range(0.0, 0.5), cz)
                    w.push(sc.moving_sphere(m, center, center2, 0.0, 1.0, 0.2))
                    counts["lambertian"] += 1
                elif choose_mat < 0.95:
                    albedo = rng.color_range(0.5, 1.0)
                    fuzz = rng.range(0.0, 0.5)
                    m = sc.metal(albedo, fuzz)
                    w.push(sc.sphere(m, center, 0.2))
                    counts["metal"] += 1
                else:
                    m = sc.dielectric(1.5)
                    w.push(sc.sphere(m, center, 0.2))
                    counts["dielectric"] += 1
    w.push(sc.sphere(sc.dielectric(1.5), (0.0, 1.0, 0.0), 1.0))
    w.push(sc.sphere(sc.lambertian(sc.tex_solid((0.4, 0.2, 0.1))), (-4.0, 1.0, 0.0), 1.0))
    w.push(sc.sphere(sc.metal((0.7, 0.6, 0.5), 0.0), (4.0, 1.0, 0.0), 1.0))
    w.done()
    return SceneSpec("random_scene", (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), 20.0, (0.7, 0.8, 1.0), 1200, 800, 500,
                     info=counts)


def two_spheres_scene(sc, seed=1, wrap_bvh=False):                                  # src/main.rs:52-63
    w = _World(sc, wrap_bvh)
    ground = sc.lambertian(sc.tex_checker((0.2, 0.3, 0.1), (0.9, 0.9, 0.9)))
    w.push(sc.sphere(ground, (0.0, -10.0, 0.0), 10.0))
    w.push(sc.sphere(ground, (0.0, 10.0, 0.0), 10.0))
    w.done()
    return SceneSpec("two_spheres", (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), 20.0, (0.7, 0.8, 1.0), 800, 450, 200)


def two_perlin_spheres_scene(sc, seed=1, wrap_bvh=False):                           # src/main.rs:65-76
    rng = HostRng(seed)
    w = _World(sc, wrap_bvh)
    ground = sc.lambertian(sc.tex_noise(*perlin_tables(rng), 4.0))
    w.push(sc.sphere(ground, (0.0, -1000.0, 0.0), 1000.0))
    w.push(sc.sphere(ground, (0.0, 2.0, 0.0), 2.0))
    w.done()
    return SceneSpec("two_perlin_spheres", (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), 20.0, (0.7, 0.8, 1.0), 800, 450, 200)


def earth_scene(sc, seed=1, wrap_bvh=False):                                        # src/main.rs:78-89
    w = _World(sc, wrap_bvh)
    earth = sc.lambertian(sc.tex_image(earth_texels()))
    w.push(sc.sphere(earth, (0.0, 0.0, 0.0), 2.0))
    w.done()
    return SceneSpec("earth", (13.0, 2.0, 3.0), (0.0, 0.0, 0.0), 20.0, (0.7, 0.8, 1.0), 800, 450, 200)


def simple_light_scene(sc, seed=1, wrap_bvh=False):                                 # src/main.rs:91-105
    rng = HostRng(seed)
    w = _World(sc, wrap_bvh)
    ground = sc.lambertian(sc.tex_noise(*perlin_tables(rng), 4.0))
    w.push(sc.sphere(ground, (0.0, -1000.0, 0.0), 1000.0))
    w.push(sc.sphere(ground, (0.0, 2.0, 0.0), 2.0))
    light = sc.diffuse_light(sc.tex_solid((4.0, 4.0, 4.0)))
    w.push(sc.xy_rect(light, 3.0, 5.0, 1.0, 3.0, -2.0))
    w.done()
    return SceneSpec("simple_light", (26.0, 3.0, 6.0), (0.0, 2.0, 0.0), 20.0, (0.0, 0.0, 0.0), 600, 600, 1000)


def _cornell_walls(sc, w, light_rgb, lx0, lx1, lz0, lz1):
    red = sc.lambertian(sc.tex_solid((0.65, 0.05, 0.05)))
    white = sc.lambertian(sc.tex_solid((0.73, 0.73, 0.73)))
    green = sc.lambertian(sc.tex_solid((0.12, 0.45, 0.15)))
    light = sc.diffuse_light(sc.tex_solid(light_rgb))
    w.push(sc.yz_rect(green, 0.0, 555.0, 0.0, 555.0, 555.0))
    w.push(sc.yz_rect(red, 0.0, 555.0, 0.0, 555.0, 0.0))
    w.push(sc.xz_rect(light, lx0, lx1, lz0, lz1, 554.0))
    w.push(sc.xz_rect(white, 0.0, 555.0, 0.0, 555.0, 0.0))
    w.push(sc.xz_rect(white, 0.0, 555.0, 0.0, 555.0, 555.0))
    w.push(sc.xy_rect(white, 0.0, 555.0, 0.0, 555.0, 555.0))
    return white


def cornell_box_scene(sc, seed=1, wrap_bvh=False):                                  # src/main.rs:107-136
    w = _World(sc, wrap_bvh)
    white = _cornell_walls(sc, w, (15.0, 15.0, 15.0), 213.0, 343.0, 227.0, 332.0)
    box1 = sc.box((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), white)
    box1 = sc.translate(sc.rotate_y(15.0, box1), (265.0, 0.0, 295.0))
    w.push(box1)
    box2 = sc.box((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), white)
    box2 = sc.translate(sc.rotate_y(-18.0, box2), (130.0, 0.0, 65.0))
    w.push(box2)
    w.done()
    return SceneSpec("cornell_box", (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), 40.0, (0.0, 0.0, 0.0), 600, 600, 1000)


def cornell_box_smoke_scene(sc, seed=1, wrap_bvh=False):                            # src/main.rs:138-171
    w = _World(sc, wrap_bvh)
    white = _cornell_walls(sc, w, (7.0, 7.0, 7.0), 113.0, 443.0, 127.0, 432.0)
    p1 = sc.isotropic(sc.tex_solid((0.0, 0.0, 0.0)))
    box1 = sc.box((0.0, 0.0, 0.0), (165.0, 330.0, 165.0), white)
    box1 = sc.translate(sc.rotate_y(15.0, box1), (265.0, 0.0, 295.0))
    w.push(sc.constant_medium(box1, 0.01, p1))
    p2 = sc.isotropic(sc.tex_solid((1.0, 1.0, 1.0)))
    box2 = sc.box((0.0, 0.0, 0.0), (165.0, 165.0, 165.0), white)
    box2 = sc.translate(sc.rotate_y(-18.0, box2), (130.0, 0.0, 65.0))
    w.push(sc.constant_medium(box2, 0.01, p2))
    w.done()
    return SceneSpec("cornell_box_smoke", (278.0, 278.0, -800.0), (278.0, 278.0, 0.0), 40.0, (0.0, 0.0, 0.0), 600, 600, 1000)


def final_scene(sc, seed=1, wrap_bvh=False):                                        # src/main.rs:173-243
    rng = HostRng(seed)
    w = _World(sc, wrap_bvh)
    ground = sc.lambertian(sc.tex_solid((0.48, 0.83, 0.53)))
    boxes1 = []
    for i in range(20):
        for j in range(20):
            ww = 100.0
            x0 = -1000.0 + i * ww
            z0 = -1000.0 + j * ww
            y1 = rng.range(1.0, 101.0)
            boxes1.append(sc.box((x0, 0.0, z0), (x0 + ww, y1, z0 + ww), ground))
    w.push(sc.bvh_node(boxes1, 0.0, 1.0))
    light = sc.diffuse_light(sc.tex_solid((7.0, 7.0, 7.0)))
    w.push(sc.xz_rect(light, 123.0, 423.0, 147.0, 412.0, 554.0))
    c1 = (400.0, 400.0, 200.0)
    c2 = (430.0, 400.0, 200.0)
    w.push(sc.moving_sphere(sc.lambertian(sc.tex_solid((0.7, 0.3, 0.1))), c1, c2, 0.0, 1.0, 50.0))
    dielectric = sc.dielectric(1.5)
    w.push(sc.sphere(dielectric, (260.0, 150.0, 45.0), 50.0))
    w.push(sc.sphere(sc.metal((0.8, 0.8, 0.9), 1.0), (0.0, 150.0, 145.0), 50.0))
    boundary = sc.sphere(dielectric, (360.0, 150.0, 145.0), 70.0)
    w.push(boundary)                                                                  # boundary.clone()
    w.push(sc.constant_medium(boundary, 0.2, sc.isotropic(sc.tex_solid((0.2, 0.4, 0.9)))))
    boundary = sc.sphere(dielectric, (0.0, 0.0, 0.0), 5000.0)
    w.push(sc.constant_medium(boundary, 0.0001, sc.isotropic(sc.tex_solid((1.0, 1.0, 1.0)))))
    emat = sc.lambertian(sc.tex_image(earth_texels()))
    w.push(sc.sphere(emat, (400.0, 200.0, 400.0), 100.0))
    pertext = sc.lambertian(sc.tex_noise(*perlin_tables(rng), 0.1))
    w.push(sc.sphere(pertext, (220.0, 280.0, 300.0), 80.0))
    white = sc.lambertian(sc.tex_solid((0.73, 0.73, 0.73)))
    boxes2 = [sc.sphere(white, rng.color_range(0.0, 165.0), 10.0) for _ in range(1000)]
    w.push(sc.translate(sc.rotate_y(15.0, sc.bvh_node(boxes2, 0.0, 1.0)), (-100.0, 270.0, 395.0)))
    w.done()
    return SceneSpec("final_scene", (478.0, 278.0, -600.0), (278.0, 278.0, 0.0), 40.0, (0.0, 0.0, 0.0), 800, 800, 10000)


def sweep_scene(sc, n_spheres, seed=None, wrap_bvh=False):
    """C5: N static spheres uniform in a cube of side 1000, 5 % volume fill, palette of 256
    materials with the C1 mix 80/15/5.  Generated with numpy (vectorised) — synthetic, no reference fn."""
    rs = np.random.RandomState(n_spheres if seed is None else seed)
    L = 1000.0
    r = L * (0.05 * 3.0 / (4.0 * math.pi * n_spheres)) ** (1.0 / 3.0)
    mats = []
    for i in range(256):
        t = i / 256.0
        if t < 0.8:
            mats.append(sc.lambertian(sc.tex_solid(tuple(rs.uniform(0.05, 0.95, 3)))))
        elif t < 0.95:
            mats.append(sc.metal(tuple(rs.uniform(0.5, 1.0, 3)), float(rs.uniform(0.0, 0.5))))
        else:
            mats.append(sc.dielectric(1.5))
    centers = rs.uniform(-L / 2, L / 2, (n_spheres, 3))
    which = rs.randint(0, 256, n_spheres)
    sc.sphere_batch(np.asarray(mats, np.int32)[which], centers, np.full(n_spheres, r))
    return SceneSpec(f"sweep_{n_spheres}", (2.2 * L, 0.6 * L, 1.1 * L), (0.0, 0.0, 0.0), 30.0, (0.7, 0.8, 1.0),
                     3840, 2160, 256, aperture=0.0, focus_dist=10.0)


SCENES = {
    "random_scene": random_scene,              # id 0 / config C1
    "two_spheres": two_spheres_scene,          # id 1 / C2a
    "two_perlin_spheres": two_perlin_spheres_scene,  # id 2 / C2b
    "earth": earth_scene,                      # id 3 / C2c
    "simple_light": simple_light_scene,        # id 4 / C3a
    "cornell_box": cornell_box_scene,          # id 5 / C3b
    "cornell_box_smoke": cornell_box_smoke_scene,    # id 6
    "final_scene": final_scene,                # id 7 / C4
}


def build(lib, name, seed=1, wrap_bvh=False):
    """-> (Scene, SceneSpec) built against `lib` (api.load_rtw(), or any other library exporting the same constructors)."""
    from .api import Scene
    sc = Scene(lib)
    spec = SCENES[name](sc, seed=seed, wrap_bvh=wrap_bvh)
    return sc, spec
