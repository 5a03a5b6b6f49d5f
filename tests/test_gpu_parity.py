"""Parity tests proper (-m gpu): the CUDA path, called through the C ABI, against the CPU oracle.

Bars (BASELINE.json north_star): kernel-level outputs within 1e-5 relative of the f64 reference arithmetic on
identical (f32-representable) inputs; discrete outputs equal except on rays the oracle itself flips under a 1e-6
relative input perturbation ("ill-conditioned"); full renders within the Monte-Carlo bound (3 sigma per pixel,
PSNR >= 40 dB at 4096 spp).  Integer/byte work (Philox words, write_color bytes, draw counts) is bit-exact.
"""
import math
import os
import zlib

import numpy as np
import pytest

from conftest import f32, q24, record

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REL = 1e-5


def rel_err(got, ref, scale):
    return np.abs(got - ref) / np.maximum(np.abs(ref), scale)


def build_both(pkg, gpu, orc, name, wrap=None, **kw):
    if wrap is None:
        wrap = name not in ("cornell_box_smoke", "final_scene")
    a, spec = pkg.scenes.build(gpu, name, **kw)
    b, _ = pkg.scenes.build(orc, name, wrap_bvh=wrap, **kw)
    b.set_media_deferred(True)
    a.commit(1, 0)
    return a, b, spec


def camera_rays(pkg, lib, spec, n, rs, W=96, H=64):
    cam = spec.camera(lib, W, H)
    g = lib.test_get_ray(cam, f32(rs.rand(n)), f32(rs.rand(n)), q24(rs, (n, 24)))
    return g


def stable(oracle_hit, o, d, tm, scale, **kw):
    """Mask of rays whose ORACLE answer (hit flag, and t within 1e-5) survives 1e-6-relative input perturbations."""
    base = oracle_hit(o, d, tm, **kw)
    ok = np.ones(len(o), bool)
    rs = np.random.RandomState(99)
    for _ in range(2):
        dd = d * (1 + 1e-6 * rs.uniform(-1, 1, d.shape))
        oo = o + 1e-6 * scale * rs.uniform(-1, 1, o.shape)
        h = oracle_hit(oo, dd, tm, **kw)
        ok &= h["hit"] == base["hit"]
        ok &= (h["mat"] == base["mat"])
        both = (h["hit"] == 1) & (base["hit"] == 1)
        ok &= ~both | (np.abs(h["t"] - base["t"]) <= 20 * REL * np.maximum(np.abs(base["t"]), 1e-3))
        ok &= ~both | (np.abs(h["normal"] - base["normal"]).max(1) < 1e-3)
    return base, ok


def compare_hits(ha, hb, ok, pos_scale, d_len, min_stable=0.99, check_uv=True, tag=""):
    record("compare_hits", tag=tag or os.environ.get("PYTEST_CURRENT_TEST", ""), stable=ok.mean(), flag_mismatch_all=np.mean(ha["hit"] != hb["hit"]))
    assert ok.mean() >= min_stable, ok.mean()
    assert np.array_equal(ha["hit"][ok], hb["hit"][ok])
    # over ALL rays (ill-conditioned included) the flags may differ only rarely
    assert np.mean(ha["hit"] != hb["hit"]) < 5e-3
    m = ok & (hb["hit"] == 1)
    if m.sum() == 0:          # e.g. rays leaving the only sphere of the earth scene
        assert np.array_equal(ha["ndraw"][ok], hb["ndraw"][ok])
        return
    assert np.array_equal(ha["mat"][m], hb["mat"][m])
    assert np.array_equal(ha["front"][m], hb["front"][m])
    # t: 1e-5 relative — or, for origins that sit (almost) on the surface, within 8 f32 quanta of the scene
    # coordinates in DISTANCE (a world-space point stored in f32 cannot be placed more precisely than that)
    dist_err = np.abs(ha["t"][m] - hb["t"][m]) * d_len[m]
    t_ok = (rel_err(ha["t"][m], hb["t"][m], 1e-3) <= REL) | (dist_err <= 8 * 2.0 ** -24 * pos_scale)
    assert t_ok.all(), (rel_err(ha["t"][m], hb["t"][m], 1e-3).max(), dist_err.max())
    assert np.mean(rel_err(ha["t"][m], hb["t"][m], 1e-3) <= REL) > 0.999
    assert rel_err(ha["p"][m], hb["p"][m], pos_scale).max() <= REL
    assert np.abs(ha["normal"][m] - hb["normal"][m]).max() <= REL
    if check_uv:
        # u wraps at the sphere seam: compare on the circle
        du = np.abs(ha["u"][m] - hb["u"][m])
        du = np.minimum(du, 1 - du)
        polar = (hb["v"][m] < 1e-3) | (hb["v"][m] > 1 - 1e-3)
        assert du[~polar].max() <= 2 * REL and np.abs(ha["v"][m] - hb["v"][m]).max() <= 5 * REL
    assert np.array_equal(ha["ndraw"][ok], hb["ndraw"][ok])


# --------------------------------------------------------------------------------------------- RNG / camera
def test_philox_bit_exact(gpu, orc):
    rs = np.random.RandomState(0)
    ctr = rs.randint(0, 2 ** 32, (100000, 4), dtype=np.uint64).astype(np.uint32)
    key = rs.randint(0, 2 ** 32, (100000, 2), dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(gpu.philox(ctr, key), orc.philox(ctr, key))
    assert list(gpu.philox([0, 0, 0, 0], [0, 0])[0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]


@pytest.mark.parametrize("name", ["random_scene", "cornell_box", "final_scene", "simple_light"])
def test_get_ray(pkg, gpu, orc, name):
    sc, spec = pkg.scenes.build(orc, name)
    rs = np.random.RandomState(1)
    n = 200000
    s, t, xi = f32(rs.rand(n)), f32(rs.rand(n)), q24(rs, (n, 24))
    ga = gpu.test_get_ray(spec.camera(gpu, 1200, 800), s, t, xi)
    gb = orc.test_get_ray(spec.camera(orc, 1200, 800), s, t, xi)
    assert np.array_equal(ga["ndraw"], gb["ndraw"]) and (gb["ndraw"] >= 3).all()
    assert np.array_equal(ga["time"], gb["time"])
    scale = np.abs(np.array(spec.look_from)).max()
    assert rel_err(ga["origin"], gb["origin"], scale).max() <= REL
    dn = np.linalg.norm(gb["dir"], axis=1, keepdims=True)
    assert (np.abs(ga["dir"] - gb["dir"]) / dn).max() <= REL


def test_aabb_is_conservative_and_agrees(gpu, orc):
    rs = np.random.RandomState(2)
    n = 300000
    c = f32(rs.uniform(-50, 50, (n, 3))); h = f32(rs.uniform(0.01, 20, (n, 3)))
    mn, mx = f32(c - h), f32(c + h)
    o = f32(rs.uniform(-100, 100, (n, 3)))
    d = f32((c + rs.uniform(-1.5, 1.5, (n, 3)) * h - o) * rs.uniform(0.1, 3, (n, 1)))
    ga = gpu.test_aabb(mn, mx, o, d, 0.001, 1e30)
    gb = orc.test_aabb(mn, mx, o, d, 0.001, 1e30)
    assert (ga >= gb).all()                      # the f32 slab test never culls what the f64 reference hits
    # disagreements only where the interval is degenerate at the 1e-5 level
    with np.errstate(divide="ignore", invalid="ignore"):
        t0, t1 = (mn - o) / d, (mx - o) / d
    lo, hi = np.minimum(t0, t1).max(1), np.maximum(t0, t1).min(1)
    lo = np.maximum(lo, 0.001)
    with np.errstate(divide="ignore", invalid="ignore"):
        oi = np.nanmax(np.abs(o / d), axis=1)
    # the device pads the exit distance by 2 ulp and by 4 ulp of max|o/d| (rounding of the fma form): agreement is
    # required wherever the f64 interval is wider than that
    degenerate = np.abs(hi - lo) <= 1e-4 * np.maximum(np.abs(hi), 1.0) + 1e-6 * oi
    assert not (ga != gb)[~degenerate].any()
    assert 0.2 < gb.mean() < 0.95


# --------------------------------------------------------------------------------------------- primitives
def _prim_cases(pkg, sc):
    m = sc.lambertian(sc.tex_solid((0.5, 0.5, 0.5)))
    m2 = sc.metal((0.5, 0.5, 0.5), 0.1)
    cases = {}
    cases["sphere"] = (sc.sphere(m, (1.5, -0.5, 2.0), 1.25), (1.5, -0.5, 2.0), 1.25)
    cases["small_sphere"] = (sc.sphere(m2, (3.7, 0.2, -4.1), 0.2), (3.7, 0.2, -4.1), 0.2)
    cases["ground_sphere"] = (sc.sphere(m, (0.0, -1000.0, 0.0), 1000.0), (0.0, 0.0, 0.0), 12.0)
    cases["moving_sphere"] = (sc.moving_sphere(m, (0.3, 0.2, -1.0), (0.3, 0.65, -1.0), 0.0, 1.0, 0.2), (0.3, 0.4, -1.0), 0.45)
    cases["xy_rect"] = (sc.xy_rect(m, 3, 5, 1, 3, -2), (4, 2, -2), 1.6)
    cases["xz_rect"] = (sc.xz_rect(m, 213, 343, 227, 332, 554), (278, 554, 280), 90)
    cases["yz_rect"] = (sc.yz_rect(m2, 0, 555, 0, 555, 555), (555, 278, 278), 400)
    cases["box"] = (sc.box((0, 0, 0), (165, 330, 165), m), (82, 165, 82), 200)
    box = sc.box((0, 0, 0), (165, 330, 165), m)
    cases["translate_rotate_box"] = (sc.translate(sc.rotate_y(15.0, box), (265, 0, 295)), (347, 165, 377), 220)
    cases["rotate_box"] = (sc.rotate_y(-18.0, sc.box((0, 0, 0), (165, 165, 165), m2)), (80, 82, 80), 160)
    cases["translate_sphere"] = (sc.translate(sc.sphere(m, (0, 0, 0), 10.0), (100, 20, -30)), (100, 20, -30), 10)
    cases["rotate_translate_rotate_sphere"] = (
        sc.rotate_y(30.0, sc.translate(sc.rotate_y(45.0, sc.moving_sphere(m, (5, 0, 0), (5, 3, 0), 0.0, 1.0, 2.0)), (10, 0, 0))),
        (12.0, 1.5, -9.0), 6.0)
    bvh = sc.bvh_node([sc.sphere(m, (x * 3.0, 0.0, z * 3.0), 1.0) for x in range(4) for z in range(4)])
    cases["bvh_node_of_spheres"] = (bvh, (4.5, 0, 4.5), 8)
    cases["translate_rotate_bvh"] = (sc.translate(sc.rotate_y(15.0, bvh), (-100, 270, 395)), (-94, 270, 398), 9)
    return cases


CASE_NAMES = ["sphere", "small_sphere", "ground_sphere", "moving_sphere", "xy_rect", "xz_rect", "yz_rect", "box",
              "translate_rotate_box", "rotate_box", "translate_sphere", "rotate_translate_rotate_sphere",
              "bvh_node_of_spheres", "translate_rotate_bvh"]


@pytest.mark.parametrize("case", CASE_NAMES)
def test_hittable_hit(pkg, gpu, orc, case):
    """Hittable::hit (src/hittable.rs:209-252) per variant / composition, fixed rays."""
    import ctypes as C
    a, b = pkg.Scene(gpu), pkg.Scene(orc)
    ta, center, radius = _prim_cases(pkg, a)[case]
    tb, _, _ = _prim_cases(pkg, b)[case]          # hittable ids are library-internal
    if case != "ground_sphere":                   # aim at the reference's own bounding box (src/hittable.rs:475-525)
        mn, mx = np.zeros(3), np.zeros(3)
        dp = C.POINTER(C.c_double)
        assert orc.f("bounding_box")(b.h, tb, 0.0, 1.0, mn.ctypes.data_as(dp), mx.ctypes.data_as(dp)) == 0
        center, radius = (mn + mx) / 2, float(np.linalg.norm(mx - mn) / 2)
    rs = np.random.RandomState(zlib.crc32(case.encode()) % 1000)
    n = 120000
    center = np.array(center, float)
    # origins on a shell around the target (some inside), directions towards a jittered point of it, un-normalised
    dirs = rs.randn(n, 3); dirs /= np.linalg.norm(dirs, axis=1, keepdims=True)
    o = f32(center + dirs * radius * rs.uniform(0.2, 6.0, (n, 1)))
    tgt = center + rs.uniform(-0.7, 0.7, (n, 3)) * radius
    d = f32((tgt - o) * rs.uniform(0.05, 2.0, (n, 1)))
    tm = f32(rs.rand(n))
    hb, ok = stable(lambda oo, dd, tt: b.test_hit(tb, oo, dd, tt), o, d, tm, radius)
    ha = a.test_hit(ta, o, d, tm)
    assert 0.1 < hb["hit"].mean() < 0.999, hb["hit"].mean()
    compare_hits(ha, hb, ok, np.abs(center).max() + radius, np.linalg.norm(d, axis=1))


def test_hit_range_edges(pkg, gpu, orc):
    """t_min / t_max handling: rays starting ON the surface (t_min = 0.001 self-intersection guard), finite t_max."""
    a, b = pkg.Scene(gpu), pkg.Scene(orc)
    for sc in (a, b):
        m = sc.dielectric(1.5)
        sc.target = sc.sphere(m, (0.0, 1.0, 0.0), 1.0)
    rs = np.random.RandomState(8)
    n = 100000
    nrm = rs.randn(n, 3); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    o = f32(np.array([0, 1.0, 0]) + nrm)                     # on the sphere (to f32 rounding)
    d = f32(rs.randn(n, 3))
    hb, ok = stable(lambda oo, dd, tt: b.test_hit(b.target, oo, dd, tt), o, d, np.zeros(n), 1.0)
    ha = a.test_hit(a.target, o, d, np.zeros(n))
    compare_hits(ha, hb, ok, 2.0, np.linalg.norm(d, axis=1), min_stable=0.97)        # measured 0.9806 (origins ON the sphere)
    inward = np.einsum("ij,ij->i", d, o - np.array([0, 1.0, 0])) < -0.2 * np.linalg.norm(d, axis=1)
    assert (ha["hit"][inward] == 1).all() and (ha["front"][inward] == 0).all()      # exits through the far side
    # finite t_max clips
    ha2 = a.test_hit(a.target, o, d, np.zeros(n), t_max=0.5); hb2 = b.test_hit(b.target, o, d, np.zeros(n), t_max=0.5)
    agree = ha2["hit"] == hb2["hit"]
    assert agree[ok].mean() > 0.9995 and (hb2["t"][hb2["hit"] == 1] <= 0.5).all()


@pytest.mark.parametrize("name", ["cornell_box_smoke", "final_scene"])
def test_constant_medium_hit(pkg, gpu, orc, name):
    """hit_constant_medium (src/hittable.rs:417-473) inside the world scan: draw gating, free-flight distance."""
    a, b, spec = build_both(pkg, gpu, orc, name)
    rs = np.random.RandomState(3)
    g = camera_rays(pkg, orc, spec, 150000, rs)
    xi = q24(rs, (150000, 4))
    xi[xi == 0] = 0.5
    o, d, tm = f32(g["origin"]), f32(g["dir"]), f32(g["time"])
    hb, ok = stable(lambda oo, dd, tt, xi=xi: b.test_hit(-1, oo, dd, tt, xi=xi), o, d, tm, 1000.0)
    ha = a.test_hit(-1, o, d, tm, xi=xi)
    # smoke: a ray can cross both boxes (2 draws).  final_scene: the fog sphere always draws; the r=70 medium never
    # does for camera rays, because its glass twin earlier in the list clamps t_max to the entry point (:436-441)
    assert hb["ndraw"].max() == (2 if name == "cornell_box_smoke" else 1)
    assert (hb["ndraw"] >= (1 if name == "final_scene" else 0)).all()
    compare_hits(ha, hb, ok, 1000.0, np.linalg.norm(d, axis=1), check_uv=False)
    med = ok & (hb["hit"] == 1) & np.all(hb["normal"] == [1, 0, 0], axis=1) & (hb["u"] == 0)
    assert med.sum() > 100            # some rays really scatter inside the media


@pytest.mark.parametrize("name", ["random_scene", "two_spheres", "two_perlin_spheres", "earth", "simple_light", "cornell_box", "final_scene"])
def test_world_bvh_hit(pkg, gpu, orc, name):
    """World closest hit through OUR flattened SAH BVH vs the oracle's hit_hittables over the reference tree:
    primary rays, then secondary rays leaving the primary hit points."""
    a, b, spec = build_both(pkg, gpu, orc, name)
    rs = np.random.RandomState(4)
    n = 150000
    g = camera_rays(pkg, orc, spec, n, rs)
    xi = q24(rs, (n, 4)); xi[xi == 0] = 0.5
    o, d, tm = f32(g["origin"]), f32(g["dir"]), f32(g["time"])
    scale = max(np.abs(np.array(spec.look_from)).max(), 10.0)
    hb, ok = stable(lambda oo, dd, tt: b.test_hit(-1, oo, dd, tt, xi=xi), o, d, tm, scale)
    ha = a.test_hit(-1, o, d, tm, xi=xi)
    compare_hits(ha, hb, ok, scale, np.linalg.norm(d, axis=1), check_uv=name in ("earth", "final_scene"))
    # secondary rays: origin = hit point (f32), direction = normal + unit vector (Lambertian-like, un-normalised)
    hit = hb["hit"] == 1
    o2 = f32(hb["p"][hit]); nn = hb["normal"][hit]
    v = rs.randn(hit.sum(), 3); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d2 = f32(nn + v); tm2 = tm[hit]; xi2 = xi[hit]
    keep = np.linalg.norm(d2, axis=1) > 0.05
    o2, d2, tm2, xi2 = o2[keep], d2[keep], tm2[keep], xi2[keep]
    hb2, ok2 = stable(lambda oo, dd, tt: b.test_hit(-1, oo, dd, tt, xi=xi2), o2, d2, tm2, scale)
    ha2 = a.test_hit(-1, o2, d2, tm2, xi=xi2)
    compare_hits(ha2, hb2, ok2, scale, np.linalg.norm(d2, axis=1), min_stable=SECONDARY_STABLE[name], check_uv=name in ("earth", "final_scene"))


# --------------------------------------------------------------------------------------------- materials / textures
def _scatter_inputs(rs, n):
    nrm = rs.randn(n, 3); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    d = rs.randn(n, 3) * rs.uniform(0.05, 12, (n, 1))
    flip = np.einsum("ij,ij->i", d, nrm) > 0
    d[flip] -= 2 * np.einsum("ij,ij->i", d, nrm)[flip, None] * nrm[flip]      # face-forward: dot(d, n) < 0
    p = rs.uniform(-20, 20, (n, 3))
    return f32(p - d), f32(d), f32(rs.rand(n)), f32(p), f32(nrm), rs.randint(0, 2, n).astype(np.int32), f32(rs.rand(n)), f32(rs.rand(n))


@pytest.mark.parametrize("kind", ["lambertian_solid", "lambertian_checker", "metal", "metal_fuzz1", "dielectric", "diffuse_light", "isotropic"])
def test_material_scatter(pkg, gpu, orc, kind):
    """Material::scatter / emitted (src/material.rs:15-94) with fixed random inputs."""
    scenes = []
    for lib in (gpu, orc):
        sc = pkg.Scene(lib)
        mats = dict(lambertian_solid=lambda: sc.lambertian(sc.tex_solid((0.4, 0.2, 0.1))),
                    lambertian_checker=lambda: sc.lambertian(sc.tex_checker((0.2, 0.5, 0.5), (0.9, 0.9, 0.9))),
                    metal=lambda: sc.metal((0.7, 0.6, 0.5), 0.3), metal_fuzz1=lambda: sc.metal((0.8, 0.8, 0.9), 1.0),
                    dielectric=lambda: sc.dielectric(1.5), diffuse_light=lambda: sc.diffuse_light(sc.tex_solid((15, 15, 15))),
                    isotropic=lambda: sc.isotropic(sc.tex_solid((0.2, 0.4, 0.9))))
        mat = mats[kind]()
        sc.push(sc.sphere(mat, (0, 0, 0), 1.0))
        scenes.append((sc, mat))
    scenes[0][0].commit(1, 0)
    rs = np.random.RandomState(6)
    n = 200000
    ro, rd, rt, p, nrm, front, u, v = _scatter_inputs(rs, n)
    xi = q24(rs, (n, 48))
    ra = scenes[0][0].test_scatter(scenes[0][1], ro, rd, rt, p, nrm, front, u, v, xi)
    rb = scenes[1][0].test_scatter(scenes[1][1], ro, rd, rt, p, nrm, front, u, v, xi)
    assert (rb["ndraw"] >= 0).all()
    same = (ra["scattered"] == rb["scattered"]) & (ra["ndraw"] == rb["ndraw"])
    # conditioning: drop inputs whose ORACLE output moves by more than 1e-5/3 under a 1e-7 relative nudge of the ray
    # direction (near-critical refraction, 1 - |r_perp|^2 -> 0 in src/math.rs:114) — no f32 evaluation can hold 1e-5 there
    dnb = np.maximum(np.linalg.norm(rb["dir"], axis=1, keepdims=True), 1e-2)
    well = np.ones(n, bool)
    for _ in range(4):
        rb2 = scenes[1][0].test_scatter(scenes[1][1], ro, rd * (1 + 1e-7 * rs.uniform(-1, 1, rd.shape)), rt, p, nrm, front, u, v, xi)
        well &= (rb2["scattered"] == rb["scattered"]) & ((np.abs(rb2["dir"] - rb["dir"]) / dnb).max(1) <= REL / 3)
    assert well.mean() > 0.995
    if kind == "dielectric":    # reflect/refract branch: side of the surface the new ray leaves on
        same &= np.sign(np.einsum("ij,ij->i", ra["dir"], nrm)) == np.sign(np.einsum("ij,ij->i", rb["dir"], nrm))
    if kind == "lambertian_checker":
        same &= np.all(ra["attenuation"] == f32(rb["attenuation"]), axis=1)
    assert same.mean() > 0.9995, same.mean()
    m = same & well & (rb["scattered"] == 1)
    if kind != "diffuse_light":
        assert m.sum() > 0.4 * n
        dn = np.maximum(np.linalg.norm(rb["dir"][m], axis=1, keepdims=True), 1e-2)
        assert (np.abs(ra["dir"][m] - rb["dir"][m]) / dn).max() <= REL
        assert np.array_equal(ra["origin"][m], rb["origin"][m]) and np.array_equal(ra["time"][m], rb["time"][m])
        assert np.abs(ra["attenuation"][m] - rb["attenuation"][m]).max() <= REL
    assert rel_err(ra["emitted"], rb["emitted"], 1.0).max() <= REL
    if kind == "diffuse_light":
        assert (ra["scattered"] == 0).all() and np.allclose(ra["emitted"], 15.0)


@pytest.mark.parametrize("kind", ["checker", "noise4", "noise0.1", "image"])
def test_texture_value(pkg, gpu, orc, kind):
    """Texture::get_color_value (src/texture.rs:30-75) + Perlin::turb/noise (src/perlin.rs:32-108)."""
    out = []
    rs = np.random.RandomState(7)
    n = 200000
    span = dict(checker=30.0, noise4=40.0, image=1.0)["noise4" if kind.startswith("noise") else kind]
    if kind == "noise0.1":
        span = 400.0
    p = f32(rs.uniform(-span, span, (n, 3)))
    u, v = f32(rs.uniform(-0.1, 1.1, n)), f32(rs.uniform(-0.1, 1.1, n))
    for lib in (gpu, orc):
        sc = pkg.Scene(lib)
        if kind == "checker":
            t = sc.tex_checker((0.2, 0.5, 0.5), (0.9, 0.9, 0.9))
        elif kind.startswith("noise"):
            t = sc.tex_noise(*pkg.scenes.perlin_tables(pkg.scenes.HostRng(3)), float(kind[5:]))
        else:
            t = sc.tex_image(pkg.scenes.earth_texels())
        sc.push(sc.sphere(sc.lambertian(t), (0, 0, 0), 1.0))
        if lib is gpu:
            sc.commit(1, 0)
        out.append(sc.test_texture(t, u, v, p))
    ga, gb = out
    if kind == "checker":
        sines = np.sin(10 * p[:, 0]) * np.sin(10 * p[:, 1]) * np.sin(10 * p[:, 2])
        ok = np.abs(sines) > 1e-4
        assert np.array_equal(ga[ok], f32(gb[ok])) and np.mean(np.any(ga != f32(gb), axis=1)) < 1e-4
    elif kind == "image":
        uu, vv = np.clip(u, 0, 1) * 1024, (1 - np.clip(v, 0, 1)) * 512
        ok = (np.abs(uu - np.round(uu)) > 1e-3) & (np.abs(vv - np.round(vv)) > 1e-3)
        assert np.abs(ga[ok] - gb[ok]).max() <= 1e-7 and np.mean(np.any(np.abs(ga - gb) > 1e-7, axis=1)) < 1e-3
    else:
        # marble = 0.5 (1 + sin(scale z + 10 turb)): f32 noise values, octave sum and phase in f64 — the north-star 1e-5 holds for every point
        err = np.abs(ga - gb).max(1)
        record("noise_texture", kind=kind, p999=np.percentile(err, 99.9), max=err.max())
        assert err.max() <= REL, (np.percentile(err, 99.9), err.max())          # measured max 9.3e-7 (f64 octave sum + phase)


# --------------------------------------------------------------------------------------------- whole paths
# fraction of paths whose segment count and radiance match the oracle's; measured on the B200 (round 2, profiles/
# r2_parity_measured.jsonl): 0.9995 / 1.0 / 0.9993 / 0.99997 / 1.0 / 0.9993 / 0.99998 / 0.9973 — the bars sit just under
PATH_BARS = dict(random_scene=0.996, two_spheres=0.997, two_perlin_spheres=0.996, earth=0.997, simple_light=0.997,
                 cornell_box=0.996, cornell_box_smoke=0.997, final_scene=0.993)
# fraction of SECONDARY rays whose oracle answer is stable under a 1e-6 input perturbation (measured 0.9905 / 0.9974 /
# 0.9966 / 0.9979 / 0.9934 / 0.8311 / 0.9143: rays that leave a wall of the Cornell box graze the other walls' edges)
SECONDARY_STABLE = dict(random_scene=0.98, two_spheres=0.99, two_perlin_spheres=0.99, earth=0.99, simple_light=0.985,
                        cornell_box=0.82, final_scene=0.90)


@pytest.mark.parametrize("name", list(PATH_BARS))
def test_paths_match_oracle_and_golden(pkg, gpu, orc, name):
    """ray_color (src/main.rs:19-38) end to end, path by path with shared Philox keys: same number of segments and
    radiance within 1e-3 relative for at least PATH_BARS of the paths (f32 vs f64 can only diverge at discrete
    decisions: rejection-loop acceptance, Schlick test, grazing hits, checker cell), equal means."""
    a, b, spec = build_both(pkg, gpu, orc, name)
    g = np.load(os.path.join(ROOT, "tests", "golden", "oracle_paths_v2.npz"))
    W, H = int(g["size"][0]), int(g["size"][1])
    p = pkg.make_params(W, H, 64, background=spec.background, seed=int(g["seed"]))
    ra, sa = a.trace_paths(spec.camera(gpu, W, H), p, g["px"], g["py"], g["sample"])
    rb, sb = g[name + "_rgb"], g[name + "_seg"]
    good = (sa == sb) & (np.abs(ra - rb).max(1) <= 1e-3 * np.maximum(1.0, np.abs(rb).max(1)))
    record("paths_golden", name=name, match=good.mean())
    assert good.mean() >= PATH_BARS[name], good.mean()
    # a bigger live sample
    rs = np.random.RandomState(12)
    n = 60000
    px, py, sm = rs.randint(0, W, n), rs.randint(0, H, n), rs.randint(0, 4096, n)
    ra, sa = a.trace_paths(spec.camera(gpu, W, H), p, px, py, sm)
    rb, sb = b.trace_paths(spec.camera(orc, W, H), p, px, py, sm)
    good = (sa == sb) & (np.abs(ra - rb).max(1) <= 1e-3 * np.maximum(1.0, np.abs(rb).max(1)))
    record("paths_live", name=name, match=good.mean())
    assert good.mean() >= PATH_BARS[name], good.mean()
    assert np.isfinite(ra).all()
    se = rb.std(0) / math.sqrt(n) * math.sqrt(2 * (1 - good.mean()) + 1e-6)   # only unmatched paths contribute noise
    assert (np.abs(ra.mean(0) - rb.mean(0)) <= 5 * se + 1e-5).all(), (ra.mean(0), rb.mean(0), se)
    assert abs(sa.mean() - sb.mean()) < 0.03 * sb.mean()


# --------------------------------------------------------------------------------------------- full renders
def _psnr(a, b):
    mse = np.mean((a - b) ** 2)
    return 99.0 if mse == 0 else 10 * math.log10(1.0 / mse)


def _display(sum_, spp):
    return np.clip(np.sqrt(np.clip(sum_ / spp, 0, None)), 0, 0.999)      # write_color without the quantisation


@pytest.mark.parametrize("name,W,H", [("random_scene", 48, 32), ("two_perlin_spheres", 48, 27), ("earth", 48, 27),
                                      ("simple_light", 32, 32), ("cornell_box", 32, 32), ("cornell_box_smoke", 24, 24),
                                      ("final_scene", 32, 32)])
def test_render_psnr_4096spp(pkg, gpu, orc, name, W, H):
    """North-star render bar: at 4096 spp and equal Philox keys the megakernel's image is within PSNR >= 40 dB of the
    oracle's, and per-pixel means within 3 sigma (sigma^2 = (s_ref^2 + s_gpu^2)/n, s_gpu ~ s_ref)."""
    spp = 4096
    a, b, spec = build_both(pkg, gpu, orc, name)
    p = pkg.make_params(W, H, spp, background=spec.background, seed=5)
    img, st = a.render(spec.camera(gpu, W, H), p)
    ref = b.render_oracle(spec.camera(orc, W, H), p, threads=0, sumsq=True)
    assert st["paths"] == W * H * spp and np.isfinite(img).all()
    psnr = _psnr(_display(img.astype(np.float64), spp), _display(ref["sum"], spp))
    record("psnr_4096", name=name, psnr=psnr)
    assert psnr >= 40.0, psnr
    mean_ref = ref["sum"] / spp
    var = np.maximum(ref["sumsq"] / spp - mean_ref ** 2, 0)
    sigma = np.sqrt(2 * var / spp)
    out = np.abs(img / spp - mean_ref) > 3 * sigma + 2e-5 * np.maximum(mean_ref, 1.0)     # + f32 accumulation of spp terms
    assert out.mean() <= 0.005, out.mean()


def test_render_unbiased_vs_independent_oracle(pkg, gpu, orc):
    """Same 3-sigma test with INDEPENDENT noise (different seeds): catches a bias that common random numbers would
    hide.  Per pixel on the bounded-radiance book-1 scene."""
    name, W, H, spp = "random_scene", 60, 40, 256
    a, b, spec = build_both(pkg, gpu, orc, name)
    img, _ = a.render(spec.camera(gpu, W, H), pkg.make_params(W, H, spp, background=spec.background, seed=101))
    ref = b.render_oracle(spec.camera(orc, W, H), pkg.make_params(W, H, spp, background=spec.background, seed=202), threads=0, sumsq=True)
    mean_ref = ref["sum"] / spp
    var = np.maximum(ref["sumsq"] / spp - mean_ref ** 2, 0)
    sigma = np.sqrt(2 * var / spp)
    z = (img / spp - mean_ref) / (sigma + 1e-4)
    assert np.mean(np.abs(z) > 3) <= 0.01, np.mean(np.abs(z) > 3)
    tot_sigma = math.sqrt((2 * var / spp).sum()) / var.size
    assert abs((img / spp).mean() - mean_ref.mean()) <= 4 * tot_sigma + 1e-4


@pytest.mark.parametrize("name", ["cornell_box", "cornell_box_smoke", "final_scene", "simple_light"])
def test_radiance_distribution_unbiased_independent(pkg, gpu, orc, name):
    """Emissive scenes are heavy-tailed (a few paths carry radiance 15), so per-pixel sigma estimates are useless at
    test-sized spp.  Two-sample z-tests on INDEPENDENT path samples instead: the mean of min(L, cap) for several caps
    (finite variance, sensitive to leaks / trapped paths) and the path-length distribution."""
    a, b, spec = build_both(pkg, gpu, orc, name)
    W, H, n = 64, 64, 400000
    rs = np.random.RandomState(21)
    px, py, sm = rs.randint(0, W, n), rs.randint(0, H, n), rs.randint(0, 1 << 20, n)
    ra, sa = a.trace_paths(spec.camera(gpu, W, H), pkg.make_params(W, H, 64, background=spec.background, seed=303), px, py, sm)
    rb, sb = b.trace_paths(spec.camera(orc, W, H), pkg.make_params(W, H, 64, background=spec.background, seed=404), px, py, sm)
    for cap in (0.5, 2.0, 20.0):
        xa, xb = np.minimum(ra, cap).mean(1), np.minimum(rb, cap).mean(1)
        z = (xa.mean() - xb.mean()) / math.sqrt(xa.var() / n + xb.var() / n)
        assert abs(z) < 4.5, (cap, z, xa.mean(), xb.mean())
    z = (sa.mean() - sb.mean()) / math.sqrt(sa.var() / n + sb.var() / n)
    assert abs(z) < 4.5, (z, sa.mean(), sb.mean())
    # trapped paths (depth exhausted) must be as rare as in the reference arithmetic
    fa, fb = np.mean(sa == 50), np.mean(sb == 50)
    assert abs(fa - fb) <= 5 * math.sqrt((fa + fb + 1e-6) / n), (fa, fb)


def test_render_is_partition_invariant_full_size(pkg, gpu):
    """BASELINE config 1 at full size (1200x800, 500 spp, depth 50).  Size-independent property: the per-pixel sum
    is the sum of deterministic per-(pixel, sample) path values, so ANY split of the samples into work units —
    direct stores (one unit per tile) or atomically accumulated chunks — gives the same image up to f32 summation
    order; and a second seed gives a statistically equal but different image."""
    sc, spec = pkg.scenes.build(gpu, "random_scene")
    sc.commit(1, 0)
    cam = spec.camera(gpu, 1200, 800)
    imgs = []
    for spu in (500, 0, 64):
        img, st = sc.render(cam, pkg.make_params(1200, 800, 500, background=spec.background, seed=1, samples_per_unit=spu))
        assert st["paths"] == 480_000_000 and st["rays"] > st["paths"] and np.isfinite(img).all()
        if spu:
            assert sum(st["units_per_device"]) == 150 * 200 * math.ceil(500 / spu)
        imgs.append(img)
    for other in imgs[1:]:
        assert np.abs(other - imgs[0]).max() <= 2e-4 * np.abs(imgs[0]).max()
    assert imgs[0].min() >= 0 and imgs[0].max() <= 500.0 + 1e-3           # radiance <= 1 in this scene
    sky = imgs[0][:40].reshape(-1, 3) / 500                               # top rows: mostly sky (0.7, 0.8, 1.0)
    assert np.allclose(np.median(sky, 0), [0.7, 0.8, 1.0], atol=1e-3)
    img2, _ = sc.render(cam, pkg.make_params(1200, 800, 500, background=spec.background, seed=2))
    assert not np.array_equal(img2, imgs[0]) and abs(img2.mean() - imgs[0].mean()) < 2e-3 * imgs[0].mean()


@pytest.mark.parametrize("name,W,H,spp", [("two_spheres", 800, 450, 200), ("two_perlin_spheres", 800, 450, 200), ("earth", 800, 450, 200),
                                          ("simple_light", 600, 600, 200), ("cornell_box", 600, 600, 200), ("final_scene", 800, 800, 64)])
def test_full_resolution_configs(pkg, gpu, orc, name, W, H, spp):
    """BASELINE configs 2-4 at full resolution (reduced spp for C3/C4 to bound the test): finite, partition-invariant,
    and the box-downsampled image agrees with an oracle render of the same scene at 1/10 resolution within MC noise."""
    a, b, spec = build_both(pkg, gpu, orc, name)
    cam = spec.camera(gpu, W, H)
    img, st = a.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=3))
    img2, _ = a.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=3, samples_per_unit=spp))
    assert np.isfinite(img).all() and st["paths"] == W * H * spp
    assert np.abs(img - img2).max() <= 2e-4 * max(np.abs(img).max(), 1.0)
    f = 10
    w, h = W // f, H // f
    small = img[:h * f, :w * f].reshape(h, f, w, f, 3).mean((1, 3)) / spp
    ospp = 256
    ref = b.render_oracle(spec.camera(orc, w, h), pkg.make_params(w, h, ospp, background=spec.background, seed=9), threads=0)
    refm = ref["sum"] / ospp
    # same radiance field sampled on two pixel grids: compare the image means and a coarse 4x4 block average
    assert abs(small.mean() - refm.mean()) <= 0.03 * refm.mean() + 2e-3
    bh, bw = h // 4, w // 4
    blk = lambda x: x[:bh * 4, :bw * 4].reshape(4, bh, 4, bw, 3).mean((1, 3))
    assert np.abs(blk(small) - blk(refm)).max() <= 0.12 * max(blk(refm).max(), 0.05)


def test_final_scene_full_config(pkg, gpu):
    """BASELINE config 4 at FULL size (800x800, 10 000 spp, depth 50 = 6.4 G paths, ~8 s): finite, non-negative, the
    per-pixel sums of the first 1 000 samples (a separate render with the same Philox keys) are a prefix of it — the
    sample index is a counter coordinate, so mean(10 000 spp) must sit within MC noise of mean(1 000 spp) — and f32
    accumulation of 10 000 terms up to 7.0 stays within 1e-3 relative of the image mean."""
    sc, spec = pkg.scenes.build(gpu, "final_scene")
    sc.commit(1, 0)
    cam = spec.camera(gpu, 800, 800)
    full, st = sc.render(cam, pkg.make_params(800, 800, 10000, background=spec.background, seed=1))
    part, _ = sc.render(cam, pkg.make_params(800, 800, 1000, background=spec.background, seed=1))
    assert st["paths"] == 6_400_000_000 and np.isfinite(full).all() and full.min() >= 0
    m_full, m_part = full.mean() / 10000, part.mean() / 1000
    assert abs(m_full - m_part) <= 0.01 * m_full
    blk = lambda x: x.reshape(8, 100, 8, 100, 3).mean((1, 3))
    assert np.abs(blk(full) / 10000 - blk(part) / 1000).max() <= 0.05 * blk(full).max() / 10000


@pytest.mark.parametrize("W,H,spp,spu", [(37, 23, 160, 0), (101, 53, 163, 0), (64, 64, 500, 0), (1200, 800, 161, 0), (40, 20, 1000, 0), (37, 23, 163, 40)])
def test_work_units_cover_every_sample_once(pkg, gpu, W, H, spp, spu):
    """Work units = (tile, sample range) in two phases (big units, then short ones): with an empty world every
    sample returns exactly the background (1, 2, 4), so each pixel sum must be spp * background EXACTLY (small integers
    in f32) - a sample range covered twice or skipped shows up in every pixel of its tile."""
    sc = pkg.Scene(gpu)
    sc.commit(1, 0)
    cam = gpu.camera_new((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, W / H, 0.1, 5.0)
    for flags in (pkg.api.RTW_FLAG_KERNEL_MEGA, pkg.api.RTW_FLAG_KERNEL_POOL):
        img, st = sc.render(cam, pkg.make_params(W, H, spp, background=(1.0, 2.0, 4.0), samples_per_unit=spu, flags=flags))
        assert st["rays"] == W * H * spp
        assert (img == np.array([1.0, 2.0, 4.0], np.float32) * spp).all()


@pytest.mark.parametrize("name", ["random_scene", "cornell_box_smoke"])
def test_progressive_and_resume_equal_one_shot(pkg, gpu, name):
    """rtw_render_progressive (replaces the progress thread src/main.rs:557-582 and the all-or-nothing output :591-596):
    the sample index is a Philox counter coordinate, so passes of any size - and a render stopped by the callback and
    resumed from its buffer - give the image of ONE rtw_render call up to f32 summation order, with the same ray count."""
    sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    W, H, spp = 93, 61, 50
    cam = spec.camera(gpu, W, H)
    p = pkg.make_params(W, H, spp, background=spec.background, seed=9)
    ref, st0 = sc.render(cam, p)
    tol = 2e-5 * max(np.abs(ref).max(), 1.0)
    seen = []
    img, st1 = sc.render_progressive(cam, p, samples_per_pass=7, progress=lambda d, t, b: seen.append((d, t, float(b.sum()))) and False)
    assert [d for d, _, _ in seen] == [7, 14, 21, 28, 35, 42, 49, 50] and all(t == spp for _, t, _ in seen)
    assert all(b > a for (_, _, a), (_, _, b) in zip(seen, seen[1:]))          # the sums grow pass by pass
    assert st1["rays"] == st0["rays"] and st1["paths"] == W * H * spp and st1["kernel_launches"] == 8
    assert np.abs(img - ref).max() <= tol
    # stop after 20 samples, then resume from the buffer
    part, st2 = sc.render_progressive(cam, p, samples_per_pass=10, progress=lambda d, t, b: d >= 20)
    assert st2["paths"] == W * H * 20
    first20, _ = sc.render(cam, pkg.make_params(W, H, 20, background=spec.background, seed=9))
    assert np.abs(part - first20).max() <= tol                                   # a prefix of the sample sequence
    full, st3 = sc.render_progressive(cam, p, first_sample=20, samples_per_pass=0, buf=part.copy())
    assert st3["paths"] == W * H * 30 and st2["rays"] + st3["rays"] == st0["rays"]
    assert np.abs(full - ref).max() <= tol
    with pytest.raises(pkg.RtwError):
        sc.render_progressive(cam, p, first_sample=spp + 1, buf=part.copy())


def test_cuda_path_reproduces_the_reference_earth_image(pkg, gpu):
    """The CUDA path against pixels the REFERENCE wrote (generated_images/earth.ppm via tests/golden/, see
    test_oracle_kat.py): globe texture orientation, framing, shading level and gamma, through rtw_render."""
    from test_oracle_kat import earth_reference_check
    sc, spec = pkg.scenes.build(gpu, "earth")
    sc.commit(1, 0)
    cam = spec.camera(gpu, 400, 225)
    img, _ = sc.render(cam, pkg.make_params(400, 225, 256, background=spec.background, seed=3))
    earth_reference_check(pkg, cam, img.astype(np.float64), 256)


def test_render_edge_cases(pkg, gpu, orc):
    # empty world: every sample returns the background (src/main.rs:37)
    sc = pkg.Scene(gpu)
    sc.commit(1, 0)
    cam = gpu.camera_new((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, 37 / 23, 0.1, 5.0)
    img, st = sc.render(cam, pkg.make_params(37, 23, 7, background=(0.25, 0.5, 0.75)))        # ragged tiles, odd spp
    assert np.allclose(img, np.array([0.25, 0.5, 0.75]) * 7, rtol=1e-6) and st["rays"] == 37 * 23 * 7
    # depth 0: black; depth 1: only directly visible emitters / background
    a, b, spec = build_both(pkg, gpu, orc, "simple_light")
    camg, camo = spec.camera(gpu, 33, 17), spec.camera(orc, 33, 17)
    img, _ = a.render(camg, pkg.make_params(33, 17, 3, max_depth=0, background=(0.1, 0.1, 0.1)))
    assert (img == 0).all()
    for depth in (1, 2):
        p = pkg.make_params(33, 17, 16, max_depth=depth, background=(0.1, 0.2, 0.3), seed=4)
        img, _ = a.render(camg, p)
        ref = b.render_oracle(camo, p, threads=0)
        assert np.abs(img - ref["sum"]).max() <= 1e-3 * max(ref["sum"].max(), 1) + 0.35    # a handful of diverged paths at most
        assert abs(img.mean() - ref["sum"].mean()) <= 0.01 * ref["sum"].mean() + 1e-3
    # minimum size, spp = 1
    img, st = a.render(spec.camera(gpu, 2, 2), pkg.make_params(2, 2, 1, background=(0, 0, 0)))
    assert img.shape == (2, 2, 3) and st["paths"] == 4
    with pytest.raises(pkg.RtwError):
        a.render(camg, pkg.make_params(1, 1, 1))
    # a queued ray packs segment (6 bits) and sample index (21 bits; pool kernel 17): out-of-range requests are refused
    for bad in (dict(max_depth=64), dict(spp=(1 << 20) + 1)):
        kw = dict(spp=1, max_depth=50); kw.update(bad)
        with pytest.raises(pkg.RtwError):
            a.render(camg, pkg.make_params(33, 17, kw["spp"], max_depth=kw["max_depth"]))
    with pytest.raises(pkg.RtwError):
        a.render(camg, pkg.make_params(33, 17, (1 << 17) + 1, flags=pkg.api.RTW_FLAG_KERNEL_POOL))
    img, _ = a.render(camg, pkg.make_params(33, 17, 2, max_depth=63))
    assert np.isfinite(img).all()
    # editing the scene after commit must be re-committed
    a.push(a.sphere(1, (0, 0, 0), 1.0))
    with pytest.raises(pkg.RtwError) as e:
        a.render(camg, pkg.make_params(8, 8, 1))
    assert e.value.code == -6


@pytest.mark.parametrize("name,W,H,spp", [("random_scene", 203, 117, 48), ("cornell_box_smoke", 96, 96, 64), ("final_scene", 100, 100, 32), ("earth", 120, 67, 32)])
def test_pool_kernel_equals_megakernel(pkg, gpu, name, W, H, spp):
    """The warp-pool (shared-memory wavefront) kernel and the megakernel schedule the SAME per-(pixel, sample) paths
    (same Philox keys, same device functions): their images must agree to f32 summation order.  Ragged sizes on purpose."""
    sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    cam = spec.camera(gpu, W, H)
    mega, st0 = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=5, flags=pkg.api.RTW_FLAG_KERNEL_MEGA))
    pool, st1 = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=5, flags=pkg.api.RTW_FLAG_KERNEL_POOL))
    assert st0["paths"] == st1["paths"]
    assert st0["rays"] == st1["rays"]
    assert np.abs(mega - pool).max() <= 2e-4 * max(np.abs(mega).max(), 1.0)


@pytest.mark.parametrize("name,W,H,spp", [("random_scene", 403, 227, 40), ("cornell_box", 150, 150, 40), ("final_scene", 160, 160, 40),
                                          ("two_perlin_spheres", 200, 113, 40), ("simple_light", 200, 113, 40), ("earth", 200, 113, 40)])
def test_tile_culling_is_conservative(pkg, gpu, name, W, H, spp):
    """Primary rays only test their tile's candidate list (interval walk of the BVH with the tile's ray bundle).  With
    RTW_FLAG_NO_TILE_CULL they traverse the BVH like every other ray: a list that missed a primitive would change the
    first hit of some path, hence the ray count and the image.  Only the f32 order of the tile atomics may differ."""
    sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    cam = spec.camera(gpu, W, H)
    a, st0 = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=11, flags=pkg.api.RTW_FLAG_KERNEL_MEGA))
    b, st1 = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=11,
                                            flags=pkg.api.RTW_FLAG_KERNEL_MEGA | pkg.api.RTW_FLAG_NO_TILE_CULL))
    assert st0["rays"] == st1["rays"]
    assert np.abs(a - b).max() <= 1e-5 * max(np.abs(a).max(), 1.0)


@pytest.mark.parametrize("name,extent", [("random_scene", 12.0), ("cornell_box", 500.0), ("final_scene", 500.0)])
def test_tile_culling_with_random_cameras(pkg, gpu, name, extent):
    """Same check as above with cameras the scenes were not designed for: inside the scene, looking anywhere, narrow
    and wide fields of view, pinhole to very large apertures, near and far focus planes, tilted up vectors - the
    interval walk over the tile's ray bundle (lens box -> focus window) must stay a superset in all of them."""
    sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    rs = np.random.RandomState(17)
    centre = np.array(spec.look_at, float)
    W, H, spp = 120, 67, 12
    for k in range(10):
        look_from = centre + rs.uniform(-extent, extent, 3) * (0.2 if k % 3 == 0 else 1.0)
        look_at = centre + rs.uniform(-extent, extent, 3) * 0.5
        vup = (0.0, 1.0, 0.0) if k % 2 == 0 else tuple(rs.uniform(-1, 1, 3))
        vfov = float(rs.choice([5.0, 20.0, 40.0, 90.0, 140.0]))
        aperture = float(rs.choice([0.0, 0.1, 2.0, 0.3 * extent]))
        focus = float(rs.choice([0.05 * extent, extent, 10.0 * extent]))
        cam = gpu.camera_new(tuple(look_from), tuple(look_at), vup, vfov, W / H, aperture, focus, 0.0, 1.0)
        p0 = pkg.make_params(W, H, spp, background=(0.7, 0.8, 1.0), seed=100 + k, flags=pkg.api.RTW_FLAG_KERNEL_MEGA)
        p1 = pkg.make_params(W, H, spp, background=(0.7, 0.8, 1.0), seed=100 + k,
                             flags=pkg.api.RTW_FLAG_KERNEL_MEGA | pkg.api.RTW_FLAG_NO_TILE_CULL)
        a, st0 = sc.render(cam, p0)
        b, st1 = sc.render(cam, p1)
        assert st0["rays"] == st1["rays"], (name, k, st0["rays"], st1["rays"])
        assert np.abs(a - b).max() <= 1e-5 * max(np.abs(a).max(), 1.0), (name, k)


def test_write_color_bit_exact(pkg, gpu, orc):
    """write_color (src/math.rs:119-132): gamma 2, clamp, *256 truncation — byte-exact against the oracle."""
    import ctypes as C
    a, spec = pkg.scenes.build(gpu, "random_scene")
    a.commit(1, 0)
    spp = 10
    img, _ = a.render(spec.camera(gpu, 200, 120), pkg.make_params(200, 120, spp, background=spec.background))
    img[0, 0] = [np.nan, -1.0, 1e9]
    n = img.shape[0] * img.shape[1]
    out_g = np.zeros(n * 3, np.uint8); out_o = np.zeros(n * 3, np.uint8)
    gpu.check(gpu.f("write_color")(img.ctypes.data_as(C.POINTER(C.c_float)), n, spp, out_g.ctypes.data_as(C.POINTER(C.c_uint8))))
    d = img.astype(np.float64)
    orc.f("write_color")(d.ctypes.data_as(C.POINTER(C.c_double)), n, spp, out_o.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert np.array_equal(out_g, out_o) and list(out_g[:3]) == [0, 0, 255]


def test_two_gpus_in_process(pkg, gpu):
    """Tiles from ONE atomic counter shared by both GPUs, finished tiles added into GPU 0's framebuffer over NVLink."""
    if gpu.f("device_count")() < 2:
        pytest.skip("needs 2 GPUs")
    sc, spec = pkg.scenes.build(gpu, "random_scene")
    sc.commit(2, 0)
    cam = spec.camera(gpu, 600, 400)
    one, st1 = sc.render(cam, pkg.make_params(600, 400, 64, background=spec.background, n_gpus=1))
    two, st2 = sc.render(cam, pkg.make_params(600, 400, 64, background=spec.background, n_gpus=2))
    assert st2["n_devices"] == 2 and min(st2["units_per_device"][:2]) > 0.2 * sum(st2["units_per_device"])
    assert np.abs(one - two).max() <= 2e-4 * one.max()


@pytest.mark.parametrize("width", [2, 8])
def test_two_gpus_in_process_device_built_scene(pkg, gpu, monkeypatch, width):
    """A scene whose BVH is built ON THE DEVICE (bvh_build.cu) is built once on the first GPU and copied to the other
    replica over NVLink: the two-GPU image equals the one-GPU image, both GPUs work."""
    if gpu.f("device_count")() < 2:
        pytest.skip("needs 2 GPUs")
    monkeypatch.setenv("RTW_BVH", str(width))
    monkeypatch.setenv("RTW_DEVICE_BUILD", "1")
    sc = pkg.Scene(gpu)
    spec = pkg.scenes.sweep_scene(sc, 50000, seed=2)
    sc.commit(2, 0)
    cam = spec.camera(gpu, 960, 540)                  # enough units that the second GPU's kernel finds work when it starts
    one, st1 = sc.render(cam, pkg.make_params(960, 540, 64, background=spec.background, n_gpus=1))
    two, st2 = sc.render(cam, pkg.make_params(960, 540, 64, background=spec.background, n_gpus=2))
    assert st2["n_devices"] == 2 and min(st2["units_per_device"][:2]) > 0.2 * sum(st2["units_per_device"])
    assert st1["rays"] == st2["rays"] and st2["n_prims"] == 50000
    assert np.abs(one - two).max() <= 2e-4 * one.max()


@pytest.mark.parametrize("scene_id,name", [(0, "random_scene"), (5, "cornell_box"), (7, "final_scene")])
def test_cpp_mirror_main_matches_python_path(pkg, gpu, scene_id, name, tmp_path):
    """The replacement main (host/rtw_main.cpp: reference constructors -> flatten -> rtw_render -> write_color -> P3)
    prints the same image as the ctypes path, byte for byte up to f32 summation order (<= 1 LSB on a few bytes)."""
    import subprocess
    exe = os.path.join(ROOT, "rust-ray-tracing-in-a-weekend_b200", "host", "rtw_main")
    earth = tmp_path / "earth.rgb"
    earth.write_bytes(pkg.scenes.earth_texels().tobytes())
    W, H, spp = 96, 64, 16
    r = subprocess.run([exe, "--scene", str(scene_id), "--width", str(W), "--height", str(H), "--spp", str(spp), "--earth", str(earth)],
                       capture_output=True, text=True, check=True)
    assert r.stdout.startswith(f"P3\n{W} {H}\n255\n\n")                  # src/main.rs:472 header, byte-compatible
    vals = np.array(r.stdout.split()[4:], dtype=np.int64).reshape(H, W, 3)
    sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    img, _ = sc.render(spec.camera(gpu, W, H), pkg.make_params(W, H, spp, background=spec.background, seed=1))
    import ctypes as C
    out = np.zeros(W * H * 3, np.uint8)
    gpu.check(gpu.f("write_color")(img.ctypes.data_as(C.POINTER(C.c_float)), W * H, spp, out.ctypes.data_as(C.POINTER(C.c_uint8))))
    diff = np.abs(vals - out.reshape(H, W, 3).astype(np.int64))
    assert diff.max() <= 1 and np.mean(diff > 0) < 0.01
    # the same render in 4 passes with progress lines on stderr, and a PNG next to the P3 text
    import struct, zlib
    png = tmp_path / "o.png"
    r2 = subprocess.run([exe, "--scene", str(scene_id), "--width", str(W), "--height", str(H), "--spp", str(spp), "--earth", str(earth),
                         "--passes", "4", "--png", str(png)], capture_output=True, text=True, check=True)
    assert "Progress: 4/16 samples" in r2.stderr and "Progress: 16/16 samples" in r2.stderr
    vals2 = np.array(r2.stdout.split()[4:], dtype=np.int64).reshape(H, W, 3)
    assert np.abs(vals2 - vals).max() <= 1 and np.mean(vals2 != vals) < 0.01
    raw = png.read_bytes()
    n_idat, = struct.unpack(">I", raw[33:37])
    lines = np.frombuffer(zlib.decompress(raw[41:41 + n_idat]), np.uint8).reshape(H, 1 + 3 * W)
    assert np.array_equal(lines[:, 1:].reshape(H, W, 3), vals2)


def test_sweep_scene_parity_and_bulk_api(pkg, gpu, orc):
    """BASELINE config 5 generator at test size (20k spheres through rtw_sphere_batch): world BVH hit parity and
    path parity against the oracle's flat world list over the same spheres."""
    a, b = pkg.Scene(gpu), pkg.Scene(orc)
    spec = pkg.scenes.sweep_scene(a, 20000, seed=3)
    pkg.scenes.sweep_scene(b, 20000, seed=3, wrap_bvh=False)
    a.commit(1, 0)
    d = a.debug_flatten()
    assert d["prims"] == 20000 and d["nodes"] > (5000 if d["bvh_width"] == 2 else 2500)
    rs = np.random.RandomState(5)
    W, H, n = 96, 54, 6000
    px, py, sm = rs.randint(0, W, n), rs.randint(0, H, n), rs.randint(0, 64, n)
    p = pkg.make_params(W, H, 64, background=spec.background, seed=2)
    ra, sa = a.trace_paths(spec.camera(gpu, W, H), p, px, py, sm)
    rb, sb = b.trace_paths(spec.camera(orc, W, H), p, px, py, sm)      # flat list of 20k spheres: ~1 s
    good = (sa == sb) & (np.abs(ra - rb).max(1) <= 1e-3 * np.maximum(1.0, np.abs(rb).max(1)))
    # far camera (|o| ~ 2500) and r = 8.4 spheres: a hit point stored in f32 is known to 3e-5, i.e. 4e-6 of a radius, and
    # every specular bounce multiplies that by distance/radius — paths decorrelate after a few bounces (5 rays/path here)
    record("paths_sweep20k", match=good.mean(), seg_match=np.mean(sa[:2000] == sb[:2000]))
    assert good.mean() >= 0.90, good.mean()                                     # measured 0.915 / 0.9185
    assert np.mean(sa[:2000] == sb[:2000]) > 0.90
    se = rb.std(0) / math.sqrt(n) * math.sqrt(2 * (1 - good.mean()))
    assert (np.abs(ra.mean(0) - rb.mean(0)) <= 5 * se + 1e-4).all()


def test_aabb_axis_parallel_rays(pkg, gpu, orc):
    """AABB::hit (src/aabb.rs:77-103) with direction components of exactly +-0: 1/d is +-inf in the reference, which
    its min/max handle; the f32 slab test replaces such a component by +-1e-30 and must never cull a box the reference
    hits — in particular boxes that STRADDLE 0 on that axis while the origin is inside the slab (inf - inf = NaN before
    the fix)."""
    rs = np.random.RandomState(17)
    n = 60000
    bmin = f32(rs.uniform(-3, 0, (n, 3))); bmax = f32(bmin + rs.uniform(0.1, 4, (n, 3)))
    o = f32(rs.uniform(-4, 4, (n, 3)))
    d = f32(rs.randn(n, 3))
    zero = rs.randint(0, 7, n)                       # bit mask of components forced to +-0 (never all three)
    for k in range(3):
        m = (zero >> k) & 1 == 1
        d[m, k] = np.where(rs.rand(m.sum()) < 0.5, 0.0, -0.0)
    d[np.all(d == 0, axis=1), 0] = 1.0
    with np.errstate(divide="ignore", invalid="ignore"):
        hb = orc.test_aabb(bmin, bmax, o, d, 0.001, 1e30)
    ha = gpu.test_aabb(bmin, bmax, o, d, 0.001, 1e30)
    assert hb.mean() > 0.05
    assert (ha[hb == 1] == 1).all(), np.mean(ha[hb == 1] == 0)          # conservative: every reference hit is kept
    assert np.mean(ha != hb) < 2e-3                                      # and (almost) nothing else gets through
    # the advisor's case: o = 0.5, d = 0 on x, box [-1, 1]
    ha = gpu.test_aabb([[-1, -1, -1]], [[1, 1, 1]], [[0.5, 0.0, -5.0]], [[0.0, 0.0, 1.0]], 0.001, 1e30)
    assert ha[0] == 1


def test_world_hit_axis_parallel_rays(pkg, gpu, orc):
    """The same through the whole traversal: axis-parallel rays into cornell_box (rects, rotated boxes) and the book-1
    scene hit what the oracle hits."""
    for name in ("cornell_box", "random_scene"):
        a, b, spec = build_both(pkg, gpu, orc, name)
        rs = np.random.RandomState(23)
        n = 20000
        lo, hi = (np.array([1.0, 1.0, -700.0]), np.array([554.0, 554.0, 554.0])) if name == "cornell_box" else (np.array([-11.0, 0.05, -11.0]), np.array([11.0, 3.0, 11.0]))
        o = f32(rs.uniform(lo, hi, (n, 3)))
        axis = rs.randint(0, 3, n)
        d = np.zeros((n, 3)); d[np.arange(n), axis] = np.where(rs.rand(n) < 0.5, 1.0, -1.0) * rs.uniform(0.5, 2.0, n)
        d = f32(d)
        hb, ok = stable(lambda oo, dd, tt: b.test_hit(-1, oo, dd, tt), o, d, np.zeros(n), float(np.abs(hi).max()))
        # (the perturbation inside `stable` keeps the zero components zero: it is multiplicative on d)
        ha = a.test_hit(-1, o, d, np.zeros(n))
        compare_hits(ha, hb, ok, float(np.abs(hi).max()), np.linalg.norm(d, axis=1), min_stable=0.985, check_uv=False)   # measured 0.998 / 0.995


def test_constant_medium_with_sphere_list_boundary(pkg, gpu, orc):
    """hit_constant_medium (src/hittable.rs:417-473) whose boundary is a BvhNode of spheres: the second probe
    boundary.hit(t1 + 0.0001, inf) must be able to return a sphere's FAR crossing."""
    out = []
    for lib in (gpu, orc):
        sc = pkg.Scene(lib)
        iso = sc.isotropic(sc.tex_solid((0.9, 0.9, 0.9)))
        dummy = sc.lambertian(sc.tex_solid((0.5, 0.5, 0.5)))
        members = [sc.sphere(dummy, (0.0, 0.0, 0.0), 2.0), sc.sphere(dummy, (5.0, 0.5, 0.0), 1.5), sc.sphere(dummy, (-4.0, 0.0, 1.0), 1.0)]
        sc.push(sc.constant_medium(sc.bvh_node(members, 0.0, 1.0), 0.8, iso))
        if lib is gpu:
            sc.commit(1, 0)
        else:
            sc.set_media_deferred(True)
        out.append(sc)
    a, b = out
    rs = np.random.RandomState(31)
    n = 80000
    o = f32(rs.uniform(-9, 9, (n, 3)))
    tgt = rs.uniform(-5, 6, (n, 3)) * np.array([1.0, 0.3, 0.3])
    d = f32((tgt - o) * rs.uniform(0.1, 1.5, (n, 1)))
    xi = q24(rs, (n, 4)); xi[xi == 0] = 0.5
    hb, ok = stable(lambda oo, dd, tt: b.test_hit(-1, oo, dd, tt, xi=xi), o, d, np.zeros(n), 10.0)
    ha = a.test_hit(-1, o, d, np.zeros(n), xi=xi)
    assert hb["hit"].mean() > 0.1                       # the media really scatter
    assert np.array_equal(ha["ndraw"][ok], hb["ndraw"][ok])
    compare_hits(ha, hb, ok, 10.0, np.linalg.norm(d, axis=1), check_uv=False)


def test_shutter_outside_moving_sphere_interval_is_refused(pkg, gpu):
    """MovingSphere boxes in the BVH span the sphere's own [time0, time1] (like the reference's BvhNode boxes,
    src/hittable.rs:480-482); a camera shutter beyond it is refused instead of silently culling."""
    sc = pkg.Scene(gpu)
    m = sc.lambertian(sc.tex_solid((0.5, 0.5, 0.5)))
    sc.push(sc.moving_sphere(m, (0, 0, 0), (0, 1, 0), 0.0, 1.0, 0.5))
    sc.commit(1, 0)
    p = pkg.make_params(16, 16, 1)
    ok_cam = gpu.camera_new((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 5.0, 0.25, 0.75)
    sc.render(ok_cam, p)
    bad_cam = gpu.camera_new((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 5.0, 0.0, 2.0)
    with pytest.raises(pkg.RtwError) as e:
        sc.render(bad_cam, p)
    assert e.value.code == -1


def test_concurrent_renders_with_different_seeds(pkg, gpu):
    """The Philox key schedule travels in the kernel parameters: two host threads rendering two scene handles with
    different seeds at the same time get exactly the images they get one after the other (no shared launch state)."""
    import threading
    jobs = []
    for seed in (11, 22):
        sc, spec = pkg.scenes.build(gpu, "random_scene")
        sc.commit(1, 0)
        jobs.append((sc, spec.camera(gpu, 96, 64), pkg.make_params(96, 64, 64, background=spec.background, seed=seed)))
    serial = [sc.render(cam, p)[0].copy() for sc, cam, p in jobs]
    for _ in range(3):
        res = [None, None]
        def run(i):
            res[i] = jobs[i][0].render(jobs[i][1], jobs[i][2])[0].copy()
        th = [threading.Thread(target=run, args=(i,)) for i in range(2)]
        [t.start() for t in th]; [t.join() for t in th]
        for i in range(2):
            assert np.abs(res[i] - serial[i]).max() <= 2e-4 * np.abs(serial[i]).max()
    assert not np.array_equal(serial[0], serial[1])


def test_staged_upload_of_a_big_scene(pkg, gpu, monkeypatch):
    """SURVEY 8f-2 "commit overlap": from 32 MB on, the rtw_sphere_batch spheres reach the device through pinned staging
    buffers filled by several host threads (csrc/bvh_build.cu staged_upload) instead of one pageable cudaMemcpy.  1 Mi + 3
    spheres (40 MB, a ragged last chunk): the scene committed that way returns exactly the hits of the one committed with
    the plain copy (RTW_STAGED_UPLOAD=0), on primary and on secondary rays, twice in a row (the buffers are reused)."""
    import ctypes
    cudart = ctypes.CDLL("libcudart.so.12")                # (already in the process: librtw.so links it)
    def free_bytes():
        fr, tot = ctypes.c_size_t(0), ctypes.c_size_t(0)
        assert cudart.cudaMemGetInfo(ctypes.byref(fr), ctypes.byref(tot)) == 0
        return fr.value
    n_sph = (1 << 20) + 3
    warm = pkg.Scene(gpu); pkg.scenes.sweep_scene(warm, 1000, seed=1); warm.commit(1, 0); warm.close()      # context, module load
    free0 = free_bytes()
    scenes = []
    for sw in ("0", "1", "1"):
        monkeypatch.setenv("RTW_STAGED_UPLOAD", sw)
        sc = pkg.Scene(gpu)
        spec = pkg.scenes.sweep_scene(sc, n_sph, seed=11)
        sc.commit(1, 0)
        scenes.append(sc)
    rs = np.random.RandomState(78)
    n = 50000
    g = camera_rays(pkg, gpu, spec, n, rs)
    xi = q24(rs, (n, 4)); xi[xi == 0] = 0.5
    o, d, tm = f32(g["origin"]), f32(g["dir"]), f32(g["time"])
    hs = [sc.test_hit(-1, o, d, tm, xi=xi) for sc in scenes]
    hit = hs[0]["hit"] == 1
    assert hit.mean() > 0.3
    o2 = f32(hs[0]["p"][hit]); v = rs.randn(hit.sum(), 3); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d2 = f32(hs[0]["normal"][hit] + 1.001 * v)
    hs2 = [sc.test_hit(-1, o2, d2, tm[hit], xi=xi[hit]) for sc in scenes]
    for group in (hs, hs2):
        x = group[0]
        for y in group[1:]:
            same = (x["hit"] == y["hit"]) & (x["mat"] == y["mat"]) & (x["t"] == y["t"]) & np.all(x["normal"] == y["normal"], axis=1)
            assert same.mean() >= 0.9999, same.mean()      # (exact ties between two spheres may resolve differently: node order comes from atomics)
    # the builder's scratch pool (csrc/bvh_build.cu) is kept across commits and trimmed when a scene is freed
    held = free0 - free_bytes()
    for sc in scenes: sc.close()
    assert free0 - free_bytes() <= 64 << 20, (free0, free_bytes(), held)


@pytest.mark.parametrize("width", [2, 8])
@pytest.mark.parametrize("name", ["random_scene", "final_scene", "cornell_box_smoke", "sweep_1", "sweep_2", "sweep_3000", "sweep_70000"])
def test_device_bvh_build_equals_host_build(pkg, gpu, monkeypatch, name, width):
    """the BVH built ON THE DEVICE (csrc/bvh_build.cu: Morton LBVH, refit, collapse, records in leaf order;
    replaces new_bvh_node, src/hittable.rs:77-130) returns the closest hits of the host-built SAH tree — any valid BVH
    over the same primitives does (src/hittable.rs:290-306 only prunes).  Both node formats; scene-graph primitives
    (rects, instances, media boundaries) and rtw_sphere_batch spheres; 1- and 2-primitive corner cases."""
    monkeypatch.setenv("RTW_BVH", str(width))
    scenes = []
    for dev in ("0", "1"):
        monkeypatch.setenv("RTW_DEVICE_BUILD", dev)
        if name.startswith("sweep_"):
            sc = pkg.Scene(gpu)
            spec = pkg.scenes.sweep_scene(sc, int(name[6:]), seed=9)
        else:
            sc, spec = pkg.scenes.build(gpu, name)
        sc.commit(1, 0)
        scenes.append(sc)
    a, b = scenes
    rs = np.random.RandomState(77)
    n = 100000
    g = camera_rays(pkg, gpu, spec, n, rs)
    xi = q24(rs, (n, 4)); xi[xi == 0] = 0.5
    o, d, tm = f32(g["origin"]), f32(g["dir"]), f32(g["time"])
    ha, hb = a.test_hit(-1, o, d, tm, xi=xi), b.test_hit(-1, o, d, tm, xi=xi)
    hit = ha["hit"] == 1
    o2 = f32(ha["p"][hit]); v = rs.randn(hit.sum(), 3); v /= np.linalg.norm(v, axis=1, keepdims=True)
    d2 = f32(ha["normal"][hit] + 1.001 * v)
    ha2, hb2 = a.test_hit(-1, o2, d2, tm[hit], xi=xi[hit]), b.test_hit(-1, o2, d2, tm[hit], xi=xi[hit])
    for x, y in ((ha, hb), (ha2, hb2)):
        same = (x["hit"] == y["hit"]) & (x["mat"] == y["mat"]) & (x["front"] == y["front"]) & (x["t"] == y["t"]) & np.all(x["normal"] == y["normal"], axis=1)
        record("device_build_equiv", name=name, width=width, same=same.mean(), hit=x["hit"].mean())
        assert same.mean() >= 0.9999, same.mean()          # only exact ties between two primitives may resolve differently
        assert np.array_equal(x["ndraw"], y["ndraw"])
    # and the rendered image: same paths, same sums (work units are cut the same way)
    W, H = 96, 54
    p = pkg.make_params(W, H, 16, background=spec.background, seed=4)
    ia, sa = a.render(spec.camera(gpu, W, H), p)
    ib, sb = b.render(spec.camera(gpu, W, H), p)
    assert sa["rays"] == sb["rays"] or abs(sa["rays"] - sb["rays"]) <= 1e-4 * sa["rays"]
    assert np.abs(ia - ib).max() <= 1e-3 * max(1.0, np.abs(ia).max())
    assert sb["n_nodes"] > 0 and sb["n_prims"] == sa["n_prims"]


@pytest.mark.parametrize("width", [2, 8])
@pytest.mark.parametrize("name", list(PATH_BARS) + ["sweep_30000"])
def test_wavefront_pipeline_equals_megakernel(pkg, gpu, monkeypatch, name, width):
    """The wavefront pipeline (csrc/rtw_wavefront.cuh: path pool in HBM, logic kernel = shade + regenerate, persistent trace
    kernel with dynamic ray fetch) schedules the SAME per-(pixel, sample) paths as the megakernel — same Philox
    coordinates, same device functions for ray_color's stages (src/main.rs:19-38) — so the images agree up to f32 summation
    order and the ray counts are equal.  Both node formats, every scene, ragged image sizes, a pool smaller than the image
    (several regeneration rounds) and a progressive pass with a sample offset."""
    monkeypatch.setenv("RTW_BVH", str(width))
    monkeypatch.setenv("RTW_WF_POOL", "20000")          # far fewer slots than paths: regeneration is exercised
    if name.startswith("sweep_"):
        sc = pkg.Scene(gpu)
        spec = pkg.scenes.sweep_scene(sc, int(name[6:]), seed=4)
    else:
        sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    W, H, spp = 101, 67, 24                             # ragged tiles on both axes
    cam = spec.camera(gpu, W, H)
    mega, sm = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=6, flags=pkg.api.RTW_FLAG_KERNEL_MEGA))
    wave, sw = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=6, flags=pkg.api.RTW_FLAG_KERNEL_WAVEFRONT))
    assert sw["rays"] == sm["rays"] and sw["paths"] == W * H * spp
    assert sum(sw["units_per_device"]) == W * H * spp                       # (wavefront: paths started per device)
    assert np.isfinite(wave).all()
    assert np.abs(wave - mega).max() <= 2e-4 * max(1.0, np.abs(mega).max())
    # max_depth 0 / 1 corner cases
    for depth in (0, 1):
        a, _ = sc.render(cam, pkg.make_params(W, H, 4, max_depth=depth, background=spec.background, flags=pkg.api.RTW_FLAG_KERNEL_MEGA))
        b, _ = sc.render(cam, pkg.make_params(W, H, 4, max_depth=depth, background=spec.background, flags=pkg.api.RTW_FLAG_KERNEL_WAVEFRONT))
        assert np.abs(a - b).max() <= 2e-4 * max(1.0, np.abs(a).max())


def test_wavefront_progressive_and_two_gpus(pkg, gpu, monkeypatch):
    """Progressive passes (absolute sample index = Philox coordinate) and, when a second GPU is there, path chunks taken by
    two GPUs from the one shared counter: the image is the one-shot single-GPU image."""
    monkeypatch.setenv("RTW_WF_POOL", "50000")
    sc, spec = pkg.scenes.build(gpu, "random_scene")
    n = 2 if gpu.f("device_count")() >= 2 else 1
    sc.commit(n, 0)
    W, H, spp = 320, 200, 48
    cam = spec.camera(gpu, W, H)
    F = pkg.api.RTW_FLAG_KERNEL_WAVEFRONT
    one, s1 = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=3, n_gpus=1, flags=F))
    prog, s2 = sc.render_progressive(cam, pkg.make_params(W, H, spp, background=spec.background, seed=3, n_gpus=1, flags=F), samples_per_pass=20)
    assert np.abs(prog - one).max() <= 2e-4 * one.max() and s2["paths"] == W * H * spp
    if n == 2:
        monkeypatch.setenv("RTW_WF_POOL", "400000")
        W, H, spp = 1200, 800, 16                      # 15.4 M paths = 15 chunks of 2^20: both GPUs get some
        cam = spec.camera(gpu, W, H)
        a, sa = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=3, n_gpus=1, flags=F))
        b, sb = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=3, n_gpus=2, flags=F))
        assert sa["rays"] == sb["rays"] and min(sb["units_per_device"][:2]) > 0
        assert np.abs(a - b).max() <= 2e-4 * a.max()


def _box_mix_scene(pkg, lib):
    """Boxes in every role the reference gives them: plain, translated, rotated + translated (src/main.rs:198-210), inside a
    BvhNode (:399-407), with a dielectric material (rays travel INSIDE the box and leave through a face) and as the boundary
    of a ConstantMedium (those stay six rects)."""
    sc = pkg.Scene(lib)
    white = sc.lambertian(sc.tex_solid((0.73, 0.73, 0.73))); red = sc.lambertian(sc.tex_solid((0.65, 0.05, 0.05)))
    glass = sc.dielectric(1.5); metal = sc.metal((0.8, 0.85, 0.88), 0.1); light = sc.diffuse_light(sc.tex_solid((7, 7, 7)))
    iso = sc.isotropic(sc.tex_solid((0.9, 0.9, 0.9)))
    sc.push(sc.xz_rect(white, -400, 400, -400, 400, 0.0))
    sc.push(sc.xz_rect(light, -150, 150, -150, 150, 500.0))
    sc.push(sc.box((-250, 0, -80), (-150, 120, 20), red))
    sc.push(sc.translate(sc.box((0, 0, 0), (90, 60, 90), glass), (-60.0, 0.001, -200.0)))
    sc.push(sc.translate(sc.rotate_y(25.0, sc.box((0, 0, 0), (100, 200, 100), white)), (120.0, 0.0, 40.0)))
    sc.push(sc.translate(sc.rotate_y(-18.0, sc.box((0, 0, 0), (80, 80, 80), metal)), (-40.0, 0.0, 90.0)))
    rs = np.random.RandomState(5)
    members = []
    for i in range(9):
        lo = np.array([rs.uniform(-300, 250), 0.0, rs.uniform(150, 330)])
        members.append(sc.box(tuple(lo), tuple(lo + rs.uniform(20, 60, 3)), white if i % 2 else red))
    sc.push(sc.bvh_node(members, 0.0, 1.0))
    sc.push(sc.constant_medium(sc.translate(sc.rotate_y(40.0, sc.box((0, 0, 0), (70, 140, 70), white)), (230.0, 0.0, -150.0)), 0.01, iso))
    cam = lambda W, H: lib.camera_new((0.0, 260.0, -750.0), (0.0, 100.0, 0.0), (0.0, 1.0, 0.0), 42.0, W / H, 0.0, 10.0, 0.0, 1.0)
    return sc, cam


@pytest.mark.parametrize("name", ["cornell_box", "final_scene", "box_mix"])
@pytest.mark.parametrize("width", [2, 8])
def test_box_as_one_leaf_equals_six_rects(pkg, gpu, monkeypatch, name, width):
    """A surface Box is ONE BVH leaf (PRIM_BOX: a slab test in object space gives the entry / exit face, whose rect record
    then describes the hit) instead of the six rects of new_box (src/hittable.rs:132-145) as six leaves.  Both forms compute
    the rects' own plane distances t = (k - o_k) / d_k, so the same paths come out — except where a ray crosses a box
    EDGE within rounding (rect: closed (a, b) interval on the hit point; slab: comparison of plane distances).  Path by
    path: >= 99.9 % bit-equal, the rest statistically irrelevant; equal ray counts to 1e-3; images agree in the mean."""
    monkeypatch.setenv("RTW_BVH", str(width))
    W, H, spp = 96, 96, 16
    res = {}
    for form in ("1", "0"):
        monkeypatch.setenv("RTW_BOX_PRIM", form)
        if name == "box_mix":
            sc, camf = _box_mix_scene(pkg, gpu); cam = camf(W, H); bg = (0.05, 0.05, 0.08)
        else:
            sc, spec = pkg.scenes.build(gpu, name); cam = spec.camera(gpu, W, H); bg = spec.background
        d = sc.debug_flatten()
        sc.commit(1, 0)
        p = pkg.make_params(W, H, spp, background=bg, seed=9, flags=pkg.api.RTW_FLAG_KERNEL_MEGA)
        img, st = sc.render(cam, p)
        ys, xs, ss = np.meshgrid(np.arange(H), np.arange(W), np.arange(spp), indexing="ij")
        rgb, seg = sc.trace_paths(cam, p, xs.ravel(), ys.ravel(), ss.ravel())
        res[form] = (img, st, rgb, seg, d)
        sc.close()
    (ia, sa, ra, ga, da), (ib, sb, rb, gb, db) = res["1"], res["0"]
    assert da["bvh_prims"] < db["bvh_prims"] and da["prims"] > db["prims"]            # one leaf per box + its six face records
    same = np.mean((ra == rb).all(1) & (ga == gb))
    assert same >= 0.999, same
    assert abs(sa["rays"] - sb["rays"]) <= 1e-3 * sb["rays"]
    assert np.isfinite(ia).all() and abs(ia.mean() - ib.mean()) <= 0.01 * ib.mean()
    # the render kernel traced what trace_paths traced (tile lists, ring, self-intersection rule with face indices)
    ref = ra.reshape(H, W, spp, 3).sum(2)[::-1]
    assert np.abs(ia - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())


@pytest.mark.parametrize("name", ["random_scene", "cornell_box", "two_perlin_spheres", "final_scene"])
def test_tiny_work_units_carry_their_last_rays_over(pkg, gpu, name):
    """Work units of ONE or TWO samples per tile: every unit ends with most of its paths still in flight, so the megakernel
    constantly carries the last rays of a unit (lanes + ring) into the next one as orphans of the old tile (csrc/rtw_api.cu
    render_kernel, RTW_CARRY; the variants with media code drain the old way).  Scheduling must not show in the result: the
    same paths as with one unit per tile — equal ray counts, images equal up to f32 summation order — and the image is the
    per-pixel sum of trace_paths' paths (the reference's accumulation, src/main.rs:519-520, :542-547)."""
    sc, spec = pkg.scenes.build(gpu, name)
    sc.commit(1, 0)
    W, H, spp = 203, 117, 12                                 # ragged tiles on both axes
    cam = spec.camera(gpu, W, H)
    F = pkg.api.RTW_FLAG_KERNEL_MEGA
    one, s1 = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=21, samples_per_unit=spp, flags=F))
    for spu in (1, 2, 5):
        img, st = sc.render(cam, pkg.make_params(W, H, spp, background=spec.background, seed=21, samples_per_unit=spu, flags=F))
        assert st["rays"] == s1["rays"] and st["paths"] == W * H * spp, (name, spu)
        assert np.isfinite(img).all()
        assert np.abs(img - one).max() <= 2e-4 * max(1.0, np.abs(one).max()), (name, spu)
    p = pkg.make_params(W, H, spp, background=spec.background, seed=21, flags=F)
    ys, xs, ss = np.meshgrid(np.arange(H), np.arange(W), np.arange(spp), indexing="ij")
    rgb, _ = sc.trace_paths(cam, p, xs.ravel(), ys.ravel(), ss.ravel())
    ref = rgb.reshape(H, W, spp, 3).sum(2)[::-1]
    assert np.abs(one - ref).max() <= 2e-4 * max(1.0, np.abs(ref).max())
