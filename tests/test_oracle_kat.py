"""Pin the CPU oracle (oracle/oracle.cpp) before trusting it.

The reference has no tests; its only in-repo known answers are the sphere_uv table in
src/math.rs:292-294.  Everything else here is the reference FORMULA re-evaluated independently in numpy
(different code shape than the C++ restatement) or a published known answer (Random123 Philox KATs).
"""
import math
import os

import numpy as np
import pytest

from conftest import q24

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sphere_uv_reference_table(orc):
    # src/math.rs:292-294.  <-1 0 0> needs -p.z = -0.0 to give u = 0 (with +0.0 the formula gives 1.0)
    pts = np.array([[1, 0, 0], [-1, 0, 0.0], [0, 1, 0], [0, -1, 0], [0, 0, 1], [0, 0, -1]], float)
    want = np.array([[0.5, 0.5], [1.0, 0.5], [0.5, 1.0], [0.5, 0.0], [0.25, 0.5], [0.75, 0.5]])
    u, v = np.zeros(6), np.zeros(6)
    import ctypes as C
    dp = C.POINTER(C.c_double)
    orc.f("sphere_uv")(6, pts.ctypes.data_as(dp), u.ctypes.data_as(dp), v.ctypes.data_as(dp))
    got = np.stack([u, v], 1)
    # u of (-1,0,0) is 0 or 1 depending on the sign of zero: both are the same point on the seam
    got[1, 0] = got[1, 0] % 1.0
    want[1, 0] = 0.0
    assert np.allclose(got, want, atol=1e-15)


def test_philox_random123_known_answers(orc):
    # Random123 kat_vectors, philox4x32-10
    kats = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
            ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
            ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0],
             [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
    for ctr, key, want in kats:
        assert list(orc.philox(ctr, key)[0]) == want


def _camera_numpy(look_from, look_at, vup, vfov, aspect, aperture, focus):
    # src/camera.rs:18-56 re-evaluated in numpy
    look_from, look_at, vup = map(lambda a: np.array(a, float), (look_from, look_at, vup))
    h = math.tan(math.radians(vfov) / 2)
    vh, vw = 2 * h, aspect * 2 * h
    w = look_from - look_at
    w /= np.linalg.norm(w)
    u = np.cross(vup, w)
    u /= np.linalg.norm(u)
    v = np.cross(w, u)
    hor, ver = focus * vw * u, focus * vh * v
    return dict(origin=look_from, horizontal=hor, vertical=ver, llc=look_from - hor / 2 - ver / 2 - focus * w, u=u, v=v, w=w,
                lens=aperture / 2)


@pytest.mark.parametrize("args", [((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0),
                                  ((278, 278, -800), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.1, 10.0),
                                  ((478, 278, -600), (278, 278, 0), (0, 1, 0), 40.0, 1.0, 0.0, 10.0)])
def test_camera_new(orc, rtw, args):
    want = _camera_numpy(*args)
    for lib in (orc, rtw):      # Camera::new is host arithmetic in both libraries
        cam = lib.camera_new(*args)
        assert np.allclose(list(cam.horizontal), want["horizontal"], rtol=1e-14, atol=1e-13)
        assert np.allclose(list(cam.vertical), want["vertical"], rtol=1e-14, atol=1e-13)
        assert np.allclose(list(cam.lower_left_corner), want["llc"], rtol=1e-14, atol=1e-12)
        assert np.allclose(list(cam.u), want["u"], atol=1e-15) and np.allclose(list(cam.w), want["w"], atol=1e-15)
        assert cam.lens_radius == want["lens"] and (cam.time0, cam.time1) == (0.0, 1.0)
    # SURVEY §8c spot value
    cam = orc.camera_new(*((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0))
    assert abs(cam.horizontal[0] - 1.189463936993608) < 1e-14 and abs(cam.lower_left_corner[2] - 3.4122032022021487) < 1e-13


def test_reflectance_refract_reflect(orc):
    import ctypes as C
    dp = C.POINTER(C.c_double)
    cos = np.array([0.5, 1.0, 0.1]); idx = np.array([1 / 1.5, 1.5, 1.5]); out = np.zeros(3)
    orc.f("reflectance")(3, cos.ctypes.data_as(dp), idx.ctypes.data_as(dp), out.ctypes.data_as(dp))
    r0 = ((1 - idx) / (1 + idx)) ** 2
    assert np.allclose(out, r0 + (1 - r0) * (1 - cos) ** 5, rtol=1e-14)      # src/material.rs:89-94
    assert abs(out[0] - 0.07) < 1e-15 and abs(out[1] - 0.04) < 1e-15
    uv = np.array([1, -1, 0.0]) / math.sqrt(2); n = np.array([0, 1, 0.0]); o = np.zeros(3)
    orc.f("refract")(uv.ctypes.data_as(dp), n.ctypes.data_as(dp), 1 / 1.5, o.ctypes.data_as(dp))
    assert np.allclose(o, [0.4714045207910316, -0.881917103688197, 0.0], atol=1e-15)   # src/math.rs:110-117
    v = np.array([1.0, -2.0, 0.5])
    orc.f("reflect")(v.ctypes.data_as(dp), n.ctypes.data_as(dp), o.ctypes.data_as(dp))
    assert np.allclose(o, [1.0, 2.0, 0.5])


def _one_scene(pkg, orc):
    sc = pkg.Scene(orc)
    return sc, sc.lambertian(sc.tex_solid((0.5, 0.5, 0.5)))


def test_sphere_hit_known_answers(pkg, orc):
    sc, mat = _one_scene(pkg, orc)
    ground = sc.sphere(mat, (0, -1000, 0), 1000)
    h = sc.test_hit(ground, [[13, 2, 3]], [[-13, -2.5, -3]])
    # src/hittable.rs:254-288 by hand: t = (-half_b - sqrt(disc)) / a
    o, d, c, r = np.array([13, 2, 3.0]), np.array([-13, -2.5, -3.0]), np.array([0, -1000, 0.0]), 1000.0
    oc = o - c; a = d @ d; hb = oc @ d; cc = oc @ oc - r * r
    t = (-hb - math.sqrt(hb * hb - a * cc)) / a
    assert h["hit"][0] == 1 and abs(h["t"][0] - t) < 1e-12 and abs(t - 0.8014040785686217) < 1e-12
    assert np.allclose(h["p"][0], o + t * d, atol=1e-12) and h["front"][0] == 1
    assert np.allclose(h["normal"][0], (o + t * d - c) / r, atol=1e-12)
    assert abs(h["u"][0] - 0.4639038422446764) < 1e-9 and abs(h["v"][0] - 0.9991566051170206) < 1e-9
    # from inside a glass sphere: far root, front_face false, normal flipped
    glass = sc.sphere(mat, (0, 1, 0), 1)
    h = sc.test_hit(glass, [[0, 1, 0]], [[0.3, 0.2, -2]])
    assert h["hit"][0] == 1 and abs(h["t"][0] - 0.4920678313051229) < 1e-12 and h["front"][0] == 0
    assert np.allclose(h["normal"][0], -np.array([0.3, 0.2, -2]) * 0.4920678313051229, atol=1e-12)
    # range rules: root < t_min or t_max < root rejects; equality accepted (src/hittable.rs:268-272)
    h = sc.test_hit(ground, [[13, 2, 3]], [[-13, -2.5, -3]], t_min=0.001, t_max=t)
    assert h["hit"][0] == 1
    h = sc.test_hit(ground, [[13, 2, 3]], [[-13, -2.5, -3]], t_min=0.001, t_max=np.nextafter(t, 0))
    assert h["hit"][0] == 0


def test_moving_sphere_center(pkg, orc):
    sc, mat = _one_scene(pkg, orc)
    ms = sc.moving_sphere(mat, (0, 0, 0), (0, 2, 0), 0.0, 1.0, 0.5)
    for tm in (0.0, 0.25, 1.0):
        h = sc.test_hit(ms, [[0, 2 * tm, -5]], [[0, 0, 1]], time=[tm])
        assert h["hit"][0] == 1 and abs(h["t"][0] - 4.5) < 1e-12     # src/hittable.rs:556-558
    bb_min, bb_max = np.zeros(3), np.zeros(3)
    import ctypes as C
    dp = C.POINTER(C.c_double)
    # the MovingSphere arm ignores the time arguments (src/hittable.rs:480-482)
    orc.f("bounding_box")(sc.h, ms, 0.0, 0.0, bb_min.ctypes.data_as(dp), bb_max.ctypes.data_as(dp))
    assert np.allclose(bb_min, [-0.5, -0.5, -0.5]) and np.allclose(bb_max, [0.5, 2.5, 0.5])


def test_aabb_known_answers(orc):
    mn, mx = [[-1, -1, -1]] * 3, [[1, 1, 1]] * 3
    o = [[0, 0, -5]] * 3
    d = [[0, 0, 1], [0, 0.3, 1], [0, 0, 1]]
    assert list(orc.test_aabb(mn[:2], mx[:2], o[:2], d[:2], 0.001, float("inf"))) == [1, 0]
    assert list(orc.test_aabb(mn[2:], mx[2:], o[2:], d[2:], 0.001, 4.0)) == [0]      # max <= min fails (src/aabb.rs:97)


def test_rects_and_box(pkg, orc):
    sc, mat = _one_scene(pkg, orc)
    xy = sc.xy_rect(mat, 3, 5, 1, 3, -2)
    h = sc.test_hit(xy, [[4, 2, 5]], [[0, 0, -1]])
    assert h["hit"][0] == 1 and h["t"][0] == 7 and h["front"][0] == 1 and list(h["normal"][0]) == [0, 0, 1]
    assert (h["u"][0], h["v"][0]) == (0.5, 0.5)                                        # src/hittable.rs:322-323
    h = sc.test_hit(xy, [[5, 3, 5]], [[0, 0, -1]])                                     # inclusive bounds (:318)
    assert h["hit"][0] == 1
    h = sc.test_hit(xy, [[5.0001, 3, 5]], [[0, 0, -1]])
    assert h["hit"][0] == 0
    xz = sc.xz_rect(mat, 0, 2, 0, 4, 1)
    h = sc.test_hit(xz, [[0.5, 3, 1]], [[0, -1, 0]])
    assert h["hit"][0] == 1 and h["front"][0] == 1 and list(h["normal"][0]) == [0, 1, 0] and (h["u"][0], h["v"][0]) == (0.25, 0.25)
    yz = sc.yz_rect(mat, 0, 2, 0, 4, 1)
    h = sc.test_hit(yz, [[-3, 0.5, 1]], [[1, 0, 0]])
    assert h["hit"][0] == 1 and h["front"][0] == 0 and list(h["normal"][0]) == [-1, 0, 0] and (h["u"][0], h["v"][0]) == (0.25, 0.25)
    box = sc.box((0, 0, 0), (1, 2, 3), mat)
    h = sc.test_hit(box, [[0.5, 1, -4]], [[0, 0, 1]])
    assert h["hit"][0] == 1 and h["t"][0] == 4 and list(h["normal"][0]) == [0, 0, -1]  # closest of six (src/hittable.rs:229-231)
    # parallel ray: t = +-inf or NaN is rejected by the comparisons (src/hittable.rs:309-313)
    h = sc.test_hit(xy, [[4, 2, 5]], [[1, 0, 0]])
    assert h["hit"][0] == 0


def test_rotate_translate_quirk(pkg, orc):
    """src/hittable.rs:409 tests the object-space ray against the world-space normal; the Translate wrapper
    (:238) re-faces the normal against the world ray.  Net: normal always opposes the ray, front_face may lie."""
    sc, mat = _one_scene(pkg, orc)
    rs = np.random.RandomState(5)
    box = sc.box((0, 0, 0), (165, 330, 165), mat)
    rot = sc.rotate_y(15.0, box)
    inst = sc.translate(rot, (265, 0, 295))
    n = 4000
    o = np.tile([278.0, 278.0, -800.0], (n, 1)) + rs.randn(n, 3) * 50
    tgt = np.stack([rs.uniform(265, 430, n), rs.uniform(0, 330, n), rs.uniform(295, 460, n)], 1)
    d = tgt - o
    h = sc.test_hit(inst, o, d)
    hit = h["hit"] == 1
    assert hit.mean() > 0.5
    assert (np.einsum("ij,ij->i", h["normal"][hit], d[hit]) < 0).all()
    # RotateY alone: the wrongly-faced normals are observable
    h2 = sc.test_hit(rot, o - np.array([265, 0, 295.0]), d)
    assert np.array_equal(h2["hit"], h["hit"]) and np.allclose(h2["t"][hit], h["t"][hit], rtol=1e-9)
    c, s = math.cos(math.radians(15)), math.sin(math.radians(15))
    d_obj = np.stack([c * d[:, 0] - s * d[:, 2], d[:, 1], s * d[:, 0] + c * d[:, 2]], 1)
    n_geo = h["normal"][hit]          # true face-forward world normal
    lie = np.einsum("ij,ij->i", d_obj[hit], n_geo) >= 0        # cases where :409 picks the wrong side
    assert np.allclose(h2["normal"][hit][~lie], n_geo[~lie], atol=1e-12)
    assert np.allclose(h2["normal"][hit][lie], -n_geo[lie], atol=1e-12)
    assert np.array_equal(h["front"][hit] == 1, ~lie)          # after Translate: front_face carries the lie


def test_constant_medium_draw_gating(pkg, orc):
    """src/hittable.rs:417-473: one draw, only when the clamped segment is non-empty."""
    sc, mat = _one_scene(pkg, orc)
    iso = sc.isotropic(sc.tex_solid((1, 1, 1)))
    sph = sc.sphere(mat, (0, 0, 0), 1.0)
    med = sc.constant_medium(sph, 2.0, iso)
    xi = [[0.5, 0.25]]
    h = sc.test_hit(med, [[0, 0, -5]], [[0, 0, 1]], xi=xi)
    want_t = 4.0 + (-1 / 2.0) * math.log(0.5) / 1.0
    assert h["hit"][0] == 1 and h["ndraw"][0] == 1 and abs(h["t"][0] - want_t) < 1e-12
    assert list(h["normal"][0]) == [1, 0, 0] and h["front"][0] == 1 and h["mat"][0] == iso
    h = sc.test_hit(med, [[0, 5, -5]], [[0, 0, 1]], xi=xi)                     # misses the boundary: no draw
    assert h["hit"][0] == 0 and h["ndraw"][0] == 0
    h = sc.test_hit(med, [[0, 0, -5]], [[0, 0, 1]], t_max=3.5, xi=xi)          # clamped segment empty: no draw
    assert h["hit"][0] == 0 and h["ndraw"][0] == 0
    h = sc.test_hit(med, [[0, 0, -5]], [[0, 0, 1]], xi=[[1e-9, 0.5]])          # free flight beyond the far side
    assert h["hit"][0] == 0 and h["ndraw"][0] == 1
    h = sc.test_hit(med, [[0, 0, 0]], [[0, 0, 2]], xi=xi)                       # origin inside; un-normalised direction
    assert h["hit"][0] == 1 and abs(h["t"][0] - (0.001 + 0.5 * math.log(2) / 2.0)) < 1e-12


def _perlin_numpy(ranvec, px, py, pz, p):
    """src/perlin.rs:32-94 in vectorised numpy (double smoothstep, once-smoothed weight vector)."""
    fl = np.floor(p)
    f = p - fl
    f = f * f * (3 - 2 * f)
    i = fl.astype(np.int64)
    ff = f * f * (3 - 2 * f)
    acc = np.zeros(len(p))
    for di in (0, 1):
        for dj in (0, 1):
            for dk in (0, 1):
                idx = px[(i[:, 0] + di) & 255] ^ py[(i[:, 1] + dj) & 255] ^ pz[(i[:, 2] + dk) & 255]
                g = ranvec[idx]
                wv = f - np.array([di, dj, dk], float)
                wt = np.where(di, ff[:, 0], 1 - ff[:, 0]) * np.where(dj, ff[:, 1], 1 - ff[:, 1]) * np.where(dk, ff[:, 2], 1 - ff[:, 2])
                acc += wt * np.einsum("ij,ij->i", g, wv)
    return acc


def test_perlin_against_numpy(pkg, orc):
    rng = pkg.scenes.HostRng(7)
    rv, px, py, pz = pkg.scenes.perlin_tables(rng)
    # src/perlin.rs:122-129 writes the INDEX `target` into p[i] (then p[target] = old p[i]): values stay in
    # 0..255 but the result is generally not a permutation
    assert ((px >= 0) & (px <= 255)).all() and len(set(px.tolist())) < 256
    assert np.allclose(np.linalg.norm(rv, axis=1), 1.0)
    sc = pkg.Scene(orc)
    tex = sc.tex_noise(rv, px, py, pz, 4.0)
    rs = np.random.RandomState(1)
    p = rs.uniform(-300, 300, (5000, 3))
    import ctypes as C
    dp = C.POINTER(C.c_double)
    noise, turb = np.zeros(len(p)), np.zeros(len(p))
    orc.f("perlin_noise")(sc.h, tex, len(p), p.ctypes.data_as(dp), noise.ctypes.data_as(dp), turb.ctypes.data_as(dp))
    assert np.allclose(noise, _perlin_numpy(rv, px, py, pz, p), atol=1e-13)
    want_turb = np.abs(sum(0.5 ** k * _perlin_numpy(rv, px, py, pz, p * 2.0 ** k) for k in range(7)))
    assert np.allclose(turb, want_turb, atol=1e-12)
    rgb = sc.test_texture(tex, None, None, p)
    assert np.allclose(rgb[:, 0], 0.5 * (1 + np.sin(4.0 * p[:, 2] + 10 * want_turb)), atol=1e-11)     # src/texture.rs:43-45


def test_checker_and_image_texture(pkg, orc):
    sc = pkg.Scene(orc)
    ck = sc.tex_checker((0.2, 0.5, 0.5), (0.9, 0.9, 0.9))
    rs = np.random.RandomState(2)
    p = rs.uniform(-20, 20, (2000, 3))
    rgb = sc.test_texture(ck, None, None, p)
    sines = np.sin(10 * p[:, 0]) * np.sin(10 * p[:, 1]) * np.sin(10 * p[:, 2])
    assert np.allclose(rgb, np.where((sines < 0)[:, None], [0.9, 0.9, 0.9], [0.2, 0.5, 0.5]))       # src/texture.rs:35-42
    img = rs.randint(0, 256, (5, 7, 3)).astype(np.uint8)
    it = sc.tex_image(img)
    u, v = rs.uniform(-0.2, 1.2, 3000), rs.uniform(-0.2, 1.2, 3000)
    rgb = sc.test_texture(it, u, v, np.zeros((3000, 3)))
    uu, vv = np.clip(u, 0, 1), 1 - np.clip(v, 0, 1)
    i, j = np.minimum((uu * 7).astype(int), 6), np.minimum((vv * 5).astype(int), 4)
    assert np.allclose(rgb, img[j, i] / 255.0, atol=1e-15)                                           # src/texture.rs:46-73
    assert np.allclose(sc.test_texture(it, [0.0], [1.0], [[0, 0, 0]])[0], img[0, 0] / 255.0)       # (u=0, v=1) = first texel


def test_scatter_draw_order(pkg, orc):
    """Lambertian consumes x,y,z per rejection iteration (src/math.rs:43-58); Metal always draws; Dielectric draws
    only if it can refract (src/material.rs:72); DiffuseLight never scatters and emits on both faces."""
    sc = pkg.Scene(orc)
    lam = sc.lambertian(sc.tex_solid((0.1, 0.2, 0.3)))
    met = sc.metal((0.7, 0.6, 0.5), 0.25)
    die = sc.dielectric(1.5)
    lit = sc.diffuse_light(sc.tex_solid((4, 4, 4)))
    iso = sc.isotropic(sc.tex_solid((0.2, 0.4, 0.9)))
    n = np.array([[0, 1, 0.0]]); p = np.array([[1, 2, 3.0]]); rd = np.array([[1, -1, 0.0]]); ro = p - rd
    # first triple rejected (|.|^2 >= 1), second accepted
    xi = np.array([[0.99, 0.99, 0.99, 0.75, 0.5, 0.25, 0.1, 0.1]])
    v = np.array([0.5, 0.0, -0.5]); unit = v / np.linalg.norm(v)
    r = sc.test_scatter(lam, ro, rd, [0.3], p, n, [1], None, None, xi)
    assert r["scattered"][0] == 1 and r["ndraw"][0] == 6 and np.allclose(r["dir"][0], n[0] + unit, atol=1e-15)
    assert np.allclose(r["attenuation"][0], [0.1, 0.2, 0.3]) and r["time"][0] == 0.3 and np.allclose(r["origin"][0], p[0])
    r = sc.test_scatter(met, ro, rd, [0.0], p, n, [1], None, None, xi)
    refl = np.array([1, 1, 0.0]) / math.sqrt(2)
    assert r["ndraw"][0] == 6 and np.allclose(r["dir"][0], refl + 0.25 * v, atol=1e-15) and r["scattered"][0] == 1
    r = sc.test_scatter(die, ro, rd, [0.0], p, n, [1], None, None, np.array([[0.999, 0.5]]))     # refracts
    assert r["ndraw"][0] == 1 and r["dir"][0][1] < 0 and np.allclose(r["attenuation"][0], 1)
    # total internal reflection from inside: no draw at all
    r = sc.test_scatter(die, ro, np.array([[1, -0.2, 0.0]]), [0.0], p, n, [0], None, None, np.array([[0.999, 0.5]]))
    assert r["ndraw"][0] == 0 and r["dir"][0][1] > 0
    for front in (0, 1):
        r = sc.test_scatter(lit, ro, rd, [0.0], p, n, [front], None, None, xi)
        assert r["scattered"][0] == 0 and r["ndraw"][0] == 0 and np.allclose(r["emitted"][0], 4)
    r = sc.test_scatter(iso, ro, rd, [0.0], p, n, [1], None, None, xi)
    assert r["ndraw"][0] == 6 and np.allclose(r["dir"][0], v, atol=1e-15)                              # not normalised (:85)


def test_get_ray_draw_order(pkg, orc):
    cam = orc.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0)
    # disk: (0.95,0.95) -> (0.9,0.9) rejected; (0.75,0.5) -> (0.5,0) accepted; then time
    xi = np.array([[0.95, 0.95, 0.75, 0.5, 0.125, 0.9]])
    r = orc.test_get_ray(cam, [0.25], [0.75], xi)
    assert r["ndraw"][0] == 5 and r["time"][0] == 0.125
    off = np.array(list(cam.u)) * (0.05 * 0.5)
    assert np.allclose(r["origin"][0], np.array([13, 2, 3.0]) + off, atol=1e-15)
    want_d = np.array(list(cam.lower_left_corner)) + 0.25 * np.array(list(cam.horizontal)) + 0.75 * np.array(list(cam.vertical)) - np.array([13, 2, 3.0]) - off
    assert np.allclose(r["dir"][0], want_d, atol=1e-14)                                               # src/camera.rs:58-66


def test_bvh_node_equals_flat_list(pkg, orc):
    a, spec = pkg.scenes.build(orc, "random_scene", wrap_bvh=False)
    b, _ = pkg.scenes.build(orc, "random_scene", wrap_bvh=True)
    cam = spec.camera(orc, 60, 40)
    rs = np.random.RandomState(3)
    g = orc.test_get_ray(cam, rs.rand(3000), rs.rand(3000), q24(rs, (3000, 16)))
    ha = a.test_hit(-1, g["origin"], g["dir"], g["time"])
    hb = b.test_hit(-1, g["origin"], g["dir"], g["time"])
    assert np.array_equal(ha["hit"], hb["hit"]) and np.array_equal(ha["t"], hb["t"]) and np.array_equal(ha["mat"], hb["mat"])
    p = pkg.make_params(60, 40, 2, background=spec.background)
    xs, ys = np.meshgrid(np.arange(60), np.arange(40))
    ra, _ = a.trace_paths(cam, p, xs.ravel(), ys.ravel(), np.zeros(2400, int))
    rb, _ = b.trace_paths(cam, p, xs.ravel(), ys.ravel(), np.zeros(2400, int))
    assert np.array_equal(ra, rb)


def test_media_deferred_is_distribution_equivalent(pkg, orc):
    """The device evaluates media after the surface closest hit; the literal reference order interleaves them.
    Same free-flight distribution => same image in expectation (3-sigma on the image mean and per-pixel z-scores)."""
    W, H, spp = 24, 24, 400
    res = []
    for deferred in (False, True):
        sc, spec = pkg.scenes.build(orc, "cornell_box_smoke")
        sc.set_media_deferred(deferred)
        p = pkg.make_params(W, H, spp, background=spec.background, seed=11 + deferred)
        r = sc.render_oracle(spec.camera(orc, W, H), p, threads=0, sumsq=True)
        mean = r["sum"] / spp
        var = np.maximum(r["sumsq"] / spp - mean ** 2, 0) / spp
        res.append((mean, var))
    z = (res[0][0] - res[1][0]) / np.sqrt(res[0][1] + res[1][1] + 1e-12)
    assert abs(res[0][0].mean() - res[1][0].mean()) < 4 * math.sqrt((res[0][1].sum() + res[1][1].sum())) / res[0][0].size
    assert np.mean(np.abs(z) > 3) < 0.02


def test_write_color(orc):
    import ctypes as C
    spp = 4
    sums = np.array([[0.25 * spp, 1.0 * spp, 2.0 * spp], [0.0, float("nan"), 1e-9]], float)
    out = np.zeros(6, np.uint8)
    orc.f("write_color")(sums.ctypes.data_as(C.POINTER(C.c_double)), 2, spp, out.ctypes.data_as(C.POINTER(C.c_uint8)))
    assert list(out) == [128, 255, 255, 0, 0, 0]        # sqrt(.25)=.5 -> 128; clamp .999*256 = 255.7 -> 255; NaN -> 0


def test_estimator_properties(pkg, orc):
    """ray_color (src/main.rs:19-38): depth 0 returns black; an empty world returns the background."""
    sc = pkg.Scene(orc)
    cam = orc.camera_new((0, 0, 5), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 5.0)
    p = pkg.make_params(8, 8, 1, background=(0.1, 0.2, 0.3))
    rgb, seg = sc.trace_paths(cam, p, [3], [4], [0])
    assert np.allclose(rgb[0], [0.1, 0.2, 0.3]) and seg[0] == 1
    p0 = pkg.make_params(8, 8, 1, max_depth=0, background=(0.1, 0.2, 0.3))
    rgb, seg = sc.trace_paths(cam, p0, [3], [4], [0])
    assert np.allclose(rgb[0], 0) and seg[0] == 0
    # a light seen directly: emitted only, both faces
    lit = sc.diffuse_light(sc.tex_solid((4, 5, 6)))
    sc.push(sc.xy_rect(lit, -10, 10, -10, 10, 0))
    for z in (5, -5):
        cam = orc.camera_new((0, 0, z), (0, 0, 0), (0, 1, 0), 40.0, 1.0, 0.0, 5.0)
        rgb, seg = sc.trace_paths(cam, p, [3], [4], [0])
        assert np.allclose(rgb[0], [4, 5, 6]) and seg[0] == 1


# ------------------------------------------------------------------------------------------------
# The one reference OUTPUT that matches a HEAD scene: generated_images/earth.ppm (400x225, older commit: same globe,
# same camera, gradient sky).  tests/golden/ref_earth_400x225_blocks5.npy = its 5x5 block means (tools/make_golden_earth.py).
# ------------------------------------------------------------------------------------------------
def earth_reference_check(pkg, cam, sums, spp):
    """8-bit image like write_color (src/math.rs:119-132) -> 5x5 block means -> compare on the globe (analytic mask,
    radius shrunk by 10 % to stay off the limb, where the old gradient sky shows through pixel coverage)."""
    ref = np.load(os.path.join(ROOT, "tests", "golden", "ref_earth_400x225_blocks5.npy"))
    W, H = 400, 225
    c8 = np.floor(np.clip(np.sqrt(np.clip(sums / spp, 0, None)), 0, 0.999) * 256)
    blk = c8.reshape(45, 5, 80, 5, 3).mean((1, 3))
    o, ll = np.array(list(cam.origin)), np.array(list(cam.lower_left_corner))
    hz, vt = np.array(list(cam.horizontal)), np.array(list(cam.vertical))
    ys, xs = np.mgrid[0:45, 0:80]
    u, v = (xs * 5 + 2.5) / (W - 1), ((H - 1) - (ys * 5 + 2.5)) / (H - 1)
    d = ll + u[..., None] * hz + v[..., None] * vt - o
    a, hb, c = (d * d).sum(-1), (d * o).sum(-1), (o * o).sum() - (2.0 * 0.9) ** 2
    mask = hb * hb - a * c > 0
    assert 800 <= mask.sum() <= 1000                       # the globe covers ~ a quarter of the frame

    def corr(p, q):
        p, q = p - p.mean(), q - q.mean()
        return float((p * q).sum() / np.sqrt((p * p).sum() * (q * q).sum()))
    for ch in range(3):
        assert corr(blk[mask][:, ch], ref[mask][:, ch]) >= 0.995, ch
        # a wrong u or v orientation (src/math.rs:296-299, src/texture.rs:48-54) decorrelates completely
        assert abs(corr(blk[:, ::-1][mask][:, ch], ref[mask][:, ch])) < 0.5
        assert abs(corr(blk[::-1][mask][:, ch], ref[mask][:, ch])) < 0.5
    # absolute level too: same texels, same albedo * sky transport, same gamma (measured 1.3 / 255)
    assert np.abs(blk[mask] - ref[mask]).mean() <= 3.0


def test_oracle_reproduces_the_reference_earth_image(pkg, orc):
    """Pins sphere hit + normal, sphere_uv, image-texture addressing (and our JPEG decode against stb_image's), the
    camera framing, Lambertian transport under the sky and write_color's gamma against pixels the reference itself wrote."""
    sc, spec = pkg.scenes.build(orc, "earth")
    cam = spec.camera(orc, 400, 225)
    spp = 32
    r = sc.render_oracle(cam, pkg.make_params(400, 225, spp, background=spec.background, seed=3), threads=0)
    earth_reference_check(pkg, cam, r["sum"], spp)
