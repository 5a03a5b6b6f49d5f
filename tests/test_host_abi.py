"""Host-side logic and the C-ABI surface (no GPU needed): the library loads, exports every symbol declared in
include/rtw.h, rejects bad input with status codes instead of panicking, flattens the reference's compositions,
and fails loudly — never falls back to a CPU path — when asked to compute without a device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_every_declared_symbol_is_exported(rtw):
    hdr = open(os.path.join(ROOT, "include", "rtw.h")).read()
    names = sorted(set(re.findall(r"\b(rtw_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 40
    missing = [n for n in names if not hasattr(rtw.dll, n)]
    assert not missing, missing


def test_product_does_not_link_or_load_the_oracle(rtw):
    import subprocess
    out = subprocess.run(["ldd", rtw.path], capture_output=True, text=True).stdout
    assert "oracle" not in out
    syms = subprocess.run(["nm", "-D", "--defined-only", rtw.path], capture_output=True, text=True).stdout
    assert "orc_" not in syms
    # no product source — native OR Python — includes, links, imports or dlopens anything under oracle/: the package
    # cannot reach the CPU path (the oracle's own loader lives in oracle/binding.py)
    pkg_dir = os.path.join(ROOT, "rust-ray-tracing-in-a-weekend_b200")
    n = 0
    for base, dirs, files in os.walk(pkg_dir):
        dirs[:] = [d for d in dirs if d not in ("__pycache__", "variants")]
        for f in files:
            if not f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h", "Makefile")):
                continue
            src = open(os.path.join(base, f), errors="ignore").read()
            n += 1
            assert "liboracle" not in src and "orc_" not in src and "load_oracle" not in src and "oracle.binding" not in src, f
            assert "oracle/" not in src, f
    assert n >= 10
    assert not hasattr(pkg_mod(), "load_oracle") and not hasattr(pkg_mod().api, "load_oracle") and not hasattr(pkg_mod().api, "ORACLE_LIB_PATH")


def pkg_mod():
    import rtw_pkg
    return rtw_pkg.load()


def test_struct_layouts(pkg):
    assert C.sizeof(pkg.api.Camera) == 24 * 8
    assert C.sizeof(pkg.api.RenderParams) == 16 + 24 + 8 + 8 + 16
    assert C.sizeof(pkg.api.Stats) == 3 * 8 + 2 * 8 + 8 * 8 + 2 * 4 + 2 * 8 + 4 * 4


def test_status_codes_instead_of_panics(pkg, rtw):
    sc = pkg.Scene(rtw)
    with pytest.raises(pkg.RtwError) as e:
        sc.sphere(1, (0, 0, 0), 1.0)              # no material registered: reference would index out of range (main.rs:26)
    assert e.value.code == -1
    with pytest.raises(pkg.RtwError):
        sc.lambertian(3)                           # unknown texture id
    m = sc.lambertian(sc.tex_solid((1, 1, 1)))
    assert m == 1                                  # 1-based handle like World::register_material (main.rs:46-49)
    assert sc.metal((1, 1, 1), 0.1) == 2
    with pytest.raises(pkg.RtwError):
        sc.translate(99, (0, 0, 0))
    with pytest.raises(pkg.RtwError):
        sc.push(1234)
    with pytest.raises(pkg.RtwError):
        sc.tex_noise(np.zeros((256, 3)), np.full(256, 300), np.zeros(256), np.zeros(256), 1.0)
    assert rtw.f("last_error")()


def test_non_finite_and_degenerate_inputs_are_rejected(pkg, rtw):
    """The reference turns these into NaN pixels or divisions by zero (src/hittable.rs:205 -1/density, :556-558
    (time - time0) / (time1 - time0), :283 / radius, src/camera.rs:34 unit_vector(0)); the ABI returns INVALID_ARG."""
    sc = pkg.Scene(rtw)
    nan, inf = float("nan"), float("inf")
    t = sc.tex_solid((1, 1, 1))
    m = sc.lambertian(t)
    s0 = sc.sphere(m, (0, 0, 0), 1.0)
    bad = [
        lambda: sc.tex_solid((nan, 0, 0)), lambda: sc.tex_checker((0, 0, 0), (0, inf, 0)),
        lambda: sc.metal((1, nan, 1), 0.0), lambda: sc.metal((1, 1, 1), inf), lambda: sc.dielectric(nan), lambda: sc.dielectric(0.0),
        lambda: sc.sphere(m, (nan, 0, 0), 1.0), lambda: sc.sphere(m, (0, 0, 0), 0.0), lambda: sc.sphere(m, (0, 0, 0), inf),
        lambda: sc.moving_sphere(m, (0, 0, 0), (0, 1, 0), 0.5, 0.5, 1.0), lambda: sc.moving_sphere(m, (0, 0, 0), (0, nan, 0), 0.0, 1.0, 1.0),
        lambda: sc.xy_rect(m, 0, 1, 0, nan, 0), lambda: sc.xz_rect(m, 0, inf, 0, 1, 0), lambda: sc.yz_rect(m, 0, 1, 0, 1, nan),
        lambda: sc.box((0, 0, 0), (1, nan, 1), m), lambda: sc.translate(s0, (0, inf, 0)), lambda: sc.rotate_y(nan, s0),
        lambda: sc.rotate_y_sincos(nan, 1.0, s0),
        lambda: sc.constant_medium(s0, 0.0, m), lambda: sc.constant_medium(s0, nan, m),
        lambda: sc.sphere_batch(np.array([m, m], np.int32), np.array([[0, 0, 0], [0, nan, 0]], float), np.array([1.0, 1.0])),
        lambda: sc.sphere_batch(np.array([m], np.int32), np.array([[0, 0, 0]], float), np.array([0.0])),
        lambda: rtw.camera_new((0, 0, 0), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0),           # look_from == look_at
        lambda: rtw.camera_new((0, 0, 1), (0, 0, 0), (0, 0, 1), 20.0, 1.5, 0.1, 10.0),           # vup parallel to the view axis
        lambda: rtw.camera_new((0, 0, 1), (0, 0, 0), (0, 1, 0), nan, 1.5, 0.1, 10.0),
    ]
    for i, fn in enumerate(bad):
        with pytest.raises(pkg.RtwError) as e:
            fn()
        assert e.value.code == -1, i
    # negative radius stays legal (the hollow-glass trick of the book), and nothing above was added to the scene
    sc.sphere(sc.dielectric(1.5), (0, 1, 0), -0.9)
    sc.push(s0)
    assert sc.debug_flatten()["prims"] == 1


def test_rotate_y_sincos_equals_rotate_y(pkg, rtw):
    """new_rotate_y stores sin/cos of the angle (src/hittable.rs:147-152); a host holding a built RotateY passes them on."""
    import math
    out = []
    for how in ("deg", "sincos"):
        sc = pkg.Scene(rtw)
        b = sc.box((0, 0, 0), (1, 2, 3), sc.lambertian(sc.tex_solid((1, 1, 1))))
        r = math.radians(15.0)
        sc.push(sc.translate(sc.rotate_y(15.0, b) if how == "deg" else sc.rotate_y_sincos(math.sin(r), math.cos(r), b), (1, 0, 2)))
        out.append(sc.debug_flatten())
    assert out[0] == out[1]


def test_bvh_node_members_cloned_by_the_reference_builder_are_emitted_once(pkg, rtw):
    """new_bvh_node puts a single-object span into BOTH children as separate clones (src/hittable.rs:96-98), so a host
    that walks the built tree hands odd-span leaves over twice.  The flattener drops members that are field-by-field
    equal to an earlier member of the same BvhNode (nested wrappers included); ConstantMedium members are kept."""
    def reference_tree(objs):                      # the shape new_bvh_node builds: [left, right] with left == right for one object
        if len(objs) == 1:
            return ("node", ("leaf", objs[0]), ("leaf", objs[0]))
        if len(objs) == 2:
            return ("node", ("leaf", objs[0]), ("leaf", objs[1]))
        mid = len(objs) // 2
        return ("node", reference_tree(objs[:mid]), reference_tree(objs[mid:]))

    def register(sc, m, tree, make):               # every leaf visit builds a NEW hittable = a Rust clone (distinct ids, equal fields)
        if tree[0] == "leaf":
            return make(sc, m, tree[1])
        return sc.bvh_node([register(sc, m, tree[1], make), register(sc, m, tree[2], make)], 0.0, 1.0)

    rs = np.random.RandomState(5)
    centers = rs.uniform(0, 165, (37, 3))
    def sphere(sc, m, i):
        return sc.translate(sc.rotate_y(15.0, sc.sphere(m, tuple(centers[i]), 10.0)), (-100.0, 270.0, 395.0)) if i % 2 else sc.sphere(m, tuple(centers[i]), 10.0)
    sc = pkg.Scene(rtw)
    m = sc.lambertian(sc.tex_solid((0.7, 0.7, 0.7)))
    sc.push(register(sc, m, reference_tree(list(range(37))), sphere))
    d = sc.debug_flatten()
    assert d["prims"] == 37 and d["dedup"] > 0, d          # prim count = object count
    # boxes (one BVH leaf + six face records each) and moving spheres dedupe the same way; a repeated id counts as a clone too
    sc = pkg.Scene(rtw)
    m = sc.lambertian(sc.tex_solid((0.7, 0.7, 0.7)))
    b1 = sc.box((0, 0, 0), (1, 1, 1), m); b2 = sc.box((0, 0, 0), (1, 1, 1), m); b3 = sc.box((0, 0, 0), (1, 1, 2), m)
    ms1 = sc.moving_sphere(m, (5, 0, 0), (5, 1, 0), 0.0, 1.0, 0.5); ms2 = sc.moving_sphere(m, (5, 0, 0), (5, 1, 0), 0.0, 1.0, 0.5)
    sc.push(sc.bvh_node([b1, b2, b3, ms1, ms2, b1], 0.0, 1.0))
    d = sc.debug_flatten()
    assert d["bvh_prims"] == 1 + 1 + 1 and d["prims"] == 7 + 7 + 1 and d["dedup"] == 3, d
    # world-level repeats are NOT merged (the reference's world list tests both), nor are ConstantMedium members
    sc = pkg.Scene(rtw)
    m = sc.lambertian(sc.tex_solid((0.7, 0.7, 0.7))); iso = sc.isotropic(sc.tex_solid((1, 1, 1)))
    s1 = sc.sphere(m, (0, 0, 0), 1.0)
    sc.push(s1); sc.push(s1)
    sc.push(sc.bvh_node([sc.constant_medium(s1, 0.1, iso), sc.constant_medium(s1, 0.1, iso)], 0.0, 1.0))
    d = sc.debug_flatten()
    assert d["bvh_prims"] == 2 and d["media"] == 2 and d["dedup"] == 0, d


def test_unsupported_nesting_is_reported(pkg, rtw):
    sc = pkg.Scene(rtw)
    m = sc.lambertian(sc.tex_solid((1, 1, 1)))
    iso = sc.isotropic(sc.tex_solid((1, 1, 1)))
    h = sc.sphere(m, (0, 0, 0), 1)
    for _ in range(5):
        h = sc.rotate_y(10, sc.translate(h, (1, 0, 0)))
    sc.push(h)
    with pytest.raises(pkg.RtwError) as e:
        sc.debug_flatten()
    assert e.value.code == -2
    sc2 = pkg.Scene(rtw)
    m = sc2.lambertian(sc2.tex_solid((1, 1, 1)))
    iso = sc2.isotropic(sc2.tex_solid((1, 1, 1)))
    med = sc2.constant_medium(sc2.sphere(m, (0, 0, 0), 1), 0.1, iso)
    sc2.push(sc2.translate(med, (1, 0, 0)))
    with pytest.raises(pkg.RtwError) as e:
        sc2.debug_flatten()
    assert e.value.code == -2
    sc3 = pkg.Scene(rtw)
    m = sc3.lambertian(sc3.tex_solid((1, 1, 1)))
    iso = sc3.isotropic(sc3.tex_solid((1, 1, 1)))
    sc3.push(sc3.constant_medium(sc3.constant_medium(sc3.sphere(m, (0, 0, 0), 1), 0.1, iso), 0.1, iso))
    with pytest.raises(pkg.RtwError) as e:
        sc3.debug_flatten()
    assert e.value.code == -2


# (records, BVH leaves): a surface Box is ONE leaf (PRIM_BOX) + its six face records behind the BVH primitives; the boxes of
# cornell_box_smoke are ConstantMedium boundaries: ONE record each (the slab test yields both crossings)
@pytest.mark.parametrize("name,prims,bvh_prims,media,xforms", [
    ("random_scene", None, None, 0, 1), ("two_spheres", 2, 2, 0, 1), ("two_perlin_spheres", 2, 2, 0, 1), ("earth", 1, 1, 0, 1),
    ("simple_light", 3, 3, 0, 1), ("cornell_box", 6 + 2 * 7, 6 + 2, 0, 3), ("cornell_box_smoke", 6 + 2, 6, 2, 3),
    ("final_scene", 400 * 7 + 1 + 1 + 2 + 1 + 2 + 1000 + 2, 400 + 1 + 1 + 2 + 1 + 2 + 1000, 2, 2)])
def test_flatten_reference_compositions(pkg, rtw, name, prims, bvh_prims, media, xforms):
    sc, spec = pkg.scenes.build(rtw, name)
    d = sc.debug_flatten()                 # also runs the structural BVH validation
    if prims is not None:
        assert d["prims"] == prims and d["bvh_prims"] == bvh_prims
    else:
        assert d["prims"] == 1 + sum(spec.info.values()) + 3
    assert d["media"] == media and d["xforms"] == xforms and d["depth"] <= 60
    assert d["bvh_prims"] <= d["prims"] and d["nodes"] >= 1


def test_scene_generators_are_seeded_and_shared(pkg, rtw, orc):
    a, sa = pkg.scenes.build(rtw, "random_scene", seed=1)
    b, sb = pkg.scenes.build(orc, "random_scene", seed=1)
    c, sc_ = pkg.scenes.build(rtw, "random_scene", seed=2)
    assert sa.info == sb.info and a.world == b.world and sa.info != sc_.info
    vals = [pkg.scenes.HostRng(1).random_double() for _ in range(2)]
    assert vals[0] == vals[1] and 0 <= vals[0] < 1
    # SplitMix64 known answer: seed 0 -> first output 0xE220A8397B1DCDAF
    assert pkg.scenes.HostRng(0).random_double() == (0xE220A8397B1DCDAF >> 11) / 2.0 ** 53


def test_empty_world_and_single_prim_flatten(pkg, rtw):
    sc = pkg.Scene(rtw)
    d = sc.debug_flatten()
    assert d["prims"] == 0 and d["nodes"] == (1 if d["bvh_width"] == 2 else 0)      # binary: one sentinel node; wide: no nodes at all
    m = sc.lambertian(sc.tex_solid((1, 1, 1)))
    sc.push(sc.sphere(m, (0, 0, 0), 1))
    d = sc.debug_flatten()
    assert d["prims"] == 1 and d["nodes"] == 1


def test_no_cpu_fallback(pkg, rtw):
    """Without a device every compute entry point must fail with NO_DEVICE — never compute on the host."""
    if rtw.f("device_count")() > 0:
        pytest.skip("a CUDA device is present")
    sc, spec = pkg.scenes.build(rtw, "two_spheres")
    with pytest.raises(pkg.RtwError) as e:
        sc.commit(1, 0)
    assert e.value.code == -5
    with pytest.raises(pkg.RtwError) as e:
        rtw.philox([0, 0, 0, 0], [0, 0])
    assert e.value.code == -5
    p = pkg.make_params(8, 8, 1)
    with pytest.raises(pkg.RtwError) as e:
        sc.render(spec.camera(rtw, 8, 8), p)
    assert e.value.code == -6          # not committed
    with pytest.raises(pkg.RtwError) as e:
        sc.test_hit(-1, [[0, 0, 0]], [[0, 0, 1]])
    assert e.value.code in (-5, -6)


def test_golden_fixture_matches_oracle(pkg, orc):
    """tests/golden/oracle_paths_v2.npz (made by tools/make_golden.py from THIS oracle) pins the restatement against
    accidental edits: same scenes, same Philox keys, bit-identical radiance."""
    path = os.path.join(ROOT, "tests", "golden", "oracle_paths_v2.npz")
    g = np.load(path)
    for name in pkg.scenes.SCENES:
        sc, spec = pkg.scenes.build(orc, name, wrap_bvh=name not in ("cornell_box_smoke", "final_scene"))
        sc.set_media_deferred(True)
        W, H = int(g["size"][0]), int(g["size"][1])
        p = pkg.make_params(W, H, 64, background=spec.background, seed=int(g["seed"]))
        rgb, seg = sc.trace_paths(spec.camera(orc, W, H), p, g["px"], g["py"], g["sample"])
        assert np.array_equal(seg, g[name + "_seg"]), name
        assert np.allclose(rgb, g[name + "_rgb"], rtol=1e-12, atol=1e-14), name


SCENE_IDS = {"random_scene": 0, "two_spheres": 1, "two_perlin_spheres": 2, "earth": 3, "simple_light": 4, "cornell_box": 5,
             "cornell_box_smoke": 6, "final_scene": 7}


@pytest.mark.parametrize("name", list(SCENE_IDS))
def test_cpp_mirror_flattens_like_the_python_binding(pkg, rtw, name):
    """host/rtw.hpp + scenes.hpp (C++ twin of the Rust shim: reference type/constructor names, World -> flatten -> C ABI)
    must hand the backend the same scene as scenes.py: same seeded streams, same primitive / node / material counts."""
    import subprocess
    exe = os.path.join(ROOT, "rust-ray-tracing-in-a-weekend_b200", "host", "rtw_main")
    if not os.path.exists(exe):
        import __graft_entry__
        __graft_entry__.build()
    out = subprocess.run([exe, "--scene", str(SCENE_IDS[name]), "--dry-run"], capture_output=True, text=True, check=True).stdout.split()
    got = dict(zip(out[0::2], map(int, out[1::2])))
    sc, _ = pkg.scenes.build(rtw, name)
    want = sc.debug_flatten()
    for k in ("prims", "bvh_prims", "nodes", "xforms", "media", "mats", "depth"):
        assert got[k] == want[k], (k, got, want)
    r = subprocess.run([exe, "--scene", "9", "--dry-run"], capture_output=True, text=True)
    assert r.returncode == 1 and "Unsupported scene selected" in r.stderr          # src/main.rs:461-463, as an error not a panic


def test_output_files_ppm_and_png(pkg, rtw, tmp_path):
    """Output stage (host only): rtw_write_ppm reproduces the reference's stdout bytes — header
    "P3\\n{W} {H}\\n255\\n\\n" (src/main.rs:472), one "r g b" line per pixel (src/math.rs:127-131), rows top to bottom —
    and rtw_write_png stores the same pixels in a valid PNG (signature, chunk CRCs, zlib stream, filter-0 scanlines)."""
    import struct, zlib
    rs = np.random.RandomState(5)
    for W, H in ((7, 5), (300, 131), (1, 1)):                    # 300x131x3 > 65535: several stored deflate blocks
        img = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
        ppm, png = tmp_path / f"a{W}.ppm", tmp_path / f"a{W}.png"
        pkg.api.write_ppm(rtw, ppm, img, W, H)
        pkg.api.write_png(rtw, png, img, W, H)
        want = f"P3\n{W} {H}\n255\n\n" + "".join(f"{r} {g} {b}\n" for r, g, b in img.reshape(-1, 3))
        assert ppm.read_text() == want
        raw = png.read_bytes()
        assert raw[:8] == b"\x89PNG\r\n\x1a\n"
        pos, chunks = 8, []
        while pos < len(raw):
            n, = struct.unpack(">I", raw[pos:pos + 4]); typ = raw[pos + 4:pos + 8]; data = raw[pos + 8:pos + 8 + n]
            crc, = struct.unpack(">I", raw[pos + 8 + n:pos + 12 + n])
            assert zlib.crc32(typ + data) == crc
            chunks.append((typ, data)); pos += 12 + n
        assert [t for t, _ in chunks] == [b"IHDR", b"IDAT", b"IEND"]
        assert struct.unpack(">IIBBBBB", chunks[0][1]) == (W, H, 8, 2, 0, 0, 0)
        lines = np.frombuffer(zlib.decompress(chunks[1][1]), np.uint8).reshape(H, 1 + 3 * W)
        assert (lines[:, 0] == 0).all() and np.array_equal(lines[:, 1:].reshape(H, W, 3), img)
    with pytest.raises(pkg.RtwError):
        pkg.api.write_png(rtw, tmp_path / "no_such_dir" / "x.png", img, 1, 1)


def test_parallel_bvh_build_equals_serial(pkg, rtw, monkeypatch):
    """The host SAH builder streams the top levels of big scenes with all threads (per-thread bins, out-of-place
    partition by prefix sums) and builds the subtrees below as tasks.  Same tree as the serial builder: same node
    count, depth and SAH cost, and the structural validation (every primitive once, inside all its ancestors) passes."""
    n = 300_000                                    # above the 2^18 threshold of the all-threads path
    out = {}
    for mode, threads in (("serial", "1"), ("parallel", "4"), ("parallel8", "8")):
        monkeypatch.setenv("RTW_BUILD_THREADS", threads)
        sc = pkg.Scene(rtw)
        pkg.scenes.sweep_scene(sc, n, seed=11)
        out[mode] = sc.debug_flatten()             # flatten + BVH + validate_bvh
        sc.close()
    for mode in ("parallel", "parallel8"):
        for k in ("prims", "nodes", "depth"):
            assert out[mode][k] == out["serial"][k], (mode, k)
        assert abs(out[mode]["sah"] - out["serial"]["sah"]) <= 1e-9 * out["serial"]["sah"]
    if out["serial"]["bvh_width"] == 2:
        assert out["serial"]["nodes"] == n - 1
    else:       # 300 k primitives are beyond the cache-resident range: 8-wide nodes, ~4.4 primitives per node (also validated)
        assert n / 8 < out["serial"]["nodes"] < n / 3 and 0 < out["serial"]["wide_depth"] <= 32


@pytest.mark.parametrize("name", list(SCENE_IDS) + ["sweep_1", "sweep_2", "sweep_9", "sweep_5000", "sweep_120000"])
def test_wide_bvh_structure_and_conservative_traversal(pkg, rtw, name):
    """The 8-wide compressed BVH (csrc/bvh_wide.h; replaces the binary BvhNode of src/hittable.rs:77-130, :290-306 for
    scenes beyond the caches).  Host-only: collapse the SAH tree, validate (every primitive in exactly one leaf slot,
    every quantised child box contains what lies below it), then trace seeded rays — random, grazing, axis-parallel with
    both signs of zero — through the quantised traversal ON THE CPU with the device's arithmetic (same source:
    rtww::wide_node_hits / wide_perm16) and require that it reaches every primitive whose box the ray really crosses."""
    if name.startswith("sweep_"):
        sc = pkg.Scene(rtw)
        pkg.scenes.sweep_scene(sc, int(name[6:]), seed=5)
    else:
        sc, _ = pkg.scenes.build(rtw, name)
    d = sc.debug_wide(400 if name == "sweep_120000" else 2500, seed=3)
    assert d["missed"] == 0 and d["rays"] > 0
    assert d["leaves_reached"] >= d["boxes_crossed"] > 0
    n = d["bvh_prims"]
    if n >= 100:        # 8-bit planes cost a few per cent of extra leaf tests (a 3-primitive scene whose single node spans the
        assert d["leaves_reached"] <= 1.35 * d["boxes_crossed"]          # r = 1000 ground sphere quantises at steps of 8: looser)
    assert (n + 7) // 8 <= d["wide_nodes"] <= max(1, n - 1) and 1 <= d["wide_depth"] <= 32
    if n >= 1000:
        assert d["wide_nodes"] < n / 3                                    # ~4 primitives per 80-byte node


def test_flatten_random_scene_graphs(pkg, rtw):
    """Random compositions of the reference's Hittable variants (src/hittable.rs:29-41): spheres, moving spheres,
    rects, boxes, Translate / RotateY wrappers up to the supported depth, BvhNodes of mixed members, media over
    spheres and (instanced) boxes.  Flatten must account for every surface primitive (a surface Box is one BVH leaf in
    front of its six face records; a Box that bounds a medium is one boundary record), count media,
    and produce a BVH that passes the structural validation; deeper wrapper chains are refused with
    RTW_ERR_UNSUPPORTED_NESTING, never mis-flattened."""
    rs = np.random.RandomState(2024)
    for trial in range(40):
        sc = pkg.Scene(rtw)
        mat = sc.lambertian(sc.tex_solid((0.5, 0.5, 0.5)))
        iso = sc.isotropic(sc.tex_solid((1.0, 1.0, 1.0)))
        n_surface = n_boundary = n_media = n_faces = 0
        too_deep = False

        def leaf():
            k = rs.randint(5)
            c = tuple(rs.uniform(-50, 50, 3))
            if k == 0:
                return sc.sphere(mat, c, float(rs.uniform(0.1, 5))), 1
            if k == 1:
                return sc.moving_sphere(mat, c, tuple(np.array(c) + rs.uniform(-1, 1, 3)), 0.0, 1.0, float(rs.uniform(0.1, 3))), 1
            if k == 2:
                a0, b0 = rs.uniform(-50, 40, 2)
                f = [sc.xy_rect, sc.xz_rect, sc.yz_rect][rs.randint(3)]
                return f(mat, a0, a0 + rs.uniform(1, 10), b0, b0 + rs.uniform(1, 10), float(rs.uniform(-50, 50))), 1
            lo = rs.uniform(-50, 40, 3)
            return sc.box(tuple(lo), tuple(lo + rs.uniform(1, 10, 3)), mat), 6

        def wrap(h, depth):
            for _ in range(depth):
                h = sc.translate(h, tuple(rs.uniform(-5, 5, 3))) if rs.randint(2) else sc.rotate_y(float(rs.uniform(-90, 90)), h)
            return h

        for _ in range(rs.randint(1, 12)):
            kind = rs.randint(4)
            depth = int(rs.choice([0, 0, 1, 2, 4, 5]))
            if kind == 0:                                            # wrapped leaf
                h, n = leaf()
                too_deep |= depth > 4
                sc.push(wrap(h, depth)); n_surface += 1; n_faces += 6 if n == 6 else 0
            elif kind == 1:                                          # BvhNode of leaves, possibly wrapped as a whole
                members = [leaf() for _ in range(rs.randint(1, 9))]
                node = sc.bvh_node([m[0] for m in members], 0.0, 1.0)
                too_deep |= depth > 4
                sc.push(wrap(node, depth)); n_surface += len(members); n_faces += sum(6 for m in members if m[1] == 6)
            elif kind == 2:                                          # medium over a sphere or an instanced box
                h, n = leaf()
                d = min(depth, 4)
                sc.push(sc.constant_medium(wrap(h, d), float(rs.uniform(0.001, 0.5)), iso)); n_boundary += 1; n_media += 1
            else:
                h, n = leaf()
                sc.push(h); n_surface += 1; n_faces += 6 if n == 6 else 0
        if too_deep:
            with pytest.raises(pkg.RtwError) as e:
                sc.debug_flatten()
            assert e.value.code == -2
        else:
            d = sc.debug_flatten()                                   # flatten + BVH + validate_bvh
            assert d["bvh_prims"] == n_surface and d["prims"] == n_surface + n_faces + n_boundary and d["media"] == n_media, trial
            if d["bvh_width"] == 2:
                assert d["nodes"] == max(1, n_surface - 1)           # one primitive per leaf
            else:
                assert (n_surface + 6) // 7 <= max(1, d["nodes"]) <= max(1, n_surface - 1)
        sc.close()


def test_public_header_is_plain_c(tmp_path):
    """include/rtw.h is the drop-in boundary: it must compile as C99 (what a Rust bindgen / cgo / ctypes user sees) and
    as C++, with no CUDA, torch or C++ types in any signature."""
    import subprocess
    hdr = os.path.join(ROOT, "include", "rtw.h")
    for args in (["gcc", "-std=c99", "-pedantic", "-x", "c"], ["g++", "-std=c++17", "-x", "c++"]):
        r = subprocess.run(args + ["-Wall", "-Wextra", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
    text = open(hdr).read()
    for banned in ("torch", "std::", "cudaStream", "at::Tensor", "#include <cuda"):
        assert banned not in text, banned
    # a C program can drive the host-only entry points
    src = tmp_path / "t.c"
    src.write_text('#include "rtw.h"\n#include <stdio.h>\nint main(void){ unsigned char px[6]={1,2,3,4,5,6};'
                   ' rtw_scene* s = rtw_scene_new(); double c[3]={0,0,0}; int t = rtw_tex_solid(s,c); int m = rtw_mat_lambertian(s,t);'
                   ' int h = rtw_sphere(s,m,c,1.0); int rc = rtw_world_push(s,h); rtw_scene_free(s);'
                   ' printf("%d %d %d %d %d\\n", t>=0, m>=1, h>=0, rc, rtw_write_ppm("o.ppm", px, 2, 1)); return 0; }\n')
    exe = tmp_path / "t"
    lib_dir = os.path.join(ROOT, "rust-ray-tracing-in-a-weekend_b200")
    r = subprocess.run(["gcc", "-std=c99", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe), "-L", lib_dir, "-lrtw",
                        f"-Wl,-rpath,{lib_dir}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe)], capture_output=True, text=True, cwd=tmp_path)
    assert r.returncode == 0 and r.stdout.split() == ["1", "1", "1", "0", "0"], (r.stdout, r.stderr)
    assert (tmp_path / "o.ppm").read_text() == "P3\n2 1\n255\n\n1 2 3\n4 5 6\n"


def test_wide_collapse_absorbs_small_subtrees(pkg, rtw, monkeypatch):
    """Collapse heuristic of the 8-wide tree (csrc/bvh_wide.h collapse_one): subtrees of more than 8 leaves are opened by
    surface area, a subtree of at most 8 leaves is either absorbed whole (all its leaves become slots of the node) or left
    as one child — never opened half-way.  Against the plain area-greedy collapse (RTW_WIDE_FILL=0) that removes the 2-3-leaf
    nodes at the bottom of the tree: fewer nodes, the same conservative traversal, and no more work per ray (host-only cost
    probe: closest-hit traversal of secondary-like rays on the CPU with the device's node test)."""
    res = {}
    for fill in ("0", "1"):
        monkeypatch.setenv("RTW_WIDE_FILL", fill)
        sc = pkg.Scene(rtw)
        pkg.scenes.sweep_scene(sc, 60000, seed=8)
        d = sc.debug_wide(300, seed=4)                       # validates the structure and the conservative traversal
        c = sc.debug_wide_cost(4000, seed=2)
        assert d["missed"] == 0 and c["rays"] == 4000 and c["hits"] > 0
        res[fill] = (d, c)
        sc.close()
    (d0, c0), (d1, c1) = res["0"], res["1"]
    n = d1["bvh_prims"]
    assert d1["wide_nodes"] < 0.7 * d0["wide_nodes"] and d1["wide_nodes"] < n / 4         # measured: 24 822 -> 11 064 nodes (5.4 primitives per node)
    # (the probe aims its rays at primitives by index, and the leaf order differs between the two trees: same distribution, not the same rays)
    assert c1["node_visits"] <= 1.05 * c0["node_visits"] and c1["prim_tests"] <= 1.05 * c0["prim_tests"]
    assert abs(c0["hits"] - c1["hits"]) <= 0.03 * c0["hits"]


def test_box_flattens_to_one_leaf_and_six_face_records(pkg, rtw, monkeypatch):
    """new_box (src/hittable.rs:132-145) on the device: ONE PRIM_BOX leaf in the BVH, its six rects as face records behind the
    BVH primitives (they describe the hit), ONE record when the box bounds a ConstantMedium; a flat box (no extent on one axis)
    and RTW_BOX_PRIM=0 keep the six rects as leaves."""
    def counts(build):
        sc = pkg.Scene(rtw)
        m = sc.lambertian(sc.tex_solid((0.5, 0.5, 0.5))); iso = sc.isotropic(sc.tex_solid((1, 1, 1)))
        build(sc, m, iso)
        d = sc.debug_flatten(); sc.close()
        return d["bvh_prims"], d["prims"], d["media"]
    solid = lambda sc, m, iso: sc.push(sc.translate(sc.rotate_y(15.0, sc.box((0, 0, 0), (1, 2, 3), m)), (1.0, 0.0, 0.0)))
    flat = lambda sc, m, iso: sc.push(sc.box((0, 0, 0), (1, 0, 3), m))
    fog = lambda sc, m, iso: sc.push(sc.constant_medium(sc.box((0, 0, 0), (1, 2, 3), m), 0.1, iso))
    assert counts(solid) == (1, 7, 0)
    assert counts(flat) == (6, 6, 0)
    assert counts(fog) == (0, 1, 1)
    monkeypatch.setenv("RTW_BOX_PRIM", "0")
    assert counts(solid) == (6, 6, 0)
    assert counts(fog) == (0, 6, 1)
