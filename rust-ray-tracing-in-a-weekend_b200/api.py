"""ctypes binding of the C ABI in include/rtw.h (librtw.so).

This is the reference-side binding a maintainer would write (the Python analogue of the Rust
`extern "C"` block shown in INTEGRATION.md): a scene is built with the reference's constructors
(src/hittable.rs:29-41, src/material.rs:6-12, src/texture.rs:4-9, src/camera.rs:18-56), committed,
and rendered.  `Lib` binds any library that exports this ABI under a symbol prefix; the package itself only ever
loads librtw.so.  (The test suite's CPU checker exports the same signatures under its own prefix; its loader lives
with it, outside this package — nothing here imports it or can reach it.)

librtw.so has NO CPU fallback: `load_rtw()` raises if the CUDA library is missing.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_HERE)

RTW_LIB_PATH = os.environ.get("RTW_LIB_PATH") or os.path.join(_HERE, "librtw.so")   # override: kernel experiments only

RTW_IPC_HANDLE_BYTES = 160
RTW_FLAG_DEVICE_OUT = 1
RTW_FLAG_KERNEL_MEGA = 2     # one path per lane (default: measured faster)
RTW_FLAG_KERNEL_POOL = 4     # warp-pool / shared-memory wavefront variant
RTW_FLAG_NO_TILE_CULL = 8    # diagnostic: no per-tile candidate lists
RTW_FLAG_KERNEL_WAVEFRONT = 16   # global-memory wavefront pipeline (path pool in HBM, dynamic ray fetch)

STATUS = {0: "OK", -1: "INVALID_ARG", -2: "UNSUPPORTED_NESTING", -3: "CUDA_ERROR", -4: "OOM",
          -5: "NO_DEVICE", -6: "NOT_COMMITTED"}


class RtwError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class Camera(C.Structure):
    """The 10 public fields of the reference Camera (src/camera.rs:4-15)."""
    _fields_ = [("origin", C.c_double * 3), ("lower_left_corner", C.c_double * 3),
                ("horizontal", C.c_double * 3), ("vertical", C.c_double * 3),
                ("u", C.c_double * 3), ("v", C.c_double * 3), ("w", C.c_double * 3),
                ("lens_radius", C.c_double), ("time0", C.c_double), ("time1", C.c_double)]


class RenderParams(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("spp", C.c_int32), ("max_depth", C.c_int32),
                ("background", C.c_double * 3), ("t_min", C.c_double), ("seed", C.c_uint64),
                ("n_gpus", C.c_int32), ("samples_per_unit", C.c_int32), ("flags", C.c_int32),
                ("reserved", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("ms_render", C.c_double), ("ms_total", C.c_double), ("ms_commit", C.c_double),
                ("paths", C.c_uint64), ("rays", C.c_uint64), ("units_per_device", C.c_uint64 * 8),
                ("n_devices", C.c_int32), ("kernel_launches", C.c_int32),
                ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("n_prims", C.c_int32), ("n_nodes", C.c_int32), ("n_materials", C.c_int32), ("n_media", C.c_int32)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


def make_params(width, height, spp, max_depth=50, background=(0.7, 0.8, 1.0), t_min=0.001, seed=1, n_gpus=0,
                samples_per_unit=0, flags=0):
    p = RenderParams()
    p.width, p.height, p.spp, p.max_depth = width, height, spp, max_depth
    p.background[:] = background
    p.t_min, p.seed, p.n_gpus, p.samples_per_unit, p.flags = t_min, seed, n_gpus, samples_per_unit, flags
    return p


def _d3(v):
    return (C.c_double * 3)(float(v[0]), float(v[1]), float(v[2]))


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def _p(a, t=C.c_double):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class Lib:
    """One loaded library exporting the include/rtw.h ABI under a symbol prefix ('rtw_' = the CUDA product)."""
    scene_cls = None          # subclasses may name a Scene subclass with extra, library-specific calls

    def __init__(self, path, prefix):
        if not os.path.exists(path):
            raise RtwError(-5, f"{path} is missing — build it first (python -c 'import __graft_entry__ as g; g.build()'). "
                           "There is no CPU fallback for the render path.")
        self.path, self.prefix = path, prefix
        self.dll = C.CDLL(path)
        self._sig()
        self._sig_extra()

    def f(self, name):
        return getattr(self.dll, self.prefix + name)

    def has(self, name):
        return hasattr(self.dll, self.prefix + name)

    def _sig(self):
        f = self.f
        f("last_error").restype = C.c_char_p
        f("scene_new").restype = C.c_void_p
        f("scene_free").argtypes = [C.c_void_p]
        f("scene_free").restype = None
        for n in ("tex_solid", "tex_checker", "tex_noise", "tex_image", "mat_lambertian", "mat_metal",
                  "mat_dielectric", "mat_diffuse_light", "mat_isotropic", "sphere", "moving_sphere", "xy_rect",
                  "xz_rect", "yz_rect", "box", "translate", "rotate_y", "constant_medium", "bvh_node",
                  "world_push", "camera_new"):
            f(n).restype = C.c_int
        dp, ip, vp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.c_void_p
        f("tex_solid").argtypes = [vp, dp]
        f("tex_checker").argtypes = [vp, dp, dp]
        f("tex_noise").argtypes = [vp, dp, ip, ip, ip, C.c_double]
        f("tex_image").argtypes = [vp, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
        f("mat_lambertian").argtypes = [vp, C.c_int]
        f("mat_metal").argtypes = [vp, dp, C.c_double]
        f("mat_dielectric").argtypes = [vp, C.c_double]
        f("mat_diffuse_light").argtypes = [vp, C.c_int]
        f("mat_isotropic").argtypes = [vp, C.c_int]
        f("sphere").argtypes = [vp, C.c_int, dp, C.c_double]
        f("moving_sphere").argtypes = [vp, C.c_int, dp, dp, C.c_double, C.c_double, C.c_double]
        f("sphere_batch").argtypes = [vp, C.c_int32, ip, dp, dp]
        f("sphere_batch").restype = C.c_int
        for n in ("xy_rect", "xz_rect", "yz_rect"):
            f(n).argtypes = [vp, C.c_int] + [C.c_double] * 5
        f("box").argtypes = [vp, dp, dp, C.c_int]
        f("translate").argtypes = [vp, C.c_int, dp]
        f("rotate_y").argtypes = [vp, C.c_double, C.c_int]
        f("constant_medium").argtypes = [vp, C.c_int, C.c_double, C.c_int]
        f("bvh_node").argtypes = [vp, ip, C.c_int32, C.c_double, C.c_double]
        f("world_push").argtypes = [vp, C.c_int]
        f("camera_new").argtypes = [dp, dp, dp] + [C.c_double] * 6 + [C.POINTER(Camera)]
        u32p = C.POINTER(C.c_uint32)
        f("test_philox").argtypes = [C.c_int32, u32p, u32p, u32p]
        f("test_get_ray").argtypes = [C.POINTER(Camera), C.c_int32, dp, dp, dp, C.c_int32, dp, dp, dp, ip]
        f("test_hit").argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, C.c_double, C.c_double, dp, C.c_int32,
                                  ip, dp, dp, dp, ip, dp, dp, ip, ip]
        f("test_aabb").argtypes = [C.c_int32, dp, dp, dp, dp, C.c_double, C.c_double, ip]
        f("test_scatter").argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, dp, dp, ip, dp, dp, dp, C.c_int32,
                                      ip, dp, dp, dp, dp, dp, ip]
        f("test_texture").argtypes = [vp, C.c_int32, C.c_int32, dp, dp, dp, dp]
        f("trace_paths").argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.c_int32, ip, ip, ip, dp, ip]

    def _sig_extra(self):
        """Entry points beyond the constructors / parity hooks: the product's commit, render, IPC and output calls."""
        f = self.f
        vp, ip, dp = C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_double)
        fp = C.POINTER(C.c_float)
        f("scene_commit").argtypes = [vp, C.c_int32, C.c_int32]
        f("render").argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.c_void_p, C.POINTER(Stats)]
        f("render_progressive").argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.c_int32, C.c_int32, C.c_void_p,
                                            C.c_void_p, C.c_void_p, C.POINTER(Stats)]
        f("render_progressive").restype = C.c_int
        f("write_color").argtypes = [fp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
        f("device_count").restype = C.c_int
        f("version").restype = C.c_char_p
        u8p = C.POINTER(C.c_uint8)
        f("shared_create").argtypes = [vp, C.c_int32, C.c_int32, u8p]
        f("shared_open").argtypes = [vp, C.c_int32, C.c_int32, u8p]
        f("shared_reset").argtypes = [vp]
        f("render_shared").argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.POINTER(Stats)]
        f("shared_read").argtypes = [vp, fp]
        f("render_shared_epoch").argtypes = [vp, C.POINTER(Camera), C.POINTER(RenderParams), C.c_int32, C.POINTER(Stats)]
        f("shared_read_epoch").argtypes = [vp, C.c_int32, fp]
        f("shared_close").argtypes = [vp]
        f("host_alloc").argtypes = [C.c_uint64]
        f("host_alloc").restype = C.c_void_p
        f("host_free").argtypes = [C.c_void_p]
        f("host_free").restype = None
        f("debug_flatten").argtypes = [vp, ip, dp]
        f("debug_flatten2").argtypes = [vp, ip, dp]
        f("debug_wide").argtypes = [vp, C.c_int32, C.c_uint64, C.POINTER(C.c_uint64)]
        f("debug_wide_cost").argtypes = [vp, C.c_int32, C.c_uint64, C.POINTER(C.c_uint64)]
        f("rotate_y_sincos").argtypes = [vp, C.c_double, C.c_double, C.c_int]
        f("rotate_y_sincos").restype = C.c_int

    def check(self, rc):
        if rc < 0:
            raise RtwError(rc, (self.f("last_error")() or b"").decode())
        return rc

    def camera_new(self, look_from, look_at, vup, vfov, aspect, aperture, focus_dist, time0=0.0, time1=1.0):
        """Camera::new (src/camera.rs:18-56)."""
        cam = Camera()
        self.check(self.f("camera_new")(_d3(look_from), _d3(look_at), _d3(vup), vfov, aspect, aperture, focus_dist,
                                         time0, time1, C.byref(cam)))
        return cam

    def pinned_image(self, height, width):
        """H x W x 3 float32 array in page-locked host memory (rtw_host_alloc); falls back to pageable memory."""
        n = height * width * 3
        p = self.f("host_alloc")(n * 4) if self.has("host_alloc") else None
        if not p:
            return np.zeros((height, width, 3), np.float32)
        buf = (C.c_float * n).from_address(p)
        arr = np.frombuffer(buf, dtype=np.float32).reshape(height, width, 3)
        arr[:] = 0
        self._pinned = getattr(self, "_pinned", []) + [p]
        return arr

    def philox(self, counter, key):
        counter = np.ascontiguousarray(counter, dtype=np.uint32).reshape(-1, 4)
        key = np.ascontiguousarray(key, dtype=np.uint32).reshape(-1, 2)
        out = np.zeros_like(counter)
        self.check(self.f("test_philox")(len(counter), _p(counter, C.c_uint32), _p(key, C.c_uint32), _p(out, C.c_uint32)))
        return out

    def test_get_ray(self, cam, s, t, xi):
        s, t, xi = _f64(s), _f64(t), _f64(xi)
        n = len(s)
        stride = xi.shape[1]
        o, d, tm, nd = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n), np.zeros(n, np.int32)
        self.check(self.f("test_get_ray")(C.byref(cam), n, _p(s), _p(t), _p(xi), stride, _p(o), _p(d), _p(tm),
                                           _p(nd, C.c_int32)))
        return dict(origin=o, dir=d, time=tm, ndraw=nd)

    def test_aabb(self, bmin, bmax, origin, direction, t_min, t_max):
        bmin, bmax, origin, direction = _f64(bmin), _f64(bmax), _f64(origin), _f64(direction)
        n = len(origin)
        out = np.zeros(n, np.int32)
        self.check(self.f("test_aabb")(n, _p(bmin), _p(bmax), _p(origin), _p(direction), t_min, t_max, _p(out, C.c_int32)))
        return out


_libs = {}


def load_rtw():
    """The CUDA product library.  Raises (never falls back) when it is not built."""
    if "rtw" not in _libs:
        _libs["rtw"] = Lib(RTW_LIB_PATH, "rtw_")
    return _libs["rtw"]


def write_ppm(lib, path, rgb8, width, height):
    """rtw_write_ppm: the reference's P3 text output (src/main.rs:472, :591-596) as a file."""
    a = np.ascontiguousarray(rgb8, np.uint8)
    assert a.size == width * height * 3
    lib.check(lib.f("write_ppm")(str(path).encode(), a.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int32(width), C.c_int32(height)))


def write_png(lib, path, rgb8, width, height):
    """rtw_write_png: 8-bit RGB PNG of the same pixels."""
    a = np.ascontiguousarray(rgb8, np.uint8)
    assert a.size == width * height * 3
    lib.check(lib.f("write_png")(str(path).encode(), a.ctypes.data_as(C.POINTER(C.c_uint8)), C.c_int32(width), C.c_int32(height)))


class Scene:
    """= reference `World` (src/main.rs:40-50): materials + hittables, built through the C ABI."""

    def __new__(cls, lib):
        return object.__new__(lib.scene_cls if cls is Scene and lib.scene_cls else cls)

    def __init__(self, lib):
        self.lib = lib
        self.h = lib.f("scene_new")()
        if not self.h:
            raise RtwError(-4, "scene_new failed")
        self.world = []

    def close(self):
        if self.h:
            self.lib.f("scene_free")(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _c(self, name, *a):
        return self.lib.check(self.lib.f(name)(self.h, *a))

    # --- Texture (src/texture.rs:4-9)
    def tex_solid(self, rgb):
        return self._c("tex_solid", _d3(rgb))

    def tex_checker(self, even, odd):
        return self._c("tex_checker", _d3(even), _d3(odd))

    def tex_noise(self, ranvec, perm_x, perm_y, perm_z, scale):
        rv = _f64(ranvec).reshape(256, 3)
        px, py, pz = _i32(perm_x), _i32(perm_y), _i32(perm_z)
        return self._c("tex_noise", _p(rv), _p(px, C.c_int32), _p(py, C.c_int32), _p(pz, C.c_int32), float(scale))

    def tex_image(self, rgb8):
        a = np.ascontiguousarray(rgb8, dtype=np.uint8)
        h, w, ch = a.shape
        assert ch == 3
        return self._c("tex_image", w, h, 3 * w, _p(a, C.c_uint8))

    # --- Material (src/material.rs:6-12); returns the 1-based MaterialHandle (src/main.rs:46-49)
    def lambertian(self, tex):
        return self._c("mat_lambertian", tex)

    def metal(self, albedo, fuzz):
        return self._c("mat_metal", _d3(albedo), float(fuzz))

    def dielectric(self, ir):
        return self._c("mat_dielectric", float(ir))

    def diffuse_light(self, tex):
        return self._c("mat_diffuse_light", tex)

    def isotropic(self, tex):
        return self._c("mat_isotropic", tex)

    # --- Hittable (src/hittable.rs:29-41)
    def sphere(self, mat, center, radius):
        return self._c("sphere", mat, _d3(center), float(radius))

    def sphere_batch(self, mats, centers, radii):
        """n static spheres straight into the world (rtw_sphere_batch) — the 1M-16M sweep."""
        m, c, r = _i32(mats), _f64(centers).reshape(-1, 3), _f64(radii)
        self._c("sphere_batch", len(m), _p(m, C.c_int32), _p(c), _p(r))

    def moving_sphere(self, mat, c0, c1, t0, t1, radius):
        return self._c("moving_sphere", mat, _d3(c0), _d3(c1), float(t0), float(t1), float(radius))

    def xy_rect(self, mat, x0, x1, y0, y1, k):
        return self._c("xy_rect", mat, float(x0), float(x1), float(y0), float(y1), float(k))

    def xz_rect(self, mat, x0, x1, z0, z1, k):
        return self._c("xz_rect", mat, float(x0), float(x1), float(z0), float(z1), float(k))

    def yz_rect(self, mat, y0, y1, z0, z1, k):
        return self._c("yz_rect", mat, float(y0), float(y1), float(z0), float(z1), float(k))

    def box(self, mn, mx, mat):
        return self._c("box", _d3(mn), _d3(mx), mat)

    def translate(self, child, offset):
        return self._c("translate", child, _d3(offset))

    def rotate_y(self, angle_deg, child):
        return self._c("rotate_y", float(angle_deg), child)

    def rotate_y_sincos(self, sin_theta, cos_theta, child):
        """RotateY as the reference stores it (src/hittable.rs:39); product ABI only."""
        return self._c("rotate_y_sincos", float(sin_theta), float(cos_theta), child)

    def constant_medium(self, child, density, phase_mat):
        return self._c("constant_medium", child, float(density), phase_mat)

    def bvh_node(self, children, t0=0.0, t1=1.0):
        ch = _i32(children)
        return self._c("bvh_node", _p(ch, C.c_int32), len(ch), float(t0), float(t1))

    def push(self, hittable):
        self._c("world_push", hittable)
        self.world.append(hittable)
        return hittable

    # --- product only
    def commit(self, n_gpus=1, first_device=0):
        return self._c("scene_commit", n_gpus, first_device)

    def debug_flatten(self):
        """Host-only: flatten + structural BVH validation (no device needed)."""
        counts = np.zeros(16, np.int32)
        sah = C.c_double(0)
        self._c("debug_flatten2", _p(counts, C.c_int32), C.byref(sah))
        keys = ("prims", "bvh_prims", "nodes", "xforms", "media", "mats", "texs", "depth", "dedup", "bvh_width", "wide_depth")
        d = dict(zip(keys, (int(x) for x in counts)))
        d["sah"] = sah.value
        return d

    def debug_wide(self, n_rays=2000, seed=1):
        """Host-only: build the 8-wide compressed BVH, validate it, and run the quantised traversal on the CPU (the
        device's arithmetic) against exact box tests.  Raises when a primitive box is missed."""
        out = np.zeros(8, np.uint64)
        self._c("debug_wide", n_rays, seed, _p(out, C.c_uint64))
        keys = ("rays", "node_visits", "leaves_reached", "boxes_crossed", "missed", "wide_nodes", "wide_depth", "bvh_prims")
        return dict(zip(keys, (int(x) for x in out)))

    def debug_wide_cost(self, n_rays=20000, seed=1):
        """Host-only tuning aid: closest-hit cost of secondary-like rays through the 8-wide tree, traversed on the CPU."""
        out = np.zeros(8, np.uint64)
        self._c("debug_wide_cost", n_rays, seed, _p(out, C.c_uint64))
        keys = ("rays", "node_visits", "prim_tests", "occupied_slots", "hits", "wide_nodes", "wide_depth", "bvh_prims")
        return dict(zip(keys, (int(x) for x in out)))

    def render(self, cam, params, out=None):
        """rtw_render: per-pixel radiance SUM, H x W x 3 float32, row 0 = top.  Returns (image, stats dict)."""
        if out is None:
            out = np.zeros((params.height, params.width, 3), np.float32)
        st = Stats()
        self._c("render", C.byref(cam), C.byref(params), out.ctypes.data_as(C.c_void_p), C.byref(st))
        return out, st.as_dict()

    def render_progressive(self, cam, params, first_sample=0, samples_per_pass=0, buf=None, progress=None):
        """rtw_render_progressive: samples [first_sample, params.spp) in passes; `buf` carries the sums of the samples
        below first_sample (resume).  progress(done, total, image_view) -> truthy to stop.  Returns (image, stats)."""
        if buf is None:
            assert first_sample == 0, "resuming needs the buffer of the interrupted render"
            buf = np.zeros((params.height, params.width, 3), np.float32)
        assert buf.dtype == np.float32 and buf.shape == (params.height, params.width, 3) and buf.flags.c_contiguous
        FN = C.CFUNCTYPE(C.c_int, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p)
        def _cb(done, total, _ptr, _user):
            return 1 if (progress is not None and progress(done, total, buf)) else 0
        cb = FN(_cb)
        st = Stats()
        self._c("render_progressive", C.byref(cam), C.byref(params), C.c_int32(first_sample), C.c_int32(samples_per_pass),
                buf.ctypes.data_as(C.c_void_p), C.cast(cb, C.c_void_p), None, C.byref(st))
        return buf, st.as_dict()

    def render_device(self, cam, params, dev_ptr):
        """rtw_render with RTW_FLAG_DEVICE_OUT: the sums stay in HBM at dev_ptr (no D2H)."""
        st = Stats()
        self._c("render", C.byref(cam), C.byref(params), C.c_void_p(dev_ptr), C.byref(st))
        return st.as_dict()

    # --- parity hooks
    def test_hit(self, target, origin, direction, time=None, t_min=0.001, t_max=float("inf"), xi=None):
        o, d = _f64(origin), _f64(direction)
        n = len(o)
        tm = _f64(time) if time is not None else np.zeros(n)
        if xi is None:
            xi = np.full((n, 1), 0.5)
        xi = _f64(xi)
        hit, front, mat, nd = (np.zeros(n, np.int32) for _ in range(4))
        t, u, v = np.zeros(n), np.zeros(n), np.zeros(n)
        p, nrm = np.zeros((n, 3)), np.zeros((n, 3))
        self._c("test_hit", target, n, _p(o), _p(d), _p(tm), t_min, t_max, _p(xi), xi.shape[1],
                _p(hit, C.c_int32), _p(t), _p(p), _p(nrm), _p(front, C.c_int32), _p(u), _p(v),
                _p(mat, C.c_int32), _p(nd, C.c_int32))
        return dict(hit=hit, t=t, p=p, normal=nrm, front=front, u=u, v=v, mat=mat, ndraw=nd)

    def test_scatter(self, mat, ray_o, ray_d, ray_t, p, normal, front, u, v, xi):
        ro, rd, pp, nn, xi = _f64(ray_o), _f64(ray_d), _f64(p), _f64(normal), _f64(xi)
        n = len(ro)
        rt = _f64(ray_t) if ray_t is not None else np.zeros(n)
        uu = _f64(u) if u is not None else np.zeros(n)
        vv = _f64(v) if v is not None else np.zeros(n)
        fr = _i32(front)
        sc, nd = np.zeros(n, np.int32), np.zeros(n, np.int32)
        oo, od, ot = np.zeros((n, 3)), np.zeros((n, 3)), np.zeros(n)
        att, em = np.zeros((n, 3)), np.zeros((n, 3))
        self._c("test_scatter", mat, n, _p(ro), _p(rd), _p(rt), _p(pp), _p(nn), _p(fr, C.c_int32), _p(uu), _p(vv),
                _p(xi), xi.shape[1], _p(sc, C.c_int32), _p(oo), _p(od), _p(ot), _p(att), _p(em), _p(nd, C.c_int32))
        return dict(scattered=sc, origin=oo, dir=od, time=ot, attenuation=att, emitted=em, ndraw=nd)

    def test_texture(self, tex, u, v, p):
        pp = _f64(p)
        n = len(pp)
        uu = _f64(u) if u is not None else np.zeros(n)
        vv = _f64(v) if v is not None else np.zeros(n)
        out = np.zeros((n, 3))
        self._c("test_texture", tex, n, _p(uu), _p(vv), _p(pp), _p(out))
        return out

    def trace_paths(self, cam, params, px, py, sample):
        px, py, sample = _i32(px), _i32(py), _i32(sample)
        n = len(px)
        rgb, seg = np.zeros((n, 3)), np.zeros(n, np.int32)
        self._c("trace_paths", C.byref(cam), C.byref(params), n, _p(px, C.c_int32), _p(py, C.c_int32),
                _p(sample, C.c_int32), _p(rgb), _p(seg, C.c_int32))
        return rgb, seg
