"""ctypes loader of the CPU oracle (oracle/liboracle.so) — TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference arm import this module.  The product
package (rust-ray-tracing-in-a-weekend_b200/) neither imports it nor knows where the oracle lives: it cannot fall back
to the CPU.  The oracle exports the scene constructors and parity hooks of include/rtw.h under the `orc_` prefix, so the
product's generic `Lib` / `Scene` bindings drive it; what only the oracle has (the reference-style multi-threaded CPU
render with event counters, the closed-form helpers used as known answers) is added here.
"""
import ctypes as C
import os

import numpy as np

import rtw_pkg

api = rtw_pkg.load().api
ORACLE_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "liboracle.so")


class OrcCounters(C.Structure):
    """Event counts of one oracle render (feeds tools/flop_model.py)."""
    _fields_ = [(n, C.c_uint64) for n in ("paths", "rays", "aabb", "sphere", "sphere_accept", "moving", "rect",
                                           "rect_accept", "translate", "rotate", "medium")] + \
               [("scatter", C.c_uint64 * 5), ("tex", C.c_uint64 * 4), ("accum", C.c_uint64), ("draws", C.c_uint64)]

    def as_dict(self):
        d = {}
        for name, _ in self._fields_:
            v = getattr(self, name)
            d[name] = list(v) if hasattr(v, "__len__") else v
        return d


class OracleScene(api.Scene):
    def render_oracle(self, cam, params, threads=0, counters=False, sumsq=False):
        """orc_render: the reference's worker loop (src/main.rs:507-548) in f64 on `threads` host threads."""
        H, W = params.height, params.width
        out = np.zeros((H, W, 3))
        sq = np.zeros((H, W, 3)) if sumsq else None
        cnt = OrcCounters() if counters else None
        secs = C.c_double(0)
        self._c("render", C.byref(cam), C.byref(params), threads, api._p(out), api._p(sq),
                C.byref(cnt) if counters else None, C.byref(secs))
        return dict(sum=out, sumsq=sq, counters=cnt.as_dict() if counters else None, seconds=secs.value)

    def set_media_deferred(self, on):
        self._c("scene_set_media_deferred", 1 if on else 0)


class OracleLib(api.Lib):
    scene_cls = OracleScene
    is_oracle = True

    def _sig_extra(self):
        f = self.f
        vp, dp = C.c_void_p, C.POINTER(C.c_double)
        f("render").argtypes = [vp, C.POINTER(api.Camera), C.POINTER(api.RenderParams), C.c_int32, dp, dp,
                                C.POINTER(OrcCounters), dp]
        f("write_color").argtypes = [dp, C.c_int32, C.c_int32, C.POINTER(C.c_uint8)]
        f("scene_set_media_deferred").argtypes = [vp, C.c_int]
        f("scene_set_build_seed").argtypes = [vp, C.c_uint64]
        f("bounding_box").argtypes = [vp, C.c_int, C.c_double, C.c_double, dp, dp]
        f("sphere_uv").argtypes = [C.c_int32, dp, dp, dp]
        f("reflectance").argtypes = [C.c_int32, dp, dp, dp]
        f("refract").argtypes = [dp, dp, C.c_double, dp]
        f("reflect").argtypes = [dp, dp, dp]
        f("perlin_noise").argtypes = [vp, C.c_int, C.c_int32, dp, dp, dp]
        f("world_clear").argtypes = [vp]


_lib = None


def load_oracle():
    global _lib
    if _lib is None:
        _lib = OracleLib(ORACLE_LIB_PATH, "orc_")
    return _lib
