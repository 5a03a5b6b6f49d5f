"""Kernel experiment driver: time the configs with the library named by RTW_LIB_PATH and check path parity."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtw_pkg
m = rtw_pkg.load()
rtw, orc = m.load_rtw(), m.api.load_oracle()
tag = os.environ.get("RTW_TAG", os.path.basename(m.api.RTW_LIB_PATH))
cfgs = {"random_scene": (1200, 800, 500), "cornell_box": (600, 600, 200), "final_scene": (800, 800, 100), "two_perlin_spheres": (800, 450, 200)}
only = sys.argv[1:] or list(cfgs)
res = {}
for name in only:
    W, H, spp = cfgs[name]
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    cam = spec.camera(rtw, W, H)
    best = 1e9
    for i in range(4):
        img, st = sc.render(cam, m.make_params(W, H, spp, background=spec.background))
        best = min(best, st["ms_render"])
    b, _ = m.scenes.build(orc, name, wrap_bvh=name not in ("final_scene",)); b.set_media_deferred(True)
    rs = np.random.RandomState(1); n = 20000
    px, py, sm = rs.randint(0, 96, n), rs.randint(0, 64, n), rs.randint(0, 64, n)
    p = m.make_params(96, 64, 64, background=spec.background)
    ra, sa = sc.trace_paths(spec.camera(rtw, 96, 64), p, px, py, sm); rb, sb = b.trace_paths(spec.camera(orc, 96, 64), p, px, py, sm)
    good = np.mean((sa == sb) & (np.abs(ra - rb).max(1) <= 1e-3 * np.maximum(1, np.abs(rb).max(1))))
    res[name] = dict(ms=round(best, 2), mpaths=round(W * H * spp / best / 1e3, 1), match=round(float(good), 4), mean=float(img.mean() / spp))
    print(f"[{tag}] {name}: {best:.2f} ms  {W*H*spp/best/1e3:.1f} Mpaths/s  path-match {good:.4f}  rays/path {st['rays']/st['paths']:.3f}", flush=True)
