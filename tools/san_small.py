"""Small renders for compute-sanitizer: every schedule and the new code paths at sizes a sanitizer run finishes in a minute.
compute-sanitizer --tool memcheck python tools/san_small.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
A = m.api
for name, W, H, spp in [("random_scene", 97, 61, 6), ("cornell_box", 64, 64, 6), ("cornell_box_smoke", 48, 48, 4), ("final_scene", 48, 48, 4)]:
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    cam = spec.camera(rtw, W, H)
    for flags, spu in [(A.RTW_FLAG_KERNEL_MEGA, 0), (A.RTW_FLAG_KERNEL_MEGA, 1), (A.RTW_FLAG_KERNEL_WAVEFRONT, 0)]:
        img, st = sc.render(cam, m.make_params(W, H, spp, background=spec.background, samples_per_unit=spu, flags=flags))
        assert np.isfinite(img).all()
        print(name, flags, spu, st["rays"], float(img.mean()), flush=True)
    sc.close()
os.environ["RTW_BIG_MIN"] = "1000"                       # device build + 8-wide nodes + wavefront on a small sweep scene
sc = m.Scene(rtw); spec = m.scenes.sweep_scene(sc, 20000, seed=3); sc.commit(1, 0)
img, st = sc.render(spec.camera(rtw, 160, 90), m.make_params(160, 90, 4, background=spec.background))
print("sweep_20000 device-built wide + wavefront", st["rays"], st["n_nodes"], float(img.mean()))
