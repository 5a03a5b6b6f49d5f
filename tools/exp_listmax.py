"""Where does scanning all primitives stop paying against the BVH walk?  n random spheres (sweep_scene layout), scan forced
on / off through RTW_LIST_MAX (read per render).  exp_listmax.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
W, H, spp = 1200, 800, 100
for n in (2, 4, 6, 8, 10, 12, 16, 24, 32):
    sc = m.Scene(rtw); spec = m.scenes.sweep_scene(sc, n, seed=3); sc.commit(1, 0)
    cam = spec.camera(rtw, W, H); res = []
    for lm in ("0", "64"):
        os.environ["RTW_LIST_MAX"] = lm
        best = 1e9
        for i in range(3):
            img, st = sc.render(cam, m.make_params(W, H, spp, background=spec.background)); best = min(best, st["ms_render"])
        res.append(best)
    print(f"n {n:3d}: BVH walk {res[0]:7.2f} ms   scan {res[1]:7.2f} ms   rays/path {st['rays'] / st['paths']:.2f}", flush=True)
    sc.close()
