import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
for name, W, H, spp in [("final_scene", 100, 100, 32), ("cornell_box_smoke", 96, 96, 64), ("random_scene", 203, 117, 48), ("cornell_box", 96, 96, 64)]:
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    cam = spec.camera(rtw, W, H)
    a, s0 = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=5, flags=2))
    a2, s00 = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=5, flags=2))
    b, s1 = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=5, flags=4))
    d = np.abs(a - b).max(2)
    print(name, "rays mega", s0["rays"], s00["rays"], "pool", s1["rays"], "max diff", d.max(), "n pix diff>1e-3", int((d > 1e-3).sum()), "mega-vs-mega", np.abs(a - a2).max(),
          "mean", a.mean() / spp, b.mean() / spp)
