import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
tag = os.environ.get("RTW_KERNEL", "default")
for name, W, H, spp in [("final_scene", 100, 100, 32), ("cornell_box_smoke", 96, 96, 64), ("random_scene", 203, 117, 48), ("earth", 120, 67, 32)]:
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    cam = spec.camera(rtw, W, H)
    a, s0 = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=5, flags=2))
    b, s1 = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=5))
    d = np.abs(a - b).max(2)
    print(f"[{tag}] {name}: rays mega {s0['rays']} other {s1['rays']} max diff {d.max():.2e} mean {a.mean()/spp:.5f} {b.mean()/spp:.5f}")
