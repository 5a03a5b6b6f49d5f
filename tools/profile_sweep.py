"""Short single-GPU command for ncu: C5 sweep scene (N spheres) at 1920x1080, few spp."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
n = int(float(sys.argv[1]) * (1 << 20)) if len(sys.argv) > 1 else 1 << 20
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 8
sc = m.Scene(rtw); spec = m.scenes.sweep_scene(sc, n); sc.commit(1, 0)
W, H = 1920, 1080
for i in range(3):
    img, st = sc.render(spec.camera(rtw, W, H), m.make_params(W, H, spp, background=spec.background))
    print(n, W, H, spp, f"ms {st['ms_render']:.2f} Mpaths/s {st['paths'] / st['ms_render'] / 1e3:.1f} Mrays/s {st['rays'] / st['ms_render'] / 1e3:.1f}")
