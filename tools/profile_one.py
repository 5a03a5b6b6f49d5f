"""Short single-GPU command for ncu: a few C1 renders at reduced spp (same kernel, same scene, same image size)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load()
rtw = m.load_rtw()
name = sys.argv[1] if len(sys.argv) > 1 else "random_scene"
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 50
sc, spec = m.scenes.build(rtw, name)
sc.commit(1, 0)
W, H = (1200, 800) if name == "random_scene" else (spec.width, spec.height)
for i in range(4):
    img, st = sc.render(spec.camera(rtw, W, H), m.make_params(W, H, spp, background=spec.background))
    print(name, W, H, spp, f"ms {st['ms_render']:.2f} Mpaths/s {st['paths'] / st['ms_render'] / 1e3:.1f}")
