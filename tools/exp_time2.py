"""Timing at closer-to-config sample counts (amortises per-unit work)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
tag = os.environ.get("RTW_TAG", "lib")
only = os.environ.get("RTW_ONLY", "").split(",") if os.environ.get("RTW_ONLY") else None     # e.g. RTW_ONLY=final_scene,earth
for name, W, H, spp in [("final_scene", 800, 800, 1000), ("cornell_box", 600, 600, 1000), ("simple_light", 600, 600, 1000), ("two_perlin_spheres", 800, 450, 200), ("earth", 800, 450, 200), ("two_spheres", 800, 450, 200), ("cornell_box_smoke", 600, 600, 1000)]:
    if only and name not in only: continue
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    cam = spec.camera(rtw, W, H); best = 1e9
    for i in range(3):
        img, st = sc.render(cam, m.make_params(W, H, spp, background=spec.background)); best = min(best, st["ms_render"])
    print(f"[{tag}] {name} {W}x{H}x{spp}: {best:.2f} ms  {W*H*spp/best/1e3:.1f} Mpaths/s rays/path {st['rays']/st['paths']:.3f}", flush=True)
