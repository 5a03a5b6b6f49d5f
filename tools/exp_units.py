"""What do small work units cost, and how much of that is the per-unit tile list?  C1 at full frame size with the default and the
8-GPU unit sizes (RTW_UNITS_PER_WARP=192), with and without tile culling (RTW_FLAG_NO_TILE_CULL).  Env is read by the library
per render, so every combination runs in its own process: exp_units.py <units_per_warp|0> <nocull 0|1>"""
import os, sys
upw, nocull = sys.argv[1], int(sys.argv[2])
if upw != "0": os.environ["RTW_UNITS_PER_WARP"] = upw
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
sc, spec = m.scenes.build(rtw, "random_scene"); sc.commit(1, 0)
W, H, spp = 1200, 800, 500
flags = m.api.RTW_FLAG_KERNEL_MEGA | (m.api.RTW_FLAG_NO_TILE_CULL if nocull else 0)
best = 1e9
for i in range(4):
    img, st = sc.render(spec.camera(rtw, W, H), m.make_params(W, H, spp, background=spec.background, flags=flags)); best = min(best, st["ms_render"])
print(f"units_per_warp {upw} nocull {nocull}: {best:.2f} ms")
