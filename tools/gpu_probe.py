"""First-contact GPU probe: parity spot checks + timing.  Run under gpurun."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtw_pkg
m = rtw_pkg.load()
rtw, orc = m.load_rtw(), m.api.load_oracle()
print("devices", rtw.f("device_count")())
rs = np.random.RandomState(0)

# philox
ctr = rs.randint(0, 2**32, (1000, 4), dtype=np.uint64).astype(np.uint32); key = rs.randint(0, 2**32, (1000, 2), dtype=np.uint64).astype(np.uint32)
print("philox equal:", np.array_equal(rtw.philox(ctr, key), orc.philox(ctr, key)))

def both(name, **kw):
    a, spec = m.scenes.build(rtw, name, **kw); b, _ = m.scenes.build(orc, name, **kw)
    b.set_media_deferred(True)
    a.commit(1, 0)
    return a, b, spec

for name in ["random_scene", "cornell_box", "final_scene", "cornell_box_smoke", "two_perlin_spheres", "earth"]:
    a, b, spec = both(name)
    W, H = 96, 64
    cam_a, cam_b = spec.camera(rtw, W, H), spec.camera(orc, W, H)
    n = 20000
    px = rs.randint(0, W, n); py = rs.randint(0, H, n); sm = rs.randint(0, 64, n)
    p = m.make_params(W, H, 64, background=spec.background)
    t0 = time.time(); ra, sa = a.trace_paths(cam_a, p, px, py, sm); t1 = time.time(); rb, sb = b.trace_paths(cam_b, p, px, py, sm); t2 = time.time()
    err = np.abs(ra - rb).max(1)
    tol = 1e-3 * np.maximum(1.0, np.abs(rb).max(1))
    print(f"{name}: paths match {np.mean(err <= tol):.4f}, seg equal {np.mean(sa == sb):.4f}, mean gpu {ra.mean():.5f} orc {rb.mean():.5f}  (gpu {t1-t0:.2f}s orc {t2-t1:.2f}s)")
    # camera-ray hit parity (world, BVH)
    s = rs.rand(n); t = rs.rand(n); xi = (rs.randint(0, 2**24, (n, 16)) / 2.0**24)
    ga = rtw.test_get_ray(cam_a, s, t, xi); gb = orc.test_get_ray(cam_b, s, t, xi)
    print("   get_ray max rel err o,d:", np.abs(ga['origin'] - gb['origin']).max(), (np.abs(ga['dir'] - gb['dir']).max(1) / np.linalg.norm(gb['dir'], axis=1)).max(), "ndraw eq", np.array_equal(ga['ndraw'], gb['ndraw']))
    o32 = ga['origin']; d32 = ga['dir']; tm = ga['time']
    xi2 = (rs.randint(1, 2**24, (n, 4)) / 2.0**24)
    ha = a.test_hit(-1, o32, d32, tm, xi=xi2); hb = b.test_hit(-1, o32, d32, tm, xi=xi2)
    same = ha['hit'] == hb['hit']
    both_hit = same & (ha['hit'] == 1)
    rel_t = np.abs(ha['t'] - hb['t'])[both_hit] / np.maximum(np.abs(hb['t'][both_hit]), 1e-3)
    print(f"   hit flag agree {same.mean():.5f}; hits {both_hit.sum()}; t rel err max {rel_t.max() if len(rel_t) else 0:.2e} p99 {np.percentile(rel_t, 99) if len(rel_t) else 0:.2e}; mat agree {np.mean(ha['mat'][both_hit] == hb['mat'][both_hit]):.5f}; normal max err {np.abs(ha['normal'] - hb['normal'])[both_hit].max() if both_hit.any() else 0:.2e}; front agree {np.mean(ha['front'][both_hit]==hb['front'][both_hit]):.5f}")

# small full render parity + timing
a, b, spec = both("random_scene")
W, H, spp = 240, 160, 64
p = m.make_params(W, H, spp, background=spec.background)
img, st = a.render(spec.camera(rtw, W, H), p)
ro = b.render_oracle(spec.camera(orc, W, H), p, threads=0)
diff = img / spp - ro['sum'] / spp
print("render small: mean gpu", img.mean() / spp, "orc", ro['sum'].mean() / spp, "rmse", np.sqrt((diff**2).mean()), "stats", st)
from PIL import Image
os.makedirs("gpurun_out", exist_ok=True)
Image.fromarray((np.sqrt(np.clip(img / spp, 0, None)).clip(0, 0.999) * 256).astype(np.uint8)).save("gpurun_out/probe_random_scene.png")

# full size timing C1
for spp in (50, 500):
    p = m.make_params(1200, 800, spp, background=spec.background)
    for rep in range(3):
        img, st = a.render(spec.camera(rtw, 1200, 800), p)
        print(f"C1 {spp}spp: ms_render {st['ms_render']:.2f} total {st['ms_total']:.2f} Mpaths/s {st['paths']/st['ms_render']/1e3:.1f} rays/path {st['rays']/st['paths']:.3f} units {st['units_per_device'][0]}")
Image.fromarray((np.sqrt(np.clip(img / spp, 0, None)).clip(0, 0.999) * 256).astype(np.uint8)).save("gpurun_out/probe_c1_500.png")
for name, (W, H, spp) in {"cornell_box": (600, 600, 100), "final_scene": (800, 800, 100), "two_perlin_spheres": (800, 450, 200), "earth": (800, 450, 200), "cornell_box_smoke": (600, 600, 100)}.items():
    a, b, spec = both(name)
    p = m.make_params(W, H, spp, background=spec.background)
    for rep in range(2):
        img, st = a.render(spec.camera(rtw, W, H), p)
    print(f"{name} {W}x{H}x{spp}: ms_render {st['ms_render']:.2f} Mpaths/s {st['paths']/st['ms_render']/1e3:.1f} rays/path {st['rays']/st['paths']:.3f}")
    Image.fromarray((np.sqrt(np.clip(img / spp, 0, None)).clip(0, 0.999) * 256).astype(np.uint8)).save(f"gpurun_out/probe_{name}.png")
