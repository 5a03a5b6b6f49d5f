"""BASELINE config 5: scaling sweep, N random spheres under the BVH, 3840x2160, 256 spp.  Reports commit
(flatten + BVH build + upload) and render separately.  Usage: sweep.py [N_millions ...] [--spp S] [--gpus G]"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
args = sys.argv[1:]
spp = 256; gpus = 1
if "--spp" in args: i = args.index("--spp"); spp = int(args[i + 1]); del args[i:i + 2]
if "--gpus" in args: i = args.index("--gpus"); gpus = int(args[i + 1]); del args[i:i + 2]
sizes = [int(float(a) * (1 << 20)) for a in args] or [1 << 20]
W, H = 3840, 2160
out = []
for n in sizes:
    sc = m.Scene(rtw)
    t0 = time.time(); spec = m.scenes.sweep_scene(sc, n); t1 = time.time()
    sc.commit(gpus, 0); t2 = time.time()
    cam = spec.camera(rtw, W, H)
    best = None
    for rep in range(2):
        img, st = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=1))
        if best is None or st["ms_render"] < best["ms_render"]: best = st
    rec = dict(n_spheres=n, gpus=gpus, spp=spp, build_scene_s=round(t1 - t0, 2), commit_s=round(t2 - t1, 2), ms_render=round(best["ms_render"], 2),
               mpaths_s=round(best["paths"] / best["ms_render"] / 1e3, 1), mrays_s=round(best["rays"] / best["ms_render"] / 1e3, 1),
               rays_per_path=round(best["rays"] / best["paths"], 3), nodes=best["n_nodes"], image_mean=float(img.mean() / spp),
               scene_bytes=best["n_nodes"] * 64 + best["n_prims"] * 64)
    out.append(rec); print(json.dumps(rec), flush=True)
    sc.close()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(f"gpurun_out/sweep_g{gpus}.json", "w"), indent=1)
