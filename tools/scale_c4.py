"""BASELINE config 4 (final_scene 800x800, 10 000 spp, depth 50 = 6.4 G paths) on 1 / 2 / 4 / 8 GPUs of one box, in-process
(rtw_scene_commit(n) + rtw_render(n_gpus=k): one scene replica per GPU, tiles x sample-chunks from one counter, finished
tiles added into GPU 0's framebuffer over NVLink).  Usage: scale_c4.py [max_gpus] [spp]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
nmax = int(sys.argv[1]) if len(sys.argv) > 1 else 8
spp = int(sys.argv[2]) if len(sys.argv) > 2 else 10000
sc, spec = m.scenes.build(rtw, "final_scene")
sc.commit(nmax, 0)
W = H = 800
cam = spec.camera(rtw, W, H)
out = []
sc.render(cam, m.make_params(W, H, 64, background=spec.background, seed=1, n_gpus=nmax))      # warm-up: peer mappings, kernels
base = None
for n in (1, 2, 4, 8):
    if n > nmax: break
    img, st = sc.render(cam, m.make_params(W, H, spp, background=spec.background, seed=1, n_gpus=n))
    base = base or st["ms_render"]
    rec = dict(config="C4 final_scene 800x800", spp=spp, n_gpus=n, ms_render=round(st["ms_render"], 2), ms_total=round(st["ms_total"], 2),
               mpaths_s=round(st["paths"] / st["ms_render"] / 1e3, 1), speedup=round(base / st["ms_render"], 3),
               units_per_device=[int(u) for u in st["units_per_device"][:n]], image_mean=float(img.mean() / spp))
    out.append(rec); print(json.dumps(rec), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/scale_c4.json", "w"), indent=1)
