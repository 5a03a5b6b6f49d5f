"""Freezes the ALGORITHMIC flops/path of each config: event counts measured by the CPU oracle on the
same scene/seed with the world wrapped in a reference-style binary BVH (src/hittable.rs:77-130), times the fixed
per-event costs of the reference arithmetic.  Output: profiles/flop_model.json (read by bench.py)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load()
from oracle.binding import load_oracle
orc = load_oracle()

COST = dict(raygen=52, aabb=27, sphere=24, sphere_accept=34, moving=12, rect=12, rect_accept=21, translate=15, rotate=33,
            medium=25, scatter=[42, 57, 56, 0, 23], tex=[0, 6, 1775, 14], accum=6)


def flops(c):
    f = COST["raygen"] * c["paths"] + COST["aabb"] * c["aabb"] + COST["sphere"] * c["sphere"] + COST["sphere_accept"] * c["sphere_accept"]
    f += COST["moving"] * c["moving"] + COST["rect"] * c["rect"] + COST["rect_accept"] * c["rect_accept"]
    f += COST["translate"] * c["translate"] + COST["rotate"] * c["rotate"] + COST["medium"] * c["medium"] + COST["accum"] * c["accum"]
    f += sum(a * b for a, b in zip(COST["scatter"], c["scatter"])) + sum(a * b for a, b in zip(COST["tex"], c["tex"]))
    return f


out = {"costs": COST, "note": "oracle event counts, reference-style BVH over the world (wrap_bvh), seed 1", "configs": {}}
for name, (W, H, spp) in {"random_scene": (300, 200, 32), "two_spheres": (200, 112, 32), "two_perlin_spheres": (200, 112, 32),
                          "earth": (200, 112, 32), "simple_light": (150, 150, 32), "cornell_box": (150, 150, 32),
                          "cornell_box_smoke": (150, 150, 32), "final_scene": (160, 160, 32)}.items():
    sc, spec = m.scenes.build(orc, name, wrap_bvh=True)
    sc.set_media_deferred(True)
    p = m.make_params(W, H, spp, background=spec.background)
    r = sc.render_oracle(spec.camera(orc, W, H), p, threads=0, counters=True)
    c = r["counters"]
    out["configs"][name] = dict(paths=c["paths"], rays_per_path=c["rays"] / c["paths"], aabb_per_ray=c["aabb"] / c["rays"],
                                sphere_per_ray=c["sphere"] / c["rays"], rect_per_ray=c["rect"] / c["rays"],
                                draws_per_path=c["draws"] / c["paths"], flops_per_path=flops(c) / c["paths"], counters=c,
                                oracle_bvh_mpaths_s=c["paths"] / r["seconds"] / 1e6)
    print(name, {k: round(v, 2) for k, v in out["configs"][name].items() if isinstance(v, float)})
json.dump(out, open(os.path.join(os.path.dirname(__file__), "..", "profiles", "flop_model.json"), "w"), indent=1)
