"""roofline.frac_device_counts: the reference-arithmetic cost table of tools/flop_model.py applied to the work THIS kernel
executes (its own BVH and tile lists) instead of the reference-style BVH the oracle counts.  Needs the instrumented
library: make -C rust-ray-tracing-in-a-weekend_b200/csrc instr, RTW_LIB_PATH=.../variants/instr.so.  Writes
profiles/flop_model.json["device_counts"]."""
import ctypes as C, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
path = os.path.join(ROOT, "profiles", "flop_model.json")
model = json.load(open(path))
COST = model["costs"]
out = {}
for name, W, H, spp in [("random_scene", 1200, 800, 50), ("cornell_box", 600, 600, 50), ("final_scene", 800, 800, 20)]:
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    cnt = (C.c_ulonglong * 8)()
    rtw.dll.rtw_debug_counters(cnt, 1)
    img, st = sc.render(spec.camera(rtw, W, H), m.make_params(W, H, spp, background=spec.background))
    rtw.dll.rtw_debug_counters(cnt, 1)
    it, alive, visits, vmax, prims, pmax, prim_primary, n_primary = [int(x) for x in cnt]
    paths = st["paths"]
    oc = model["configs"][name]["counters"]; op = oc["paths"]
    # secondary rays: node visits x 2 boxes, primitive tests; primary rays: the tile's candidate list per path (stats)
    list_mean = prim_primary / max(n_primary, 1)
    sec_box, sec_prim = 2.0 * visits / paths, prims / paths
    prim_tests = sec_prim + prim_primary / paths
    moving_share = oc["moving"] / max(oc["sphere"], 1)
    rect_share = oc["rect"] / max(oc["sphere"] + oc["rect"], 1)
    f = COST["raygen"] + COST["aabb"] * sec_box
    f += prim_tests * ((1 - rect_share) * (COST["sphere"] + COST["moving"] * moving_share) + rect_share * COST["rect"])
    f += (COST["sphere_accept"] * oc["sphere_accept"] + COST["rect_accept"] * oc["rect_accept"] + COST["translate"] * oc["translate"] + COST["rotate"] * oc["rotate"]
          + COST["medium"] * oc["medium"] + COST["accum"] * oc["accum"] + sum(a * b for a, b in zip(COST["scatter"], oc["scatter"]))
          + sum(a * b for a, b in zip(COST["tex"], oc["tex"]))) / op
    out[name] = dict(flops_per_path=f, box_tests_per_path=sec_box, prim_tests_per_path=prim_tests, node_visits_per_secondary_ray=visits / max(alive, 1),
                     lanes_alive_per_iteration=alive / max(it, 1), traversal_lane_efficiency=visits / max(32 * vmax, 1), rays_per_path=st["rays"] / paths,
                     tile_list_mean=list_mean, what=f"{W}x{H}x{spp}, instrumented build")
    print(name, json.dumps(out[name]))
model["device_counts"] = out
json.dump(model, open(path, "w"), indent=1)
