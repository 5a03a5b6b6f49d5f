"""Host-only: cost of secondary-like rays through the 8-wide tree of the sweep scene (collapse heuristic A/B: RTW_WIDE_FILL=0|1).
Usage: wide_cost.py [N_millions] [n_rays]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
n = int(float(sys.argv[1]) * (1 << 20)) if len(sys.argv) > 1 else 1 << 18
rays = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
sc = m.Scene(rtw); m.scenes.sweep_scene(sc, n)
d = sc.debug_wide_cost(rays, 1)
print({**d, "visits_per_ray": round(d["node_visits"] / d["rays"], 2), "prim_tests_per_ray": round(d["prim_tests"] / d["rays"], 2),
       "slots_per_visit": round(d["occupied_slots"] / d["node_visits"], 2), "prims_per_node": round(d["bvh_prims"] / d["wide_nodes"], 2)})
