"""Generates tests/golden/oracle_paths_v2.npz: per-path radiance of fixed (pixel, sample) paths for every reference
scene, computed by the CPU oracle (media_deferred order = the device's).  The reference itself cannot run here
(Rust toolchain absent), so this fixture pins the ORACLE, and gives the GPU tests a committed target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import rtw_pkg
m = rtw_pkg.load()
from oracle.binding import load_oracle
orc = load_oracle()
rs = np.random.RandomState(2024)
W, H, n, seed = 96, 64, 1500, 7
px, py, sm = rs.randint(0, W, n), rs.randint(0, H, n), rs.randint(0, 64, n)
out = dict(size=np.array([W, H]), seed=np.array(seed), px=px.astype(np.int32), py=py.astype(np.int32), sample=sm.astype(np.int32))
for name in m.scenes.SCENES:
    sc, spec = m.scenes.build(orc, name, wrap_bvh=name not in ("cornell_box_smoke", "final_scene"))
    sc.set_media_deferred(True)
    p = m.make_params(W, H, 64, background=spec.background, seed=seed)
    rgb, seg = sc.trace_paths(spec.camera(orc, W, H), p, px, py, sm)
    out[name + "_rgb"], out[name + "_seg"] = rgb, seg
    print(name, rgb.mean(), seg.mean())
np.savez_compressed(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "oracle_paths_v2.npz"), **out)
