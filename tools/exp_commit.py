"""Commit time of a big rtw_sphere_batch scene with and without the staged upload (csrc/bvh_build.cu staged_upload).
Usage: exp_commit.py [N_millions ...]   (RTW_TIMING=1 prints the builder's phases)"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
for n in [int(float(a) * (1 << 20)) for a in sys.argv[1:]] or [16 << 20]:
    sc = m.Scene(rtw)
    spec = m.scenes.sweep_scene(sc, n)
    for sw in ("0", "1", "0", "1", "1"):
        os.environ["RTW_STAGED_UPLOAD"] = sw
        t0 = time.time(); sc.commit(1, 0); t1 = time.time()
        print(f"[commit] {n} spheres staged={sw}: {1e3 * (t1 - t0):.1f} ms", flush=True)
    sc.close()
