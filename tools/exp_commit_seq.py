"""The bench's big-scene sequence (1 M -> close -> 4 M -> close -> 16 M -> close, twice) with the commit's phases (RTW_TIMING=1)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
for rep in range(2):
    for n in (1 << 20, 4 << 20, 16 << 20):
        sc = m.Scene(rtw)
        spec = m.scenes.sweep_scene(sc, n)
        t0 = time.time(); sc.commit(1, 0); t1 = time.time()
        print(f"[seq] rep {rep} {n} spheres: commit {1e3 * (t1 - t0):.1f} ms", flush=True)
        if os.environ.get("SEQ_RENDER"):             # a 4K wavefront render in between, like the bench
            spp = int(os.environ["SEQ_RENDER"])
            img, st = sc.render(spec.camera(rtw, 3840, 2160), m.make_params(3840, 2160, spp, background=spec.background, seed=1))
            print(f"[seq] rep {rep} {n} spheres: render {st['ms_render']:.1f} ms", flush=True)
        t0 = time.time(); sc.close(); t1 = time.time()
        print(f"[seq] rep {rep} {n} spheres: close {1e3 * (t1 - t0):.1f} ms", flush=True)
