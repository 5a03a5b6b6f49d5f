"""Warp timeline of one launch (instrumented build): how long after the AVERAGE warp does the LAST warp finish (the tail of
the frame), and how spread are the starts.  RTW_LIB_PATH=.../variants/instr.so [RTW_EMULATE_RANKS=8] python tools/diag_tail.py"""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
for name, W, H, spp in [("random_scene", 1200, 800, 500), ("final_scene", 800, 800, 200)]:
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    out = (C.c_ulonglong * 6)()
    for rep in range(3):
        rtw.dll.rtw_debug_timeline(out)
        img, st = sc.render(spec.camera(rtw, W, H), m.make_params(W, H, spp, background=spec.background))
        rtw.dll.rtw_debug_timeline(out)
        t0, t1s, tsum, tend, n, tfirst = [int(x) for x in out]
        mean_end = tsum / max(n, 1)
        print(f"{name} {W}x{H}x{spp}: kernel {st['ms_render']:.2f} ms (instrumented); warps {n}; starts spread {(t1s - t0) / 1e6:.3f} ms; "
              f"first warp done at {(tfirst - t0) / 1e6:.2f} ms, mean {(mean_end - t0) / 1e6:.2f} ms, last {(tend - t0) / 1e6:.2f} ms "
              f"-> tail (last - mean) {(tend - mean_end) / 1e6:.3f} ms, idle share {(tend - mean_end) / max(tend - t0, 1) * 100:.2f} %", flush=True)
