"""Lane-efficiency diagnosis: needs the instrumented library (make -C rust-ray-tracing-in-a-weekend_b200/csrc instr) and
RTW_LIB_PATH=.../variants/instr.so.  Prints, per warp iteration, live lanes and mean vs max traversal length."""
import os, sys, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw = m.load_rtw()
for name, W, H, spp in [("random_scene", 1200, 800, 50), ("cornell_box", 600, 600, 40), ("final_scene", 800, 800, 20)]:
    sc, spec = m.scenes.build(rtw, name); sc.commit(1, 0)
    out = (C.c_ulonglong * 8)()
    rtw.dll.rtw_debug_counters(out, 1)
    img, st = sc.render(spec.camera(rtw, W, H), m.make_params(W, H, spp, background=spec.background))
    rtw.dll.rtw_debug_counters(out, 1)
    it, alive, vs, vm, ps, pm = [int(x) for x in out][:6]
    print(f"{name}: warp-iterations {it:.3e} alive/iter {alive/it:.2f} node visits: sum/iter {vs/it:.1f} max/iter {vm/it:.1f} -> lane eff {vs/(32*vm):.3f} (mean per alive lane {vs/alive:.2f}); prim tests sum/iter {ps/it:.1f} max/iter {pm/it:.2f} -> eff {ps/(32*pm):.3f}")
