"""Path-level parity diagnosis for one scene: 200 k individual (pixel, sample) paths traced by the CUDA path
(rtw_trace_paths) and by the oracle with the same Philox keys; prints the share of identical paths and the segment-count
histogram of the ones that diverge.  Usage: diag_paths2.py [scene]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw, orc = m.load_rtw(), __import__("oracle.binding", fromlist=["x"]).load_oracle()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell_box"
a, spec = m.scenes.build(rtw, name); a.commit(1, 0)
b, _ = m.scenes.build(orc, name, wrap_bvh=name not in ("final_scene", "cornell_box_smoke")); b.set_media_deferred(True)
W, H = 96, 64
rs = np.random.RandomState(12); n = 200000
px, py, sm = rs.randint(0, W, n), rs.randint(0, H, n), rs.randint(0, 4096, n)
ca, cb = spec.camera(rtw, W, H), spec.camera(orc, W, H)
p = m.make_params(W, H, 64, max_depth=50, background=spec.background, seed=7)
ra, sa = a.trace_paths(ca, p, px, py, sm); rb, sb = b.trace_paths(cb, p, px, py, sm)
print("seg hist gpu", np.bincount(sa, minlength=51)[[1,2,3,5,10,20,30,40,49,50]], "orc", np.bincount(sb, minlength=51)[[1,2,3,5,10,20,30,40,49,50]])
print("n seg>=40 gpu", np.sum(sa >= 40), "orc", np.sum(sb >= 40), " mean", sa.mean(), sb.mean())
bad = np.where((sa == 50) & (sb < 50))[0][:8]
print("gpu-50 paths:", [(int(px[i]), int(py[i]), int(sm[i]), int(sb[i])) for i in bad])
# trace one of them by depth to find where it diverges, and dump the hit sequence through test_hit
for i in bad[:3]:
    prev = None
    for depth in range(1, 12):
        pp = m.make_params(W, H, 64, max_depth=depth, background=spec.background, seed=7)
        r1, s1 = a.trace_paths(ca, pp, [px[i]], [py[i]], [sm[i]]); r2, s2 = b.trace_paths(cb, pp, [px[i]], [py[i]], [sm[i]])
        print(f"   path {i} depth {depth}: gpu seg {s1[0]} rad {r1[0][0]:.4f} | orc seg {s2[0]} rad {r2[0][0]:.4f}")
        if s1[0] != s2[0] and prev: break
        prev = True
