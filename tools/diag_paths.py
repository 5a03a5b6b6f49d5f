import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, rtw_pkg
m = rtw_pkg.load(); rtw, orc = m.load_rtw(), m.api.load_oracle()
name = sys.argv[1] if len(sys.argv) > 1 else "cornell_box"
a, spec = m.scenes.build(rtw, name); a.commit(1, 0)
b, _ = m.scenes.build(orc, name, wrap_bvh=name not in ("final_scene", "cornell_box_smoke")); b.set_media_deferred(True)
W, H = 96, 64
rs = np.random.RandomState(12); n = 60000
px, py, sm = rs.randint(0, W, n), rs.randint(0, H, n), rs.randint(0, 4096, n)
ca, cb = spec.camera(rtw, W, H), spec.camera(orc, W, H)
for depth in (1, 2, 3, 4, 6, 10, 50):
    p = m.make_params(W, H, 64, max_depth=depth, background=spec.background, seed=7)
    ra, sa = a.trace_paths(ca, p, px, py, sm); rb, sb = b.trace_paths(cb, p, px, py, sm)
    good = (sa == sb) & (np.abs(ra - rb).max(1) <= 1e-3 * np.maximum(1, np.abs(rb).max(1)))
    print(f"depth {depth:2d}: match {good.mean():.4f} seg gpu {sa.mean():.4f} orc {sb.mean():.4f} rad gpu {ra.mean():.5f} orc {rb.mean():.5f}  seg-only-mismatch {np.mean(sa != sb):.4f}")
    if depth == 2:
        bad = np.where(sa != sb)[0][:12]
        for i in bad: print("   ", px[i], py[i], sm[i], "seg", sa[i], sb[i], "rad", ra[i], rb[i])
# first-hit comparison for camera rays of mismatching paths at depth 2
p = m.make_params(W, H, 64, max_depth=2, background=spec.background, seed=7)
ra, sa = a.trace_paths(ca, p, px, py, sm); rb, sb = b.trace_paths(cb, p, px, py, sm)
bad = np.where(sa != sb)[0]
print("n bad at depth 2:", len(bad))
d = ra - rb
print("depth2: max abs diff", np.abs(d).max(), "n nonzero>1e-6", np.sum(np.abs(d).max(1) > 1e-6), "sum diff", d.sum(0), "n light paths", np.sum(rb.max(1) > 1))
idx = np.argsort(-np.abs(d).max(1))[:10]
for i in idx: print("   ", px[i], py[i], sm[i], "seg", sa[i], sb[i], "gpu", ra[i], "orc", rb[i])
vals, cnt = np.unique(np.round(rb[:, 0], 4), return_counts=True)
print("orc distinct values", list(zip(vals[:12], cnt[:12])))
vals, cnt = np.unique(np.round(ra[:, 0], 4), return_counts=True)
print("gpu distinct values", list(zip(vals[:12], cnt[:12])))
