"""Fixture from the reference's OWN output: /root/reference/generated_images/earth.ppm (400x225, P3, written by an older
commit of the reference: same globe, same camera, gradient sky instead of HEAD's constant background).  It is the one
artifact in the repository that shows what the reference's sphere_uv (src/math.rs:288-300) + image texture addressing
(src/texture.rs:46-73) + camera framing put on screen.  We keep 5x5 block means (45 x 80 x 3, float32, 43 KB), not the
file itself.  Run in the build container only (the GPU box has no /root/reference):  python tools/make_golden_earth.py"""
import os, sys
import numpy as np
src = "/root/reference/generated_images/earth.ppm"
tok = open(src).read().split()
assert tok[0] == "P3" and (int(tok[1]), int(tok[2]), int(tok[3])) == (400, 225, 255)
img = np.array(tok[4:], dtype=np.uint8).reshape(225, 400, 3)
blocks = img.astype(np.float32).reshape(45, 5, 80, 5, 3).mean((1, 3))
out = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ref_earth_400x225_blocks5.npy")
np.save(out, blocks)
print("wrote", out, blocks.shape, "mean", blocks.mean((0, 1)))
