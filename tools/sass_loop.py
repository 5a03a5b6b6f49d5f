"""Print the BVH node loop of one kernel variant from the SASS (first backward branch after the slab FFMA ..., 1.0000004).
Usage: sass_loop.py <lib.so> <mangled kernel substring>"""
import re, subprocess, sys
out = subprocess.run(["cuobjdump", "-sass", sys.argv[1]], capture_output=True, text=True).stdout.split("\n")
on = False; ins = []
for l in out:
    if "Function :" in l: on = sys.argv[2] in l
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if on and m: ins.append((int(m.group(1), 16), m.group(2).strip()))
k = next(i for i, (a, t) in enumerate(ins) if re.search(r"FFMA .*1\.00000[0-9]+", t))
for j in range(k, len(ins)):
    m = re.search(r"BRA\s+(?:P\d, )?0x([0-9a-f]+)", ins[j][1])
    if m and int(m.group(1), 16) < ins[k][0]:
        tgt = int(m.group(1), 16); break
i0 = next(i for i, (a, t) in enumerate(ins) if a == tgt)
print("loop", hex(tgt), "..", hex(ins[j][0]), "=", j - i0 + 1, "instructions; kernel total", len(ins))
for a, t in ins[i0:j + 1]: print(f"  {a:05x}  {t}")
