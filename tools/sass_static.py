"""Static SASS instruction count of one kernel by source function / line (needs -lineinfo).
Usage: sass_static.py <lib.so> <mangled kernel substring> [lines]"""
import collections, os, re, subprocess, sys, tempfile
lib, key = sys.argv[1], sys.argv[2]
show_lines = len(sys.argv) > 3
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, capture_output=True)
cubin = max((f for f in os.listdir(tmp) if f.endswith(".cubin")), key=lambda f: os.path.getsize(os.path.join(tmp, f)))   # the render kernels' module
sass = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
inside = False; cur = None; cnt = collections.Counter()
for l in sass:
    if l.startswith(".text."):
        inside = key in l; continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (os.path.basename(m.group(1)), int(m.group(2))); continue
    if re.match(r'\s+/\*[0-9a-f]{4,}\*/\s+\S', l): cnt[cur] += 1
srcs = {}
def fn(f, line):
    if f not in srcs:
        path = os.path.join(root, "rust-ray-tracing-in-a-weekend_b200", "csrc", f)
        st = []
        if os.path.exists(path):
            for i, l in enumerate(open(path), 1):
                m = re.match(r'(?:template\s*<[^>]*>\s*)?(?:RTW_DEV|__global__|static|inline)\s+[\w:<>\*&\s]+?\s+(\w+)\s*\(', l)
                if m: st.append((i, m.group(1)))
                m = re.match(r'\s+RTW_DEV\s+[\w:<>\*&\s]+?\s+(\w+)\s*\(', l)
                if m: st.append((i, m.group(1)))
        srcs[f] = st
    name = f
    for s, n in srcs[f]:
        if s <= line: name = f + ":" + n
        else: break
    return name
by = collections.Counter()
for (f, ln), c in cnt.items(): by[fn(f, ln)] += c
print("total", sum(cnt.values()))
for k, v in by.most_common(45): print(f"{v:5d} {k}")
if show_lines:
    print("--- top lines")
    for (f, ln), c in cnt.most_common(40): print(f"{c:5d} {f}:{ln}")
