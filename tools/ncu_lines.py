"""Aggregate an `ncu --page source --csv --print-source cuda,sass` export by CUDA source line:
instructions executed, avg active threads, share of samples.  Usage: ncu_lines.py file.csv [top]"""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur_file = None
agg = collections.OrderedDict()
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    if r[0] != "":   # a CUDA source line summary row
        ix = {h: i for i, h in enumerate(hdr)}
        key = (cur_file, int(r[0]))
        try:
            inst = int(r[7]); thr = int(r[8]); samp = int(r[6])
        except ValueError:
            continue
        a = agg.setdefault(key, [0, 0, 0, r[1].strip()[:90]])
        a[0] += inst; a[1] += thr; a[2] += samp
tot_i = sum(a[0] for a in agg.values()); tot_t = sum(a[1] for a in agg.values()); tot_s = sum(a[2] for a in agg.values())
print(f"total warp-inst {tot_i:.3e} thread-inst {tot_t:.3e} avg threads {tot_t / max(tot_i,1):.2f} samples {tot_s}")
for (f, ln), a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
    print(f"{f}:{ln:4d} inst {100 * a[0] / tot_i:5.2f}% thr/inst {a[1] / max(a[0],1):5.1f} samp {100 * a[2] / max(tot_s,1):5.2f}%  {a[3]}")
