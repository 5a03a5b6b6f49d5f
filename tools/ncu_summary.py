"""Turn a .ncu-rep (ncu --set full) into the summaries committed under profiles/:
  <tag>_ncu.json   selected raw metrics        <tag>_lines.txt  per-CUDA-source-line instruction / lane-occupancy table
and refresh profiles/ncu_summary.json (read by bench.py for roofline.traffic).
Usage: ncu_summary.py gpurun_out/prof.ncu-rep <tag> <kernel-key> "<what>" """
import csv, json, os, subprocess, sys
rep, tag, key, what = sys.argv[1:5]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, vals = rows[0], rows[1], rows[2]
keep = ['gpu__time_duration.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'sm__cycles_elapsed.avg.per_second', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__sass_thread_inst_executed_op_fadd_pred_on.sum', 'sm__sass_thread_inst_executed_op_fmul_pred_on.sum',
        'sm__sass_thread_inst_executed_op_ffma_pred_on.sum', 'sm__sass_thread_inst_executed_op_dfma_pred_on.sum', 'sm__inst_executed.sum', 'smsp__thread_inst_executed.sum']
keep += [h for h in hdr if 'issue_stalled' in h and h.endswith('.ratio')]
d = {h: {"value": vals[i], "unit": units[i]} for i, h in enumerate(hdr) if h in keep}
json.dump({"what": what, "report": os.path.basename(rep), "metrics": d}, open(os.path.join(root, "profiles", tag + "_ncu.json"), "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
tmp = "/tmp/_ncu_src.csv"; open(tmp, "w").write(src)
lines = subprocess.run([sys.executable, os.path.join(root, "tools", "ncu_lines.py"), tmp, "45"], capture_output=True, text=True).stdout
open(os.path.join(root, "profiles", tag + "_lines.txt"), "w").write(what + "\n" + lines)
def num(k):
    v = d.get(k, {}).get("value", "0").replace(",", ""); u = d.get(k, {}).get("unit", "")
    f = float(v) if v else 0.0
    return f * {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "byte": 1}.get(u, 1)
summ_path = os.path.join(root, "profiles", "ncu_summary.json")
summ = json.load(open(summ_path)) if os.path.exists(summ_path) else {}
summ[key] = {"dram_bytes_per_launch": num('dram__bytes_read.sum') + num('dram__bytes_write.sum'), "from": tag + "_ncu.json", "what": what,
             "lanes_per_warp_inst": float(d['smsp__thread_inst_executed_per_inst_executed.ratio']['value']),
             "issue_active_pct": float(d['smsp__issue_active.avg.pct_of_peak_sustained_active']['value'])}
json.dump(summ, open(summ_path, "w"), indent=1)
print(json.dumps(summ[key])); print(lines[:600])
