#!/usr/bin/env python
"""bench.py — the headline benchmark of the render hot path.

Workload (BASELINE.json configs[0], the README render): book-1 random-spheres scene (src/main.rs:245-289, seeded),
1200 x 800, 500 spp, max depth 50  ->  480 M paths per step.  One "step" = one whole render.

  python bench.py [--gpus N] [--steps K] [--warmup W]            our arm (CUDA megakernel through the C ABI)
  python bench.py --impl reference ...                           reference arm: the reference's own CPU algorithm
                                                                 (f64, flat world list, all host threads) — the oracle
                                                                 port, because the Rust toolchain is absent.
N > 1 is launched by torchrun, one rank per GPU.  torch.distributed only carries the IPC handle, the barriers and
the max-over-ranks of the timing; pixels never travel through a collective (see dist.py).

Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

import rtw_pkg  # noqa: E402
from oracle.binding import load_oracle  # noqa: E402  (cpu_baseline / reference arm only: the checker, never the product)

SCENE, W, H, SPP, DEPTH = "random_scene", 1200, 800, 500, 50
README_MPATHS = 480.0 / 4200.0          # README.md:6 — 1 h 10 min on 10 CPU threads (older commit of the scene)
FP32_LANES = 148 * 128                  # B200: 148 SMs x 128 FP32 lanes


def load_json(path, default=None):
    try:
        return json.load(open(path))
    except Exception:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.idx = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._pump, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.05 and len(r) >= 9] or [r for _, r in self.rows if len(r) >= 9]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                if r[col].lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "power_w_max": max(float(r[3]) for r in rows),
                "samples": len(rows), "reasons": sorted(reasons)}


def cpu_baseline(m, threads=0, budget_s=12.0, impl_steps=None):
    """The reference's own algorithm on the host cores: f64, FLAT world list (src/main.rs:25 / hittable.rs:43-55),
    all threads, same scene/camera/depth, reduced spp (throughput is spp-independent)."""
    orc = load_oracle()
    sc, spec = m.scenes.build(orc, SCENE, wrap_bvh=False)
    cores = int(orc.f("hardware_threads")()) if threads <= 0 else threads
    # calibrate on a 1/16-area image, 1 spp
    p = m.make_params(W // 4, H // 4, 1, max_depth=DEPTH, background=spec.background)
    r = sc.render_oracle(spec.camera(orc, W // 4, H // 4), p, threads=cores)
    rate = (W // 4) * (H // 4) / max(r["seconds"], 1e-6)
    spp = max(1, min(SPP, int(budget_s * rate / (W * H))))
    return sc, spec, orc, cores, spp


def run_reference(args, m):
    rank, _, world = m.dist.env_rank() if hasattr(m, "dist") else (0, 0, 1)
    if rank != 0:
        return 0
    sc, spec, orc, cores, spp = cpu_baseline(m, budget_s=6.0)
    cam = spec.camera(orc, W, H)
    p = m.make_params(W, H, spp, max_depth=DEPTH, background=spec.background)
    times = []
    for i in range(args.warmup + args.steps):
        r = sc.render_oracle(cam, p, threads=cores)
        if i >= args.warmup:
            times.append(r["seconds"])
    total = sum(times)
    val = W * H * spp * len(times) / total / 1e6
    sample = f"{W}x{H}, {spp} of {SPP} spp per step ({W * H * spp / 1e6:.2f} M paths), depth {DEPTH}, flat world list, f64"
    line = {"impl": "reference", "metric": "Mpaths/s, 1200x800 500spp depth 50 (book-1 random spheres)", "value": val, "unit": "Mpaths/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": val / README_MPATHS, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"book-1 random_scene (seed 1, 485 spheres) {W}x{H} {SPP}spp depth {DEPTH}", "paths_per_step": W * H * spp,
                       "paths_full_workload": W * H * SPP, "sample": sample + " - each step is a bounded sample of the workload, throughput is spp-independent"},
            "cpu_baseline": {"value": val, "unit": "Mpaths/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "Mpaths/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0, "extrapolated_full_render_s": W * H * SPP / (val * 1e6)}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="rtw", choices=["rtw", "reference"])
    ap.add_argument("--spp", type=int, default=SPP, help="debug only; the reported config is 500")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the other BASELINE.json configs (C2-C5) after the headline")
    ap.add_argument("--sweep", default="1,4,16", help="C5 sphere counts in Mi (comma separated; empty = skip)")
    args = ap.parse_args()
    m = rtw_pkg.load()
    from rtw_b200 import dist
    m.dist = dist
    if args.impl == "reference":
        return run_reference(args, m)

    import torch
    rank, local_rank, world = dist.env_rank()
    if world != max(1, args.gpus) and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: librtw has no CPU fallback")
    comm = dist.Comm()
    rtw = m.load_rtw()
    spp = args.spp

    sc, spec = m.scenes.build(rtw, SCENE)
    sc.commit(1, local_rank)
    cam = spec.camera(rtw, W, H)
    prm = m.make_params(W, H, spp, max_depth=DEPTH, background=spec.background, seed=1, n_gpus=world)
    n_paths = W * H * spp
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=comm.device)      # > 126 MB L2

    # ---------------- resident arm: framebuffer stays in HBM ------------------------------------------
    if world == 1:
        fb = torch.zeros(H * W * 3, dtype=torch.float32, device=comm.device)
        prm_dev = m.make_params(W, H, spp, max_depth=DEPTH, background=spec.background, seed=1, flags=m.api.RTW_FLAG_DEVICE_OUT)
        step = lambda: sc.render_device(cam, prm_dev, fb.data_ptr())          # noqa: E731
    else:
        shared = dist.SharedRender(comm, sc, W, H)
        step = lambda: shared.step(cam, prm)                                   # noqa: E731

    warmup_run = max(3, args.warmup)                      # never fewer than 3 untimed steps, whatever was asked
    for _ in range(warmup_run):
        st = step()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    wall, dev_ms, rays = 0.0, 0.0, 0
    t_begin = time.time()
    for _ in range(args.steps):
        flush.fill_(1)                                   # L2 flush between timed iterations
        comm.barrier()                                   # barrier + cuda synchronize
        t0 = time.perf_counter()
        st = step()                                      # blocking: returns after the kernel's stream is synchronised
        if world == 1:
            comm.barrier()                               # (N > 1: the step itself ends with the barrier + cuda synchronize)
        wall += time.perf_counter() - t0
        dev_ms += st["ms_render"]
        rays += st["rays"]
    t_end = time.time()
    clocks = sampler.stop(t_begin, t_end) if rank == 0 else None
    wall = comm.reduce_max(wall)
    dev_ms = comm.reduce_max(dev_ms)
    rays = comm.reduce_sum(rays)
    value = n_paths * args.steps / wall / 1e6
    kernel_ms = dev_ms / args.steps

    # ---------------- e2e arm: host buffers in, host framebuffer out, every step ----------------------
    e2e_wall, h2d, d2h = 0.0, 0, 0
    host_img = None
    host_out = rtw.pinned_image(H, W)                   # caller-allocated, page-locked host framebuffer, reused every step
    blob_bytes = int(sc_blob_bytes(sc))
    for i in range(2 + args.steps):
        comm.barrier()
        t0 = time.perf_counter()
        sc.commit(1, local_rank)                         # flatten + BVH + H2D upload of the scene blob
        if world == 1:
            host_img, st2 = sc.render(cam, prm, out=host_out)   # render + D2H of the H x W x 3 f32 sums
            h2d_i, d2h_i = blob_bytes + st2["h2d_bytes"], st2["d2h_bytes"]
        else:
            st2 = shared.step(cam, prm)
            if rank == 0:
                host_img = shared.read(host_out)
            h2d_i, d2h_i = blob_bytes, (W * H * 12 if rank == 0 else 0)
        comm.barrier()
        if i >= 2:
            e2e_wall += time.perf_counter() - t0
            h2d, d2h = h2d_i, d2h_i
    t0 = time.perf_counter()
    for _ in range(5):
        sc.commit(1, local_rank)
    commit_ms = 1e3 * (time.perf_counter() - t0) / 5          # flatten + SAH BVH + upload, reported separately
    e2e_wall = comm.reduce_max(e2e_wall)
    h2d = comm.reduce_sum(h2d)
    d2h = comm.reduce_sum(d2h)
    e2e_value = n_paths * args.steps / e2e_wall / 1e6
    # ---------------- A27 check (src/main.rs:542-547): the image all ranks add up over NVLink == one GPU's image ----------
    mgpu = None
    if world > 1:
        lo_spp = 32
        prm_lo = m.make_params(W, H, lo_spp, max_depth=DEPTH, background=spec.background, seed=7, n_gpus=world)
        shared.step(cam, prm_lo)
        if rank == 0:
            img_all = shared.read().copy()
            img_one, _ = sc.render(cam, m.make_params(W, H, lo_spp, max_depth=DEPTH, background=spec.background, seed=7))
            diff = float(np.abs(img_all - img_one).max())
            tol = 2e-4 * float(np.abs(img_one).max())
            mgpu = {"max_abs_diff": diff, "tol": tol, "ok": bool(diff <= tol and np.isfinite(img_all).all()), "spp": lo_spp,
                    "what": f"{W}x{H} radiance sums: {world} ranks into rank 0's framebuffer (peer atomics) vs rank 0 alone; f32 summation order is the only difference"}
        comm.barrier()
        shared.close()

    # ---------------- the other BASELINE.json configs at this N (device-timed, not the headline) -----------------------
    configs = None
    if not args.no_configs:
        configs = run_configs(m, rtw, comm, dist, local_rank, world, args)

    if rank == 0:
        peaks = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {})
        model = load_json(os.path.join(ROOT, "profiles", "flop_model.json"), {"configs": {}})
        fpp = model["configs"].get(SCENE, {}).get("flops_per_path", 5960.0)
        sm_max = float(peaks.get("sm_max_mhz", 1965.0))
        derived_peak = FP32_LANES * 2 * sm_max * 1e6 / 1e12
        fp32 = measure_fp32_peak()                                             # tools/fp32_peak: FFMA chains, this GPU, now
        peak_tflops = fp32["fp32_tflops_measured"] if fp32 else derived_peak
        peak_source = (f"measured in this run by tools/fp32_peak ({fp32['how']}); derived 148 SM x 128 lanes x 2 x {sm_max:.0f} MHz = {derived_peak:.2f}"
                       if fp32 else f"DERIVED 148 SM x 128 lanes x 2 x sm_max_mhz {sm_max:.0f} (MEASURED_PEAKS.json clock; tools/fp32_peak not built)")
        achieved = fpp * n_paths / world / (kernel_ms * 1e-3) / 1e12          # per GPU, dominant kernel
        # the same cost table applied to what THIS kernel executes (its own BVH, tile lists): tools/flop_model_device.py
        fpp_dev = model.get("device_counts", {}).get(SCENE, {}).get("flops_per_path")
        ncu = load_json(os.path.join(ROOT, "profiles", "ncu_summary.json"), {})
        line = {
            "metric": "Mpaths/s, 1200x800 500spp depth 50 (book-1 random spheres)", "value": value, "unit": "Mpaths/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "warmup_steps_run": warmup_run, "ms_per_step": 1e3 * wall / args.steps,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": value / README_MPATHS, "dtype": "f32 (+f64 sphere discriminant)",
            "data": "synthetic",
            "config": {"workload": f"book-1 random_scene (seed 1, 485 spheres) {W}x{H} {spp}spp depth {DEPTH}", "paths_per_step": n_paths,
                       "l2": "256 MiB buffer written between timed iterations", "parallelism": f"tiles x sample-chunks from one atomic counter over {world} GPU(s)",
                       "vs_baseline_ref": "README.md:6 (10 CPU threads, older commit of the scene)"},
            "wall_time_s": wall / args.steps, "kernel_ms_per_step": kernel_ms, "commit_ms": commit_ms, "rays_per_path": rays / (n_paths * args.steps),
            "mrays_per_s": rays / wall / 1e6,
            "e2e": {"value": e2e_value, "unit": "Mpaths/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": 1e3 * e2e_wall / args.steps, "includes": "rtw_scene_commit (flatten + BVH build + H2D upload of the scene) + rtw_render (kernel + D2H of the H x W x 3 f32 sums into a pinned host buffer), every step",
                    "excludes": "building the scene GRAPH (the ~1 400 constructor calls through the C ABI, done once before the loop: host-only bookkeeping, no device work)"},
            "gpu_launches": args.steps * world,
            "clocks": clocks,
            "roofline": {"bound": "fp32", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s", "frac": achieved / peak_tflops,
                         "traffic": ncu.get("render_kernel", {}).get("dram_bytes_per_launch"),
                         "kernel": "render_kernel", "flops_per_path": fpp,
                         "flops_per_path_is": "REFERENCE-ALGORITHM-EQUIVALENT work (SURVEY 8d): oracle event counts with the world in a reference-style median-split BVH x the reference's f64 op costs (profiles/flop_model.json) - not the instructions this kernel executes",
                         "frac_device_counts": (fpp_dev * n_paths / world / (kernel_ms * 1e-3) / 1e12 / peak_tflops) if fpp_dev else None,
                         "flops_per_path_device_counts": fpp_dev,
                         "peak_source": peak_source + "; FP32 pipe / issue is the bound, no tensor- or HBM-bound work on this path",
                         "peak_derived_tflops": derived_peak,
                         "hbm_view": {"algorithmic_bytes_per_launch": W * H * 12 + int(sc_blob_bytes(sc)), "peak_gbs": peaks.get("hbm_gbs")}},
        }
        if not args.no_cpu_baseline:
            scb, specb, orc, cores, bspp = cpu_baseline(m)
            r = scb.render_oracle(specb.camera(orc, W, H), m.make_params(W, H, bspp, max_depth=DEPTH, background=specb.background), threads=cores)
            v = W * H * bspp / r["seconds"] / 1e6
            line["cpu_baseline"] = {"value": v, "unit": "Mpaths/s", "cores": cores, "kind": "port",
                                    "sample": f"{W}x{H}, {bspp} of {SPP} spp ({W * H * bspp / 1e6:.2f} M paths, {r['seconds']:.1f} s), depth {DEPTH}, flat world list, f64",
                                    "extrapolated_full_render_s": n_paths / (v * 1e6)}
        if host_img is not None:
            line["image_mean"] = float(host_img.mean() / spp)
        if mgpu is not None:
            line["multi_gpu_check"] = mgpu
        if configs is not None:
            line["configs"] = configs
        print(json.dumps(line))
    ok = comm.reduce_max(0.0 if (mgpu is None or mgpu["ok"]) else 1.0) == 0.0
    comm.close()
    if not ok:
        sys.stderr.write("multi_gpu_check FAILED: the multi-GPU image differs from the single-GPU image\n")
        return 3
    return 0


def run_configs(m, rtw, comm, dist, local_rank, world, args):
    """BASELINE.json configs[1..4] at the run's N: C2a/b/c 800x450x200, C3a/b/c 600x600x1000, C4 final_scene 800x800 (timed at
    1000 of its 10 000 spp: Mpaths/s is spp-independent), C5 the 1 M / 4 M / 16 M sphere sweep at 3840x2160x256.
    Device-timed (CUDA events around the kernel, max over ranks); multi-rank renders go through the same shared-framebuffer
    path as the headline.  Returns {name: {...}} on rank 0 (other ranks: partial)."""
    out = {}
    small = [("C2a two_spheres", "two_spheres", 800, 450, 200), ("C2b two_perlin_spheres", "two_perlin_spheres", 800, 450, 200),
             ("C2c earth", "earth", 800, 450, 200), ("C3a simple_light", "simple_light", 600, 600, 1000),
             ("C3b cornell_box", "cornell_box", 600, 600, 1000), ("C3c cornell_box_smoke", "cornell_box_smoke", 600, 600, 1000),
             ("C4 final_scene (1000 of 10000 spp)", "final_scene", 800, 800, 1000)]
    sweep = [int(float(x) * (1 << 20)) for x in args.sweep.split(",") if x]

    def timed(sc, spec, Wc, Hc, sppc, reps, warm):
        cam = spec.camera(rtw, Wc, Hc)
        prm = m.make_params(Wc, Hc, sppc, max_depth=DEPTH, background=spec.background, seed=1, n_gpus=world)
        sh = dist.SharedRender(comm, sc, Wc, Hc) if world > 1 else None
        best, rays, mean = [], 0, None
        for i in range(warm + reps):
            if i < warm and sppc > 64:       # a short warm-up pass (same kernel, same scene; fills caches and instruction memory)
                p_i = m.make_params(Wc, Hc, 16, max_depth=DEPTH, background=spec.background, seed=1, n_gpus=world)
            else:
                p_i = prm
            if sh is None:
                img, st = sc.render(cam, p_i)
            else:
                st = sh.step(cam, p_i)
            if i >= warm:
                best.append(comm.reduce_max(st["ms_render"]))
                rays = comm.reduce_sum(st["rays"])
        if sh is not None:
            if comm.rank == 0:
                mean = float(sh.read().mean() / sppc)
            sh.close()
        else:
            mean = float(img.mean() / sppc)
        ms = sum(best) / len(best)
        n = Wc * Hc * sppc
        return {"mpaths_per_s": n / ms / 1e3, "ms": ms, "rays_per_path": rays / n, "paths": n, "image_mean": mean}

    for label, name, Wc, Hc, sppc in small:
        sc, spec = m.scenes.build(rtw, name)
        sc.commit(1, local_rank)
        out[label] = dict(timed(sc, spec, Wc, Hc, sppc, reps=3, warm=2), workload=f"{name} {Wc}x{Hc} {sppc}spp depth {DEPTH}")
        sc.close()
    for n_sph in sweep:
        sc = m.Scene(rtw)
        t0 = time.perf_counter()
        spec = m.scenes.sweep_scene(sc, n_sph)
        t1 = time.perf_counter()
        sc.commit(1, local_rank)
        commit_s = comm.reduce_max(time.perf_counter() - t1)
        r = timed(sc, spec, 3840, 2160, 256, reps=1, warm=1)
        st = sc.debug_stats() if hasattr(sc, "debug_stats") else {}
        out[f"C5 sweep {n_sph >> 20}M spheres"] = dict(r, workload=f"{n_sph} random spheres 3840x2160 256spp depth {DEPTH}", commit_s=commit_s,
                                                       scene_build_host_s=t1 - t0, **st)
        sc.close()
    return out


def measure_fp32_peak():
    exe = os.path.join(ROOT, "tools", "fp32_peak")
    if not os.path.exists(exe):
        return None
    try:
        r = subprocess.run([exe], capture_output=True, text=True, timeout=60)
        d = json.loads(r.stdout.strip().splitlines()[-1])
        return d if r.returncode == 0 and d.get("fp32_tflops_measured", 0) > 1 else None
    except Exception:
        return None


def sc_blob_bytes(sc):
    d = sc.debug_flatten()
    return d["nodes"] * 64 + d["prims"] * 64 + d["xforms"] * 160 + d["media"] * 16 + d["mats"] * 32 + d["texs"] * 48


if __name__ == "__main__":
    sys.exit(main())
