#!/usr/bin/env python
"""bench.py — headline benchmark of the zraytrace hot path on B200.

Metric (BASELINE.json): Mrays/s, device-timed, whole job over N GPUs; a "step" is one full render of the
workload (default c5 = the README headline: 7-spheres, 1000x1000, 1000 spp, depth 30, spp split across
ranks, one NCCL reduce of the fp32 accumulators).  rays = the reference's `rays_processed` counter
(raytrace.zig:69).  See DESIGN.md "Measurement" for every field of the JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl zrt|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...     (one rank per GPU)
Without torchrun, --gpus N > 1 drives the N devices from this ONE process (zrt_multi_create: ncclCommInitAll).
Either way the per-step path is libzrt only: zrt_multi_render = trace + ncclReduce + 1/spp (+ copy home for e2e).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

from zraytrace_b200 import _abi as A  # noqa: E402  (ctypes structs only: does not load libzrt.so)

# workload -> (scene index, variant, width, height, spp, depth, aspect, x_limit, bvh flags, description)
WORKLOADS = {
    "c1": dict(scene=1, variant=0, w=200, h=200, spp=100, depth=30, desc="7-spheres 200x200 100spp depth30 list"),
    "c2": dict(scene=3, variant=0, w=512, h=512, spp=256, depth=30, desc="teapot+ground 512x512 256spp depth30 BVH"),
    "c3": dict(scene=2, variant=1, w=1024, h=1024, spp=512, depth=30, desc="bunny(dielectric 1.52)+ground 1024x1024 512spp depth30 BVH"),
    "c4": dict(scene=5, variant=2, w=1920, h=1080, spp=1024, depth=30, aspect=16 / 9, x_limit=A.ZRT_XLIMIT_WIDTH,
               desc="goat-substitute(bunny x64)+Man, image textures, 1920x1080 1024spp depth30 BVH"),
    "c5": dict(scene=1, variant=0, w=1000, h=1000, spp=1000, depth=30, desc="7-spheres 1000x1000 1000spp depth30 list (README headline)"),
}

# Algorithmic FP32 operation model of the reference path (1 op per add/sub/mul/div/sqrt/
# min/max/compare-select as written in the reference, transcendentals not counted, no FMA credit).
OPS = dict(sample=30 + 3, sphere_test=18, sphere_sqrt=5, sphere_accept=28, triangle_test=45, triangle_accept=15,
           box_test=24, lambertian=19 + 3, metal=36 + 3, dielectric_reflect=52 + 3, dielectric_refract=76 + 3,
           background=21)


def algorithmic_ops(stats, counters):
    s = stats
    return (OPS["sample"] * counters["samples_processed"] + OPS["sphere_test"] * s["sphere_tests"]
            + OPS["sphere_sqrt"] * s["sphere_sqrt"] + OPS["sphere_accept"] * s["sphere_accepts"]
            + OPS["triangle_test"] * s["triangle_tests"] + OPS["triangle_accept"] * s["triangle_accepts"]
            + OPS["box_test"] * s["box_tests"] + OPS["lambertian"] * s["lambertian"]
            + OPS["metal"] * (s["metal"] + s["metal_absorbed"]) + OPS["dielectric_reflect"] * s["dielectric_reflect"]
            + OPS["dielectric_refract"] * s["dielectric_refract"] + OPS["background"] * s["background"])


# Per-launch numbers of the dominant kernel from the committed `ncu --set full` captures named here (a number measured
# under a profiler is only ever used for these fields): dram bytes read + written, l1tex / lts bytes, issue-slot
# utilisation and active lanes per instruction.  Filled from profiles/ by hand after each capture.
NCU = {
    "c5": {"traffic": 71932672 + 392275456, "issue_active_pct": 83.7, "active_lanes": 27.25,
           "source": "profiles/r2_w_c5_pool3_final_summary.txt (1 GPU, k_trace_pool3<7,128,7>, the full 1000 spp launch; 384 MB of the "
                     "writes are the 32 partial-sum slices)"},
    "c2": {"traffic": 6677248 + 61849856, "issue_active_pct": 78.9, "active_lanes": 17.61,
           "source": "profiles/r2_w_c2_ws_final_summary.txt (k_trace_ws, the full 256 spp launch)"},
    "c3": {"traffic": 21672960 + 380491776, "issue_active_pct": 82.5, "active_lanes": 20.41,
           "source": "profiles/r2_w_c3_ws_final_summary.txt (k_trace_ws, 128 spp launch: the 32 slices are the same 403 MB at any sample count)"},
    "c4": {"traffic": 124312064 + 796394240, "issue_active_pct": 79.1, "active_lanes": 20.33,
           "source": "profiles/r2_w_c4_ws_final_summary.txt (k_trace_ws, 128 spp launch: the 32 slices are the same 796 MB at any sample count)"},
}


class ClockSampler:
    """SM clock, power and clock-event (throttle) reasons of the listed GPUs, sampled through NVML (what nvidia-smi
    reads) every few ms from a thread that is started at least a second before the timed region; only the samples
    whose timestamps fall inside the region count.  Falls back to `nvidia-smi -lms` if pynvml is missing."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_indices, period_s=0.005):
        self.idx = list(gpu_indices)
        self.period = period_s
        self.samples = []  # (t, gpu, sm_mhz, max_mhz, power_w, reasons_mask)
        self._stop = threading.Event()
        self._thread = None
        self._smi = None
        self.backend = None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            hs = [(i, pynvml.nvmlDeviceGetHandleByIndex(i)) for i in self.idx]
            mx = {i: pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM) for i, h in hs}
            reasons = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def loop():
                while not self._stop.is_set():
                    for i, h in hs:
                        try:
                            self.samples.append((time.perf_counter(), i, pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                                 mx[i], pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0, int(reasons(h))))
                        except Exception:
                            pass
                    time.sleep(self.period)

            self._thread = threading.Thread(target=loop, daemon=True)
            self._thread.start()
            self.backend = "nvml"
        except Exception:
            self._start_smi()

    def _start_smi(self):
        q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        self._f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self._smi = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=self._f, stderr=subprocess.DEVNULL)
            self.backend = "nvidia-smi"
        except OSError:
            self._smi = None

    def stop(self, t0=None, t1=None):
        if self._thread is not None:
            self._stop.set()
            self._thread.join()
            rows = [s for s in self.samples if t0 is None or t0 <= s[0] <= t1]
            sm = [s[2] for s in rows]
            masks = [s[5] for s in rows]
            reasons = sorted({n for n, bit in self.REASONS for mk in masks if mk & bit})
            return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max((s[3] for s in rows), default=None),
                    "reasons": reasons, "samples": len(rows), "samples_total": len(self.samples),
                    "power_w_max": max((s[4] for s in rows), default=None), "backend": "nvml", "period_ms": 1e3 * self.period}
        if self._smi is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no NVML, no nvidia-smi"], "samples": 0}
        time.sleep(0.1)
        self._smi.terminate()
        self._smi.wait()
        self._f.flush()
        self._f.seek(0)
        sm, mx, power, reasons = [], [], [], set()
        for line in self._f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 8 or not c[0].isdigit() or int(c[0]) not in self.idx:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self._f.name)
        busy = [s for s, p in zip(sm, power) if p > 250.0] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(busy), "samples_total": len(sm),
                "power_w_max": max(power) if power else None, "backend": "nvidia-smi -lms 50 (samples under load, no timestamps)"}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def params_for(wl, **kw):
    return A.make_params(wl["w"], wl["h"], wl["spp"], wl["depth"], bvh=True, x_limit=wl.get("x_limit", A.ZRT_XLIMIT_HEIGHT),
                         seed=42, **kw)


def oracle_scene(wl_name, wl):
    """The workload's scene for the CPU legs, built WITHOUT libzrt where a pure-Python builder exists
    (tests/scenes_py.py restates scenes.zig); c4's subdivided mesh only exists in the C++ host mirror."""
    from tests import scenes_py
    make = {"c1": scenes_py.three_balls, "c5": scenes_py.three_balls, "c2": scenes_py.teapot_and_ball,
            "c3": lambda: scenes_py.bunny_and_ball(dielectric=True)}.get(wl_name)
    if make is not None:
        sc, cam = make()
        return sc, cam, "tests/scenes_py (pure Python)"
    from zraytrace_b200 import host
    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
    return hs, hs.camera, "zraytrace_b200.host (C++ host mirror inside libzrt: c4's subdivided mesh has no Python builder)"


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args, wl_name, wl):
    """--impl reference: the reference's CPU path (the oracle port; the Zig original cannot be built in this
    image) on the host cores, same workload plane, bounded spp per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import zro_py

    cores = os.cpu_count() or 1
    sc, cam, scene_src = oracle_scene(wl_name, wl)
    # bounded sample: same image plane and depth, spp reduced so one step is a few seconds on this host
    target_rays = 6e6 * cores * (1.0 if wl_name in ("c1", "c5") else 0.02)
    rays_per_sample = 2.15 if wl_name in ("c1", "c5") else 1.6
    spp = int(max(1, min(wl["spp"], target_rays / (wl["w"] * wl["h"] * rays_per_sample))))
    p = A.make_params(wl["w"], wl["h"], spp, wl["depth"], bvh=True, x_limit=wl.get("x_limit", 0), seed=42)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, c, _ = zro_py.render(sc, cam, p, rng=zro_py.RNG_CTR, traversal=zro_py.TRAVERSAL_REF,
                                math=zro_py.MATH_LIBM, threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays = c.rays_processed
    total = sum(times)
    value = rays * len(times) / total / 1e6
    # the literal single-thread figure (sequential Xoroshiro stream, as the reference really runs), bounded to ~5 s
    heavy = wl_name not in ("c1", "c5")
    div = (16 if wl_name == "c4" else 4) if heavy else 4
    p1 = A.make_params(wl["w"] // div, wl["h"] // div, 1 if heavy else 8, wl["depth"], bvh=True, x_limit=wl.get("x_limit", 0), seed=42)
    t0 = time.perf_counter()
    _, c1, _ = zro_py.render(sc, cam, p1, rng=zro_py.RNG_REF, traversal=zro_py.TRAVERSAL_REF, math=zro_py.MATH_LIBM)
    one_thread = c1.rays_processed / (time.perf_counter() - t0) / 1e6
    sample = (f"oracle port of the reference CPU path (Zig original not buildable here), {wl['w']}x{wl['h']} plane, "
              f"{spp} spp per step instead of {wl['spp']}, depth {wl['depth']}, {cores} threads over scanlines, "
              f"literal aabb.zig traversal, glibc transcendentals; scene from {scene_src}")
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl_name}: {wl['desc']}", "spp_per_step": spp},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample,
                             "single_thread_literal": {"value": one_thread, "unit": "Mrays/s", "cores": 1,
                                                       "sample": f"sequential Xoroshiro128+ stream, {p1.width}x{p1.height} plane, "
                                                                 f"{p1.samples_per_pixel} spp: how the reference itself runs (README.md:11)"}},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------- CPU legs of our arm
def algorithmic_model(wl_name, wl):
    """Event counts of the oracle (interval-carrying traversal of the REFERENCE tree, counter RNG) on a bounded sample of
    the workload -> algorithmic FP32 ops and bytes per ray.  All host cores, about a second."""
    from oracle import zro_py

    sc, cam, _ = oracle_scene(wl_name, wl)
    heavy = wl_name not in ("c1", "c5")
    div = (8 if wl_name == "c4" else 2) if heavy else 1
    spp = 2 if not heavy else 4
    p = A.make_params(wl["w"] // div, wl["h"] // div, spp, wl["depth"], bvh=True, x_limit=wl.get("x_limit", 0), seed=42)
    _, c, st = zro_py.render(sc, cam, p, rng=zro_py.RNG_CTR, traversal=zro_py.TRAVERSAL_TIGHT, threads=os.cpu_count() or 1)
    ops_per_ray = algorithmic_ops(st.as_dict(), c.as_dict()) / c.rays_processed
    bytes_per_ray = (64 * st.box_passes + 48 * st.triangle_tests + 16 * st.sphere_tests
                     + 12 * c.samples_processed + 4 * st.texture_lookups) / c.rays_processed
    d = st.as_dict()
    events = {"ops": algorithmic_ops(d, c.as_dict()), "rays": c.rays_processed, "box_tests": d["box_tests"],
              "triangle_tests": d["triangle_tests"], "sphere_tests": d["sphere_tests"]}
    return ops_per_ray, bytes_per_ray, f"{p.width}x{p.height} plane at {spp} spp, {c.rays_processed} rays", events


def cpu_baseline(wl_name, wl):
    """Oracle timed on ONE host core in its most literal mode (sequential Xoroshiro stream, libm, literal
    traversal): the reference is single-threaded (README.md:11).  Bounded sample, ~10-20 s."""
    from oracle import zro_py

    sc, cam, _ = oracle_scene(wl_name, wl)
    heavy = wl_name not in ("c1", "c5")
    spp = 1 if heavy else max(1, min(wl["spp"], int(60e6 / (wl["w"] * wl["h"] * 2.15))))
    div = 16 if wl_name == "c4" else 4  # literal aabb.zig traversal visits ~10^3 nodes per ray
    w, h = (wl["w"] // div, wl["h"] // div) if heavy else (wl["w"], wl["h"])
    p = A.make_params(w, h, spp, wl["depth"], bvh=True, x_limit=wl.get("x_limit", 0), seed=42)
    t0 = time.perf_counter()
    _, c, st = zro_py.render(sc, cam, p, rng=zro_py.RNG_REF, traversal=zro_py.TRAVERSAL_REF, math=zro_py.MATH_LIBM)
    dt = time.perf_counter() - t0
    return {"value": c.rays_processed / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": (f"oracle port, single thread, sequential Xoroshiro128+ stream, glibc math, literal aabb.zig traversal; "
                       f"{w}x{h} plane at {spp} spp (full workload is {wl['w']}x{wl['h']} at {wl['spp']} spp), {dt:.1f} s"),
            "rays_per_sample": c.rays_processed / c.samples_processed}


# ------------------------------------------------------------------------------------------- our arm
def roofline_for(wl_name, kernel_name, kernel_ms, rays_launch, ops_per_ray, bytes_per_ray, model_sample, peaks, stats=None, model_events=None):
    """The `roofline` object of one workload: the bound that applies, algorithmic work per launch / kernel time."""
    peaks_file, peaks_src = load_peaks()
    ncu = NCU.get(wl_name, {})
    common = {"kernel": kernel_name, "kernel_ms": kernel_ms, "rays_per_launch": rays_launch,
              "algorithmic_ops_per_ray": ops_per_ray, "algorithmic_bytes_per_ray": bytes_per_ray, "model_sample": model_sample,
              "traffic": ncu.get("traffic"), "traffic_source": ncu.get("source"), "k0": peaks,
              "hbm_peak_gbs": peaks_file.get("hbm_gbs"), "hbm_peak_source": peaks_src,
              "mrays_per_s_kernel_only": rays_launch / (kernel_ms * 1e-3) / 1e6}
    for k in ("issue_active_pct", "active_lanes", "l1tex_bytes", "lts_bytes"):
        if k in ncu:
            common["ncu_" + k] = ncu[k]
    if wl_name in ("c2", "c3", "c4") and stats is not None and stats.rays:
        # BVH workloads.  The scene (<= 36 MB of nodes and triangle planes) lives in L1 / L2 and HBM is idle; ncu shows the kernels
        # bound by instruction issue (issue slots 61-77 % busy at 17-21 active lanes, L1 hit rate 69-88 %), not by L2 bandwidth.
        # The roof is therefore the same FP32-issue roof as on the sphere scenes.  `achieved` counts the algorithmic
        # operations of the traversal the DEVICE runs (its own event counts from the instrumented build: 2 box tests per node
        # visit, triangle / sphere tests) plus the shading events of the oracle model; the traversal terms of the reference
        # tree (oracle, interval-carrying test: what SURVEY 8(d) defines) are reported beside it, as are the byte-side figures.
        m = model_events
        trav_ref = OPS["box_test"] * m["box_tests"] + OPS["triangle_test"] * m["triangle_tests"] + OPS["sphere_test"] * m["sphere_tests"]
        shade_ops_per_ray = (m["ops"] - trav_ref) / m["rays"]
        trav_dev_per_ray = (OPS["box_test"] * 2 * stats.node_visits + OPS["triangle_test"] * stats.triangle_tests
                            + OPS["sphere_test"] * stats.sphere_tests) / stats.rays
        dev_ops_per_ray = shade_ops_per_ray + trav_dev_per_ray
        achieved = dev_ops_per_ray * rays_launch / (kernel_ms * 1e-3) / 1e12
        peak = peaks["fp32_nofma_ops"] / 1e12
        dev_bytes = (64 * stats.node_visits + 48 * stats.triangle_tests + 32 * stats.sphere_tests + 12 * stats.samples
                     + 4 * stats.texture_lookups) / stats.rays
        r = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
             "peak_source": "FP32 issue without FMA credit, measured in this run by zrt_measure_peaks",
             "device_algorithmic_ops_per_ray": dev_ops_per_ray,
             "device_events_per_ray": {"node_visits": stats.node_visits / stats.rays, "triangle_tests": stats.triangle_tests / stats.rays,
                                       "sphere_tests": stats.sphere_tests / stats.rays},
             "device_bytes_per_ray_sah_tree": dev_bytes,
             "device_l1_request_gbs": dev_bytes * rays_launch / (kernel_ms * 1e-3) / 1e9,
             "l2_read_peak_gbs": peaks["l2_read_gbs"],
             "reference_tree_ops_per_ray": ops_per_ray, "reference_tree_bytes_per_ray": bytes_per_ray}
        r.update(common)
        return r
    achieved = ops_per_ray * rays_launch / (kernel_ms * 1e-3) / 1e12
    peak = peaks["fp32_nofma_ops"] / 1e12
    r = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
         "peak_source": "measured in this run by zrt_measure_peaks (FMUL/FADD chains, no FMA credit: parity forbids "
                        "contraction); MEASURED_PEAKS.json has no FP32-issue figure"}
    r.update(common)
    return r


def kernel_name_for(wl_name, flags):
    if wl_name in ("c1", "c5"):
        wl = WORKLOADS[wl_name]
        auto_pool = wl["w"] * wl["h"] * wl["spp"] >= (1 << 24)  # the library's rule for sphere-only scenes (zrt_api.cu makePlan)
        pool = (flags & A.ZRT_FLAG_KERNEL_POOL or auto_pool) and not flags & A.ZRT_FLAG_KERNEL_THREAD
        return "k_trace_pool3<7,128,7>" if pool else "k_trace<SPHERES,7>"
    return "k_trace_ws / k_trace<BVH>"


def side_configs(Z, host, peaks, flags):
    """c1..c4 at N = 1, a few steps each: the other BASELINE configurations in the driver's record."""
    out = {}
    for name in ("c1", "c2", "c3", "c4"):
        wl = WORKLOADS[name]
        try:
            hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0)).pin()
            p = params_for(wl, flags=flags if name == "c1" else 0)
            with Z.HostImage((wl["h"], wl["w"], 3)) as him:
                with Z.Scene(hs, device=0) as sc:
                    for _ in range(3):
                        sc.render(hs.camera, p, out=him.array)
                    ks, rays = [], 0
                    for _ in range(3):
                        _, c, tm = sc.render(hs.camera, p, out=him.array)
                        ks.append(tm.kernel_ms + tm.resolve_ms)
                        rays = c.rays_processed
                    stats = sc.trace_statistics(hs.camera, p) if name != "c1" else None
                e2e = []
                for _ in range(3):
                    t0 = time.perf_counter()
                    with Z.Scene(hs, device=0) as sc2:
                        sc2.render(hs.camera, p, out=him.array)
                    e2e.append(time.perf_counter() - t0)
            ops, byts, sample, events = algorithmic_model(name, wl)
            ms = float(np.mean(ks))
            rf = roofline_for(name, kernel_name_for(name, p.flags), ms, rays, ops, byts, sample, peaks, stats, events)
            out[name] = {"workload": wl["desc"], "ms": ms, "Mrays/s": rays / ms / 1e3, "bound": rf["bound"], "frac": rf["frac"],
                         "e2e_ms": 1e3 * float(np.mean(e2e)), "rays_per_step": rays,
                         "algorithmic_ops_per_ray": rf.get("device_algorithmic_ops_per_ray", ops),
                         "reference_tree_ops_per_ray": ops, "reference_tree_bytes_per_ray": byts,
                         "device_bytes_per_ray_sah_tree": rf.get("device_bytes_per_ray_sah_tree"),
                         "device_events_per_ray": rf.get("device_events_per_ray")}
            hs.close()
        except Exception as e:  # a side config must never cost the headline line
            out[name] = {"error": repr(e)}
    return out


def run_zrt(args, wl_name, wl):
    import torch
    import torch.distributed as dist

    from zraytrace_b200 import distributed as D
    from zraytrace_b200 import host
    from zraytrace_b200 import lib as Z

    torchrun = int(os.environ.get("WORLD_SIZE", "1")) > 1
    procs = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if Z.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; libzrt has no CPU path")
    world = procs if torchrun else max(1, args.gpus)       # ranks of the libzrt group = GPUs
    my_devices = [local_rank] if torchrun else list(range(world))
    if max(my_devices) >= Z.device_count():
        raise SystemExit(f"bench.py: --gpus {world} but only {Z.device_count()} devices are visible")
    torch.cuda.set_device(my_devices[0])
    if torchrun:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0)).pin()  # page-locked texels
    flags = (A.ZRT_FLAG_BVH_REFERENCE if args.reftree else 0) | KERNEL_FLAGS[args.kernel]
    params = params_for(wl, flags=flags)
    grp = D.create_group(hs, local_rank) if torchrun else Z.MultiScene(hs, devices=my_devices)
    flush = [torch.empty(256 << 20, dtype=torch.uint8, device=torch.device("cuda", d)) for d in my_devices]  # > 126 MB L2
    peaks = Z.measure_peaks(my_devices[0]) if rank == 0 else None

    def sync_all():
        for d in my_devices:
            torch.cuda.synchronize(d)
        if torchrun:
            dist.barrier()
            torch.cuda.synchronize()

    def flush_l2():
        for f in flush:
            f.zero_()
        for d in my_devices:
            torch.cuda.synchronize(d)
        if torchrun:
            dist.barrier()  # the ranks start a step together: rank 0's device time then measures the job, not the skew

    for _ in range(max(args.warmup, 3) if args.warmup else 0):
        grp.render(hs.camera, params, to_host=False)
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(1.0)  # the sampler is up and has history before the timed region starts
    sync_all()
    launches0 = grp.launch_count()
    t_wall0 = time.perf_counter()
    dev_ms, kern_ms, cnt = [], [], None
    for _ in range(args.steps):
        flush_l2()  # L2 flush between timed iterations, outside the per-step device-event pair
        _, c, tm = grp.render(hs.camera, params, to_host=False)  # device events on libzrt's stream: first launch -> scaled image
        dev_ms.append(tm.total_ms)
        kern_ms.append(tm.kernel_ms)
        cnt = c
    sync_all()
    t_wall1 = time.perf_counter()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None
    launches = grp.launch_count() - launches0
    t = torch.tensor([sum(dev_ms), float(launches)], dtype=torch.float64, device=torch.device("cuda", my_devices[0]))
    if torchrun:
        tmax = t.clone()
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        total_ms, launches = float(tmax[0].item()), int(t[1].item())
    else:
        total_ms = float(t[0].item())
    rays_per_step = int(cnt.rays_processed)  # valid on rank 0 (reduced)

    # kernel-only time of rank 0's share (for the roofline), CUDA events on the launching stream
    p_rank = D.rank_params(params, rank, world)
    dev0 = torch.device("cuda", my_devices[0])
    accum = torch.empty((wl["h"], wl["w"], 3), dtype=torch.float32, device=dev0)
    cnt_k = torch.zeros(6, dtype=torch.int64, device=dev0)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kernel_ms, stats = [], None
    with Z.Scene(hs, device=my_devices[0]) as scene:
        for i in range(4):
            flush[0].zero_()
            k0.record()
            scene.render_device(hs.camera, p_rank, accum.data_ptr(), cnt_k.data_ptr(), torch.cuda.current_stream().cuda_stream)
            k1.record()
            torch.cuda.synchronize()
            if i:
                kernel_ms.append(k0.elapsed_time(k1))
        if wl_name in ("c2", "c3", "c4") and rank == 0:
            stats = scene.trace_statistics(hs.camera, p_rank)
    kernel_ms = float(np.mean(kernel_ms))
    rays_rank = int(cnt_k.cpu().numpy().astype(np.uint64)[5])

    # end to end through the public API with HOST buffers, every step: scene flatten + upload (H2D), render, reduce,
    # image + counters back into the caller's page-locked image (D2H)
    h2d = hs.upload_bytes() + 256
    d2h = wl["w"] * wl["h"] * 12 + 48
    host_image = Z.HostImage((wl["h"], wl["w"], 3)) if rank == 0 else None
    e2e_times = []
    for i in range(2 + args.steps):
        sync_all()
        t0 = time.perf_counter()
        if world == 1:
            with Z.Scene(hs, device=my_devices[0]) as sc2:  # zrt_scene_create: flatten + H2D
                sc2.render(hs.camera, params, out=host_image.array)  # zrt_render: kernels + D2H into the host buffer
        else:
            grp.reload(hs)  # zrt_scene_create on every rank's device: flatten + H2D
            grp.render(hs.camera, params, out=host_image.array if rank == 0 else None)
        sync_all()
        if i >= 2:
            e2e_times.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_times))

    if rank != 0:
        if torchrun:
            dist.destroy_process_group()
        return

    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl_name}: {wl['desc']}",
                       "parallelism": f"spp-split x{world} + 1 ncclReduce (libzrt zrt_multi_render; "
                                      + ("one process per GPU, torchrun" if torchrun else "one process drives all GPUs") + ")",
                       "kernel": kernel_name_for(wl_name, flags),
                       "rays_per_step": rays_per_step, "samples_per_step": int(cnt.samples_processed),
                       "l2": "256 MiB device memset between timed steps (outside the per-step event pairs); "
                             "scene data is <= a few MB and stays cache resident by design",
                       "timing": "per step: CUDA events on libzrt's stream, first launch -> reduced and scaled image on rank 0; "
                                 "sum over steps, max over ranks",
                       "bvh": ("reference topology" if args.reftree else "binned SAH over the reference's surviving primitives") if wl_name in ("c2", "c3", "c4") else "none (surface list)", "seed": 42,
                       "trace_ms_per_step_slowest_rank": float(np.mean(kern_ms)), "nccl_version": Z.nccl_version() if world > 1 else None,
                       "wall_s_timed_region": t_wall1 - t_wall0},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": rays_per_step / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "steps": len(e2e_times),
                    "path": "zrt_scene_create (H2D: page-locked texels, pageable primitive arrays) + zrt_render into a page-locked host image (D2H) per step" if world == 1 else
                            "zrt_multi_reload (zrt_scene_create per rank) + zrt_multi_render (trace, ncclReduce, 1/spp, D2H into a page-locked host image) per step"},
            "published_reference": {"value": 3.47, "unit": "Mrays/s", "note": "README.md:49-61, unknown CPU, 1 thread"}}
    ops_per_ray, bytes_per_ray, model_sample, model_events = algorithmic_model(wl_name, wl)
    if world == 1 and not args.no_cpu:
        line["cpu_baseline"] = cpu_baseline(wl_name, wl)
    line["roofline"] = roofline_for(wl_name, kernel_name_for(wl_name, flags), kernel_ms, rays_rank, ops_per_ray, bytes_per_ray,
                                    model_sample, peaks, stats, model_events)
    if world == 1 and wl_name == "c5" and not args.no_configs:
        line["configs"] = side_configs(Z, host, peaks, flags)
    emit(line)
    if torchrun:
        dist.destroy_process_group()


KERNEL_FLAGS = {"auto": 0, "thread": A.ZRT_FLAG_KERNEL_THREAD, "warp": A.ZRT_FLAG_KERNEL_WARP, "pool": A.ZRT_FLAG_KERNEL_POOL}
_JSON_FD = None


def emit(line):
    """The one JSON line goes to the process's ORIGINAL stdout; everything else any library prints while the
    benchmark runs (NCCL prints its version banner on stdout) has been pointed at stderr by main()."""
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)  # C-level stdout of every library -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="c5", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="zrt", choices=["zrt", "reference"])
    ap.add_argument("--kernel", default="auto", choices=sorted(KERNEL_FLAGS), help="force a kernel variant (default: the library's choice)")
    ap.add_argument("--reftree", action="store_true", help="BVH workloads: traverse the reference topology, not the SAH tree")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1..c4 side block of the default line")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_zrt(args, args.workload, wl)


if __name__ == "__main__":
    main()
