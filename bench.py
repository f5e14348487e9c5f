#!/usr/bin/env python
"""bench.py — headline benchmark of the zraytrace hot path on B200.

Metric (BASELINE.json): Mrays/s, device-timed, whole job over N GPUs; a "step" is one full render of the
workload (default c5 = the README headline: 7-spheres, 1000x1000, 1000 spp, depth 30, spp split across
ranks, one NCCL reduce of the fp32 accumulators).  rays = the reference's `rays_processed` counter
(raytrace.zig:69).  See DESIGN.md "Measurement" for every field of the JSON line.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c1|c2|c3|c4|c5] [--impl zrt|reference]
  python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...
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

from zraytrace_b200 import _abi as A  # noqa: E402

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


# dram__bytes_read.sum + dram__bytes_write.sum of ONE k_trace launch at the bench size, from the committed
# `ncu --set full` capture named here (a number measured under a profiler is only ever used for this field)
NCU_TRAFFIC = {"c5": {"bytes": 3894528 + 44411136, "source": "profiles/r1_v10_c5_k_trace.txt (1 GPU, 1000 spp)"}}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 100 ms while the timed region runs."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_indices):
        self.idx = set(gpu_indices)
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        self.p.wait()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, power = [], [], set(), []
        for line in self.f:
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9 or not c[0].isdigit() or int(c[0]) not in self.idx:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2])); power.append(float(c[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.f.name)
        busy = [s for s, p in zip(sm, power) if p > 250.0] or sm
        return {"sm_mhz": float(np.median(busy)) if busy else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "power_w_max": max(power) if power else None}


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return json.load(open(path)), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback (B200_PROFILING.md)"


def params_for(wl, **kw):
    return A.make_params(wl["w"], wl["h"], wl["spp"], wl["depth"], bvh=True, x_limit=wl.get("x_limit", A.ZRT_XLIMIT_HEIGHT),
                         seed=42, **kw)


# ------------------------------------------------------------------------------------------- reference arm
def run_reference(args, wl_name, wl):
    """--impl reference: the reference's CPU path (the oracle port; the Zig original cannot be built in this
    image) on the host cores, same workload plane, bounded spp per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import zro_py
    from zraytrace_b200 import host

    cores = os.cpu_count() or 1
    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0))
    # bounded sample: same image plane and depth, spp reduced so one step is a few seconds on this host
    target_rays = 6e6 * cores * (1.0 if wl_name in ("c1", "c5") else 0.02)
    rays_per_sample = 2.15 if wl_name in ("c1", "c5") else 1.6
    spp = int(max(1, min(wl["spp"], target_rays / (wl["w"] * wl["h"] * rays_per_sample))))
    p = A.make_params(wl["w"], wl["h"], spp, wl["depth"], bvh=True, x_limit=wl.get("x_limit", 0), seed=42)
    times, rays = [], 0
    for i in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        _, c, _ = zro_py.render(hs, hs.camera, p, rng=zro_py.RNG_CTR, traversal=zro_py.TRAVERSAL_REF,
                                math=zro_py.MATH_LIBM, threads=cores)
        dt = time.perf_counter() - t0
        if i >= args.warmup:
            times.append(dt)
            rays = c.rays_processed
    total = sum(times)
    value = rays * len(times) / total / 1e6
    sample = (f"oracle port of the reference CPU path (Zig original not buildable here), {wl['w']}x{wl['h']} plane, "
              f"{spp} spp per step instead of {wl['spp']}, depth {wl['depth']}, {cores} threads over scanlines, "
              f"literal aabb.zig traversal, glibc transcendentals")
    line = {"impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl_name}: {wl['desc']}", "spp_per_step": spp},
            "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------- CPU baseline leg
def cpu_baseline(wl_name, wl, hs):
    """Oracle timed on ONE host core in its most literal mode (sequential Xoroshiro stream, libm, literal
    traversal): the reference is single-threaded (README.md:11).  Bounded sample, ~10-20 s."""
    from oracle import zro_py

    heavy = wl_name not in ("c1", "c5")
    spp = 1 if heavy else max(1, min(wl["spp"], int(60e6 / (wl["w"] * wl["h"] * 2.15))))
    div = 16 if wl_name == "c4" else 4  # literal aabb.zig traversal visits ~10^3 nodes per ray
    w, h = (wl["w"] // div, wl["h"] // div) if heavy else (wl["w"], wl["h"])
    p = A.make_params(w, h, spp, wl["depth"], bvh=True, x_limit=wl.get("x_limit", 0), seed=42)
    t0 = time.perf_counter()
    _, c, st = zro_py.render(hs, hs.camera, p, rng=zro_py.RNG_REF, traversal=zro_py.TRAVERSAL_REF, math=zro_py.MATH_LIBM)
    dt = time.perf_counter() - t0
    # event counts for the algorithmic-op model come from the tight traversal (the reference's own visit
    # counts include the Q4 defect and are not the algorithmic figure)
    if heavy:
        _, c2, st2 = zro_py.render(hs, hs.camera, p, rng=zro_py.RNG_CTR, traversal=zro_py.TRAVERSAL_TIGHT)
    else:
        c2, st2 = c, st
    ops_per_ray = algorithmic_ops(st2.as_dict(), c2.as_dict()) / c2.rays_processed
    bytes_per_ray = (64 * st2.box_passes + 48 * st2.triangle_tests + 16 * st2.sphere_tests
                     + 12 * c2.samples_processed + 4 * st2.texture_lookups) / c2.rays_processed
    return {"value": c.rays_processed / dt / 1e6, "unit": "Mrays/s", "cores": 1, "kind": "port",
            "sample": (f"oracle port, single thread, sequential Xoroshiro128+ stream, glibc math, literal aabb.zig traversal; "
                       f"{w}x{h} plane at {spp} spp (full workload is {wl['w']}x{wl['h']} at {wl['spp']} spp), {dt:.1f} s"),
            "rays_per_sample": c.rays_processed / c.samples_processed}, ops_per_ray, bytes_per_ray


# ------------------------------------------------------------------------------------------- our arm
def run_zrt(args, wl_name, wl):
    import torch
    import torch.distributed as dist

    from zraytrace_b200 import distributed as D
    from zraytrace_b200 import host
    from zraytrace_b200 import lib as Z

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if Z.device_count() < 1:
        raise SystemExit("bench.py: no CUDA device visible; libzrt has no CPU path")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    hs = host.HostScene(wl["scene"], variant=wl["variant"], aspect_ratio=wl.get("aspect", 1.0)).pin()  # page-locked texels
    flags = A.ZRT_FLAG_BVH_REFERENCE if args.reftree else 0
    params = params_for(wl, flags=flags)
    scene = Z.Scene(hs, device=local_rank)
    accum = torch.empty((wl["h"], wl["w"], 3), dtype=torch.float32, device=dev)
    counters = torch.zeros(6, dtype=torch.int64, device=dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    peaks = Z.measure_peaks(local_rank) if rank == 0 else None

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step():
        D.render_distributed(scene, hs.camera, params, accum, counters)

    for _ in range(max(args.warmup, 3) if args.warmup else 0):
        step()
    barrier()
    launches0 = scene.launch_count()
    sampler = ClockSampler(range(world)) if rank == 0 else None
    if sampler:
        sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    t_wall0 = time.perf_counter()
    for s, e in ev:
        flush.zero_()  # L2 flush between timed iterations, outside the per-step event pair
        s.record()
        step()
        e.record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    clocks = sampler.stop() if sampler else None
    my_ms = sum(s.elapsed_time(e) for s, e in ev)
    launches = scene.launch_count() - launches0
    t = torch.tensor([my_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t.item())
    cnt = counters.cpu().numpy().astype(np.uint64)  # valid on rank 0 (reduced)
    rays_per_step = int(cnt[5])

    # kernel-only time of this rank's share (for the roofline), CUDA events on the launching stream
    p_rank = D.rank_params(params, rank, world)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    cnt_k = torch.zeros(6, dtype=torch.int64, device=dev)
    kernel_ms = []
    for _ in range(3):
        flush.zero_()
        k0.record()
        scene.render_device(hs.camera, p_rank, accum.data_ptr(), cnt_k.data_ptr(), torch.cuda.current_stream().cuda_stream)
        k1.record()
        torch.cuda.synchronize()
        kernel_ms.append(k0.elapsed_time(k1))
    kernel_ms = float(np.mean(kernel_ms))
    rays_rank = int(cnt_k.cpu().numpy().astype(np.uint64)[5])

    # end to end through the public API with HOST buffers: scene upload (H2D), render, image + counters back
    # (D2H) every step; N>1: every rank renders its share and rank 0 receives the reduced image
    e2e_times = []
    h2d = hs.upload_bytes() + 256
    d2h = wl["w"] * wl["h"] * 12 + 48
    pinned = torch.empty((wl["h"], wl["w"], 3), dtype=torch.float32).pin_memory() if world > 1 else None
    host_image = Z.HostImage((wl["h"], wl["w"], 3)) if world == 1 else None  # the caller's result image (zrt_pinned_alloc)
    for i in range(2 + min(args.steps, 3)):
        barrier()
        t0 = time.perf_counter()
        if world == 1:
            with Z.Scene(hs, device=local_rank) as sc2:  # zrt_scene_create: flatten + H2D
                img, c_e2e, _ = sc2.render(hs.camera, params, out=host_image.array)  # zrt_render: kernels + D2H into the host buffer
        else:
            with Z.Scene(hs, device=local_rank) as sc2:
                a2, c2 = D.render_distributed(sc2, hs.camera, params)
                if rank == 0:
                    pinned.copy_(a2, non_blocking=True)
                    c2.cpu()
                torch.cuda.synchronize()
        barrier()
        if i >= 2:
            e2e_times.append(time.perf_counter() - t0)
    e2e_s = float(np.mean(e2e_times))

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = rays_per_step * args.steps / (total_ms * 1e-3) / 1e6
    peaks_file, peaks_src = load_peaks()
    line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{wl_name}: {wl['desc']}", "parallelism": f"spp-split x{world} + 1 NCCL reduce",
                       "rays_per_step": rays_per_step, "samples_per_step": int(cnt[4]),
                       "l2": "256 MiB device memset between timed steps (outside the per-step event pairs); "
                             "scene data is <= a few MB and stays cache resident by design",
                       "bvh": ("reference topology" if flags else "binned SAH over the reference's surviving primitives") if wl_name in ("c2", "c3", "c4") else "none (surface list)", "seed": 42,
                       "wall_s_timed_region": t_wall},
            "gpu_launches": int(launches), "clocks": clocks,
            "e2e": {"value": rays_per_step / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                    "path": "zrt_scene_create (H2D: page-locked texels, pageable primitive arrays) + zrt_render into a page-locked host image (D2H) per step" if world == 1 else
                            "zrt_scene_create + zrt_render_device + NCCL reduce + D2H per step"},
            "published_reference": {"value": 3.47, "unit": "Mrays/s", "note": "README.md:49-61, unknown CPU, 1 thread"}}
    cpu = None
    if world == 1 and not args.no_cpu:
        cpu, ops_per_ray, bytes_per_ray = cpu_baseline(wl_name, wl, hs)
        line["cpu_baseline"] = cpu
    else:
        ops_per_ray = {"c1": 201.0, "c5": 200.94}.get(wl_name)  # measured by the cpu_baseline leg at N=1 (oracle events)
        bytes_per_ray = None
    is_bvh = wl_name in ("c2", "c3", "c4")
    if is_bvh:
        # byte side: event counts from the instrumented build of the same kernel on this rank's share
        st = scene.trace_statistics(hs.camera, p_rank)
        alg_bytes = (64 * st.node_visits + 48 * st.triangle_tests + 32 * st.sphere_tests + 12 * st.samples
                     + 4 * st.texture_lookups)
        achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
        peak = peaks["l2_read_gbs"]
        line["roofline"] = {"bound": "l2", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                            "traffic": None, "kernel": "k_trace", "kernel_ms": kernel_ms, "rays_per_launch": rays_rank,
                            "algorithmic_bytes_per_ray": alg_bytes / max(st.rays, 1),
                            "events_per_ray": {"node_visits": st.node_visits / max(st.rays, 1),
                                               "triangle_tests": st.triangle_tests / max(st.rays, 1),
                                               "sphere_tests": st.sphere_tests / max(st.rays, 1)},
                            "peak_source": "L2-resident 128-bit read bandwidth measured in this run by zrt_measure_peaks "
                                           "(MEASURED_PEAKS.json has no L2 figure); scene data is L1/L2 resident, HBM is idle",
                            "k0": peaks, "hbm_peak_gbs": peaks_file.get("hbm_gbs"), "hbm_peak_source": peaks_src,
                            "fp32_ops_per_ray_reference_tree": ops_per_ray,
                            "mrays_per_s_kernel_only": rays_rank / (kernel_ms * 1e-3) / 1e6}
    elif ops_per_ray:
        achieved = ops_per_ray * rays_rank / (kernel_ms * 1e-3) / 1e12
        peak = peaks["fp32_nofma_ops"] / 1e12
        line["roofline"] = {"bound": "fp32", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                            "traffic": NCU_TRAFFIC.get(wl_name, {}).get("bytes") if world == 1 else None,
                            "traffic_source": NCU_TRAFFIC.get(wl_name, {}).get("source"),
                            "kernel": "k_trace", "kernel_ms": kernel_ms, "rays_per_launch": rays_rank,
                            "algorithmic_ops_per_ray": ops_per_ray, "algorithmic_bytes_per_ray": bytes_per_ray,
                            "peak_source": "measured in this run by zrt_measure_peaks (FMUL/FADD chains, no FMA credit: "
                                           "parity forbids contraction); MEASURED_PEAKS.json has no FP32-issue figure",
                            "k0": peaks, "hbm_peak_gbs": peaks_file.get("hbm_gbs"), "hbm_peak_source": peaks_src,
                            "mrays_per_s_kernel_only": rays_rank / (kernel_ms * 1e-3) / 1e6}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


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
    ap.add_argument("--reftree", action="store_true", help="BVH workloads: traverse the reference topology, not the SAH tree")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, args.workload, wl)
    else:
        run_zrt(args, args.workload, wl)


if __name__ == "__main__":
    main()
