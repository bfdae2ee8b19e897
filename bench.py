#!/usr/bin/env python
"""Headline benchmark: Mpixel-samples/s of the path-tracing hot path on BASELINE.json's config C2
(Cornell box 1920x1080, 16 bounces, importance sampling + direct lighting, 128 spp per launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c2|c3|c4|c5]

A "step" is one launch (= one optixLaunch(w, h, 1) of the reference, PathTracerMain.cpp:197): samplesPerPixel
samples for every pixel folded into the running average; step i uses currentFrameIdx = i, so K = 8 steps are the
1024 spp of config C2.  Prints ONE JSON line on rank 0 (contract in the task statement; extra keys: roofline,
cpu_baseline, e2e, clocks, gpu_launches, grays_per_s).

For N > 1 (torchrun, one rank per GPU) the scene and LBVH are replicated, 8x4-pixel tiles are interleaved over
ranks.  The K launches of the image are dealt to the ranks as WHOLE frames (frame k -> rank k % N, K // N rounds; the
mean image of a frame depends only on (pixel, frame index), pathTracerPrograms.cu:721); the K % N leftover frames are
rendered by all ranks together, 8x4-pixel tiles interleaved.  The timed region ends with ONE exchange (NCCL all-gather of
the per-frame means) and the reference's running average applied in frame order (pt_fold_frames), after which every
rank holds the complete image, bit-identical to K launches on one GPU (acgpathtracing_b200/multirank.py).
`--partition tiles` forces the pure tile partition instead.  Total work is fixed as N grows ("strong" scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
_emit = print          # replaced in __main__: the JSON line goes to the real stdout, everything else to stderr
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (scene, width, height, spp per launch, maxDepth, NEE, importance sampling, description)
    "c1": ("cornell", 512, 512, 1, 3, False, True, "Cornell-32 512x512, 1 spp, 3 bounces, IS, no NEE (configs[0])"),
    "c2": ("cornell", 1920, 1080, 128, 16, True, True, "Cornell-32 1920x1080, 16 bounces, IS+NEE, 128 spp/launch (configs[1]; 8 launches = 1024 spp)"),
    "c3": ("material", 1920, 1080, 128, 28, True, True, "Cornell + dielectric + conductor spheres 1080p, 28 bounces, IS+NEE, 128 spp/launch (configs[2])"),
    "c4": ("tess1000", 1920, 1080, 128, 8, True, True, "2.0 M-triangle displaced height field in the Cornell box, 1080p, 8 bounces, 128 spp/launch (configs[3])"),
    "c5": ("cornell", 3840, 2160, 128, 28, True, True, "Cornell-32 3840x2160, 28 bounces, IS+NEE, 128 spp/launch (configs[4])"),
}

# Algorithmic work per event (SURVEY.md 8d / DESIGN.md "Roofline accounting"); FMA = 2 flop.
FLOP_NODE, FLOP_TRI = 24, 45
FLOP_SAMPLE, FLOP_SEGMENT, FLOP_NEE = 40, 20, 60
FLOP_HIT = (110, 190, 120)
BYTES_NODE, BYTES_TRI = 64, 48


def algorithmic_work(c, nee):
    flops = (FLOP_NODE * c["n_node"] + FLOP_TRI * c["n_tri"] + FLOP_SAMPLE * c["n_samples"] +
             FLOP_SEGMENT * c["n_radiance"] + sum(f * n for f, n in zip(FLOP_HIT, c["n_hit"])))
    if nee:
        flops += FLOP_NEE * (c["n_hit"][0] + c["n_hit"][1])
    return flops, BYTES_NODE * c["n_node"] + BYTES_TRI * c["n_tri"]


def make_scene_arrays(name):
    from acgpathtracing_b200 import scenes as S
    if name == "cornell":
        return S.cornell32()
    if name == "material":
        return S.material_scene(3)
    if name.startswith("tess"):
        return S.tessellated_scene(int(name[4:]))
    raise ValueError(name)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self.proc = None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def stop(self):
        if self.proc:
            self.proc.terminate()
        self.join(timeout=2)
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
                power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                continue
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm),
                "power_w_max": max(power)}


# ================================================================================================================
# reference arm: the reference's own programs on the host cores
# ================================================================================================================
def cpu_reference_run(workload, steps, warmup, prefer_ref=True, target_step_s=2.0):
    """Times the reference's CPU implementation of the path on this box's host cores with all hardware threads:
    oracle/_ref/libptref.so (the reference's own pathTracerPrograms.cu compiled for the host) when it exists, else
    the oracle port.  Each step is a bounded sample of the workload: the full frame at a reduced spp."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    scene_name, w, h, spp, md, nee, imp, desc = WORKLOADS[workload]
    arr = make_scene_arrays(scene_name)
    port = O.load_port()
    ref = O.load_ref() if prefer_ref else None
    cores = port.ptor_hardware_threads()
    use_bvh = arr.n_tri > 64            # brute force over <= 64 triangles is the faster CPU choice (measured)
    osc = None if ref is not None else O.OracleScene(port, arr)

    def one(frame, s_spp, region=None):
        p = O.make_params(port, w, h, s_spp, md, frame=frame, nee=nee, importance=imp)
        t0 = time.perf_counter()
        if ref is not None:
            O.ref_render(ref, arr, p, use_bvh=use_bvh, nthreads=cores, region=region)
        else:
            osc.render(p, math_mode=0, use_bvh=use_bvh, nthreads=cores, region=region, want_fb=True)
        return time.perf_counter() - t0

    rows = max(8, h // 8)
    t = one(0, 1, region=(0, w, (h - rows) // 2, (h + rows) // 2))      # calibration strip from the image centre
    rate = w * rows / max(t, 1e-6)
    s_spp = int(max(1, min(spp, round(target_step_s * rate / (w * h)))))
    for i in range(warmup):
        one(i, s_spp)
    times = [one(i, s_spp) for i in range(steps)]
    total = sum(times)
    value = w * h * s_spp * steps / total / 1e6
    kind = "reference" if ref is not None else "port"
    sample = (f"{desc}: full {w}x{h} frame at {s_spp} spp per step ({steps} steps, frames 0..{steps - 1}), "
              f"{'CPU BVH traversal (the tree the GPU walks)' if use_bvh else 'brute-force'} intersection, {cores} threads, glibc libm")
    return dict(value=value, unit="Mpixel-samples/s", cores=cores, kind=kind, sample=sample,
                ms_per_step=1e3 * total / steps, spp=s_spp)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    scene_name, w, h, spp, md, nee, imp, desc = WORKLOADS[args.workload]
    r = cpu_reference_run(args.workload, args.steps, args.warmup)
    out = {
        "impl": "reference", "metric": "Mpixel-samples/s", "value": r["value"], "unit": "Mpixel-samples/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_step"],
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "width": w, "height": h, "max_depth": md, "spp_per_step": r["spp"],
                   "direct_lighting": nee, "importance_sampling": imp},
        "cpu_baseline": {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]},
        "e2e": {"value": r["value"], "unit": r["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _emit(json.dumps(out))


# ================================================================================================================
# B200 arm
# ================================================================================================================
def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from acgpathtracing_b200 import api
    from acgpathtracing_b200.types import PathTraceParams

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    if args.gpus > 1 and world == 1:
        raise SystemExit("for --gpus N > 1 launch with: python -m torch.distributed.run --nproc-per-node N bench.py --gpus N ...")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    scene_name, w, h, spp, md, nee, imp, desc = WORKLOADS[args.workload]
    if args.spp:
        spp = args.spp
    arr = make_scene_arrays(scene_name)
    scene = api.Scene(arr)
    if args.pipeline:
        scene.set_pipeline(args.pipeline)
    if args.tree:
        scene.set_tree(args.tree)
    info = scene.info()
    stream = torch.cuda.current_stream().cuda_stream

    acc = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
    fb = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
    from acgpathtracing_b200.multirank import FrameParallel, ImageExchange
    exchange = ImageExchange(w, h, rank, world, dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    frame_mode = world > 1 and args.partition == "auto"

    def params(frame, acc_ptr=None, fb_ptr=None):
        p = api.make_params(w, h, spp, md, frame=frame, nee=nee, importance=imp, scene=scene)
        p.accumulationBuffer = acc.data_ptr() if acc_ptr is None else acc_ptr
        p.frameBuffer = fb.data_ptr() if fb_ptr is None else fb_ptr
        return p

    def render_raw(frame, r, n, out):
        api.launch_raw(params(frame, out.data_ptr(), 0), r, n, stream)

    fp = FrameParallel(w, h, rank, world, dev, args.steps, render_raw) if frame_mode else None

    def my_steps():
        """The launches this rank issues for one image: [(frame, kind)]."""
        return fp.my_frames() if frame_mode else [(i, "seq") for i in range(args.steps)]

    def step(frame, kind="seq"):
        """One launch of the hot path with device-resident buffers; returns the number of our kernels launched."""
        if kind != "seq":
            return fp.render(frame, kind)
        p = params(frame)
        if args.emulate_ranks > 1:
            api.launch_partition(p, 0, args.emulate_ranks, stream)   # tuning aid: one rank's share of an N-rank job
        elif world == 1:
            api.launch(p, stream)
        else:
            api.launch_partition(p, rank, world, stream)
        return 1

    def finish():
        """N > 1: the one exchange of the image; returns the number of our kernels launched."""
        if frame_mode:
            return fp.finish(acc, fb, stream)
        return exchange.gather(acc, fb, stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------------------
    if frame_mode:                                    # at least args.warmup launches per rank, in whole images
        reps = -(-args.warmup // max(len(my_steps()), 1))
        for _ in range(max(reps, 1)):
            for f, kind in my_steps():
                step(f, kind)
    else:
        for i in range(args.warmup):
            step(i)
    finish()
    barrier()

    # ---- timed region: the image's K launches (this rank's share of them), CUDA events around each, L2 flushed between
    # launches, then the exchange ---------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    todo = my_steps()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(len(todo) + 1)]
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    for i, (f, kind) in enumerate(todo):
        flush_buf.fill_(i & 0xFF)                     # evict L2 (untimed)
        ev[i][0].record()
        launches += step(f, kind)
        ev[i][1].record()
    ev[-1][0].record()
    launches += finish()                              # N > 1: exchange (+ ordered fold), timed
    ev[-1][1].record()
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    exchange_ms = step_ms[-1]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    clocks = sampler.stop() if sampler else None
    checksum = float(acc[..., :3].double().mean().item())
    acc_after_timed = acc.clone()           # for the bit-for-bit comparison with the host-buffer path below

    pixel_samples = float(w) * h * spp * args.steps / max(args.emulate_ranks, 1)
    value = pixel_samples / (total_ms * 1e-3) / 1e6

    # ---- work counters for the timed frames (instrumented kernel, untimed) -> Grays/s and algorithmic flops ----
    scene.enable_counters(True)
    scene.counters(reset=True)
    for f, kind in my_steps():
        step(f, kind)
    finish()
    barrier()
    c = scene.counters(reset=True)
    sched = scene.sched_stats(reset=True)
    scene.enable_counters(False)
    ct = torch.tensor([c["n_samples"], c["n_radiance"], c["n_shadow"], c["n_node"], c["n_tri"], *c["n_hit"], c["n_miss"]],
                      dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(ct)
    ct = [int(x) for x in ct.tolist()]
    counters = dict(n_samples=ct[0], n_radiance=ct[1], n_shadow=ct[2], n_node=ct[3], n_tri=ct[4], n_hit=ct[5:8], n_miss=ct[8])
    flops, tbytes = algorithmic_work(counters, nee)
    rays = counters["n_radiance"] + counters["n_shadow"]

    # ---- end to end through the host-buffer entry point (pt_launch_host): pinned host buffers, H2D + D2H in the
    # timed region every step -------------------------------------------------------------------------------------
    e2e = None
    if world == 1:
        h_acc = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
        h_fb = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
        up = down = 0
        import ctypes as C

        def host_step(frame):
            p = api.make_params(w, h, spp, md, frame=frame, nee=nee, importance=imp, scene=scene)
            p.accumulationBuffer, p.frameBuffer = h_acc.data_ptr(), h_fb.data_ptr()
            a, b = C.c_uint64(), C.c_uint64()
            api._check(api.lib.pt_launch_host(C.byref(p), C.byref(a), C.byref(b)))
            return a.value, b.value

        for i in range(min(args.warmup, 2)):
            host_step(i)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for i in range(args.steps):
            a, b = host_step(i)       # synchronous: returns after the D2H copies completed
            up += a
            down += b
        t_e2e = time.perf_counter() - t0
        e2e = {"value": pixel_samples / t_e2e / 1e6, "unit": "Mpixel-samples/s",
               "h2d_bytes_per_step": up // args.steps, "d2h_bytes_per_step": down // args.steps,
               "ms_per_step": 1e3 * t_e2e / args.steps,
               "bit_identical_to_device_path": bool(torch.equal(h_acc.view(torch.int32), acc_after_timed.cpu().view(torch.int32)))}
    else:
        # every rank: per launch the params record goes up and the launch's result comes down to a pinned host buffer
        # (tile mode: previous accumulation up, accumulation + frame buffer down); after the last launch the exchange and
        # a read of the complete image
        h_acc = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
        h_fb = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()
        up = down = 0

        def host_step(i, f, kind):
            nonlocal up, down
            if kind == "seq":
                if f > 0:
                    acc.copy_(h_acc, non_blocking=True)
                    up += w * h * 16
                step(f, kind)
                h_acc.copy_(acc, non_blocking=True)
                h_fb.copy_(fb, non_blocking=True)
                down += w * h * 20
            else:
                step(f, kind)
                src = fp.mine[f // world] if kind == "whole" else fp.shared[f - fp.rounds * world]
                h_acc.copy_(src, non_blocking=True)
                down += w * h * 16
            up += 168
            torch.cuda.synchronize()

        def host_finish():
            nonlocal down
            finish()
            h_acc.copy_(acc, non_blocking=True)
            h_fb.copy_(fb, non_blocking=True)
            down += w * h * 20
            torch.cuda.synchronize()

        for i, (f, kind) in enumerate(todo[:1]):
            host_step(i, f, kind)
        host_finish()
        barrier()
        up = down = 0
        t0 = time.perf_counter()
        for i, (f, kind) in enumerate(todo):
            host_step(i, f, kind)
        host_finish()
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_e2e = float(t_e2e.item())
        e2e = {"value": pixel_samples / t_e2e / 1e6, "unit": "Mpixel-samples/s",
               "h2d_bytes_per_step": up // args.steps, "d2h_bytes_per_step": down // args.steps,
               "ms_per_step": 1e3 * t_e2e / args.steps,
               "note": "bytes moved by rank 0 divided by the image's K steps; includes the final full-image read",
               "bit_identical_to_device_path": bool(torch.equal(h_acc.view(torch.int32), acc_after_timed.cpu().view(torch.int32)))}

    # ---- roofline denominators measured on this device -----------------------------------------------------------
    roofline = None
    cpu = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        dinfo = api.device_info()
        fp32_meas = api.bench_fp32_peak()
        l2_meas = api.bench_l2_bandwidth()
        sm_max = (clocks or {}).get("sm_max_mhz") or peaks.get("sm_max_mhz") or 1965.0
        fp32_nominal = dinfo["sm_count"] * 128 * 2 * sm_max * 1e6 / 1e12
        # average duration of one render kernel on rank 0 (its own launches, exchange excluded)
        kernel_ms = sum(step_ms[:-1]) / max(len(todo), 1)
        per_launch_share = 1.0 / args.steps if (world == 1 or frame_mode) else 1.0 / args.steps / world
        if frame_mode and fp.leftover:        # mixed whole / tile launches: use the job-wide average per launch issued
            per_launch_share = 1.0 / (world * len(todo))
        ach_tflops = flops * per_launch_share / (kernel_ms * 1e-3) / 1e12
        ach_gbs = tbytes * per_launch_share / (kernel_ms * 1e-3) / 1e9
        roofline = {
            "kernel": "render_persistent_kernel" if args.pipeline == 1 else "render_wavefront_kernel",
            "bound": "fp32", "achieved": ach_tflops, "peak": fp32_meas, "unit": "TFLOP/s", "frac": ach_tflops / fp32_meas,
            "peak_source": "measured live by pt_bench_fp32_peak (dependent-free FFMA loop); MEASURED_PEAKS.json has no FP32 figure",
            "peak_nominal": fp32_nominal,
            # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed ncu --set full capture of this
            # command on C2 (profiles/r1_v3g_sah_wavefront_ncu_summary.txt: 682,496 + 522,752 B); other workloads: not captured
            "traffic": 1205248 if (args.workload == "c2" and args.pipeline != 1 and not args.spp and not args.tree) else None,
            "l2": {"achieved": ach_gbs, "peak": l2_meas, "unit": "GB/s", "frac": ach_gbs / l2_meas,
                   "note": "node/triangle record bytes; served from shared memory when the scene is staged there"},
            "hbm": {"achieved": (w * h * 36) / (kernel_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"),
                    "unit": "GB/s", "note": "image traffic only: 16 B read + 16 B write + 4 B framebuffer per pixel per launch"},
            "algorithmic_flops_per_launch": flops / args.steps, "algorithmic_bytes_per_launch": tbytes / args.steps,
            "algorithmic_note": "work of the SPECIFIED traversal (every ray through the tree named in config.pipeline.tree), counted by "
                                "the instrumented kernel and equal to the oracle's counts; the timed kernel resolves camera rays of "
                                "tiny scenes from per-pixel candidate lists instead, which lowers the work done, not the work credited",
            "per_gpu": world > 1,
        }
        if world == 1 and not args.no_cpu_baseline:
            r = cpu_reference_run(args.workload, steps=1, warmup=0, prefer_ref=False, target_step_s=12.0)
            cpu = {"value": r["value"], "unit": r["unit"], "cores": r["cores"], "kind": r["kind"], "sample": r["sample"]}

    if rank == 0:
        out = {
            "metric": "Mpixel-samples/s", "value": value, "unit": "Mpixel-samples/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "scene": scene_name, "triangles": info["n_tris"], "width": w, "height": h,
                       "spp_per_step": spp, "max_depth": md, "direct_lighting": nee, "importance_sampling": imp,
                       "frames": f"0..{args.steps - 1}",
                       "parallelism": f"pixel-tiles x{world}" + (f" (EMULATED rank 0 of {args.emulate_ranks}: not a bench value)" if args.emulate_ranks > 1 else ""),
                       "l2": "flushed between timed steps (256 MiB fill)", "pipeline": dict(scene.last_launch_info(), tree=scene.tree_in_use()),
                       "exchange": "none (1 GPU)" if world == 1 else (
                           f"frames dealt to ranks ({fp.rounds} whole frame(s) per rank, {len(fp.leftover)} tile-partitioned); one "
                           f"all-gather of per-frame means + ordered fold per image, {fp.bytes_received()} B received per rank, "
                           f"{exchange_ms:.3f} ms on rank 0" if frame_mode else
                           f"one all-gather of owned pixels per image, {exchange.bytes_per_rank()} B received per rank, {exchange_ms:.3f} ms on rank 0"),
                       "lbvh_build_ms": info["build_ms"]},
            "grays_per_s": rays / (total_ms * 1e-3) / 1e9,
            "rays_per_sample": rays / max(counters["n_samples"], 1),
            "counters": counters, "sched_rank0": sched, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
            "e2e": e2e, "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "image_mean_radiance": checksum,
        }
        _emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=0, help="override samples per pixel per launch")
    ap.add_argument("--pipeline", type=int, default=0, help="0 auto, 1 persistent state machine, 2 CTA-local wavefront")
    ap.add_argument("--tree", type=int, default=0, help="0 auto (sweep-SAH tree for scenes of 2..8192 triangles, PLOC above), 1 LBVH, 2 SAH, 3 PLOC")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--partition", default="auto", choices=["auto", "tiles"],
                    help="N > 1: auto = whole frames per rank + tile-partitioned leftovers; tiles = every launch tile-partitioned")
    ap.add_argument("--emulate-ranks", type=int, default=0,
                    help="tuning aid (1 GPU): render only rank 0's tiles of an R-rank partition; value counts those pixels only")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    # stdout carries exactly one line, the JSON: anything a library prints there while we run (NCCL's version banner)
    # goes to stderr instead
    _json_fd = os.dup(1)
    os.dup2(2, 1)
    _emit = lambda line: os.write(_json_fd, (line + "\n").encode())
    main()
