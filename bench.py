#!/usr/bin/env python
"""Headline benchmark: Mpixel-samples/s of the path-tracing hot path on BASELINE.json's config C2
(Cornell box 1920x1080, 16 bounces, importance sampling + direct lighting, 128 spp per launch).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c1|c2|c3|c4|c5]

A "step" is one launch (= one optixLaunch(w, h, 1) of the reference, PathTracerMain.cpp:197): samplesPerPixel
samples for every pixel folded into the running average; step i uses currentFrameIdx = i, so K = 8 steps are the
1024 spp of config C2.  Prints ONE JSON line on rank 0 (contract in the task statement; extra keys: roofline,
cpu_baseline, e2e, clocks, gpu_launches, grays_per_s).

For N > 1 (torchrun, one rank per GPU) every rank joins a pt_group and the K launches of the image are ONE call of
pt_launch_multi (include/ptb200.h; csrc/multigpu.cu): scene and acceleration structure replicated, whole frames dealt to the
ranks (frame k -> rank k % N; the mean image of a frame depends only on (pixel, frame index), pathTracerPrograms.cu:721),
the K % N leftover frames cut into (frame, tile-class range) sub-shares so that every rank renders exactly K / N frames'
worth of pixels; slice j of every rendered frame goes to rank j over NCCL, rank j applies the reference's running average
to its slice in frame order, the folded slices are gathered: every rank ends with the complete image, bit-identical to K
launches on one GPU.  `--partition tiles` issues K calls of one launch each instead (every launch split N ways -- the
reference's one-optixLaunch-per-displayed-frame usage).  Total work is fixed as N grows ("strong" scaling).
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


def _load_profile_facts():
    """profiles/facts.json: per workload the DRAM bytes and warp instructions of ONE launch, read by tools/ncu_facts.py from
    the committed `ncu --set full` summaries under profiles/ (file named per entry)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "facts.json")))
    except Exception:
        return {}


PROFILE_FACTS = _load_profile_facts()


# node step of the executed traversal: (flop, bytes) -- two exact boxes in 64 B, two quantised boxes in 32 B, four in 64 B
NODE_STEP = {"binary": (FLOP_NODE, BYTES_NODE), "binary-quant": (FLOP_NODE, 32), "wide4-quant": (2 * FLOP_NODE, 64)}


def algorithmic_work(c, nee, traversal="binary"):
    flop_node, bytes_node = NODE_STEP.get(traversal, NODE_STEP["binary"])
    flops = (flop_node * c["n_node"] + FLOP_TRI * c["n_tri"] + FLOP_SAMPLE * c["n_samples"] +
             FLOP_SEGMENT * c["n_radiance"] + sum(f * n for f, n in zip(FLOP_HIT, c["n_hit"])))
    if nee:
        flops += FLOP_NEE * (c["n_hit"][0] + c["n_hit"][1])
    return flops, bytes_node * c["n_node"] + BYTES_TRI * c["n_tri"]


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
    group = None
    torch.zeros(1, device=dev)                           # CUDA context up before anything is timed
    torch.cuda.synchronize()
    group_init_ms = None
    if world > 1:
        from acgpathtracing_b200.multirank import rank_group
        t0 = time.perf_counter()
        group = rank_group(rank, world)                  # NCCL communicator inside libptb200.so
        group_init_ms = 1e3 * (time.perf_counter() - t0)
    t_create = time.perf_counter()
    scene = group.create_scene(arr) if world > 1 else api.Scene(arr)
    scene_create_ms = 1e3 * (time.perf_counter() - t_create)          # host wall-clock of pt_scene_create: upload + trees
    if args.pipeline:
        scene.set_pipeline(args.pipeline)
    if args.tree:
        scene.set_tree(args.tree)
    info = scene.info()
    stream = torch.cuda.current_stream().cuda_stream

    acc = torch.zeros((h, w, 4), dtype=torch.float32, device=dev)
    fb = torch.zeros((h, w, 4), dtype=torch.uint8, device=dev)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    per_launch_calls = world > 1 and args.partition == "tiles"

    def params(frame, acc_ptr=None, fb_ptr=None):
        p = api.make_params(w, h, spp, md, frame=frame, nee=nee, importance=imp, scene=scene)
        p.accumulationBuffer = acc.data_ptr() if acc_ptr is None else acc_ptr
        p.frameBuffer = fb.data_ptr() if fb_ptr is None else fb_ptr
        return p

    def step(frame):
        """One launch of the hot path with device-resident buffers (N = 1); returns the number of our kernels launched."""
        p = params(frame)
        if args.emulate_ranks > 1:
            api.launch_partition(p, 0, args.emulate_ranks, stream)   # tuning aid: one rank's share of an N-rank job
        else:
            api.launch(p, stream)
        return 1

    phases = {}

    def image():
        """N > 1: the K launches of the image through pt_launch_multi; returns the kernels this rank launched and leaves the
        device time of the phases (summed over the calls) in `phases`."""
        phases.clear()
        if not per_launch_calls:
            group.launch(params(0), args.steps, stream)          # asynchronous; collect_phases() reads the timing afterwards
            return
        for i in range(args.steps):                              # a display loop: one launch per call, synchronised per frame
            group.launch(params(i), 1, stream)                   # like the reference's CUDA_SYNC_CHECK (PathTracerMain.cpp:209)
            collect_phases()

    def collect_phases():
        for key, v in group.last_timing(0).items():              # synchronises this rank's stream
            phases[key] = phases.get(key, 0) + v
        phases["calls"] = phases.get("calls", 0) + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- warm-up ------------------------------------------------------------------------------------------------
    if world > 1:                                     # at least args.warmup launches, in whole images
        for _ in range(max(1, -(-args.warmup // max(args.steps, 1)))):
            image()
    else:
        for i in range(args.warmup):
            step(i)
    barrier()

    # ---- timed region: the image's K launches, CUDA events on the launching stream (pt_launch_multi is stream-ordered),
    # N = 1: L2 flushed between launches ------------------------------------------------------------------------------------
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.25)
    launches = 0
    barrier()
    t_wall0 = time.perf_counter()
    if world == 1:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
        for i in range(args.steps):
            flush_buf.fill_(i & 0xFF)                 # evict L2 (untimed)
            ev[i][0].record()
            launches += step(i)
            ev[i][1].record()
    else:
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))]
        flush_buf.fill_(1)
        ev[0][0].record()
        image()
        ev[0][1].record()
        torch.cuda.synchronize()
        if not per_launch_calls:
            collect_phases()
        launches += phases["kernels"]
    barrier()
    t_wall = time.perf_counter() - t_wall0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    n_launch = torch.tensor([launches], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(n_launch)
    total_ms = float(total_ms.item())
    launches = int(n_launch.item())
    phases_timed = dict(phases) if world > 1 else None
    clocks = sampler.stop() if sampler else None
    checksum = float(acc[..., :3].double().mean().item())
    acc_after_timed = acc.clone()           # for the bit-for-bit comparison with the host-buffer path below

    pixel_samples = float(w) * h * spp * args.steps / max(args.emulate_ranks, 1)
    value = pixel_samples / (total_ms * 1e-3) / 1e6

    # ---- work counters for the timed frames (instrumented kernel, untimed) -> Grays/s and algorithmic flops.  Two passes:
    # the SPECIFIED work (every ray through the tree named in config.pipeline.tree; equals the oracle's counts) and the
    # EXECUTED work of the timed kernel (camera rays of tiny scenes resolved from candidate lists, quantised nodes) ----
    def counted(mode):
        scene.enable_counters(mode)
        scene.counters(reset=True)
        if world > 1:
            image()
        else:
            for i in range(args.steps):
                step(i)
        barrier()
        c = scene.counters(reset=True)
        sched = scene.sched_stats(reset=True)
        scene.enable_counters(False)
        ct = torch.tensor([c["n_samples"], c["n_radiance"], c["n_shadow"], c["n_node"], c["n_tri"], *c["n_hit"], c["n_miss"]],
                          dtype=torch.int64, device=dev)
        if world > 1:
            dist.all_reduce(ct)
        ct = [int(x) for x in ct.tolist()]
        return dict(n_samples=ct[0], n_radiance=ct[1], n_shadow=ct[2], n_node=ct[3], n_tri=ct[4], n_hit=ct[5:8], n_miss=ct[8]), sched

    counters, sched = counted(1)
    counters_exec, _ = counted(2)
    traversal_exec = scene.last_launch_info()["traversal"]      # of the mode-2 launches = of the timed ones
    flops, tbytes = algorithmic_work(counters, nee)
    flops_exec, tbytes_exec = algorithmic_work(counters_exec, nee, traversal_exec)
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
        # every rank: pt_launch_multi_host with pinned host buffers -- the image's K launches in one call, the complete image
        # downloaded on every rank
        h_acc = torch.zeros((h, w, 4), dtype=torch.float32).pin_memory()
        h_fb = torch.zeros((h, w, 4), dtype=torch.uint8).pin_memory()

        def host_image():
            up = down = 0
            calls = [(i, 1) for i in range(args.steps)] if per_launch_calls else [(0, args.steps)]
            for f0, k in calls:
                p = api.make_params(w, h, spp, md, frame=f0, nee=nee, importance=imp, scene=scene)
                p.accumulationBuffer, p.frameBuffer = h_acc.data_ptr(), h_fb.data_ptr()
                a, b = group.launch_host(p, k)
                up += a
                down += b
            return up, down

        host_image()
        barrier()
        t0 = time.perf_counter()
        up, down = host_image()
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        t_e2e = float(t_e2e.item())
        e2e = {"value": pixel_samples / t_e2e / 1e6, "unit": "Mpixel-samples/s",
               "h2d_bytes_per_step": up // args.steps, "d2h_bytes_per_step": down // args.steps,
               "ms_per_step": 1e3 * t_e2e / args.steps,
               "note": "pt_launch_multi_host on every rank; bytes moved by rank 0 divided by the image's K steps",
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
        # Render-kernel time on rank 0 and the share of the image's work it did there.  N = 1: the K launches, one kernel each.
        # N > 1: the render phase of pt_launch_multi (CUDA events inside the library); the plan gives every rank 1/N of the work.
        if world == 1:
            render_ms, n_kernels, share = sum(step_ms), args.steps, 1.0
        else:
            render_ms, n_kernels, share = phases_timed["render_ms"], max(phases_timed["kernels"] - phases_timed["calls"], 1), 1.0 / world
        kernel_ms = render_ms / n_kernels
        facts = PROFILE_FACTS.get(args.workload) if (args.pipeline != 1 and not args.spp and not args.tree and world == 1) else None
        sm_count = dinfo["sm_count"]
        sm_mhz = (clocks or {}).get("sm_mhz") or sm_max

        def rates(fl, by):
            tf = fl * share / (render_ms * 1e-3) / 1e12
            gb = by * share / (render_ms * 1e-3) / 1e9
            return {"flops_per_launch": fl / args.steps, "bytes_per_launch": by / args.steps, "tflops": tf, "fp32_frac": tf / fp32_meas,
                    "record_gbs": gb, "l2_frac": gb / l2_meas}

        spec, exe = rates(flops, tbytes), rates(flops_exec, tbytes_exec)
        # which roofline bounds the kernel: scenes staged in shared memory have no memory traffic to speak of (FP32 / issue
        # bound); scenes read from global memory are bound by the record fetches through L1/L2
        in_smem = scene.last_launch_info()["smem_bytes"] > 205 * 1024
        bound = "fp32" if in_smem else "l2"
        roofline = {
            "kernel": "render_persistent_kernel" if args.pipeline == 1 else "render_wavefront_kernel",
            "bound": bound,
            "achieved": exe["tflops"] if bound == "fp32" else exe["record_gbs"],
            "peak": fp32_meas if bound == "fp32" else l2_meas,
            "unit": "TFLOP/s" if bound == "fp32" else "GB/s",
            "frac": exe["fp32_frac"] if bound == "fp32" else exe["l2_frac"],
            "achieved_is": "EXECUTED algorithmic work of the timed kernel (counter mode 2) / its CUDA-event time",
            "peak_source": ("measured live by pt_bench_fp32_peak (dependent-free FFMA loop)" if bound == "fp32" else
                            "measured live by pt_bench_l2_bandwidth (streaming read of an L2-resident buffer)") +
                           "; MEASURED_PEAKS.json has no such figure",
            "peak_nominal_fp32": fp32_nominal,
            "kernel_ms": kernel_ms,
            "executed": dict(exe, traversal=traversal_exec),
            "specified": dict(spec, note="every ray through the tree named in config.pipeline.tree, exact nodes: the oracle's counts"),
            # dram__bytes_read.sum + dram__bytes_write.sum per launch from the committed ncu --set full capture of this workload
            "traffic": facts["dram_bytes"] if facts else None,
            "traffic_source": facts["file"] if facts else None,
            "hbm": {"achieved": (facts["dram_bytes"] if facts else w * h * 36) / (kernel_ms * 1e-3) / 1e9, "peak": peaks.get("hbm_gbs"),
                    "unit": "GB/s", "note": "measured DRAM bytes of the ncu capture / this run's kernel time" if facts else
                    "image traffic only: 16 B read + 16 B write + 4 B framebuffer per pixel per launch"},
            "per_gpu": world > 1,
        }
        if facts:
            # issue-slot roofline: warp instructions the kernel issues per launch (ncu smsp__inst_executed.sum) against the
            # schedulers' capacity, 4 per SM and clock
            peak_ginst = sm_count * 4 * sm_mhz * 1e6 / 1e9
            ach_ginst = facts["warp_inst"] / (kernel_ms * 1e-3) / 1e9
            roofline["issue"] = {"warp_inst_per_launch": facts["warp_inst"], "achieved": ach_ginst, "peak": peak_ginst,
                                 "unit": "Gwarp-inst/s", "frac": ach_ginst / peak_ginst, "sm_mhz": sm_mhz,
                                 "source": facts["file"] + " (smsp__inst_executed.sum; deterministic for a given frame)"}
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
                       "l2": "flushed between timed steps (256 MiB fill)" if world == 1 else
                             "flushed before the call; the K launches of one pt_launch_multi call write K distinct scratch frames (> L2 together)",
                       "pipeline": dict(scene.last_launch_info(), tree=scene.tree_in_use()),
                       "exchange": "none (1 GPU)" if world == 1 else (
                           ("K calls of pt_launch_multi with one launch each (every launch split N ways)" if per_launch_calls else
                            "one pt_launch_multi call: whole frames per rank + class-range sub-shares of the leftover frames") +
                           f"; slices over NCCL send/recv, ordered fold per slice, gather; rank 0 device time: render "
                           f"{phases_timed['render_ms']:.3f} ms, exchange + fold {phases_timed['exchange_fold_ms']:.3f} ms, gather {phases_timed['gather_ms']:.3f} ms"),
                       "scene_create_ms": scene_create_ms, "group_init_ms": group_init_ms,
                       "lbvh_build_ms": info["build_ms"]},
            "grays_per_s": rays / (total_ms * 1e-3) / 1e9,
            "rays_per_sample": rays / max(counters["n_samples"], 1),
            "counters": counters, "counters_executed": counters_exec, "sched_rank0": sched, "wall_ms_per_step_incl_flush": 1e3 * t_wall / args.steps,
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
