#!/usr/bin/env python
"""bench.py - the hot path on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W

Workload (config.workload): BASELINE.json configs[2] - the 1 M-triangle procedurally tessellated
glossy/dielectric scene at 1920x1080 - the configuration the 1/2/4/8-GPU metric and the
">= 1 Grays/s" target are quoted on.  One STEP = one full pass of the per-pixel estimator over
the frame at `--spp` samples per pixel: primary rays + G-buffer + every direct and indirect
sample + firefly clamp + resolve (renderPixel over all pixels, src/render.cpp:448-551).
The working set of a step (path / shadow queues, ~3 GB) is far larger than L2, and the
accumulators are re-zeroed every step, so no step can reuse cached results of the previous one.

metric = Mrays/s (one ray = one BVH::rayHit invocation, closest-hit or occlusion);
pixel-samples/s and time-to-spp are reported beside it.
  value : device-resident - scene already in HBM, no host copies in the timed region
  e2e   : through the C ABI the reference would bind (rm_scene_upload + rm_render) with HOST
          buffers: scene H2D and G-buffer/radiance-plane D2H inside the timed region
Multi-GPU: samples are sharded by interleaved index (rank, world) with no data-path collective
during sampling; the fp32 accumulators are reduced over NCCL at the end of every step (inside
the timed region).  scaling = strong (the frame and spp are fixed as N grows)... per the
contract's vocabulary we report "strong".
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np

# The headline workload is configs[2]; --workload config2 / config4 run the same bench on BASELINE.json's other GPU configs.
WORKLOADS = {
    "config3": dict(name="configs[2]: 1M-triangle glossy/dielectric synthetic scene (990,744 tris), 1920x1080", spp=1024),
    "config2": dict(name="configs[1]: Sponza-scale synthetic mesh (257,778 tris) with a 2048x1024 HDR sky, 1024^2 albedo maps, 1920x1080", spp=256),
    "config4": dict(name="configs[3]: texture-heavy scene (491,368 tris, 32 materials x 2048^2 RGBA8 albedo + RGB8 normal maps with mips, alpha cut-outs), 3840x2160", spp=4096),
}
WORKLOAD = WORKLOADS["config3"]["name"]
CPU_SAMPLE = dict(width=960, height=540, spp=16)     # bounded sample of the same scene/camera for the CPU legs (~10-20 s of host time)


def build_workload(spp, workload="config3"):
    """Scene + args of the workload.  Under torchrun the ranks of one box share one generated scene through a pickle in
    /tmp (LOCAL_RANK 0 writes it): the texture-heavy scene takes a minute of numpy to generate."""
    import pickle
    from raym0nade_b200 import scenes
    make = {"config3": lambda: scenes.glossy_dielectric(1_000_000, 1920, 1080, spp),
            "config2": lambda: scenes.sponza_scale(260_000, 1920, 1080, spp, tex_size=1024),
            "config4": lambda: scenes.texture_heavy(500_000, 3840, 2160, spp)}[workload]
    world, local_rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
    if world == 1 or workload == "config3":
        return make()
    path = "/tmp/rm_bench_scene_%s_%d_%s.pkl" % (workload, spp, os.environ.get("MASTER_PORT", "0"))
    if local_rank == 0:
        scene, args = make()
        with open(path + ".tmp", "wb") as f:
            pickle.dump((scene, args), f, protocol=4)
        os.replace(path + ".tmp", path)
        return scene, args
    while not os.path.exists(path):
        time.sleep(0.5)
    with open(path, "rb") as f:
        return pickle.load(f)


class ClockSampler(threading.Thread):
    """SM clock / throttle reasons during the timed region (NVML, 200 ms period)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if not self.nv:
            return
        nv = self.nv
        names = {nv.nvmlClocksEventReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksEventReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 nv.nvmlClocksEventReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksEventReasonSwPowerCap: "sw_power_cap"} \
            if hasattr(nv, "nvmlClocksEventReasonHwSlowdown") else \
            {nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown", nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
             nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown", nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap"}
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
                r = get(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.2)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


class StdoutToStderr:
    """The reference's C++ code prints progress on stdout (std::cout in Model / BVH); stdout carries the JSON line only."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)
        return False


def cpu_reference_run(scene, args, threads):
    """The reference's own renderPixel loop (oracle/_ref) on a bounded sample of the workload."""
    from oracle import refbind
    a = args.replace(**CPU_SAMPLE)
    with StdoutToStderr():
        R = refbind.RefScene(scene)
        out = R.render(a, threads=threads, seed_base=0)
        R.close()
    return out, a


def run_reference(opt, rank, world):
    if rank != 0:
        return
    from oracle import refbind
    base = {"impl": "reference", "metric": "Mrays/s", "unit": "Mrays/s", "n_gpus": opt.gpus, "steps": opt.steps, "warmup": opt.warmup,
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic"}
    if not refbind.available("plain"):
        print(json.dumps(dict(base, unavailable="oracle/_ref not built (needs the reference sources: make -C oracle ref)")))
        return
    threads = os.cpu_count() or 1
    scene, args = build_workload(opt.spp, opt.workload)
    a = args.replace(**CPU_SAMPLE)
    rays, secs = 0, 0.0
    with StdoutToStderr():
        R = refbind.RefScene(scene)
        for _ in range(opt.warmup):
            R.render(a, threads=threads, seed_base=0)
        for i in range(opt.steps):
            o = R.render(a, threads=threads, seed_base=1000 * (i + 1))
            rays += o["rays"]
            secs += o["seconds"]
    value = rays / secs / 1e6
    sample = "same scene and camera at %dx%d, %d spp per step (cost is linear in pixels x spp)" % (a.width, a.height, a.spp)
    print(json.dumps(dict(base, value=value, ms_per_step=1e3 * secs / opt.steps,
                          pixel_samples_per_s=a.width * a.height * a.spp * opt.steps / secs,
                          config={"workload": WORKLOAD, "spp": opt.spp, "timed_sample": sample, "threads": threads},
                          cpu_baseline={"value": value, "unit": "Mrays/s", "cores": threads, "kind": "reference", "sample": sample},
                          e2e={"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0})))


def run_ours(opt, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from raym0nade_b200.api import Context, Model
    from raym0nade_b200.ctypes_defs import HITINFO_DTYPE, RADIANCE_DTYPE

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL may print its version banner on stdout when the communicator is created; keep stdout for the JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    scene, args = build_workload(opt.spp, opt.workload)
    npix = args.width * args.height
    model = Model(scene)
    stream = torch.cuda.current_stream().cuda_stream
    ctx = Context(local_rank, stream=stream).upload(model)
    ctx.set_option("exact_secondary", 1 if opt.exact_secondary else 0)

    # the frame reduction runs inside the library (rm_reduce: NCCL on the context's stream); torch.distributed only
    # carries the 128-byte unique id to the ranks and provides the barriers.  --reduce torch drives the same three steps
    # through torch.distributed collectives on the library's device pointers instead.
    if world > 1 and opt.reduce == "cabi":
        uid = [Context.comm_unique_id().tobytes() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(np.frombuffer(uid[0], np.uint8), rank, world)

    def reduce_across_ranks():
        if opt.reduce == "cabi":
            ctx.reduce(0)
            return
        from raym0nade_b200 import multi_gpu
        multi_gpu.reduce_frame(multi_gpu.ContextAccum(ctx), dist, rank, world,
                               lambda buf: torch.as_tensor(multi_gpu.DevPtr(buf), device=dev))

    def step(seed):
        ctx.trace_primary(args, download=False)
        ctx.gbuffer(args, download=False)
        ctx.render_samples(args, sample_begin=rank, sample_stride=world, seed=seed, reset=True)
        if world > 1:
            reduce_across_ranks()
        if rank == 0:
            ctx.resolve(args, download=False)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # --- work per step (B, T per ray and per kernel kind), untimed, counting build of the kernels
    ctx.set_option("count_tests", 1)
    ctx.stats_reset()
    step(1)
    counted = ctx.stats_kernels()
    # the same rays through the reference's own tree in the reference's order: the B and T SURVEY.md section 8d defines the
    # per-ray bytes on (bounce and shadow rays normally traverse the secondary-ray tree, which tests fewer boxes and triangles)
    ctx.set_option("exact_secondary", 1)
    ctx.stats_reset()
    step(1)
    counted_ref = ctx.stats_kernels()
    ctx.set_option("exact_secondary", 1 if opt.exact_secondary else 0)
    ctx.set_option("count_tests", 0)

    for i in range(opt.warmup):
        step(100 + i)
    ctx.set_option("time_kernels", 1)
    barrier()
    ctx.stats_reset()
    clocks = ClockSampler(local_rank)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(opt.steps):
        step(1)                     # same seed as the counted pass: identical rays, so its B/T apply exactly
    e1.record()
    barrier()
    clk = clocks.result()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    st = ctx.stats()
    kinds = ctx.stats_kernels()
    ctx.set_option("time_kernels", 0)
    rays = torch.tensor([float(st["rays"])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(rays, op=dist.ReduceOp.SUM)
    ms_total, rays_total = float(ms.item()), float(rays.item())
    value = rays_total / (ms_total * 1e-3) / 1e6

    # --- end to end through the C ABI with host buffers (scene upload + render + download)
    g_host = torch.empty(npix * HITINFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(HITINFO_DTYPE)
    p_host = [torch.empty(npix * RADIANCE_DTYPE.itemsize, dtype=torch.uint8).pin_memory().numpy().view(RADIANCE_DTYPE) for _ in range(4)]
    e2e_steps = max(1, min(opt.steps, 3))

    def e2e_step(seed):
        ctx.upload(model)
        if world == 1:
            ctx.render_into(args, seed, g_host, p_host)
        else:
            step(seed)
            if rank == 0:
                ctx.download_resolved(g_host, p_host)

    e2e_step(7)
    barrier()
    ctx.stats_reset()
    t0 = time.perf_counter()
    for i in range(e2e_steps):
        e2e_step(1)
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    e2e_rays = torch.tensor([float(ctx.stats()["rays"])], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(e2e_rays, op=dist.ReduceOp.SUM)
    e2e_value = float(e2e_rays.item()) / float(e2e_s.item()) / 1e6

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # roofline of the dominant traversal kernel: algorithmic bytes = 32 B per box test + 36 B per triangle test
        per_kind = {}
        for k in ("primary", "paths", "shadow"):
            c, t = counted[k], kinds[k]
            if t["launches"] == 0 or c["rays"] == 0:
                continue
            bytes_per_step = 32.0 * c["box"] + 36.0 * c["tri"]
            gbs = bytes_per_step * opt.steps / (t["ms"] * 1e-3) / 1e9
            per_kind[k] = {"ms_per_step": t["ms"] / opt.steps, "launches_per_step": t["launches"] / opt.steps, "rays_per_step": c["rays"],
                           "box_tests_per_ray": c["box"] / c["rays"], "tri_tests_per_ray": c["tri"] / c["rays"],
                           "bytes_per_ray": bytes_per_step / c["rays"], "achieved_gbs": gbs,
                           "mrays_per_s_kernel_only": c["rays"] * opt.steps / (t["ms"] * 1e-3) / 1e6}
            cr = counted_ref[k]
            if cr["rays"]:
                ref_bytes = 32.0 * cr["box"] + 36.0 * cr["tri"]
                per_kind[k]["reference_order"] = {"box_tests_per_ray": cr["box"] / cr["rays"], "tri_tests_per_ray": cr["tri"] / cr["rays"],
                                                  "bytes_per_ray": ref_bytes / cr["rays"],
                                                  "equivalent_gbs": ref_bytes * opt.steps / (t["ms"] * 1e-3) / 1e9}
        per_kind["shade"] = {"ms_per_step": kinds["shade"]["ms"] / opt.steps, "launches_per_step": kinds["shade"]["launches"] / opt.steps}
        dom = max((k for k in per_kind if k != "shade"), key=lambda k: per_kind[k]["ms_per_step"])
        # DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json)
        # per launch like `achieved`: the capture's DRAM bytes per ray x the rays of this run's average launch
        traffic, traffic_note = None, None
        rays_per_launch = counted[dom]["rays"] / max(kinds[dom]["launches"] / opt.steps, 1)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom, {})
            if "dram_bytes_per_ray" in tj:
                traffic = tj["dram_bytes_per_ray"] * rays_per_launch
                traffic_note = "ncu: %.1f DRAM bytes per ray on a %d-ray launch (profiles/traffic.json) x %.0f rays per launch here" % (
                    tj["dram_bytes_per_ray"], tj["rays_per_launch"], rays_per_launch)
            else:
                traffic = tj.get("dram_bytes_per_launch")
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": {"primary": "k_trace<PrimaryJob>", "paths": "k_trace<PathJob>", "shadow": "k_trace<ShadowJob>"}[dom],
                    "achieved": per_kind[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per_kind[dom]["achieved_gbs"] / peak,
                    "traffic": traffic, "traffic_note": traffic_note, "rays_per_launch": rays_per_launch, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": 32.0 * counted[dom]["box"] / max(kinds[dom]["launches"] / opt.steps, 1) + 36.0 * counted[dom]["tri"] / max(kinds[dom]["launches"] / opt.steps, 1),
                    "tree": "reference tree, reference visit order" if (opt.exact_secondary or dom == "primary") else
                            "secondary-ray tree (binned SAH over the same triangles, leaves <= 4): fewer box and triangle tests per ray than the "
                            "reference order, whose figures are under kernels.*.reference_order",
                    "note": "achieved = (32 B x box tests + 36 B x triangle tests) this kernel actually performed / its CUDA-event time; "
                            "the scene (%.0f MB in HBM, traversal streams %.0f MB) is largely L2-resident on B200 (126 MB L2), see DESIGN.md"
                            % (ctx.scene_bytes() / 1e6, (scene.n_faces * 48 + model.desc.n_nodes * 32) / 1e6)}
        # reference CPU path on this box's host cores, bounded sample of the same workload
        cpu = None
        if world == 1 and not opt.no_cpu:
            try:
                threads = os.cpu_count() or 1
                o, a = cpu_reference_run(scene, args, threads)
                cpu = {"value": o["rays"] / o["seconds"] / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "reference",
                       "sample": "same scene and camera at %dx%d, %d spp (%.1f s)" % (a.width, a.height, a.spp, o["seconds"]),
                       "pixel_samples_per_s": a.width * a.height * a.spp / o["seconds"]}
            except Exception as e:                                   # oracle missing on this box
                cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
        spp_d = int(np.float32(args.spp) * np.float32(args.P_Direct))
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": opt.steps, "warmup": opt.warmup,
                "ms_per_step": ms_total / opt.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "spp": args.spp, "spp_direct": spp_d, "P_Direct": args.P_Direct,
                           "parallelism": "samples interleaved over %d GPU(s), NCCL reduce of fp32 accumulators per step (%s)" % (world, "rm_reduce" if opt.reduce == "cabi" else "torch.distributed"),
                           "secondary_rays": "reference tree and order" if opt.exact_secondary else "secondary-ray tree (same triangles, same tests, binned-SAH topology)",
                           "l2_policy": "per-step working set (queues + accumulators, >3 GB) exceeds L2; accumulators re-zeroed each step"},
                "pixel_samples_per_s": npix * args.spp * opt.steps / (ms_total * 1e-3),
                "time_to_spp_s": {str(args.spp): ms_total / opt.steps * 1e-3},
                "rays_per_step": rays_total / opt.steps,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(ctx.scene_h2d_bytes()),
                        "d2h_bytes_per_step": int(npix * (HITINFO_DTYPE.itemsize + 4 * RADIANCE_DTYPE.itemsize)),
                        "ms_per_step": float(e2e_s.item()) * 1e3 / e2e_steps, "steps": e2e_steps,
                        "note": "rm_scene_upload + rm_render per step; outputs land in pinned host memory, the prepared scene is pageable"},
                "gpu_launches": int(st["launches"]),
                "clocks": clk, "roofline": roofline, "kernels": per_kind, "cpu_baseline": cpu}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="config3", choices=sorted(WORKLOADS))
    ap.add_argument("--spp", type=int, default=int(os.environ.get("RM_BENCH_SPP", "0")), help="samples per pixel of one step (default: the config's)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--exact-secondary", action="store_true", help="bounce and shadow rays through the reference's own tree in the reference's order (default: the secondary-ray tree)")
    ap.add_argument("--reduce", default="cabi", choices=["cabi", "torch"], help="N > 1: rm_reduce (NCCL inside the library) or torch.distributed collectives")
    opt = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[opt.workload]["name"]
    if opt.spp <= 0:
        opt.spp = WORKLOADS[opt.workload]["spp"]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if opt.warmup < 3 and opt.impl == "ours":
        opt.warmup = 3
    if opt.impl == "reference":
        run_reference(opt, rank, world)
    else:
        run_ours(opt, rank, world, local_rank)


if __name__ == "__main__":
    main()
