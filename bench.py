#!/usr/bin/env python
"""bench.py - the hot path on BASELINE.json's headline workload.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference --gpus N --steps K --warmup W
    python bench.py --workload config2|config4|config5 ...   (BASELINE.json's other GPU configs on the same harness)

Workload (config.workload): BASELINE.json configs[2] - the 1 M-triangle procedurally tessellated
glossy/dielectric scene at 1920x1080 - the configuration the 1/2/4/8-GPU metric and the
">= 1 Grays/s" target are quoted on.  One STEP = one full pass of the per-pixel estimator over
the frame at `--spp` samples per pixel: primary rays + G-buffer + every direct and indirect
sample + firefly clamp + resolve (renderPixel over all pixels, src/render.cpp:448-551).
The working set of a step (path / shadow queues, ~3 GB) is far larger than L2, and the
accumulators are re-zeroed every step, so no step can reuse cached results of the previous one.

metric = Mrays/s (one ray = one BVH::rayHit invocation, closest-hit or occlusion);
pixel-samples/s and time-to-spp are reported beside it.
  value           : device-resident - scene already in HBM, no host copies in the timed region
  e2e             : through the C ABI the reference would bind, with HOST buffers, nothing cached between steps: per step
                    rm_scene_upload (scene H2D) + the render (which first rebuilds the secondary-ray tree on the device) + the frame D2H
  e2e_first_frame : the first frame of a fresh process on the host clock, context creation and scene preparation (reference tree built on the device) included
Multi-GPU: samples are sharded by interleaved index (rank, world) with no data-path collective
during sampling; per step the fp32 accumulators are exchanged over NCCL inside the timed region
(rm_reduce_scatter: every rank ends up with, finalises and - end to end - downloads its own slice of
the frame into one pinned host frame the ranks share).  scaling = strong (the frame and spp are fixed as N grows).
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
    "config5": dict(name="configs[4]: 5M-triangle scene, 3840x2160, primary rays + FXAA", spp=0),
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


class SharedFrame:
    """One host frame (G-buffer + the four radiance planes) that every rank of the box maps and pins: POSIX shared memory,
    registered with the CUDA driver in each process, so that a rank's rm_resolve_slice DMAs its slice of the frame straight
    into the place rank 0 reads the whole frame from."""

    def __init__(self, npix, rank, world, tag, barrier):
        from multiprocessing import shared_memory
        import torch
        from raym0nade_b200.ctypes_defs import HITINFO_DTYPE, RADIANCE_DTYPE
        self.bytes = npix * (HITINFO_DTYPE.itemsize + 4 * RADIANCE_DTYPE.itemsize)
        name = "rm_bench_frame_%s" % tag
        if rank == 0:
            try:
                shared_memory.SharedMemory(name=name).unlink()
            except Exception:
                pass
            self.shm = shared_memory.SharedMemory(name=name, create=True, size=self.bytes)
        barrier()
        if rank != 0:
            self.shm = shared_memory.SharedMemory(name=name)
            try:                                        # rank 0 owns the segment: the attaching ranks must not unlink it at exit
                from multiprocessing import resource_tracker
                resource_tracker.unregister(self.shm._name, "shared_memory")
            except Exception:
                pass
        self.buf = np.frombuffer(self.shm.buf, np.uint8, self.bytes)
        self.rt = torch.cuda.cudart()
        err = self.rt.cudaHostRegister(self.buf.ctypes.data, self.bytes, 0)
        self.registered = int(err) == 0
        off = npix * HITINFO_DTYPE.itemsize
        self.gbuffer = self.buf[:off].view(HITINFO_DTYPE)
        self.planes = [self.buf[off + k * npix * RADIANCE_DTYPE.itemsize: off + (k + 1) * npix * RADIANCE_DTYPE.itemsize].view(RADIANCE_DTYPE) for k in range(4)]
        self.rank = rank

    def close(self, barrier):
        if self.registered:
            self.rt.cudaHostUnregister(self.buf.ctypes.data)
        self.gbuffer = self.planes = self.buf = None
        barrier()
        try:
            self.shm.close()
            if self.rank == 0:
                self.shm.unlink()
        except Exception:
            pass


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

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    scene, args = build_workload(opt.spp, opt.workload)
    npix = args.width * args.height
    stream = torch.cuda.current_stream().cuda_stream
    out_bytes = npix * (HITINFO_DTYPE.itemsize + 4 * RADIANCE_DTYPE.itemsize)

    # --- the first frame of a fresh process, timed on the host clock, everything included: context creation, the preparation of
    # the scene (rm_prepare_scene_device: the reference-topology tree built on the device; mip chains, light tables and the permuted
    # face streams on the host - the "model loading" the north star keeps there), rm_scene_upload (H2D + the secondary-ray tree
    # built on the device), the render with cold caches and first-touch allocations, and the download of the frame.  Single rank
    # only (N > 1 adds nothing to it).
    first_frame = None
    t0 = time.perf_counter()
    model = Model(scene)
    t_prepare = time.perf_counter() - t0
    if world == 1 and not opt.no_first_frame:
        g_host = np.zeros(npix, HITINFO_DTYPE)
        p_host = [np.zeros(npix, RADIANCE_DTYPE) for _ in range(4)]
        ff_spp = min(args.spp, opt.first_frame_spp)
        a_ff = args.replace(spp=ff_spp)
        t1 = time.perf_counter()
        c0 = Context(local_rank, stream=stream)
        t2 = time.perf_counter()
        m0 = Model(scene, c0)              # rm_prepare_scene_device: the reference's tree built on the device (gpu_ref_bvh.cu), the rest on the host
        t2b = time.perf_counter()
        c0.upload(m0)
        c0.synchronize()
        t3 = time.perf_counter()
        c0.render_into(a_ff, 1, g_host, p_host)
        t4 = time.perf_counter()
        tree = c0.tree_info()
        c0.close()
        m0.close()
        first_frame = {"spp": ff_spp, "total_s": t4 - t1, "context_create_s": t2 - t1, "prepare_scene_device_tree_s": t2b - t2,
                       "scene_upload_s": t3 - t2b, "render_incl_secondary_tree_build_and_download_s": t4 - t3, "secondary_tree": tree,
                       "prepare_scene_host_s": t_prepare,
                       "note": "fresh context, nothing cached: rm_context_create + rm_prepare_scene_device (reference tree built on the device; mips, lights, "
                               "permuted face streams on the host) + rm_scene_upload + rm_render (its first call builds the secondary-ray tree on the device: sweep SAH) into pageable "
                               "host arrays, at %d spp; prepare_scene_host_s = the all-host rm_prepare_scene of the same scene, for comparison (not in "
                               "total_s)" % ff_spp}
        del g_host, p_host

    ctx = Context(local_rank, stream=stream)
    if opt.tree_builder >= 0:
        ctx.set_option("tree_builder", opt.tree_builder)
    ctx.upload(model)
    ctx.set_option("exact_secondary", 1 if opt.exact_secondary else 0)

    # the frame reduction runs inside the library (NCCL on the context's stream); torch.distributed only carries the
    # 128-byte unique id to the ranks and provides the barriers.  Default: rm_reduce_scatter - every rank ends up with the
    # summed accumulators of its own slice of the frame, finalises it and (end to end) downloads it over its own PCIe link.
    # --reduce root / torch: the round-1 paths (ncclReduce to rank 0, or torch.distributed collectives on the library's pointers).
    if world > 1 and opt.reduce in ("scatter", "root"):
        uid = [Context.comm_unique_id().tobytes() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        ctx.comm_init(np.frombuffer(uid[0], np.uint8), rank, world)

    def reduce_and_resolve(frame):
        """exchange + resolve; `frame` = SharedFrame / (gbuffer, planes) host arrays to download into, or None"""
        g, pl = (frame.gbuffer, frame.planes) if frame is not None else (None, (None, None, None, None))
        if world == 1 or opt.reduce == "scatter":
            if world > 1:
                ctx.reduce_scatter()
            ctx.resolve_slice(args, g, pl)
            return
        if opt.reduce == "root":
            ctx.reduce(0)
        else:
            from raym0nade_b200 import multi_gpu
            multi_gpu.reduce_frame(multi_gpu.ContextAccum(ctx), dist, rank, world,
                                   lambda buf: torch.as_tensor(multi_gpu.DevPtr(buf), device=dev))
        if rank == 0:
            ctx.resolve(args, download=False)
            if frame is not None:
                ctx.download_resolved(g, pl)

    def step(seed, frame=None):
        ctx.trace_primary(args, download=False)
        ctx.gbuffer(args, download=False)
        ctx.render_samples(args, sample_begin=rank, sample_stride=world, seed=seed, reset=True)
        reduce_and_resolve(frame)

    for i in range(opt.warmup):
        step(100 + i)
    # ("tree_builder" 2 only: the device-resident figure is for a scene that has been on the device for a while - let the background
    # refinement of the secondary-ray tree finish so that the counted and the timed steps traverse the same tree)
    ctx.set_option("tree_wait", 1)
    barrier()

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
    step(99)                                    # one more untimed step with the plain kernels before the clock starts
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
    tree_info = ctx.tree_info()                 # the tree the timed steps traversed

    # --- end to end through the C ABI with HOST buffers: every step stages the scene again (rm_scene_upload: H2D of the
    # prepared scene; the secondary-ray tree is rebuilt on the device by the render that follows - nothing is cached between steps), renders, and brings
    # the frame to host memory.  N > 1: the frame lands in one pinned shared-memory frame every rank maps; each rank writes
    # its slice of it (rm_reduce_scatter + rm_resolve_slice), rank 0 owns the whole frame after the closing barrier.
    frame = SharedFrame(npix, rank, world, os.environ.get("MASTER_PORT", str(os.getpid())), barrier)
    e2e_steps = max(1, min(opt.steps, 3))

    ctx.set_option("tree_cache", 0)             # every upload builds its trees anew (nothing keyed by geometry hash survives an upload)
    if not opt.pageable_scene:
        model.pin()                             # the prepared scene is page-locked once (outside the timed loop): each step's upload is a DMA out of it

    def e2e_step(seed):
        ctx.upload(model)
        step(seed, frame)

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
    frame_ok = bool(np.isfinite(frame.planes[2]["radiance"]).all()) if rank == 0 else True
    frame_pinned = frame.registered
    # the upload alone (scene H2D + device tree build), host clock
    t0 = time.perf_counter()
    ctx.upload(model)
    ctx.synchronize()
    upload_s = time.perf_counter() - t0
    frame.close(barrier)

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
        # DRAM / L2 bytes per launch of the dominant kernel from the committed `ncu --set full` capture (profiles/traffic.json),
        # per launch like `achieved`: the capture's bytes per ray x the rays of this run's average launch
        traffic, traffic_note, l2 = None, None, None
        rays_per_launch = counted[dom]["rays"] / max(kinds[dom]["launches"] / opt.steps, 1)
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get(dom, {})
            if "dram_bytes_per_ray" in tj:
                traffic = tj["dram_bytes_per_ray"] * rays_per_launch
                traffic_note = "ncu: %.1f DRAM bytes per ray on a %d-ray launch (profiles/traffic.json) x %.0f rays per launch here" % (
                    tj["dram_bytes_per_ray"], tj["rays_per_launch"], rays_per_launch)
            else:
                traffic = tj.get("dram_bytes_per_launch")
            if "l2_bytes_per_ray" in tj:
                # what actually feeds this kernel: the scene is L2-resident, so the honest ceiling is the L2's, not HBM's
                l2_gbs = tj["l2_bytes_per_ray"] * per_kind[dom]["mrays_per_s_kernel_only"] * 1e6 / 1e9
                l2 = {"bytes_per_ray": tj["l2_bytes_per_ray"], "achieved_gbs": l2_gbs, "hit_rate": tj.get("l2_hit_rate"),
                      "pct_of_peak_under_ncu": tj.get("l2_pct_of_peak"), "l1_pct_of_peak_under_ncu": tj.get("l1_pct_of_peak"),
                      "issue_slots_busy_pct_under_ncu": tj.get("issue_pct"), "threads_per_instruction_under_ncu": tj.get("thr_per_inst"),
                      "note": "lts__t_bytes.sum per ray from the same ncu capture x this run's kernel-only ray rate; the kernel is bound by instruction "
                              "issue (see issue_slots_busy_pct / threads_per_instruction), neither by L2 nor by DRAM"}
        except Exception:
            pass
        sec_tree = "reference tree, reference visit order" if opt.exact_secondary else (
            "4-wide secondary-ray tree over the same triangles (%s, %d records of 64 B, %d levels; 8-bit quantised child boxes, conservative slab test, "
            "the reference's triangle test)" % ({"sweep_sah": "built on the device at every upload: top-down sweep SAH (gpu_sah_bvh.cu) + collapse",
                                                 "ploc+host refinement": "built on the device (Morton sort + PLOC + collapse), then replaced by the host binned-SAH tree built in the background",
                                                 "ploc": "built on the device: Morton sort + PLOC + collapse", "host": "host binned SAH + collapse"}[tree_info["builder"]],
                                                tree_info["nodes"], tree_info["levels"]))
        roofline = {"bound": "hbm", "kernel": {"primary": "k_trace<PrimaryJob>", "paths": "k_trace<PathJob>", "shadow": "k_trace<ShadowJob>"}[dom],
                    "achieved": per_kind[dom]["achieved_gbs"], "peak": peak, "unit": "GB/s", "frac": per_kind[dom]["achieved_gbs"] / peak,
                    "traffic": traffic, "traffic_note": traffic_note, "l2": l2, "rays_per_launch": rays_per_launch, "peak_source": peak_src,
                    "algorithmic_bytes_per_launch": (32.0 * counted[dom]["box"] + 36.0 * counted[dom]["tri"]) / max(kinds[dom]["launches"] / opt.steps, 1),
                    "tree": "reference tree, reference visit order" if dom == "primary" else sec_tree,
                    "note": "achieved = (32 B x child-box tests + 36 B x triangle tests) this kernel actually performed (SURVEY.md 8(d)'s per-test bytes on the "
                            "kernel's own, smaller test counts; the same rays in the reference's order are under kernels.*.reference_order) / its CUDA-event time; "
                            "the scene (%.0f MB in HBM) is largely L2-resident on B200 (126 MB L2): `traffic` and `l2` say what moved where"
                            % (ctx.scene_bytes() / 1e6)}
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
        reduce_name = {"scatter": "rm_reduce_scatter: ncclAllReduce of the clamp side data + ncclReduceScatter of the fp32 accumulators, every rank resolves its slice",
                       "root": "rm_reduce: ncclReduce to rank 0", "torch": "torch.distributed"}[opt.reduce]
        line = {"metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": opt.steps, "warmup": opt.warmup,
                "ms_per_step": ms_total / opt.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "spp": args.spp, "spp_direct": spp_d, "P_Direct": args.P_Direct,
                           "parallelism": "samples interleaved over %d GPU(s)%s" % (world, "" if world == 1 else ", per step: " + reduce_name),
                           "secondary_rays": sec_tree,
                           "l2_policy": "per-step working set (queues + accumulators, >3 GB) exceeds L2; accumulators re-zeroed each step"},
                "pixel_samples_per_s": npix * args.spp * opt.steps / (ms_total * 1e-3),
                "time_to_spp_s": {str(args.spp): ms_total / opt.steps * 1e-3},
                "rays_per_step": rays_total / opt.steps,
                "e2e": {"value": e2e_value, "unit": "Mrays/s", "h2d_bytes_per_step": int(ctx.scene_h2d_bytes()),
                        "d2h_bytes_per_step": int(out_bytes) if world == 1 else int(out_bytes / world),
                        "d2h_bytes_per_step_all_ranks": int(out_bytes),
                        "ms_per_step": float(e2e_s.item()) * 1e3 / e2e_steps, "steps": e2e_steps,
                        "time_to_spp_s": {str(args.spp): float(e2e_s.item()) / e2e_steps},
                        "scene_upload_ms": upload_s * 1e3, "frame_finite": frame_ok, "frame_pinned": frame_pinned,
                        "scene_pinned": not opt.pageable_scene,
                        "note": "per step and per rank: rm_scene_upload (prepared scene from " + ("pageable" if opt.pageable_scene else "page-locked") + " host memory; the secondary-ray tree is rebuilt on the device by the render that follows, "
                                "nothing cached) + primary + G-buffer + the rank's sample shard + exchange + resolve + download into a pinned host frame"
                                + ("" if world == 1 else " shared by the ranks (each DMAs its 1/%d slice)" % world)},
                "e2e_first_frame": first_frame,
                "gpu_launches": int(st["launches"]),
                "clocks": clk, "roofline": roofline, "kernels": per_kind, "cpu_baseline": cpu}
        print(json.dumps(line))
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


def run_config5(opt):
    """BASELINE.json configs[4]: primary rays (bit-exact {tri_idx, t}) over the 5 M-triangle scene at 3840x2160 + FXAA of the shaded
    frame.  One GPU.  value = primary Mrays/s device-resident; the scene (240 MB of triangle records + 34 MB of nodes) does not
    fit the 126 MB L2, so this is the one traversal whose HBM roofline is physical."""
    import torch
    from raym0nade_b200 import scenes
    from raym0nade_b200.api import Context, Model
    from oracle import refbind
    torch.cuda.set_device(0)
    stream = torch.cuda.current_stream().cuda_stream
    scene, args = scenes.five_million(5_000_000, 3840, 2160)
    npix = args.width * args.height
    model = Model(scene)
    ctx = Context(0, stream=stream).upload(model)
    ctx.set_option("count_tests", 1)
    ctx.stats_reset()
    ctx.trace_primary(args, download=False)
    counted = ctx.stats_kernels()["primary"]
    ctx.set_option("count_tests", 0)
    for _ in range(opt.warmup):
        ctx.trace_primary(args, download=False)
        ctx.gbuffer(args, download=False)
    torch.cuda.synchronize()
    ctx.set_option("time_kernels", 1)
    ctx.stats_reset()
    clocks = ClockSampler(0)
    clocks.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(opt.steps):
        ctx.trace_primary(args, download=False)
        ctx.gbuffer(args, download=False)
    e1.record()
    torch.cuda.synchronize()
    clk = clocks.result()
    ms = e0.elapsed_time(e1)
    kind = ctx.stats_kernels()["primary"]
    st = ctx.stats()
    ctx.set_option("time_kernels", 0)
    value = npix * opt.steps / (ms * 1e-3) / 1e6
    # FXAA alone on the shaded frame of this scene (base colour x exposure through Photo::shade + gamma), device buffers
    a0 = args.replace(spp=0)
    ctx.render_samples(a0, seed=1)
    ctx.resolve(a0, download=False)
    shaded = ctx.postprocess(a0, refbind.SHADE["BaseColor"])
    frames = [torch.from_numpy(np.roll(shaded, 17 * k, axis=1).copy()).to("cuda") for k in range(6)]        # 6 x 100 MB > L2
    out = torch.empty_like(frames[0])
    evs = []
    for k in range(3 + 24):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        ctx.fxaa_device(frames[k % 6].data_ptr(), out.data_ptr(), args.width, args.height)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    fx = sorted(a.elapsed_time(b) for a, b in evs[3:])
    fx_ms = fx[len(fx) // 2]
    # end to end: scene upload + primary hits to the host + G-buffer + shade + gamma + FXAA to the host
    tri = torch.empty(npix, dtype=torch.int32).pin_memory().numpy()
    tt = torch.empty(npix, dtype=torch.float32).pin_memory().numpy()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(opt.steps, 3))
    for _ in range(e2e_steps):
        ctx.upload(model)
        ctx.trace_primary_into(args, tri, tt)
        ctx.render_samples(a0, seed=1)
        ctx.resolve(a0, download=False)
        rgb = ctx.postprocess(a0, refbind.SHADE["BaseColor"] | refbind.SHADE["DoFXAA"])
    e2e_s = (time.perf_counter() - t0) / e2e_steps
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_per_launch = 32.0 * counted["box"] + 36.0 * counted["tri"]
    k_ms = kind["ms"] / max(kind["launches"], 1)
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json"))).get("primary_config5", {}).get("dram_bytes_per_launch")
    except Exception:
        pass
    cpu = None
    if not opt.no_cpu:
        try:
            threads = os.cpu_count() or 1
            with StdoutToStderr():
                R = refbind.RefScene(scene)
                t1 = time.perf_counter()
                R.trace_primary(args.replace(width=960, height=540), threads=threads)
                sec = time.perf_counter() - t1
                R.close()
            cpu = {"value": 960 * 540 / sec / 1e6, "unit": "Mrays/s", "cores": threads, "kind": "reference", "sample": "primary rays of the same scene and camera at 960x540 (%.1f s)" % sec}
        except Exception as e:
            cpu = {"value": None, "unit": "Mrays/s", "cores": 0, "kind": "reference", "sample": "unavailable: %s" % e}
    print(json.dumps({
        "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": 1, "steps": opt.steps, "warmup": opt.warmup, "ms_per_step": ms / opt.steps,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "configs[4]: 5M-triangle scene (%d tris), 3840x2160, primary rays + G-buffer per step; FXAA timed alone" % scene.n_faces,
                   "l2_policy": "the traversal streams (%.0f MB) exceed the 126 MB L2; FXAA rotates 6 distinct 100 MB inputs" % ((scene.n_faces * 48 + model.desc.n_nodes * 32) / 1e6)},
        "e2e": {"value": npix / e2e_s / 1e6, "unit": "Mrays/s", "h2d_bytes_per_step": int(ctx.scene_h2d_bytes()), "d2h_bytes_per_step": int(npix * 8 + npix * 12),
                "ms_per_step": e2e_s * 1e3, "note": "rm_scene_upload + rm_trace_primary to pinned host arrays + G-buffer + shade + gamma + FXAA to the host"},
        "gpu_launches": int(st["launches"]), "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "k_trace<PrimaryJob>", "achieved": bytes_per_launch / (k_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                     "frac": bytes_per_launch / (k_ms * 1e-3) / 1e9 / peak, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_per_launch,
                     "box_tests_per_ray": counted["box"] / counted["rays"], "tri_tests_per_ray": counted["tri"] / counted["rays"],
                     "tree": "reference tree, reference visit order (bit-exact tri_idx / t)"},
        "kernels": {"primary": {"ms_per_launch": k_ms, "mrays_per_s_kernel_only": counted["rays"] / (k_ms * 1e-3) / 1e6},
                    "fxaa": {"us_per_launch": fx_ms * 1e3, "algorithmic_bytes": 24.0 * npix, "achieved_gbs": 24.0 * npix / (fx_ms * 1e-3) / 1e9,
                             "frac": 24.0 * npix / (fx_ms * 1e-3) / 1e9 / peak, "roofline_us": 24.0 * npix / peak / 1e3}},
        "cpu_baseline": cpu}))
    ctx.close()


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
    ap.add_argument("--reduce", default="scatter", choices=["scatter", "root", "torch"],
                    help="N > 1: rm_reduce_scatter (default: every rank resolves and downloads its slice), rm_reduce to rank 0, or torch.distributed collectives")
    ap.add_argument("--tree-builder", type=int, default=-1, help="secondary-ray tree: 0 host binned SAH, 1 device PLOC, 2 device PLOC + background refinement by the host builder, 3 device sweep SAH (library default)")
    ap.add_argument("--no-first-frame", action="store_true", help="skip the e2e_first_frame measurement")
    ap.add_argument("--pageable-scene", action="store_true", help="end-to-end loop: upload the scene from pageable host memory (round-2 behaviour) instead of page-locking it once")
    ap.add_argument("--first-frame-spp", type=int, default=1 << 30, help="spp of the first-frame measurement (default: the step's)")
    opt = ap.parse_args()
    global WORKLOAD
    WORKLOAD = WORKLOADS[opt.workload]["name"]
    if opt.spp <= 0:
        opt.spp = WORKLOADS[opt.workload]["spp"]
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if opt.warmup < 3 and opt.impl == "ours":
        opt.warmup = 3
    if opt.workload == "config5" and opt.impl == "ours":
        if rank == 0:
            run_config5(opt)
        return
    if opt.impl == "reference":
        run_reference(opt, rank, world)
    else:
        run_ours(opt, rank, world, local_rank)


if __name__ == "__main__":
    main()
