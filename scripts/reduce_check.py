"""Multi-GPU parity of the in-library exchange (rm_comm_init + rm_reduce), run under torchrun with one rank per GPU:
every rank renders its interleaved sample shard, rm_reduce combines them on rank 0, and rank 0 compares the resolved
frame with the frame it renders alone (same seed => same sample set; only fp32 summation order differs).
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 scripts/reduce_check.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from raym0nade_b200 import scenes
from raym0nade_b200.api import Context, Model

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("gloo")                      # only carries the unique id and the verdict: the data path is rm_reduce
scene, args = scenes.cornell_box(128, 128, 48)
ctx = Context(local)
ctx.set_option("tree_builder", 1)           # the same tree on every rank and in the single-GPU render it is compared with
ctx.upload(Model(scene))
uid = [Context.comm_unique_id().tobytes() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
ctx.comm_init(np.frombuffer(uid[0], np.uint8), rank, world)
ok = True
for seed in (3, 4):                                  # twice: the communicator and the accumulators are reused
    ctx.trace_primary(args, download=False); ctx.gbuffer(args, download=False)
    ctx.render_samples(args, rank, world, seed=seed, reset=True)
    ctx.reduce(0)
    if rank == 0:
        combined = ctx.resolve(args)
        ctx.render_samples(args, 0, 1, seed=seed, reset=True)
        alone = ctx.resolve(args)
        alone["gbuffer"] = ctx.resolved(args)["gbuffer"]
        for k in ("Dd", "Ds", "Id", "Is"):
            a, b = alone[k]["radiance"].astype(np.float64), combined[k]["radiance"].astype(np.float64)
            good = np.isfinite(b).all() and np.abs(a - b).max() <= 2e-4 * (1.0 + np.abs(a).max()) and abs(a.sum() - b.sum()) <= 1e-5 * abs(a.sum()) + 1e-6
            if not good: print("MISMATCH", k, np.abs(a - b).max(), a.sum(), b.sum())
            ok = ok and bool(good)
    # the same frame through rm_reduce_scatter: every rank finalises its own slice and writes it into a whole-frame array;
    # the slices of all ranks, gathered, must be the frame rank 0 renders alone
    ctx.render_samples(args, rank, world, seed=seed, reset=True)
    ctx.reduce_scatter()
    first, count = ctx.frame_slice()
    npix = args.width * args.height
    from raym0nade_b200.ctypes_defs import HITINFO_DTYPE, RADIANCE_DTYPE
    g = np.zeros(npix, HITINFO_DTYPE)
    planes = [np.zeros(npix, RADIANCE_DTYPE) for _ in range(4)]
    ctx.resolve_slice(args, g, planes)
    mine = [(first, count, [p[first:first + count].copy() for p in planes], g[first:first + count].copy())]
    gathered = [None] * world
    dist.all_gather_object(gathered, mine[0])
    if rank == 0:
        covered = np.zeros(npix, np.int32)
        for k, name in enumerate(("Dd", "Ds", "Id", "Is")):
            whole = np.zeros(npix, RADIANCE_DTYPE)
            for (f, c, pl, gg) in gathered:
                whole[f:f + c] = pl[k]
                if k == 0: covered[f:f + c] += 1
            a, b = alone[name]["radiance"].astype(np.float64), whole["radiance"].astype(np.float64)
            good = np.isfinite(b).all() and np.abs(a - b).max() <= 2e-4 * (1.0 + np.abs(a).max()) and abs(a.sum() - b.sum()) <= 1e-5 * abs(a.sum()) + 1e-6
            if not good: print("SCATTER MISMATCH", name, np.abs(a - b).max(), a.sum(), b.sum())
            ok = ok and bool(good)
        ok = ok and bool((covered == 1).all())
        gw = np.zeros(npix, HITINFO_DTYPE)
        for (f, c, pl, gg) in gathered: gw[f:f + c] = gg
        same_g = np.array_equal(gw["baseColor"], alone["gbuffer"]["baseColor"], equal_nan=True) and np.array_equal(gw["position"], alone["gbuffer"]["position"], equal_nan=True)
        if not same_g: print("SCATTER MISMATCH gbuffer")
        ok = ok and bool(same_g)
    ctx.synchronize()
verdict = [ok]
dist.broadcast_object_list(verdict, src=0)
ctx.close()
dist.destroy_process_group()
if rank == 0: print("rm_reduce over %d ranks: %s" % (world, "OK" if verdict[0] else "FAILED"))
sys.exit(0 if verdict[0] else 1)
