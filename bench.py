#!/usr/bin/env python
"""bench.py -- body-steps/sec of the pdphyzx world step on B200 (BASELINE.json metric).

A "step" is one pxBatchStep(b, FUSED) over every world of the batch: FUSED substeps of
pxWorldStepWithDt (reference src/px_world.c:410-420) fused into one kernel launch; FUSED
defaults to 5, one fired px->world->step frame (px_world.c:427-437).  The workload is the
north_star's target scene: WORLDS worlds per GPU (weak scaling: 8192/GPU = 65,536 on 8 GPUs)
of 256 mixed-shape bodies (252 dynamic: 1/3 circles, 1/3 boxes, 1/3 3-8-gons; 4 static),
10 solver iterations, PDPHYZX_UNIT_M, penned piles pre-settled on the device before timing.

value   = body-substeps/s, device-timed with CUDA events on the launching stream, state
          resident in HBM, whole job (all ranks), max over ranks.
e2e     = same metric through the C ABI with HOST buffers (pxBatchStepHost): every step copies
          the mutable blocks pinned-host -> HBM, steps, copies them back and waits.
--impl reference = the reference's own CPU step (oracle/_ref, the unmodified sources compiled
          under the PlaydateAPI host stub) on all host threads, bounded sample of the workload.

One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

BODIES_PER_WORLD = 256
N_DYNAMIC = 252
# algorithmic HBM bytes per body-substep at K=1: mutable state read+written
# (44 B x 2) + immutable (28 B) + shape (circle 0, box 64, polygon 16 B/vertex avg 5.5 verts)
ALG_BYTES = {"circles": 116.0, "mixed": (116.0 + 180.0 + (116.0 + 16 * 5.5)) / 3.0}


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--worlds", type=int, default=8192, help="worlds per GPU")
    ap.add_argument("--mix", default="mixed", choices=["mixed", "circles"])
    ap.add_argument("--iterations", type=int, default=10)
    ap.add_argument("--fused", type=int, default=5, help="substeps per launch (K)")
    ap.add_argument("--settle", type=int, default=1000, help="untimed substeps to let the piles form")
    ap.add_argument("--threads", type=int, default=0, help="threads per world-CTA (0 = library default)")
    ap.add_argument("--seed", type=int, default=1)
    ap.add_argument("--max-manifolds", type=int, default=0, help="per-world manifold capacity (0 = library default)")
    ap.add_argument("--max-pairs", type=int, default=0, help="per-world AABB-pass pair capacity (0 = library default)")
    ap.add_argument("--cpu-worlds", type=int, default=0, help="worlds in the CPU baseline sample (0 = 2 per core)")
    ap.add_argument("--e2e-chunks", type=int, default=5, help="world ranges pipelined per host-buffer step")
    ap.add_argument("--control", action="store_true",
                    help="BASELINE config 5: polygon stacks, PDPHYZX_UNIT_MM, every step the host queues one applyImpulse per "
                         "world and a setGravity before stepping --fused substeps (compare --fused 1 with --fused 8)")
    ap.add_argument("--config", type=int, default=0, choices=[0, 2, 3, 4, 5],
                    help="BASELINE.json configs as stated: 2 = 4,096 circle-pile worlds; 3 = 16,384 mixed worlds, 10 it.; "
                         "4 = 65,536 mixed worlds IN TOTAL, 20 it., sharded over --gpus (strong scaling); 5 = 32,768 polygon-stack "
                         "worlds, UNIT_MM, host control every step (use --fused 1 and --fused 8); 0 = the default line")
    ap.add_argument("--sharded", action="store_true",
                    help="single process: one PxBatchGroup (C ABI) over --gpus devices instead of one process per GPU")
    ap.add_argument("--no-aos", action="store_true", help="skip the e2e_aos leg (PxWorld structs in, PxWorld structs out)")
    ap.add_argument("--timed-iterations", type=int, default=-1,
                    help="experiment: switch every world to this many solver iterations after settling (0 = time phases A-D+F alone)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--profile", action="store_true", help="per-phase SM-cycle breakdown (adds clock reads)")
    args = ap.parse_args()
    args.scaling = "weak"
    if args.config == 2:
        args.worlds, args.mix, args.iterations = 4096, "circles", 10
    elif args.config == 3:
        args.worlds, args.mix, args.iterations = 16384, "mixed", 10
    elif args.config == 4:
        world_size = int(os.environ.get("WORLD_SIZE", "1")) if not args.sharded else 1
        n = world_size if not args.sharded else 1
        args.total_worlds = 65536
        args.worlds = 65536 // max(n, 1) if not args.sharded else 65536
        args.mix, args.iterations, args.scaling = "mixed", 20, "strong"
    elif args.config == 5:
        args.worlds, args.control = 32768, True
        # impulse-driven polygon stacks pile deeper than the settled bench scene: capacities above the library defaults,
        # so that no world drops a manifold or a pair (worlds_overflowed must read 0)
        args.max_manifolds = args.max_manifolds or 1024
        args.max_pairs = args.max_pairs or 2304
    return args


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.lines = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.index)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, smax, reasons = [], [], set()
        for ln in self.lines:
            p = [x.strip() for x in ln.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1]))
                smax.append(float(p[2]))
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(smax) if smax else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_facts(args):
    """profiles/*_kernel_facts.json written by profiles/summarize.py from an ncu capture of the
    default command line; only used when this run IS that configuration."""
    if (args.worlds, args.mix, args.iterations, args.fused, args.threads) != (8192, "mixed", 10, 5, 0):
        return None
    import glob
    files = sorted(glob.glob(os.path.join(ROOT, "profiles", "*_kernel_facts.json")))
    if not files:
        return None
    with open(files[-1]) as f:
        facts = json.load(f)
    facts["_file"] = os.path.basename(files[-1])
    return facts


def build_scenes(args, rank: int, n_worlds: int, first: int = 0):
    from pdphyzx_b200 import scenes
    base = rank * args.worlds + first
    if args.control:
        return [scenes.stack_scene(base + w, seed=args.seed, n_dynamic=N_DYNAMIC, iterations=args.iterations, cols=21)
                for w in range(n_worlds)]
    return [scenes.pile_scene(base + w, seed=args.seed, mix=args.mix, iterations=args.iterations, n_dynamic=N_DYNAMIC)
            for w in range(n_worlds)]


def bind_to_gpu_numa_node(index: int):
    """Pin this process to the CPUs NVML reports as local to GPU `index`, BEFORE the pinned staging
    buffers are allocated and first touched: with one process per GPU the host <-> device copies of
    pxBatchStepHost then stay on the GPU's own NUMA node.  Returns the CPU count, or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        ncpu = os.cpu_count() or 1
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        cpus = [i for i in range(ncpu) if (mask[i // 64] >> (i % 64)) & 1]
        if cpus and len(cpus) < ncpu:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


CHUNK = 2048  # worlds built on the host at a time when the AoS form of the whole batch would be large


def make_batch(pxb, args, rank, device, stream):
    """The batch of args.worlds worlds on `device`.  Up to 8192 worlds are built in one go and the
    host PxWorld structs are kept (e2e_aos needs them); larger batches (configs 3-5) are created
    empty and uploaded chunk by chunk (pxBatchCreate + pxBatchUpload), 340 KB of AoS per world."""
    flags = pxb.FLAG_PROFILE if args.profile else 0
    if args.worlds <= 8192:
        worlds = pxb.HostWorlds(build_scenes(args, rank, args.worlds))
        b = pxb.Batch(worlds, device=device, stream=stream, threads=args.threads, flags=flags,
                      max_manifolds=args.max_manifolds, max_pairs=args.max_pairs)
        return b, worlds
    b = None
    for first in range(0, args.worlds, CHUNK):
        n = min(CHUNK, args.worlds - first)
        hw = pxb.HostWorlds(build_scenes(args, rank, n, first))
        if b is None:
            d = pxb.make_desc(hw.unit)
            pxb._check(pxb.core().pxBatchFitDesc(hw.pointer_array(), n, pxb.C.byref(d)), "pxBatchFitDesc")
            b = pxb.Batch(None, n_worlds=args.worlds, unit=hw.unit, device=device, stream=stream, threads=args.threads,
                          flags=flags, max_dynamic=d.maxDynamic, max_static=d.maxStatic,
                          max_vertices=max(d.maxVertices + 256, N_DYNAMIC * 6), max_manifolds=args.max_manifolds,
                          max_pairs=args.max_pairs)
        b.upload(hw, first=first)
        b.sync()
        hw.close()
    return b, None


def workload_name(args, n_gpus):
    if args.config == 4:
        return (f"BASELINE config 4: 65,536 worlds in total x {BODIES_PER_WORLD} mixed bodies ({N_DYNAMIC} dynamic + 4 static), "
                f"20 iterations, PDPHYZX_UNIT_M, sharded over {n_gpus} GPU(s) = {args.worlds} worlds/GPU, penned piles settled "
                f"{args.settle} substeps")
    if args.control:
        return (f"{args.worlds} worlds/GPU x polygon stacks ({N_DYNAMIC} dynamic bodies), {args.iterations} iterations, "
                f"PDPHYZX_UNIT_MM, host-driven applyImpulse per world + setGravity before every step of "
                f"{args.fused} substeps (BASELINE config 5)")
    return (f"{args.worlds} worlds/GPU x {BODIES_PER_WORLD} {args.mix} bodies ({N_DYNAMIC} dynamic + 4 static), "
            f"{args.iterations} iterations, PDPHYZX_UNIT_M, penned pile settled {args.settle} substeps "
            f"(north_star target scene: 65,536 worlds on 8 GPUs)")


# ------------------------------------------------------------------------ reference arm
def run_reference(args):
    """The unmodified reference (oracle/_ref) on the host cores: --impl reference."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import refbind
    cores = os.cpu_count() or 1
    n_worlds = args.cpu_worlds or cores
    variant = "wide"
    if not refbind.available(variant, 1.0):
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libpxref_wide_m.so was not built"}))
        return 0
    scenes = build_scenes(args, 0, n_worlds)
    worlds = [refbind.RefWorld(s, variant) for s in scenes]
    settle = min(args.settle, 400)  # bounded: settling is untimed but costs CPU seconds
    refbind.time_substeps(worlds, settle, cores)
    for _ in range(args.warmup):
        refbind.time_substeps(worlds, args.fused, cores)
    t = 0.0
    for _ in range(args.steps):
        t += refbind.time_substeps(worlds, args.fused, cores)
    body_substeps = n_worlds * BODIES_PER_WORLD * args.fused * args.steps
    value = body_substeps / t
    sample = f"{n_worlds} worlds x {args.fused * args.steps} substeps after {settle} settle substeps, {cores} pthreads"
    out = {
        "impl": "reference", "metric": "body_steps_per_sec", "value": value, "unit": "body-substeps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, args.gpus), "substeps_per_step": args.fused,
                   "body_frame_steps_per_sec": value / 5.0},
        "cpu_baseline": {"value": value, "unit": "body-substeps/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "body-substeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out))
    return 0


# ------------------------------------------------------------------------------ our arm
def cpu_baseline(args, batch, blocks):
    """Reference CPU step timed on the host cores from the device's settled state."""
    from oracle import refbind
    cores = os.cpu_count() or 1
    n_worlds = min(args.cpu_worlds or 2 * cores, batch.n)
    if not refbind.available("wide", 1.0):
        return None
    scenes = build_scenes(args, 0, n_worlds)
    worlds = [refbind.RefWorld(s, "wide") for s in scenes]
    ncol = worlds[0].lib.ref_body_row_floats()
    L = batch.L
    for w, rw in enumerate(worlds):
        nd = int(blocks.header[w]["nDynamic"])
        order = blocks.order[w, :nd].copy()
        rows = np.zeros((255, ncol), np.float32)
        dyn = blocks.dyn[w]
        rows[:, :14][: min(255, L.maxDynamic)] = dyn.T[:255]
        # the reference keeps an AABB per body: pos -/+ radius (px_body.c:95-102)
        rad = blocks.dync[w][5][:255]
        rows[:, 14] = rows[:, 0] - rad
        rows[:, 15] = rows[:, 1] - rad
        rows[:, 16] = rows[:, 0] + rad
        rows[:, 17] = rows[:, 1] + rad
        rw.set_bodies(rows, blocks.flags[w, :255].copy(), order, True)
    substeps = max(5, min(50, int(2.0e6 * cores / (n_worlds * BODIES_PER_WORLD * 10))))
    refbind.time_substeps(worlds, 2, cores)
    best = None
    for _ in range(3):
        t = refbind.time_substeps(worlds, substeps, cores)
        best = t if best is None else min(best, t)
    value = n_worlds * BODIES_PER_WORLD * substeps / best
    return {"value": value, "unit": "body-substeps/s", "cores": cores, "kind": "reference",
            "sample": f"first {n_worlds} worlds of the batch from the device's settled state, {substeps} substeps, "
                      f"best of 3, {cores} pthreads over disjoint worlds"}


def run_ours(args):
    import torch
    import torch.distributed as dist
    from pdphyzx_b200 import batch as pxb

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the product has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    numa_cpus = None
    if world_size > 1:
        numa_cpus = bind_to_gpu_numa_node(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    # a real (non-default) torch stream, made current, so torch.cuda.Event times the stream
    # the kernels are launched on
    tstream = torch.cuda.Stream()
    torch.cuda.set_stream(tstream)
    b, worlds = make_batch(pxb, args, rank, local_rank, tstream.cuda_stream)
    info = b.kernel_info()

    # settle the piles (untimed), in chunks so a hung kernel cannot run for minutes
    done = 0
    while done < args.settle:
        k = min(50, args.settle - done)
        b.step(k)
        done += k
    b.sync()
    if args.timed_iterations >= 0:
        b.set_iterations(0xFFFFFFFF, args.timed_iterations)
    for _ in range(max(args.warmup, 3)):
        b.step(args.fused)
    b.sync()
    prof0 = prof0_raw = nsub0 = None
    if args.profile:  # phase shares of the settled piles only: subtract what the free fall accumulated
        prof0_raw = b.profile()
        prof0 = prof0_raw.astype(np.float64)
        nsub0 = float(b.download_blocks().header["substeps"].mean())

    def barrier():
        torch.cuda.synchronize()
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: CUDA events on the launching stream ------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launches0 = b.launches
    barrier()
    ctl = None
    if args.control:
        rng = np.random.default_rng(args.seed + rank)
        ctl = (np.arange(b.n, dtype=np.uint32), rng.integers(0, N_DYNAMIC, b.n).astype(np.uint32),
               rng.uniform(-2.0, 2.0, (b.n, 2)).astype(np.float32))

    def control(i):
        if ctl is not None:  # the commands are queued on the host and applied by a kernel before the substeps
            b.apply_impulses(ctl[0], ctl[1], ctl[2])
            b.set_gravity(0xFFFFFFFF, 0.0, 9.8 + 0.001 * (i % 7))

    control(0)
    b.step(args.fused)
    barrier()
    launches0 = b.launches
    t_host0 = time.perf_counter()
    e0.record()
    for i in range(args.steps):
        control(i)
        b.step(args.fused)
    e1.record()
    barrier()
    host_ms = 1e3 * (time.perf_counter() - t_host0)
    ms = e0.elapsed_time(e1)
    gpu_launches = b.launches - launches0
    clocks = sampler.stop()
    # per-kernel device time, live: CUDA events around EVERY launch of a few extra steps
    # (pxBatchStepTimedKernels); feeds the roofline of the dominant kernel below
    pipe = b.pipeline_info()
    kt = {"sweep_ms": 0.0, "sweep_launches": 0, "sched_ms": 0.0, "sched_launches": 0, "solve_ms": 0.0, "solve_launches": 0}
    for _ in range(3):
        control(0)
        r = b.step_timed_kernels(args.fused)
        for k_ in kt:
            kt[k_] += r[k_]

    # ---- end to end through the C ABI with host buffers -------------------------------
    # pxBatchStepHost: pinned host blocks -> HBM -> FUSED substeps -> pinned host blocks, per world
    # range, copies of one range pipelined under the substeps of another; the host waits for
    # (and can read) the result of every step
    e2e_ms = None
    mut_bytes = b.n * b.L.mutStride
    if not args.no_e2e:
        b.mutable_d2h()
        b.sync()
        for _ in range(2):
            b.step_host(args.fused, args.e2e_chunks)
            b.sync()
        barrier()
        t0 = time.perf_counter()
        e0.record()
        for _ in range(args.steps):
            b.step_host(args.fused, args.e2e_chunks)
            b.sync()  # the host reads the result of every step
        e1.record()
        barrier()
        e2e_ms = max(e0.elapsed_time(e1), 1e3 * (time.perf_counter() - t0))

    # ---- end to end through PxWorld structs: what a px->world->step over N host worlds costs ----
    # pxBatchUpload (AoS -> SoA on the host threads, both blocks host -> HBM) -> pxBatchStep ->
    # pxBatchReadback (HBM -> host, SoA -> AoS into the same PxWorld structs), waited for every step
    aos_ms = None
    if worlds is not None and not args.no_e2e and not args.no_aos and not args.control:
        b.readback(worlds)  # the host structs still hold the initial scene: bring the settled piles over first
        for _ in range(2):
            b.upload(worlds)
            b.step(args.fused)
            b.readback(worlds)
        barrier()
        n_aos = max(3, min(args.steps, 10))
        t0 = time.perf_counter()
        for _ in range(n_aos):
            b.upload(worlds)
            b.step(args.fused)
            b.readback(worlds)
        aos_ms = 1e3 * (time.perf_counter() - t0) / n_aos
        barrier()

    if world_size > 1:
        t = torch.tensor([ms, e2e_ms or 0.0, aos_ms or 0.0], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms, e2e_max, aos_max = float(t[0]), float(t[1]), float(t[2])
        e2e_ms = e2e_max if e2e_ms is not None else None
        aos_ms = aos_max if aos_ms is not None else None

    blocks = b.download_blocks()
    stats = np.array([blocks.header["statManifolds"].mean(), blocks.header["statPairs"].mean(),
                      blocks.header["statSleeping"].mean(), float((blocks.header["statOverflow"] != 0).sum())])
    if world_size > 1:
        # the only collective of the whole run: an end-of-run statistics reduction over NCCL
        t = torch.tensor(stats, device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        stats = t.cpu().numpy()
        stats[:3] /= world_size

    if rank == 0:
        total_worlds = args.worlds * world_size
        body_substeps = total_worlds * BODIES_PER_WORLD * args.fused * args.steps
        value = body_substeps / (ms * 1e-3)
        peak, peak_src = measured_peak_hbm()
        alg = ALG_BYTES[args.mix]
        # Roofline of the dominant kernel.  Fused pipeline: the one kernel, which streams the state once
        # per launch whatever K is.  Split pipeline: the sweep kernel (phases F + A-D: the one that streams
        # the state; the solver kernel takes about as long and is bound by the shared-memory data pipe,
        # see roofline_lsu) -- one launch advances every body by ONE substep, so its algorithmic bytes are
        # bodies x B_alg(K = 1).  Launch durations: CUDA events around every launch, measured live above.
        sweep_ms = kt["sweep_ms"] / max(kt["sweep_launches"], 1)
        solve_ms = kt["solve_ms"] / max(kt["solve_launches"], 1)
        sched_ms = kt["sched_ms"] / max(kt["sched_launches"], 1)
        alg_bytes = args.worlds * BODIES_PER_WORLD * alg
        achieved = alg_bytes / (sweep_ms * 1e-3) / 1e9
        out = {
            "metric": "body_steps_per_sec", "value": value, "unit": "body-substeps/s", "n_gpus": world_size,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {
                "workload": workload_name(args, world_size), "substeps_per_step": args.fused,
                **({"baseline_config": args.config} if args.config else {}),
                **({"host_cpus_per_rank_numa_local": numa_cpus} if numa_cpus else {}),
                "body_frame_steps_per_sec": value / 5.0, "l2": "inputs larger than L2 (state > 126 MB per GPU)"
                if args.worlds * (b.L.mutStride + b.L.immStride) > 126e6 else "inputs smaller than L2",
                "pipeline": (("split: sweep kernel (phase F of the previous substep + phases A-D) and solver kernel (derives the "
                              "substep's periodic visit schedule, then replays it per iteration) per substep" if pipe.get("static_schedule") else
                              "split: sweep kernel (phase F of the previous substep + phases A-D) and dataflow-solver kernel per substep")
                             if pipe["split"] else "fused: one kernel, K substeps"),
                "threads_per_world_cta": info["threads"], "ctas_per_sm": info["ctas_per_sm"],
                "smem_bytes_per_cta": info["smem_bytes"], "solver_ctas_per_sm": pipe["solve_ctas_per_sm"],
                "schedule_ctas_per_sm": pipe["sched_ctas_per_sm"], "seed": args.seed,
                "avg_manifolds_per_world": stats[0], "avg_pairs_per_world": stats[1],
                "avg_sleeping_per_world": stats[2], "worlds_overflowed": int(stats[3]),
                **({"host_wall_ms_per_step": host_ms / args.steps} if args.control else {}),
            },
            "clocks": clocks, "gpu_launches": int(gpu_launches),
            "kernels": {"sweep_ms_per_launch": sweep_ms, "sweep_launches_per_step": kt["sweep_launches"] // 3,
                        "finish_ms_per_launch": sched_ms, "finish_launches_per_step": kt["sched_launches"] // 3,
                        "solve_ms_per_launch": solve_ms, "solve_launches_per_step": kt["solve_launches"] // 3,
                        "timed_with": "CUDA events around every launch (pxBatchStepTimedKernels), 3 steps"},
            "roofline": {"bound": "hbm", "kernel": "pxStepKernel (sweep)" if pipe["split"] else "pxStepKernel (fused)",
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": None, "peak_source": peak_src,
                         "algorithmic_bytes_per_launch": alg_bytes, "launch_ms": sweep_ms,
                         "note": "HBM is the nominal bound of this path (SURVEY 8d) but not the binding one: non-FMA "
                                 "fp32 issue slots at 19.6 of 32 lanes (sweep kernel) and the per-world dependent chain "
                                 "of the order-preserving solver (visit stream out of L2) are; see roofline_issue, "
                                 "roofline_lsu, DESIGN.md 3.5 and profiles/"},
        }
        facts = kernel_facts(args)
        if facts and pipe["split"] == ("sweep" in facts.get("kernels", {})):
            # ncu of this very command line (profiles/summarize.py): DRAM bytes of one launch of the dominant
            # kernel, the data-pipe utilisation of both kernels, and the warp instructions one world-substep
            # costs -> issue-slot utilisation at the rate measured live above
            ks = facts["kernels"]
            dom = ks["sweep"] if pipe["split"] else ks["fused"]
            # the capture may hold launches over a part of the batch (large batches are stepped as two halves on two
            # streams): scaled to the worlds of the launch that `launch_ms` and the algorithmic bytes refer to
            out["roofline"]["traffic"] = dom["dram_bytes_per_launch"] * args.worlds / facts["worlds"]
            out["roofline"]["traffic_source"] = "profiles/" + facts["_file"]
            sm_hz = (clocks.get("sm_mhz") or 1965.0) * 1e6
            n_sms = info.get("num_sms", 148)
            issue_peak = n_sms * 4 * sm_hz  # one warp instruction per scheduler per cycle
            instr = sum(k_["warp_instructions_per_world_substep"] for k_ in ks.values())
            issue_rate = instr * (value / world_size) / BODIES_PER_WORLD
            out["roofline_issue"] = {
                "bound": "issue", "achieved": issue_rate / 1e12, "peak": issue_peak / 1e12, "unit": "T warp-instr/s",
                "frac": issue_rate / issue_peak, "warp_instructions_per_world_substep": instr,
                "per_kernel": {n_: {"warp_instructions_per_world_substep": k_["warp_instructions_per_world_substep"],
                                    "lanes_per_instruction": k_["lanes_per_instruction"], "issue_active_pct": k_["issue_active_pct"]}
                               for n_, k_ in ks.items()},
                "note": "per GPU; instruction counts from ncu (smsp__inst_executed.sum), rate from this run"}
            # the binding unit: wavefronts through the L1 / shared-memory data pipe, one per cycle per SM
            wf = sum((k_["shared_wavefronts"] + k_["global_load_wavefronts"]) / facts["worlds"] / k_["substeps_per_launch"]
                     for k_ in ks.values())
            wf_rate = wf * (value / world_size) / BODIES_PER_WORLD
            out["roofline_lsu"] = {
                "bound": "l1-shared-data-pipe", "achieved": wf_rate / 1e12, "peak": n_sms * sm_hz / 1e12, "unit": "T wavefronts/s",
                "frac": wf_rate / (n_sms * sm_hz), "wavefronts_per_world_substep": wf,
                "per_kernel_pct_of_peak_under_ncu": {n_: k_["l1_data_pipe_wavefronts_pct"] for n_, k_ in ks.items()},
                "note": "shared-memory + global-load wavefronts per world-substep from ncu "
                        "(l1tex__data_pipe_lsu_wavefronts_mem_shared, l1tex__t_output_wavefronts_pipe_lsu_mem_global_op_ld) x the "
                        "rate measured live; peak = one wavefront per cycle per SM"}
        if args.profile:
            pr = b.profile().astype(np.float64) - prof0
            nsub = float(blocks.header["substeps"].mean()) - nsub0
            tot = pr[:, :6].sum() + pr[:, 12].sum()
            out["phase_share"] = {n: float(pr[:, i].sum() / tot) for i, n in enumerate(
                ("A_sort", "B_sweep", "C_narrowphase", "D_lists", "E_solver", "F_integrate"))}
            out["solver_rounds_per_substep"] = float(pr[:, 6].mean() / nsub)
            out["solver_cycles_per_round"] = float(pr[:, 4].sum() / max(pr[:, 6].sum(), 1))
            out["cycles_per_world_substep"] = float(pr[:, :6].sum(axis=1).mean() / nsub)
            c = pr[:, 8:12].sum(axis=0) / pr[:, 2].sum()
            out["narrowphase_stage_share"] = {"C1_circle_pairs": float(c[0]), "C2a_quick_reject": float(c[1]), "C2b_sat": float(c[2]),
                                              "C3_clip": float(c[3]), "C4_manifold_init": float(1.0 - c.sum())}
            if pipe.get("static_schedule"):
                # static schedule: relaxation + sort/stream-build cycles of phase D (slots 12, 14; included in D_lists'
                # neighbour slot 12 of the phase clock), relaxation votes and period (slot 13)
                raw13 = b.profile()[:, 13] - prof0_raw[:, 13]
                out["schedule"] = {
                    "sweep_kernel_hand_over_share_of_cycles": float(pr[:, 12].sum() / tot),
                    "relaxation_cycles_per_substep": float(pr[:, 15].mean() / nsub),
                    "schedule_kernel_cycles_per_substep": float(pr[:, 14].mean() / nsub),
                    "relaxation_passes_per_substep": float((raw13 >> np.uint64(32)).mean() / nsub),
                    "period_rounds": float((raw13 & np.uint64(0xFFFFFFFF)).mean() / nsub),
                    "replay_chunks_per_substep": float(pr[:, 6].mean() / nsub)}
            raw7 = b.profile()[:, 7] - prof0_raw[:, 7]
            out["polygon_pairs_per_substep"] = float((raw7 >> np.uint64(32)).mean() / nsub)
            out["polygon_pairs_after_quick_reject_per_substep"] = float((raw7 & np.uint64(0xFFFFFFFF)).mean() / nsub)
        if e2e_ms is not None:
            out["e2e"] = {"value": body_substeps / (e2e_ms * 1e-3), "unit": "body-substeps/s",
                          "h2d_bytes_per_step": int(mut_bytes), "d2h_bytes_per_step": int(mut_bytes),
                          "ms_per_step": e2e_ms / args.steps}
        if aos_ms is not None:
            aos_bytes = int(b.n * (b.L.mutStride + b.L.immStride))
            out["e2e_aos"] = {"value": args.worlds * world_size * BODIES_PER_WORLD * args.fused / (aos_ms * 1e-3),
                              "unit": "body-substeps/s", "ms_per_step": aos_ms, "h2d_bytes_per_step": aos_bytes,
                              "d2h_bytes_per_step": int(mut_bytes), "host_threads": min(os.cpu_count() or 1, (b.n + 63) // 64, 64),
                              "note": "pxBatchUpload -> pxBatchStep -> pxBatchReadback over host PxWorld structs (AoS, 664-byte "
                                      "bodies): the host-side AoS <-> SoA conversion dominates; e2e above moves pre-packed SoA blocks"}
        if world_size == 1 and not args.no_cpu:
            try:
                cb = cpu_baseline(args, b, blocks)
            except Exception as exc:  # the baseline must never take the GPU number down with it
                cb = {"value": None, "unit": "body-substeps/s", "cores": os.cpu_count(), "kind": "reference",
                      "sample": f"failed: {exc}"}
            if cb:
                out["cpu_baseline"] = cb
        print(json.dumps(out))
    b.close()
    if worlds is not None:
        worlds.close()
    if world_size > 1:
        dist.destroy_process_group()
    return 0


def run_sharded(args):
    """--sharded: ONE process, one PxBatchGroup (C ABI) over --gpus devices; the calling thread fans
    every step out, the devices run side by side.  Device-timed with CUDA events per device, the
    slowest device counts.  The comparison line for the one-process-per-GPU mode."""
    import torch
    from pdphyzx_b200 import batch as pxb
    n_gpus = args.gpus
    if torch.cuda.device_count() < n_gpus:
        raise SystemExit(f"bench.py --sharded: {n_gpus} GPUs asked for, {torch.cuda.device_count()} visible")
    total = args.worlds if args.config == 4 else args.worlds * n_gpus
    devices = list(range(n_gpus))
    g = None
    for first in range(0, total, CHUNK):
        n = min(CHUNK, total - first)
        hw = pxb.HostWorlds(build_scenes(args, 0, n, first))
        if g is None:
            d = pxb.make_desc(hw.unit)
            pxb._check(pxb.core().pxBatchFitDesc(hw.pointer_array(), n, pxb.C.byref(d)), "pxBatchFitDesc")
            g = pxb.BatchGroup(None, devices, threads=args.threads, n_worlds=total, unit=hw.unit, max_dynamic=d.maxDynamic,
                               max_static=d.maxStatic, max_vertices=max(d.maxVertices + 256, N_DYNAMIC * 6),
                               max_manifolds=args.max_manifolds, max_pairs=args.max_pairs)
        g.upload(hw, first)
        g.sync()
        hw.close()
    done = 0
    while done < args.settle:
        k = min(50, args.settle - done)
        g.step(k)
        done += k
    for _ in range(max(args.warmup, 3)):
        g.step(args.fused)
    g.sync()
    samplers = [ClockSampler(i) for i in devices[:1]]
    for s_ in samplers:
        s_.start()
    shards = [g.shard(s)[0] for s in range(g.n_shards)]
    launches0 = g.launches()
    for h in shards:
        pxb._check(g.lib.pxBatchTimerStart(h), "pxBatchTimerStart")
    for _ in range(args.steps):
        g.step(args.fused)
    ms = 0.0
    for h in shards:
        t = pxb.C.c_float(0)
        pxb._check(g.lib.pxBatchTimerStop(h, pxb.C.byref(t)), "pxBatchTimerStop")
        ms = max(ms, float(t.value))
    gpu_launches = g.launches() - launches0
    clocks = samplers[0].stop()
    # the host buffers must hold the settled state the device holds (they still hold the uploaded one)
    for s_ in range(g.n_shards):
        h, first_w, n_w = g.shard(s_)
        pxb._check(g.lib.pxBatchMutableD2H(h, None, 0, n_w), "pxBatchMutableD2H")
    g.sync()
    for _ in range(2):
        g.step_host(args.fused, args.e2e_chunks)
        g.sync()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        g.step_host(args.fused, args.e2e_chunks)
        g.sync()
    e2e_ms = 1e3 * (time.perf_counter() - t0)
    L = g.shard_layout(0)
    body_substeps = total * BODIES_PER_WORLD * args.fused * args.steps
    out = {
        "metric": "body_steps_per_sec", "value": body_substeps / (ms * 1e-3), "unit": "body-substeps/s", "n_gpus": n_gpus,
        "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": args.scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_name(args, n_gpus), "substeps_per_step": args.fused,
                   "mode": f"single process, PxBatchGroup over {n_gpus} devices ({g.n_shards} shards)",
                   "l2": "inputs larger than L2 (state > 126 MB per GPU)"},
        "clocks": clocks, "gpu_launches": int(gpu_launches),
        "e2e": {"value": body_substeps / (e2e_ms * 1e-3), "unit": "body-substeps/s",
                "h2d_bytes_per_step": int(total * L.mutStride), "d2h_bytes_per_step": int(total * L.mutStride),
                "ms_per_step": e2e_ms / args.steps},
    }
    print(json.dumps(out))
    g.close()
    return 0


def main():
    args = parse_args()
    if args.control:  # the control bench reports the device-timed rate and the host-side queueing time only
        args.no_cpu = args.no_e2e = True
    if args.impl == "reference":
        return run_reference(args)
    if args.sharded:
        return run_sharded(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
