#!/usr/bin/env python
"""bench.py -- particle-steps/s of the fused Gerris particulate hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config C2|C3|C5] [--particles n]
                  [--two-way] [--impl b200|reference]

A "step" is one gfs_particle_list_event-equivalent pass over the resident
particle batch: the per-field-update cell pass (vertex velocity table +
vorticity table; the fluid field changes every step in a coupled run, so it
is redone every step here even though the synthetic field is frozen), the
fused locate+interpolate+force+integrate kernel, a re-sort by cell every
--resort steps, and with --two-way the void-fraction/force deposition plus an
NCCL all-reduce of the deposited field.  At N=1 the workload is BASELINE
config C2 (128^3 uniform octree, frozen Taylor-Green field, 10 M particles,
drag+lift+buoyancy).  For N>1 particles are sharded (weak scaling: every rank
holds the full per-GPU batch), tree and field are replicated.

Runs longer than 100 steps are timed in segments of 100 steps, each started from the initial
cloud (restored outside the CUDA-event brackets), because the sedimenting cloud starts leaving
the closed box after ~150 steps and the run would otherwise time early-outs; `ms_per_step` is
the sum of the segment times over the number of steps.

Prints ONE JSON line (rank 0).  `value` is whole-job particle-steps/s with
inputs resident in HBM; `e2e` is the same metric through the host-buffer
C-ABI call (H2D of the particle arrays and the field, D2H of the new state
inside the timed region); `roofline` describes the fused step kernel -- its
launch duration is measured live, with CUDA events on the kernel's stream inside
the library, around every 4th launch of the timed region (--kernel-timer-every;
the events are stream operations between the cell pass and the step kernel and
cost 1.5 % of a step when every launch carries them; `kernel_launches` is the
number of launches averaged over);
`cpu_baseline` is the reference's own object code (oracle/_ref/libgfsrefobj:
modules/particulatecommon.c, src/event.c, src/particle.c, src/fluid.c, src/ftt.c
compiled unmodified) running gfs_event_do on a GfsParticleList on one host core
-- the reference is serial per MPI rank.

--impl reference times the same object code on all host cores the way the
reference scales, as independent single-threaded ranks (forked processes, each
with its share of the particles and the whole replicated domain).
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
import __graft_entry__ as entry  # noqa: E402

BYTES_PER_PARTICLE_STEP = {3: 112, 2: 80}
COLS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")


def measured_traffic(world_name, n_local, dim):
    """dram__bytes_read.sum + dram__bytes_write.sum of the step kernel per launch, from the
    committed `ncu --set full` capture (profiles/step_kernel_traffic.json), scaled per particle"""
    try:
        with open(os.path.join(ROOT, "profiles", "step_kernel_traffic.json")) as f:
            t = json.load(f)
        if t.get("config") == world_name and t.get("dim") == dim:
            return t["dram_bytes_per_particle"] * n_local
    except Exception:
        pass
    return None


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """SM clock and throttle reasons sampled through NVML (nvidia_ml_py) from a
    thread while the timed region runs (nvidia-smi's start-up alone is longer
    than a 100-step region)."""

    def __init__(self, gpu: int, period_s: float = 0.002):
        self.gpu, self.period, self.rows = gpu, period_s, []
        self._stop = threading.Event()
        self._thread = None
        self._h = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(gpu)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._h = None
            self.max_sm = None

    def _loop(self):
        nv, h = self._nv, self._h
        while not self._stop.is_set():
            try:
                t0 = time.perf_counter()
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                try:
                    why = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    why = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.rows.append((float(sm), int(why), t0, time.perf_counter()))
            except Exception:
                pass
            time.sleep(self.period)

    def start(self):
        if self._h is not None:
            self._thread = threading.Thread(target=self._loop, daemon=True)
            self._thread.start()

    def stop(self, t_begin=None, t_end=None):
        """statistics over the samples whose NVML queries overlap [t_begin, t_end] (host clock).  The
        headline region is a few milliseconds and one query can take that long on a busy GPU: when no
        query overlaps it the window is widened (50 ms on either side -- the warm-up steps before and the
        two-way region after run the same kernels --, then the whole run) and `window` says so."""
        self._stop.set()
        if self._thread:
            self._thread.join()
        window = "all samples"
        if t_begin is not None:
            window = "timed region"
            rows = [r for r in self.rows if r[3] >= t_begin and r[2] <= t_end]
            if not rows:
                window = "timed region +- 50 ms"
                rows = [r for r in self.rows if r[3] >= t_begin - 0.05 and r[2] <= t_end + 0.05]
            if not rows:
                window = "whole run (no NVML query overlapped the timed region)"
                rows = list(self.rows)
            self.rows = rows
        names = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown",
                 0x4: "sw_power_cap"}
        sm = [r[0] for r in self.rows]
        reasons = sorted({n for r in self.rows for bit, n in names.items() if r[1] & bit})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": self.max_sm,
                "reasons": reasons, "samples": len(sm), "window": window}


def build_world(worlds, name, n_particles, level=0):
    if name == "C2":
        return worlds.make_c2(level=level or 7, n_particles=n_particles or 10_000_000)
    if name == "C3":
        return worlds.make_c3(n_particles=n_particles or 10_000_000)
    if name == "C4":
        return worlds.make_c4(n_particles=n_particles or 50_000_000)
    if name == "C5":
        return worlds.make_c5(n_particles=n_particles or 200_000_000)
    if name == "C1":
        return worlds.make_c1(n_particles=n_particles or 1000)
    raise SystemExit(f"unknown config {name}")


# --------------------------------------------------------------------------
# reference arm: the reference's CPU path (oracle/_ref), all host threads

def oracle_world(ora, worlds, name):
    """Reference-built tree + analytic field for the config, oracle only."""
    sp = worlds.spec(name)
    sim = ora.Sim(sp.dim, nvar=4)
    m = sp.meta
    if m["field"] == "lid":
        for s in range(4):
            sim.add_boundary(s)
        sim.refine_uniform(m["level"])
    elif m["field"] == "tg":
        sim.refine_uniform(m["level"])
    else:
        sim.refine_ring(m["levels"][0], m["levels"][1], 0.25, 1.5)
        sim.corner_sweep()
    sim.finalize()
    ptr, pos, level, leaf = sim.export_cells()
    f = worlds.field_of(sp, pos)
    for i in range(sp.dim):
        sim.set_values(i, ptr, f[i])
    return sp, sim


def time_oracle(ora, worlds, helpers_params, sp, sim, n_sample, steps, warmup, threads, pattern=0):
    parts = worlds.make_particles(sp, n_sample)
    plist = ora.ParticleList(sim, *[parts[k] for k in COLS])
    par = helpers_params(sp, pattern)
    for _ in range(warmup):
        plist.step(par, threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        plist.step(par, threads)
    dt = time.perf_counter() - t0
    return n_sample * steps / dt, dt


def time_refobj(ora, worlds, helpers_params, sp, sim, n_sample, steps, warmup, first=0, timers=True):
    """particle-steps/s of the reference's own GfsParticleList event (cull, per-particle
    GfsEvent gating + timers, GfsParticulate events, gfs_particle_bc) on one core"""
    parts = worlds.make_particles(sp, first + n_sample)
    par = helpers_params(sp, 0)
    rs = ora.RefSim(sim)
    rs.configure(par, timers)
    rl = ora.RefParticleList(rs, *[parts[k][first:] for k in COLS], par)
    if warmup:
        rl.event(warmup)
    t0 = time.perf_counter()
    rl.event(steps)
    dt = time.perf_counter() - t0
    left = len(rl)
    rs.close()
    return n_sample * steps / dt, dt, left


def _refobj_rank(ora, worlds, mk, sp, sim, first, n, steps, warmup, barrier, out):
    parts = worlds.make_particles(sp, first + n)
    par = mk(sp, 0)
    rs = ora.RefSim(sim)
    rs.configure(par, True)
    rl = ora.RefParticleList(rs, *[parts[k][first:] for k in COLS], par)
    if warmup:
        rl.event(warmup)
    barrier.wait()
    t0 = time.perf_counter()
    rl.event(steps)
    out.put(time.perf_counter() - t0)
    os._exit(0)


def time_refobj_ranks(ora, worlds, mk, sp, sim, n_per_rank, steps, warmup, ranks):
    """`ranks` forked single-threaded processes, each running the reference's list event
    on its own n_per_rank particles of the cloud; the domain (built before the fork) is
    shared copy-on-write.  Returns (aggregate particle-steps/s, slowest rank's seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    barrier, out = ctx.Barrier(ranks), ctx.SimpleQueue()
    procs = [ctx.Process(target=_refobj_rank, args=(ora, worlds, mk, sp, sim, r * n_per_rank, n_per_rank,
                                                    steps, warmup, barrier, out)) for r in range(ranks)]
    for p in procs:
        p.start()
    times = [out.get() for _ in procs]
    for p in procs:
        p.join()
    dt = max(times)
    return ranks * n_per_rank * steps / dt, dt


def oracle_params_factory(ora):
    fmap = {1: ora.FORCE_DRAG, 2: ora.FORCE_LIFT, 3: ora.FORCE_BUOY}

    def make(sp, pattern):
        return ora.step_params(sp.dt, [fmap[f] for f in sp.forces], rho=sp.rho, mu=sp.mu, g=sp.g, pattern=pattern)
    return make


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    pkg = entry.load_package()          # worlds.spec / analytic fields only: no product compute
    ora = entry.load_oracle()
    worlds = pkg.worlds
    cores = len(os.sched_getaffinity(0))
    sp, sim = oracle_world(ora, worlds, args.config)
    mk = oracle_params_factory(ora)
    # size the per-step sample so that one step takes about a second and the whole run
    # (steps + warmup) about a minute and a half at most
    rate1, _, _ = time_refobj(ora, worlds, mk, sp, sim, 10_000, 1, 1)
    per_step_s = min(1.0, 90.0 / max(args.steps + args.warmup, 1))
    n_rank = int(min(sp.n_particles // cores, max(2_000, rate1 * per_step_s)))
    value, dt = time_refobj_ranks(ora, worlds, mk, sp, sim, n_rank, args.steps, args.warmup, cores)
    n_sample = n_rank * cores
    line = {
        "impl": "reference", "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(args.config), "sample_particles_per_step": n_sample},
        "cpu_baseline": {"value": value, "unit": "particle-steps/s", "cores": cores, "kind": "reference",
                         "sample": f"{n_sample} particles/step of the {args.config} cloud x {args.steps} steps "
                                   f"on the full tree, as {cores} single-threaded ranks of {n_rank} particles; "
                                   "the reference's own object code (particulatecommon.c, event.c, particle.c, "
                                   "fluid.c, ftt.c unmodified): gfs_event_do on a GfsParticleList",
                         "one_rank": rate1},
        "e2e": {"value": value, "unit": "particle-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_name(cfg):
    return {
        "C1": "C1: 2D lid-style level-6 quadtree, 1k particles, drag",
        "C2": "C2: 3D uniform level-7 octree (128^3), frozen Taylor-Green field, 10M one-way particles, drag+lift+buoyancy",
        "C3": "C3: 3D adaptive octree levels 5-9 around a vortex ring, 10M particles, drag+lift+buoyancy",
        "C4": "C4: two-way, adaptive octree levels 6-10, 50M particles, deposit + all-reduce",
        "C5": "C5: 3D adaptive octree levels 6-10, 200M particles, drag+lift+buoyancy",
    }[cfg]


# --------------------------------------------------------------------------
# B200 arm

def device_cloud(torch, world, n_local, rank, device):
    """The ring-config cloud (half uniform, half Gaussian around the core; v0 = analytic fluid
    velocity) drawn ON THE DEVICE with torch's generator -- plumbing for the 200 M-particle C5
    line, where drawing and shipping the cloud from the host would dominate the run.  Returns the
    8 SoA columns as CUDA tensors."""
    g = torch.Generator(device=device)
    g.manual_seed(world.seed * 1000 + rank)
    f64 = dict(dtype=torch.float64, device=device, generator=g)
    n_uni = n_local // 2
    n_g = n_local - n_uni
    pos = []
    theta = torch.rand(n_g, **f64) * (2.0 * np.pi)
    for a in range(3):
        uni = torch.rand(n_uni, **f64) * 0.9 - 0.45
        off = torch.randn(n_g, **f64) * 0.08
        if a == 0:
            off += 0.25 * torch.cos(theta)
        elif a == 1:
            off += 0.25 * torch.sin(theta)
        pos.append(torch.cat([uni, off]).clamp_(-0.49, 0.49))
        del uni, off
    del theta
    m = world.meta
    d = torch.rand(n_local, **f64) * (m["d_p"][1] - m["d_p"][0]) + m["d_p"][0]
    vol = d ** 3 * (np.pi / 6.0)
    mass = vol * m["rho_p"]
    del d
    # worlds.vortex_ring, restated for tensors
    x, y, z = pos
    R, gamma, a0 = 0.25, 1.0, 0.05
    rho = torch.sqrt(x * x + y * y)
    dr = rho - R
    s2 = dr * dr + z * z
    s2c = torch.clamp(s2, min=1e-300)
    k = gamma / (2.0 * np.pi) * (1.0 - torch.exp(-s2 / (a0 * a0))) / s2c
    ur, w = -k * z, k * dr
    rs = torch.clamp(rho, min=1e-300)
    vel = [ur * x / rs, ur * y / rs, w]
    return pos + vel + [mass, vol]


def fill_from_device(torch, ctx, cols, device):
    """copy CUDA tensors into the context's SoA columns (device to device)"""
    n = int(cols[0].numel())
    ctx.particles_resize(n)
    for ptr, src in zip(ctx.particles_device_ptrs(), cols):
        dst = torch.as_tensor(_Alias(ptr, n), device=device)
        dst.copy_(src)
    torch.cuda.synchronize()


class _Alias:
    def __init__(self, ptr, count):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3}


def stats_ms(ms):
    a = np.asarray(ms, dtype=np.float64)
    if a.size == 0:
        return None
    return {"min": float(a.min()), "median": float(np.median(a)), "max": float(a.max()), "n": int(a.size)}


def run_b200(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world_size = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    if world_size > 1:
        # torch.distributed is rendezvous plumbing only (the NCCL id, the max over ranks of the
        # timings): every data-path collective is issued by the C library on its own communicator
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    def allmax(v):
        if world_size == 1:
            return float(v)
        t = torch.tensor([v], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def allsum(v):
        if world_size == 1:
            return float(v)
        t = torch.tensor([v], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    pkg = entry.load_package()
    capi, worlds = pkg.capi, pkg.worlds
    peak, peak_src = measured_peak()
    ctx = capi.Context(local_rank)
    # The kernel-time events of the library are stream operations BETWEEN the cell pass and the step
    # kernel: around every launch they cost 1.5 % of a C2 step (0.2885 vs 0.2835 ms).  They are
    # recorded around every 4th launch of the timed region (--kernel-timer-every 1: all of them);
    # roofline.kernel_launches says how many launches the average is over.
    ctx.timer_sampling(args.kernel_timer_every)
    uid = [capi.comm_unique_id() if (rank == 0 and world_size > 1) else None]
    if world_size > 1:
        dist.broadcast_object_list(uid, src=0)
    comm = capi.Comm.init_rank(ctx, uid[0], rank, world_size)
    stream = torch.cuda.ExternalStream(ctx.stream, device=device)

    def order_particles():
        """one GPU: sort by cell.  Several: every rank takes a contiguous slice of the GLOBAL cell
        order -- its tables shrink to 1/N of the tree and its deposits to its own slice"""
        if world_size == 1 and ordered[0]:
            ctx.sort()                   # the one slice stays the whole tree: a plain re-sort
        else:
            comm.rebalance()
            ordered[0] = True

    ordered = [False]

    def barrier():
        if world_size > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ctx.synchronize()

    def timed_steps(one_step, steps, tail=None):
        """`steps` calls of one_step(i) between CUDA events on the context's stream, plus an event
        after every step; returns (total ms, [ms per step])"""
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for i in range(steps):
            one_step(i)
            if i == steps - 1 and tail is not None:
                tail()
            ev[i + 1].record(stream)
        barrier()
        per = [ev[i].elapsed_time(ev[i + 1]) for i in range(steps)]
        return ev[0].elapsed_time(ev[steps]), per

    # ======================================================================
    # headline: one-way, weak scaling (every rank holds the config's full batch)
    world = build_world(worlds, args.config, args.particles, args.level)
    n_share = world.n_particles
    ctx.upload_tree(world.tree)
    ctx.upload_field(world.u, world.v, world.w)
    parts = worlds.make_particles(world, n_share * world_size, rank, world_size)
    ctx.particles_upload(**parts)
    order_particles()
    ctx.synchronize()
    n_local = ctx.count
    total_particles = int(allsum(n_local))
    par = world.step_params()

    def one_step(i):
        ctx.refresh_field()
        ctx.step(par)
        if args.resort and (i + 1) % args.resort == 0:
            order_particles()

    # the sampler thread starts before the warm-up (NVML's first calls are slow); only the
    # samples taken while the timed region runs are kept
    sampler = ClockSampler(local_rank, period_s=0.0005)
    if rank == 0:
        sampler.start()
    # The synthetic cloud sediments (rho_p/rho = 1000, g = -1): past ~150 steps particles start
    # leaving through the bottom wall and the kernel would early-out for them.  Long runs are
    # therefore timed in SEGMENTS of at most 100 steps, each started from the initial cloud
    # (re-uploaded and re-ordered OUTSIDE the timed events); every segment must end with
    # >= 99.9 % of its particles inside or the run aborts.
    SEG = 100

    def restore():
        ctx.particles_upload(**parts)
        order_particles()
        barrier()

    def check_inside(what):
        before = ctx.count
        removed = ctx.cull()
        frac = 1.0 - removed / max(before, 1)
        if frac < 0.999:
            raise SystemExit(f"bench.py: only {frac:.4f} of the particles are still inside the domain "
                             f"after {what} -- the workload is invalid")
        return frac

    for i in range(min(args.warmup, SEG)):
        one_step(i)
    barrier()
    if args.warmup > 10:
        restore()
    ctx.timer_reset()
    t_region0 = time.perf_counter()
    launches0 = capi.kernel_launches()               # counted inside the library, per launch
    ms, done, n_segments, inside_frac, untimed_launches, per_step = 0.0, 0, 0, 1.0, 0, []
    while done < args.steps:
        m = min(SEG, args.steps - done)
        if done:
            skip = capi.kernel_launches()
            restore()
            untimed_launches += capi.kernel_launches() - skip      # untimed: not part of the claim
        seg_ms, per = timed_steps(lambda i: one_step(done + i), m)
        ms += seg_ms
        per_step += per
        done += m
        n_segments += 1
        skip = capi.kernel_launches()
        inside_frac = min(inside_frac, check_inside("a timed segment"))
        untimed_launches += capi.kernel_launches() - skip
    gpu_launches = capi.kernel_launches() - launches0 - untimed_launches
    kernel_ms, kernel_launches = ctx.timer_read()
    t_region1 = time.perf_counter()
    n_sorts = (args.steps // args.resort) if args.resort else 0
    ms = allmax(ms)
    value = total_particles * args.steps / (ms * 1e-3)

    # ======================================================================
    # two-way: the same step with both deposits fused into the step kernel, and the sum of the
    # deposited field over the ranks (gfsb200_deposit_allreduce) -- the path north_star names for
    # multi-GPU.  Timed like the headline; the last exchange must have landed inside the region.
    two_way = None
    if not args.no_two_way:
        restore()
        par2 = world.step_params(fuse_deposit=True)

        def two_way_step(i):
            ctx.refresh_field()
            ctx.step(par2)
            comm.deposit_allreduce()

        for i in range(3):
            two_way_step(i)
        comm.deposit_wait()
        barrier()
        ctx.timer_reset()
        comm.exchange_stats()
        k2 = min(args.steps, SEG)
        t2_ms, t2_per = timed_steps(two_way_step, k2, tail=comm.deposit_wait)
        t2_kernel_ms, _ = ctx.timer_read()
        x_ms, x_n, x_bytes = comm.exchange_stats()
        t2_ms = allmax(t2_ms)
        # in-bench check of the exchanged field: volume conservation over ALL ranks and the same
        # bytes on every rank
        a = world.arrays
        field = [ctx.download_deposit(c) for c in range(1 + world.dim)]
        cons = float(np.sum(field[0] * a.h ** world.dim))
        state = ctx.particles_download()
        inside = ctx.locate(state["x"], state["y"], state["z"]) >= 0
        vp = allsum(float(state["volume"][inside].sum()))
        import hashlib
        digest = hashlib.sha1(b"".join(np.ascontiguousarray(f).tobytes() for f in field)).hexdigest()
        digests = [digest]
        if world_size > 1:
            digests = [None] * world_size
            dist.all_gather_object(digests, digest)
        check_inside("the two-way region")
        b2 = 136 if world.dim == 3 else 96
        ach2 = n_local * b2 / (t2_kernel_ms * 1e-3) / 1e9 if t2_kernel_ms > 0 else 0.0
        whole2 = n_local * b2 / (t2_ms / k2 * 1e-3) / 1e9
        two_way = {
            "value": total_particles * k2 / (t2_ms * 1e-3), "unit": "particle-steps/s", "steps": k2,
            "ms_per_step": t2_ms / k2, "step_ms": stats_ms(t2_per),
            "what": "per step: cell pass + ONE fused kernel (locate, interpolate, forces, integrate, "
                    "then void fraction and on-fluid force deposited at the new state) + "
                    "gfsb200_deposit_allreduce; the last exchange lands inside the timed region",
            "kernel": "step_kernel_wpipe<%d,...,DEP> (fused step + deposit)" % world.dim,
            "kernel_ms": t2_kernel_ms,
            "roofline": {"bound": "hbm", "algorithmic_bytes_per_particle_step": b2, "achieved": ach2,
                         "peak": peak, "unit": "GB/s", "frac": ach2 / peak,
                         "whole_step_achieved": whole2, "whole_step_frac": whole2 / peak},
            "exchange": {"mode": ("owner slices: remote fp64 reductions over NVLink peer memory in the deposit "
                                  "kernel + all-gather of the slices by the copy engines" if comm.peer_access
                                  else "ncclAllReduce of the whole buffer"),
                         "peer_access": bool(comm.peer_access), "device_ms": x_ms, "n": x_n,
                         "bytes_sent_per_rank": int(x_bytes),
                         "busbw_GBs": (x_bytes / (x_ms * 1e-3) / 1e9) if x_ms > 0 else None,
                         "field_bytes": int((1 + world.dim) * a.n_cells * 8),
                         "overlapped": "runs on a communication stream behind the deposit; the next "
                                       "step's cell pass and step kernel do not wait for it"},
            "check": {"sum_field_times_cell_volume": cons, "sum_particle_volume_all_ranks": vp,
                      "rel_err": abs(cons - vp) / vp if vp else None,
                      "conserved_1e-12": bool(vp and abs(cons - vp) <= 1e-12 * vp),
                      "identical_on_all_ranks": len(set(digests)) == 1},
        }
        if two_way["check"]["rel_err"] is None or two_way["check"]["rel_err"] > 1e-9 or \
                not two_way["check"]["identical_on_all_ranks"]:
            raise SystemExit(f"bench.py: the exchanged deposit failed its check: {two_way['check']}")

    # ======================================================================
    # e2e: host buffers in, host buffers out, through the C-ABI
    restore()
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    host = {k: torch.from_numpy(np.ascontiguousarray(parts[k])).pin_memory().numpy() for k in COLS if parts[k] is not None}
    fields = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory().numpy()
              for f in (world.u, world.v, world.w) if f is not None]
    n_host = len(host["x"])
    h2d = sum(a.nbytes for a in host.values()) + sum(a.nbytes for a in fields)
    d2h = n_host * 8 * 2 * world.dim

    def e2e_step():
        ctx.upload_field(*fields)                     # mirror U,V,W + cell pass
        # host particle arrays in, updated host particle arrays out (in place),
        # streamed through the device in chunks on three streams
        ctx.step_host(par, host["x"], host["y"], host.get("z"), host["vx"], host["vy"], host.get("vz"),
                      host["mass"], host["volume"])

    e2e_step()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = allmax(time.perf_counter() - t0)
    e2e_value = allsum(n_host) * e2e_steps / e2e_s

    # ---- e2e, resident list: what a coupled run does on the steps where nothing on the host
    # touches the particle objects -- the host solver's new U,V,W go up ONCE (rank 0, pinned) and
    # reach the other GPUs over NVLink (gfsb200_broadcast_field), the device runs
    # gfs_particle_list_event (cull + step + BCs) on the resident list, and only the event's
    # result (the number of particles removed) comes back
    restore()

    def resident_step():
        if rank == 0:
            comm.broadcast_field(0, *fields)
        else:
            comm.broadcast_field(0, *[None] * len(fields))
        return ctx.particle_list_event(par)

    resident_step()
    barrier()
    res_steps = max(e2e_steps, 20)
    t0 = time.perf_counter()
    for _ in range(res_steps):
        resident_step()
    barrier()
    res_s = allmax(time.perf_counter() - t0)
    res_value = total_particles * res_steps / res_s
    field_bytes = int(sum(a.nbytes for a in fields))

    # ======================================================================
    # the other configs, short runs on the same context (C3: the non-lattice path -- child0
    # descent + leaf_vtx loads; C5: STRONG scaling, 200 M particles over the N GPUs; 2D: the
    # 80 B/particle-step roofline)
    clocks = sampler.stop(t_region0, t_region1) if rank == 0 else None
    configs = {}
    if not args.no_configs:
        def balance_by_cost(w, step_fn, rounds=2):
            """N > 1, adaptive tree: equal numbers of particles are not equal work (a particle in a deep
            leaf costs more than one in a shallow leaf).  Every rank times its kernel over a few steps,
            the shares of the next gfsb200_comm_rebalance are made inversely proportional to the measured
            time per particle, twice -- outside the timed region, like the re-sort."""
            if world_size == 1:
                return None
            share = np.full(world_size, 1.0 / world_size)
            for _ in range(rounds):
                ctx.timer_sampling(1)
                ctx.timer_reset()
                for i in range(4):
                    step_fn(i)
                if comm_pending[0]:
                    comm.deposit_wait()
                barrier()
                k_ms, _ = ctx.timer_read()
                ctx.timer_sampling(args.kernel_timer_every)
                t = torch.zeros(world_size, device=device, dtype=torch.float64)
                t[rank] = k_ms
                dist.all_reduce(t)
                t = t.cpu().numpy()
                share = share * (t.mean() / t)
                share /= share.sum()
                comm.set_shares(share)
                comm.rebalance()
                barrier()
            return [float(x) for x in share]

        comm_pending = [False]

        def short_run(name, w, n_here, steps, fill, two=False, balance=False):
            ctx.upload_tree(w.tree)
            ctx.upload_field(w.u, w.v, w.w)
            fill()
            order_particles()
            p1 = w.step_params()

            def st(i):
                ctx.refresh_field()
                ctx.step(p1)
            for i in range(3):
                st(i)
            barrier()
            shares = balance_by_cost(w, st) if balance else None
            ctx.timer_reset()
            t_ms, per = timed_steps(st, steps)
            k_ms, _ = ctx.timer_read()
            frac_in = check_inside(name)
            t_ms = allmax(t_ms)
            n_tot = int(allsum(ctx.count))
            bps1 = BYTES_PER_PARTICLE_STEP[w.dim]
            out = {"workload": workload_name(name), "particles_total": n_tot, "particles_this_gpu": ctx.count,
                   "cells": int(w.arrays.n_cells), "leaves": int(w.arrays.n_leaves),
                   "levels": {int(k): int(v) for k, v in w.level_histogram().items()},
                   "steps": steps, "ms_per_step": t_ms / steps, "step_ms": stats_ms(per),
                   "value": n_tot * steps / (t_ms * 1e-3), "unit": "particle-steps/s",
                   "step_kernel_ms": k_ms,
                   "roofline_frac": ctx.count * bps1 / (k_ms * 1e-3) / 1e9 / peak if k_ms > 0 else None,
                   "algorithmic_bytes_per_particle_step": bps1, "particles_inside_at_end": frac_in}
            if two:
                p2 = w.step_params(fuse_deposit=True)

                def st2(i):
                    ctx.refresh_field()
                    ctx.step(p2)
                    comm.deposit_allreduce()
                for i in range(2):
                    st2(i)
                comm.deposit_wait()
                barrier()
                if balance:
                    comm_pending[0] = True
                    shares2 = balance_by_cost(w, st2)
                    comm_pending[0] = False
                ctx.timer_reset()
                comm.exchange_stats()
                t_ms2, per2 = timed_steps(st2, steps, tail=comm.deposit_wait)
                k_ms2, _ = ctx.timer_read()
                x_ms2, _, x_b2 = comm.exchange_stats()
                t_ms2 = allmax(t_ms2)
                out["two_way"] = {"ms_per_step": t_ms2 / steps, "value": n_tot * steps / (t_ms2 * 1e-3),
                                  "kernel_ms": k_ms2, "exchange_device_ms": x_ms2,
                                  "bytes_sent_per_rank": int(x_b2),
                                  "roofline_frac": ctx.count * 136 / (k_ms2 * 1e-3) / 1e9 / peak if k_ms2 > 0 else None}
                if balance and shares2:
                    out["two_way"]["shares"] = shares2
            if balance and shares:
                out["shares"] = shares
                out["sharding"] = ("slices of the depth-first leaf order; shares inversely proportional to the "
                                   "measured kernel time per particle (gfsb200_comm_set_shares), set outside the timed region")
                comm.set_shares(None)
            return out

        w3 = worlds.make_c3(n_particles=10_000_000)
        if world_size == 1:
            p3 = worlds.make_particles(w3)
            configs["C3"] = short_run("C3", w3, len(p3["x"]), 20, lambda: ctx.particles_upload(**p3), two=True)
            del p3
        w5 = worlds.make_c5()
        n5 = args.c5_particles // world_size
        cols5 = device_cloud(torch, w5, n5, rank, device)
        configs["C5"] = short_run("C5", w5, n5, 10, lambda: fill_from_device(torch, ctx, cols5, device), two=True,
                                  balance=world_size > 1)
        configs["C5"]["scaling"] = "strong: %d particles in total, 1/N per GPU; cloud drawn on the device" % (n5 * world_size)
        del cols5
        torch.cuda.empty_cache()
        if world_size == 1:
            w2d = worlds.make_c1(level=10, n_particles=10_000_000)
            p2d = worlds.make_particles(w2d)
            configs["2D"] = short_run("C1", w2d, len(p2d["x"]), 20, lambda: ctx.particles_upload(**p2d))
            configs["2D"]["workload"] = "2D lid-style level-10 quadtree (1024^2, four ghost layers), 10M particles, drag"
            del p2d

    if rank == 0:
        bps = BYTES_PER_PARTICLE_STEP[world.dim]
        achieved = n_local * bps / (kernel_ms * 1e-3) / 1e9 if kernel_ms > 0 else 0.0
        line = {
            "metric": "particle-steps/sec", "value": value, "unit": "particle-steps/s",
            "n_gpus": world_size, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "step_ms": stats_ms(per_step),
            "config": {"workload": workload_name(args.config), "particles_per_gpu": n_local,
                       "particles_total": total_particles,
                       "cells": int(world.arrays.n_cells), "leaves": int(world.arrays.n_leaves),
                       "vertices": int(world.arrays.n_vertices), "two_way": False,
                       "sharding": ("one GPU: sorted by cell" if world_size == 1 else
                                    "contiguous slices of the globally cell-sorted cloud (gfsb200_comm_rebalance)"),
                       "resort_every": args.resort, "sorts_in_timed_region": n_sorts,
                       "timed_segments": n_segments,
                       "cell_pass_every_step": True, "particles_inside_at_end": inside_frac,
                       "l2": "per-step particle stream (%.0f MB) exceeds the 126 MB L2" % (n_local * bps / 1e6)},
            "roofline": {"bound": "hbm", "kernel": "%s<%d,...> (TMA-staged fused locate+interpolate+force+integrate)" % ("step_kernel_wpipe" if not os.environ.get("GFSB200_STEP_MODE") else "step_kernel (GFSB200_STEP_MODE=%s)" % os.environ["GFSB200_STEP_MODE"], world.dim),
                         "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "frac_of_8TBs_spec": achieved / 8000.0,
                         "traffic": measured_traffic(args.config, n_local, world.dim), "peak_source": peak_src,
                         "algorithmic_bytes_per_particle_step": bps,
                         "kernel_ms": kernel_ms, "kernel_launches": kernel_launches,
                         "kernel_timing": "CUDA events inside the library around every %s launch of the timed "
                                          "region (%d of its %d step-kernel launches)" % (
                                              "" if args.kernel_timer_every == 1 else "%dth" % args.kernel_timer_every,
                                              kernel_launches, args.steps),
                         "whole_step_frac": n_local * bps / (ms / args.steps * 1e-3) / 1e9 / peak},
            "e2e": {"value": e2e_value, "unit": "particle-steps/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "what": "per step: gfsb200_upload_field (U,V,W from pinned host memory, cell pass) + "
                            "gfsb200_step_host (8 particle columns H2D, fused step, 6 columns D2H, chunked "
                            "on three streams); wall clock"},
            "e2e_resident": {"value": res_value, "unit": "particle-steps/s",
                             "h2d_bytes_per_step": field_bytes, "h2d_on": "rank 0 only",
                             "nvlink_bytes_per_step_per_gpu": field_bytes if world_size > 1 else 0,
                             "d2h_bytes_per_step": 8, "steps": res_steps,
                             "what": "per step: gfsb200_broadcast_field (U,V,W from rank 0's pinned host memory "
                                     "over PCIe once, ncclBroadcast over NVLink to the other GPUs, cell pass) + "
                                     "gfsb200_particle_list_event on the device-resident list (cull, fused step, "
                                     "particle BCs), returning the removed count; wall clock.  The full-sync "
                                     "`e2e` above is the headline; this is the same API with the list left on "
                                     "the device between steps"},
            "two_way": two_way,
            "configs": configs,
            "gpu_launches": int(gpu_launches),
            "clocks": clocks,
        }
        if not args.no_cpu_baseline and world_size == 1:
            line["cpu_baseline"] = cpu_baseline(args, worlds)
        print(json.dumps(line), flush=True)
    comm.close()
    ctx.close()
    if world_size > 1:
        dist.destroy_process_group()


def cpu_baseline(args, worlds):
    """The reference's own object code on one host core, bounded sample."""
    ora = entry.load_oracle()
    sp, sim = oracle_world(ora, worlds, args.config)
    mk = oracle_params_factory(ora)
    rate, _, _ = time_refobj(ora, worlds, mk, sp, sim, 20_000, 1, 0)
    n = int(min(sp.n_particles, max(20_000, rate * 5.0)))          # ~5 s per step
    value, dt, left = time_refobj(ora, worlds, mk, sp, sim, n, 3, 0)
    value_port, _ = time_oracle(ora, worlds, mk, sp, sim, n, 3, 0, 1)
    threads = ora.load(3).ora_max_threads()
    value_mt, _ = time_oracle(ora, worlds, mk, sp, sim, n, 3, 1, threads)
    value_fused, _ = time_oracle(ora, worlds, mk, sp, sim, n, 3, 0, 1, pattern=1)
    return {"value": value, "unit": "particle-steps/s", "cores": 1, "kind": "reference",
            "sample": f"{n} particles x 3 steps of the {args.config} cloud on the full tree; the reference's "
                      "own object code (modules/particulatecommon.c, src/event.c, src/particle.c, src/fluid.c, "
                      "src/ftt.c compiled unmodified): gfs_event_do on a GfsParticleList, serial as the "
                      "reference is per MPI rank",
            "port_1core": {"value": value_port, "note": "restated oracle (bit-identical results), same call "
                                                        "pattern without the GtsObject/event layer"},
            "port_all_cores": {"value": value_mt, "cores": threads, "note": "restated oracle, OpenMP over particles"},
            "port_fused_1core": {"value": value_fused, "note": "one locate + one interpolation set per particle"},
            "host_cpus": os.cpu_count()}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="C2")
    ap.add_argument("--particles", type=int, default=0, help="particles per GPU (default: the config's)")
    ap.add_argument("--no-two-way", action="store_true", help="skip the two-way block")
    ap.add_argument("--no-configs", action="store_true", help="skip the C3 / C5 / 2D lines")
    ap.add_argument("--c5-particles", type=int, default=200_000_000, help="total particles of the C5 strong-scaling line")
    ap.add_argument("--level", type=int, default=0, help="experiment: override the C2 tree level")
    ap.add_argument("--resort", type=int, default=100,
                    help="re-sort particles by cell every R steps (0: never); particles cross a cell "
                         "every ~15 steps at most in these configs and the step kernel slows by ~1 %% "
                         "over 100 steps without a re-sort, while one re-sort costs ~0.6 ms")
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--kernel-timer-every", type=int, default=4,
                    help="CUDA events around every n-th step-kernel launch (1: every launch)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
