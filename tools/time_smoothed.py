#!/usr/bin/env python
"""Times gfsb200_deposit_force_smoothed (GfsSourceParticulate with its kernel) on
a C2/C3-style world: python tools/time_smoothed.py [C2|C3] [n_particles] [rkernel/h]"""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds
import torch

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
n = int(float(sys.argv[2])) if len(sys.argv) > 2 else 1_000_000
rk_h = [float(v) for v in sys.argv[3].split(",")] if len(sys.argv) > 3 else [1.0, 2.0]
w = worlds.make_c2(n_particles=n) if cfg == "C2" else worlds.make_c3(n_particles=n)
ctx = capi.Context(0)
ctx.upload_tree(w.tree)
ctx.upload_field(w.u, w.v, w.w)
parts = worlds.make_particles(w)
ctx.particles_upload(**parts)
ctx.sort()
hmin = 0.5 ** w.arrays.max_level
par = w.step_params()
for r in rk_h:
    rk = r * hmin
    for rep in range(2):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ctx.deposit_force_smoothed(par, rk, capi.KERNEL_GAUSSIAN, 1.0, 1e-5 / (r * r + 0.1))
        ctx.synchronize()
        dt = time.perf_counter() - t0
    corr = None
    ctx.deposit_force_smoothed(par, rk, capi.KERNEL_GAUSSIAN, 1.0, 1e-5 / (r * r + 0.1), record_norm=True)
    corr, vol = ctx.download_kernel_norm()
    leaves = vol / hmin ** 3
    print(json.dumps({"config": cfg, "particles": n, "rkernel_over_hmin": r, "ms": dt * 1e3,
                      "particles_per_s": n / dt, "mean_cells_per_particle_volume_units": float(leaves.mean())}))

# the oracle (reference traversal object code + restated callbacks) on a small sample, one core
if os.environ.get("SMOOTHED_ORACLE", "1") == "1" and cfg == "C2":
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import helpers
    ora = helpers.ora
    w2 = worlds.make_c2(level=6, n_particles=2000)          # 64^3: the oracle tree build stays short
    sim, ptrs = helpers.matched_oracle(w2)
    p2 = worlds.make_particles(w2)
    plist = ora.ParticleList(sim, *[p2[k] for k in ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")])
    h6 = 0.5 ** 6
    for r in rk_h:
        t0 = time.perf_counter()
        plist.deposit_force_smoothed(helpers.oracle_params(w2), 4, r * h6, ora.Kernel(ora.KERNEL_GAUSSIAN, 1.0, 1e-5, 1, 0))
        dt = time.perf_counter() - t0
        print(json.dumps({"oracle_cpu_1core": True, "tree": "64^3", "particles": 2000, "rkernel_over_hmin": r,
                          "particles_per_s": 2000 / dt}))
