#!/usr/bin/env python
"""A/B of the fused step kernel's launch modes (GFSB200_STEP_MODE) on one world:
python tools/ab_step_modes.py C2|C3 mode[,mode...] [steps]
Prints, per mode, the mean step-kernel time (CUDA events inside the library), the HBM
fraction (112 B per particle-step over MEASURED_PEAKS hbm_gbs) and a checksum of the state
after the run (must be identical across modes: same arithmetic, different staging)."""
import sys, os, json, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds

cfg = sys.argv[1]
modes = [int(m) for m in sys.argv[2].split(",")]
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 60
w = worlds.make_c2() if cfg == "C2" else worlds.make_c3()
parts = worlds.make_particles(w, w.n_particles)
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
par = w.step_params()
first = None
for mode in modes:
    os.environ["GFSB200_STEP_MODE"] = str(mode)
    ctx = capi.Context(0)
    ctx.upload_tree(w.tree)
    ctx.upload_field(w.u, w.v, w.w)
    ctx.particles_upload(**parts)
    ctx.sort()
    for _ in range(10):
        ctx.refresh_field(); ctx.step(par)
    ctx.synchronize()
    ctx.timer_reset()
    for _ in range(steps):
        ctx.refresh_field(); ctx.step(par)
    ctx.synchronize()
    ms, n = ctx.timer_read()
    got = ctx.particles_download()
    h = hashlib.sha1(b"".join(np.ascontiguousarray(got[k]).tobytes() for k in ("x", "y", "z", "vx", "vy", "vz"))).hexdigest()[:12]
    if first is None:
        first = got
    dmax = max(float(np.max(np.abs(got[c] - first[c])) / np.max(np.abs(first[c]))) for c in ("x", "y", "z", "vx", "vy", "vz"))
    k = ms          # timer_read returns the mean over its n launches
    print(json.dumps({"config": cfg, "mode": mode, "kernel_ms": round(k, 5),
                      "frac": round(112 * len(parts["x"]) / (k * 1e-3) / 1e9 / peak, 4), "state_sha1": h, "max_diff_vs_first_mode": dmax}), flush=True)
    del ctx
