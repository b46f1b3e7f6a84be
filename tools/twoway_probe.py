#!/usr/bin/env python
"""Times the step kernel alone and the fused step + deposit kernel on one config, for A/B runs of
library builds (GFSB200_LIB=...) and for ncu captures:
  python tools/twoway_probe.py C2|C3|C5 [steps] [n_particles]
Prints one JSON line: kernel ms of both flavours (CUDA events inside the library), HBM fractions
(112 / 136 B per particle-step over MEASURED_PEAKS hbm_gbs) and a checksum of state + field."""
import sys, os, json, hashlib
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds

cfg = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 40
npart = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000_000
w = {"C2": worlds.make_c2, "C3": worlds.make_c3, "C5": worlds.make_c5,
     "2D": lambda n_particles: worlds.make_c1(level=10, n_particles=n_particles)}[cfg](n_particles=npart)
parts = worlds.make_particles(w, npart)
try:
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    peak = 6533.8
ctx = capi.Context(0)
ctx.upload_tree(w.tree)
ctx.upload_field(w.u, w.v, w.w)
out = {"config": cfg, "lib": os.environ.get("GFSB200_LIB", "default"), "particles": npart}
b1, b2 = (112, 136) if w.dim == 3 else (80, 96)
for name, par, b in (("step", w.step_params(), b1), ("fused", w.step_params(fuse_deposit=True), b2)):
    ctx.particles_upload(**parts)
    ctx.sort()
    for _ in range(5):
        ctx.refresh_field(); ctx.step(par)
    ctx.synchronize()
    ctx.timer_reset()
    for _ in range(steps):
        ctx.refresh_field(); ctx.step(par)
    ctx.synchronize()
    ms, n = ctx.timer_read()
    out[name + "_ms"] = round(ms, 5)
    out[name + "_frac"] = round(b * npart / (ms * 1e-3) / 1e9 / peak, 4)
got = ctx.particles_download()
h = hashlib.sha1(b"".join(np.ascontiguousarray(got[k]).tobytes() for k in ("x", "y", "z", "vx", "vy", "vz") if got[k] is not None))
f = [ctx.download_deposit(c) for c in range(1 + w.dim)]
out["state_sha1"] = h.hexdigest()[:12]
out["field_abs_sum"] = [float(np.abs(a).sum()) for a in f]
print(json.dumps(out), flush=True)
