#!/usr/bin/env python
"""Times the force-recording flavour of the step kernel (what the GModule launches):
python tools/rec_probe.py C2|C3 [steps]"""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import __graft_entry__ as entry
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds
cfg = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
w = {"C2": worlds.make_c2, "C3": worlds.make_c3}[cfg]()
parts = worlds.make_particles(w, w.n_particles)
ctx = capi.Context(0)
ctx.upload_tree(w.tree)
ctx.upload_field(w.u, w.v, w.w)
out = {"config": cfg, "lib": os.environ.get("GFSB200_LIB", "default")}
for name, kw in (("plain", {}), ("rec_forces", {"record_forces": True}), ("rec_forces_cells", {"record_forces": True, "record_cells": True}),
                 ("rec_forces_escapes", {"record_forces": True, "track_escapes": True})):
    par = w.step_params(**kw)
    ctx.particles_upload(**parts)
    ctx.sort()
    for _ in range(3):
        ctx.refresh_field(); ctx.step(par)
    ctx.synchronize(); ctx.timer_reset()
    for _ in range(steps):
        ctx.refresh_field(); ctx.step(par)
    ctx.synchronize()
    ms, n = ctx.timer_read()
    out[name + "_ms"] = round(ms, 5)
print(json.dumps(out), flush=True)
