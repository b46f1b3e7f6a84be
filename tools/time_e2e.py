#!/usr/bin/env python
"""e2e (host buffers in/out) step time versus the streaming chunk size: python tools/time_e2e.py"""
import sys, os, time, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds
import torch
COLS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")
w = worlds.make_c2()
ctx = capi.Context(0)
ctx.upload_tree(w.tree)
ctx.upload_field(w.u, w.v, w.w)
parts = worlds.make_particles(w)
host = {k: torch.from_numpy(np.ascontiguousarray(parts[k])).pin_memory().numpy() for k in COLS}
fields = [torch.from_numpy(np.ascontiguousarray(f)).pin_memory().numpy() for f in (w.u, w.v, w.w)]
par = w.step_params()
n = len(parts["x"])
for chunk in [1 << 20, 1 << 19, 1 << 18, 1 << 17, 1 << 16, 1 << 21]:
    for rep in range(2):
        ctx.synchronize()
        t0 = time.perf_counter()
        for _ in range(3):
            ctx.upload_field(*fields)
            ctx.step_host(par, host["x"], host["y"], host["z"], host["vx"], host["vy"], host["vz"],
                          host["mass"], host["volume"], chunk=chunk)
        ctx.synchronize()
        dt = (time.perf_counter() - t0) / 3
    print(json.dumps({"chunk": chunk, "ms_per_step": dt * 1e3, "particle_steps_per_s": n / dt}))
