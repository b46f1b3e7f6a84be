#!/usr/bin/env python
"""Times the per-field-update cell pass (vertex + vorticity tables) alone:
python tools/time_cellpass.py [C2|C3]"""
import sys, os, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds

cfg = sys.argv[1] if len(sys.argv) > 1 else "C2"
w = worlds.make_c2(n_particles=1000) if cfg == "C2" else worlds.make_c3(n_particles=1000)
ctx = capi.Context(0)
ctx.upload_tree(w.tree)
ctx.upload_field(w.u, w.v, w.w)
for _ in range(int(os.environ.get("CELLPASS_WARM", "20"))):
    ctx.refresh_field()
ctx.synchronize()
n = int(os.environ.get("CELLPASS_ITERS", "300"))
t0 = time.perf_counter()
for _ in range(n):
    ctx.refresh_field()
ctx.synchronize()
dt = (time.perf_counter() - t0) / n
a = w.arrays
bytes_min = a.n_cells * 24 + a.n_vertices * 32 + a.n_leaves * 32
print(json.dumps({"config": cfg, "cells": int(a.n_cells), "vertices": int(a.n_vertices),
                  "cell_pass_ms": dt * 1e3, "min_bytes": int(bytes_min),
                  "GBps_of_min_bytes": bytes_min / dt / 1e9}))
