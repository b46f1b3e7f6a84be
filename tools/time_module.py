#!/usr/bin/env python
"""Times one GfsParticleList event through the drop-in GModule (libgfsrefmod: the module
source linked with the reference's object code) against the same event in the unmodified
reference (libgfsrefobj), on the C2 tree.  usage: time_module.py [n_particles] [steps]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                      # noqa: E402
import __graft_entry__ as entry   # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
config = sys.argv[3] if len(sys.argv) > 3 else "C2"
pkg = entry.load_package()
ora = entry.load_oracle()
worlds = pkg.worlds
sp, sim = bench.oracle_world(ora, worlds, config)
mk = bench.oracle_params_factory(ora)
par = mk(sp, 0)
parts = worlds.make_particles(sp, n)
for module in (True, False):
    m = n if module else min(n, 100_000)
    rs = ora.RefSim(sim, module=module)
    rs.configure(par)
    t0 = time.perf_counter()
    rl = ora.RefParticleList(rs, *[parts[k][:m] for k in bench.COLS], par)
    t_build = time.perf_counter() - t0
    rl.event(1)                                   # first event: flatten + stencils + upload
    t0 = time.perf_counter()
    rl.event(steps)
    dt = time.perf_counter() - t0
    print(f"{'module' if module else 'reference'}: {m} particles, list built in {t_build:.2f} s, "
          f"{1e3 * dt / steps:.2f} ms/event, {m * steps / dt:.3e} particle-steps/s, {len(rl)} left", flush=True)
    rs.close()
