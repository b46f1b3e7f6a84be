"""Generates the golden fixtures of tests/golden/ by running the ORACLE
(reference ftt.c/fluid.c object code + restated particulate layer) in this
container:   python tests/golden/make_golden.py

Each fixture is a small .npz holding seeded inputs and the oracle's outputs
for one world: located cell (as level + exact centre, which is independent of
any cell numbering), interpolated velocity, per-leaf vorticity and corner
values at sample leaves, and the particle state after 1 and 10 steps.

The reference itself ships no golden vector for this path (SURVEY.md section
4); these pin (a) the oracle against regressions of the restatement/shim and
of the prebuilt oracle/_ref/*.so that travels to the GPU box, and (b) the
CUDA path against numbers that were produced here, where /root/reference
exists.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import helpers  # noqa: E402
from helpers import capi, worlds, ora  # noqa: E402

COLS = ("x", "y", "z", "vx", "vy", "vz", "mass", "volume")


def cases():
    yield "c1_l5", worlds.make_c1(level=5, n_particles=400)
    yield "tg_l4", worlds.make_c2(level=4, n_particles=400)
    yield "ring_3_6", worlds.make_ring("ring", 3, 6, 400, 3003)


def build(name, w):
    sim, ptrs = helpers.matched_oracle(w)
    idx = helpers.PtrIndex(ptrs)
    a = w.arrays
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    rng = np.random.default_rng({"c1_l5": 11, "tg_l4": 12, "ring_3_6": 13}[name])
    pts = worlds.adversarial_points(a, rng, 300)
    rnd = [rng.uniform(-0.55, 0.55, 300) for _ in range(w.dim)] + ([None] if w.dim == 2 else [])
    qx = np.concatenate([pts[0], rnd[0]]); qy = np.concatenate([pts[1], rnd[1]])
    qz = None if w.dim == 2 else np.concatenate([pts[2], rnd[2]])
    cell = idx(sim.locate(qx, qy, qz))
    out = dict(dim=w.dim, qx=qx, qy=qy, qz=np.zeros(0) if qz is None else qz,
               loc_level=np.where(cell >= 0, a.level[np.maximum(cell, 0)], -1).astype(np.int32),
               loc_pos=np.where((cell >= 0)[:, None], a.pos[np.maximum(cell, 0)], np.nan))
    for c in range(w.dim):
        out[f"interp{c}"] = sim.interpolate(c, qx, qy, qz)
    leaves = a.box_leaves
    pick = leaves[rng.choice(len(leaves), min(200, len(leaves)), replace=False)]
    out["leaf_level"] = a.level[pick].astype(np.int32)
    out["leaf_pos"] = a.pos[pick]
    out["vort"] = sim.vorticity(ptrs[pick])
    for c in range(w.dim):
        out[f"corner{c}"] = sim.corner_values(c, ptrs[pick])
    parts = worlds.make_particles(w)
    for k in COLS:
        out["p_" + k] = np.zeros(0) if parts[k] is None else parts[k]
    for steps in (1, 10):
        cells, st = helpers.oracle_step(sim, ptrs, w, parts, steps=steps)
        for k in ("x", "y", "z", "vx", "vy", "vz", "fx", "fy", "fz"):
            out[f"s{steps}_{k}"] = np.zeros(0) if st[k] is None else st[k]
        out[f"s{steps}_cell_level"] = a.level[cells].astype(np.int32)
        out[f"s{steps}_cell_pos"] = a.pos[cells]
    out["n_cells"] = a.n_cells
    out["n_leaves"] = a.n_leaves
    out["n_vertices"] = a.n_vertices
    return out


if __name__ == "__main__":
    for name, w in cases():
        data = build(name, w)
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **data)
        print(name, os.path.getsize(path), "bytes")
