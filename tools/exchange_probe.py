#!/usr/bin/env python
"""Times gfsb200_deposit_allreduce on the C2 field, one process per GPU (torch.distributed.run):
alone (nothing else on the GPU) and behind the fused step + deposit kernel of the next step, in the
owner-slice mode and (GFSB200_EXCHANGE=1) as ncclAllReduce.  Rank 0 prints one JSON line."""
import json, os, sys, time
import numpy as np
import torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry

rank, size, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
pkg = entry.load_package()
capi, worlds = pkg.capi, pkg.worlds
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
cfg = sys.argv[2] if len(sys.argv) > 2 else "C2"
w = worlds.make_c2(n_particles=n) if cfg == "C2" else worlds.make_c5(n_particles=n)
ctx = capi.Context(local)
ctx.upload_tree(w.tree)
ctx.upload_field(w.u, w.v, w.w)
uid = [capi.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
comm = capi.Comm.init_rank(ctx, uid[0], rank, size)
parts = worlds.make_particles(w, n * size, rank, size)
ctx.particles_upload(**parts)
comm.rebalance()
par, parf = w.step_params(), w.step_params(fuse_deposit=True)
out = {"ranks": size, "config": cfg, "peer_access": comm.peer_access, "mode": os.environ.get("GFSB200_EXCHANGE", "0"),
       "local_red_gpu_scope": bool(os.environ.get("GFSB200_LOCAL_RED_GPU_SCOPE")),
       "particles_per_rank": ctx.count}
# (a) the exchange alone
for _ in range(3):
    ctx.deposit_volume(); comm.deposit_allreduce(); comm.deposit_wait(); ctx.synchronize()
comm.exchange_stats()
dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    ctx.deposit_volume(); comm.deposit_allreduce(); comm.deposit_wait(); ctx.synchronize()
wall = (time.perf_counter() - t0) / 20
ms, cnt, sent = comm.exchange_stats()
out["alone"] = {"transfer_ms": ms, "bytes_sent_per_rank": sent, "GBs": sent / (ms * 1e-3) / 1e9 if ms > 0 else None,
                "wall_ms_per_iteration_incl_deposit_volume": wall * 1e3}
# (b) behind the next step's kernels
for _ in range(3):
    ctx.refresh_field(); ctx.step(parf); comm.deposit_allreduce()
comm.deposit_wait(); ctx.synchronize(); comm.exchange_stats(); ctx.timer_reset(); dist.barrier()
t0 = time.perf_counter()
for _ in range(20):
    ctx.refresh_field(); ctx.step(parf); comm.deposit_allreduce()
comm.deposit_wait(); ctx.synchronize()
wall = (time.perf_counter() - t0) / 20
ms, cnt, sent = comm.exchange_stats()
kms, _ = ctx.timer_read()
now = ctx.particles_download()
cc = ctx.locate(now["x"], now["y"], now["z"])
owner = comm.owner_table(w.arrays.n_cells)
out["fused_kernel_ms"] = kms
out["drifters_fraction"] = float(np.mean((cc >= 0) & (owner[np.maximum(cc, 0)] != rank)))
out["overlapped"] = {"transfer_ms": ms, "GBs": sent / (ms * 1e-3) / 1e9 if ms > 0 else None, "wall_ms_per_step": wall * 1e3}
res = [None] * size
dist.all_gather_object(res, out)
if rank == 0:
    out["alone_transfer_ms_all_ranks"] = [round(r["alone"]["transfer_ms"], 4) for r in res]
    out["overlapped_transfer_ms_all_ranks"] = [round(r["overlapped"]["transfer_ms"], 4) for r in res]
    print(json.dumps(out), flush=True)
comm.close(); ctx.close()
dist.destroy_process_group()
