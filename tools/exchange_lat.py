#!/usr/bin/env python
"""Latency of the step's exchange alone (no step kernel in between): NCCL path (pack + ncclAllReduce through the C ABI +
unpack) vs the one-kernel NVLink peer-memory path, on the partitioned bench workload.  torchrun, N ranks."""
import os, sys, statistics, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from shine_mapping_b200 import SdfTrainer, dist as sdist
rank, world, local = sdist.init_from_env("nccl")
dev = torch.device("cuda", local); torch.cuda.set_device(dev)
cfg, octree, decoder, pool, plan, comm, p2p, info = bench.build_partitioned_workload(str(dev), rank, world, 1024, exchange="p2p")
out = {}
for name, kw in (("nccl", dict(comm=comm)), ("p2p", dict(comm=comm, p2p=p2p))):
    tr = SdfTrainer(cfg, octree, decoder, shard_mode="spatial", boundary=plan, **kw)
    tr.zero_grad()
    for _ in range(20): tr.all_reduce_grads()
    torch.cuda.synchronize(); sdist.barrier(dev)
    reps = 200
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): tr.all_reduce_grads()
    e1.record(); torch.cuda.synchronize()
    out[name] = sdist.max_over_ranks(e0.elapsed_time(e1) / reps * 1e3, dev)
if rank == 0:
    print(f"world {world}: exchange of {plan.total_floats} floats back to back: nccl path {out['nccl']:.1f} us, p2p kernel {out['p2p']:.1f} us"
          f" (includes the host's per-call cost when it is the limiter); p2p timeouts {p2p.timeouts()}")
torch.distributed.destroy_process_group()
