#!/usr/bin/env python
"""Pinned host -> device copy rate of this box as a function of where the pinned buffer was first touched
(development aid for the e2e leg): unpinned process, process bound to the GPU's NUMA node, bound to another node."""
import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from shine_mapping_b200 import dist as sdist

dev = torch.device("cuda", 0); torch.cuda.set_device(dev)
node = sdist.gpu_numa_node(0)
all_cpus = sorted(os.sched_getaffinity(0))
print("gpu numa node", node, "cpus", len(all_cpus))


def node_cpus(n):
    try:
        return sorted(sdist._parse_cpulist(open(f"/sys/devices/system/node/node{n}/cpulist").read()) & set(all_cpus))
    except Exception:
        return []


def rate(tag, mb=12.4, nbuf=14, reps=40):
    n = int(mb * 1e6 / 4)
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    for h in host: h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    for i in range(5): d.copy_(host[i % nbuf], non_blocking=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(reps): d.copy_(host[i % nbuf], non_blocking=True)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{tag:28s} {nbuf:3d} x {mb:.1f} MB buffers: {ms:.3f} ms/copy = {mb / ms:.1f} GB/s")


def rate_split(tag, nstreams, mb=12.4, nbuf=14, reps=40):
    n = int(mb * 1e6 / 4)
    host = [torch.empty(n, dtype=torch.float32).pin_memory() for _ in range(nbuf)]
    for h in host: h.fill_(1.0)
    d = torch.empty(n, dtype=torch.float32, device=dev)
    streams = [torch.cuda.Stream(dev) for _ in range(nstreams)]
    cuts = [n * i // nstreams for i in range(nstreams + 1)]

    def one(i):
        h = host[i % nbuf]
        for k, st in enumerate(streams):
            with torch.cuda.stream(st):
                d[cuts[k]:cuts[k + 1]].copy_(h[cuts[k]:cuts[k + 1]], non_blocking=True)
    for i in range(5): one(i)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(reps): one(i)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) / reps * 1e3
    print(f"{tag:28s} split over {nstreams} streams: {ms:.3f} ms/copy = {mb / ms:.1f} GB/s")


rate("as launched", nbuf=14); rate("as launched", nbuf=4)
for k in (1, 2, 4): rate_split("as launched", k)
nodes = [int(x[4:]) for x in os.listdir("/sys/devices/system/node") if x.startswith("node") and x[4:].isdigit()]
for n in sorted(nodes):
    cpus = node_cpus(n)
    if not cpus: continue
    os.sched_setaffinity(0, cpus)
    rate(f"bound to node {n}{' (GPU)' if n == node else ''}", nbuf=14)
os.sched_setaffinity(0, all_cpus)
