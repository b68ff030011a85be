import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from shine_mapping_b200 import Decoder, FeatureOctree, SdfTrainer, synth
dev = torch.device("cuda", 0); cfg = bench.workload_config(str(dev)); torch.manual_seed(42)
octree, decoder = FeatureOctree(cfg), Decoder(cfg)
pool = synth.build_scene_map(cfg, octree, n_azimuth=2048, n_frames=1, seed=42, device=str(dev))
n = len(pool); c, l, w = pool.get_batch(n)
ch, lh = c.cpu().pin_memory(), l.cpu().pin_memory()
tr = SdfTrainer(cfg, octree, decoder)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for chunks, graph in ((2, False), (2, True), (4, True), (8, True), (16, True)):
    for _ in range(3): tr.step_from_host(ch, lh, chunks=chunks, use_graph=graph)
    ts = []
    for k in range(10):
        flush.fill_(k); torch.cuda.synchronize(); t0 = time.perf_counter(); tr.step_from_host(ch, lh, chunks=chunks, use_graph=graph); ts.append(time.perf_counter() - t0)
    ts.sort(); print(f"chunks {chunks} graph={graph}: median {ts[5]*1e3:.3f} ms  min {ts[0]*1e3:.3f} ms  -> {n/ts[5]/1e9:.2f} Gpts/s")
# raw H2D
torch.cuda.synchronize(); d = torch.empty_like(c)
t0 = time.perf_counter()
for _ in range(20): d.copy_(ch, non_blocking=True)
torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / 20
print(f"raw H2D {ch.numel()*4/1e6:.1f} MB in {dt*1e3:.3f} ms = {ch.numel()*4/dt/1e9:.1f} GB/s")
