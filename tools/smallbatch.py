import os, sys, time, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from shine_mapping_b200 import Decoder, FeatureOctree, SdfTrainer, synth
from shine_mapping_b200.batch_loop import _GraphedIteration
dev = torch.device("cuda", 0)
cfg = bench.workload_config(str(dev)); cfg.bs = int(os.environ.get("BS", "4096"))
torch.manual_seed(42)
octree, decoder = FeatureOctree(cfg), Decoder(cfg)
pool = synth.build_scene_map(cfg, octree, n_azimuth=512, n_frames=1, seed=42, device=str(dev))
tr = SdfTrainer(cfg, octree, decoder)
def t(fn, name, n=200):
    for _ in range(10): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name:36s} gpu {e0.elapsed_time(e1)/n*1e3:8.1f} us   wall {(time.perf_counter()-t0)/n*1e6:8.1f} us")
batch = pool.get_batch(cfg.bs)
t(lambda: pool.get_batch(cfg.bs), "get_batch")
t(lambda: tr.forward_backward(*batch), "forward_backward (incl loss.zero_)")
t(lambda: tr.optimizer_step(True), "adam host-step")
t(lambda: tr.optimizer_step(True, device_step=True), "adam device-step (bump+adam)")
def eager():
    b = pool.get_batch(cfg.bs); tr.forward_backward(*b); tr.optimizer_step(True)
t(eager, "eager iteration")
g = _GraphedIteration(tr, pool, cfg.bs); g.run()
t(g.run, "graph replay iteration")
