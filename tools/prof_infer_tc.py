import os, sys, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import bench
from shine_mapping_b200 import Decoder, FeatureOctree, sdf_infer, synth
dev = torch.device("cuda", 0); cfg = bench.workload_config(str(dev)); torch.manual_seed(42)
octree, decoder = FeatureOctree(cfg), Decoder(cfg)
pool = synth.build_scene_map(cfg, octree, n_azimuth=2048, n_frames=1, seed=42, device=str(dev))
coord, _, _ = pool.get_batch(len(pool))
for _ in range(3): sdf_infer(octree, decoder, coord, tcgen05=True)
torch.cuda.synchronize(); torch.cuda.profiler.start()
sdf_infer(octree, decoder, coord, tcgen05=True); sdf_infer(octree, decoder, coord)
torch.cuda.synchronize(); torch.cuda.profiler.stop()
