"""Synthetic LiDAR data for the hot path (there is no dataset on the GPU box).

* `raycast_scene`   — analytic scene (ground plane, two walls, axis-aligned boxes) hit by an HDL-64-like
                      scan pattern from a sensor origin -> hit points in metres.
* `sample_rays`     — the training-sample contract of the reference's `dataSampler.sample`
                      (utils/data_sampler.py:18-139): per hit, `surface_sample_n` samples uniformly within
                      +-surface_sample_range of the hit and `free_sample_n` samples in free space; label = signed
                      displacement along the ray in scaled units (positive behind the surface, :109-111); weight
                      +1 for surface samples, -1 for free-space samples (:104); ray-wise output order (:123-134).
* `SamplePool`      — the device sample pools and `get_batch()` of `LiDARDataset`
                      (dataset/lidar_dataset.py:104-113,431-448): `torch.randint` gather.
* `build_scene_map` — scans -> samples -> `octree.update(surface samples)` (dataset/lidar_dataset.py:204-218).
"""
from __future__ import annotations

import math

import torch

from .config import SHINEConfig


def lidar_directions(n_azimuth: int, n_elev: int = 64, elev_min_deg: float = -24.8, elev_max_deg: float = 2.0,
                     device="cpu") -> torch.Tensor:
    """Unit ray directions [n_elev * n_azimuth, 3] of an HDL-64-like spinning LiDAR."""
    el = torch.linspace(math.radians(elev_min_deg), math.radians(elev_max_deg), n_elev, device=device)
    az = torch.arange(n_azimuth, device=device, dtype=torch.float32) * (2 * math.pi / n_azimuth)
    el, az = torch.meshgrid(el, az, indexing="ij")
    d = torch.stack((torch.cos(el) * torch.cos(az), torch.cos(el) * torch.sin(az), torch.sin(el)), -1)
    return d.reshape(-1, 3)


def default_boxes(device="cpu") -> torch.Tensor:
    """[K, 6] axis-aligned boxes (xmin, ymin, zmin, xmax, ymax, zmax) in metres: parked-car / kiosk sized."""
    return torch.tensor([
        [6.0, -5.5, -1.7, 10.0, -3.7, -0.2], [14.0, 3.5, -1.7, 18.5, 5.4, -0.1], [-9.0, -6.0, -1.7, -5.0, -4.2, 0.0],
        [24.0, -6.5, -1.7, 27.0, -4.0, 1.0], [-20.0, 4.0, -1.7, -16.0, 6.0, -0.2], [35.0, 2.0, -1.7, 38.0, 6.5, 1.5],
        [48.0, -6.0, -1.7, 52.0, -3.8, -0.3], [62.0, 3.0, -1.7, 66.0, 5.0, 0.2], [77.0, -6.8, -1.7, 80.0, -4.4, 0.8],
        [91.0, 3.6, -1.7, 95.0, 5.6, -0.2]], dtype=torch.float32, device=device)


def raycast_scene(origin: torch.Tensor, dirs: torch.Tensor, boxes: torch.Tensor | None = None,
                  min_range: float = 3.0, max_range: float = 50.0, ground_z: float = -1.7,
                  wall_y: float = 8.0, wall_top: float = 4.3) -> torch.Tensor:
    """First hit of every ray with {ground plane, walls y=+-wall_y, boxes}; returns the hit points [M,3] (metres)
    whose range lies in [min_range, max_range]."""
    o = origin.reshape(1, 3).to(dirs)
    inf = torch.full((dirs.shape[0],), float("inf"), device=dirs.device)
    dz = dirs[:, 2]
    t_best = torch.where(dz < -1e-6, (ground_z - o[0, 2]) / dz.clamp(max=-1e-6), inf)
    for sign in (1.0, -1.0):
        dy = dirs[:, 1] * sign
        t = torch.where(dy > 1e-6, (wall_y - sign * o[0, 1]) / dy.clamp(min=1e-6), inf)
        z_hit = o[0, 2] + t * dz
        t = torch.where((z_hit >= ground_z) & (z_hit <= wall_top), t, inf)
        t_best = torch.minimum(t_best, t)
    if boxes is not None and boxes.numel():
        safe = torch.where(dirs.abs() < 1e-9, torch.full_like(dirs, 1e-9), dirs)
        inv = (1.0 / safe).unsqueeze(1)                                   # [R,1,3]
        t0 = (boxes[None, :, :3] - o[:, None, :]) * inv
        t1 = (boxes[None, :, 3:] - o[:, None, :]) * inv
        t_near = torch.minimum(t0, t1).amax(-1)
        t_far = torch.maximum(t0, t1).amin(-1)
        hit = (t_far >= t_near) & (t_near > 0)
        t_box = torch.where(hit, t_near, torch.full_like(t_near, float("inf"))).amin(1)
        t_best = torch.minimum(t_best, t_box)
    keep = (t_best >= min_range) & (t_best <= max_range)
    return o + dirs[keep] * t_best[keep].unsqueeze(1)


def sample_rays(points_scaled: torch.Tensor, origin_scaled: torch.Tensor, config: SHINEConfig,
                generator: torch.Generator | None = None):
    """-> coord [M,3] (scaled, in [-1,1]), sdf_label [M] (scaled), weight [M] (+1 surface / -1 free), ray-wise
    ordered: for every ray its surface samples then its free-space samples."""
    dev = points_scaled.device
    shift = points_scaled - origin_scaled
    dist = torch.linalg.norm(shift, dim=1, keepdim=True)                              # [R,1]
    R = shift.shape[0]
    ns, nf = config.surface_sample_n, config.free_sample_n
    rng = config.surface_sample_range_m * config.scale

    def rand(*shape):
        return torch.rand(*shape, device=dev, generator=generator)

    surf_disp = (rand(R, ns) - 0.5) * 2.0 * rng                                       # [R,ns]
    surf_ratio = surf_disp / dist + 1.0
    free_max = config.free_sample_end_dist_m * config.scale / dist + 1.0
    free_ratio = rand(R, nf) * (free_max - config.free_sample_begin_ratio) + config.free_sample_begin_ratio
    free_disp = (free_ratio - 1.0) * dist
    ratio = torch.cat((surf_ratio, free_ratio), 1)                                    # [R, ns+nf]
    disp = torch.cat((surf_disp, free_disp), 1)
    coord = (shift.unsqueeze(1) * ratio.unsqueeze(2) + origin_scaled.reshape(1, 1, 3)).reshape(-1, 3)
    weight = torch.ones(R, ns + nf, device=dev)
    weight[:, ns:] = -1.0
    return coord.contiguous(), disp.reshape(-1).contiguous(), weight.reshape(-1).contiguous()


class SamplePool:
    """coord / sdf_label / weight pools + `get_batch()` (dataset/lidar_dataset.py:431-448).

    `sort_morton()` puts the pool in Morton order of the sample coordinates (once, when the map is built).  From then on
    `get_batch()` draws the SAME random index multiset as the reference (`torch.randint`, with replacement) and hands the
    samples out in ascending index order, i.e. in Morton order: neighbouring points of a batch then touch the same octree
    nodes, which is what the gather (L1 hits) and the voxel-grouped scatter of the training kernel feed on
    (`SdfTrainer.forward_backward(..., morton_ordered=pool.ordered)`).  The loss of a batch does not depend on its order."""

    def __init__(self, device):
        self.device = device
        self.coord_pool = torch.empty(0, 3, device=device)
        self.sdf_label_pool = torch.empty(0, device=device)
        self.weight_pool = torch.empty(0, device=device)
        self.ordered = False

    def append(self, coord, label, weight):
        self.coord_pool = torch.cat((self.coord_pool, coord.to(self.device)))
        self.sdf_label_pool = torch.cat((self.sdf_label_pool, label.to(self.device)))
        self.weight_pool = torch.cat((self.weight_pool, weight.to(self.device)))
        self.ordered = False

    def __len__(self):
        return self.sdf_label_pool.shape[0]

    def sort_morton(self, level: int = 16, octree=None):
        """Reorder the pool along the Z-order curve of a 2^level grid over [-1, 1]^3 (16: kaolin's int16 coordinates).
        With `octree` (the map the pool belongs to) the samples that see no node on any level — free space: their features
        are 0 whatever the tables hold — go behind all others, each part in Z-order: the tiles of a batch that take the
        step kernel's zero-tile shortcut then sit at the end of the batch, where the kernel's strided tile schedule hands
        every warp the same share of them."""
        from .feature_octree import points_to_morton, quantize_points
        if len(self):
            key = points_to_morton(quantize_points(self.coord_pool, level))
            if octree is not None:
                key = key | ((~octree.sees_a_node(self.coord_pool)).long() << 62)      # Morton keys use 48 bits
            order = torch.argsort(key)
            self.coord_pool = self.coord_pool[order].contiguous()
            self.sdf_label_pool = self.sdf_label_pool[order].contiguous()
            self.weight_pool = self.weight_pool[order].contiguous()
        self.ordered = True
        return self

    def get_batch(self, bs: int, generator: torch.Generator | None = None, ordered: bool | None = None):
        """ordered: None = Morton order iff the pool is sorted; False = the reference's order (as drawn)."""
        index = torch.randint(0, len(self), (bs,), device=self.device, generator=generator)
        if self.ordered and ordered is not False:
            index = torch.sort(index).values
        elif ordered:
            raise ValueError("ordered batches need a pool in Morton order: call sort_morton() first")
        return self.coord_pool[index, :], self.sdf_label_pool[index], self.weight_pool[index]


def generate_scans(config: SHINEConfig, n_azimuth: int, n_frames: int = 1, frame_step_m: float = 1.0, seed: int = 42,
                   device=None, origin_x0: float = 0.0):
    """Scan the analytic scene from `n_frames` poses along +x and sample every scan like the reference's sampler.
    -> list of (coord, sdf_label, weight, hits_scaled) per frame."""
    device = device or config.device
    gen = torch.Generator(device=device)
    gen.manual_seed(seed)
    dirs = lidar_directions(n_azimuth, device=device)
    boxes = default_boxes(device)
    boxes[:, 0] += origin_x0
    boxes[:, 3] += origin_x0
    frames = []
    for f in range(n_frames):
        origin = torch.tensor([origin_x0 + f * frame_step_m, 0.0, 0.0], device=device)
        hits = raycast_scene(origin, dirs, boxes, min_range=config.min_range, max_range=config.pc_radius)
        coord, label, weight = sample_rays(hits * config.scale, origin * config.scale, config, gen)
        frames.append((coord, label, weight, hits * config.scale))
    return frames


def build_scene_map(config: SHINEConfig, octree, n_azimuth: int, n_frames: int = 1, frame_step_m: float = 1.0,
                    seed: int = 42, device=None, origin_x0: float = 0.0):
    """Scan the analytic scene from `n_frames` poses along +x, sample every scan, grow the octree from the
    surface samples (weight > 0; dataset/lidar_dataset.py:212-218) and return the SamplePool."""
    device = device or config.device
    pool = SamplePool(device)
    for coord, label, weight, hits in generate_scans(config, n_azimuth, n_frames, frame_step_m, seed, device, origin_x0):
        if config.octree_from_surface_samples:
            octree.update(coord[weight > 0, :])
        else:
            octree.update(hits)
        pool.append(coord, label, weight)
    return pool
