"""Spatial partition of ONE map over the GPUs of a box (the reference is single-GPU, so this is new
design on top of its data structures).

* Samples are assigned to ranks by the Morton key of their voxel at the COARSEST featured level
  (`tree_level_world - tree_level_feat + 1`): everything a sample touches on every featured level lies inside that
  voxel, so a rank that owns a contiguous key range owns every node its samples can hit.  The ranges are balanced over
  the pool's key histogram (`balanced_key_bounds`).
* Each rank grows its own `FeatureOctree` from its own surface samples = the global octree restricted to its range.
* Corner rows are shared between neighbouring voxels (reference model/feature_octree.py:131-137), so the corners on
  a face between two ranges exist on both ranks.  `BoundaryPlan` lists, per level, the corners held by more than one
  rank in a globally agreed order; every step their gradients are summed through a compact exchange buffer that also
  carries the decoder gradients (ONE all-reduce), which keeps the duplicates bit-identical to each other and the
  N-rank step equal to the single-GPU step on the same global batch.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from .feature_octree import FeatureOctree, points_to_morton, quantize_points


def coarse_keys(coord: torch.Tensor, level: int) -> torch.Tensor:
    """Morton key of every sample's voxel at `level` (same arithmetic as the kernels; works on CPU and CUDA)."""
    return points_to_morton(quantize_points(coord, level))


def balanced_key_bounds(keys: torch.Tensor, world: int) -> torch.Tensor:
    """[world + 1] int64 thresholds: rank r owns keys in [bounds[r], bounds[r+1]).  Cut points sit between distinct
    keys (a voxel is never split) where the cumulative sample count crosses r/world of the pool."""
    uniq, counts = torch.unique(keys, return_counts=True)
    cum = torch.cumsum(counts, 0)
    total = int(cum[-1])
    bounds = [0]
    for r in range(1, world):
        target = total * r // world
        j = int(torch.searchsorted(cum, torch.tensor(target, device=cum.device, dtype=cum.dtype), right=False))
        j = min(max(j + 1, 1), uniq.numel() - 1) if uniq.numel() > 1 else 0
        bounds.append(max(int(uniq[j]), bounds[-1]))
    bounds.append(1 << 62)
    return torch.tensor(bounds, dtype=torch.int64)


def owner_of(keys: torch.Tensor, bounds: torch.Tensor) -> torch.Tensor:
    """Rank that owns each key."""
    return torch.bucketize(keys, bounds[1:-1].to(keys.device), right=True)


class BoundaryPlan:
    """Which corner rows are duplicated across ranks, and where they sit in the exchange buffer.

    Built from the per-rank, per-level corner keys (`octree corner_morton_by_row`): pure function of those key sets,
    so every rank computes the same plan after an all-gather (and the single-process tests can build it for all ranks
    at once)."""

    def __init__(self, rank: int, per_rank_level_keys: list[list[torch.Tensor]], feature_dim: int, dec_floats: int):
        world = len(per_rank_level_keys)
        n_levels = len(per_rank_level_keys[rank])
        self.rank, self.world, self.feature_dim = rank, world, feature_dim
        self.dec_floats = dec_floats
        self.rows, self.slots, self.owned, self.counts, self.offsets, self.inverse = [], [], [], [], [], []
        self.holders = []                  # per level, per shared corner: bit r set <=> rank r holds a row of it
        off = dec_floats
        for lvl in range(n_levels):        # coarse -> fine, like hier_features
            keys_all = torch.cat([per_rank_level_keys[r][lvl].cpu() for r in range(world)])
            ranks_all = torch.cat([torch.full((per_rank_level_keys[r][lvl].numel(),), r, dtype=torch.int64)
                                   for r in range(world)])
            uniq, inv, cnt = torch.unique(keys_all, return_inverse=True, return_counts=True)
            shared = uniq[cnt > 1]                                     # sorted: the agreed order
            mine = per_rank_level_keys[rank][lvl].cpu()
            pos = torch.searchsorted(shared, mine).clamp_(max=max(shared.numel() - 1, 0))
            is_b = (shared[pos] == mine) if shared.numel() else torch.zeros_like(mine, dtype=torch.bool)
            rows = torch.nonzero(is_b).flatten()
            slots = pos[rows]
            # owner of a shared corner = the lowest rank that holds it (used to make the duplicates' VALUES identical)
            first = torch.full((uniq.numel(),), world, dtype=torch.int64)
            first.scatter_reduce_(0, inv, ranks_all, reduce="amin")
            owner_of_shared = first[cnt > 1]
            held = torch.zeros(uniq.numel(), dtype=torch.int64)
            held.scatter_add_(0, inv, torch.ones_like(ranks_all) << ranks_all)   # a rank lists a key once: sum == or
            self.holders.append(held[cnt > 1].to(torch.int32))
            inv = torch.full((int(shared.numel()),), -1, dtype=torch.int32)
            inv[slots] = rows.to(torch.int32)
            self.inverse.append(inv)
            self.rows.append(rows.to(torch.int32)); self.slots.append(slots.to(torch.int32))
            self.owned.append(owner_of_shared[slots] == rank)
            self.counts.append(int(shared.numel()))
            self.offsets.append(off)
            off += int(shared.numel()) * feature_dim
        self.total_floats = off                      # decoder segment + every level's boundary rows

    def to(self, device):
        self.rows = [t.to(device) for t in self.rows]
        self.slots = [t.to(device) for t in self.slots]
        self.owned = [t.to(device) for t in self.owned]
        self.inverse = [t.to(device) for t in self.inverse]
        self.holders = [t.to(device) for t in self.holders]
        return self

    def inverse_descriptor(self) -> _abi.ShineBoundaryInverse:
        d = _abi.ShineBoundaryInverse()
        for lvl, inv in enumerate(self.inverse):
            d.row_of_slot[lvl] = inv.data_ptr() if inv.numel() else None
            d.slots[lvl] = int(inv.numel())
            d.holders[lvl] = self.holders[lvl].data_ptr() if inv.numel() else None
        return d

    def descriptor(self, tables, buf_base_offset: int = 0) -> _abi.ShineBoundary:
        """C descriptor for `tables` (coarse -> fine list of [rows, F] tensors).  Offsets are relative to the buffer
        passed to pack/unpack (which starts at the decoder segment)."""
        d = _abi.ShineBoundary()
        for lvl, t in enumerate(tables):
            b = d.lv[lvl]
            b.table, b.rows, b.slots = t.data_ptr(), self.rows[lvl].data_ptr(), self.slots[lvl].data_ptr()
            b.offset, b.count = self.offsets[lvl] + buf_base_offset, int(self.rows[lvl].numel())
        return d

    # ---- exchange through a buffer laid out [decoder | boundary rows ...] ------------------------------------------

    def pack(self, tables, buf):
        if buf.is_cuda:
            d = self.descriptor(tables)
            _abi.check(_abi.lib().shine_boundary_pack(C.byref(d), len(tables), self.feature_dim, _abi.ptr(buf),
                                                      _abi.stream_ptr(buf.device)), "shine_boundary_pack")
        else:        # host logic of the gloo tests
            for lvl, t in enumerate(tables):
                seg = buf[self.offsets[lvl]:self.offsets[lvl] + self.counts[lvl] * self.feature_dim].view(-1, self.feature_dim)
                seg[self.slots[lvl].long()] = t[self.rows[lvl].long()]

    def unpack(self, tables, buf):
        if buf.is_cuda:
            d = self.descriptor(tables)
            _abi.check(_abi.lib().shine_boundary_unpack(C.byref(d), len(tables), self.feature_dim, _abi.ptr(buf),
                                                        _abi.stream_ptr(buf.device)), "shine_boundary_unpack")
        else:
            for lvl, t in enumerate(tables):
                seg = buf[self.offsets[lvl]:self.offsets[lvl] + self.counts[lvl] * self.feature_dim].view(-1, self.feature_dim)
                t[self.rows[lvl].long()] = seg[self.slots[lvl].long()]

    @torch.no_grad()
    def unify_values(self, tables, all_reduce):
        """Make the duplicates of every shared corner hold the OWNER's values (features at start-up, Adam moments are
        zero): owner writes, others contribute zeros, one sum all-reduce, everybody copies back."""
        dev = tables[0].device
        buf = torch.zeros(self.total_floats, dtype=torch.float32, device=dev)
        self.pack(tables, buf)
        for lvl in range(len(tables)):       # zero the slots this rank does not own
            seg = buf[self.offsets[lvl]:self.offsets[lvl] + self.counts[lvl] * self.feature_dim].view(-1, self.feature_dim)
            not_owned = self.slots[lvl][~self.owned[lvl]].long()
            seg[not_owned] = 0.0
        all_reduce(buf)
        self.unpack(tables, buf)


def corner_keys_of(octree: FeatureOctree) -> list[torch.Tensor]:
    """Per featured level (coarse -> fine) the Morton key of every corner row (the trash row has none)."""
    return [octree._levels[octree.free_level_num + k].corner_morton_by_row
            for k in range(octree.featured_level_num)]


def gather_corner_keys(octree: FeatureOctree, group=None) -> list[list[torch.Tensor]]:
    """all_gather of every rank's per-level corner keys (set-up time, variable sizes)."""
    import torch.distributed as dist
    mine = [t.cpu() for t in corner_keys_of(octree)]
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [mine]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, mine, group=group)
    return out


def partition_pool(coord, label, weight, config, world: int, bounds: torch.Tensor | None = None):
    """Split one global sample pool by the Morton key of every sample's voxel at the coarsest featured level.
    -> (bounds, [ (coord, label, weight) per rank ])."""
    level = config.tree_level_world - config.tree_level_feat + 1
    keys = coarse_keys(coord, level)
    if bounds is None:
        bounds = balanced_key_bounds(keys.cpu(), world)
    owner = owner_of(keys, bounds)
    parts = []
    for r in range(world):
        m = owner == r
        parts.append((coord[m].contiguous(), label[m].contiguous(), weight[m].contiguous()))
    return bounds, parts


def build_rank_map(config, octree: FeatureOctree, part, device=None):
    """Grow `octree` from this rank's surface samples (what LiDARDataset does with the whole pool,
    dataset/lidar_dataset.py:212-218) and wrap its samples in a SamplePool."""
    from .synth import SamplePool
    coord, label, weight = part
    octree.update(coord[weight > 0, :])
    pool = SamplePool(device or coord.device)
    pool.append(coord, label, weight)
    return pool


def decoder_segment_floats(decoder) -> int:
    """Size of the decoder segment of SdfTrainer's flat gradient buffer (every tensor padded to 4 floats)."""
    return sum((p.numel() + 3) & ~3 for p in decoder.fused_params() if p is not None)
