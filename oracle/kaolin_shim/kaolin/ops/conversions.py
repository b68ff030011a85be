"""kaolin.ops.conversions.unbatched_pointcloud_to_spc (call site: reference model/feature_octree.py:116)."""
import torch
from . import spc as _spc


class _Spc:
    def __init__(self, point_hierarchies, pyramids):
        self.point_hierarchies = point_hierarchies
        self.pyramids = pyramids


def unbatched_pointcloud_to_spc(pointcloud, level, features=None):
    q = _spc.quantize_points(pointcloud.contiguous(), level)
    leaf = torch.unique(_spc.points_to_morton(q))  # sorted Morton order
    per_level = []
    counts = []
    for l in range(level + 1):
        m = torch.unique(leaf >> (3 * (level - l)))
        per_level.append(_spc.morton_to_points(m))
        counts.append(m.shape[0])
    counts_t = torch.tensor(counts + [0], dtype=torch.int32)
    offsets = torch.zeros(level + 2, dtype=torch.int32)
    offsets[1:] = torch.cumsum(torch.tensor(counts, dtype=torch.int32), 0)
    pyramid = torch.stack((counts_t, offsets), 0).unsqueeze(0)  # [1, 2, level+2]
    return _Spc(torch.cat(per_level, 0), pyramid)
