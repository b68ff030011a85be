"""kaolin.ops.spc subset (call sites: reference model/feature_octree.py:88-89,97,123,131,134,162-164,203-204)."""
import torch


def quantize_points(x, level):
    # kaolin doc: floor(clamp(res * (x + 1) / 2, 0, res - 1)) as int16
    res = 2 ** level
    return torch.floor(torch.clamp(res * (x + 1.0) / 2.0, 0, res - 1.0)).short()


def _spread3(v):
    # v: int64 tensor holding 16-bit values -> bits spaced by 3
    out = torch.zeros_like(v)
    for i in range(16):
        out |= ((v >> i) & 1) << (3 * i)
    return out


def points_to_morton(points):
    # x -> bit 3i+2, y -> bit 3i+1, z -> bit 3i  (kaolin spc convention)
    p = points.long()
    shape = p.shape[:-1]
    p = p.reshape(-1, 3)
    m = (_spread3(p[:, 0]) << 2) | (_spread3(p[:, 1]) << 1) | _spread3(p[:, 2])
    return m.reshape(shape)


def _compact3(m):
    out = torch.zeros_like(m)
    for i in range(16):
        out |= ((m >> (3 * i)) & 1) << i
    return out


def morton_to_points(mortons):
    m = mortons.long()
    shape = m.shape
    m = m.reshape(-1)
    pts = torch.stack((_compact3(m >> 2), _compact3(m >> 1), _compact3(m)), dim=-1)
    return pts.reshape(*shape, 3).short()


_CORNER_OFFSETS = [[(i >> 2) & 1, (i >> 1) & 1, i & 1] for i in range(8)]


def points_to_corners(points):
    off = torch.tensor(_CORNER_OFFSETS, dtype=points.dtype, device=points.device)
    return points.unsqueeze(-2) + off
