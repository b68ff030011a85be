"""Incremental mapping loop with the regularisation-based continual-learning term — the caller of the hot path for
BASELINE config 4, after reference shine_incre.py:86-194 and utils/incre_learning.py:8-40.

Per frame: the pool holds this frame's samples only, `octree.update(surface, incremental_on=True)` grows the map and
snapshots `features_last_frame` / extends `importance_weight`; the optimiser state is rebuilt (reference
shine_incre.py:108-109); `iters` x { get_batch -> fused fwd+loss(sum)+bwd -> + lambda_forget * d(reg)/d(features)
-> Adam }; then `cal_feature_importance` sweeps the frame's pool and accumulates |dL/dfeature| into the importance.

The BCE part is the fused sm_100a step; the regulariser (model/feature_octree.py:246-255) and the importance update touch
only the rows the batch touched: `shine_mark_touched` collects them (bitmap + compact list, no unique()/sort) and
`shine_regularization_apply` / `shine_importance_accumulate` run over that list.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _abi
from .config import SHINEConfig
from .decoder import Decoder
from .feature_octree import FeatureOctree
from .trainer import SdfTrainer


class TouchedRows:
    """Device scratch of the per-touched-row passes: one bitmap word per 32 rows and a compact row list per level
    (bottom-up like the C descriptors).  Rebuilt when `octree.update()` re-grows the tables."""

    def __init__(self, octree: FeatureOctree):
        dev = octree.hier_features[0].device
        L = octree.featured_level_num
        self.counts = torch.zeros(L, dtype=torch.int32, device=dev)
        self.bitmaps, self.rows = [], []
        self.desc = _abi.ShineTouched()
        self.rows_sig = tuple(int(p.shape[0]) for p in octree.hier_features)
        for i in range(L):
            k = L - i - 1
            n_rows = int(octree.hier_features[k].shape[0])
            bm = torch.zeros((n_rows + 31) // 32, dtype=torch.int32, device=dev)
            rl = torch.empty(n_rows, dtype=torch.int32, device=dev)
            self.bitmaps.append(bm); self.rows.append(rl)
            lv = self.desc.lv[i]
            lv.bitmap, lv.rows = bm.data_ptr(), rl.data_ptr()
            lv.count = self.counts.data_ptr() + 4 * i
            lv.capacity = n_rows

    @staticmethod
    def of(octree: FeatureOctree) -> "TouchedRows":
        t = getattr(octree, "_touched_rows", None)
        if t is None or t.rows_sig != tuple(int(p.shape[0]) for p in octree.hier_features) or \
                t.counts.device != octree.hier_features[0].device:
            t = TouchedRows(octree)
            octree._touched_rows = t
        return t


def _row_tables(octree: FeatureOctree, writable: bool) -> _abi.ShineRowTables:
    aux = _abi.ShineRowTables()
    L = octree.featured_level_num
    for i in range(L):
        k = L - i - 1
        aux.last[i] = octree.features_last_frame[k].data_ptr()
        aux.importance[i] = octree.importance_weight[k].data_ptr()
        aux.importance_rw[i] = octree.importance_weight[k].data_ptr() if writable else None
    return aux


def add_regularization(trainer: SdfTrainer, octree: FeatureOctree, lambda_forget: float, coord=None) -> torch.Tensor:
    """reg = cal_regularization() over the rows touched by the last batch (model/feature_octree.py:246-255) and
    grads += 2 lambda Omega (f - f_last) on those rows — two launches (mark, apply), no unique()/sort, nothing dense."""
    coord = octree._last_coord if coord is None else coord
    if coord is None:
        raise _abi.ShineB200Error("add_regularization needs the batch of the last step (octree._last_coord)")
    dev = trainer.flat_grad.device
    t = TouchedRows.of(octree)
    od = octree._descriptor(None, trainer.table_grads)
    reg = torch.zeros((), device=dev)
    t.counts.zero_()
    lib, st = _abi.lib(), _abi.stream_ptr(dev)
    _abi.check(lib.shine_mark_touched(C.byref(od), _abi.ptr(coord), coord.shape[0], C.byref(t.desc), st),
               "shine_mark_touched")
    aux = _row_tables(octree, writable=False)
    _abi.check(lib.shine_regularization_apply(C.byref(od), C.byref(t.desc), C.byref(aux), 2.0 * lambda_forget,
                                              _abi.ptr(reg), 1, st), "shine_regularization_apply")
    return reg


@torch.no_grad()
def cal_feature_importance(trainer: SdfTrainer, octree: FeatureOctree, coord_pool, label_pool, bs: int, down_rate: int = 1):
    """utils/incre_learning.py:8-40 on the fused kernel: per pool stride, one unweighted step, then
    importance[u] += |dL/dfeature[u]| over the rows that stride touched (which also re-zeroes their gradients)."""
    n = coord_pool.shape[0]
    interval = bs * down_rate
    dev = trainer.flat_grad.device
    t = TouchedRows.of(octree)
    aux = _row_tables(octree, writable=True)
    lib, st = _abi.lib(), _abi.stream_ptr(dev)
    trainer.zero_grad()
    for head in range(0, n, interval):
        c = coord_pool[head:min(head + interval, n):down_rate].contiguous()
        l = label_pool[head:min(head + interval, n):down_rate].contiguous()
        trainer.forward_backward(c, l, weighted=False)     # utils/incre_learning.py:33: weight=None
        od = octree._descriptor(None, trainer.table_grads)
        t.counts.zero_()
        _abi.check(lib.shine_mark_touched(C.byref(od), _abi.ptr(c), c.shape[0], C.byref(t.desc), st), "shine_mark_touched")
        _abi.check(lib.shine_importance_accumulate(C.byref(od), C.byref(t.desc), C.byref(aux), 1, 1, st),
                   "shine_importance_accumulate")
    trainer.zero_grad()          # decoder segment + loss accumulator


def run_shine_mapping_incremental(config: SHINEConfig, octree: FeatureOctree, decoder: Decoder, frames, iters=None,
                                  log=None):
    """frames: iterable of (coord, sdf_label, weight) sample sets, one per scan (what `process_frame` leaves in the
    pools).  Returns per-frame dicts with first/last loss."""
    if config.continual_learning_reg:
        config.loss_reduction = "sum"          # reference shine_incre.py:77-78
    iters = config.iters if iters is None else iters
    dev = None
    history = []
    for fid, (coord, label, weight) in enumerate(frames):
        if fid == config.freeze_after_frame:   # reference shine_incre.py:97-101
            for child in decoder.children():
                for p in child.parameters():
                    p.requires_grad = False
        surface = coord[weight > 0, :]
        octree.update(surface, incremental_on=config.continual_learning_reg)        # lidar_dataset.py:212-218
        trainer = SdfTrainer(config, octree, decoder)                               # fresh Adam state per frame
        dev = trainer.flat_grad.device
        trainer.zero_grad()
        first = last = None
        n = coord.shape[0]
        for it in range(iters):
            index = torch.randint(0, n, (config.bs,), device=dev)
            c, l, w = coord[index], label[index], weight[index]
            loss = trainer.forward_backward(c, l, w)
            total = loss.clone()
            if config.continual_learning_reg:
                total = total + config.lambda_forget * add_regularization(trainer, octree, config.lambda_forget, c)
            trainer.optimizer_step(zero_grad=True)
            if it == 0:
                first, bce_first = float(total), float(loss)
        last, bce_last = float(total), float(loss)
        if config.continual_learning_reg:
            cal_feature_importance(trainer, octree, coord, label, config.bs, config.cal_importance_weight_down_rate)
        history.append({"frame": fid, "loss_first": first, "loss_last": last, "bce_first": bce_first, "bce_last": bce_last,
                        "rows": [int(p.shape[0]) for p in octree.hier_features]})
        if log:
            log(history[-1])
    return history
