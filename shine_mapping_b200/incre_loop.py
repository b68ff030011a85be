"""Incremental mapping loop with the regularisation-based continual-learning term — the caller of the hot path for
BASELINE config 4, after reference shine_incre.py:86-194 and utils/incre_learning.py:8-40.

Per frame: the pool holds this frame's samples only, `octree.update(surface, incremental_on=True)` grows the map and
snapshots `features_last_frame` / extends `importance_weight`; the optimiser state is rebuilt (reference
shine_incre.py:108-109); `iters` x { get_batch -> fused fwd+loss(sum)+bwd -> + lambda_forget * d(reg)/d(features)
-> Adam }; then `cal_feature_importance` sweeps the frame's pool and accumulates |dL/dfeature| into the importance.

The BCE part is the fused sm_100a step; the regulariser (model/feature_octree.py:246-255) touches only the unique rows of
the batch and is applied with a handful of torch GPU ops on the (lazily materialised) `hierarchical_indices`.
"""
from __future__ import annotations

import math

import torch

from .config import SHINEConfig
from .decoder import Decoder
from .feature_octree import FeatureOctree
from .trainer import SdfTrainer


def add_regularization(trainer: SdfTrainer, octree: FeatureOctree, lambda_forget: float) -> torch.Tensor:
    """loss += lambda * cal_regularization(); grads += lambda * 2 Omega (f - f_last) on the unique rows of the last batch."""
    reg = torch.zeros((), device=trainer.flat_grad.device)
    idx = octree.hierarchical_indices          # bottom-up [N,8] per level, materialised by shine_get_indices
    for i in range(octree.featured_level_num):
        k = octree.featured_level_num - i - 1
        u = idx[i].flatten().unique()           # includes -1 (the trash row), like the reference
        f = octree.hier_features[k].data
        diff = f[u] - octree.features_last_frame[k][u]
        w = octree.importance_weight[k][u]
        reg = reg + (w * diff * diff).sum()
        trainer.table_grads[k].index_add_(0, torch.where(u < 0, u + f.shape[0], u), (2.0 * lambda_forget) * w * diff)
    return reg


@torch.no_grad()
def cal_feature_importance(trainer: SdfTrainer, octree: FeatureOctree, coord_pool, label_pool, bs: int, down_rate: int = 1):
    """utils/incre_learning.py:8-40 on the fused kernel: importance += |dL/dfeature| per pool stride."""
    n = coord_pool.shape[0]
    interval = bs * down_rate
    for head in range(0, n, interval):
        c = coord_pool[head:min(head + interval, n):down_rate].contiguous()
        l = label_pool[head:min(head + interval, n):down_rate].contiguous()
        trainer.zero_grad()
        trainer.forward_backward(c, l, weighted=False)     # utils/incre_learning.py:33: weight=None
        for k in range(len(octree.importance_weight)):
            octree.importance_weight[k] += trainer.table_grads[k].abs()
            octree.importance_weight[k][-1] *= 0
    trainer.zero_grad()


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
                octree._last_coord, octree._hier_idx = c, []      # the batch whose unique rows are regularised
                total = total + config.lambda_forget * add_regularization(trainer, octree, config.lambda_forget)
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
