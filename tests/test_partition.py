"""Host logic of the spatial partition (CPU, no GPU): balanced Morton-prefix split, local octree == global octree
restricted to the range, BoundaryPlan exchange == single-process gradient, with the oracle standing in for the kernels;
and the same exchange through a real 2-process gloo all-reduce."""
import os
import socket

import numpy as np
import pytest
import torch

from tests.parity_utils import DEC_KEYS, orc
from tests.partition_utils import check_rank_against_global, global_oracle_step, global_scene


def _rank_setup(cfg, pool, batch, world, dec_floats=1380):
    """Partition pool + batch, build every rank's package octree (CPU) and oracle octree."""
    from shine_mapping_b200 import FeatureOctree
    from shine_mapping_b200.partition import BoundaryPlan, coarse_keys, corner_keys_of, owner_of, partition_pool
    bounds, parts = partition_pool(*pool, cfg, world)
    level = cfg.tree_level_world - cfg.tree_level_feat + 1
    b_owner = owner_of(coarse_keys(batch[0], level), bounds)
    ranks = []
    for r in range(world):
        c, l, w = parts[r]
        octree = FeatureOctree(cfg)
        octree.update(c[w > 0])
        o = orc.OracleOctree(cfg.tree_level_world, cfg.tree_level_feat, cfg.feature_dim, cfg.feature_std, cfg.poly_int_on)
        o.update(c[w > 0])
        for lvl in range(cfg.tree_level_world + 1):      # package update() == oracle update() on this rank's range
            assert octree.nodes_lookup_tables[lvl] == o.nodes_lookup_tables[lvl]
        m = b_owner == r
        ranks.append({"octree": octree, "oracle": o, "keys": corner_keys_of(octree), "batch": (batch[0][m], batch[1][m])})
    plans = [BoundaryPlan(r, [x["keys"] for x in ranks], cfg.feature_dim, dec_floats) for r in range(world)]
    return bounds, parts, ranks, plans


def test_balanced_bounds_partition_every_sample_once():
    from shine_mapping_b200.partition import balanced_key_bounds, coarse_keys, owner_of
    cfg, pool, batch, dec = global_scene()
    keys = coarse_keys(pool[0], cfg.tree_level_world - cfg.tree_level_feat + 1)
    for world in (1, 2, 3, 8):
        bounds = balanced_key_bounds(keys, world)
        assert bounds.numel() == world + 1 and bool((bounds[1:] >= bounds[:-1]).all())
        owner = owner_of(keys, bounds)
        assert int(owner.min()) >= 0 and int(owner.max()) <= world - 1
        sizes = torch.bincount(owner, minlength=world).float()
        # a voxel is never split, so balance is limited by the biggest coarse voxel
        biggest = torch.unique(keys, return_counts=True)[1].max().item()
        assert float(sizes.max() - sizes.min()) <= 2.5 * biggest + 1, (world, sizes.tolist())
        # contiguous key ranges
        for r in range(world):
            k = keys[owner == r]
            if k.numel():
                assert int(k.min()) >= int(bounds[r]) and int(k.max()) < int(bounds[r + 1])


def test_holder_masks_name_exactly_the_ranks_that_list_a_shared_corner():
    """BoundaryPlan.holders (what the peer-memory exchange reads instead of every rank's buffer): bit r of a slot's
    mask <=> rank r's plan lists that slot; every mask has >= 2 bits; all ranks agree on the masks."""
    cfg, pool, batch, dec = global_scene()
    world = 3
    _, _, _, plans = _rank_setup(cfg, pool, batch, world)
    for lvl in range(len(plans[0].counts)):
        masks = plans[0].holders[lvl].long()
        assert masks.numel() == plans[0].counts[lvl]
        for p in plans[1:]:
            assert torch.equal(p.holders[lvl].long(), masks)
        want = torch.zeros_like(masks)
        for r, p in enumerate(plans):
            want[p.slots[lvl].long()] |= 1 << r
            assert torch.equal(p.inverse[lvl][p.slots[lvl].long()].long(), p.rows[lvl].long())
        assert torch.equal(want, masks)
        if masks.numel():
            bits = sum(((masks >> r) & 1) for r in range(world))
            assert int(bits.min()) >= 2


@pytest.mark.parametrize("world", [2, 3])
def test_partitioned_step_equals_single_process_step(world):
    """Sum over ranks of (local step on the rank's share of a fixed global batch) + boundary exchange == the single
    process step on the whole batch: loss, decoder gradients, and every table row (matched by corner key)."""
    cfg, pool, batch, dec = global_scene()
    o_glob, key_to_row, res_glob = global_oracle_step(cfg, pool, batch, dec)
    bounds, parts, ranks, plans = _rank_setup(cfg, pool, batch, world)
    n_global = batch[0].shape[0]
    assert sum(x["batch"][0].shape[0] for x in ranks) == n_global
    assert sum(p.counts[-1] for p in plans[:1]) > 0, "the ranges do not touch: the test would not exercise the boundary"
    F = cfg.feature_dim
    bufs, loss, dec_sum, local = [], 0.0, None, []
    for r, x in enumerate(ranks):
        o = x["oracle"]
        # local features = the global ones of the same corners (duplicates identical by construction)
        feats = []
        for lvl, keys in enumerate(x["keys"]):
            rows = torch.tensor([key_to_row[lvl][int(k)] for k in keys.tolist()], dtype=torch.long)
            f = torch.cat((o_glob.hier_features[lvl].detach()[rows], torch.zeros(1, F)))
            feats.append(f.clone().requires_grad_(True))
        o.hier_features = feats
        d = {k: v.detach().clone().requires_grad_(True) for k, v in dec.items()}
        res = orc.train_step(o, d, x["batch"][0], x["batch"][1], None, float(cfg.sigma_sigmoid), False, "sum")
        grads = [g / n_global for g in res["table_grads"]]
        loss += float(res["loss"]) / n_global
        dflat = torch.cat([res["dec_grads"][k].reshape(-1) for k in DEC_KEYS]) / n_global
        dec_sum = dflat if dec_sum is None else dec_sum + dflat
        buf = torch.zeros(plans[r].total_floats)
        plans[r].pack(grads, buf)
        bufs.append(buf); local.append(grads)
    total = torch.stack(bufs).sum(0)                       # what the all-reduce leaves on every rank
    for r, x in enumerate(ranks):
        plans[r].unpack(local[r], total)
        check_rank_against_global([g.numpy() for g in local[r]], x["keys"], key_to_row, res_glob, grad_rel=1e-5)
    want_dec = torch.cat([res_glob["dec_grads"][k].reshape(-1) for k in DEC_KEYS])
    assert float((dec_sum - want_dec).abs().max()) <= 1e-5 * float(want_dec.abs().max())
    assert abs(loss - float(res_glob["loss"])) <= 1e-5 * abs(float(res_glob["loss"]))


def test_unify_values_gives_every_duplicate_the_owners_features():
    cfg, pool, batch, dec = global_scene()
    bounds, parts, ranks, plans = _rank_setup(cfg, pool, batch, 2)
    tables = [[p.detach().clone() for p in x["octree"].hier_features] for x in ranks]
    bufs = []

    class Collect:          # stand-in for the collective: run both ranks' "all_reduce" in two passes
        def __init__(self): self.sent, self.total = [], None
    col = Collect()

    def capture(buf): col.sent.append(buf.clone())
    for r in range(2):
        plans[r].unify_values(tables[r], capture)           # pass 1: what each rank would contribute
    col.total = torch.stack(col.sent).sum(0)
    tables = [[p.detach().clone() for p in x["octree"].hier_features] for x in ranks]
    for r in range(2):
        plans[r].unify_values(tables[r], lambda buf: buf.copy_(col.total))
    # the shared corners now hold identical values on both ranks, equal to rank 0's original ones
    for lvl in range(cfg.tree_level_feat):
        k0, k1 = ranks[0]["keys"][lvl], ranks[1]["keys"][lvl]
        common = set(k0.tolist()) & set(k1.tolist())
        pos0 = {int(k): i for i, k in enumerate(k0.tolist())}; pos1 = {int(k): i for i, k in enumerate(k1.tolist())}
        orig0 = ranks[0]["octree"].hier_features[lvl].detach()
        for k in list(common)[:200]:
            assert torch.equal(tables[0][lvl][pos0[k]], tables[1][lvl][pos1[k]])
            assert torch.equal(tables[0][lvl][pos0[k]], orig0[pos0[k]])


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gloo_worker(rank, world, port, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    import torch.distributed as dist
    from shine_mapping_b200 import FeatureOctree, dist as sdist
    from shine_mapping_b200.partition import (BoundaryPlan, coarse_keys, gather_corner_keys, owner_of, partition_pool)
    sdist.init_from_env("gloo")
    torch.set_num_threads(1)
    cfg, pool, batch, dec = global_scene()
    bounds, parts = partition_pool(*pool, cfg, world)
    c, l, w = parts[rank]
    octree = FeatureOctree(cfg)
    octree.update(c[w > 0])
    plan = BoundaryPlan(rank, gather_corner_keys(octree), cfg.feature_dim, 1380)      # all_gather_object over gloo
    o = orc.OracleOctree(cfg.tree_level_world, cfg.tree_level_feat, cfg.feature_dim, cfg.feature_std, cfg.poly_int_on)
    o.update(c[w > 0])
    # identical duplicates: owner's values through the real collective
    tables = [p.detach().clone() for p in octree.hier_features]
    plan.unify_values(tables, lambda b: dist.all_reduce(b))
    o.hier_features = [t.clone().requires_grad_(True) for t in tables]
    m = owner_of(coarse_keys(batch[0], cfg.tree_level_world - cfg.tree_level_feat + 1), bounds) == rank
    n_global = batch[0].shape[0]
    d = {k: v.detach().clone().requires_grad_(True) for k, v in dec.items()}
    res = orc.train_step(o, d, batch[0][m], batch[1][m], None, float(cfg.sigma_sigmoid), False, "sum")
    grads = [g / n_global for g in res["table_grads"]]
    buf = torch.zeros(plan.total_floats)
    buf[:1377 + 3] = torch.cat([torch.cat([res["dec_grads"][k].reshape(-1) / n_global,
                                           torch.zeros((-res["dec_grads"][k].numel()) % 4)]) for k in DEC_KEYS])
    plan.pack(grads, buf)
    dist.all_reduce(buf)                  # ONE collective: [decoder | boundary rows]
    plan.unpack(grads, buf)
    torch.save({"tables": tables, "grads": grads, "dec": buf[:1380].clone(), "keys": [k.clone() for k in
                [octree._levels[octree.free_level_num + k].corner_morton_by_row for k in range(cfg.tree_level_feat)]]},
               os.path.join(out_dir, f"rank{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_process_gloo_exchange_keeps_duplicates_identical(tmp_path):
    """World-size-2 run over gloo: after the ONE all-reduce of [decoder | boundary rows] both ranks hold the same
    decoder gradient and the same gradient on every shared corner row; values were unified through the collective."""
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    a, b = (torch.load(tmp_path / f"rank{r}.pt") for r in range(2))
    assert torch.equal(a["dec"], b["dec"]) and float(a["dec"].abs().max()) > 0
    shared_rows = 0
    for lvl in range(len(a["keys"])):
        pa = {int(k): i for i, k in enumerate(a["keys"][lvl].tolist())}
        pb = {int(k): i for i, k in enumerate(b["keys"][lvl].tolist())}
        for k in set(pa) & set(pb):
            assert torch.equal(a["grads"][lvl][pa[k]], b["grads"][lvl][pb[k]])
            assert torch.equal(a["tables"][lvl][pa[k]], b["tables"][lvl][pb[k]])
            shared_rows += 1
    assert shared_rows > 0
