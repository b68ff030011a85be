"""world_size-2 gloo tests (CPU) of the multi-GPU host logic in shine_mapping_b200/dist.py: the sharding rule and the
data-parallel gradient algebra (per-point scale 1/N_global + ONE sum all-reduce of the flat gradient buffer ==
single-process gradient of the global batch), exercised with the CPU oracle standing in for the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tests.parity_utils import make_case, oracle_from_case, orc


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _flat(res):
    parts = [g.reshape(-1) for g in res["table_grads"]] + [res["dec_grads"][k].reshape(-1) for k in sorted(res["dec_grads"])]
    return torch.cat([torch.as_tensor(p) for p in parts])


def _worker(rank, world, port, mode, out_dir):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from shine_mapping_b200 import dist as sdist
    r, w, _ = sdist.init_from_env("gloo")
    assert (r, w) == (rank, world)
    torch.set_num_threads(1)
    if mode == "replicated":
        case = make_case(n_points=1200, n_batch=1500, feat_levels=3, seed=9)
        o, dec = oracle_from_case(case)
        n = case["coord"].shape[0]
        b, e = sdist.shard_range(n, rank, world)
        coord, label = torch.from_numpy(case["coord"][b:e]), torch.from_numpy(case["label"][b:e])
        res = orc.train_step(o, dec, coord, label, None, case["cfg"]["sigma"], False, "sum")
        flat = _flat(res) / n                      # per-point scale 1/N_global (what loss_scale does in the kernel)
        sdist.all_reduce_sum(flat)
        loss = res["loss"].reshape(1) / n
        sdist.all_reduce_sum(loss)
    else:   # spatial: every rank owns its own block (own scene / table), only the decoder segment is exchanged
        case = make_case(n_points=1000, n_batch=800, feat_levels=2, seed=20 + rank)
        shared = make_case(n_points=1000, n_batch=800, feat_levels=2, seed=20)     # same decoder on every rank
        case["dec"] = shared["dec"]
        o, dec = oracle_from_case(case)
        n_global = 816 * world
        res = orc.train_step(o, dec, torch.from_numpy(case["coord"]), torch.from_numpy(case["label"]), None,
                             case["cfg"]["sigma"], False, "sum")
        flat = torch.cat([res["dec_grads"][k].reshape(-1) for k in sorted(res["dec_grads"])]) / n_global
        sdist.all_reduce_sum(flat)
        loss = res["loss"].reshape(1) / n_global
        sdist.all_reduce_sum(loss)
    assert sdist.max_over_ranks(float(rank), "cpu") == world - 1
    sdist.barrier()
    if rank == 0:
        np.save(os.path.join(out_dir, "flat.npy"), flat.numpy())
        np.save(os.path.join(out_dir, "loss.npy"), loss.numpy())
    dist.destroy_process_group()


def test_shard_range_is_a_balanced_partition():
    from shine_mapping_b200.dist import shard_range
    for n in (0, 1, 7, 16, 1000003):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - b for b, e in spans]
            assert max(sizes) - min(sizes) <= 1


@pytest.mark.timeout(300)
def test_replicated_data_parallel_equals_single_process(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, "replicated", str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "flat.npy"); got_loss = float(np.load(tmp_path / "loss.npy")[0])
    case = make_case(n_points=1200, n_batch=1500, feat_levels=3, seed=9)
    o, dec = oracle_from_case(case)
    res = orc.train_step(o, dec, torch.from_numpy(case["coord"]), torch.from_numpy(case["label"]), None,
                         case["cfg"]["sigma"], False, "mean")
    want = _flat(res).numpy()
    assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max() + 1e-12
    assert abs(got_loss - float(res["loss"])) <= 1e-6 * abs(float(res["loss"]))


@pytest.mark.timeout(300)
def test_spatial_blocks_share_only_decoder_gradients(tmp_path):
    port = _free_port()
    mp.spawn(_worker, args=(2, port, "spatial", str(tmp_path)), nprocs=2, join=True)
    got = np.load(tmp_path / "flat.npy")
    shared = make_case(n_points=1000, n_batch=800, feat_levels=2, seed=20)
    total = None
    for rank in range(2):
        case = make_case(n_points=1000, n_batch=800, feat_levels=2, seed=20 + rank)
        case["dec"] = shared["dec"]
        o, dec = oracle_from_case(case)
        res = orc.train_step(o, dec, torch.from_numpy(case["coord"]), torch.from_numpy(case["label"]), None,
                             case["cfg"]["sigma"], False, "sum")
        flat = torch.cat([res["dec_grads"][k].reshape(-1) for k in sorted(res["dec_grads"])]) / (816 * 2)
        total = flat if total is None else total + flat
    assert np.abs(got - total.numpy()).max() <= 1e-5 * np.abs(total.numpy()).max() + 1e-12
