"""Parity of the sm_100a path (through the C ABI behind the package classes) against the CPU oracle and the
reference-minted golden vectors.  Tolerances are stated in tests/parity_utils.py."""
import ctypes as C

import numpy as np
import pytest
import torch

from tests.parity_utils import (DEC_KEYS, GOLDEN_NAMES, build_cuda_models, compare_step, load_golden, make_case,
                                run_cuda_step, run_oracle_step, sort_case_morton)

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(scope="module", autouse=True)
def _lib(built_lib):
    assert torch.cuda.is_available()
    return built_lib


# ---- against the reference's own outputs -----------------------------------------------------------------

@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_cuda_matches_reference_golden(name):
    case, exp = load_golden(name)
    got = run_cuda_step(case, DEV)
    if case["cfg"]["decoder_frozen"]:
        got["dec_grads"] = {}
    print(name, compare_step(got, exp))


# ---- against the oracle on seeded synthetic batches -----------------------------------------------------

@pytest.mark.parametrize("levels,poly,weighted,reduction,frames", [
    (1, True, False, "mean", 1), (2, True, False, "mean", 1), (3, False, True, "sum", 2),
    (4, True, False, "mean", 1), (4, True, True, "mean", 2), (4, False, False, "sum", 1),
    (6, True, False, "mean", 1), (8, True, True, "sum", 2),        # > 4 levels: the LMAX = 8 kernel instantiations
])
def test_fused_step_matches_oracle(levels, poly, weighted, reduction, frames):
    case = make_case(n_points=2500, n_batch=3000, feat_levels=levels, seed=10 + levels, n_frames=frames,
                     poly=poly, weighted=weighted, reduction=reduction)
    print(compare_step(run_cuda_step(case, DEV), run_oracle_step(case)))


@pytest.mark.parametrize("levels,poly,weighted,reduction,ordered", [
    (1, True, False, "mean", True), (2, False, True, "sum", True), (3, True, False, "mean", True),
    (4, True, True, "mean", True), (4, True, False, "mean", False), (3, False, False, "sum", False),
    (6, True, False, "mean", True),      # > 4 levels: the flag falls back to the general kernel
])
def test_grouped_scatter_matches_oracle(levels, poly, weighted, reduction, ordered):
    """SHINE_FLAG_MORTON_ORDERED: per-run tensor-core reduction of the table gradients.  Batches in Morton order (runs of
    equal node, several per tile) and in random order (the hint is wrong: every level of every tile takes the per-point
    path) must both match the oracle; a frozen decoder exercises the kernel flavour without the staging area."""
    case = make_case(n_points=2500, n_batch=6000, feat_levels=levels, seed=40 + levels, n_frames=1,
                     poly=poly, weighted=weighted, reduction=reduction)
    if ordered:
        case = sort_case_morton(case)
    want = run_oracle_step(case)
    print(compare_step(run_cuda_step(case, DEV, morton_ordered=True), want))
    want_f = dict(want); want_f["dec_grads"] = {}
    print(compare_step(run_cuda_step(case, DEV, morton_ordered=True, freeze_decoder=True), want_f))


@pytest.mark.parametrize("ordered,weighted,reduction,frozen", [(True, False, "mean", False), (False, True, "sum", False),
                                                              (True, True, "mean", True)])
def test_all_miss_tiles_match_oracle(ordered, weighted, reduction, frozen):
    """Whole 16-point tiles of free-space samples that miss every level (a third of the tiles of an ordered C2 batch): the
    kernel gives them Decoder.sdf(0) and folds their decoder gradients in by linearity.  Loss, predictions and every
    gradient must equal the oracle's, with the block of far points in the middle of a batch (order drawn) and spread by
    the Morton sort (ordered), general and grouped kernels, weighted / sum, and with a frozen decoder."""
    case = make_case(n_points=2500, n_batch=1500, feat_levels=3, seed=91, weighted=weighted, reduction=reduction)
    rng = np.random.default_rng(7)
    far = np.concatenate([rng.uniform(0.55, 0.95, size=(700, 3)), rng.uniform(-0.95, -0.6, size=(420, 3))]).astype(np.float32)
    k = 640                                     # a tile-aligned block of far points inside the batch
    case["coord"] = np.concatenate([case["coord"][:k], far, case["coord"][k:]]).astype(np.float32)
    case["label"] = np.concatenate([case["label"][:k], rng.uniform(-0.2, 0.2, size=far.shape[0]).astype(np.float32),
                                    case["label"][k:]])
    case["weight"] = np.concatenate([case["weight"][:k], rng.uniform(0.5, 1.5, size=far.shape[0]).astype(np.float32),
                                     case["weight"][k:]])
    if ordered:
        case = sort_case_morton(case)
    want = run_oracle_step(case)
    missed_everywhere = np.logical_and.reduce([(idx == -1).all(axis=1) for idx in want["indices"]])
    assert int(missed_everywhere.sum()) >= far.shape[0]                     # the far points really miss every level
    if not ordered:
        assert bool(missed_everywhere[k:k + far.shape[0]].all())
    if frozen:
        want = dict(want); want["dec_grads"] = {}
    for flag in (False, True):
        print(compare_step(run_cuda_step(case, DEV, morton_ordered=flag, freeze_decoder=frozen), want))


def test_all_miss_batch():
    """Every point misses every level: only virtual tiles reach the decoder."""
    case = make_case(n_points=2500, n_batch=64, feat_levels=2, seed=93)
    rng = np.random.default_rng(3)
    n = 1000
    case["coord"] = rng.uniform(0.6, 0.9, size=(n, 3)).astype(np.float32)
    case["label"] = rng.uniform(-0.1, 0.1, size=n).astype(np.float32)
    case["weight"] = np.ones(n, dtype=np.float32)
    print(compare_step(run_cuda_step(case, DEV), run_oracle_step(case)))


def test_grouped_scatter_dense_runs():
    """Many samples per voxel (16-point tiles inside ONE node at every level, runs crossing tile borders)."""
    case = make_case(n_points=2500, n_batch=64, feat_levels=3, seed=77)
    rng = np.random.default_rng(5)
    base = case["coord"][rng.integers(0, 64, size=24)]
    coord = (base[:, None, :] + rng.uniform(-2e-4, 2e-4, size=(24, 200, 3))).reshape(-1, 3).astype(np.float32)
    case["coord"] = coord
    case["label"] = rng.uniform(-0.05, 0.05, size=coord.shape[0]).astype(np.float32)
    case["weight"] = np.ones(coord.shape[0], dtype=np.float32)
    case = sort_case_morton(case)
    print(compare_step(run_cuda_step(case, DEV, morton_ordered=True), run_oracle_step(case)))


@pytest.mark.parametrize("n_batch", [0, 1, 15, 16, 17, 255])
def test_ragged_batch_sizes(n_batch):
    """Tile tails: 16 stragglers are always appended, so N = n_batch + 16 covers 16..271."""
    case = make_case(n_points=1500, n_batch=n_batch, feat_levels=2, seed=5)
    print(compare_step(run_cuda_step(case, DEV), run_oracle_step(case)))


def test_two_pass_mode_matches_oracle():
    case = make_case(n_points=2000, n_batch=2000, feat_levels=3, seed=21)
    print(compare_step(run_cuda_step(case, DEV, single_pass=False), run_oracle_step(case)))


def test_unfused_class_surface_matches_oracle():
    """query_feature kernel + torch MLP + torch loss (the reference call sequence verbatim)."""
    case = make_case(n_points=2000, n_batch=2000, feat_levels=4, seed=22)
    print(compare_step(run_cuda_step(case, DEV, unfused=True), run_oracle_step(case)))


@pytest.mark.parametrize("feature_dim", [4, 16])
def test_class_surface_other_feature_dims(feature_dim):
    """The generic query kernels take any feature_dim = 4*LP (here LP = 1 and 4); the fused decoder kernel is F=8 only
    and must say so."""
    from shine_mapping_b200 import _abi, sdf_bce_step
    case = make_case(n_points=2000, n_batch=2000, feat_levels=3, seed=25, feature_dim=feature_dim)
    print(compare_step(run_cuda_step(case, DEV, unfused=True), run_oracle_step(case)))
    cfg, octree, dec = build_cuda_models(case, DEV)
    with pytest.raises(_abi.ShineB200Error, match="feature_dim=8"):
        sdf_bce_step(octree, dec, torch.from_numpy(case["coord"]).to(DEV), torch.from_numpy(case["label"]).to(DEV), 0.01)


def test_plain_tf32_flag_is_close():
    case = make_case(n_points=2000, n_batch=2000, feat_levels=2, seed=23)
    got, want = run_cuda_step(case, DEV, tf32x1=True), run_oracle_step(case)
    print(compare_step(got, want, pred_atol=5e-3, pred_rtol=5e-3, grad_rel=3e-2))


def test_frozen_decoder_gives_only_table_grads():
    from shine_mapping_b200 import sdf_bce_step
    case = make_case(n_points=2000, n_batch=2000, feat_levels=4, seed=24)
    cfg, octree, dec = build_cuda_models(case, DEV, freeze_decoder=True)
    coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
    loss = sdf_bce_step(octree, dec, coord, label, case["cfg"]["sigma"])
    loss.backward()
    want = run_oracle_step(case)
    for p, w in zip(octree.hier_features, want["table_grads"]):
        g = p.grad.cpu().numpy()[:-1]
        assert np.abs(g - w[:-1]).max() <= 2e-4 * np.abs(w).max() + 1e-10
    assert all(p.grad is None for p in dec.parameters())


def test_empty_batch_is_ok():
    from shine_mapping_b200 import sdf_infer
    case = make_case(n_points=1500, n_batch=0, feat_levels=2, seed=5)
    cfg, octree, dec = build_cuda_models(case, DEV)
    empty = torch.empty(0, 3, device=DEV)
    assert octree.query_feature(empty).shape == (0, 8)
    assert [t.shape for t in octree.get_indices(empty)] == [(0, 8)] * 2
    assert sdf_infer(octree, dec, empty).shape == (0,)


def test_infer_and_mask_match_step_pred():
    from shine_mapping_b200 import sdf_infer
    case = make_case(n_points=2500, n_batch=3000, feat_levels=3, seed=31)
    cfg, octree, dec = build_cuda_models(case, DEV)
    coord = torch.from_numpy(case["coord"]).to(DEV)
    want = run_oracle_step(case)
    for lvl in range(3):
        pred, mask = sdf_infer(octree, dec, coord, mask_level=lvl)
        assert np.abs(pred.cpu().numpy() - want["pred"]).max() < 2e-5
        assert np.array_equal(mask.cpu().numpy(), (want["indices"][lvl] >= 0).all(1))
    case6 = make_case(n_points=2500, n_batch=3000, feat_levels=7, seed=37)       # LMAX = 8 instantiation
    cfg6, octree6, dec6 = build_cuda_models(case6, DEV)
    want6 = run_oracle_step(case6)
    pred6, mask6 = sdf_infer(octree6, dec6, torch.from_numpy(case6["coord"]).to(DEV), mask_level=5)
    assert np.abs(pred6.cpu().numpy() - want6["pred"]).max() < 2e-5
    assert np.array_equal(mask6.cpu().numpy(), (want6["indices"][5] >= 0).all(1))


@pytest.mark.parametrize("levels,n_batch", [(4, 3000), (2, 100), (3, 40000), (6, 3000)])
def test_tcgen05_infer_matches_oracle(levels, n_batch):
    """tcgen05.mma / TMEM decoder (SHINE_FLAG_TCGEN05) vs the oracle and vs the mma.sync kernel, incl. the mask."""
    from shine_mapping_b200 import sdf_infer
    case = make_case(n_points=2500, n_batch=n_batch, feat_levels=levels, seed=90 + levels)
    cfg, octree, dec = build_cuda_models(case, DEV)
    coord = torch.from_numpy(case["coord"]).to(DEV)
    want = run_oracle_step(case)
    pred, mask = sdf_infer(octree, dec, coord, mask_level=0, tcgen05=True)
    torch.cuda.synchronize()
    assert np.abs(pred.cpu().numpy() - want["pred"]).max() < 2e-5
    assert np.array_equal(mask.cpu().numpy(), (want["indices"][0] >= 0).all(1))
    ref = sdf_infer(octree, dec, coord)
    assert (pred - ref).abs().max() < 1e-5


@pytest.mark.parametrize("levels,poly", [(2, True), (4, True), (3, False), (6, True)])
def test_eikonal_through_class_surface_matches_oracle(levels, poly):
    """ekional_loss_on (reference shine_batch.py:141-142,183-185): d pred / d coord with create_graph=True through
    query_feature's coordinate-gradient kernels, and the second backward through the tangent kernels."""
    from oracle import shine_oracle as orc
    from shine_mapping_b200 import SdfTrainer
    from shine_mapping_b200.batch_loop import eikonal_iteration
    from tests.parity_utils import oracle_from_case
    case = make_case(n_points=2000, n_batch=1500, feat_levels=levels, seed=100 + levels, poly=poly)
    cfg, octree, dec = build_cuda_models(case, DEV)
    cfg.ekional_loss_on, cfg.weight_e = True, 0.1
    coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
    weight = torch.from_numpy(case["weight"]).to(DEV)
    tr = SdfTrainer(cfg, octree, dec)
    tr.zero_grad()
    total, eik, g = eikonal_iteration(cfg, octree, dec, tr, coord, label, weight)
    o, odec = oracle_from_case(case)
    want = orc.train_step_eikonal(o, odec, torch.from_numpy(case["coord"]), torch.from_numpy(case["label"]),
                                  torch.from_numpy(case["weight"]), case["cfg"]["sigma"], 0.1)
    gw = want["g"].numpy()
    assert np.abs(g.cpu().numpy() - gw).max() <= 1e-4 * np.abs(gw).max() + 1e-7
    assert abs(float(eik) - float(want["eikonal"])) <= 1e-4 * abs(float(want["eikonal"])) + 1e-7
    assert abs(float(total) - float(want["loss"])) <= 1e-4 * abs(float(want["loss"]))
    for k, gt in enumerate(want["table_grads"]):
        got = tr.table_grads[k].cpu().numpy()
        assert np.abs(got - gt.numpy())[:-1].max() <= 1e-3 * np.abs(gt.numpy()).max() + 1e-9, k
    for name, p in zip(DEC_KEYS, dec.fused_params()):
        gt = want["dec_grads"][name].numpy()
        assert np.abs(p.grad.cpu().numpy() - gt).max() <= 1e-3 * np.abs(gt).max() + 1e-9, name


def test_eikonal_matches_reference_golden():
    """The CUDA eikonal path against the outputs of the reference's own classes (tests/golden/ref_eikonal_l3.npz)."""
    import json
    import os
    from shine_mapping_b200 import SdfTrainer
    from shine_mapping_b200.batch_loop import eikonal_iteration
    from tests.parity_utils import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "ref_eikonal_l3.npz"))
    cfg_j = json.loads(str(z["cfg_json"]))
    case = {"cfg": cfg_j, "frames": [z["frame_0"]], "tables": [z[f"table_{k}"] for k in range(cfg_j["tree_level_feat"])],
            "dec": {k: z["dec_" + k] for k in DEC_KEYS}, "coord": z["coord"], "label": z["label"], "weight": z["weight"]}
    cfg, octree, dec = build_cuda_models(case, DEV)
    cfg.ekional_loss_on, cfg.weight_e = True, cfg_j["weight_e"]
    tr = SdfTrainer(cfg, octree, dec)
    tr.zero_grad()
    total, eik, g = eikonal_iteration(cfg, octree, dec, tr, torch.from_numpy(z["coord"]).to(DEV),
                                      torch.from_numpy(z["label"]).to(DEV), torch.from_numpy(z["weight"]).to(DEV))
    assert np.abs(g.cpu().numpy() - z["exp_g"]).max() <= 1e-4 * np.abs(z["exp_g"]).max() + 1e-7
    assert abs(float(eik) - float(z["exp_eikonal"])) <= 1e-4 * abs(float(z["exp_eikonal"]))
    assert abs(float(total) - float(z["exp_loss"])) <= 1e-4 * abs(float(z["exp_loss"]))
    for k in range(cfg_j["tree_level_feat"]):
        want = z[f"exp_tgrad_{k}"]
        assert np.abs(tr.table_grads[k].cpu().numpy() - want)[:-1].max() <= 1e-3 * np.abs(want).max() + 1e-9
    for name, p in zip(DEC_KEYS, dec.fused_params()):
        want = z["exp_dgrad_" + name]
        assert np.abs(p.grad.cpu().numpy() - want).max() <= 1e-3 * np.abs(want).max() + 1e-9


def _fused_eikonal(case, weight_e):
    from shine_mapping_b200 import SdfTrainer
    cfg, octree, dec = build_cuda_models(case, DEV)
    cfg.ekional_loss_on, cfg.weight_e = True, weight_e
    coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
    weight = torch.from_numpy(case["weight"]).to(DEV)
    tr = SdfTrainer(cfg, octree, dec)
    tr.zero_grad()
    g = torch.empty(coord.shape[0], 3, device=DEV)
    bce, eik = tr.forward_backward_eikonal(coord, label, weight, grad_out=g)
    torch.cuda.synchronize()
    return tr, dec, float(bce) + weight_e * float(eik), float(eik), g.cpu().numpy()


@pytest.mark.parametrize("levels,poly", [(2, True), (3, False), (4, True)])
def test_fused_eikonal_step_matches_oracle(levels, poly):
    """ONE launch (shine_sdf_bce_eikonal_step) against the oracle's autograd double backward."""
    from oracle import shine_oracle as orc
    from tests.parity_utils import oracle_from_case
    case = make_case(n_points=2000, n_batch=1500, feat_levels=levels, seed=100 + levels, poly=poly)
    tr, dec, total, eik, g = _fused_eikonal(case, 0.1)
    o, odec = oracle_from_case(case)
    want = orc.train_step_eikonal(o, odec, torch.from_numpy(case["coord"]), torch.from_numpy(case["label"]),
                                  torch.from_numpy(case["weight"]), case["cfg"]["sigma"], 0.1)
    gw = want["g"].numpy()
    assert np.abs(g - gw).max() <= 1e-4 * np.abs(gw).max() + 1e-7
    assert abs(eik - float(want["eikonal"])) <= 1e-4 * abs(float(want["eikonal"])) + 1e-7
    assert abs(total - float(want["loss"])) <= 1e-4 * abs(float(want["loss"]))
    for k, gt in enumerate(want["table_grads"]):
        got = tr.table_grads[k].cpu().numpy()
        assert np.abs(got - gt.numpy())[:-1].max() <= 1e-3 * np.abs(gt.numpy()).max() + 1e-9, k
    for name, gd in zip(DEC_KEYS, tr.dec_grads):
        gt = want["dec_grads"][name].numpy()
        assert np.abs(gd.cpu().numpy() - gt).max() <= 1e-3 * np.abs(gt).max() + 1e-9, name


def test_fused_eikonal_step_matches_reference_golden_and_beats_class_surface():
    """The fused entry against the outputs of the reference's own classes (tests/golden/ref_eikonal_l3.npz), and its
    time against the class-surface composition (query kernels + cuBLAS + autograd double backward)."""
    import json
    import os
    from shine_mapping_b200 import SdfTrainer
    from shine_mapping_b200.batch_loop import eikonal_iteration
    from tests.parity_utils import GOLDEN_DIR
    z = np.load(os.path.join(GOLDEN_DIR, "ref_eikonal_l3.npz"))
    cfg_j = json.loads(str(z["cfg_json"]))
    case = {"cfg": cfg_j, "frames": [z["frame_0"]], "tables": [z[f"table_{k}"] for k in range(cfg_j["tree_level_feat"])],
            "dec": {k: z["dec_" + k] for k in DEC_KEYS}, "coord": z["coord"], "label": z["label"], "weight": z["weight"]}
    tr, dec, total, eik, g = _fused_eikonal(case, cfg_j["weight_e"])
    assert np.abs(g - z["exp_g"]).max() <= 1e-4 * np.abs(z["exp_g"]).max() + 1e-7
    assert abs(eik - float(z["exp_eikonal"])) <= 1e-4 * abs(float(z["exp_eikonal"]))
    assert abs(total - float(z["exp_loss"])) <= 1e-4 * abs(float(z["exp_loss"]))
    for k in range(cfg_j["tree_level_feat"]):
        want = z[f"exp_tgrad_{k}"]
        assert np.abs(tr.table_grads[k].cpu().numpy() - want)[:-1].max() <= 1e-3 * np.abs(want).max() + 1e-9
    for name, gd in zip(DEC_KEYS, tr.dec_grads):
        want = z["exp_dgrad_" + name]
        assert np.abs(gd.cpu().numpy() - want).max() <= 1e-3 * np.abs(want).max() + 1e-9
    # timing at the KITTI batch size (config/kitti/kitti_batch.yaml: batch_size 16384)
    big = make_case(n_points=3000, n_batch=16384, feat_levels=4, seed=7)
    cfg, octree, dec2 = build_cuda_models(big, DEV)
    cfg.ekional_loss_on, cfg.weight_e = True, 0.1
    coord = torch.from_numpy(big["coord"]).to(DEV); label = torch.from_numpy(big["label"]).to(DEV)
    weight = torch.from_numpy(big["weight"]).to(DEV)
    tr2 = SdfTrainer(cfg, octree, dec2)

    def timed(fn, reps=10):
        for _ in range(3):
            tr2.zero_grad(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            tr2.zero_grad(); fn()
        e1.record(); torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    t_fused = timed(lambda: tr2.forward_backward_eikonal(coord, label, weight))
    t_class = timed(lambda: eikonal_iteration(cfg, octree, dec2, tr2, coord, label, weight))
    print(f"eikonal step, {coord.shape[0]} points: fused {t_fused:.3f} ms, class surface {t_class:.3f} ms, x{t_class / t_fused:.1f}")
    assert t_class > 2.0 * t_fused


def test_points_to_morton_bit_exact():
    from shine_mapping_b200 import _abi
    from oracle import shine_oracle as orc
    g = torch.Generator().manual_seed(3)
    x = (torch.rand(200000, 3, generator=g) * 2.4 - 1.2)
    x[:8] = torch.tensor([[-1, -1, -1], [1, 1, 1], [0, 0, 0], [1 - 2 ** -12, 2 ** -12, -2 ** -12], [0.5, -0.5, 0.25],
                          [-1.0000001, 0.99999994, 0.9999999], [3, -3, 0], [2 ** -11, 2 ** -10, 2 ** -9]])
    xd = x.to(DEV).contiguous()
    for level in (1, 5, 9, 12, 15):
        out = torch.empty(x.shape[0], dtype=torch.int64, device=DEV)
        _abi.check(_abi.lib().shine_points_to_morton(_abi.ptr(xd), x.shape[0], level, _abi.ptr(out),
                                                     _abi.stream_ptr()), "morton")
        want = orc.points_to_morton(orc.quantize_points(x.numpy(), level))
        assert np.array_equal(out.cpu().numpy(), want), level


# ---- size-independent properties at BASELINE scale ------------------------------------------------------

def test_large_batch_properties():
    """1M points, L=4: (a) interpolation weights sum to 1 => per level, the column sums of the table gradient
    equal the column sums of dL/dfeature over the points that hit that level; (b) two runs give identical
    indices and (up to atomic ordering) identical gradients; (c) pred of the step == pred of the inference
    kernel bit-for-bit."""
    from shine_mapping_b200 import SdfTrainer, sdf_infer, synth
    from tests.parity_utils import make_config
    from shine_mapping_b200 import Decoder, FeatureOctree
    torch.manual_seed(1)
    cfg = make_config(4, device=DEV, pc_radius=50.0)
    octree, dec = FeatureOctree(cfg), Decoder(cfg)
    pool = synth.build_scene_map(cfg, octree, n_azimuth=512, n_frames=2, seed=1, device=DEV)
    n = 1 << 20
    coord, label, weight = pool.get_batch(n)
    tr = SdfTrainer(cfg, octree, dec)
    pred = torch.empty(n, device=DEV)
    tr.zero_grad(); tr.forward_backward(coord, label, weight, pred_out=pred)
    g1 = tr.flat_grad.clone()
    tr.zero_grad(); tr.forward_backward(coord, label, weight)
    g2 = tr.flat_grad.clone()
    assert (g1 - g2).abs().max() <= 1e-4 * g1.abs().max()
    assert torch.equal(pred, sdf_infer(octree, dec, coord))
    # (a) with the unfused autograd path giving dL/dfeature
    feat = octree.query_feature(coord).detach().requires_grad_(True)
    loss = torch.nn.functional.binary_cross_entropy_with_logits(
        dec.sdf(feat), torch.sigmoid(label / cfg.sigma_sigmoid))
    loss.backward()
    idx = octree.get_indices(coord)
    for i in range(4):
        hit = (idx[i] >= 0).all(1)
        want = feat.grad[hit].double().sum(0)
        got = tr.table_grads[4 - i - 1].double().sum(0)
        assert torch.allclose(got, want, rtol=2e-3, atol=1e-7), (i, got, want)
    # loss value agrees with the torch composition
    assert abs(float(tr.loss) - float(loss)) <= 2e-5 * abs(float(loss))


# ---- optimizer + trainer ---------------------------------------------------------------------------------

def test_adam_kernel_matches_torch_adam():
    from shine_mapping_b200 import SdfTrainer
    case = make_case(n_points=2000, n_batch=4096, feat_levels=3, seed=41)
    cfg, octree, dec = build_cuda_models(case, DEV)
    cfg.lr, cfg.weight_decay, cfg.lr_level_reduce_ratio = 0.01, 1e-7, 0.7
    cfg2, octree2, dec2 = build_cuda_models(case, DEV)
    coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
    tr = SdfTrainer(cfg, octree, dec)
    # torch reference optimiser with the reference grouping (utils/tools.py:57-83)
    groups = [{"params": list(dec2.fused_params()), "lr": cfg.lr, "weight_decay": cfg.weight_decay}]
    lr = cfg.lr
    feats = list(octree2.parameters())
    for i in range(3):
        groups.append({"params": feats[3 - i - 1], "lr": lr}); lr *= cfg.lr_level_reduce_ratio
    opt = torch.optim.Adam(groups, betas=(0.9, 0.99), eps=cfg.adam_eps)
    from shine_mapping_b200 import sdf_bce_step
    for it in range(5):
        tr.zero_grad() if it == 0 else None
        tr.train_step(coord, label)
        opt.zero_grad(set_to_none=True)
        sdf_bce_step(octree2, dec2, coord, label, cfg.sigma_sigmoid).backward()
        opt.step()
    for a, b in zip(octree.hier_features, octree2.hier_features):
        assert torch.allclose(a[:-1], b[:-1], rtol=2e-3, atol=2e-5), (a[:-1] - b[:-1]).abs().max()
    for a, b in zip(dec.fused_params(), dec2.fused_params()):
        assert torch.allclose(a, b, rtol=2e-3, atol=2e-5)


def test_training_reduces_loss():
    from shine_mapping_b200 import SdfTrainer
    case = make_case(n_points=3000, n_batch=8192, feat_levels=4, seed=42)
    cfg, octree, dec = build_cuda_models(case, DEV)
    cfg.lr = 0.01
    coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
    tr = SdfTrainer(cfg, octree, dec)
    tr.zero_grad()
    first = float(tr.train_step(coord, label))
    for _ in range(60):
        last = float(tr.train_step(coord, label))
    assert last < 0.9 * first, (first, last)


def test_cpu_tensor_is_rejected_loudly():
    from shine_mapping_b200 import _abi
    case = make_case(n_points=1500, n_batch=10, feat_levels=2, seed=5)
    cfg, octree, dec = build_cuda_models(case, DEV)
    with pytest.raises(_abi.ShineB200Error):
        octree.query_feature(torch.zeros(4, 3))


def test_abi_rejects_bad_arguments():
    from shine_mapping_b200 import _abi
    lib = _abi.lib()
    d = _abi.ShineOctree()
    d.num_levels = 0
    assert lib.shine_query_fwd(C.byref(d), None, 4, None, None) == -1
    assert b"invalid" in lib.shine_error_string(-1)


# ---- the tcgen05 / TMEM training kernel (csrc/shine_train_tc.cu) ---------------------------------------------------------

def _trainer_step(case, tcgen05, freeze=False):
    from shine_mapping_b200 import SdfTrainer
    cfg, octree, dec = build_cuda_models(case, DEV, freeze_decoder=freeze)
    coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
    weight = torch.from_numpy(case["weight"]).to(DEV)
    tr = SdfTrainer(cfg, octree, dec, tcgen05=tcgen05)
    tr.zero_grad()
    pred = torch.empty(coord.shape[0], device=DEV)
    loss = tr.forward_backward(coord, label, weight, pred_out=pred)
    torch.cuda.synchronize()
    return {
        "indices": [t.cpu().numpy() for t in octree.get_indices(coord)],
        "feature": octree.query_feature(coord).detach().cpu().numpy(),
        "pred": pred.cpu().numpy(), "loss": float(loss),
        "table_grads": [g.detach().cpu().numpy().copy() for g in tr.table_grads],
        "dec_grads": {} if freeze else {k: g.detach().cpu().numpy().copy() for k, g in zip(DEC_KEYS, tr.dec_grads)},
    }


@pytest.mark.parametrize("levels,poly,weighted,reduction,n_batch", [
    (4, True, False, "mean", 3000), (2, True, False, "mean", 100), (3, False, True, "sum", 5000),
    (4, True, True, "mean", 60000),          # > 148 tiles of 128 points: several rounds per CTA, both gather groups busy
    (1, True, False, "mean", 0),             # 16 stragglers only: one partial tile
])
def test_tcgen05_train_step_matches_oracle(levels, poly, weighted, reduction, n_batch):
    """SHINE_FLAG_TCGEN05 on shine_sdf_bce_step: decoder forward / dgrad / wgrad as tcgen05.mma on 128-point tiles
    (operands in shared memory, accumulators in TMEM), same tolerances as the mma.sync kernel."""
    from tests.parity_utils import drop_relu_kink_points
    case = make_case(n_points=2500, n_batch=n_batch, feat_levels=levels, seed=40 + levels, poly=poly, weighted=weighted,
                     reduction=reduction, n_frames=2 if n_batch > 10000 else 1)
    case, dropped = drop_relu_kink_points(case)
    print("points on a ReLU kink dropped:", dropped, compare_step(_trainer_step(case, True), run_oracle_step(case)))


def test_tcgen05_train_step_frozen_decoder():
    case = make_case(n_points=2500, n_batch=4000, feat_levels=4, seed=47)
    got = _trainer_step(case, True, freeze=True)
    want = run_oracle_step(case)
    want["dec_grads"] = {}
    print(compare_step(got, want))


def test_two_queries_before_one_backward():
    """the reference queries `coord_near` while the first query's graph is still alive (shine_batch.py:155-160);
    re-zeroing the trash row inside query_feature must not invalidate the tensors autograd saved for the first query."""
    case = make_case(n_points=1500, n_batch=600, feat_levels=3, seed=12)
    cfg, octree, dec = build_cuda_models(case, DEV)
    c = torch.from_numpy(case["coord"]).to(DEV)
    f1 = octree.query_feature(c[:300])
    f2 = octree.query_feature(c[300:600] + 1e-4)
    (f1.sum() + 2.0 * f2.sum()).backward()
    g = [p.grad.clone() for p in octree.hier_features]
    for p in octree.hier_features:
        p.grad = None
    octree.query_feature(c[:300]).sum().backward()
    ga = [p.grad.clone() for p in octree.hier_features]
    for p in octree.hier_features:
        p.grad = None
    (2.0 * octree.query_feature(c[300:600] + 1e-4).sum()).backward()
    for a, b, p in zip(g, ga, octree.hier_features):
        assert torch.allclose(a, b + p.grad, rtol=1e-5, atol=1e-7)
