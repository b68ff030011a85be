"""The plain-C restatement (oracle/shine_oracle.c) against the numpy/torch oracle and the reference-run goldens: two
independently written CPU restatements must agree bit-exactly on indices and to fp32 round-off on values."""
import numpy as np
import pytest

from oracle import c_oracle
from tests.parity_utils import GOLDEN_NAMES, load_golden, make_case, oracle_from_case, orc, run_oracle_step


def _check(case, want):
    o, _ = oracle_from_case(case)
    got = c_oracle.forward(o, case["tables"], case["dec"], case["coord"], case["label"], case["cfg"]["sigma"],
                           case["cfg"]["poly_int_on"])
    for a, b in zip(got["indices"], want["indices"]):
        assert np.array_equal(a, b)
    assert np.abs(got["feature"] - want["feature"]).max() <= 2e-6
    assert np.abs(got["pred"] - want["pred"]).max() <= 2e-5
    return got


@pytest.mark.parametrize("name", GOLDEN_NAMES)
def test_c_oracle_matches_reference_goldens(name):
    case, exp = load_golden(name)
    got = _check(case, exp)
    c = case["cfg"]
    if not c["weighted"]:
        n = case["coord"].shape[0]
        loss = got["loss_sum"] / n if c["reduction"] == "mean" else got["loss_sum"]
        assert abs(loss - exp["loss"]) <= 1e-5 * abs(exp["loss"])


@pytest.mark.parametrize("levels,poly", [(1, True), (4, False), (6, True)])
def test_c_oracle_matches_python_oracle(levels, poly):
    case = make_case(n_points=1500, n_batch=800, feat_levels=levels, seed=200 + levels, poly=poly, n_frames=2)
    want = run_oracle_step(case)
    got = _check(case, want)
    assert abs(got["loss_sum"] / case["coord"].shape[0] - want["loss"]) <= 1e-5 * abs(want["loss"])


def test_c_morton_matches_numpy_on_edge_coordinates():
    rng = np.random.default_rng(0)
    x = (rng.random((50000, 3)) * 2.6 - 1.3).astype(np.float32)
    x[:6] = [[-1, -1, -1], [1, 1, 1], [0, 0, 0], [0.99999994, -0.99999994, 1.0000001], [3, -3, 0.5], [2 ** -12, -2 ** -12, 0]]
    for level in (1, 7, 12, 15):
        assert np.array_equal(c_oracle.points_to_morton(x, level), orc.points_to_morton(orc.quantize_points(x, level)))
