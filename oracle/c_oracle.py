"""ctypes wrapper of oracle/shine_oracle.c (plain-C forward restatement).  *** TEST INFRASTRUCTURE ***

`build()` compiles it with gcc into oracle/_build/ (git-ignored; travels to the GPU box with the snapshot).
`forward(case)` runs quantise -> Morton -> sorted-key lookup -> blend -> MLP -> BCE on a parity case dict
(tests/parity_utils.py) whose node tables come from an OracleOctree replay."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "shine_oracle.c")
LIB = os.path.join(HERE, "_build", "libshine_oracle_c.so")
_lib = None


def build(force: bool = False) -> str:
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(SRC):
        os.makedirs(os.path.dirname(LIB), exist_ok=True)
        subprocess.run(["gcc", "-O2", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"], check=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.orc_forward.restype = C.c_double
    return _lib


def points_to_morton(xyz: np.ndarray, level: int) -> np.ndarray:
    xyz = np.ascontiguousarray(xyz, dtype=np.float32)
    out = np.empty(xyz.shape[0], dtype=np.int64)
    lib().orc_points_to_morton(xyz.ctypes.data_as(C.c_void_p), C.c_int64(xyz.shape[0]), C.c_int32(level),
                               out.ctypes.data_as(C.c_void_p))
    return out


def forward(oracle_octree, tables, dec: dict, coord: np.ndarray, label: np.ndarray, sigma: float, poly: bool):
    """-> dict(indices [L][n,8], feature [n,F], pred [n], loss_sum).  `oracle_octree`: an OracleOctree whose dict node
    tables define the map; `tables`: coarse->fine feature arrays; `dec`: decoder arrays keyed like the state dict."""
    L, W = oracle_octree.featured_level_num, oracle_octree.max_level
    coord = np.ascontiguousarray(coord, dtype=np.float32)
    label = np.ascontiguousarray(label, dtype=np.float32)
    n, F = coord.shape[0], tables[0].shape[1]
    keys, ids, tabs = [], [], []
    for i in range(L):
        t = oracle_octree.nodes_lookup_tables[W - i]
        k = np.array(sorted(t.keys()), dtype=np.int64)
        keys.append(k)
        ids.append(np.ascontiguousarray(np.array([t[int(m)] for m in k], dtype=np.int32).reshape(-1, 8)))
        tabs.append(np.ascontiguousarray(tables[L - i - 1], dtype=np.float32))
    ptr = lambda a: a.ctypes.data_as(C.c_void_p)   # noqa: E731
    arr = lambda xs: (C.c_void_p * len(xs))(*[x.ctypes.data for x in xs])   # noqa: E731
    nkeys = np.array([k.shape[0] for k in keys], dtype=np.int64)
    w = {k: np.ascontiguousarray(v, dtype=np.float32) for k, v in dec.items()}
    H = w["layers.0.weight"].shape[0]
    idx = np.empty((L, n, 8), dtype=np.int64); feat = np.empty((n, F), dtype=np.float32); pred = np.empty(n, dtype=np.float32)
    loss = lib().orc_forward(ptr(coord), ptr(label), C.c_int64(n), C.c_int32(W), C.c_int32(L), C.c_int32(F),
                             C.c_int32(1 if poly else 0), arr(keys), ptr(nkeys), arr(ids), arr(tabs),
                             ptr(w["layers.0.weight"]), ptr(w["layers.0.bias"]), ptr(w["layers.1.weight"]),
                             ptr(w["layers.1.bias"]), ptr(w["lout.weight"]), ptr(w["lout.bias"]), C.c_int32(H),
                             C.c_float(sigma), ptr(idx), ptr(feat), ptr(pred))
    return {"indices": [idx[i] for i in range(L)], "feature": feat, "pred": pred, "loss_sum": float(loss)}
