import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import __graft_entry__ as ge; ge.build()
from tests.parity_utils import make_case, run_oracle_step, DEC_KEYS, build_cuda_models
from shine_mapping_b200 import SdfTrainer
DEV = "cuda:0"
for n_batch in (3000, 60000):
    case = make_case(n_points=2500, n_batch=n_batch, feat_levels=4, seed=44, n_frames=2 if n_batch > 10000 else 1)
    want = run_oracle_step(case)
    for tc in (False, True):
        cfg, octree, dec = build_cuda_models(case, DEV)
        coord = torch.from_numpy(case["coord"]).to(DEV); label = torch.from_numpy(case["label"]).to(DEV)
        tr = SdfTrainer(cfg, octree, dec, tcgen05=tc); tr.zero_grad()
        loss = tr.forward_backward(coord, label, None); torch.cuda.synchronize()
        print(f"n={coord.shape[0]} tcgen05={tc} loss {float(loss):.7f} want {want['loss']:.7f}")
        for k, g in zip(DEC_KEYS, tr.dec_grads):
            w = want["dec_grads"][k]; g = g.detach().cpu().numpy()
            print(f"   {k:18s} max|got| {np.abs(g).max():.4e} max|want| {np.abs(w).max():.4e} rel {np.abs(g - w).max() / np.abs(w).max():.3e}")
        if tc:
            g = tr.dec_grads[0].detach().cpu().numpy(); w = want["dec_grads"]["layers.0.weight"]
            print("   dW1 got[0:3]", g[:3]); print("   dW1 want[0:3]", w[:3])
            print("   dW1 got^T-like? ", np.abs(g - w).max(), "col-perm check", [float(np.abs(g[:, (c + 4) % 8] - w[:, c]).max()) for c in range(2)])
        for lvl, (a, b) in enumerate(zip(tr.table_grads, want["table_grads"])):
            a = a.detach().cpu().numpy()
            print(f"   table {lvl} rel {np.abs(a[:-1] - b[:-1]).max() / np.abs(b).max():.3e}")
