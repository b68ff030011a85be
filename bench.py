#!/usr/bin/env python
"""bench.py — SDF training points/sec (forward + backward) of the fused sm_100a step, with roofline,
CPU baseline and end-to-end (host buffers) numbers.  Contract: see the task statement / DESIGN.md §Measurement.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # our arm (N>1: launched by torchrun)
    python bench.py --impl reference [--steps K] [--warmup W]      # the reference's CPU algorithm (oracle port)

Workload (BASELINE.json configs[1], "C2"): one synthetic MaiCity-like HDL-64 scan (64 x 2048 rays) of the
analytic street scene -> reference-style ray samples (3 surface + 3 free per hit) -> 4-level octree (leaf 0.2 m,
world level 12, F=8) + geo_decoder_8dim architecture (8->32->32->1, random init, trainable) + BCE loss.
One step = fwd+bwd over a whole-scan batch (N = number of samples of the scan, drawn with torch.randint like
LiDARDataset.get_batch).  N>1 GPUs: every rank owns its own spatial block (its own scan 100 m further along the
street, own table shard) and only the 1 377 decoder gradients are all-reduced (BASELINE configs[4] layout) — weak scaling.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "sdf_train_points_per_sec_fwd_bwd"
UNIT = "points/s"
L, F = 4, 8
BYTES_PER_POINT = 24 + L * 40 + 3 * L * 8 * F * 4   # 3 256 B at L=4, F=8
HBM_FALLBACK_GBS = 6650.0


def workload_config(device):
    from shine_mapping_b200 import SHINEConfig
    return SHINEConfig(  # config/maicity/maicity_batch.yaml values, tree_level_feat 4 per BASELINE.json
        name="c2_maicity_like_single_scan", device=device, tree_level_world=12, tree_level_feat=L, leaf_vox_size=0.2,
        feature_dim=F, poly_int_on=True, surface_sample_range_m=0.15, surface_sample_n=3,
        free_sample_begin_ratio=0.3, free_sample_end_dist_m=0.8, free_sample_n=3, sigma_sigmoid_m=0.05,
        min_range=1.5, pc_radius=50.0, lr=0.01, weight_decay=1e-7, loss_weight_on=False, loss_reduction="mean",
        geo_mlp_level=2, geo_mlp_hidden_dim=32)


def build_workload(device, rank, world, n_azimuth):
    from shine_mapping_b200 import Decoder, FeatureOctree, synth
    cfg = workload_config(device)
    torch.manual_seed(42)   # decoder init identical on every rank (it is replicated)
    octree, decoder = FeatureOctree(cfg), Decoder(cfg)
    x0 = (rank - (world - 1) / 2.0) * 100.0
    pool = synth.build_scene_map(cfg, octree, n_azimuth=n_azimuth, n_frames=1, seed=42 + rank, device=device,
                                 origin_x0=x0)
    return cfg, octree, decoder, pool


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return HBM_FALLBACK_GBS, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (every 20 ms, timestamped; only the
    samples that fall inside [t0, t1] are reported)."""
    Q = ("timestamp,clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", "-i", str(index), f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "20"], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def wait_first_sample(self, timeout=5.0):
        t0 = time.time()
        while self.p is not None and time.time() - t0 < timeout:
            if os.path.getsize(self.f.name) > 0:
                return
            time.sleep(0.02)

    def stop(self, t0, t1):
        import datetime
        if self.p is None:
            return None
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush(); self.f.seek(0)
        sm, mx, reasons, total = [], 0.0, set(), 0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                clk, cmax = float(parts[1]), float(parts[2])
            except ValueError:
                continue
            total += 1
            if ts < t0 - 0.02 or ts > t1 + 0.02:
                continue
            sm.append(clk); mx = max(mx, cmax)
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0, "samples_total": total}
        return {"sm_mhz": statistics.median(sm), "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm),
                "window_s": round(t1 - t0, 3)}


def oracle_from_octree(octree, decoder):
    """CPU oracle (Python dict tables + torch CPU tensors) holding the same map as `octree`."""
    from oracle import shine_oracle as orc
    o = orc.OracleOctree(octree.max_level, octree.featured_level_num, octree.feature_dim, octree.feature_std,
                         octree.polynomial_interpolation)
    o.nodes_lookup_tables = octree.nodes_lookup_tables
    o.corners_lookup_tables = octree.corners_lookup_tables
    o.hier_features = [p.detach().cpu().clone().requires_grad_(True) for p in octree.hier_features]
    dec = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in decoder.state_dict().items()
           if not k.startswith("nclass_out")}
    return orc, o, dec


def pick_threads(orc, o, dec, batch, sigma):
    """The oracle's torch-CPU ops are small: too many threads hurt.  Try a few counts, keep the fastest."""
    best, best_t = None, None
    cands = sorted({min(c, os.cpu_count() or 1) for c in (8, 16, 32, os.cpu_count() or 1)})
    for c in cands:
        torch.set_num_threads(c)
        orc.train_step(o, dec, *batch, sigma, False, "mean")
        t0 = time.perf_counter()
        orc.train_step(o, dec, *batch, sigma, False, "mean")
        dt = time.perf_counter() - t0
        if best_t is None or dt < best_t:
            best, best_t = c, dt
    torch.set_num_threads(best)
    return best


def time_oracle(orc, o, dec, batches, sigma, steps, warmup):
    ts = []
    for i in range(warmup + steps):
        c, l, w = batches[i % len(batches)]
        t0 = time.perf_counter()
        orc.train_step(o, dec, c, l, w, sigma, False, "mean")
        dt = time.perf_counter() - t0
        if i >= warmup:
            ts.append(dt)
    return ts


def run_reference(args):
    """The reference's own CPU algorithm for the path (oracle port: Python-dict Morton lookup + torch CPU
    gather/MLP/BCE/autograd, all host threads), on a bounded sample of the C2 workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    torch.set_num_threads(os.cpu_count() or 1)
    cfg, octree, decoder, pool = build_workload("cpu", 0, 1, args.n_azimuth)
    orc, o, dec = oracle_from_octree(octree, decoder)
    sample = args.ref_sample
    gen = torch.Generator().manual_seed(7)
    batches = [pool.get_batch(sample, gen) for _ in range(2)]
    pick_threads(orc, o, dec, tuple(t[:20000] for t in batches[0]), cfg.sigma_sigmoid)
    ts = time_oracle(orc, o, dec, batches, cfg.sigma_sigmoid, args.steps, args.warmup)
    sec = statistics.mean(ts)
    value = sample / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": cfg.name, "n_azimuth": args.n_azimuth, "pool_samples": len(pool),
                   "tree_level_feat": L, "feature_dim": F, "decoder": "geo_decoder_8dim arch 8-32-32-1",
                   "loss": "sdf_bce mean", "points_per_step": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"{sample} randint-drawn samples of the {len(pool)}-sample C2 scan per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_ours(args):
    from shine_mapping_b200 import SdfTrainer, _abi, dist as sdist
    rank, world, local = sdist.init_from_env("nccl")
    if world != args.gpus:
        if rank == 0:
            print(f"[bench] note: --gpus {args.gpus} but WORLD_SIZE={world}; using WORLD_SIZE", file=sys.stderr)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device: the hot path has no CPU fallback")
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    cfg, octree, decoder, pool = build_workload(str(dev), rank, world, args.n_azimuth)
    n = len(pool) if args.points <= 0 else args.points
    trainer = SdfTrainer(cfg, octree, decoder, shard_mode="spatial")
    n_global = n * world
    gen = torch.Generator(device=dev).manual_seed(1000 + rank)
    batches = [pool.get_batch(n, gen) for _ in range(4)]
    host = [tuple(t.cpu().pin_memory() for t in b[:2]) for b in batches]
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step(b):
        trainer.zero_grad()
        trainer.forward_backward(b[0], b[1], None, n_norm=n_global)
        trainer.all_reduce_grads()

    for i in range(max(args.warmup, 3)):
        step(batches[i % 4])
    torch.cuda.synchronize(dev)

    # ---- `value`: inputs resident in HBM, CUDA events on the launching stream, L2 flushed between steps ----
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.wait_first_sample()
    sdist.barrier(dev); torch.cuda.synchronize(dev)     # after the sampler start-up so that no rank enters late
    launches0 = _abi.LAUNCHES["count"]
    t_clk0 = time.time()
    if args.cuda_profiler:
        torch.cuda.profiler.start()                   # ncu --profile-from-start off: only the timed steps
    for k in range(args.steps):
        flush_buf.fill_(k & 0xFF)                     # > L2 (126 MB): evicts tables, inputs and gradients
        b = batches[k % 4]
        ev[k][0].record()
        trainer.zero_grad()
        ev[k][1].record()
        trainer.forward_backward(b[0], b[1], None, n_norm=n_global)
        ev[k][2].record()
        trainer.all_reduce_grads()
        ev[k][3].record()
    torch.cuda.synchronize(dev); sdist.barrier(dev)
    if args.cuda_profiler:
        torch.cuda.profiler.stop()
    launches = _abi.LAUNCHES["count"] - launches0
    step_ms = statistics.mean(e[0].elapsed_time(e[3]) for e in ev)
    kern_ms = statistics.mean(e[1].elapsed_time(e[2]) for e in ev)
    step_ms = sdist.max_over_ranks(step_ms, dev)
    kern_ms_max = sdist.max_over_ranks(kern_ms, dev)
    # nvidia-smi cannot sample faster than ~20 ms: keep the SAME steps running (not counted) so that the sampled
    # window under load is ~0.5 s.  The count is derived from the rank-agreed step time: every rank issues the same
    # number of collectives.
    n_cont = max(0, min(20000, int(500.0 / max(step_ms, 0.02)) - args.steps))
    for k in range(n_cont):
        step(batches[k % 4])
        if k % 64 == 63:
            torch.cuda.synchronize(dev)
    torch.cuda.synchronize(dev)
    clocks = sampler.stop(t_clk0, time.time()) if sampler else None
    value = n_global / (step_ms * 1e-3)

    # ---- `e2e`: host (pinned) buffers through SdfTrainer.step_from_host, wall clock incl. H2D + loss D2H -------
    for i in range(max(4, args.warmup)):            # every pinned host batch once: its CUDA graph is captured untimed
        trainer.step_from_host(*host[i % 4])
    e2e_ts = []
    sdist.barrier(dev); torch.cuda.synchronize(dev)
    for k in range(args.steps):
        flush_buf.fill_(k & 0xFF)
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        trainer.step_from_host(*host[k % 4])          # ends with loss.item(): device -> host read
        if world > 1:
            trainer.all_reduce_grads(); torch.cuda.synchronize(dev)
        e2e_ts.append(time.perf_counter() - t0)
    e2e_sec = sdist.max_over_ranks(statistics.mean(e2e_ts), dev)
    e2e_value = n_global / e2e_sec

    if rank != 0:
        return
    peak, peak_src = peaks()
    achieved = n * BYTES_PER_POINT / (kern_ms_max * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "r01_step_kernel_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = float(json.load(open(tpath))["dram_bytes_per_point"]) * n
        except Exception:
            traffic = None
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": cfg.name, "n_azimuth": args.n_azimuth, "points_per_step_per_gpu": n,
                   "global_points_per_step": n_global, "tree_level_feat": L, "feature_dim": F,
                   "table_rows": [int(p.shape[0]) for p in octree.hier_features],
                   "decoder": "geo_decoder_8dim arch 8-32-32-1, trainable, 3xTF32 mma.sync",
                   "loss": "sdf_bce mean", "parallelism": f"spatial-block x{world} + decoder-grad all-reduce",
                   "l2": "flushed between timed steps (256 MiB write, not timed)",
                   "step": "grad memset + fused fwd+loss+bwd kernel (+ all-reduce when N>1); no optimizer"},
        "roofline": {"bound": "hbm", "kernel": "sdf_fused_kernel<3,train,dec_grad,4>",
                     "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": traffic, "peak_source": peak_src, "kernel_ms": kern_ms_max,
                     "algorithmic_bytes_per_point": BYTES_PER_POINT},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": n * 16, "d2h_bytes_per_step": 4,
                "ms_per_step": e2e_sec * 1e3},
        "gpu_launches": launches,
        "clocks": clocks,
    }
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        orc, o, dec = oracle_from_octree(octree, decoder)
        sample = args.ref_sample
        cb = [tuple(t[:sample].cpu() for t in b) for b in batches[:2]]
        pick_threads(orc, o, dec, tuple(t[:20000] for t in cb[0]), cfg.sigma_sigmoid)
        ts = time_oracle(orc, o, dec, cb, cfg.sigma_sigmoid, 3, 1)
        line["cpu_baseline"] = {"value": sample / statistics.mean(ts), "unit": UNIT, "cores": torch.get_num_threads(),
                                "kind": "port", "sample": f"first {sample} points of the step's batch, 1 warm-up + 3 "
                                                           "timed oracle steps (Python-dict lookup + torch CPU autograd)"}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-azimuth", type=int, default=2048, help="rays per ring of the synthetic scan (C2: 2048)")
    ap.add_argument("--points", type=int, default=0, help="points per step per GPU (default: the whole scan)")
    ap.add_argument("--ref-sample", type=int, default=100000, help="points per oracle step (bounded CPU sample)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cuda-profiler", action="store_true", help="cudaProfilerStart/Stop around the timed steps")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        import contextlib
        import __graft_entry__ as ge
        if int(os.environ.get("LOCAL_RANK", "0")) == 0:
            with contextlib.redirect_stdout(sys.stderr):      # stdout carries exactly one JSON line
                ge.build()
        run_ours(args)
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
