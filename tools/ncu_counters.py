#!/usr/bin/env python
"""profiles/r02_kernel_counters.json from the two `ncu --set full` captures of bench.py (C2 step, HBM-bound leg).

usage: python tools/ncu_counters.py c2=gpurun_out/r02_step_c2_grouped.ncu-rep c2_random=gpurun_out/r02_step_c2.ncu-rep \
           hbm=gpurun_out/r02_step_hbm.ncu-rep
(c2 = the kernel of the bench's timed steps: Morton-ordered batches, voxel-grouped scatter; c2_random = the general kernel
on batches in the order drawn; hbm = the general kernel on the 328 MB map).  Entries not named keep their old values.
Reads the raw page (`ncu -i REP --page raw --csv --print-units base`), keeps the per-launch counters bench.py turns into
rooflines and writes the details pages next to the json (profiles/r02_step_{c2,hbm}_details.csv).
"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
POINTS = {"c2": 776616, "c2_random": 776616, "hbm": 1048576}
SMS = 148


def raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--print-units", "base"], capture_output=True, text=True,
                         check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    head, last = rows[0], rows[-1]
    return dict(zip(head, last))


def num(d, key):
    if key not in d:                       # some counters only exist with a section prefix (SM_A.TriageCompute.<name>)
        key = next(k for k in d if k.endswith("." + key))
    return float(d[key].replace(",", ""))


def main():
    reps = dict(a.split("=", 1) for a in sys.argv[1:])
    out_path = os.path.join(ROOT, "profiles", "r02_kernel_counters.json")
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    for name, rep in reps.items():
        d = raw(rep)
        pts = POINTS[name]
        rd, wr = num(d, "dram__bytes_read.sum"), num(d, "dram__bytes_write.sum")
        res[name] = {
            "kernel": d["Kernel Name"], "points": pts,
            "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_point": (rd + wr) / pts,
            "lsu_wavefronts_per_point": num(d, "l1tex__data_pipe_lsu_wavefronts.avg") * SMS / pts,
            "duration_us_under_ncu": num(d, "gpu__time_duration.sum") / 1e3,
            "lsu_wavefront_pct_of_peak": num(d, "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            "inst_executed": num(d, "smsp__inst_executed.sum"),
            "registers": num(d, "launch__registers_per_thread"),
            "l2_hit_pct": num(d, "lts__t_sector_hit_rate.pct"),
            "tensor_pipe_pct": num(d, "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
            "stall_cycles_per_issue": {k.split("issue_stalled_")[1].split("_per_issue")[0]: round(num(d, k), 3) for k in d
                                       if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")
                                       and "not_issued" not in k and num(d, k) >= 0.1},
            "source": f"profiles/r02_step_{name}_details.csv (ncu --set full of bench.py, gpurun {os.path.basename(rep)}; raw page)",
        }
        det = subprocess.run(["ncu", "-i", rep, "--page", "details", "--csv"], capture_output=True, text=True, check=True).stdout
        open(os.path.join(ROOT, "profiles", f"r02_step_{name}_details.csv"), "w").write(det)
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
