"""Turn the raw ncu exports of one bench.py capture into the tracked summaries under profiles/.

    ncu -i gpurun_out/rNN_prof.ncu-rep --page raw --csv > gpurun_out/rNN_raw.csv
    python tools/make_profiles.py --raw gpurun_out/rNN_raw.csv [--raw ...] --launches gpurun_out/rNN_launches.csv --tag r1

Writes profiles/<tag>_ncu_full_summary.md, <tag>_ncu_dram_traffic_c2.json, <tag>_ncu_launches_c2.csv and
<tag>_ncu_launch_shares_c2.txt.  Several --raw files are merged (later files win per kernel).
"""
import argparse
import csv
import json
import os
import re
import shutil
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
METRICS = [
    "Block Size", "Grid Size", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def short(name):
    m = re.search(r"(?:jps::)?(\w+)\s*(?:<|\()", name)
    return m.group(1) if m else name


def read_raw(path):
    rows = list(csv.reader(open(path)))
    hdr, units = rows[0], rows[1]
    out = OrderedDict()
    for r in rows[2:]:
        d = {h: (v, u) for h, v, u in zip(hdr, r, units)}
        out[d["Kernel Name"][0]] = d
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--raw", action="append", required=True)
    ap.add_argument("--launches")
    ap.add_argument("--tag", default="r1")
    ap.add_argument("--suffix", default="c2", help="workload suffix of the output files (c2, c4_rank, ...)")
    ap.add_argument("--title", default="bench.py C2 (Np=1e8 lognormal, TSC, N=512)")
    a = ap.parse_args()
    kernels = OrderedDict()
    for p in a.raw:
        for k, d in read_raw(p).items():
            kernels[short(k)] = (k, d, os.path.basename(p))
    prof = os.path.join(ROOT, "profiles")
    summary = f"{a.tag}_ncu_full_summary.md" if a.suffix == "c2" else f"{a.tag}_ncu_full_summary_{a.suffix}.md"
    with open(os.path.join(prof, summary), "w") as f:
        f.write(f"# ncu --set full --clock-control none --import-source on, {a.title}\n\n"
                "One launch per kernel inside the `jps_timed` NVTX range.  Exported with "
                "`ncu -i <rep> --page raw --csv` and condensed by tools/make_profiles.py.  Durations under ncu are\n"
                "cold-cache and serialised; DESIGN.md quotes bench.py's CUDA-event times.  Tensor pipes are idle by "
                "design\n(nothing on this path is a dense contraction).\n")
        traffic = {}
        for s, (full, d, src) in kernels.items():
            f.write(f"\n## {full.replace('jps::', '')}\n\n(capture: {src})\n\n")
            for m in METRICS:
                if m in d:
                    v, u = d[m]
                    f.write(f"- `{m}`: {v} {u}\n")
            rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
            if rd and wr:
                traffic[s] = float(rd[0]) * UNIT_SCALE.get(rd[1], 1.0) + float(wr[0]) * UNIT_SCALE.get(wr[1], 1.0)
    with open(os.path.join(prof, f"{a.tag}_ncu_dram_traffic_{a.suffix}.json"), "w") as f:
        json.dump(traffic, f, indent=1)
    if a.launches:
        shutil.copy(a.launches, os.path.join(prof, f"{a.tag}_ncu_launches_{a.suffix}.csv"))
        tot = OrderedDict()
        nsteps = 0
        for r in csv.reader(l for l in open(a.launches) if l.startswith('"')):
            if r[0] == "ID":
                idx = {h: i for i, h in enumerate(r)}
                continue
            name = r[idx["Kernel Name"]].replace("jps::", "")
            name = re.sub(r"\(.*", "", name).replace("void ", "")
            tot[name] = tot.get(name, 0.0) + float(r[idx["Metric Value"]]) * 1e-6
            if "pk_finalize" in name:
                nsteps += 1
        nsteps = max(nsteps, 1)
        s = sum(tot.values())
        with open(os.path.join(prof, f"{a.tag}_ncu_launch_shares_{a.suffix}.txt"), "w") as f:
            f.write(f"ncu launch list (profiles/{a.tag}_ncu_launches_{a.suffix}.csv), {nsteps} steps of {a.title}; "
                    "share of the summed kernel time\n")
            for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
                f.write(f"{k[:48]:<48s} {v / nsteps:7.3f} ms/step  {100 * v / s:5.1f}%\n")
    print(json.dumps(traffic))


if __name__ == "__main__":
    main()
