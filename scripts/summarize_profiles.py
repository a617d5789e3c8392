"""Turns the ncu artefacts that came back in gpurun_out/ into the small, tracked summaries under profiles/.

    python scripts/summarize_profiles.py launches gpurun_out/launches_X.csv profiles/r01_launches_X.csv
    python scripts/summarize_profiles.py report   gpurun_out/prof_X.ncu-rep  profiles/r01_prof_X.json

`launches`: per-kernel totals of an `ncu --metrics gpu__time_duration.sum` launch list (cold-cache, serialised:
compare shares).  `report`: the metrics the roofline / issue-slot arguments in DESIGN.md cite, per captured launch
of an `ncu --set full` report (read with `ncu -i ... --page raw --csv`, no GPU needed)."""
import collections
import csv
import json
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.avg.pct_of_peak_sustained_active",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
]


def launches(src, dst):
    t = collections.defaultdict(float)
    n = collections.Counter()
    rows = [l for l in open(src) if not l.startswith("==")]
    for r in csv.DictReader(rows):
        if r["Metric Name"] == "gpu__time_duration.sum":
            k = r["Kernel Name"].split("(")[0]
            t[k] += float(r["Metric Value"]) / 1e6
            n[k] += 1
    tot = sum(t.values())
    with open(dst, "w") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "launches", "total_ms", "share_pct"])
        for k in sorted(t, key=t.get, reverse=True):
            w.writerow([k, n[k], "%.4f" % t[k], "%.2f" % (100 * t[k] / tot)])
        w.writerow(["TOTAL", sum(n.values()), "%.4f" % tot, "100.00"])


def report(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = csv.reader(out.splitlines())
    hdr = next(rd)
    units = next(rd)
    res = []
    for r in rd:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        e = {"kernel": d["Kernel Name"], "grid": d.get("Grid Size"), "block": d.get("Block Size")}
        for k in KEEP:
            if k in d and d[k] not in ("", "n/a"):
                e[k] = {"value": float(d[k].replace(",", "")), "unit": u[k]}
        res.append(e)
    json.dump({"source": src, "launches": res}, open(dst, "w"), indent=1)


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
