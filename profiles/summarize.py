"""Turn gpurun_out/ ncu artefacts into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_r01.csv profiles/r01_launches.txt
    python profiles/summarize.py raw gpurun_out/prof_r01.ncu-rep profiles/r01_kernels.txt
"""
import collections
import csv
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.sum", "smsp__inst_executed.sum",
    "sm__inst_executed_pipe_tensor_subpipe_dmma.avg.pct_of_peak_sustained_active",
    "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared_op_ld.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__maximum_warps_per_active_cycle_pct",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
]


def launches(src, dst):
    with open(src) as f:
        lines = [ln for ln in f if not ln.startswith("==")]
    tot = collections.OrderedDict()
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r["Metric Unit"], 1.0)
        t = tot.setdefault(r["Kernel Name"][:90], [0, 0.0])
        t[0] += 1
        t[1] += v
    s = sum(v[1] for v in tot.values())
    with open(dst, "w") as out:
        out.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src})\n")
        out.write("# per-launch times are cold-cache and serialised: compare SHARES\n")
        out.write(f"# total {s / 1e3:.3f} ms over {sum(v[0] for v in tot.values())} launches\n")
        out.write(f"{'us':>12} {'share':>7} {'n':>5} {'us/launch':>11}  kernel\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
            out.write(f"{v[1]:12.1f} {100 * v[1] / s:6.1f}% {v[0]:5d} {v[1] / v[0]:11.1f}  {k}\n")


def raw(src, dst):
    txt = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True,
                         text=True, check=True).stdout
    rd = list(csv.reader(txt.splitlines()))
    hdr, units = rd[0], rd[1]
    with open(dst, "w") as out:
        out.write(f"# ncu --set full --clock-control none --import-source on  ({src})\n")
        seen = collections.Counter()
        for r in rd[2:]:
            name = r[hdr.index("Kernel Name")]
            seen[name[:40]] += 1
            if seen[name[:40]] > 2:
                continue
            out.write("---\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    out.write(f"{k} [{units[i]}] = {r[i][:100]}\n")


if __name__ == "__main__":
    {"launches": launches, "raw": raw}[sys.argv[1]](sys.argv[2], sys.argv[3])
