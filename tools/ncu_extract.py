"""Compact, commit-sized extracts of Nsight Compute reports (run where `ncu` is installed; no GPU needed):

    python tools/ncu_extract.py gpurun_out/prof.ncu-rep            # one block of key metrics per profiled launch
    python tools/ncu_extract.py --launches gpurun_out/launches.csv # launch list (--metrics gpu__time_duration.sum,
                                                                   # dram__bytes_read.sum,dram__bytes_write.sum --csv)
                                                                   # -> markdown table of per-kernel shares
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_st.sum", "l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
]


def report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("## %s  grid %s block %s" % (r[col["Kernel Name"]], r[col["Grid Size"]], r[col["Block Size"]]))
        print("```")
        for k in KEYS:
            if k in col:
                print("%-90s %-10s %s" % (k, units[col[k]], r[col[k]]))
        print("```")


def launches(path):
    txt = open(path).read()
    start = txt.index('"ID"')
    rows = list(csv.DictReader(io.StringIO(txt[start:])))
    agg = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        a = agg.setdefault(name, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
        v = float(r["Metric Value"].replace(",", ""))
        m, u = r["Metric Name"], r["Metric Unit"]
        if m == "gpu__time_duration.sum":
            a["n"] += 1
            a["us"] += v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1.0)
        elif m.startswith("dram__bytes_"):
            mb = v * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
            a["rd" if "read" in m else "wr"] += mb
    tot = sum(a["us"] for a in agg.values()) or 1.0
    print("| kernel | launches | total us | us/launch | share | DRAM read MB/launch | DRAM write MB/launch |")
    print("|---|---:|---:|---:|---:|---:|---:|")
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
        n = max(a["n"], 1)
        print("| %s | %d | %.1f | %.1f | %.1f%% | %.1f | %.1f |" % (name, a["n"], a["us"], a["us"] / n, 100 * a["us"] / tot,
                                                                  a["rd"] / n, a["wr"] / n))


if __name__ == "__main__":
    if len(sys.argv) == 3 and sys.argv[1] == "--launches":
        launches(sys.argv[2])
    elif len(sys.argv) == 2:
        report(sys.argv[1])
    else:
        sys.exit(__doc__)
