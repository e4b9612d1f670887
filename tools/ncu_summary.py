"""Condenses an .ncu-rep (read with `ncu -i ... --page raw --csv`) into the handful of metrics DESIGN.md and
bench.py's roofline refer to.  Usage: python tools/ncu_summary.py <report.ncu-rep> [kernel-regex] > profiles/x.txt"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_sectors.sum", "lts__t_sector_hit_rate.pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__t_sector_hit_rate.pct", "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum",
    "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
]


def main():
    rep = sys.argv[1]
    pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units = rows[0], rows[1]
    unit = dict(zip(hdr, units))
    print("# %s" % rep)
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        name = d.get("Kernel Name", "")
        if pat and not pat.search(name):
            continue
        print("\n## %s" % name[:110])
        for k in KEYS:
            if k in d:
                print("%-80s %14s %s" % (k, d[k], unit.get(k, "")))
        if "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum" in d and "l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum" in d:
            try:
                s = float(d["l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
                q = float(d["l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum"].replace(",", ""))
                print("%-80s %14.2f" % ("sectors per global load request", s / q))
            except Exception:
                pass


if __name__ == "__main__":
    main()
