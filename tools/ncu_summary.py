"""Condensed view of an .ncu-rep (run where ncu is installed): duration, DRAM traffic, throughputs,
occupancy limiters and the warp-stall breakdown per captured launch. Usage: python tools/ncu_summary.py file.ncu-rep"""
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "launch__shared_mem_per_block_static",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "smsp__inst_executed.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "smsp__sass_average_data_bytes_per_sector_mem_global_op_ld.pct", "smsp__sass_average_data_bytes_per_sector_mem_global_op_st.pct",
]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print("== %s" % d.get("Kernel Name", "")[:110])
        for k in KEYS:
            if d.get(k, "") not in ("", "n/a"):
                print("  %-78s %16s %s" % (k, d[k][:16], u.get(k, "")))
        stalls = []
        for k in hdr:
            m = re.match(r"smsp__average_warps_issue_stalled_(.*)_per_issue_active.ratio", k)
            if m and d[k] not in ("", "n/a"):
                stalls.append((float(d[k]), m.group(1)))
        stalls.sort(reverse=True)
        print("  stalls per issue: " + ", ".join("%s %.2f" % (n, v) for v, n in stalls[:7]))


if __name__ == "__main__":
    main()
