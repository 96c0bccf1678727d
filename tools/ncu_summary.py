#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) into profiles/<name>.csv: one row per captured launch with the metrics the
roofline discussion uses.  usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/r01_post_physics"""
import csv
import io
import json
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__bytes.sum.per_second",
        "dram__cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__m_xbar2l1tex_read_sectors_mem_global_op_tma_ld.sum", "l1tex__m_xbar2l1tex_read_sectors_mem_lg_op_ld.sum",
        "l1tex__m_l1tex2xbar_write_sectors_mem_global_op_tma_st.sum", "l1tex__m_l1tex2xbar_write_sectors_mem_lg_op_st.sum",
        "lts__t_sector_hit_rate.pct", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio"]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}


def main(rep, out):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel"] + KEYS)
        w.writerow(["(unit)"] + [units[hdr.index(k)] if k in hdr else "" for k in KEYS])
        for r in rows[2:]:
            w.writerow([r[hdr.index("Kernel Name")]] + [r[hdr.index(k)] if k in hdr else "" for k in KEYS])
    traffic = []
    for r in rows[2:]:
        t = 0.0
        for k in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            t += float(r[hdr.index(k)].replace(",", "")) * UNIT[units[hdr.index(k)]]
        traffic.append(t)
    print(json.dumps({"launches": len(traffic), "dram_bytes_per_launch": sum(traffic) / len(traffic)}))


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
