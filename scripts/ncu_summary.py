#!/usr/bin/env python
"""Key metrics of `ncu --set full` reports (run here, no GPU needed):  python scripts/ncu_summary.py a.ncu-rep [b.ncu-rep ...]"""
import csv
import io
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "lts__t_sector_hit_rate.pct",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    out = {}
    for path in sys.argv[1:]:
        raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(io.StringIO(raw)))
        hdr, units = rows[0], rows[1]
        for row in rows[2:]:
            d = dict(zip(hdr, row))
            u = dict(zip(hdr, units))
            name = d["Kernel Name"].split("(")[0]
            rec = {k: (d[k] + " " + u[k]).strip() for k in WANT if k in d}
            stalls = {k.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(v.replace(",", "")) for k, v in d.items()
                      if k.startswith("smsp__pcsamp_warps_issue_stalled_") and "not_issued" not in k and v}
            tot = sum(stalls.values()) or 1.0
            rec["top_stalls_pct"] = {k: round(100 * v / tot, 1) for k, v in sorted(stalls.items(), key=lambda kv: -kv[1])[:5]}
            try:
                rd, wr = float(d["dram__bytes_read.sum"].replace(",", "")), float(d["dram__bytes_write.sum"].replace(",", ""))
                scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_read.sum"]]
                scale_w = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1}[u["dram__bytes_write.sum"]]
                rec["dram_traffic_bytes"] = int(rd * scale + wr * scale_w)
            except (KeyError, ValueError):
                pass
            out[name] = rec
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
