#!/usr/bin/env python
"""Compact text summary of an .ncu-rep (one block per launch: duration, DRAM bytes, achieved GB/s, pipe / issue /
L1 / occupancy metrics) for profiles/.   python tools/ncu_summary.py report.ncu-rep [algorithmic_bytes_json] > out.txt"""
import csv
import io
import json
import subprocess
import sys

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram read"),
    ("dram__bytes_write.sum", "dram write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram throughput % of peak"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots active %"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "l1tex throughput %"),
    ("l1tex__t_sector_hit_rate.pct", "l1 hit rate %"),
    ("lts__t_sector_hit_rate.pct", "l2 hit rate %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % (elapsed)"),
    ("l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem operand-read pipe %"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed", "smem LSU wavefronts %"),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu pipe %"),
    ("sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "xu pipe %"),
]


def main():
    rep = sys.argv[1]
    alg = json.load(open(sys.argv[2])) if len(sys.argv) > 2 else {}
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    print(f"# {rep}: ncu --set full --clock-control none, one block per captured launch")
    for r in rows[2:]:
        name = r[hdr.index("Kernel Name")]
        print(f"\n{name}")
        vals = {}
        for m, label in METRICS:
            if m in hdr:
                i = hdr.index(m)
                vals[m] = (r[i], units[i])
                print(f"  {label:34s} {r[i]:>16s} {units[i]}")
        try:
            t, tu = vals["gpu__time_duration.sum"]
            t = float(t) * {"us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(tu, 1e-6)
            scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}
            rd = float(vals["dram__bytes_read.sum"][0]) * scale.get(vals["dram__bytes_read.sum"][1], 1.0)
            wr = float(vals["dram__bytes_write.sum"][0]) * scale.get(vals["dram__bytes_write.sum"][1], 1.0)
            print(f"  {'dram traffic / duration':34s} {(rd + wr) / t / 1e9:16.1f} GB/s  ({(rd + wr) / 1e6:.1f} MB)")
            for key, (b, what) in alg.items():
                if key in name:
                    print(f"  {'ALGORITHMIC bytes / duration':34s} {b / t / 1e9:16.1f} GB/s  ({b / 1e6:.1f} MB: {what})"
                          f" = {b / t / 1e9 / 6545 * 100:.1f} % of the measured 6545 GB/s")
        except Exception as e:  # noqa: BLE001
            print("  (no derived figures:", e, ")")


if __name__ == "__main__":
    main()
