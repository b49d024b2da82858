#!/usr/bin/env python
"""Summarise an .ncu-rep (read here with `ncu -i`, no GPU needed) into the few numbers the roofline argument uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [kernel-substring] > profiles/rNN_<kernel>.md"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of ncu peak)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__registers_per_thread", "registers/thread"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem/block"),
    ("launch__shared_mem_per_block_static", "static smem/block"),
    ("launch__occupancy_limit_shared_mem", "occupancy limit (smem) blocks/SM"),
    ("launch__occupancy_limit_registers", "occupancy limit (regs) blocks/SM"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe %"),
]


def main():
    rep = sys.argv[1]
    pat = sys.argv[2] if len(sys.argv) > 2 else ""
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    print("# ncu summary of `%s`\n" % rep)
    for r in rows[2:]:
        if pat and pat not in r[ki]:
            continue
        print("## %s\n" % r[ki].split("(")[0])
        print("| metric | value | unit |\n|---|---|---|")
        for key, label in KEYS:
            if key in hdr:
                i = hdr.index(key)
                print("| %s (`%s`) | %s | %s |" % (label, key, r[i], units[i]))
        stalls = [(float(r[i]), h) for i, h in enumerate(hdr)
                  if "issue_stalled" in h and h.endswith("_per_issue_active.ratio") and r[i] not in ("", "n/a")]
        stalls.sort(reverse=True)
        print("\nTop warp stall reasons (warps stalled per issue-active cycle): " +
              ", ".join("%s %.2f" % (h.split("issue_stalled_")[1].split("_per_")[0], v) for v, h in stalls[:6]))
        try:
            rd = float(r[hdr.index("dram__bytes_read.sum")]); wr = float(r[hdr.index("dram__bytes_write.sum")])
            u = units[hdr.index("dram__bytes_read.sum")]
            dur = float(r[hdr.index("gpu__time_duration.sum")]); du = units[hdr.index("gpu__time_duration.sum")]
            print("\nDRAM traffic per launch: %.3f %s read + %.3f %s written = %.3f %s in %.3f %s" % (rd, u, wr, u, rd + wr, u, dur, du))
        except Exception:
            pass
        print()


if __name__ == "__main__":
    main()
