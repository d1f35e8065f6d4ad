#!/usr/bin/env python
"""Summarise .ncu-rep captures (ncu --set full) into a markdown table: python tools/ncu_summary.py out.md title rep1 rep2 ..."""
import csv
import subprocess
import sys

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__registers_per_thread", "regs/thread"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed", "tensor pipe active % of elapsed"),
    ("sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor operand fetch (smem) active % of elapsed"),
    ("sm__cycles_elapsed.avg.per_second", "SM clock"),
    ("l1tex__data_pipe_lsu_wavefronts.sum", "L1 LSU wavefronts"),
    ("sm__inst_executed_pipe_tmem.avg.pct_of_peak_sustained_active", "TMEM pipe inst %"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slots busy %"),
    ("smsp__inst_executed.sum", "warp instructions"),
]


def main():
    out, title, reps = sys.argv[1], sys.argv[2], sys.argv[3:]
    lines = ["# " + title, ""]
    for rep in reps:
        raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(raw.splitlines()))
        hdr, units = rows[0], rows[1]
        idx = {h: i for i, h in enumerate(hdr)}
        lines += ["## " + rep.split("/")[-1], ""]
        def gname(r):
            g = r[idx["launch__grid_size"]] if "launch__grid_size" in idx else ""
            return r[idx["Kernel Name"]].split("(")[0].replace("void ", "").replace("lt::", "")[:30] + " #%s" % r[idx["ID"]]
        lines += ["| metric | " + " | ".join(gname(r) for r in rows[2:]) + " |",
                  "|---|" + "---:|" * len(rows[2:])]
        for key, name in KEYS:
            if key not in idx:
                continue
            i = idx[key]
            lines.append("| %s [%s] | " % (name, units[i]) + " | ".join(r[i] for r in rows[2:]) + " |")
        lines.append("")
    open(out, "w").write("\n".join(lines))


if __name__ == "__main__":
    main()
