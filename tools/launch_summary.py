#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list of
tools/profile_step.py --repeat 2 into a per-kernel markdown table (second forward only: the first one also packs weights).

    python tools/launch_summary.py launches.csv out.md "title" [traffic.json]
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    src, out, title = sys.argv[1], sys.argv[2], sys.argv[3]
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"')) if len(r) >= 15]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = OrderedDict()
    for r in rows[1:]:
        d = launches.setdefault(int(r[ix["ID"]]), {"name": r[ix["Kernel Name"]]})
        d[r[ix["Metric Name"]]] = float(r[ix["Metric Value"]].replace(",", ""))
        d["unit:" + r[ix["Metric Name"]]] = r[ix["Metric Unit"]]
    ids = sorted(launches)
    lt = [i for i in ids if "lt::" in launches[i]["name"]]
    # second forward = second half of the lt:: launches (the first forward additionally runs the weight packing kernels)
    coord = [i for i in lt if "coord_volume_kernel" in launches[i]["name"]]
    start = coord[-1] if coord else lt[len(lt) // 2]
    sel = [i for i in ids if i >= start]
    agg = OrderedDict()
    foreign = 0
    for i in sel:
        d = launches[i]
        name = re.sub(r"\(.*", "", d["name"]).replace("void ", "")
        if "lt::" not in name:
            foreign += 1
            continue
        t_ns = d.get("gpu__time_duration.sum", 0.0)
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d.get("unit:gpu__time_duration.sum", "ns"), 1e-3)
        a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
        a[0] += 1
        a[1] += t_ns * scale
        for k, j in (("dram__bytes_read.sum", 2), ("dram__bytes_write.sum", 3)):
            u = d.get("unit:" + k, "byte")
            a[j] += d.get(k, 0.0) * {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6)
    tot = sum(a[1] for a in agg.values())
    lines = ["# " + title, "", "| kernel | launches | total us | share | DRAM read MB | DRAM write MB |", "|---|---:|---:|---:|---:|---:|"]
    for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("| %s | %d | %.1f | %.1f%% | %.1f | %.1f |" % (name, a[0], a[1], 100 * a[1] / tot, a[2], a[3]))
    lines += ["", "Total %.1f us over %d launches of our kernels; %d launches of other kernels (torch fills / copies) in the window."
              % (tot, sum(a[0] for a in agg.values()), foreign)]
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))
    if len(sys.argv) > 4:
        # per-launch DRAM traffic (read + write bytes) per kernel family, for bench.py's `roofline*.traffic`
        import json
        fam = {}
        for name, a in agg.items():
            short = name.split("::")[-1].split("<")[0]
            key = ("unproject" if "unproject" in short else "softargmax" if ("softargmax" in short or short.startswith("stream_")) else short)
            f = fam.setdefault(key, [0, 0.0])
            f[0] += a[0] if key not in ("unproject", "softargmax") else 0
            f[1] += (a[2] + a[3]) * 1e6
        traffic = {"_source": "%s (ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum, one steady-state forward, config #2, B=8): %s"
                   % (src.split("/")[-1], title),
                   "unit": "bytes per launch (read + write); unproject / softargmax: all kernels of the stage together"}
        for key, (cnt, nbytes) in fam.items():
            traffic[key] = nbytes / max(cnt, 1)
        json.dump(traffic, open(sys.argv[4], "w"), indent=1)


if __name__ == "__main__":
    main()
