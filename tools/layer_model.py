#!/usr/bin/env python
"""Per-layer bounds of the tcgen05 conv path vs the measured per-launch times of a bench timeline.

    python tools/layer_model.py profiles/r01e_timeline_tc.json [out.md]

For every conv shape of the step (label `N32 1x24x24 Cin256 Cout1024 k111 s1` = batch, output grid, channels, kernel,
stride) three lower bounds are computed from the measured machine constants:
  mma     3-product math at the measured floor: 384 cycles per (128 x 128 x 32) chunk, scaled by the N tile
  ingest  operand bytes through the per-SM L2 ingest port, 64 B/clk: A 16 KB + B Nt*128 B per chunk and CTA
  hbm     compulsory DRAM bytes (input + output, 4 B per element; + the residual read of the bottleneck expand layers) at 6.57 TB/s
and compared with the measured time; `x bound` is measured / max(bounds).  148 SMs at 1.965 GHz, perfect balance assumed.
"""
import collections
import json
import math
import re
import sys

SMS, GHZ, HBM = 148, 1.965, 6571.9e9


def parse(desc):
    m = re.match(r"N(\d+) (\d+)x(\d+)x(\d+) Cin(\d+) Cout(\d+) k(\d)(\d)(\d) s(\d)", desc)
    n, d, h, w, cin, cout, kd, kh, kw, s = map(int, m.groups())
    return n, d, h, w, cin, cout, kd * kh * kw, s


def bounds(desc, kernel):
    n, d, h, w, cin, cout, taps, s = parse(desc)
    pos = n * d * h * w
    cinp, coutp = -(-cin // 32) * 32, -(-cout // 16) * 16
    nt = coutp if coutp <= 128 else 128
    m_tiles, n_tiles, chunks = -(-pos // 128), -(-coutp // nt), taps * cinp // 32
    if kernel == "conv_fold":       # kw folded into N: per (kd, kh) step two slices of N = 2*NF + NF columns, x windows keep WX-K+1 of WX
        k = round(taps ** (1 / 3))
        nc = -(-cout // 16) * 16
        nf = k * nc
        wx = 32 if (k == 7 and w >= 32) else 16
        tiles = n * d * -(-h // (128 // wx)) * -(-w // (wx - k + 1))
        mma = tiles * k * k * 2 * (2 * nf + nf) / 2 / SMS / (GHZ * 1e3)
        ingest = tiles * k * (wx * (128 // wx + k - 1) * 128) / 64 / SMS / (GHZ * 1e3)
    else:
        mma = m_tiles * n_tiles * chunks * 384 * (nt / 128) / SMS / (GHZ * 1e3)                 # us
        ingest = m_tiles * n_tiles * chunks * (16384 + nt * 128) / 64 / SMS / (GHZ * 1e3)
    in_pos = pos * (s ** (2 if d == 1 else 3))
    hbm_b = 4.0 * (in_pos * cinp + pos * max(coutp, 32))
    if taps == 1 and s == 1 and d == 1 and cout == 4 * cin:
        hbm_b += 4.0 * pos * coutp          # bottleneck expand layers add the identity (one more read of the output size)
    return mma, ingest, hbm_b / HBM * 1e6


def main():
    tl = json.load(open(sys.argv[1]))
    agg = collections.OrderedDict()
    for r in tl:
        if r["kernel"] not in ("conv_tc", "conv_fold"):
            continue
        a = agg.setdefault((r["kernel"], r["desc"]), [0, 0.0])
        a[0] += 1
        a[1] += r["ms"] * 1e3
    rows = []
    for (kernel, desc), (cnt, us) in agg.items():
        mma, ing, hbm = bounds(desc, kernel)
        per = us / cnt
        lim = max(mma, ing, hbm)
        which = "mma" if lim == mma else ("ingest" if lim == ing else "hbm")
        rows.append((us - lim * cnt, kernel, desc, cnt, per, mma, ing, hbm, which, per / lim))
    rows.sort(reverse=True)
    tot = sum(r[4] * r[3] for r in rows)
    gap = sum(r[0] for r in rows)
    lines = ["# Per-layer bounds vs measured (%s)" % sys.argv[1], "",
             "Total conv time %.2f ms; sum of the per-layer lower bounds %.2f ms; gap %.2f ms.  Columns in us per launch." % (tot / 1e3, (tot - gap) / 1e3, gap / 1e3), "",
             "| kernel | layer | launches | measured | mma | ingest | hbm | bound | x bound | gap (all launches, us) |", "|---|---|---:|---:|---:|---:|---:|---|---:|---:|"]
    for g, kernel, desc, cnt, per, mma, ing, hbm, which, ratio in rows:
        lines.append("| %s | %s | %d | %.1f | %.1f | %.1f | %.1f | %s | %.2f | %.0f |" % (kernel, desc, cnt, per, mma, ing, hbm, which, ratio, g))
    text = "\n".join(lines) + "\n"
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(text)
    print(text)


if __name__ == "__main__":
    main()
