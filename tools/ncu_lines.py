#!/usr/bin/env python
"""Per-source-line summary of an `ncu --page source --csv --print-source cuda,sass` dump.

usage: ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > X.csv; python tools/ncu_lines.py X.csv [top_n]
Prints the share of stall samples / executed instructions per file and the hottest source lines with their three
dominant stall reasons (lines carrying a source line number, i.e. the per-line roll-up rows of the dump)."""
import csv
import sys


def main(path, top=40):
    rows = list(csv.reader(open(path)))
    sec, hdr, data = None, None, []
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            sec = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or r[2] != "-":          # keep the per-line roll-up rows (Address == "-")
            continue
        d = {}
        for k, v in zip(hdr, r):
            d.setdefault(k, v)
        try:
            smp, inst = float(d["# Samples"] or 0), float(d["Instructions Executed"] or 0)
        except ValueError:
            continue
        data.append((sec, int(d["Line No"]), r[1].strip(), smp, inst, d))
    ts, ti = sum(x[3] for x in data), sum(x[4] for x in data)
    print(f"total samples {ts:.0f}, warp instructions {ti:.0f}")
    for f in sorted(set(x[0] for x in data)):
        s, i = sum(x[3] for x in data if x[0] == f), sum(x[4] for x in data if x[0] == f)
        print(f"  {f:20s} {100 * s / ts:5.1f}% samples {100 * i / ti:5.1f}% instructions")
    tot = {}
    for x in data:
        for k, v in x[5].items():
            if k.startswith("stall_") and "Not" not in k:
                tot[k] = tot.get(k, 0) + float(v or 0)
    print("stalls:", " ".join(f"{k[6:]}={100 * v / ts:.1f}%" for k, v in sorted(tot.items(), key=lambda kv: -kv[1])[:10]))
    for x in sorted(data, key=lambda x: -x[3])[:top]:
        st = {k: float(v or 0) for k, v in x[5].items() if k.startswith("stall_") and "Not" not in k}
        top3 = sorted(st.items(), key=lambda kv: -kv[1])[:3]
        print(f"{x[0]}:{x[1]:4d} {100 * x[3] / ts:5.2f}%s {100 * x[4] / ti:5.2f}%i  {x[2][:78]:78s} | "
              + " ".join(f"{k[6:]}={v:.0f}" for k, v in top3))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
