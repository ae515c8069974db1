#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from `ncu -i x.ncu-rep --page source --csv` exports
(runs here, no GPU): which instruction the warps wait at, and why.

    python tools/ncu_hot_sass.py gpurun_out/r2c10/full_cfg2_fwd.source.csv [--top 12]
"""
import argparse
import csv


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--top", type=int, default=12)
    args = ap.parse_args()
    rows = list(csv.reader(open(args.csv)))
    kernel = rows[0][1]
    hdr = rows[1]
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    c_src, c_all = hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)")
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "(Not Issued)" not in h]
    total = sum(int(r[c_all] or 0) for r in data)
    print(f"`{kernel}` — {total:,} warp-stall samples over {len(data)} SASS instructions\n")
    print("| share | SASS | dominant stall reasons |")
    print("|---:|---|---|")
    for r in sorted(data, key=lambda r: -int(r[c_all] or 0))[: args.top]:
        n = int(r[c_all] or 0)
        why = sorted(((int(r[i] or 0), hdr[i][6:]) for i in stall_cols), reverse=True)[:2]
        why_s = ", ".join(f"{w} {100 * c / max(n, 1):.0f} %" for c, w in why if c)
        print(f"| {100 * n / max(total, 1):.1f} % | `{r[c_src].strip()}` | {why_s} |")
    by = {}
    for r in data:
        for i in stall_cols:
            by[hdr[i][6:]] = by.get(hdr[i][6:], 0) + int(r[i] or 0)
    tot = sum(by.values())
    print("\nAll instructions, by reason: " + ", ".join(f"{k} {100 * v / tot:.1f} %" for k, v in sorted(by.items(), key=lambda kv: -kv[1])[:6]))


if __name__ == "__main__":
    main()
