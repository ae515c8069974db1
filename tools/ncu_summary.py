#!/usr/bin/env python
"""Turn `ncu --set full` reports into the tracked evidence under profiles/ (runs here, no GPU):

    python tools/ncu_summary.py --tag r2_final --out profiles/r2_final_ncu.md \
        --traffic profiles/ncu_traffic.json  cfg2_reddit_n128_fp32:fwd=gpurun_out/x/full_cfg2_fwd.ncu-rep ...

Each positional argument is `<workload>:<op>=<report>`.  The markdown gets one column per report
with the metrics the roofline discussion needs; the traffic file gets, per workload and op,
dram__bytes_read.sum + dram__bytes_write.sum and the L2 (lts__t_sectors x 32 B) traffic per launch —
bench.py reads it for `roofline.traffic` instead of a hand-typed constant."""
import argparse
import csv
import io
import json
import os
import subprocess

METRICS = [
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput % of peak"),
    ("lts__t_sectors.sum", "L2 sectors, all sources (x 32 B)"),
    ("lts__t_sectors_srcunit_tex_op_read.sum", "L2 read sectors from SMs (x 32 B)"),
    ("lts__t_sectors_srcunit_ltcfabric.sum", "L2 sectors via the die-to-die fabric (x 32 B)"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum", "L2 -> SM read bytes"),
    ("l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed", "L2 -> SM read, % of SM-side peak"),
    ("lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed", "L2 slice output port busy % (avg over slices)"),
    ("lts__lts2xbar_cycles_active.max.pct_of_peak_sustained_elapsed", "L2 slice output port busy % (busiest slice)"),
    ("lts__t_sector_hit_rate.pct", "L2 hit rate %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput % of peak"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit rate %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "L1/TEX throughput % of peak"),
    ("smsp__inst_executed.sum", "warp instructions"),
    ("sm__inst_executed.avg.per_cycle_active", "IPC (active)"),
    ("smsp__average_warp_latency_per_inst_issued.ratio", "warp cycles per issued instruction"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "of which long-scoreboard"),
    ("launch__registers_per_thread", "registers / thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
]
TO_BYTES = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TO_MS = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1, "msecond": 1, "s": 1e3, "second": 1e3, "nsecond": 1e-6}


def read_report(path):
    if path.endswith(".csv"):      # `ncu -i x.ncu-rep --page raw --csv` already run (on the GPU box, to keep gpurun_out small)
        out = open(path).read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    return d


def fnum(x):
    return float(x.replace(",", ""))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", required=True)
    ap.add_argument("--out", required=True)
    ap.add_argument("--traffic", default=None)
    ap.add_argument("--note", default="")
    ap.add_argument("reports", nargs="+")
    args = ap.parse_args()
    cols, traffic = [], {}
    for spec in args.reports:
        key, _, path = spec.partition("=")
        wl, _, op = key.partition(":")
        d = read_report(path)
        cols.append((wl, op, path, d))
        rd, wr = d["dram__bytes_read.sum"], d["dram__bytes_write.sum"]
        l2 = d.get("lts__t_sectors.sum", ("0", "sector"))
        l2 = (str(fnum(l2[0]) * 32), "byte")
        dur = d["gpu__time_duration.sum"]
        traffic.setdefault(wl, {})[op] = {
            "kernel": d["Kernel Name"][0],
            "dram_bytes": int(fnum(rd[0]) * TO_BYTES[rd[1]] + fnum(wr[0]) * TO_BYTES[wr[1]]),
            "l2_bytes": int(fnum(l2[0]) * TO_BYTES.get(l2[1], 1)),
            "duration_ms_under_ncu": fnum(dur[0]) * TO_MS.get(dur[1], 1),
            "l2_hit_pct": fnum(d["lts__t_sector_hit_rate.pct"][0]),
            "l2_throughput_pct_ncu": fnum(d["lts__throughput.avg.pct_of_peak_sustained_elapsed"][0]),
            "l2_slice_output_busy_pct_avg": fnum(d["lts__lts2xbar_cycles_active.avg.pct_of_peak_sustained_elapsed"][0]),
            "l2_slice_output_busy_pct_max": fnum(d["lts__lts2xbar_cycles_active.max.pct_of_peak_sustained_elapsed"][0]),
            "dram_throughput_pct_ncu": fnum(d["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][0]),
            "report": os.path.basename(path), "tag": args.tag,
        }
    with open(args.out, "w") as f:
        f.write(f"# ncu `--set full` summary — {args.tag}\n\n")
        if args.note:
            f.write(args.note + "\n\n")
        f.write("Generated by `tools/ncu_summary.py` from the `.ncu-rep` files named in the last row (scratch, "
                "`gpurun_out/`); one launch per column, `--clock-control none`, captured after the same command "
                "exited 0 without ncu.\n\n")
        f.write("| metric | " + " | ".join(f"{wl} {op}" for wl, op, _, _ in cols) + " |\n")
        f.write("|---|" + "---:|" * len(cols) + "\n")
        f.write("| kernel | " + " | ".join("`" + d["Kernel Name"][0].replace("|", "\\|")[:90] + "`" for *_, d in cols) + " |\n")
        for m, label in METRICS:
            cells = []
            for *_, d in cols:
                v, u = d.get(m, ("n/a", ""))
                try:
                    x = fnum(v)
                    v = f"{x:,.3f}".rstrip("0").rstrip(".") if abs(x) < 1e6 else f"{x:,.0f}"
                except ValueError:
                    pass
                cells.append(f"{v} {u}".strip())
            f.write(f"| {label} (`{m}`) | " + " | ".join(cells) + " |\n")
        f.write("| report | " + " | ".join(os.path.basename(p) for _, _, p, _ in cols) + " |\n")
    if args.traffic:
        old = {}
        if os.path.exists(args.traffic):
            try:
                old = json.load(open(args.traffic))
            except Exception:
                old = {}
        old = {k: v for k, v in old.items() if isinstance(v, dict)}
        for wl, opsd in traffic.items():
            old.setdefault(wl, {}).update(opsd)
        old["_comment"] = ("per workload / op: DRAM (dram__bytes_read.sum + dram__bytes_write.sum) and L2 (lts__t_sectors.sum x 32 B) "
                           "bytes of ONE launch of the named kernel, written by tools/ncu_summary.py from ncu --set full "
                           "reports; bench.py reads dram_bytes for roofline.traffic (N=1 only)")
        json.dump(old, open(args.traffic, "w"), indent=1, sort_keys=True)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
