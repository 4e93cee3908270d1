#!/usr/bin/env python
"""Summarise an .ncu-rep (ncu --set full) per kernel launch: duration, DRAM bytes, throughput %,
executed warp instructions, issue activity, achieved occupancy, registers, shared memory.
usage: ncu_summary.py <rep> [--md out.md] [--traffic profiles/traffic.json]"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "ns",
    "dram__bytes_read.sum": "dram_rd",
    "dram__bytes_write.sum": "dram_wr",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "smsp__inst_executed.sum": "warp_inst",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active": "alu_pct",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "launch__shared_mem_per_block_static": "smem_static",
    "launch__grid_size": "grid",
    "launch__block_size": "block",
    "launch__waves_per_multiprocessor": "waves",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "threads_per_inst",
}


def to_bytes(v, unit):
    mul = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return float(v) * mul


def to_ns(v, unit):
    mul = {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "s": 1e9, "second": 1e9, "nsecond": 1}.get(unit, 1)
    return float(v) * mul


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        d = {"kernel": r[col["Kernel Name"]], "id": r[col["ID"]]}
        for m, k in WANT.items():
            if m in col and r[col[m]] not in ("", "n/a"):
                v = r[col[m]].replace(",", "")
                u = units[col[m]]
                try:
                    d[k] = to_ns(v, u) if k == "ns" else to_bytes(v, u) if k.startswith("dram_") and k != "dram_pct" else float(v)
                except ValueError:
                    pass
        launches.append(d)
    md = ["| # | kernel | grid x block | time (us) | DRAM rd+wr (MB) | DRAM GB/s | DRAM % | warp inst (M) | issue % | ALU % | warps active % | thr/inst | regs | smem (B) | L2 hit % |",
          "|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
    for d in launches:
        byt = d.get("dram_rd", 0) + d.get("dram_wr", 0)
        us = d.get("ns", 0) / 1e3
        md.append("| {id} | `{k}` | {g:.0f} x {b:.0f} | {us:.1f} | {mb:.2f} | {gbs:.0f} | {dp:.1f} | {wi:.2f} | {ip:.1f} | {ap:.1f} | {wa:.1f} | {ti:.1f} | {rg:.0f} | {sm:.0f} | {l2:.1f} |".format(
            id=d["id"], k=d["kernel"][:60], g=d.get("grid", 0), b=d.get("block", 0), us=us, mb=byt / 1e6,
            gbs=byt / max(d.get("ns", 1), 1), dp=d.get("dram_pct", 0), wi=d.get("warp_inst", 0) / 1e6, ip=d.get("issue_pct", 0),
            ap=d.get("alu_pct", 0), wa=d.get("warps_active_pct", 0), ti=d.get("threads_per_inst", 0), rg=d.get("regs", 0),
            sm=d.get("smem_dyn", 0) + d.get("smem_static", 0), l2=d.get("l2_hit_pct", 0)))
    text = "\n".join(md)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(text + "\n")
    else:
        print(text)
    if "--json" in sys.argv:
        json.dump(launches, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)


if __name__ == "__main__":
    main()
