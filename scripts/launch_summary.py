#!/usr/bin/env python
"""Aggregate an ncu launch list (ncu --metrics gpu__time_duration.sum --csv --log-file ...) by kernel:
launch count, total and mean duration, share of the summed GPU time; repo kernels (namespace elfb200)
and library kernels (cuDNN / cuBLAS / ATen: the network) are totalled separately.
usage: launch_summary.py <launches.csv> [--md out.md]"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.reader(lines)
    hdr = next(rd)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    for r in rd:
        if len(r) <= vi:
            continue
        ns = float(r[vi].replace(",", "")) * {"ns": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6, "nsecond": 1}.get(r[ui], 1)
        rows.append((r[ki], ns))
    agg = collections.OrderedDict()
    for k, ns in rows:
        short = re.sub(r"\(.*", "", k).replace("void ", "")
        short = re.sub(r"<.*", "<...>", short) if not short.startswith("elfb200::") else re.sub(r"\(.*", "", k).replace("void ", "")
        a = agg.setdefault(short, [0, 0.0])
        a[0] += 1
        a[1] += ns
    tot = sum(a[1] for a in agg.values())
    ours = sum(a[1] for k, a in agg.items() if k.startswith("elfb200::") or k.startswith("k_"))
    n_ours = sum(a[0] for k, a in agg.items() if k.startswith("elfb200::") or k.startswith("k_"))
    md = [f"{len(rows)} launches, {tot / 1e6:.1f} ms of summed kernel time (serialised under the profiler, cold caches).",
          f"Repo kernels (`elfb200::*`): {n_ours} launches, {ours / 1e6:.2f} ms = {100 * ours / tot:.2f} % of the GPU time; "
          f"library kernels (cuDNN / cuBLAS / ATen -- the policy/value network and its copies): {100 - 100 * ours / tot:.2f} %.",
          "", "| kernel | launches | total ms | mean us | share % |", "|---|---|---|---|---|"]
    for k, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:40]:
        md.append(f"| `{k[:90]}` | {n} | {ns / 1e6:.3f} | {ns / n / 1e3:.1f} | {100 * ns / tot:.2f} |")
    text = "\n".join(md)
    if "--md" in sys.argv:
        open(sys.argv[sys.argv.index("--md") + 1], "w").write(text + "\n")
    print(text)


if __name__ == "__main__":
    main()
