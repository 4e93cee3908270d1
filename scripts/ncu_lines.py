#!/usr/bin/env python
"""Attribute executed warp instructions of one kernel in an .ncu-rep to CUDA source lines.

Joins ncu's SASS source page (per-instruction 'Instructions Executed') with nvdisasm
--print-line-info of the same cubin, matched by instruction offset within the function.
usage: ncu_lines.py <rep> <lib.so> <mangled-kernel-substring> [top]
"""
import csv
import io
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def main():
    rep, so, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hi = next(i for i, r in enumerate(rows) if "Address" in r and "Source" in r)
    hdr = rows[hi]
    ai, si, ci = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed")
    sti = hdr.index("Warp Stall Sampling (All Samples)")
    inst = [(int(r[ai], 16), r[si].strip(), int(r[ci] or 0), int(r[sti] or 0)) for r in rows[hi + 1:] if len(r) > ci and r[ai].startswith("0x")]
    base = inst[0][0]
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(so)], cwd=tmp, capture_output=True)
    dis = ""
    for cubin in sorted(f for f in os.listdir(tmp) if f.endswith(".cubin")):
        dis += subprocess.run(["nvdisasm", "--print-line-info", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout
    sec = re.split(r"//-+ \.text\.", dis)
    body = next(s for s in sec if s.startswith("_Z") and kern in s.split(" ", 1)[0])
    line_of = {}
    cur = ("?", 0)
    for ln in body.splitlines():
        m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(\S.*);", ln)
        if m:
            line_of[int(m.group(1), 16)] = cur
    per = defaultdict(lambda: [0, 0])
    tot = sum(i[2] for i in inst)
    tot_s = sum(i[3] for i in inst)
    for addr, _, n, st in inst:
        k = line_of.get(addr - base, ("?", 0))
        per[k][0] += n
        per[k][1] += st
    srcs = {}
    print(f"total warp instructions {tot}, stall samples {tot_s}")
    for (f, l), (n, st) in sorted(per.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            for d in ("elf_b200/csrc", "include"):
                p = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), d, f)
                if os.path.exists(p):
                    srcs[f] = open(p).read().splitlines()
        text = srcs.get(f, [""] * (l + 1))[l - 1].strip() if l and f in srcs and l <= len(srcs[f]) else ""
        print(f"{100 * n / tot:5.1f}% inst {100 * st / max(tot_s, 1):5.1f}% stall  {f}:{l:<4d} {text[:100]}")


if __name__ == "__main__":
    main()
