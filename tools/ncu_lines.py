#!/usr/bin/env python3
"""Per-source-line view of an .ncu-rep captured with --import-source on: joins ncu's per-instruction SASS page
(samples, executed instructions) with the line table nvdisasm prints for the same kernel in the built library.
usage: tools/ncu_lines.py <rep> <lib.so> <mangled-kernel-substring> [top]
The library must be the binary that was profiled (instruction offsets are matched one to one)."""
import csv
import os
import re
import subprocess
import sys
import tempfile
from collections import defaultdict


def sass_rows(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h = rows[1]
    ix = {k: i for i, k in enumerate(h)}
    res = []
    for r in rows[2:]:
        if len(r) < len(h):
            continue
        res.append((int(r[ix["Address"]], 16), r[ix["Source"]].strip(), float(r[ix["# Samples"]] or 0),
                    float(r[ix["Instructions Executed"]] or 0), {k: float(r[ix[k]] or 0) for k in h if k.startswith("stall_") and "Not Issued" not in k}))
    return res


def line_table(lib, kernel):
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, capture_output=True)
    for f in sorted(os.listdir(d)):
        if not f.endswith(".cubin"):
            continue
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, f)], capture_output=True, text=True).stdout
        m = re.search(r"^\.text\.(\S*%s\S*):$" % re.escape(kernel), txt, re.M)
        if not m:
            continue
        body = txt[m.end():]
        nxt = body.find("\n//--------------------- .")
        body = body[:nxt] if nxt > 0 else body
        table, cur = {}, ("?", 0)
        for ln in body.splitlines():
            mm = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
            if mm:
                cur = (os.path.basename(mm.group(1)), int(mm.group(2)))
                continue
            mi = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*);", ln)
            if mi:
                table[int(mi.group(1), 16)] = cur
        return m.group(1), table
    raise SystemExit("kernel not found in " + lib)


def main():
    rep, lib, kernel = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    rows = sass_rows(rep)
    name, table = line_table(lib, kernel)
    base = rows[0][0]
    by = defaultdict(lambda: [0.0, 0.0, defaultdict(float)])
    tot_s = tot_i = 0.0
    for addr, src, samp, inst, stalls in rows:
        key = table.get(addr - base, ("?", 0))
        by[key][0] += samp
        by[key][1] += inst
        for k, v in stalls.items():
            by[key][2][k] += v
        tot_s += samp
        tot_i += inst
    print("kernel %s: %d SASS instructions, %.0f samples, %.0f warp instructions executed" % (name, len(rows), tot_s, tot_i))
    srcs = {}
    for (f, l), (s, i, st) in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            for root in ("cupoch_b200/csrc", "include"):
                p = os.path.join(root, f)
                if os.path.exists(p):
                    srcs[f] = open(p).read().splitlines()
        text = srcs.get(f, [""] * (l + 1))[l - 1].strip()[:90] if f in srcs and l - 1 < len(srcs[f]) else ""
        top_st = ", ".join("%s %.0f%%" % (k.replace("stall_", ""), 100 * v / max(s, 1)) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:2])
        print("%5.1f%% samples %5.1f%% instr  %s:%d  [%s]  %s" % (100 * s / tot_s, 100 * i / tot_i, f, l, top_st, text))


if __name__ == "__main__":
    main()
