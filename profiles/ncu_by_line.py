"""Attribute ncu per-SASS-instruction metrics to CUDA source lines using nvdisasm line info.

usage: python profiles/ncu_by_line.py <ncu source csv (sass)> <cubin> <kernel substring> [top]
"""
import csv
import re
import subprocess
import sys


def main():
    src_csv, cubin, kname = sys.argv[1], sys.argv[2], sys.argv[3]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
    # locate the kernel's text section
    line_of = {}
    cur_line, in_k = None, False
    for l in dis:
        if l.startswith("//--------------------- .text."):
            in_k = kname in l
            cur_line = None
            continue
        if not in_k:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            cur_line = (m.group(1).split("/")[-1], int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
        if m and cur_line:
            line_of[int(m.group(1), 16)] = cur_line
    rows = list(csv.reader(open(src_csv)))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = min(int(r[ci["Address"]], 16) for r in data)
    agg = {}
    tot_i = tot_s = 0.0
    for r in data:
        off = int(r[ci["Address"]], 16) - base
        key = line_of.get(off, ("?", 0))
        a = agg.setdefault(key, [0.0, 0.0])
        i, s = float(r[ci["Instructions Executed"]] or 0), float(r[ci["# Samples"]] or 0)
        a[0] += i
        a[1] += s
        tot_i += i
        tot_s += s
    srcs = {}
    print(f"total warp-instructions {tot_i:.0f}, samples {tot_s:.0f}")
    for (f, ln), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
        if f not in srcs:
            try:
                import glob
                cands = glob.glob(f"/root/repo/**/{f}", recursive=True)
                srcs[f] = open(cands[0]).read().splitlines() if cands else []
            except Exception:
                srcs[f] = []
        text = srcs[f][ln - 1].strip()[:100] if srcs[f] and 0 < ln <= len(srcs[f]) else ""
        print(f"{f}:{ln:<5d} inst {i / tot_i * 100:5.1f}%  samp {s / tot_s * 100:5.1f}%  {text}")


if __name__ == "__main__":
    main()
