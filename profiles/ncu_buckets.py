"""Bucket ncu per-instruction counts by source line ranges: python profiles/ncu_buckets.py src.csv cubin kernel file.cu 'lo-hi:name,...'"""
import csv, re, subprocess, sys
src_csv, cubin, kname, fname, spec = sys.argv[1:6]
buckets = []
for part in spec.split(","):
    rng, name = part.split(":")
    lo, hi = rng.split("-")
    buckets.append((int(lo), int(hi), name))
dis = subprocess.run(["nvdisasm", "--print-line-info", cubin], capture_output=True, text=True).stdout.splitlines()
line_of, cur, ink = {}, None, False
for l in dis:
    if l.startswith("//--------------------- .text."):
        ink = kname in l; cur = None; continue
    if not ink: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(\S.*?);", l)
    if m and cur: line_of[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv))); hdr = rows[1]; ci = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
base = min(int(r[ci["Address"]], 16) for r in data)
bs, tot, tots = {}, 0, 0
for r in data:
    off = int(r[ci["Address"]], 16) - base
    f, ln = line_of.get(off, ("?", 0))
    i = float(r[ci["Instructions Executed"]] or 0); s = float(r[ci["# Samples"]] or 0)
    name = "other:" + f
    if f == fname:
        for lo, hi, nm in buckets:
            if lo <= ln <= hi: name = nm; break
    b = bs.setdefault(name, [0, 0]); b[0] += i; b[1] += s; tot += i; tots += s
print(f"total warp-instructions {tot:.0f}")
for k, (i, s) in sorted(bs.items(), key=lambda kv: -kv[1][0]):
    print(f"{k:40s} inst {i / tot * 100:5.1f}%  samples {s / tots * 100:5.1f}%")
