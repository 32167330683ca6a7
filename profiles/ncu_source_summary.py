"""Summarise `ncu --page source --csv` output per source line: instructions executed and stall samples.

usage: ncu -i X.ncu-rep --page source --csv > src.csv ; python profiles/ncu_source_summary.py src.csv [top_n]
The SASS view has no line column in CSV mode, so this uses `--print-source sass,cuda`-less CSV: rows are SASS
instructions; we aggregate by the CUDA source line when ncu provides it, else list the top SASS instructions.
"""
import csv
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    rows = list(csv.reader(open(path)))
    # first line: kernel name; second: header
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    tot_inst = sum(float(r[ci["Instructions Executed"]] or 0) for r in data)
    tot_samp = sum(float(r[ci["# Samples"]] or 0) for r in data)
    print(f"total warp instructions {tot_inst:.0f}, stall samples {tot_samp:.0f}")
    data.sort(key=lambda r: -float(r[ci["# Samples"]] or 0))
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    for r in data[:top]:
        st = sorted(((float(r[ci[c]] or 0), c) for c in stall_cols), reverse=True)[:2]
        print(f"{r[ci['Address']][-6:]:>6s} samp {float(r[ci['# Samples']] or 0) / tot_samp * 100:5.1f}% inst {float(r[ci['Instructions Executed']] or 0) / tot_inst * 100:5.1f}% "
              f"thr {r[ci['Avg. Threads Executed']]:>5s}  {r[ci['Source']][:70]:70s} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f}")


if __name__ == "__main__":
    main()
