#!/bin/bash
# one `ncu --set full` capture of one kernel of one workload, summarised into gpurun_out/ (copy the summaries you keep into profiles/):
#   profiles/tools/ncu_capture.sh <tag> <workload> <kernel regex> [skip]
# writes gpurun_out/<tag>.ncu-rep (kept small: one launch), <tag>_details.txt (section pages) and <tag>_raw.csv (all metrics)
tag=$1; wl=$2; rx=$3; skip=${4:-1}
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:$rx -s $skip -c 1 -f -o gpurun_out/$tag \
    python bench.py --workload $wl --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/${tag}.log 2>&1
ncu -i gpurun_out/$tag.ncu-rep --page details > gpurun_out/${tag}_details.txt 2>/dev/null
ncu -i gpurun_out/$tag.ncu-rep --page raw --csv > gpurun_out/${tag}_raw.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/${tag}_raw.csv")))
if len(rows) >= 3:
    hdr, vals = rows[0], rows[-1]
    want = ["Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active"]
    units = rows[1]
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(w, "=", vals[i], units[i])
PY
