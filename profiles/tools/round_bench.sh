#!/bin/bash
# the bench lines kept under profiles/ for a round: profiles/tools/round_bench.sh r02  (writes gpurun_out/<tag>_bench_*.json)
tag=${1:-r02}
mkdir -p gpurun_out
for wl in cfg3 cfg2 cfg5 cfg4; do
  python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/${tag}_bench_${wl}_n1.json 2> gpurun_out/${tag}_bench_${wl}_n1.err
  tail -c 600 gpurun_out/${tag}_bench_${wl}_n1.json | head -c 10 >/dev/null
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${tag}_bench_${wl}_n1.json").read().strip().splitlines()[-1])
    print("${wl}", "ms/step %.3f" % d["ms_per_step"], "e2e %.2f ms" % d["e2e"]["ms_per_step"], "frac %.3f" % d["roofline"]["frac"], "cpu", d.get("cpu_baseline", {}).get("value"), d["clocks"])
except Exception as e:
    print("${wl}", "failed", e)
PY
done
python bench.py --impl reference --workload cfg3 --steps 2 --warmup 1 > gpurun_out/${tag}_bench_cfg3_reference.json 2> gpurun_out/${tag}_bench_cfg3_reference.err
tail -c 800 gpurun_out/${tag}_bench_cfg3_reference.json
