#!/bin/bash
# stability of the final code: the GPU suite three times in a row (fresh process each), smoke, wall time of the default bench
mkdir -p gpurun_out
for i in 1 2 3; do
  timeout 600 python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/stab_pytest_$i.log 2>&1; echo "pytest[$i] rc=$?"; tail -1 gpurun_out/stab_pytest_$i.log
done
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/stab_smoke.log 2>&1; echo "smoke rc=$?"
s=$(date +%s); timeout 900 python bench.py > gpurun_out/stab_bench.json 2> gpurun_out/stab_bench.err; echo "bench rc=$? wall $(( $(date +%s) - s )) s"
s=$(date +%s); timeout 600 python bench.py --impl reference > gpurun_out/stab_ref.json 2> gpurun_out/stab_ref.err; echo "ref rc=$? wall $(( $(date +%s) - s )) s"
python - <<PY
import json
d=json.loads(open("gpurun_out/stab_bench.json").read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phase_ms"], d["loss"], d["clocks"])
r=json.loads(open("gpurun_out/stab_ref.json").read().strip().splitlines()[-1]); print(r["value"], r.get("cpu_baseline"))
PY
