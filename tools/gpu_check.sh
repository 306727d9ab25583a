#!/bin/bash
# Standard GPU check run under gpurun: parity tests, a short bench line, and the ncu launch list of two resident passes.
# usage: tools/gpu_check.sh <tag> [full-capture kernel regex]
tag=${1:-chk}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/${tag}_tests.log
python bench.py --steps 20 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_launches.csv python tools/profile_pass.py C3 2 > gpurun_out/${tag}_ncu1.log 2>&1
if [ -n "$2" ]; then
  ncu --set full --clock-control none --import-source on -k regex:$2 -c 4 -o gpurun_out/${tag}_full python tools/profile_pass.py C3 2 > gpurun_out/${tag}_ncu2.log 2>&1
fi
tail -4 gpurun_out/${tag}_tests.log
python tools/agg_launches.py gpurun_out/${tag}_launches.csv | head -12
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),round(d["e2e"]["ms_per_step"],2),"stages",{k:round(v,3) for k,v in d["stages_ms"].items()},"frac",round(d["roofline"]["frac"],4))
PY
