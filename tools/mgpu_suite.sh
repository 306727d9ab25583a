#!/bin/bash
# Multi-GPU evidence run (one box, N GPUs visible): parity pytest, exchange-step comparison, C3 and C5 bench lines.
# usage: tools/mgpu_suite.sh <tag> "<world sizes>"
tag=${1:-mg}; sizes=${2:-"2 4 8"}
mkdir -p gpurun_out
python -m pytest tests/test_gpu_multi.py -m gpu -q 2>&1 | tail -4 > gpurun_out/${tag}_multi_tests.log
cat gpurun_out/${tag}_multi_tests.log
port=29540
for n in $sizes; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port tools/allreduce_compare.py 2>/dev/null | grep world >> gpurun_out/${tag}_allreduce.log
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 10 --warmup 3 2>/dev/null | grep '^{' > gpurun_out/${tag}_c3_${n}gpu.json
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --config C5 --steps 2 --warmup 1 2>/dev/null | grep '^{' > gpurun_out/${tag}_c5_${n}gpu.json
done
python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${tag}_c3_1gpu.json
python bench.py --config C5 --steps 2 --warmup 1 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/${tag}_c5_1gpu.json
cat gpurun_out/${tag}_allreduce.log
python - <<PY
import json,glob
for cfg in ("c3","c5"):
    for n in (1,2,4,8):
        try:
            d=json.loads(open(f"gpurun_out/${tag}_{cfg}_{n}gpu.json").read().strip().splitlines()[-1])
            print(cfg,n,"value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"stages",{k:round(v,3) for k,v in d["stages_ms"].items()},"e2e",(round(d["e2e"]["ms_per_step"],2) if d.get("e2e") else None), d.get("parity"))
        except Exception as e:
            print(cfg,n,"missing",e)
PY
