#!/bin/bash
# Quick GPU iteration: GPU tests, then a short stage-timed bench, then (optional) ncu --set full of chosen kernels.
# Usage: bash tools/gpu_quick.sh TAG [kernel regex for ncu | none] [bench args...]
TAG=${1:-q}; RX=${2:-none}; shift; shift
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -4 $O/pytest_gpu_$TAG.log
timeout 900 python bench.py --no-e2e --no-cpu --no-stream "$@" > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?"
python - <<PY
import json
try:
    j = json.load(open("$O/bench_$TAG.json"))
    print("value", round(j["value"]), "ms/step", round(j["ms_per_step"], 1), {k: round(v / j["steps"], 1) for k, v in j["roofline"]["stage_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("$O/bench_$TAG.err").read()[-2000:])
PY
if [ "$RX" != "none" ]; then
  timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$RX" -c 6 -f -o $O/prof_$TAG \
      python bench.py --steps 1 --warmup 1 --workload chain48 --streams 256 --seconds 10 --no-e2e --no-cpu > $O/ncu_full_$TAG.log 2>&1; echo "ncu rc=$?"
fi
