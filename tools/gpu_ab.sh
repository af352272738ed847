#!/bin/bash
# A/B of the synthesis kernel variants (VP_SYNTH = stream | rows | rows32): parity numbers on the vocoder-only cases
# (breathy + clean voice) and stage times of the default workload. Usage: bash tools/gpu_ab.sh TAG
TAG=${1:-ab}; O=gpurun_out; mkdir -p $O
for v in stream rows rows32; do
  echo "== VP_SYNTH=$v"
  VP_SYNTH=$v timeout 300 python tools/gpu_check.py voc 2>&1 | grep -E "^\{|mismatch" | cut -c1-400
  VP_SYNTH=$v timeout 600 python bench.py --no-e2e --no-cpu --no-stream --no-parity --steps 3 --warmup 3 > $O/bench_${v}_$TAG.json 2> $O/bench_${v}_$TAG.err
  python - <<PY
import json
try:
    j = json.load(open("$O/bench_${v}_$TAG.json"))
    print("$v value", round(j["value"]), "ms/step", round(j["ms_per_step"], 1), {k: round(x / j["steps"], 1) for k, x in j["roofline"]["stage_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("$O/bench_${v}_$TAG.err").read()[-1500:])
PY
done
