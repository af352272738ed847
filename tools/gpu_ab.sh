#!/bin/bash
# A/B on one box: GPU tests of the build that travelled, then for every variant ("name:nvcc defines:env") a rebuild on the
# box and a short stage-timed bench of the default workload. Usage: bash tools/gpu_ab.sh TAG "base::" "u5:-DAV_STAGE_UNROLL=5:" ...
TAG=${1:-ab}; shift; O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_gpu_$TAG.log
for spec in "$@"; do
  name=${spec%%:*}; rest=${spec#*:}; defs=${rest%%:*}; envs=${rest#*:}
  if [ -n "$defs" ] || [ "$name" != "base" ]; then (cd vocoderproject_b200 && VP_NVCC_EXTRA="$defs" python build.py --force > /dev/null 2>&1; echo "build $name rc=$?"); fi
  env $envs timeout 600 python bench.py --no-e2e --no-cpu --no-stream --no-parity --no-sub --steps 3 --warmup 3 > $O/bench_${name}_$TAG.json 2> $O/bench_${name}_$TAG.err
  python - <<PY
import json
try:
    j = json.load(open("$O/bench_${name}_$TAG.json"))
    print("$name value", round(j["value"]), "ms/step", round(j["ms_per_step"], 1), {k: round(x / j["steps"], 1) for k, x in j["roofline"]["stage_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("$O/bench_${name}_$TAG.err").read()[-1500:])
PY
done
