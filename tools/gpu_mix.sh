#!/bin/bash
# Executed instruction mix + DRAM bytes of every kernel of one pass, for the four workloads (small streams x seconds so
# that ncu's per-kernel replay stays short). Usage: bash tools/gpu_mix.sh TAG   -> gpurun_out/mix_<workload>_TAG.csv
TAG=${1:-r02}; O=gpurun_out; mkdir -p $O
M=$(python -c "import sys; sys.path.insert(0,'tools'); import ncu_mix; print(ncu_mix.METRICS)")
for wl in chain48 chain44 voc44 pitch44; do
  timeout 900 ncu --metrics $M --clock-control none -c 300 --csv --log-file $O/mix_${wl}_$TAG.csv \
      python bench.py --steps 1 --warmup 1 --workload $wl --streams 256 --seconds 10 --no-e2e --no-cpu --no-stream --no-parity --no-sub > $O/mix_${wl}_$TAG.log 2>&1
  echo "mix $wl rc=$?"
done
