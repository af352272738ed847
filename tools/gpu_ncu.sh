#!/bin/bash
# ncu evidence for one round: launch list (all kernels of a short bench run) + one --set full capture of the hot kernels.
# Usage: bash tools/gpu_ncu.sh TAG [kernel regex]
TAG=${1:-r01}; RX=${2:-"k_voc_synth|k_pitch_frame|k_yin|k_voc_autocorr|k_voc_levinson|k_marks|k_pitch_iir"}
O=gpurun_out; mkdir -p $O
SMALL="--workload chain48 --streams 256 --seconds 10 --no-e2e --no-cpu --no-stream"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 1 $SMALL > $O/ncu_launch_$TAG.log 2>&1; echo "launch list rc=$?"
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:$RX" -c 16 -f -o $O/prof_$TAG \
    python bench.py --steps 1 --warmup 1 $SMALL > $O/ncu_full_$TAG.log 2>&1; echo "ncu full rc=$?"
ls -la $O | tail -8
