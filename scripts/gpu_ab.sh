#!/bin/bash
# usage (on the GPU box): scripts/gpu_ab.sh <tag> [ncu] [VAR=val ...]* -- parity tests, bench A/B over knob settings (one run per
# extra argument, "default" always first), optionally an ncu launch list of the default configuration
tag=${1:-ab}; shift
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest exit $?" | tee -a gpurun_out/pytest_$tag.log
tail -3 gpurun_out/pytest_$tag.log
run() { label=$1; shift; env "$@" timeout 300 scripts/bench_brief.sh "$label" 2>&1 | tail -1 | tee -a gpurun_out/ab_$tag.log; }
run default LSS_DUMMY=1
for kv in "$@"; do
  if [ "$kv" = ncu ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv \
      python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --e2e-steps 10 --no-e2e-graph > gpurun_out/ncu_$tag.log 2>&1
    python scripts/summarize_launches.py gpurun_out/launches_$tag.csv 2>&1 | tee gpurun_out/launch_summary_$tag.txt | head -20
  else
    run "$kv" $kv
  fi
done
