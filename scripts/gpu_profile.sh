#!/bin/bash
# usage (on the GPU box): scripts/gpu_profile.sh <tag> -- the round's evidence: default bench line, reference arm, ncu launch list,
# one `ncu --set full` capture of the kernels of a step.  Every ncu pass runs after the same command exited 0 without ncu.
tag=${1:-prof}
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench exit $?"; tail -c 400 gpurun_out/bench_$tag.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err; echo "ref exit $?"
CMD="python bench.py --steps 20 --warmup 3 --no-graph --no-cpu-baseline --no-gpu-reference --no-train --no-e2e"
$CMD > gpurun_out/plain_$tag.log 2>&1 || exit 1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $CMD > gpurun_out/ncu_$tag.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_$tag.csv > gpurun_out/launch_summary_$tag.txt 2>&1; head -14 gpurun_out/launch_summary_$tag.txt
# full capture: skip the warm-up steps' launches, then two steps' worth of our kernels
timeout 900 ncu --set full --import-source on --clock-control none -k regex:'^k_' --launch-skip 24 --launch-count 8 \
  -o gpurun_out/prof_$tag -f $CMD > gpurun_out/ncu_full_$tag.log 2>&1; echo "ncu full exit $?"
ls -la gpurun_out/prof_$tag.ncu-rep
