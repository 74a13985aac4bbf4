#!/bin/bash
# DEV: does a THROTTLED zero-fill let a latency-bound kernel next to it run at its own speed?
for cfgs in "3 8 0" "3 8 48" "3 8 100" "3 8 200" "3 16 96" "3 16 200" "3 32 190" "3 4 100" "3 4 200" "3 2 200"; do
  set -- $cfgs
  echo "mode=$1 chunk_kb=$2 pad=$3: $(LSS_ZERO_MODE=$1 LSS_ZERO_CHUNK_KB=$2 LSS_ZERO_PAD_KB=$3 python scripts/bench_runplan_quick.py cfg2 300 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k.replace('_us',''):v for k,v in d.items() if k in ('plan_us','zero_us','gather_us','zero||plan_us','zero||gather_us','zero||lift_us','zero||bwd_us','step_us')})")"
done
