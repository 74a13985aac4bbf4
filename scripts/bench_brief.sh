#!/bin/bash
# usage: scripts/bench_brief.sh <label> [bench.py args...]  -- prints value + in-step stage times
label=$1; shift
python bench.py --steps 1000 --no-cpu-baseline "$@" 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$label', d['value'], 'e2e', d['e2e']['value'], d['roofline']['stage_us_in_step'])"
