cd $GRAFT_REPO_ROOT
KEYS="('forward_us','fwd_ordered_us','bwd_us','step_us','step_kept_plan_us')"
run_prev() { python - "$@" <<'PY'
import sys, runpy
from lss_carla_b200 import _lib
_lib.SO_PATH = _lib.SO_PATH.replace("liblss_b200.so", "liblss_b200_ab_prev.so")
sys.argv = ["q"] + sys.argv[1:]
runpy.run_path("scripts/bench_runplan_quick.py", run_name="__main__")
PY
}
for i in 1 2; do
echo PREV; run_prev cfg2 2000 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in $KEYS})"
echo NEW; python scripts/bench_runplan_quick.py cfg2 2000 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in $KEYS})"
done
echo PREV4; run_prev cfg4 300 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in $KEYS})"
echo NEW4; python scripts/bench_runplan_quick.py cfg4 300 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in $KEYS})"
LSS_TIMELINE=1 python scripts/bench_runplan_quick.py cfg2 300 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d.get('timeline_us'))"
python -m pytest tests/test_runplan_gpu.py -x -q 2>&1 | tail -3
