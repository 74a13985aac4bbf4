p() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], 'e2e', d['e2e']['value'], d['e2e']['steps'])"; }
for r in 1 2 3; do
python bench.py --no-cpu-baseline --steps 1000 2>/dev/null | p plain
LSS_PIPE_WC=1 python bench.py --no-cpu-baseline --steps 1000 2>/dev/null | p wc
done
