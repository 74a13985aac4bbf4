p() { python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], 'e2e', d['e2e']['value'], d['e2e']['steps'])"; }
for r in 1 2 3; do
LSS_PIPE_H2D_STREAMS=1 python bench.py --no-cpu-baseline --steps 1000 2>/dev/null | p h2d1
LSS_PIPE_H2D_STREAMS=2 python bench.py --no-cpu-baseline --steps 1000 2>/dev/null | p h2d2
done
LSS_PIPE_H2D_STREAMS=4 python bench.py --no-cpu-baseline --steps 1000 --e2e-streams 8 2>/dev/null | p h2d4_depth8
