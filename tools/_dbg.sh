for v in 4 8 12 16; do
LIBXSMM_B200_EXEC_PANELS=$v timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu --others "" 2>/dev/null | tail -1 | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('panels=$v', round(r['value']), r['e2e']['ms_per_step'], round(r['e2e']['value']))"
done
