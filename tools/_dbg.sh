for v in 0 1; do
LIBXSMM_B200_EXEC_TWO_UP=$v timeout 300 python bench.py --steps 5 --warmup 3 2>/dev/null | tail -1 | python -c "
import json,sys; r=json.loads(sys.stdin.read()); print('two_up=$v', r['value'], r['ms_per_step'], r['e2e']['ms_per_step'], r['e2e']['value'])"
done
LIBXSMM_B200_EXEC_TWO_UP=1 LIBXSMM_B200_EXEC_TRACE=1 timeout 300 python bench.py --steps 3 --warmup 3 2>&1 >/dev/null | tail -8
