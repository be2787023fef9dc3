for sp in 0 1 0 1; do
echo "TC16_SPARSE=$sp: $(timeout -s KILL 120 env LIBXSMM_B200_TC16_SPARSE=$sp python bench.py --others '' --sharded '' --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['e2e']['ms_per_step'], d['e2e']['value'], d['ms_per_step'])")"
done
LIBXSMM_B200_EXEC_TRACE=1 timeout -s KILL 120 python bench.py --others '' --sharded '' --no-cpu --steps 3 2>&1 >/dev/null | grep "exec_host step" | tail -9
