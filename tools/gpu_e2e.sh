mkdir -p gpurun_out
for sp in 0 1; do
echo "EXEC_SPLIT=$sp: $(timeout 120 env LIBXSMM_B200_EXEC_SPLIT=$sp python bench.py --others '' --sharded '' --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['e2e']['ms_per_step'], d['e2e']['value'])")"
done
LIBXSMM_B200_EXEC_TRACE=1 timeout 120 python bench.py --others '' --sharded '' --no-cpu --steps 3 2>&1 >/dev/null | grep "exec_host step" | tail -8
timeout 300 python -m pytest tests/test_spmdm_gpu.py -m gpu -q -x -k "exec_host" 2>&1 | tail -2
timeout 300 python bench.py --workload soa --others '' --sharded '' --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print('soa', d['ms_per_step'], d['value'], d['roofline']['kernel'], d['roofline']['frac'])"
