mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_pyfr_operators.py tests/test_reference_samples_gpu.py tests/test_fsspmdm_gpu.py tests/test_widening.py -m gpu -q -x > gpurun_out/pytest_fs.log 2>&1; tail -15 gpurun_out/pytest_fs.log
timeout 300 python bench.py --workload c3-hex --others c3-tet,c3-b1 --sharded '' --no-cpu > gpurun_out/bench_hex.json 2> gpurun_out/bench_hex.err; echo "rc=$?"; tail -c 400 gpurun_out/bench_hex.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_hex.json'))
print('c3-hex', d['ms_per_step'], d['roofline']['kernel'], d['roofline']['frac'])
for k,v in d.get('other_workloads',{}).items(): print(k, v.get('ms_per_step'), v.get('roofline',{}).get('kernel'), v.get('roofline',{}).get('frac'), v.get('error'))
PY
