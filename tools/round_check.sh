# what the driver runs at round end, on one GPU: the GPU tests, smoke(), both bench arms
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_r02_ref.json 2> gpurun_out/bench_r02_ref.err; cut -c1-700 gpurun_out/bench_r02_ref.json
timeout 900 python bench.py > gpurun_out/bench_r02.json 2> gpurun_out/bench_r02.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_r02.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r02.json'))
print('value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3), 'cpu', round(d['cpu_baseline']['value'],1))
for k,v in d['column_sharded'].items(): print(' sharded', k, round(v['value']), round(v['per_gpu_hbm_frac'],3))
for k,v in d['other_workloads'].items():
    if 'error' in v: print(' other', k, 'ERROR', v['error']); continue
    print(' other', k, round(v['value']), 'ms', round(v['ms_per_step'],4), v['roofline']['kernel'], v['roofline']['bound'], round(v['roofline']['frac'],3), 'e2e', round(v.get('e2e',{}).get('value',0)), 'cpu', round(v.get('cpu_baseline',{}).get('value',0),1))
PY
