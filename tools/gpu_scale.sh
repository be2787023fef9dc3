# usage: bash tools/gpu_scale.sh N    (run under gpurun --gpus N)
N=$1
mkdir -p gpurun_out
if [ "$N" -gt 1 ]; then
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tools/pcie_probe.py > gpurun_out/pcie_n$N.json 2> gpurun_out/pcie_n$N.err
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
else
  timeout 300 python tools/pcie_probe.py > gpurun_out/pcie_n1.json 2> gpurun_out/pcie_n1.err
  timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 --others '' --no-cpu > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
fi
echo "rc=$?"; cat gpurun_out/pcie_n$N.json; tail -c 300 gpurun_out/bench_n$N.err
python - <<PY
import json
d=json.load(open('gpurun_out/bench_n$N.json'))
print('N', d['n_gpus'], 'c2 value', round(d['value']), 'ms', round(d['ms_per_step'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['ms_per_step'],3))
for k,v in d['column_sharded'].items(): print(k, round(v['value']), round(v['ms_per_step'],4), round(v['per_gpu_hbm_frac'],3), 'e2e', round(v['e2e']['value']), round(v['e2e']['ms_per_step'],2))
PY
