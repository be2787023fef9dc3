mkdir -p gpurun_out
timeout 120 python tools/pcie_probe.py 2>/dev/null | tail -1 | tee gpurun_out/pcie_n1.json
for np in 4 8 12 16; do
echo "PANELS=$np: $(timeout 120 env LIBXSMM_B200_EXEC_PANELS=$np python bench.py --others '' --sharded '' --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['e2e']['ms_per_step'], d['e2e']['value'])")"
done
LIBXSMM_B200_EXEC_TRACE=1 timeout 120 python bench.py --others '' --sharded '' --no-cpu --steps 3 2>&1 >/dev/null | grep "exec_host step" | tail -8
