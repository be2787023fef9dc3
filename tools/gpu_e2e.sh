# end-to-end call (libxsmm_spmdm_exec_host) on C2 against the number of column panels of its pipeline
mkdir -p gpurun_out
for np in ${PANELS:-8 16 12 6 8}; do
echo "EXEC_PANELS=$np: $(timeout 120 env LIBXSMM_B200_EXEC_PANELS=$np python bench.py --others '' --sharded '' --no-cpu 2>/dev/null | python -c "import json,sys; d=json.load(sys.stdin); print(d['e2e']['ms_per_step'], d['e2e']['value'])")"
done
