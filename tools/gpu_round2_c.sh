mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -25 gpurun_out/pytest_gpu.log
echo "K2s c2: $(timeout 60 env LIBXSMM_B200_SPMDM_TC=0 python tools/time_spmdm.py c2 20 2>&1 | tail -2 | head -1)"
echo "K2s c1: $(timeout 60 env LIBXSMM_B200_SPMDM_TC=0 python tools/time_spmdm.py c1 20 2>&1 | tail -2 | head -1)"
timeout 600 python bench.py > gpurun_out/bench_a.json 2> gpurun_out/bench_a.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/bench_a.err
