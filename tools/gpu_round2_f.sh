mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_spmdm_gpu.py tests/test_spmdm_tc_gpu.py tests/test_hardening_gpu.py tests/test_spmdm_sweep_gpu.py -m gpu -q -x > gpurun_out/pytest_k1.log 2>&1; tail -5 gpurun_out/pytest_k1.log
for sp in 0 1; do
echo "SPLIT=$sp c2: $(timeout 60 env LIBXSMM_B200_K1_SPLIT=$sp python tools/time_spmdm.py c2 20 2>&1 | tail -2 | tr '\n' ' ')"
echo "SPLIT=$sp c1: $(timeout 60 env LIBXSMM_B200_K1_SPLIT=$sp python tools/time_spmdm.py c1 20 2>&1 | tail -2 | tr '\n' ' ')"
echo "SPLIT=$sp c4: $(timeout 60 env LIBXSMM_B200_K1_SPLIT=$sp python tools/time_spmdm.py c4 20 2>&1 | tail -2 | tr '\n' ' ')"
done
