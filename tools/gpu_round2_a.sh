set -x
mkdir -p gpurun_out
timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=1 python tools/time_spmdm.py c2 20 2>&1 | tail -2
timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=2 python tools/time_spmdm.py c2 20 2>&1 | tail -2
timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=4 python tools/time_spmdm.py c2 20 2>&1 | tail -2
timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=0 python tools/time_spmdm.py c1 20 2>&1 | tail -2
timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=4 python tools/time_spmdm.py c1 20 2>&1 | tail -2
timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=1 python tools/time_spmdm.py c1 20 2>&1 | tail -2
timeout 300 python -m pytest tests/test_spmdm_gpu.py -x -q > gpurun_out/pytest_spmdm.log 2>&1; tail -15 gpurun_out/pytest_spmdm.log
