timeout 600 python -m pytest tests/test_spmdm_tc_gpu.py tests/test_spmdm_gpu.py -m gpu -x -q -k "bf16" 2>&1 | tail -5
LIBXSMM_B200_TC16_PAIR=0 timeout 300 python -m pytest tests/test_spmdm_tc_gpu.py -m gpu -x -q -k "bf16_tensor_core_branch" 2>&1 | tail -3
export LIBXSMM_B200_SPMDM_TC=1
LIBXSMM_B200_TC16_DBG=64 timeout 120 python tools/time_spmdm.py c2 1 2>&1 | tail -9
timeout 120 python tools/time_spmdm.py c2 20 2>&1 | tail -2
LIBXSMM_B200_TC16_PAIR=0 timeout 120 python tools/time_spmdm.py c2 20 2>&1 | tail -2
