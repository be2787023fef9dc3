timeout 120 python tools/time_spmdm.py c2 30 2>&1 | tail -2
timeout 900 python -m pytest tests/test_spmdm_tc_gpu.py -m gpu -x -q 2>&1 | tail -3
