# K4s: parity tests, then the C2 step (events between the kernels / nothing between the launches)
timeout 600 python -m pytest tests/test_spmdm_tc16s_gpu.py -q -x 2>&1 | tail -2
for d in ${K4S_FLAGS:-0}; do for i in 1 2; do echo -n "debug=$d "; LIBXSMM_B200_K4S_DEBUG=$d timeout 100 python tools/time_spmdm.py c2 20 2>/dev/null | head -1; done; done
timeout 100 python tools/time_step_plain.py c2 40 2>/dev/null | tail -1
