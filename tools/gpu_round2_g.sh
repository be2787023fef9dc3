mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_mm_dispatch.py tests/test_csr_soa.py tests/test_hardening_gpu.py -m gpu -q -x > gpurun_out/pytest_new.log 2>&1; tail -12 gpurun_out/pytest_new.log
