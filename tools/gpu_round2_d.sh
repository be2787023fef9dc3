mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharding_nccl_gpu.py tests/test_hardening_gpu.py -m gpu -q -x -k "nccl or two_devices" > gpurun_out/pytest_2gpu.log 2>&1; tail -15 gpurun_out/pytest_2gpu.log
timeout 600 python bench.py --gpus 2 --others '' > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "bench rc=$?"; tail -c 800 gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json | head -c 3000
