set -x
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; tail -3 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_r01.json 2> gpurun_out/bench_r01.err; tail -c 600 gpurun_out/bench_r01.json
timeout 600 python bench.py --impl reference > gpurun_out/bench_r01_ref.json 2> gpurun_out/bench_r01_ref.err; cat gpurun_out/bench_r01_ref.json
