mkdir -p gpurun_out
# (a) K4s (structured-sparse tensor-core kernel) on C2
timeout -s KILL 120 python tools/time_spmdm.py c2 3 > gpurun_out/plain_c2.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:spmdm_compute_tc16s -s 2 -c 1 -f -o gpurun_out/r02_c2_tc16s python tools/time_spmdm.py c2 3 > gpurun_out/ncu_c2s.log 2>&1
# (b) K1x with the structured-sparse words on C2
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:spmdm_slice_bf16x -s 2 -c 1 -f -o gpurun_out/r02_c2_k1x_sp python tools/time_spmdm.py c2 3 > gpurun_out/ncu_k1x_sp.log 2>&1
# (c) launch list of the bench command
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --others '' --sharded '' --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c2_bench.csv python bench.py --steps 2 --warmup 3 --others '' --sharded '' --no-cpu > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/r02_c2_tc16s.ncu-rep gpurun_out/r02_c2_k1x_sp.ncu-rep gpurun_out/r02_launches_c2_bench.csv; tail -n 2 gpurun_out/ncu_c2s.log
