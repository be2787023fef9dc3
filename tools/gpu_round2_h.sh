# end of round 2: ncu captures of the kernels changed last (K4f, K4s, the tuned fp64 baked kernel), launch list of the bench command,
# then what the driver runs (tools/round_check.sh)
mkdir -p gpurun_out
export LIBXSMM_B200_FSSPMDM_TC=1
timeout -s KILL 120 python tools/time_fs_dense.py 1.0 > gpurun_out/plain_k4f.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:fsspmdm_tc_kernel -s 2 -c 1 -f -o gpurun_out/r02_k4f python tools/time_fs_dense.py 1.0 > gpurun_out/ncu_k4f.log 2>&1
unset LIBXSMM_B200_FSSPMDM_TC
timeout -s KILL 120 python tools/time_spmdm.py c2 3 > gpurun_out/plain_c2.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:spmdm_compute_tc16s -s 2 -c 1 -f -o gpurun_out/r02_c2_tc16s_b python tools/time_spmdm.py c2 3 > gpurun_out/ncu_c2s.log 2>&1
timeout -s KILL 200 python tools/time_fs.py c3-tet 3 > gpurun_out/plain_tet.log 2>&1 &&
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:fs_baked -s 14 -c 1 -f -o gpurun_out/r02_c3tet_baked python tools/time_fs.py c3-tet 3 > gpurun_out/ncu_tet.log 2>&1
timeout -s KILL 300 python bench.py --steps 2 --warmup 3 --others '' --sharded '' --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c2_bench.csv python bench.py --steps 2 --warmup 3 --others '' --sharded '' --no-cpu > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -4
bash tools/round_check.sh
