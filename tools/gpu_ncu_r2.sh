mkdir -p gpurun_out
# (a) K4q on C1 (the kernel the default dispatch runs there)
timeout 120 python tools/time_spmdm.py c1 3 > gpurun_out/plain_c1.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:spmdm_compute_tcq -s 2 -c 1 -f -o gpurun_out/r02_c1_tcq python tools/time_spmdm.py c1 3 > gpurun_out/ncu_c1.log 2>&1
# (b) the strip kernel on p4/hex/m0 (c3-hex)
timeout 120 python tools/time_fs.py c3-hex 3 > gpurun_out/plain_hex.log 2>&1 &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fs_baked -s 2 -c 1 -f -o gpurun_out/r02_c3hex_strip python tools/time_fs.py c3-hex 3 > gpurun_out/ncu_hex.log 2>&1
# (c) launch list of the bench command
timeout 300 python bench.py --steps 2 --warmup 3 --others '' --sharded '' --no-cpu > gpurun_out/plain_bench.log 2>&1 &&
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r02_launches_c2_bench.csv python bench.py --steps 2 --warmup 3 --others '' --sharded '' --no-cpu > gpurun_out/ncu_bench.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r02_launches_c2_bench.csv; tail -2 gpurun_out/ncu_c1.log gpurun_out/ncu_hex.log
