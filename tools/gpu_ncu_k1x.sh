mkdir -p gpurun_out
timeout -s KILL 300 ncu --set full --clock-control none --import-source on -k regex:spmdm_slice_bf16x -s 2 -c 1 -f -o gpurun_out/r02_c2_k1x_spread python tools/time_spmdm.py c2 3 > gpurun_out/ncu_k1x_spread.log 2>&1
ls -la gpurun_out/r02_c2_k1x_spread.ncu-rep
