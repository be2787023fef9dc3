mkdir -p gpurun_out
for dbg in 1 2; do for cl in 1 2 4; do
echo "DEBUG=$dbg CL=$cl: $(timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=$cl LIBXSMM_B200_K2S_DEBUG=$dbg python tools/time_spmdm.py c2 20 2>&1 | tail -2 | head -1)"
done; done
echo "C4 CL=1: $(timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=1 python tools/time_spmdm.py c4-nnn 10 2>&1 | tail -2 | head -1)"
echo "C4 K2: $(timeout 60 env LIBXSMM_B200_SPMDM_TC=0 LIBXSMM_B200_K2S=0 python tools/time_spmdm.py c4-nnn 10 2>&1 | tail -2 | head -1)"
