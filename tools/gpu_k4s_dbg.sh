# K4s: which stage bounds it?  (LIBXSMM_B200_K4S_DEBUG: 1 no worker work, 2 no B loads, 4 no C stores, 8 no overflow pass, 16 no MMAs)
for d in ${K4S_FLAGS:-0 8 9 10 12 24 31}; do echo -n "debug=$d  "; LIBXSMM_B200_K4S_DEBUG=$d timeout 100 python tools/time_spmdm.py c2 20 2>/dev/null | head -1; done
