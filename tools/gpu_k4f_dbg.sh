# K4f (sfsspmdm, dense float operator on tcgen05): parity of the tensor-core branch, then which stage bounds it.
# LIBXSMM_B200_K4F_EPI: 2 C through TMA store boxes (default), 1 per-thread stores in 8-row chunks, 0 round-1 epilogue; LIBXSMM_B200_K4F_PFMODE=0:
# L2 prefetch through the TMA unit; LIBXSMM_B200_K4F_PF: prefetch distance in tiles of the CTA; LIBXSMM_B200_K4F_DEBUG (results wrong): 1 no b_lo,
# 2 no C stores, 4 no MMAs, 8 no B loads.  FS_BETA=1: beta = 1.
timeout 300 python -m pytest tests/test_fsspmdm_gpu.py tests/test_mm_dispatch.py -q -x -k "tensor_core or branch_choice or dispatch" 2>&1 | tail -3
run() { echo -n "$* : "; env "$@" LIBXSMM_B200_FSSPMDM_TC=1 timeout 120 python tools/time_fs_dense.py 1.0 2>/dev/null | grep "TC=1"; }
run FS_BETA=0
run FS_BETA=1
run FS_BETA=0 LIBXSMM_B200_K4F_EPI=1
run FS_BETA=1 LIBXSMM_B200_K4F_EPI=1
for d in ${K4F_FLAGS:-2 13}; do run LIBXSMM_B200_K4F_DEBUG=$d; done
