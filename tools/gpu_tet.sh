# dense-branch fp64 operators (PyFR tet family): the emitter variant chosen by timing at create (LIBXSMM_VERBOSE=1 names it) against
# the plain form (LIBXSMM_B200_FSSPMDM_TUNE=0)
for op in ${TET_OPS:-p4/tet/m6 p5/tet/m0 p4/tet/m3 p5/tet/m460 p6/tri/m132 p6/tet/m0}; do
  echo "== $op: tuned"; LIBXSMM_VERBOSE=1 timeout 300 python tools/time_fs_op.py $op 10 2>&1 | grep -E "variant|beta="
  echo "== $op: plain form"; LIBXSMM_B200_FSSPMDM_TUNE=0 timeout 300 python tools/time_fs_op.py $op 10 2>&1 | grep -E "beta="
done
