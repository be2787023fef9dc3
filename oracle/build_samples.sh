#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- compiles the reference's two drivers of this path, UNMODIFIED and from where they lie under
# $REF, against the PRODUCT library: samples/spmdm/spmdm.c and samples/pyfr/pyfr_driver_asp_reg.c.  libxsmm_b200.so comes
# first on the link line, so the 14 hot-path entry points (libxsmm_spmdm_*, libxsmm_[sd]fsspmdm_*) resolve to the CUDA
# implementation; the compiled reference (oracle/_ref/libxsmm_ref.so) only supplies the service symbols the drivers also
# use (libxsmm_rng_*, libxsmm_timer_*, libxsmm_init ...).  Outputs go to oracle/_ref/samples/ (git-ignored, travels to the
# GPU box); tests/test_reference_samples_gpu.py runs them there and checks the errors they print.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref/samples"
PROD="$HERE/../libxsmm-1_b200/lib"
if [ ! -d "$REF/samples" ]; then
  echo "build_samples: $REF not present (GPU box?) -- using prebuilt files in $OUT" >&2
  exit 0
fi
[ -f "$HERE/_ref/libxsmm_ref.so" ] || "$HERE/build_ref.sh" avx2
mkdir -p "$OUT"
INC="-I$HERE/_ref/include -I$REF/include"
RPATH='-Wl,-rpath,$ORIGIN/..:$ORIGIN/../../../libxsmm-1_b200/lib'
gcc -O2 -fopenmp -fcommon -w $INC "$REF/samples/spmdm/spmdm.c" -o "$OUT/spmdm_b200" \
    -L"$PROD" -l:libxsmm_b200.so -L"$HERE/_ref" -l:libxsmm_ref.so $RPATH -lm
gcc -O2 -fopenmp -fcommon -w $INC "$REF/samples/pyfr/pyfr_driver_asp_reg.c" "$HERE/dgemm_shim.c" -o "$OUT/pyfr_b200" \
    -L"$PROD" -l:libxsmm_b200.so -L"$HERE/_ref" -l:libxsmm_ref.so $RPATH -lm
echo "build_samples: wrote $OUT/spmdm_b200 and $OUT/pyfr_b200"
