#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED reference (hanzz2007/libxsmm-1,
# LIBXSMM master-1.12-4) from its sources where they lie under $REF into oracle/_ref/.
# Nothing is copied into the repository: oracle/_ref/ is git-ignored (it still
# travels to the GPU box with gpurun).  The reference's own build system (its
# Makefile/Makefile.inc) is NOT used: every src/*.c is handed to gcc directly.
# The two configuration headers LIBXSMM expects (libxsmm.h, libxsmm_config.h) are
# plain $VAR substitutions of src/template/*.h; they are produced with the
# reference's own substitution scripts using their built-in defaults and written
# to oracle/_ref/include only.
#
#   usage: oracle/build_ref.sh [avx2|avx512]
#     avx2   (default) static target SSE4.2 -> runtime picks the AVX2 spmdm
#            instantiation (bn=48), exactly what `make` with GCC gives.
#     avx512 adds -mavx512{f,cd,dq,bw,vl} -mfma -> AVX-512 instantiation (bn=96).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
FLAVOR="${1:-avx2}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
  echo "build_ref: $REF not present (GPU box?) -- using prebuilt files in $OUT" >&2
  exit 0
fi
case "$FLAVOR" in
  avx2)   TARGET_FLAGS="-msse4.2"; SUFFIX="" ;;
  avx512) TARGET_FLAGS="-mavx512f -mavx512cd -mavx512dq -mavx512bw -mavx512vl -mfma"; SUFFIX="_avx512" ;;
  *) echo "unknown flavor $FLAVOR" >&2; exit 2 ;;
esac
OBJ="$OUT/obj$SUFFIX"
mkdir -p "$OUT/include" "$OBJ"
PY="${PYTHON:-python3}"
"$PY" "$REF/scripts/libxsmm_config.py"    "$REF/src/template/libxsmm_config.h" > "$OUT/include/libxsmm_config.h"
"$PY" "$REF/scripts/libxsmm_interface.py" "$REF/src/template/libxsmm.h"        > "$OUT/include/libxsmm.h"
: > "$OUT/include/libxsmm_dispatch.h"   # no statically generated kernels (default build has none either)
CFLAGS="-w -O2 -fPIC -fcommon -DNDEBUG -DLIBXSMM_BUILD -D__BLAS=0 $TARGET_FLAGS -I$OUT/include -I$REF/include -I$REF/src"
SRCS=$(ls "$REF"/src/*.c | grep -v -e 'libxsmm_ext' -e 'gemm_driver' -e 'libxsmm_python' -e 'libxsmm_perf')
JOBS="${JOBS:-$(nproc)}"
printf '%s\n' $SRCS | xargs -P "$JOBS" -I{} sh -c \
  'o="'"$OBJ"'/$(basename {} .c).o"; [ "$o" -nt {} ] || gcc '"$CFLAGS"' -c {} -o "$o"'
gcc -shared -fcommon -o "$OUT/libxsmm_ref$SUFFIX.so" "$OBJ"/*.o -lm -ldl -lpthread -lrt
# our own OpenMP driver (oracle/ref_driver.c) that calls the reference the way
# samples/spmdm/spmdm.c:88-111 and samples/pyfr/pyfr_driver_asp_reg.c:297-308 do.
DRV="$OUT/drv$SUFFIX"; mkdir -p "$DRV"
gcc -O2 -fPIC -fopenmp -fcommon -c -I"$HERE/../include" -o "$DRV/ref_driver.o" "$HERE/ref_driver.c"
gcc -O2 -fPIC -fopenmp -fcommon -w -c -I"$OUT/include" -I"$REF/include" -o "$DRV/ref_driver_soa.o" "$HERE/ref_driver_soa.c"
gcc -shared -fopenmp -fcommon -o "$OUT/libref_driver$SUFFIX.so" "$DRV/ref_driver.o" "$DRV/ref_driver_soa.o" \
    -L"$OUT" -l:libxsmm_ref$SUFFIX.so -Wl,-rpath,'$ORIGIN' -lm
echo "build_ref: wrote $OUT/libxsmm_ref$SUFFIX.so and $OUT/libref_driver$SUFFIX.so"
