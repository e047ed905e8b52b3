#!/usr/bin/env bash
# Build the UNMODIFIED reference yahmm (Cython) into oracle/_ref/ as a CPython extension.
# TEST INFRASTRUCTURE ONLY: the product never imports anything from oracle/.
# Source is read where it lies (/root/reference, read-only); only derived files land in oracle/_ref/
# (git-ignored, but shipped to the GPU box by gpurun so the CPU baseline can run there).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${PROTPORE_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python3}"
mkdir -p "$OUT"
EXT="$($PY -c 'import sysconfig;print(sysconfig.get_config_var("EXT_SUFFIX"))')"
SO="$OUT/yahmm$EXT"
# the segmenter of SURVEY 8f-1 (PyPore/cparsers.pyx, FastStatSplit); same recipe, its own extension
SO2="$OUT/cparsers$EXT"
if [ -f "$REF/PyPore/cparsers.pyx" ] && { [ ! -f "$SO2" ] || [ "$REF/PyPore/cparsers.pyx" -nt "$SO2" ] || [ "${1:-}" = "--force" ]; }; then
  $PY -m cython -2 "$REF/PyPore/cparsers.pyx" -o "$OUT/cparsers.c"
  gcc -O2 -fPIC -shared -w -I"$($PY -c 'import numpy;print(numpy.get_include())')" \
      -I"$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')" "$OUT/cparsers.c" -o "$SO2"
  echo "oracle/_ref: built $SO2"
fi
# the segment aligner of SURVEY 8f-4 (PyPore/calignment.pyx, cSegmentAligner)
SO3="$OUT/calignment$EXT"
if [ -f "$REF/PyPore/calignment.pyx" ] && { [ ! -f "$SO3" ] || [ "$REF/PyPore/calignment.pyx" -nt "$SO3" ] || [ "${1:-}" = "--force" ]; }; then
  $PY -m cython -2 "$REF/PyPore/calignment.pyx" -o "$OUT/calignment.c"
  gcc -O2 -fPIC -shared -w -I"$($PY -c 'import numpy;print(numpy.get_include())')" \
      -I"$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')" "$OUT/calignment.c" -o "$SO3"
  echo "oracle/_ref: built $SO3"
fi
if [ -f "$SO" ] && [ "$SO" -nt "$REF/yahmm/yahmm.pyx" ] && [ "${1:-}" != "--force" ]; then
  echo "oracle/_ref: up to date ($SO)"; exit 0
fi
if [ ! -f "$REF/yahmm/yahmm.pyx" ]; then
  echo "oracle/_ref: reference source not present at $REF (GPU box uses the prebuilt .so)"; exit 0
fi
# language_level 2: the source uses print statements (yahmm.pyx:2348, 2374, 2424, 2439, 4028)
$PY -m cython -2 "$REF/yahmm/yahmm.pyx" -o "$OUT/yahmm.c"
NPINC="$($PY -c 'import numpy;print(numpy.get_include())')"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
gcc -O2 -fPIC -shared -w -I"$NPINC" -I"$PYINC" "$OUT/yahmm.c" -o "$SO"
echo "oracle/_ref: built $SO"
