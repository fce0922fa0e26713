#!/bin/bash
# TEST INFRASTRUCTURE ONLY: compiles the CUDA sources with g++ against the CPU SIMT emulator
# (tests/emul/cuda_emul.h) into tests/emul/libsperr_emul.so so kernel logic can be debugged in a
# container without a GPU. Never shipped, never loaded by the product.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../sperr_b200/csrc"
CXX=/usr/bin/g++
# experiments: SPERR_EMUL_DEFS="-DSPERR_REC_PREFIX=1" SPERR_EMUL_TAG=recprefix tests/emul/build.sh
# builds tests/emul/libsperr_emul_<tag>.so from objects of its own (scripts/build_variants.sh)
TAG="${SPERR_EMUL_TAG:+_$SPERR_EMUL_TAG}"
FLAGS="-std=c++17 -O1 -g -fPIC -ffp-contract=off -DSPERR_EMUL $SPERR_EMUL_DEFS -I$HERE -I$SRC -include $HERE/cuda_emul.h -Wall -Wno-unknown-pragmas -Wno-unused-function"
OBJS=""
OBJDIR="$HERE/obj$TAG"
mkdir -p "$OBJDIR"
pids=()
for f in $SRC/*.cu $SRC/*.cpp $HERE/cuda_emul.cpp; do
  o="$OBJDIR/$(basename $f).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC $HERE $HERE/../../include -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o" 2>/dev/null | head -1)" ]; then
    ( $CXX $FLAGS -x c++ -c "$f" -o "$o.tmp" 2> "$o.log" && mv "$o.tmp" "$o" || { rm -f "$o"; grep -m8 -E "error" "$o.log" >&2; exit 1; } ) &
    pids+=($!)
  fi
  OBJS="$OBJS $o"
done
fail=0
for p in "${pids[@]}"; do wait $p || fail=1; done
[ $fail = 0 ] || { echo "emul build FAILED" >&2; exit 1; }
$CXX -shared -o "$HERE/libsperr_emul$TAG.so" $OBJS -lpthread
echo "built $HERE/libsperr_emul$TAG.so"
