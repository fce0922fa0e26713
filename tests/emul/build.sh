#!/bin/bash
# TEST INFRASTRUCTURE ONLY: compiles the CUDA sources with g++ against the CPU SIMT emulator
# (tests/emul/cuda_emul.h) into tests/emul/libsperr_emul.so so kernel logic can be debugged in a
# container without a GPU. Never shipped, never loaded by the product.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../sperr_b200/csrc"
CXX=/usr/bin/g++
FLAGS="-std=c++17 -O1 -g -fPIC -ffp-contract=off -DSPERR_EMUL -I$HERE -I$SRC -include $HERE/cuda_emul.h -Wall -Wno-unknown-pragmas -Wno-unused-function"
OBJS=""
mkdir -p "$HERE/obj"
for f in $SRC/*.cu $SRC/*.cpp $HERE/cuda_emul.cpp; do
  o="$HERE/obj/$(basename $f).o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ -n "$(find $SRC $HERE -maxdepth 1 \( -name '*.h' -o -name '*.cuh' \) -newer "$o" 2>/dev/null | head -1)" ]; then
    $CXX $FLAGS -x c++ -c "$f" -o "$o" &
  fi
  OBJS="$OBJS $o"
done
wait
$CXX -shared -o "$HERE/libsperr_emul.so" $OBJS -lpthread
echo "built $HERE/libsperr_emul.so"
